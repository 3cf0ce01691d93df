"""Summarise an `ncu --page source --csv --print-source sass` export: hot SASS lines by executed count.

usage: python tools/sass_hot.py file.csv [min_pct]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
nxt = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Address"), len(rows))   # first launch only
data = [r for r in rows[hi + 1:nxt] if len(r) == len(h)]
ia, ism, ith = h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
num = lambda s: int(s) if s.isdigit() else 0
tot = sum(num(r[ia]) for r in data)
tots = sum(num(r[ism]) for r in data)
print("total warp inst", tot, "samples", tots, "lines", len(data))
for i, r in enumerate(data):
    n, s = num(r[ia]), num(r[ism])
    if n > tot * thr / 100 or s > tots * thr / 100:
        print(f"{i:4d} {100*n/tot:5.2f}% s={100*s/max(tots,1):5.2f}% thr={r[ith]:>5} {r[1].strip()[:100]}")
