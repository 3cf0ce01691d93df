"""Group an `ncu --page source --csv --print-source sass` export into code regions of equal execution count
(= loop nests): instructions per region, executions, share of all executed warp instructions.
usage: python tools/sass_regions.py file.csv [min_share_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
nxt = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Address"), len(rows))   # first launch only
data = [r for r in rows[hi + 1:nxt] if len(r) == len(h)]
ia = h.index("Instructions Executed")
num = lambda s: int(s) if s.isdigit() else 0
tot = sum(num(r[ia]) for r in data)
print("total warp instructions", tot, "SASS lines", len(data))
prev, start, acc, out = None, 0, 0, []
for i, r in enumerate(data):
    n = num(r[ia])
    if prev is None or abs(n - prev) > 0.03 * max(prev, 1):
        if prev is not None:
            out.append((start, i - 1, prev, acc))
        start, acc, prev = i, 0, n
    acc += n
out.append((start, len(data) - 1, prev, acc))
for s, e, n, a in out:
    if a > thr / 100 * tot:
        print(f"lines {s:4d}-{e:4d} ({e - s + 1:3d} instr)  executed {n:>10d} x  share {100 * a / tot:5.1f}%   first: {data[s][1].strip()[:70]}")
