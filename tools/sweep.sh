#!/bin/bash
# usage: tools/sweep.sh OUTFILE VAR=v ... ; runs bench.py once per assignment and appends stage times
out=$1; shift
for kv in "$@"; do
  env $kv python bench.py --no-cpu-baseline --no-other-workloads --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms']
print('$kv', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k: round(x,3) for k,x in s.items()})" >> $out
done
