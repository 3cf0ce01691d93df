#!/bin/bash
# usage: tools/variants.sh ENVVAR v0 v1 ...   -> stage times of bench.py per value of ENVVAR
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --no-cpu-baseline --no-other-workloads --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('$var=$v', 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k: round(x,3) for k,x in s.items()})"
done
