#!/bin/bash
# usage: tools/sweep_multi.sh NGPUS OUTFILE VAR=v ... ; torchrun bench per assignment, appends stage times
n=$1; out=$2; shift; shift
port=29600
for kv in "$@"; do
  port=$((port+1))
  env $kv python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms']
print('$kv', 'gpus', d['n_gpus'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k: round(x,3) for k,x in s.items() if k in ('per_gaussian_forward','per_gaussian_backward','rotation_backward','blend_backward')})" >> $out
done
