#!/bin/bash
# One multi-GPU box job: N>1 parity of the splat exchange, bench in peer and nccl modes, reference arm at the same N.
# usage: gpurun --gpus N --timeout 900 -- 'bash tools/multi_round.sh TAG N'
TAG=$1; N=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 400 python -m pytest tests/test_exchange_multi_gpu.py -m gpu -q -s > $OUT/pytest_exchange_${N}gpu.log 2>&1; tail -4 $OUT/pytest_exchange_${N}gpu.log
port=29520
for mode in peer nccl; do
  port=$((port+1))
  ADGS_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline > $OUT/bench_${N}gpu_$mode.jsonl 2> $OUT/bench_${N}gpu_$mode.err
  python - "$OUT/bench_${N}gpu_$mode.jsonl" $mode <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "Mpix/s", d["value"], "e2e", d["e2e"]["ms_per_step"],
          {k: round(v, 3) for k, v in d["stage_ms"].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
  bench.py --impl reference --gpus $N --steps 5 --warmup 3 > $OUT/bench_${N}gpu_reference.jsonl 2> $OUT/bench_${N}gpu_reference.err
cut -c1-220 $OUT/bench_${N}gpu_reference.jsonl
