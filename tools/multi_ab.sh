#!/bin/bash
# A/B of environment knobs at N GPUs: one short bench.py run per argument ("-" = defaults).
# usage: gpurun --gpus N --timeout 900 -- 'bash tools/multi_ab.sh TAG N "ENV1" "ENV2" ...'
TAG=$1; N=$2; shift 2; OUT=gpurun_out/$TAG; mkdir -p $OUT
port=29540
for envs in "$@"; do
  [ "$envs" = "-" ] && envs=""
  name=$(echo "${envs:-default}" | tr ' =' '__')
  port=$((port+1))
  env $envs timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 40 --warmup 10 --no-cpu-baseline > $OUT/bench_${N}gpu_$name.jsonl 2> $OUT/bench_${N}gpu_$name.err
  python - "$OUT/bench_${N}gpu_$name.jsonl" "${envs:-default}" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d["stage_ms"]
    print(sys.argv[2], "| gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "Mpix/s", d["value"], "e2e", d["e2e"]["ms_per_step"],
          "stage sum", round(sum(s.values()), 3), {k: round(v, 3) for k, v in s.items() if k in
          ("per_gaussian_forward", "per_gaussian_backward", "rotation_backward", "blend_forward", "blend_backward", "depth_sort", "fills")})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
