#!/bin/bash
# A/B of tuning knobs on one box: each line = one bench.py run (100 steps) with the given environment.
# usage: gpurun --timeout 600 -- 'bash tools/ab_round.sh TAG "ENV1" "ENV2" ...'   ("-" = no extra environment)
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for envs in "$@"; do
  [ "$envs" = "-" ] && envs=""
  name=$(echo "${envs:-default}" | tr ' =' '__')
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-other-workloads --steps 100 --warmup 10 > $OUT/bench_$name.jsonl 2> $OUT/bench_$name.err
  python - "$OUT/bench_$name.jsonl" "${envs:-default}" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], {k: v for k, v in d["stage_ms"].items() if k in ("blend_backward", "blend_forward", "fills", "per_gaussian_backward")}, "train", d.get("train_iteration", {}).get("ms_per_iteration"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  i=$((i+1))
done
