#!/bin/bash
# One GPU-box job: GPU parity tests, smoke, both bench arms, ncu launch list and one full capture.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG [skip-tests|tests] [no-full]'
TAG=${1:-rX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
fi
timeout 600 python bench.py > $OUT/bench_1gpu.jsonl 2> $OUT/bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.jsonl 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --no-cpu-baseline --no-other-workloads --steps 2 --warmup 3 > $OUT/launches_bench.log 2>&1; echo "ncu launches rc=$?"
[ "$3" = "no-full" ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_|fused_|rotation_backward' \
  --launch-skip 20 --launch-count 5 -o $OUT/full python bench.py --no-cpu-baseline --no-other-workloads --steps 2 --warmup 3 > $OUT/full_bench.log 2>&1; echo "ncu full rc=$?"
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench_1gpu.jsonl | cut -c1-400
