#!/bin/bash
# compute-sanitizer over a 10k-Gaussian 64x48 forward+backward of the strict, fused and exchange paths.
# usage (GPU box): bash tools/sanitize.sh TAG   -> gpurun_out/TAG/{memcheck,racecheck,synccheck}_*.log
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  for which in strict fused exchange; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tests/sanitize_scene.py $which > $OUT/${tool}_$which.log 2>&1
    echo "$tool $which rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}_$which.log | tail -1)"
  done
done
