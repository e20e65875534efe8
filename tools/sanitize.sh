#!/bin/bash
# compute-sanitizer over the loss kernels (all three tools) and one small training step of every loss + an eval forward
# (memcheck + initcheck + racecheck).  Logs: gpurun_out/<tag>_sanitizer_*.log; summary lines are echoed.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck initcheck racecheck; do
  echo "== $tool: loss tests"
  timeout 300 $CS --tool $tool --print-limit 30 --error-exitcode 0 \
      python -m pytest tests/test_gpu_train.py -x -q -k "losses" > $O/${TAG}_sanitizer_${tool}_loss.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/${TAG}_sanitizer_${tool}_loss.log | tail -3
  echo "== $tool: small training step"
  timeout 300 $CS --tool $tool --print-limit 30 --error-exitcode 0 \
      python tools/train_step_small.py > $O/${TAG}_sanitizer_${tool}_step.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^done|^tversky|^focal" $O/${TAG}_sanitizer_${tool}_step.log | tail -5
done
