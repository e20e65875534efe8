#!/bin/bash
# One GPU-box call of round 2: parity tests, the default bench line (train + infer + library bar + cpu baseline), the
# reference arm, sanitizer runs, ncu launch list of the training step.   usage: bash tools/gpu_round2.sh [tag] [what...]
TAG=${1:-r02}
shift
WHAT=${@:-tests bench ref sanitize launches}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
for w in $WHAT; do case $w in
tests)
  echo "== pytest -m gpu"
  timeout 1500 python -m pytest tests -m gpu -x -q -rA > $O/${TAG}_pytest_gpu.log 2>&1
  grep -E "^(default-geometry|double_conv|down|up|inconv|oracle|cuda  ) " $O/${TAG}_pytest_gpu.log | cut -c1-1500; tail -12 $O/${TAG}_pytest_gpu.log | cut -c1-400 ;;
tests_all)
  echo "== pytest -m gpu (no -x)"
  timeout 1800 python -m pytest tests -m gpu -q -rA > $O/${TAG}_pytest_gpu.log 2>&1
  grep -E "^(default-geometry|double_conv|down|up|inconv|oracle|cuda  ) " $O/${TAG}_pytest_gpu.log | cut -c1-1500
  grep -E "^(FAILED|ERROR)|^E  +(Assert|assert)" $O/${TAG}_pytest_gpu.log | cut -c1-600 | head -40; tail -3 $O/${TAG}_pytest_gpu.log ;;
poison)
  echo "== pytest -m gpu with NaN-poisoned workspaces / outputs (FABRIC_B200_POISON=1)"
  FABRIC_B200_POISON=1 timeout 1500 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest_gpu_poison.log 2>&1
  grep -E "^(FAILED|ERROR)|^E  +(Assert|assert)" $O/${TAG}_pytest_gpu_poison.log | cut -c1-400 | head; tail -2 $O/${TAG}_pytest_gpu_poison.log ;;
bench)
  echo "== bench (default: train)"
  timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err ;;
ref)
  echo "== reference arm"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>/dev/null; tail -c 400 $O/${TAG}_bench_reference.json ;;
sanitize)
  bash tools/sanitize.sh $TAG ;;
launches)
  echo "== launch list (train step)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_train.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1 > $O/${TAG}_launches_train.log 2>&1
  python tools/summarize_launches.py $O/${TAG}_launches_bench_train.csv > $O/${TAG}_launches_train_summary.md 2>&1; head -40 $O/${TAG}_launches_train_summary.md ;;
prof)
  bash tools/prof_round2.sh $TAG ;;
smoke)
  echo "== smoke"
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 ;;
esac; done
du -sh $O
