#!/bin/bash
# ncu --set full over EVERY kernel of one training step and of one eval forward (64 pairs of 13x256x256), exported as
# raw-page CSV (gz) -- the .ncu-rep files stay on the box (too big to bring back).   usage: bash tools/prof_round2.sh [tag]
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
F="--steps 1 --warmup 3 --no-cpu-baseline --no-library --no-scene --no-infer --no-small"
for what in train infer; do
  FABRIC_B200_PROFILE_STEP=$what timeout 1200 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${TAG}_full_$what \
      python bench.py $F > $O/${TAG}_full_$what.log 2>&1
  tail -2 $O/${TAG}_full_$what.log | cut -c1-200
  ncu -i /tmp/${TAG}_full_$what.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/${TAG}_ncu_full_${what}_step.raw.csv.gz
  ls -la /tmp/${TAG}_full_$what.ncu-rep $O/${TAG}_ncu_full_${what}_step.raw.csv.gz
done
