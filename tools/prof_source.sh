#!/bin/bash
# Source-level ncu captures (--set full --import-source on) of two kernels, exported as source-page + raw-page CSV (gz):
#   the eval-mode 64 -> 64 @ 256^2 encoder conv (inc.c2: pool + date product epilogue, BN folded) and the quad BN-backward reduce.
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
FB_MODE=lean FB_FOLD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 2 -c 1 -f -o /tmp/${TAG}_inc_c2 \
    python tools/prof_conv.py inc.c2 4 > $O/${TAG}_src_inc_c2.log 2>&1
tail -2 $O/${TAG}_src_inc_c2.log | cut -c1-200
ncu -i /tmp/${TAG}_inc_c2.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_ncu_source_inc_c2_eval.csv.gz
ncu -i /tmp/${TAG}_inc_c2.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/${TAG}_ncu_raw_inc_c2_eval.csv.gz
F="--steps 1 --warmup 3 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd2q -s 30 -c 2 -f -o /tmp/${TAG}_bwd2q python bench.py $F > $O/${TAG}_src_bwd2q.log 2>&1
tail -2 $O/${TAG}_src_bwd2q.log | cut -c1-200
ncu -i /tmp/${TAG}_bwd2q.ncu-rep --page source --csv 2>/dev/null | gzip > $O/${TAG}_ncu_source_bn_bwd2q.csv.gz
ncu -i /tmp/${TAG}_bwd2q.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/${TAG}_ncu_raw_bn_bwd2q.csv.gz
ls -la $O/${TAG}_ncu_*.gz
