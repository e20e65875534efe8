#!/bin/bash
# One GPU-box call: parity tests, both bench workloads, ncu launch lists, ncu --set full captures of the headline kernels.
# Everything lands in gpurun_out/ (merged back; the .ncu-rep files are exported to CSV and removed to stay under the
# 64 MiB return limit).   usage: bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/${TAG}_pytest_gpu.log
grep -q " passed" $O/${TAG}_pytest_gpu.log && ! grep -q "failed" $O/${TAG}_pytest_gpu.log || { echo "tests failed: stopping"; exit 1; }
echo "== bench infer"
timeout 600 python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench_infer.json 2> $O/${TAG}_bench_infer.err; tail -c 300 $O/${TAG}_bench_infer.json
echo "== bench train"
timeout 600 python bench.py --workload train --steps 20 --warmup 5 > $O/${TAG}_bench_train.json 2> $O/${TAG}_bench_train.err; tail -c 300 $O/${TAG}_bench_train.json
echo "== launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_infer.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/${TAG}_launches_infer.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_bench_train.csv \
    python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/${TAG}_launches_train.log 2>&1
echo "== ncu --set full (headline layers)"
for L in ${FULL_LAYERS-down1.c2 down1.c1 inc.c2 inc.c1 down2.c2 down3.c2 up4.c1}; do
    timeout 300 ncu --set full --clock-control none -k regex:conv3x3_umma -s 3 -c 1 -f -o $O/${TAG}_full_${L} \
        python tools/prof_conv.py $L 5 > $O/${TAG}_full_${L}.log 2>&1
    tail -1 $O/${TAG}_full_${L}.log
    ncu -i $O/${TAG}_full_${L}.ncu-rep --page raw --csv > $O/${TAG}_full_${L}.raw.csv 2>/dev/null && rm -f $O/${TAG}_full_${L}.ncu-rep
done
du -sh $O
