#!/bin/bash
# ncu --set full capture of one elementwise training kernel inside the train bench: tools/prof_train_kernel.sh <regex> <skip>
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-40} -c 2 -f -o gpurun_out/prof_$1 \
    python bench.py --workload train --steps 1 --warmup 2 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_$1.log 2>&1
tail -2 gpurun_out/prof_$1.log
