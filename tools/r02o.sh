#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02o_pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)|^E  +" $O/r02o_pytest_gpu.log | cut -c1-400 | head -20; tail -2 $O/r02o_pytest_gpu.log
echo "== prof_wgrad"
timeout 300 python tools/prof_wgrad.py up3.c1 up4.c1 2>&1 | tail -4
echo "== A/B quad BN backward: 8 vs 4 channels per thread"
bash tools/ab.sh FABRIC_B200_BWD2Q_V r02o 8 4 2
echo "== A/B wgrad operand swap"
bash tools/ab.sh FABRIC_B200_WGRAD_SWAP r02o 0 1 2
