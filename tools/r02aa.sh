#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest train + round2"
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_round2.py -m gpu -x -q > $O/r02aa_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|^E  +" $O/r02aa_pytest.log | cut -c1-300 | head; tail -2 $O/r02aa_pytest.log
bash tools/ab.sh FABRIC_B200_CONV_FIX r02aa 0 1 2
