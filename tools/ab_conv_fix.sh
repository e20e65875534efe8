#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest parity (eval path)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02z_pytest_parity.log 2>&1; grep -E "^(FAILED|ERROR)|^E  +" $O/r02z_pytest_parity.log | cut -c1-300 | head; tail -2 $O/r02z_pytest_parity.log
for v in 0 1 2 0 1 2; do
  FABRIC_B200_CONV_FIX=$v timeout 600 python bench.py --workload infer --steps 30 --warmup 5 --no-cpu-baseline --no-library --no-scene --no-small --e2e-steps 1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); L=d['layers']
        print('FIX=$v ms/step %.3f pairs/s %.0f clk %s'%(d['ms_per_step'], d['value'], d['clocks']['sm_mhz']), {k:round(v['ms'],3) for k,v in L.items() if '64@' in k or '256@' in k or '512@' in k}, {k:round(v['frac_of_burst_peak'],3) for k,v in d['roofline']['encoder_double_conv'].items()})"
done
