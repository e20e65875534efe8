#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest wgrad + round2"
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02t_pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)|^E  +" $O/r02t_pytest_gpu.log | cut -c1-300 | head -20; tail -2 $O/r02t_pytest_gpu.log
for hp in 0 1; do echo "== prof_wgrad HP=$hp"; FABRIC_B200_WGRAD_HP=$hp timeout 300 python tools/prof_wgrad.py inc.c1 inc.c2 up3.c2 up4.c2 2>&1 | grep -v Warn | cut -c1-110; done
echo "== bench x2"
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f pairs/s %.1f clk %s'%(d['ms_per_step'], d['value'], d['clocks']['sm_mhz']), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['by_kind'].items()})"; done
