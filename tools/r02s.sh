#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02s_pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)|^E  +" $O/r02s_pytest_gpu.log | cut -c1-400 | head -20; tail -2 $O/r02s_pytest_gpu.log
F="--steps 1 --warmup 3 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"bn_bwd_apply|bn_bwd_reduce|bn_head|bn_apply_head" --csv --log-file $O/r02s_bn.csv python bench.py $F > /dev/null 2>&1
python - $O/r02s_bn.csv <<'PY'
import csv,sys
from collections import defaultdict
rows=list(csv.reader(open(sys.argv[1])))
i=next(k for k,r in enumerate(rows) if r and r[0]=='ID')
h=rows[i]; d=defaultdict(dict)
for r in rows[i+1:]:
    if len(r)<len(h): continue
    rec=dict(zip(h,r)); d[rec['ID']]['k']=rec['Kernel Name'][:48]; d[rec['ID']][rec['Metric Name']]=rec['Metric Value']
ids=sorted(d,key=int)
n=len(ids)//4
tot=defaultdict(float)
for k in ids[-n:]:
    r=d[k]; t=float(r['gpu__time_duration.sum'])/1e3; b=(float(r['dram__bytes_read.sum'])+float(r['dram__bytes_write.sum']))/1e9
    tot[r['k']]+=t
    print('%-50s %8.1f us %6.2f GB %6.0f GB/s issue %s inst %s'%(r['k'],t,b,b/t*1e6 if t else 0,r['smsp__issue_active.avg.pct_of_peak_sustained_active'],r['smsp__inst_executed.sum']))
print(dict(tot))
PY
echo "== bench x3 (lean)"
for i in 1 2 3; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f pairs/s %.1f clk %s'%(d['ms_per_step'], d['value'], d['clocks']['sm_mhz']))"; done
