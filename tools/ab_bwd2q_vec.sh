#!/bin/bash
O=gpurun_out; mkdir -p $O
F="--steps 1 --warmup 3 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1"
for v in 8 4; do
  FABRIC_B200_BWD2Q_V=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:bn_bwd2q --csv --log-file $O/r02p_bwd2q_v$v.csv python bench.py $F > /dev/null 2>&1
  python - $O/r02p_bwd2q_v$v.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
i=next(k for k,r in enumerate(rows) if r and r[0]=='ID')
h=rows[i]; 
from collections import defaultdict
d=defaultdict(dict)
for r in rows[i+1:]:
    if len(r)<len(h): continue
    rec=dict(zip(h,r)); d[rec['ID']]['k']=rec['Kernel Name'][:60]; d[rec['ID']][rec['Metric Name']]=rec['Metric Value']
ids=sorted(d,key=int)[-8:]
for k in ids: print(sys.argv[1][-8:], d[k])
PY
done
