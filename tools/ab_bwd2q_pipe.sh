#!/bin/bash
O=gpurun_out; mkdir -p $O
F="--steps 1 --warmup 3 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1"
for v in 4 7; do
  FABRIC_B200_BWD2Q_V=$v timeout 600 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:bn_bwd2q --csv --log-file $O/r02v_bwd2q_v$v.csv python bench.py $F > /dev/null 2>&1
  python - $O/r02v_bwd2q_v$v.csv <<'PY'
import csv,sys
from collections import defaultdict
rows=list(csv.reader(open(sys.argv[1])))
i=next(k for k,r in enumerate(rows) if r and r[0]=='ID')
h=rows[i]; d=defaultdict(dict)
for r in rows[i+1:]:
    if len(r)<len(h): continue
    rec=dict(zip(h,r)); d[rec['ID']]['k']=rec['Kernel Name'][22:52]; d[rec['ID']][rec['Metric Name']]=rec['Metric Value']
ids=sorted(d,key=int)[-8:]
print(sys.argv[1][-8:], [(d[k]['k'][16:30], round(float(d[k]['gpu__time_duration.sum'])/1e3,1), d[k]['smsp__issue_active.avg.pct_of_peak_sustained_active']) for k in ids], 'total', round(sum(float(d[k]['gpu__time_duration.sum']) for k in ids)/1e3,1))
PY
done
