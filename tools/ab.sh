#!/bin/bash
# A/B on one box: the training bench (headline only) with an environment switch off / on, twice each, interleaved.
# usage: bash tools/ab.sh ENV_NAME [tag] [value_a value_b] [reps]
V=$1; TAG=${2:-ab}; VA=${3:-0}; VB=${4:-1}; REPS=${5:-3}
O=gpurun_out; mkdir -p $O
F="--steps 20 --warmup 5 --no-cpu-baseline --no-library --no-scene --no-infer --no-small --e2e-steps 1"
for rep in $(seq 1 $REPS); do for val in $VA $VB; do
  env $V=$val timeout 600 python bench.py $F > $O/${TAG}_${V}_${val}_$rep.json 2>/dev/null
  python - $O/${TAG}_${V}_${val}_$rep.json $V $val <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], '=', sys.argv[3], 'ms/step %.3f'%d['ms_per_step'], 'pairs/s %.1f'%d['value'], 'clk', d['clocks']['sm_mhz'])
PY
done; done
