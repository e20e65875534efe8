#!/bin/bash
# Same 2-GPU box: N=1, N=2 (segmented async all-reduce), N=2 with one all-reduce after backward; then the 2-GPU tests.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
F="--steps 20 --warmup 5 --no-cpu-baseline --no-library --no-scene --no-infer --no-small"
run() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1].split('/')[-1], 'N=%d'%d['n_gpus'], 'pairs/s %.1f'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.1f'%d['e2e']['value'], 'raw %.1f'%d['e2e_raw_uint16']['value'])
PY
}
timeout 600 python bench.py --gpus 1 $F > $O/${TAG}_scal_n1.json 2>/dev/null; run $O/${TAG}_scal_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 $F > $O/${TAG}_scal_n2.json 2>/dev/null; run $O/${TAG}_scal_n2.json
FABRIC_B200_NO_OVERLAP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 $F > $O/${TAG}_scal_n2_nooverlap.json 2>/dev/null; run $O/${TAG}_scal_n2_nooverlap.json
echo "== 2-GPU tests"
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -rA -k "nccl or exact_global or another_device" > $O/${TAG}_pytest_n2.log 2>&1
grep -E "^(exact-global)" $O/${TAG}_pytest_n2.log | cut -c1-400; grep -E "^(FAILED|ERROR)|^E  +" $O/${TAG}_pytest_n2.log | cut -c1-400 | head; tail -2 $O/${TAG}_pytest_n2.log
