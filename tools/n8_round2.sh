#!/bin/bash
# Eight-GPU call (charged 8x): lean N=1 training line, the driver's exact N=8 command, then the multi-GPU tests.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo; nproc
run() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        print(sys.argv[1].split('/')[-1], 'N=%d'%d['n_gpus'], 'pairs/s %.1f'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.1f'%d['e2e']['value'],
              'raw %.1f'%d['e2e_raw_uint16']['value'], 'numa', d.get('numa_node'), 'clk', (d.get('clocks') or {}).get('sm_mhz'))
        if d.get('infer'):
            i=d['infer']; print('   infer %.1f'%i['value'], 'e2e %.1f'%i['e2e']['value'], 'raw-mask %.1f'%i['e2e_raw_uint16_mask']['value'])
        if d.get('scene'):
            print('   scene', {k: d['scene'].get(k) for k in ('value','ms_per_scene','tiles_this_rank','sharding','unavailable')})
PY
}
F="--steps 20 --warmup 5 --no-cpu-baseline --no-library --no-scene --no-small"
echo "== N=1 (lean)"
timeout 300 python bench.py --gpus 1 $F > $O/${TAG}_n8box_n1.json 2> $O/${TAG}_n8box_n1.err; run $O/${TAG}_n8box_n1.json
echo "== N=8 (the driver's command)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err
run $O/${TAG}_bench_n8.json; grep -v "^W\|^\s*$\|\*\*\*" $O/${TAG}_bench_n8.err | tail -5
echo "== multi-GPU tests"
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -rA -k "nccl or exact_global or another_device" > $O/${TAG}_pytest_n8box.log 2>&1
grep -E "^(exact-global)" $O/${TAG}_pytest_n8box.log | cut -c1-400; grep -E "^(FAILED|ERROR)|^E  +" $O/${TAG}_pytest_n8box.log | cut -c1-400 | head; tail -2 $O/${TAG}_pytest_n8box.log
