#!/bin/bash
# Two-GPU call: NCCL tests (DataParallelStep segments, exact-global equivalence, device guard), the default bench line under
# torchrun (training step with the gradient all-reduce), and the scene workload sharded by row bands.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader; nvidia-smi topo -m | head -12
echo "== pytest (2-GPU tests)"
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -rA -k "nccl or exact_global or another_device or fused_update or trajectory" > $O/${TAG}_pytest_n2.log 2>&1
grep -E "^(exact-global|oracle|cuda  )" $O/${TAG}_pytest_n2.log | cut -c1-400; grep -E "^(FAILED|ERROR)|^E  +" $O/${TAG}_pytest_n2.log | cut -c1-500 | head -30; tail -3 $O/${TAG}_pytest_n2.log
echo "== bench N=2 (train)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
tail -c 400 $O/${TAG}_bench_n2.json; tail -5 $O/${TAG}_bench_n2.err
echo "== bench N=2 (scene)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload scene --scene 10000 > $O/${TAG}_bench_scene_n2.json 2> $O/${TAG}_bench_scene_n2.err
tail -c 700 $O/${TAG}_bench_scene_n2.json; tail -3 $O/${TAG}_bench_scene_n2.err
echo "== reference arm under torchrun"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -c 300
