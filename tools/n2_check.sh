timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r01g_bench_infer_n2.json 2> gpurun_out/r01g_n2.err; tail -c 600 gpurun_out/r01g_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 5 --workload train > gpurun_out/r01g_bench_train_n2.json 2>> gpurun_out/r01g_n2.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r01g_bench_infer.json
timeout 300 python bench.py --steps 15 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/r01g_bench_train.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 1 --workload scene --scene 6000 > gpurun_out/r01g_bench_scene_n2.json 2>> gpurun_out/r01g_n2.err
python - <<EOP
import json
for f in ["r01g_bench_infer.json","r01g_bench_infer_n2.json","r01g_bench_train.json","r01g_bench_train_n2.json","r01g_bench_scene_n2.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), d.get("roofline") and round(d["roofline"]["achieved"]), d.get("e2e") and round(d["e2e"]["value"]))
    except Exception as e: print(f, "ERR", e)
EOP
