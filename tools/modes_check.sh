# development checks: parity tests + both bench workloads
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r01q_bench_infer.json
timeout 300 python bench.py --steps 20 --warmup 5 --workload train > gpurun_out/r01q_bench_train.json
python - <<EOP
import json
for f in ["r01q_bench_infer.json","r01q_bench_train.json"]:
    d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), round(d["roofline"]["conv_share_of_step"],3), "e2e", round(d["e2e"]["value"]), d.get("e2e_raw_uint16") and round(d["e2e_raw_uint16"]["value"]), d["clocks"]["sm_mhz"], round(d["cpu_baseline"]["value"],2))
EOP
