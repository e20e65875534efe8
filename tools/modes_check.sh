# development checks: parity tests + training bench
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --workload train > gpurun_out/r01r_bench_train.json
python - <<EOP
import json
for f in ["r01r_bench_train.json"]:
    d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3), round(d["roofline"]["conv_share_of_step"],3), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], round(d["cpu_baseline"]["value"],2))
    print(sum(v["ms"]*v["launches_per_step"] for k,v in d["layers"].items() if k.startswith("wgrad")))
EOP
