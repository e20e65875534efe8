# development checks: parity tests, then single-layer timings (FB_MODE = plain | pool | prod | lean | stats, FB_TUNE = tuning dict)
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r01m_bench_infer.json; python - <<EOP
import json
d=json.loads(open("gpurun_out/r01m_bench_infer.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["conv_share_of_step"], "e2e", d["e2e"]["value"], "raw", d["e2e_raw_uint16"]["value"])
for k,v in d["layers"].items(): print(k, round(v["ms"],3), round(v["tflops"]))
EOP
