# development checks: parity tests, then single-layer timings (FB_MODE = plain | pool | prod | lean | stats, FB_TUNE = tuning dict)
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for M in lean; do FB_MODE=$M timeout 120 python tools/prof_conv.py down1.c2 10 2>&1 | tail -1; done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r01l_bench_infer.json; python - <<EOP
import json
d=json.loads(open("gpurun_out/r01l_bench_infer.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["conv_share_of_step"], "e2e", d["e2e"]["value"], "raw", d["e2e_raw_uint16"])
EOP
timeout 300 python bench.py --steps 3 --warmup 1 --workload scene --scene 10000 > gpurun_out/r01l_bench_scene.json; tail -c 900 gpurun_out/r01l_bench_scene.json
