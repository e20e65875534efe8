# epilogue-mode timings of single layers (development tool): plain / pool / prod / lean / stats x epilogue warps
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for L in ${LAYERS_M-inc.c2 down1.c2}; do for M in plain lean stats; do for E in 4 8; do
  FB_TUNE="dict(epi_warps=$E)" FB_MODE=$M timeout 120 python tools/prof_conv.py $L 10 2>&1 | tail -1; done; done; done
for L in inc.c1 up4.c1 up3.c1 up2.c1 down1.c1; do for E in 4 8; do FB_TUNE="dict(epi_warps=$E)" timeout 120 python tools/prof_conv.py $L 10 2>&1 | tail -1; done; done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r01f_bench_infer.json; python - <<EOP
import json
d=json.loads(open("gpurun_out/r01f_bench_infer.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["conv_share_of_step"], d["e2e"]["value"])
for k,v in d["layers"].items(): print(k, round(v["ms"],3), round(v["tflops"]))
EOP
