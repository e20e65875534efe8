"""Single-layer conv launches for ncu captures (development tool).
    python tools/prof_conv.py <layer> [iters]      layer in tools/gpu_check.LAYERS, e.g. down1.c2
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fabric_b200 import ops  # noqa: E402
from tools.gpu_check import LAYERS  # noqa: E402

name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = int(os.environ.get("FB_BENCH_B", "64"))
tune = eval(os.environ.get("FB_TUNE", "{}"))
(_, G, H, cin, cout), = [l for l in LAYERS if l[0] == name]
cp = ops.cpad(cin)
x5 = torch.randn(G, B, H, H, cp, device="cuda").bfloat16()
w = torch.randn(cout, cin, 3, 3, device="cuda") / (3.0 * cin ** 0.5)
wp = ops.pack_conv_weight(w, 0)
scale = torch.ones(cout, device="cuda")
shift = torch.zeros(cout, device="cuda")
out = torch.empty(G, B, H, H, cout, device="cuda", dtype=torch.bfloat16)
mode = os.environ.get("FB_MODE", "plain")   # plain | pool | prod | lean (= pool + prod, no main output) | stats
kw = {}
if mode in ("pool", "lean"):
    kw["pool"] = True
if mode in ("prod", "lean"):
    kw["prod_out"] = torch.empty(1, B, H, H, 2 * cout, device="cuda", dtype=torch.bfloat16)
if mode == "lean":
    kw["store_main"] = False
    out = None
if mode == "stats":
    kw["stats"] = True


if os.environ.get("FB_FOLD", "0") == "1":      # eval form: scale folded into the weights, shift preloaded into TMEM
    wp = ops.pack_conv_weight(w, 0, scale=scale)
    scale = None
    kw["shift_in_acc"] = True


def run(n):
    for _ in range(n):
        ops.conv3x3(x5, wp, cout, scale, shift, relu=True, tune=tune, out=out, **kw)


run(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(iters)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("done", name, mode, tune, "ms=%.4f tflops=%.1f" % (ms, 2.0 * G * B * H * H * 9 * cin * cout / ms / 1e9))
