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
for _ in range(iters):
    ops.conv3x3(x5, wp, cout, scale, shift, relu=True, tune=tune, out=out)
torch.cuda.synchronize()
print("done", name, tune)
