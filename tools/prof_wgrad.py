"""Weight-gradient kernel timings on the real layer shapes (development tool): first form (wide = 1) vs second form (wide = 2).
    python tools/prof_wgrad.py [layer ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fabric_b200 import ops  # noqa: E402
from tools.gpu_check import LAYERS  # noqa: E402

B = int(os.environ.get("FB_BENCH_B", "64"))
want = sys.argv[1:]
for name, G, H, cin, cout in LAYERS + [("up2.c2", 1, 64, 128, 128), ("up3.c2", 1, 128, 64, 64), ("up4.c2", 1, 256, 64, 64)]:
    if want and name not in want:
        continue
    cp = ops.cpad(cin)
    x5 = torch.randn(G, B, H, H, cp, device="cuda").bfloat16()
    dz = torch.randn(G, B, H, H, cout, device="cuda").bfloat16()
    res = {}
    swappable = cout == 64 and cp % 128 == 0 and cp == cin
    for wide in (1, 2) + ((11, 12) if swappable else ()):       # 11 / 12: the same two forms with the operands swapped
        kw = dict(wide=wide % 10, swap=wide > 10)
        for _ in range(2):
            dw = ops.conv3x3_wgrad(dz, x5, cin, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dw = ops.conv3x3_wgrad(dz, x5, cin, **kw)
        e1.record()
        torch.cuda.synchronize()
        res[wide] = (e0.elapsed_time(e1) / 5, dw)
    fl = 2.0 * G * B * H * H * 9 * cin * cout
    d = ((res[1][1] - res[2][1]).norm() / res[1][1].norm()).item()
    print(f"{name:9s} {cin:4d}->{cout:4d} @{H:3d}  wide1 {res[1][0]:.3f} ms {fl / res[1][0] / 1e9:6.0f} TF   wide2 {res[2][0]:.3f} ms "
          f"{fl / res[2][0] / 1e9:6.0f} TF   rel diff {d:.2e}" +
          (f"   swapped: wide1 {res[11][0]:.3f} ms {fl / res[11][0] / 1e9:6.0f} TF  wide2 {res[12][0]:.3f} ms {fl / res[12][0] / 1e9:6.0f} TF"
           f"  rel diff {((res[11][1] - res[1][1]).norm() / res[1][1].norm()).item():.2e}" if swappable else ""), flush=True)
