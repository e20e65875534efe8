"""In-situ timeline of the training step (development tool): torch.profiler (CUPTI) over a few steady-state steps --
per-kernel durations as they run INSIDE the step (power-capped clocks, warm L2, side stream), idle gaps on the main stream.
    python tools/timeline.py [steps]"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from fabric_b200 import BiDateNet  # noqa: E402
from fabric_b200.distributed import DataParallelStep  # noqa: E402
from fabric_b200.metrics import TverskyLoss  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = BiDateNet(13, 2).to(dev).train()
g = torch.Generator(device=dev).manual_seed(1)
x1 = torch.randn(64, 13, 256, 256, device=dev, generator=g)
x2 = torch.randn(64, 13, 256, 256, device=dev, generator=g)
labels = (torch.rand(64, 256, 256, device=dev, generator=g) < 0.1).long()
crit = TverskyLoss(alpha=0.1, beta=0.9)
dp = DataParallelStep(model)


def step():
    dp.zero_grad()
    loss = crit(model(x1, x2), labels)
    loss.backward()
    dp.sync_and_step(0.01)


for _ in range(8):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time > 0]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
span = (t1 - t0) / steps
agg = defaultdict(lambda: [0.0, 0])
for e in ev:
    k = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0][:60]
    agg[k][0] += e.device_time
    agg[k][1] += 1
busy = sum(v[0] for v in agg.values()) / steps
# union of busy intervals (any stream)
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
union, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e in iv[1:]:
    if s > cur_e:
        union += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
union += cur_e - cur_s
print(f"steps {steps}: span {span / 1e3:.3f} ms/step, sum of kernel time {busy / 1e3:.3f} ms/step, GPU busy (union) "
      f"{union / steps / 1e3:.3f} ms/step, idle {(span - union / steps) / 1e3:.3f} ms/step, launches {len(ev) // steps}")
print("| kernel | launches/step | ms/step | share of span |\n|---|---|---|---|")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if t / steps / span > 0.002:
        print(f"| `{k}` | {n / steps:.0f} | {t / steps / 1e3:.3f} | {100 * t / steps / span:.1f} % |")
