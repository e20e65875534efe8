"""Host-side cost of one eager training step at the reference's default geometry (32 pairs of 13x90x90; development tool):
cProfile over steady-state steps, top functions by own time.   python tools/host_overhead.py [steps]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fabric_b200 import BiDateNet  # noqa: E402
from fabric_b200.distributed import DataParallelStep  # noqa: E402
from fabric_b200.metrics import TverskyLoss  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = BiDateNet(13, 2).to(dev).train()
x1, x2 = torch.randn(32, 13, 90, 90, device=dev), torch.randn(32, 13, 90, 90, device=dev)
labels = (torch.rand(32, 90, 90, device=dev) < 0.1).long()
crit = TverskyLoss(alpha=0.1, beta=0.9)
dp = DataParallelStep(model)


def step():
    dp.zero_grad()
    loss = crit(model(x1, x2), labels)
    loss.backward()
    dp.sync_and_step(0.01)


for _ in range(10):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
t_host = time.perf_counter() - t0          # host time to ENQUEUE the steps
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"{steps} steps: host enqueue {1e3 * t_host / steps:.3f} ms/step, wall {1e3 * t_all / steps:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
