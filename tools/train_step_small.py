"""One small training step + eval forward + every loss (config 1 shapes), for compute-sanitizer runs."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fabric_b200 import BiDateNet, metrics
from fabric_b200.distributed import DataParallelStep
from oracle import bidatenet_oracle as O

dev = torch.device("cuda:0")
model = BiDateNet(13, 2)
model.load_state_dict(O.make_state_dict(seed=0))
model = model.to(dev).train()
x1, x2, labels = O.make_inputs(2, 32, seed=1)
x1, x2, labels = x1.to(dev), x2.to(dev), labels.to(dev)
dp = DataParallelStep(model)
for name, crit in (("tversky", metrics.TverskyLoss(0.1, 0.9)), ("dice", metrics.dice_loss), ("jaccard", metrics.jaccard_loss),
                   ("focal", metrics.FocalLoss(2.0)), ("ce", metrics.cross_entropy_loss)):
    dp.zero_grad()
    loss = crit(model(x1, x2), labels)
    loss.backward()
    dp.sync_and_step(1e-3)
    print(name, float(loss))
model.eval()
with torch.no_grad():
    print("eval", float(model(x1, x2).abs().sum()))
# kernels the planner only picks at large sizes, forced here: halo-P weight gradient (stem and 64 -> 64), operand swap,
# quad BatchNorm backward with the software pipeline (even map, several quads per thread), plain-case main loops
from fabric_b200 import ops
torch.manual_seed(0)
for cin, cout, wide, swap in ((13, 64, 4, False), (64, 64, 4, False), (128, 64, 3, True)):
    x5 = torch.randn(2, 2, 32, 24, ops.cpad(cin), device=dev).bfloat16()
    dz = torch.randn(2, 2, 32, 24, cout, device=dev).bfloat16()
    print("wgrad", cin, cout, float(ops.conv3x3_wgrad(dz, x5, cin, wide=wide, swap=swap).abs().sum()))
G, B, H, W, C = 2, 2, 96, 96, 64
z = torch.randn(G, B, H, W, C, device=dev).bfloat16()
scale, shift = torch.rand(G, C, device=dev) + 0.5, torch.randn(G, C, device=dev) * 0.3
mean, invstd, gam = torch.zeros(G, C, device=dev), torch.ones(G, C, device=dev), torch.ones(C, device=dev)
gcat = torch.randn(1, B, H, W, 2 * C, device=dev).bfloat16()
gp = torch.randn(G, B, H // 2, W // 2, C, device=dev).bfloat16()
print("bn_bwd quad", float(ops.bn_relu_bwd(z, None, gcat, True, gp, scale, shift, mean, invstd, gam)[0].float().abs().sum()))
ga = torch.randn(G, B, H, W, C, device=dev).bfloat16()
print("bn_bwd plain", float(ops.bn_relu_bwd(z, None, ga, False, None, scale, shift, mean, invstd, gam)[0].float().abs().sum()))
torch.cuda.synchronize()
print("done")
