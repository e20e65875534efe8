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
torch.cuda.synchronize()
print("done")
