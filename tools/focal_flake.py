"""Development probe: repeat the fused focal / CE loss on the test's inputs and report any run that differs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fabric_b200 import ops
from oracle import bidatenet_oracle as O
torch.manual_seed(1)
B, H, W = 3, 24, 40
logits = torch.randn(B, 2, H, W) * 2
labels = (torch.rand(B, H, W) < 0.2).long()
v = float(O.focal_loss(logits, labels, 2.0))
lg, lb = logits.cuda(), labels.cuda()
vals = []
for i in range(200):
    if i % 3 == 0:
        ops.seg_loss_fwd_bwd("tversky", lg, lb, 0.1, 0.9, 2.0, 1e-7)
    loss, dl = ops.seg_loss_fwd_bwd("focal", lg, lb, 0.1, 0.9, 2.0, 1e-7)
    vals.append(float(loss))
print("oracle", v, "gpu min/max", min(vals), max(vals), "distinct", sorted(set(vals))[:6], "cpu threads", torch.get_num_threads())
