"""Data-parallel training step for one process per GPU (replaces the reference's single-process
``nn.DataParallel``, utils/helpers.py:333-335): patch pairs shard over ranks, weights are replicated, and after the
local backward ONE NCCL all-reduce over a flat fp32 bucket averages all gradients and the BatchNorm running
statistics.  BatchNorm normalises with per-rank batch statistics, like nn.DataParallel's replicas.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class DataParallelStep:
    def __init__(self, model, process_group=None):
        self.model = model
        self.pg = process_group
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.stats = [b for n, b in model.named_buffers() if n.endswith("running_mean") or n.endswith("running_var")]
        n = sum(p.numel() for p in self.params) + sum(b.numel() for b in self.stats)
        dev = self.params[0].device
        self.bucket = torch.empty(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for t in self.params + self.stats:
            self.views.append(self.bucket[off:off + t.numel()].view_as(t))
            off += t.numel()

    @property
    def world(self):
        return dist.get_world_size(self.pg) if dist.is_initialized() else 1

    def sync(self):
        """Call after loss.backward(): average gradients and BN running statistics across ranks with ONE all-reduce."""
        if self.world == 1:
            return
        srcs = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params] + list(self.stats)
        torch._foreach_copy_(self.views, srcs)
        dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.pg)
        self.bucket.mul_(1.0 / self.world)
        np_ = len(self.params)
        torch._foreach_copy_([p.grad for p in self.params if p.grad is not None],
                             [v for p, v in zip(self.params, self.views[:np_]) if p.grad is not None])
        torch._foreach_copy_(self.stats, self.views[np_:])

    def broadcast_parameters(self, src: int = 0):
        """Make every rank start from rank `src`'s weights (nn.DataParallel broadcasts replica 0 each forward)."""
        if self.world == 1:
            return
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            dist.broadcast(t.data, src, group=self.pg)
