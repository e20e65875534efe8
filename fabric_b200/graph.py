"""The whole training step as ONE CUDA graph (CUDA streams and graphs instead of a tracing compiler).

A training step of BiDateNet is ~180 kernel launches; at the benchmark shape (64 pairs of 13x256x256, ~27 ms of GPU work)
the host keeps ahead of the device, but at the shapes the reference actually trains with by default (patch 90, batch 32:
metadata.json:32-33,40) or at BASELINE configs[0] (2 pairs of 13x32x32) the step is bound by the host: every launch plans
its tiling, encodes up to five TMA descriptors and crosses ctypes.  ``GraphedTrainStep`` captures forward + loss + backward
+ the fused optimizer update (reference train.py:88-95) once -- descriptors, grids and workspace addresses are baked into
the graph's kernel nodes, torch's allocator serves the capture from a private pool -- and replays it with one
``cudaGraphLaunch`` per step.  Inputs are copied into static buffers (the only per-step work on the stream besides the
graph); the loss comes back as a device scalar.

Single-GPU and data-parallel (NCCL collectives are capturable) alike; the batch shape is fixed at capture time.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, criterion, dp, lr: float, example, warmup: int = 3):
        """``example`` = (x_d1, x_d2, labels) device tensors of the batch shape to capture; ``dp`` the
        ``fabric_b200.distributed.DataParallelStep`` that owns the model's gradients and the fused update."""
        self.model, self.criterion, self.dp, self.lr = model, criterion, dp, float(lr)
        dev = example[0].device
        self.device = dev
        self.static = [torch.empty_like(t) for t in example]
        for s, t in zip(self.static, example):
            s.copy_(t)
        # the step changes parameters and running statistics: capture must start from (and leave) the state it found
        snapshot = {k: v.detach().clone() for k, v in model.state_dict().items()}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(snapshot[k])
        if dp._managed:
            dp._manage_packed()         # repack from the restored weights, in place (the graph holds these addresses)
        dp.invalidate()
        self.replays = 0

    def _eager_step(self):
        self.dp.zero_grad()
        loss = self.criterion(self.model(self.static[0], self.static[1]), self.static[2])
        loss.backward()
        self.dp.sync_and_step(self.lr)
        return loss.detach()

    def __call__(self, x_d1, x_d2, labels):
        """one training step on this batch (same shapes / dtypes as the captured example); returns the loss (device scalar,
        valid until the next call)"""
        for s, t in zip(self.static, (x_d1, x_d2, labels)):
            if s.shape != t.shape or s.dtype != t.dtype:
                raise ValueError(f"GraphedTrainStep was captured for {tuple(s.shape)} {s.dtype}, got {tuple(t.shape)} {t.dtype}")
            if t.data_ptr() != s.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        self.dp.invalidate()        # eval-mode caches keyed on tensor versions cannot see the graph's writes
        return self.loss
