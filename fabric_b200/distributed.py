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

    # ---- fused optimizer step -------------------------------------------------------------
    def _chunk_table(self, grads):
        """device table of {param ptr, grad ptr, n} records (<= 64 Ki elements each) for fabric_b200_sgd_step"""
        import struct
        key = tuple(g.data_ptr() for g in grads)
        if getattr(self, "_chunk_key", None) == key:
            return self._chunks, self._n_chunks
        recs = bytearray()
        n_chunks = 0
        for p, g in zip(self.params, grads):
            n, off = p.numel(), 0
            while off < n:
                k = min(65536, n - off)
                recs += struct.pack("<QQii", p.data_ptr() + 4 * off, g.data_ptr() + 4 * off, k, 0)
                off += k
                n_chunks += 1
        t = torch.frombuffer(recs, dtype=torch.uint8).clone()
        self._chunks, self._n_chunks, self._chunk_key = t.to(self.params[0].device), n_chunks, key
        return self._chunks, n_chunks

    def sync_and_step(self, lr: float):
        """loss.backward() -> this: gradients (+ BN running statistics) go into the flat bucket, ONE all-reduce (skipped
        for world == 1), then ONE multi-tensor SGD kernel reads the reduced gradients straight from the bucket
        (p -= lr/world * g).  Replaces dp.sync(); optimizer.step() for the reference's plain SGD (train.py:55,95)."""
        from . import _lib, ops
        if not self.params[0].is_cuda:
            raise RuntimeError("sync_and_step needs the CUDA library; use sync() + torch.optim.SGD on CPU")
        np_ = len(self.params)
        world = self.world
        torch._foreach_copy_(self.views[:np_], [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params])
        if world > 1:
            torch._foreach_copy_(self.views[np_:], list(self.stats))
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.pg)
            torch._foreach_mul_(self.views[np_:], 1.0 / world)
            torch._foreach_copy_(self.stats, self.views[np_:])
        chunks, n = self._chunk_table(self.views[:np_])       # bucket pointers never change: built once
        _lib.check(_lib.load().fabric_b200_sgd_step(chunks.data_ptr(), n, float(lr), 1.0 / world,
                                                    torch.cuda.current_stream().cuda_stream), "sgd_step")
        ops._count()
        self._bump()

    def _bump(self):
        # the kernel updated the parameters behind torch's back (no version bump): drop packed-weight caches
        for m in self.model.modules():
            c = m.__dict__.get("_fb_cache")
            if c is not None:
                c._store.clear()

    def broadcast_parameters(self, src: int = 0):
        """Make every rank start from rank `src`'s weights (nn.DataParallel broadcasts replica 0 each forward)."""
        if self.world == 1:
            return
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            dist.broadcast(t.data, src, group=self.pg)
