"""Data-parallel training step for one process per GPU (replaces the reference's single-process
``nn.DataParallel``, utils/helpers.py:333-335): patch pairs shard over ranks, weights are replicated, and the local
backward is followed by the NCCL all-reduce of a flat fp32 bucket holding all gradients and the BatchNorm running
statistics.  BatchNorm normalises with per-rank batch statistics, like nn.DataParallel's replicas.

No eager PyTorch on the step's hot path:

* the bucket IS the gradient storage: ``p.grad`` of every parameter is a view of it and the backward kernels (wgrad
  reduce, BatchNorm backward, 1x1 head backward) write their results straight into those views
  (``fabric_b200.autograd``: gradient sink), so nothing is packed or unpacked around the collective;
* the BatchNorm running statistics LIVE in the bucket's tail (the module buffers are re-pointed at it), so they are
  averaged by the same all-reduce in place;
* the bucket is laid out in BACKWARD order and cut into a few segments; the all-reduce of a segment is launched (async, on
  NCCL's stream) as soon as its last gradient has been written, so only the small encoder tail is exposed after backward;
* ONE kernel then applies plain SGD to every parameter straight from the reduced bucket, scales the running statistics by
  1/world, and refreshes the packed bf16 copies of the conv weights (forward and data-gradient layouts) that the next
  step's tcgen05 kernels read (``fabric_b200_train_step_update``).

Exact-global modes (off by default -- they put latency-bound collectives on the critical path):
``sync_bn=True`` all-reduces the BatchNorm moment sums (forward) and the (sum dy, sum dy*xhat) pairs (backward), which
makes N ranks x B pairs identical to 1 rank x N*B pairs up to reduction order.
"""
from __future__ import annotations

import struct

import torch
import torch.distributed as dist

# close an all-reduce segment once it holds at least this many elements (the last one takes the rest + the BN statistics)
SEGMENT_MIN_ELEMS = 1 << 20


def _block_of(name: str) -> str:
    return name.split(".", 1)[0]


class DataParallelStep:
    def __init__(self, model, process_group=None, overlap: bool = True, manage_packed: bool = True, exact: bool = False):
        self.model = model
        self.pg = process_group
        # exact-global mode: SyncBN + loss on the global batch; gradients are then SUMMED (each rank holds its share of the
        # global gradient), BatchNorm affine gradients arrive pre-divided by world (see ops.bn_relu_bwd)
        self.exact = bool(exact)
        if self.exact:
            from . import ops
            ops.EXACT = ops.ExactGlobal(process_group)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        # backward order for BiDateNet (so that finished gradients form a growing prefix of the bucket)
        try:
            from .autograd import BACKWARD_ORDER
            order = {b: i for i, b in enumerate(BACKWARD_ORDER)}
            if all(_block_of(n) in order for n, _ in named):
                named.sort(key=lambda np_: order[_block_of(np_[0])])
            else:
                order = None
        except Exception:
            order = None
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.stats = [b for n, b in model.named_buffers() if n.endswith("running_mean") or n.endswith("running_var")]
        n_grad = sum(p.numel() for p in self.params)
        n = n_grad + sum(b.numel() for b in self.stats)
        dev = self.params[0].device
        self.device = dev
        self.bucket = torch.zeros(n, dtype=torch.float32, device=dev)
        self.n_grad = n_grad
        self.views, self.offsets = [], []
        off = 0
        for t in self.params + self.stats:
            self.views.append(self.bucket[off:off + t.numel()].view_as(t))
            self.offsets.append(off)
            off += t.numel()
        np_ = len(self.params)
        # gradients: p.grad IS the bucket view (zeroed once; conv biases in front of a train-mode BN stay zero forever)
        for p, v in zip(self.params, self.views[:np_]):
            p.grad = v
        self.sink = {p: v for p, v in zip(self.params, self.views[:np_])}
        # BatchNorm running statistics move into the bucket tail: one all-reduce averages them in place
        with torch.no_grad():
            for b, v in zip(self.stats, self.views[np_:]):
                v.copy_(b)
                b.set_(v)
        # all-reduce segments [lo, hi) over the bucket, closed at block boundaries (BiDateNet) in backward order
        self.segments, self.seg_after_block = [], {}
        if order is not None and overlap:
            lo, cur = 0, 0
            blocks = [_block_of(nm) for nm in self.names]
            for i, p in enumerate(self.params):
                cur += p.numel()
                last_of_block = i + 1 == len(blocks) or blocks[i + 1] != blocks[i]
                if last_of_block and cur - lo >= SEGMENT_MIN_ELEMS and i + 1 < len(blocks):
                    self.seg_after_block[blocks[i]] = len(self.segments)
                    self.segments.append((lo, cur))
                    lo = cur
            self.segments.append((lo, n))            # the tail: remaining gradients + the BN statistics
        else:
            self.segments.append((0, n))
        self._pending = []
        self._managed = False
        self._table = None
        model.__dict__["_fb_dp"] = self
        if manage_packed and dev.type == "cuda":
            self._manage_packed()

    def use_compute_stream(self):
        """Make a HIGH-PRIORITY stream the current stream of this device (call once, before the training loop).  The backward
        pass puts the weight-gradient launches on a default-priority side stream (fabric_b200.autograd); with the main chain
        (data gradient -> BatchNorm backward -> ...) on a high-priority stream its kernels are scheduled first whenever both
        wait for SMs, and the weight gradients fill in behind them.  Measured: -0.13 .. -0.24 ms per step (three A/B pairs,
        tools/ab.sh FABRIC_B200_HIPRI).  Returns the stream."""
        if self.device.type != "cuda":
            return None
        if getattr(self, "compute_stream", None) is None:
            self.compute_stream = torch.cuda.Stream(self.device, priority=-1)
        cur = torch.cuda.current_stream(self.device)
        if cur != self.compute_stream:
            self.compute_stream.wait_stream(cur)
            torch.cuda.set_stream(self.compute_stream)
        return self.compute_stream

    # ------------------------------------------------------------------------------------------------ basics
    @property
    def world(self):
        return dist.get_world_size(self.pg) if dist.is_initialized() else 1

    def zero_grad(self):
        """Gradients are OVERWRITTEN by the fabric_b200 backward kernels (never accumulated), so there is nothing to clear
        when the model is a fabric_b200.BiDateNet.  For any other model (plain autograd accumulates into p.grad) the
        gradient part of the bucket is zeroed with one memset."""
        if not self._fabric_model():
            self.bucket[:self.n_grad].zero_()
        for p, v in self.sink.items():
            if p.grad is not v:
                p.grad = v

    def _fabric_model(self):
        from .bidate_model import BiDateNet
        return isinstance(self.model, BiDateNet)

    def wants_block(self, name: str) -> bool:
        """does an all-reduce segment end with block ``name`` (and is there anyone to reduce with)?"""
        return self.world > 1 and name in self.seg_after_block

    def block_done(self, name: str):
        """Called by the backward pass when every gradient of block ``name`` has been written: launches the all-reduce of
        the bucket segment that ends there (async; overlaps the rest of backward)."""
        i = self.seg_after_block.get(name)
        if i is None or self.world == 1:
            return
        lo, hi = self.segments[i]
        self._pending.append(dist.all_reduce(self.bucket[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def _gather_stray_grads(self):
        """Make p.grad the bucket view again.  fabric_b200 models: the backward kernels already wrote into the views
        (whatever p.grad currently is -- e.g. None after ``optimizer.zero_grad(set_to_none=True)``), so only re-attach.
        Plain-autograd models: copy gradients that live elsewhere into the bucket (not the fabric_b200 hot path)."""
        fabric = self._fabric_model()
        for p, v in self.sink.items():
            if p.grad is v:
                continue
            if not fabric:
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)
            p.grad = v

    def _reduce(self):
        """finish the step's collective: the segments not yet launched by block_done (the tail), then wait for all"""
        if self.world == 1:
            return
        launched = len(self._pending)
        if launched == 0:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.pg)        # ONE all-reduce (no overlap hooks ran)
        else:
            for lo, hi in self.segments[launched:]:
                self._pending.append(dist.all_reduce(self.bucket[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
            for w in self._pending:
                w.wait()
        self._pending = []

    def sync(self):
        """Call after loss.backward(): average gradients and BN running statistics across ranks (generic path: the
        result is left in p.grad / the module buffers for any torch optimizer)."""
        self._gather_stray_grads()
        if self.world == 1:
            return
        self._reduce()
        if self.exact:
            self.bucket[self.n_grad:].mul_(1.0 / self.world)     # gradients are already the global ones; statistics: mean
        else:
            self.bucket.mul_(1.0 / self.world)

    # ------------------------------------------------------------------------------------------------ fused step
    def _manage_packed(self):
        """Own the packed bf16 copies of every conv weight (forward [Cout][9][CinPad] and data-gradient [Cin][9][Cout]
        layouts): the fused update kernel rewrites them in place after each step, so a training step launches no pack
        kernels at all."""
        from . import ops
        from .unet_parts import double_conv
        self._packed = {}
        fresh = False
        for m in self.model.modules():
            if isinstance(m, double_conv):
                old = m.__dict__.get("_fb_managed") or {}
                fresh = fresh or not old
                managed = {}
                for idx in (0, 3):
                    w = m.conv[idx].weight
                    # refresh IN PLACE when the copies exist: a captured CUDA graph has their addresses baked in
                    managed[("w", idx)] = ops.pack_conv_weight(w, 0, out=old.get(("w", idx)))
                    managed[("wd", idx)] = ops.pack_conv_weight(w, 1, out=old.get(("wd", idx)))
                    managed[("v", idx)] = w._version
                    self._packed[w] = (managed[("w", idx)], managed[("wd", idx)])
                m.__dict__["_fb_managed"] = managed
        self._managed = True
        if fresh:
            self._table = None          # new buffers: the update kernel's record table must be rebuilt

    def _update_table(self):
        """device table of 64-byte records for fabric_b200_train_step_update (<= 64 Ki elements each), built once"""
        if self._table is not None:
            return self._table
        recs = bytearray()
        n_chunks = 0
        np_ = len(self.params)

        def emit(pptr, gptr, n, mode, wf=0, wd=0, cout=0, cin=0, cinpad=0):
            nonlocal n_chunks, recs
            off = 0
            while off < n:
                k = min(65536, n - off)
                recs += struct.pack("<QQQQiiiiiiii", pptr + 4 * off, gptr + 4 * off, wf, wd, k, mode, off, cout, cin, cinpad, 0, 0)
                off += k
                n_chunks += 1
        def emit_tiles(pptr, gptr, wf, wd, cout, cin, cinpad):
            # mode 3: one record per (16 output channels x <= 128 input channels) tile of a 3x3 conv weight
            nonlocal n_chunks, recs
            for co0 in range(0, cout, 16):
                for ci0 in range(0, cin, 128):
                    recs += struct.pack("<QQQQiiiiiiii", pptr, gptr, wf, wd, min(16, cout - co0), 3, co0, cout, cin, cinpad,
                                        ci0, min(128, cin - ci0))
                    n_chunks += 1
        for p, v in zip(self.params, self.views[:np_]):
            pk = getattr(self, "_packed", {}).get(p) if self._managed else None
            if pk is not None:
                cout, cin = p.shape[0], p.shape[1]
                emit_tiles(p.data_ptr(), v.data_ptr(), pk[0].data_ptr(), pk[1].data_ptr(), cout, cin, pk[0].shape[2])
            else:
                emit(p.data_ptr(), v.data_ptr(), p.numel(), 0)
        n_sgd = n_chunks
        for b, v in zip(self.stats, self.views[np_:]):
            emit(v.data_ptr(), v.data_ptr(), v.numel(), 1)
        t = torch.frombuffer(recs, dtype=torch.uint8).clone().to(self.device)
        self._table = (t, n_sgd, n_chunks)
        return self._table

    def sync_and_step(self, lr: float):
        """loss.backward() -> this.  Finishes the all-reduce (skipped for world == 1), then ONE kernel: plain SGD
        p -= lr/world * g straight from the reduced bucket (reference train.py:55,95: optim.SGD(lr), no momentum / weight
        decay), running statistics *= 1/world, packed bf16 conv weights refreshed."""
        from . import _lib, ops
        if self.device.type != "cuda":
            raise RuntimeError("sync_and_step needs the CUDA library; use sync() + torch.optim.SGD on CPU")
        self._gather_stray_grads()
        world = self.world
        self._reduce()
        table, n_sgd, n_all = self._update_table()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().fabric_b200_train_step_update(
                table.data_ptr(), n_all if world > 1 else n_sgd, float(lr), 1.0 if self.exact else 1.0 / world, 1.0 / world,
                torch.cuda.current_stream().cuda_stream), "train_step_update")
        ops._count()
        self.invalidate()

    def close(self):
        """detach from the model (and leave exact-global mode)"""
        from . import ops
        if self.exact:
            ops.EXACT = None
        self.model.__dict__.pop("_fb_dp", None)

    def invalidate(self):
        """the kernel updated the parameters behind torch's back (no version bump): drop version-keyed caches"""
        mods = self.__dict__.get("_mods")
        if mods is None:      # (the module tree is fixed: walking it every step cost 0.3 ms at small shapes)
            mods = self._mods = list(self.model.modules())
        for m in mods:
            c = m.__dict__.get("_fb_cache")
            if c is not None:
                c.clear()

    def broadcast_parameters(self, src: int = 0):
        """Make every rank start from rank `src`'s weights (nn.DataParallel broadcasts replica 0 each forward)."""
        if self.world > 1:
            with torch.no_grad():
                for t in list(self.model.parameters()) + list(self.model.buffers()):
                    dist.broadcast(t, src, group=self.pg)
        self.invalidate()
        if self._managed:
            self._manage_packed()          # repack from the received weights
