"""Host-buffer inference loop: the B200 version of the reference's full-scene prediction loop
(train.py:187-201): patches live in HOST memory (numpy / pinned torch), are copied to the device in
sub-batches, run through ``BiDateNet`` and the logits (or the argmax change mask, train.py:199) come back to the
host.  Unlike the reference's loop (synchronous ``.to(dev)`` -> forward -> ``.cpu()`` per batch), the copies of
sub-batch i+1 and the read-back of sub-batch i-1 overlap the compute of sub-batch i on separate CUDA streams.
"""
from __future__ import annotations

from typing import Optional

import torch


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated: first-touch then places the staging buffers in that node's memory, and the threads that issue the copies
    run next to it.  With eight ranks on a two-socket host, unbound ranks put every staging buffer on the node the
    launcher happened to run on and the H2D streams of the far GPUs cross the socket interconnect.  Best effort (sysfs):
    returns the node id, or None when the topology cannot be read / there is a single node."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        if cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class HostPipeline:
    """Reusable pinned staging + streams for ``predict_patches``."""

    def __init__(self, model, chunk: int = 16, n_channels: int = 13, size: int = 256, return_logits: bool = True):
        self.model = model
        self.chunk = chunk
        self.dev = next(model.parameters()).device
        self.copy_in = torch.cuda.Stream(self.dev)
        self.copy_out = torch.cuda.Stream(self.dev)
        self.return_logits = return_logits
        self.shape = (chunk, n_channels, size, size)
        self.bufs = [dict(x1=torch.empty(self.shape, device=self.dev), x2=torch.empty(self.shape, device=self.dev),
                          ready=torch.cuda.Event(), done=torch.cuda.Event(), out=None, copied=torch.cuda.Event())
                     for _ in range(2)]

    @torch.no_grad()
    def run(self, x1_host: torch.Tensor, x2_host: torch.Tensor, out_host: torch.Tensor):
        """x*_host: [N,C,S,S] fp32 (pinned for async copies); out_host: [N,2,S,S] fp32 or [N,S,S] uint8 (mask)."""
        n = x1_host.shape[0]
        main = torch.cuda.current_stream(self.dev)
        if self.bufs[0]["x1"].dtype != x1_host.dtype:   # raw uint16 rasters (model.set_input_normalisation) or fp32 patches
            main.synchronize()
            for b in self.bufs:
                b["x1"] = torch.empty(self.shape, dtype=x1_host.dtype, device=self.dev)
                b["x2"] = torch.empty(self.shape, dtype=x1_host.dtype, device=self.dev)
        h2d = d2h = 0
        for i, lo in enumerate(range(0, n, self.chunk)):
            hi = min(lo + self.chunk, n)
            b = self.bufs[i & 1]
            k = hi - lo
            with torch.cuda.stream(self.copy_in):
                self.copy_in.wait_event(b["done"])          # compute that last used this buffer has finished
                b["x1"][:k].copy_(x1_host[lo:hi], non_blocking=True)
                b["x2"][:k].copy_(x2_host[lo:hi], non_blocking=True)
                b["ready"].record(self.copy_in)
            h2d += 2 * x1_host[lo:hi].numel() * x1_host.element_size()
            main.wait_event(b["ready"])
            main.wait_event(b["copied"])                    # previous result of this slot has left the device
            logits = self.model(b["x1"][:k], b["x2"][:k])
            res = logits if self.return_logits else logits.argmax(1).to(torch.uint8)
            b["out"] = res
            b["done"].record(main)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(b["done"])
                out_host[lo:hi].copy_(res, non_blocking=True)
                b["copied"].record(self.copy_out)
            res.record_stream(self.copy_out)
            d2h += res.numel() * res.element_size()
        main.wait_stream(self.copy_out)
        return h2d, d2h


class BatchFeeder:
    """Host -> device hand-off for the training loop (reference train.py:80-84: the DataLoader yields host batches and the
    loop calls ``.to(device)`` synchronously before every step).  Here the copy of batch i+1 runs on a side stream into
    the other half of a double buffer while step i computes; ``next()`` returns device tensors that are safe to use on
    the current stream.  Feed it pinned host tensors (``DataLoader(pin_memory=True)``)."""

    def __init__(self, device):
        self.dev = torch.device(device)
        self.stream = torch.cuda.Stream(self.dev)
        self.slots = [dict(t=None, ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
        self.i = 0
        self.pending = None

    def _stage(self, host_batch):
        slot = self.slots[self.i & 1]
        self.i += 1
        if slot["t"] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot["t"], host_batch)):
            slot["t"] = [torch.empty(h.shape, dtype=h.dtype, device=self.dev) for h in host_batch]
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(slot["free"])       # the step that last read this slot has finished
            for d, h in zip(slot["t"], host_batch):
                d.copy_(h, non_blocking=True)
            slot["ready"].record(self.stream)
        return slot

    def prefetch(self, host_batch):
        """Start copying ``host_batch`` (a tuple of host tensors); returns the number of bytes queued."""
        self.pending = self._stage(host_batch)
        return sum(h.numel() * h.element_size() for h in host_batch)

    def next(self):
        """Device tensors of the batch given to the last ``prefetch``; call ``release()`` after the step is enqueued."""
        slot, self.pending = self.pending, None
        torch.cuda.current_stream(self.dev).wait_event(slot["ready"])
        self.current = slot
        return slot["t"]

    def release(self):
        self.current["free"].record(torch.cuda.current_stream(self.dev))


def predict_patches(model, p1, p2, batch_size: int = 16, return_logits: bool = False,
                    pipeline: Optional[HostPipeline] = None):
    """Drop-in for the loop at reference train.py:187-201: ``p1``/``p2`` are host arrays [N,13,S,S] fp32 (as
    produced by ``generate_patches``); returns the host change mask [N,S,S] uint8 (or fp32 logits)."""
    x1 = torch.as_tensor(p1)
    x2 = torch.as_tensor(p2)
    if not x1.is_pinned():
        x1, x2 = x1.pin_memory(), x2.pin_memory()
    n, c, s, _ = x1.shape
    if pipeline is None:
        pipeline = HostPipeline(model, batch_size, c, s, return_logits)
    shape = (n, 2, s, s) if return_logits else (n, s, s)
    out = torch.empty(shape, dtype=torch.float32 if return_logits else torch.uint8).pin_memory()
    pipeline.run(x1, x2, out)
    torch.cuda.current_stream().synchronize()
    return out
