"""Host-side mirror of the reference's ``models/bidate_model.py`` (BiDateNet): same constructor, attribute
names and ``state_dict`` keys; ``forward(x_d1, x_d2)`` takes NCHW fp32 [B,13,H,W] pairs and returns NCHW fp32
logits [B,2,H,W] exactly like reference models/bidate_model.py:22-40, but runs as ~25 sm_100a kernel launches:

* both dates go through the weight-shared encoder as ONE launch per conv (date group dim G=2; in training
  mode the BatchNorm moments are still kept per date, as in the reference which calls the encoder twice);
* MaxPool2d is fused into the producing conv's epilogue, relu(d2*d1) + bilinear upsample + pad + concat into one
  decoder-input kernel, outconv into the last conv's epilogue.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .unet_parts import _CacheInvalidation, down, outconv, up, inconv


class BiDateNet(_CacheInvalidation, nn.Module):
    def __init__(self, n_channels, n_classes):
        super(BiDateNet, self).__init__()
        if n_channels > 16:
            raise NotImplementedError("fabric_b200 BiDateNet supports up to 16 input bands (the reference uses 13)")
        self.inc = inconv(n_channels, 64)
        self.down1 = down(64, 128)
        self.down2 = down(128, 256)
        self.down3 = down(256, 512)
        self.down4 = down(512, 512)

        self.up1 = up(1024, 256)
        self.up2 = up(512, 128)
        self.up3 = up(256, 64)
        self.up4 = up(128, 64)
        self.outc = outconv(64, n_classes)
        self.fuse_head = True
        self.fuse_product = True

    def __setstate__(self, state):
        """a whole-model pickle written by the reference (train.py:222) has none of this repo's extra attributes"""
        super().__setstate__(state)
        self.__dict__.setdefault("fuse_head", True)
        self.__dict__.setdefault("fuse_product", True)

    def __getstate__(self):
        s = dict(self.__dict__)
        s.pop("_fb_dp", None)          # a DataParallelStep (process group, bucket) is not part of the model
        return s

    def set_input_normalisation(self, mean, std):
        """Per-band mean / std of the loader's z-score (reference utils/dataloaders.py:94-99, constants in
        metadata.json:4-29).  With them set, ``forward`` also accepts RAW uint16 rasters [B,13,H,W]: the normalisation runs
        inside the pack kernel and the host ships half the bytes.  Not part of ``state_dict`` (the reference has no such
        entry), but the tensors follow ``.to(device)``."""
        mean = torch.as_tensor(mean, dtype=torch.float32).flatten()
        std = torch.as_tensor(std, dtype=torch.float32).flatten()
        dev = next(self.parameters()).device
        self.register_buffer("_fb_in_mean", mean.to(dev).contiguous(), persistent=False)
        self.register_buffer("_fb_in_inv_std", (1.0 / std).to(dev).contiguous(), persistent=False)

    def pack_pair(self, x_d1, x_d2, aug=None):
        """Both dates into one NHWC5 bf16 tensor [2,B,H,W,16].  ``aug`` int32 [B,3] (rot90 turns, flip rows, flip columns)
        applies the loader's augmentation while packing (use ``ops.augment_labels`` with the same rows for the labels)."""
        b, c, h, w = x_d1.shape
        x5 = ops._empty((2, b, h, w, ops.cpad(c)), dtype=torch.bfloat16, device=x_d1.device)
        if aug is not None:
            mean, inv_std = getattr(self, "_fb_in_mean", None), getattr(self, "_fb_in_inv_std", None)
            if x_d1.dtype == torch.uint16 and mean is None:
                raise RuntimeError("raw uint16 input needs BiDateNet.set_input_normalisation(mean, std) first")
            if x_d1.dtype != torch.uint16:
                mean = inv_std = None
            ops.pack_input_aug(x_d1.contiguous(), aug, mean, inv_std, out=x5[0])
            ops.pack_input_aug(x_d2.contiguous(), aug, mean, inv_std, out=x5[1])
            return x5
        if x_d1.dtype == torch.uint16:
            mean, inv_std = getattr(self, "_fb_in_mean", None), getattr(self, "_fb_in_inv_std", None)
            if mean is None:
                raise RuntimeError("raw uint16 input needs BiDateNet.set_input_normalisation(mean, std) first")
            ops.pack_input_raw(x_d1.contiguous(), mean, inv_std, out=x5[0])
            ops.pack_input_raw(x_d2.contiguous(), mean, inv_std, out=x5[1])
            return x5
        ops.pack_input(x_d1.contiguous(), out=x5[0])
        ops.pack_input(x_d2.contiguous(), out=x5[1])
        return x5

    def forward_packed(self, x5):
        """Eval-mode forward on an already packed pair tensor; returns NCHW fp32 logits."""
        if not getattr(self, "fuse_product", True):
            return self._forward_packed_unfused(x5)
        _, b, h, w, _ = x5.shape
        dev = x5.device

        def cat(level, cs, cl):     # decoder input [1,B,H/2^l,W/2^l,cs+cl]; skip half filled by the encoder epilogue
            return ops._empty((1, b, h >> level, w >> level, cs + cl), dtype=torch.bfloat16, device=dev)
        cat4, cat3, cat2, cat1 = cat(0, 64, 64), cat(1, 128, 128), cat(2, 256, 256), cat(3, 512, 512)
        # the full-resolution encoder outputs are consumed only through their pooled copy and the fused product, so the
        # 64- and 128-wide levels do not write them at all (the date-0 tile waits in shared memory for its partner)
        e1 = self.inc.run5(x5, pool=True, prod_out=cat4, keep_main=False)      # models/bidate_model.py:23,29 (+ :38 skip)
        e2 = self.down1.run5(e1["pool"], pool=True, prod_out=cat3, keep_main=False)   # :24,30 (+ :37)
        e3 = self.down2.run5(e2["pool"], pool=True, prod_out=cat2)             # :25,31 (+ :36)
        e4 = self.down3.run5(e3["pool"], pool=True, prod_out=cat1)             # :26,32 (+ :35)
        e5 = self.down4.run5(e4["pool"])                                       # :27,33
        x = self.up1.run5(e5["y"], None, cat5=cat1)["y"]                       # :35
        x = self.up2.run5(x, None, cat5=cat2)["y"]                             # :36
        x = self.up3.run5(x, None, cat5=cat3)["y"]                             # :37
        if self.fuse_head:
            return self.up4.run5(x, None, cat5=cat4, head=self.outc.head(), keep_main=False)["logits"]   # :38-39
        x = self.up4.run5(x, None, cat5=cat4)["y"]                             # :38
        return self.outc.run5(x)                                               # :39

    def _forward_packed_unfused(self, x5):
        """Same network with the decoder inputs built by the stand-alone kernel (cross-check of the fused epilogue)."""
        e1 = self.inc.run5(x5, pool=True)
        e2 = self.down1.run5(e1["pool"], pool=True)
        e3 = self.down2.run5(e2["pool"], pool=True)
        e4 = self.down3.run5(e3["pool"], pool=True)
        e5 = self.down4.run5(e4["pool"])
        x = self.up1.run5(e5["y"], e4["y"])["y"]
        x = self.up2.run5(x, e3["y"])["y"]
        x = self.up3.run5(x, e2["y"])["y"]
        if self.fuse_head:
            return self.up4.run5(x, e1["y"], head=self.outc.head(), keep_main=False)["logits"]
        x = self.up4.run5(x, e1["y"])["y"]
        return self.outc.run5(x)

    def forward(self, x_d1, x_d2, aug=None):
        """Reference signature ``forward(x_d1, x_d2)`` (models/bidate_model.py:22).  Optional ``aug`` int32 [B,3]: the
        loader's rot90 / flip augmentation fused into the input pack (see ``pack_pair``)."""
        if x_d1.shape != x_d2.shape:
            raise ValueError("x_d1 and x_d2 must have the same shape")
        if not x_d1.is_cuda:
            raise RuntimeError("fabric_b200.BiDateNet runs on sm_100 CUDA devices only (no CPU fallback); "
                               "call .cuda() on the model and inputs")
        if getattr(self, "_is_replica", False) or next(self.parameters(), None) is None:
            raise RuntimeError("this BiDateNet is an nn.DataParallel replica (reference utils/helpers.py:335 wraps the model): "
                               "fabric_b200 runs one process per GPU -- drop the nn.DataParallel wrapper and use "
                               "fabric_b200.distributed.DataParallelStep under torchrun")
        if self.training and torch.is_grad_enabled():
            from .autograd import bidatenet_train_forward
            return bidatenet_train_forward(self, x_d1, x_d2, aug)
        if self.training:
            raise RuntimeError("BiDateNet in .train() mode under torch.no_grad(): call .eval() for inference "
                               "(batch-statistics BatchNorm without a backward pass is not a path the reference uses)")
        with torch.cuda.device(x_d1.device):
            return self.forward_packed(self.pack_pair(x_d1, x_d2, aug))
