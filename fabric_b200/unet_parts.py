"""Host-side mirror of the reference's ``models/unet_parts.py``: same class names, constructor signatures,
sub-module names and ``state_dict`` keys (so ``torch.save(model)`` / published weights round-trip), but the
arithmetic runs in the sm_100a kernels behind the C ABI (``include/fabric_b200.h``).

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` objects below are parameter containers only -- their ``forward`` is never
called.  Each block has two entry points:

* ``forward(x)``: the reference's calling convention, NCHW fp32 in / NCHW fp32 out (reference
  models/unet_parts.py:21-23,31-33,44-46,64-80,88-90);
* ``run5(...)``: the internal fast path on NHWC5 bf16 tensors used by ``BiDateNet.forward`` so that no layout
  change happens between blocks.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _PackCache:
    """Packed bf16 weights and folded BN (scale, shift), rebuilt when the source tensors change.

    The key is (data_ptr, _version, device) of every source tensor.  ``_version`` does NOT move for writes through
    ``.data`` / raw kernels (``p.data.copy_``, EMA updates, ``dist.broadcast(t.data)``, this repo's fused SGD kernel), so
    the cache is ALSO dropped on every event after which such a write is plausible: ``module.train()`` / ``.eval()``,
    ``load_state_dict``, ``_apply`` (``.to`` / ``.cuda`` / ``.half``), and explicitly through ``invalidate_caches()``
    (``DataParallelStep`` calls it after ``broadcast_parameters`` and after every fused optimizer step).  A training-mode
    forward never trusts the cache at all: weights change every step, so it repacks (or uses the packed copies the fused
    optimizer kernel maintains, see ``fabric_b200.distributed``)."""

    def __init__(self):
        self._store = {}

    def get(self, key, tensors, build, extra=None):
        sig = tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors if t is not None) + (extra,)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        val = build()
        self._store[key] = (sig, val)
        return val

    def clear(self):
        self._store.clear()


class _CacheInvalidation:
    """Mixin for the block modules: drop packed-weight caches whenever the parameters may have been rewritten."""

    def invalidate_caches(self):
        for m in self.modules():
            c = m.__dict__.get("_fb_cache")
            if c is not None:
                c.clear()
        return self

    def train(self, mode=True):
        self.invalidate_caches()
        return super().train(mode)

    def _apply(self, fn, *a, **kw):
        self.invalidate_caches()
        for m in self.modules():
            m.__dict__.pop("_fb_managed", None)      # packed copies owned by a DataParallelStep point at the old storage
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        self.invalidate_caches()
        return r

    def _load_from_state_dict(self, *a, **kw):        # also reached when a PARENT module's load_state_dict recurses
        r = super()._load_from_state_dict(*a, **kw)
        c = self.__dict__.get("_fb_cache")
        if c is not None:
            c.clear()
        return r


class double_conv(_CacheInvalidation, nn.Module):
    '''(conv => BN => ReLU) * 2   -- reference models/unet_parts.py:8-23'''

    def __init__(self, in_ch, out_ch):
        super(double_conv, self).__init__()
        if out_ch % 64:
            raise NotImplementedError("fabric_b200 double_conv needs out_ch to be a multiple of 64")
        self.conv = nn.Sequential(
            nn.Conv2d(in_ch, out_ch, 3, padding=1),
            nn.BatchNorm2d(out_ch),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_ch, out_ch, 3, padding=1),
            nn.BatchNorm2d(out_ch),
            nn.ReLU(inplace=True)
        )
        self.in_ch, self.out_ch = in_ch, out_ch
        self.tune1 = None   # optional fb_conv_tuning overrides (dict), used by the benchmarks
        self.tune2 = None
        # eval mode: fold the BatchNorm scale into the packed weights and start the accumulator at the shift, so the conv
        # epilogue is convert + ReLU + store (False = per-element scale / shift in the epilogue, one packed copy for all)
        self.fold_bn = True

    # caches are not part of the module state (and must not be pickled)
    def _cache(self) -> _PackCache:
        c = self.__dict__.get("_fb_cache")
        if c is None:
            c = _PackCache()
            self.__dict__["_fb_cache"] = c
        return c

    def __getstate__(self):
        s = dict(self.__dict__)
        s.pop("_fb_cache", None)
        s.pop("_fb_managed", None)
        return s

    def __setstate__(self, state):
        """Whole-model pickles written by the REFERENCE (train.py:222 ``torch.save(model)``) resolve to this class through
        the ``models/`` shim but carry only the reference's attributes: derive the rest."""
        super().__setstate__(state)
        conv = self.__dict__["_modules"]["conv"]
        d = self.__dict__
        d.setdefault("in_ch", conv[0].weight.shape[1])
        d.setdefault("out_ch", conv[0].weight.shape[0])
        d.setdefault("tune1", None)
        d.setdefault("tune2", None)
        d.setdefault("fold_bn", True)

    def _packed(self, idx, training=False):
        """bf16 [Cout][9][CinPad] forward weights of conv ``idx``.  Training forwards never trust the version-keyed cache
        (see _PackCache): they use the copies the fused optimizer kernel keeps fresh, else repack."""
        conv = self.conv[idx]
        managed = self.__dict__.get("_fb_managed")
        if managed is not None and managed.get(("v", idx)) == conv.weight._version:
            return managed[("w", idx)]       # (a torch optimizer that updated the weight bumps _version: not ours any more)
        if training:
            return ops.pack_conv_weight(conv.weight, 0)
        return self._cache().get(("w", idx), [conv.weight], lambda: ops.pack_conv_weight(conv.weight, 0))

    def _packed_dgrad(self, idx):
        """bf16 [Cin][9][Cout] tap-flipped weights: conv3x3 with them is the data gradient of conv ``idx``"""
        conv = self.conv[idx]
        managed = self.__dict__.get("_fb_managed")
        if managed is not None and managed.get(("v", idx)) == conv.weight._version:
            return managed[("wd", idx)]
        return ops.pack_conv_weight(conv.weight, 1)

    def _folded(self, idx):
        conv, bn = self.conv[idx], self.conv[idx + 1]
        return self._cache().get(("bn", idx), [bn.weight, bn.bias, bn.running_mean, bn.running_var, conv.bias],
                                 lambda: ops.bn_fold_eval(bn, conv.bias), extra=bn.__dict__.get("_fb_stats_epoch", 0))

    def _packed_folded(self, idx):
        """(bf16 weights with the eval BatchNorm scale folded in, shift) for conv ``idx``."""
        conv, bn = self.conv[idx], self.conv[idx + 1]
        scale, shift = self._folded(idx)
        w = self._cache().get(("w_folded", idx), [conv.weight, bn.weight, bn.running_var],
                              lambda: ops.pack_conv_weight(conv.weight, 0, scale=scale),
                              extra=bn.__dict__.get("_fb_stats_epoch", 0))
        return w, shift

    def run5(self, x5, pool=False, head=None, keep_main=True, prod_out=None):
        """NHWC5 bf16 in -> dict(y=..., pool=..., logits=...).  Eval mode: BatchNorm (running statistics), the conv
        bias and ReLU ride in the conv epilogue; ``pool`` adds the fused MaxPool2d(2) copy for the next ``down``;
        ``head`` = (weight[2,64], bias[2]) fuses ``outconv`` into the second conv."""
        if self.training:
            raise RuntimeError("run5 is the eval fast path; training goes through fabric_b200.autograd "
                               "(BiDateNet.forward / double_conv.forward in .train() mode)")
        if getattr(self, "fold_bn", True):
            w1, h1 = self._packed_folded(0)
            w2, h2 = self._packed_folded(3)
            mid = ops.conv3x3(x5, w1, self.out_ch, None, h1, relu=True, tune=self.tune1, true_cin=self.in_ch,
                              shift_in_acc=True)["y"]
            return ops.conv3x3(mid, w2, self.out_ch, None, h2, relu=True, pool=pool, head=head, store_main=keep_main,
                               tune=self.tune2, prod_out=prod_out, shift_in_acc=True)
        s1, h1 = self._folded(0)
        s2, h2 = self._folded(3)
        mid = ops.conv3x3(x5, self._packed(0), self.out_ch, s1, h1, relu=True, tune=self.tune1, true_cin=self.in_ch)["y"]
        return ops.conv3x3(mid, self._packed(3), self.out_ch, s2, h2, relu=True, pool=pool, head=head,
                           store_main=keep_main, tune=self.tune2, prod_out=prod_out)

    def forward(self, x):
        """reference models/unet_parts.py:21-23: NCHW fp32 in / out; differentiable in training mode (batch statistics,
        running-stat update, gradients for x and all parameters)"""
        if not x.is_cuda:
            raise RuntimeError("fabric_b200 blocks run on sm_100 CUDA devices only (no CPU fallback)")
        if self.training:
            from .autograd import double_conv_train_forward
            return double_conv_train_forward(self, x)
        with torch.cuda.device(x.device):
            x5 = ops.pack_input(x, c_pad=ops.cpad(self.in_ch)).unsqueeze(0)
            y = self.run5(x5)["y"]
            return ops.unpack_output(y[0])


class inconv(_CacheInvalidation, nn.Module):
    '''reference models/unet_parts.py:26-33'''

    def __init__(self, in_ch, out_ch):
        super(inconv, self).__init__()
        self.conv = double_conv(in_ch, out_ch)

    def run5(self, x5, **kw):
        return self.conv.run5(x5, **kw)

    def forward(self, x):
        x = self.conv(x)
        return x


class down(_CacheInvalidation, nn.Module):
    '''MaxPool2d(2) => double_conv   -- reference models/unet_parts.py:36-46.
    On the fast path the pooling is done by the PREVIOUS block's conv epilogue (``pool=True``), so ``run5``
    takes the already pooled tensor.'''

    def __init__(self, in_ch, out_ch):
        super(down, self).__init__()
        self.mpconv = nn.Sequential(
            nn.MaxPool2d(2),
            double_conv(in_ch, out_ch)
        )

    def run5(self, pooled5, **kw):
        return self.mpconv[1].run5(pooled5, **kw)

    def forward(self, x):
        # stand-alone use (reference :44-46): the pooling glue is torch's (differentiable), the double_conv is ours
        return self.mpconv[1](torch.nn.functional.max_pool2d(x, 2))


class up(_CacheInvalidation, nn.Module):
    '''bilinear x2 => pad => cat([skip, up]) => double_conv   -- reference models/unet_parts.py:49-80'''

    def __init__(self, in_ch, out_ch, bilinear=True):
        super(up, self).__init__()
        if not bilinear:
            raise NotImplementedError("the reference only ever builds up(..., bilinear=True) (models/bidate_model.py:16-19)")
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self.conv = double_conv(in_ch, out_ch)

    def run5(self, low5, skip5, cat5=None, **kw):
        """low5: [2,B,h,w,C] (both dates; their product is taken on the fly) or [1,B,h,w,C];
        skip5: [2,B,H,W,Cs] encoder activations of both dates (relu(d2*d1) is fused), or None when ``cat5`` already
        holds the fused skip half (written by the encoder conv's product epilogue)."""
        if skip5 is not None:
            cat5 = ops.build_up_input(skip5, low5)
        else:
            ops.build_up_input(None, low5, out=cat5)
        return self.conv.run5(cat5, **kw)

    def forward(self, x1, x2):
        """reference calling convention (:64-80): x1 = low-res tensor, x2 = skip, both NCHW fp32.  Stand-alone use: the
        upsample / pad / concat glue is torch's (differentiable); the double_conv runs on the tcgen05 kernels."""
        x1 = self.up(x1)
        dy, dx = x2.size(2) - x1.size(2), x2.size(3) - x1.size(3)
        x1 = torch.nn.functional.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        return self.conv(torch.cat([x2, x1], dim=1))


class outconv(_CacheInvalidation, nn.Module):
    '''1x1 conv head   -- reference models/unet_parts.py:83-90'''

    def __init__(self, in_ch, out_ch):
        super(outconv, self).__init__()
        if out_ch != 2:
            raise NotImplementedError("fabric_b200 outconv is built for n_classes == 2 (models/bidate_model.py via helpers.py:334)")
        self.conv = nn.Conv2d(in_ch, out_ch, 1)

    def head(self):
        """(weight [2,C] fp32, bias [2]) for fusion into the preceding conv's epilogue."""
        w = self.conv.weight.detach()
        return w.reshape(w.shape[0], w.shape[1]).contiguous(), self.conv.bias.detach()

    def run5(self, x5):
        return ops.outconv(x5, self.conv.weight, self.conv.bias)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("fabric_b200 blocks run on sm_100 CUDA devices only (no CPU fallback)")
        if self.training and torch.is_grad_enabled():
            from .autograd import outconv_train_forward
            return outconv_train_forward(self, x)
        with torch.cuda.device(x.device):
            x5 = ops.pack_input(x, c_pad=x.shape[1]).unsqueeze(0)
            return self.run5(x5)
