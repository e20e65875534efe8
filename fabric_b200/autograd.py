"""Training-mode forward and backward of BiDateNet as ONE ``torch.autograd.Function`` over the C-ABI kernels
(reference: autograd through models/bidate_model.py:22-40 at train.py:91-94).

Forward (per double_conv): conv (tcgen05, raw output z + BatchNorm moment partials from the epilogue) ->
bn_finalize (batch statistics per date group, running-stat update) -> bn_apply (+ReLU, +MaxPool copy).
Backward: outconv_bwd -> per block [bn_relu_bwd (with the product-fusion / max-pool adjoints fused into its reads) ->
wgrad (tcgen05) + dgrad (the forward conv kernel with tap-flipped weights)] -> up_input_bwd between decoder stages.
Everything stays NHWC bf16 on the device; gradients of the fp32 parameters come back in nn.Conv2d / BatchNorm layout.
"""
from __future__ import annotations

import torch

from . import ops

KEEP_SAVED = False   # tests: keep the stored forward tensors of the last training forward in LAST_SAVED
LAST_SAVED = None


def _dc_forward(dc, x5, pool, prod_out=None):
    """double_conv in training mode.  Returns (a2, pooled, saved).  ``prod_out``: decoder input whose skip half
    receives relu(a2[date 1] * a2[date 0]) from the BN-apply kernel."""
    c1, b1, c2, b2 = dc.conv[0], dc.conv[1], dc.conv[3], dc.conv[4]
    g, b, h, w, _ = x5.shape
    n = b * h * w
    r1 = ops.conv3x3(x5, dc._packed(0), dc.out_ch, stats=True, tune=dc.tune1, true_cin=dc.in_ch)
    s1 = ops.bn_finalize(r1["stats"], b1, c1.bias, n, g)
    a1, _ = ops.bn_apply_relu(r1["y"], s1[0], s1[1])
    r2 = ops.conv3x3(a1, dc._packed(3), dc.out_ch, stats=True, tune=dc.tune2)
    s2 = ops.bn_finalize(r2["stats"], b2, c2.bias, n, g)
    a2, pooled = ops.bn_apply_relu(r2["y"], s2[0], s2[1], pool=pool, prod_out=prod_out)
    saved = dict(x=x5, z1=r1["y"], a1=a1, z2=r2["y"], a2=a2, s1=s1, s2=s2)
    return a2, pooled, saved


def _dc_backward(dc, sv, ga, mul_other, gp, need_dx, grads):
    """Backward of one double_conv.  ga / gp: gradient sources for its output activation (see ops.bn_relu_bwd)."""
    c1, b1, c2, b2 = dc.conv[0], dc.conv[1], dc.conv[3], dc.conv[4]
    need_a = mul_other or gp is not None
    dz2, dg2, db2 = ops.bn_relu_bwd(sv["z2"], sv["a2"] if need_a else None, ga, mul_other, gp, *sv["s2"], b2.weight)
    grads[c2.weight] = ops.conv3x3_wgrad(dz2, sv["a1"], dc.out_ch)
    grads[c2.bias] = torch.zeros_like(c2.bias)          # a conv bias in front of a train-mode BN has zero gradient
    grads[b2.weight], grads[b2.bias] = dg2, db2
    w2d = dc._cache().get(("wd", 3), [c2.weight], lambda: ops.pack_conv_weight(c2.weight, 1))
    da1 = ops.conv3x3(dz2, w2d, dc.out_ch)["y"]
    del dz2
    dz1, dg1, db1 = ops.bn_relu_bwd(sv["z1"], None, da1, False, None, *sv["s1"], b1.weight)
    del da1
    grads[c1.weight] = ops.conv3x3_wgrad(dz1, sv["x"], dc.in_ch)
    grads[c1.bias] = torch.zeros_like(c1.bias)
    grads[b1.weight], grads[b1.bias] = dg1, db1
    if not need_dx:
        return None
    w1d = dc._cache().get(("wd", 0), [c1.weight], lambda: ops.pack_conv_weight(c1.weight, 1))
    return ops.conv3x3(dz1, w1d, sv["x"].shape[4])["y"]


class _BiDateNetTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x_d1, x_d2, aug, *params):
        x5 = model.pack_pair(x_d1, x_d2, aug)
        _, b, h, w, _ = x5.shape
        dev = x5.device

        def cat(level, cs, cl):     # decoder input; the skip half (relu(d2*d1)) is written by the encoder's BN-apply kernel
            return torch.empty((1, b, h >> level, w >> level, cs + cl), dtype=torch.bfloat16, device=dev)
        cat4, cat3, cat2, cat1 = cat(0, 64, 64), cat(1, 128, 128), cat(2, 256, 256), cat(3, 512, 512)
        sv = {}
        e1, p1, sv["inc"] = _dc_forward(model.inc.conv, x5, True, cat4)                 # bidate_model.py:23,29 (+:38 skip)
        e2, p2, sv["down1"] = _dc_forward(model.down1.mpconv[1], p1, True, cat3)        # :24,30 (+:37)
        e3, p3, sv["down2"] = _dc_forward(model.down2.mpconv[1], p2, True, cat2)        # :25,31 (+:36)
        e4, p4, sv["down3"] = _dc_forward(model.down3.mpconv[1], p3, True, cat1)        # :26,32 (+:35)
        e5, _, sv["down4"] = _dc_forward(model.down4.mpconv[1], p4, False)              # :27,33
        ops.build_up_input(None, e5, out=cat1)                                          # :35 upsampled half
        u1, _, sv["up1"] = _dc_forward(model.up1.conv, cat1, False)
        ops.build_up_input(None, u1, out=cat2)                                          # :36
        u2, _, sv["up2"] = _dc_forward(model.up2.conv, cat2, False)
        ops.build_up_input(None, u2, out=cat3)                                          # :37
        u3, _, sv["up3"] = _dc_forward(model.up3.conv, cat3, False)
        ops.build_up_input(None, u3, out=cat4)                                          # :38
        u4, _, sv["up4"] = _dc_forward(model.up4.conv, cat4, False)
        logits = ops.outconv(u4, model.outc.conv.weight, model.outc.conv.bias)   # :39
        ctx.model, ctx.sv, ctx.params = model, sv, params
        if KEEP_SAVED:
            global LAST_SAVED
            LAST_SAVED = sv
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model, sv, params = ctx.model, ctx.sv, ctx.params
        grads = {}
        dlogits = dlogits.contiguous().float()
        oc = model.outc.conv
        du4, grads[oc.weight], grads[oc.bias] = ops.outconv_bwd(dlogits, sv["up4"]["a2"], oc.weight)
        # decoder, top down: d(cat) = [d(skip product) | d(upsampled low)]
        dcat4 = _dc_backward(model.up4.conv, sv["up4"], du4, False, None, True, grads)
        e1 = sv["inc"]["a2"]
        du3 = ops.up_input_bwd(dcat4, e1.shape[4], e1.shape[2] // 2, e1.shape[3] // 2)
        dcat3 = _dc_backward(model.up3.conv, sv["up3"], du3, False, None, True, grads)
        e2 = sv["down1"]["a2"]
        du2 = ops.up_input_bwd(dcat3, e2.shape[4], e2.shape[2] // 2, e2.shape[3] // 2)
        dcat2 = _dc_backward(model.up2.conv, sv["up2"], du2, False, None, True, grads)
        e3 = sv["down2"]["a2"]
        du1 = ops.up_input_bwd(dcat2, e3.shape[4], e3.shape[2] // 2, e3.shape[3] // 2)
        dcat1 = _dc_backward(model.up1.conv, sv["up1"], du1, False, None, True, grads)
        e4 = sv["down3"]["a2"]
        dp5 = ops.up_input_bwd(dcat1, e4.shape[4], e4.shape[2] // 2, e4.shape[3] // 2)   # d relu(x5_d2 * x5_d1)
        # encoder, bottom up: each level's output gets the product-fusion gradient (times the other date's
        # activation) plus the gradient flowing back through the max pool from the level below
        gp4 = _dc_backward(model.down4.mpconv[1], sv["down4"], dp5, True, None, True, grads)
        gp3 = _dc_backward(model.down3.mpconv[1], sv["down3"], dcat1, True, gp4, True, grads)
        gp2 = _dc_backward(model.down2.mpconv[1], sv["down2"], dcat2, True, gp3, True, grads)
        gp1 = _dc_backward(model.down1.mpconv[1], sv["down1"], dcat3, True, gp2, True, grads)
        _dc_backward(model.inc.conv, sv["inc"], dcat4, True, gp1, False, grads)
        ctx.sv = None
        return (None, None, None, None) + tuple(grads.get(p) for p in params)


def bidatenet_train_forward(model, x_d1, x_d2, aug=None):
    params = tuple(model.parameters())
    return _BiDateNetTrain.apply(model, x_d1.contiguous(), x_d2.contiguous(), aug, *params)
