"""Training-mode forward and backward of BiDateNet as ONE ``torch.autograd.Function`` over the C-ABI kernels
(reference: autograd through models/bidate_model.py:22-40 at train.py:91-94).

Forward (per double_conv): conv (tcgen05, raw output z + BatchNorm moment partials from the epilogue) ->
bn_finalize (batch statistics per date group, running-stat update) -> bn_apply (+ReLU, +MaxPool copy).
Backward: outconv_bwd -> per block [bn_relu_bwd (with the product-fusion / max-pool adjoints fused into its reads) ->
wgrad (tcgen05) + dgrad (the forward conv kernel with tap-flipped weights)] -> up_input_bwd between decoder stages.
Everything stays NHWC bf16 on the device; gradients of the fp32 parameters come back in nn.Conv2d / BatchNorm layout.

Gradient sink: when a ``fabric_b200.distributed.DataParallelStep`` owns the model, every parameter gradient is written
by the producing kernel STRAIGHT into that step's flat all-reduce bucket (``p.grad`` is a view of it) and the Function
returns ``None`` for the parameters -- no gradient tensor is allocated, copied or accumulated by torch.
"""
from __future__ import annotations

import os

import torch

from . import ops

KEEP_SAVED = False   # tests: keep the stored forward tensors of the last training forward in LAST_SAVED
LAST_SAVED = None
# the data-gradient conv of c2 masks its output with the ReLU of BatchNorm 1 and reduces (sum dy, sum dy*xhat) in its
# epilogue, so BatchNorm 1's backward is ONE pass over (dy, z) instead of a reduce pass + an apply pass (False = the
# stand-alone kernels, kept as the cross-check)
FUSE_BN_BWD_REDUCE = True
FUSE_BN_BWD_MAX_CH = 64
# the product-fused encoder activations a2 (levels inc .. down3) are not stored; BatchNorm-2's backward recomputes them from z2
RECOMPUTE_ENCODER_ACT = True
# up4's last BatchNorm + ReLU and `outconv` run as one pass forward and two passes backward (dlogits -> dz directly)
FUSE_HEAD = True
# weight-gradient launches go to a side stream: wgrad(c) only needs dz and the saved conv input, while the main stream carries
# on with the data gradient -> BatchNorm backward chain; both kinds of kernel fill the GPU, so what overlaps is each kernel's
# tail (partly filled last wave) and the launch bubbles of the small finalize kernels in between
WGRAD_SIDE_STREAM = os.environ.get("FABRIC_B200_WGRAD_SIDE", "1") != "0"
_SIDE = {}
_PENDING = {}


def _wgrad(dz5, x5, cin_true, out):
    if not WGRAD_SIDE_STREAM:
        return ops.conv3x3_wgrad(dz5, x5, cin_true, out=out)
    dev = dz5.device
    main = torch.cuda.current_stream(dev)
    side = _SIDE.get(dev.index)
    if side is None:
        side = _SIDE[dev.index] = torch.cuda.Stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        dw = ops.conv3x3_wgrad(dz5, x5, cin_true, out=out)
    # The caching allocator must not hand these blocks to later main-stream work while the side stream still reads them:
    # they stay referenced until the streams join (Tensor.record_stream would do, but it defers every reuse to an event
    # poll and measured pathological -- one run in two took 1.8x longer, the allocator falling back to cudaMalloc).
    _PENDING.setdefault(dev.index, []).append((dz5, x5, dw))
    return dw


def _join_side(dev):
    side = _SIDE.get(dev.index)
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)
    _PENDING.pop(dev.index, None)      # (freed blocks go back to the main stream's pool: ordered after the join)

# backward order of the blocks (gradients of a block are complete when its _dc_backward returns)
BACKWARD_ORDER = ("outc", "up4", "up3", "up2", "up1", "down4", "down3", "down2", "down1", "inc")


def _dc_forward(dc, x5, pool, prod_out=None, head=None):
    """double_conv in training mode.  Returns (a2, pooled, saved).  ``prod_out``: decoder input whose skip half
    receives relu(a2[date 1] * a2[date 0]) from the BN-apply kernel.  ``head`` = the `outconv` module: its 1x1 conv is fused
    into the last BN-apply pass (then ``pooled`` carries the logits)."""
    c1, b1, c2, b2 = dc.conv[0], dc.conv[1], dc.conv[3], dc.conv[4]
    g, b, h, w, _ = x5.shape
    n = b * h * w
    r1 = ops.conv3x3(x5, dc._packed(0, training=True), dc.out_ch, stats=True, tune=dc.tune1, true_cin=dc.in_ch)
    s1 = ops.bn_finalize(r1["stats"], b1, c1.bias, n, g)
    a1, _ = ops.bn_apply_relu(r1["y"], s1[0], s1[1])
    r2 = ops.conv3x3(a1, dc._packed(3, training=True), dc.out_ch, stats=True, tune=dc.tune2)
    s2 = ops.bn_finalize(r2["stats"], b2, c2.bias, n, g)
    # encoder levels 1-4: the activation is consumed only through its pooled copy and the date product, and backward
    # recomputes it from z2 (bit-identically) -- it never touches HBM.  (KEEP_SAVED: tests want to look at it.)
    skip_a = RECOMPUTE_ENCODER_ACT and pool and prod_out is not None and not KEEP_SAVED
    if head is not None:
        # (the head's backward recomputes the activation from z2: it is stored only for tests that look at it)
        a2, pooled = ops.bn_apply_relu_head(r2["y"], s2[0], s2[1], head.conv.weight, head.conv.bias, write_a=KEEP_SAVED)
    else:
        a2, pooled = ops.bn_apply_relu(r2["y"], s2[0], s2[1], pool=pool, prod_out=prod_out, write_a=not skip_a)
    saved = dict(x=x5, z1=r1["y"], a1=a1, z2=r2["y"], a2=a2, s1=s1, s2=s2)
    return a2, pooled, saved


def _dc_backward(dc, sv, ga, mul_other, gp, need_dx, grads, sink=None, head=None):
    """Backward of one double_conv.  ga / gp: gradient sources for its output activation (see ops.bn_relu_bwd).
    ``sink``: {parameter: gradient tensor to write into} (the data-parallel bucket views) or None.  ``head``: the `outconv`
    module whose forward was fused behind this block -- then ``ga`` is dL/dlogits (fp32 NCHW)."""
    c1, b1, c2, b2 = dc.conv[0], dc.conv[1], dc.conv[3], dc.conv[4]
    sink = sink or {}
    need_a = mul_other or gp is not None
    if head is not None:
        oc = head.conv
        dz2, dg2, db2, grads[oc.weight], grads[oc.bias] = ops.bn_head_bwd(
            ga, sv["z2"], sv["s2"], b2.weight, oc.weight, dgamma_out=sink.get(b2.weight), dbeta_out=sink.get(b2.bias),
            dw_out=sink.get(oc.weight), db_out=sink.get(oc.bias))
    else:
        dz2, dg2, db2 = ops.bn_relu_bwd(sv["z2"], sv["a2"] if need_a else None, ga, mul_other, gp, *sv["s2"], b2.weight,
                                        dgamma_out=sink.get(b2.weight), dbeta_out=sink.get(b2.bias))
    grads[c2.weight] = _wgrad(dz2, sv["a1"], dc.out_ch, sink.get(c2.weight))
    # a conv bias in front of a train-mode BN has zero gradient (the sink's slot was zeroed once and is never written)
    grads[c2.bias] = sink[c2.bias] if c2.bias in sink else torch.zeros_like(c2.bias)
    grads[b2.weight], grads[b2.bias] = dg2, db2
    # (fused for the 64-wide layers -- the 256 x 256 and 128 x 128 tensors, where the stand-alone reduce costs most.  The
    #  128-wide data gradients keep their weight slab resident only without the prefetch buffer, and measured faster with the
    #  stand-alone reduce pass (0.38 + 0.2 ms against 0.66 ms); on the 256-wide tiles the fused form reads z with plain loads.)
    if FUSE_BN_BWD_REDUCE and dc.out_ch <= FUSE_BN_BWD_MAX_CH:
        r = ops.conv3x3(dz2, dc._packed_dgrad(3), dc.out_ch, tag="dgrad", bnbwd=(sv["z1"], sv["s1"]))
        del dz2
        dz1, dg1, db1 = ops.bn_bwd_from_partials(sv["z1"], r["y"], r["stats"], sv["s1"], b1.weight,
                                                 dgamma_out=sink.get(b1.weight), dbeta_out=sink.get(b1.bias))
        del r
    else:
        da1 = ops.conv3x3(dz2, dc._packed_dgrad(3), dc.out_ch, tag="dgrad")["y"]
        del dz2
        dz1, dg1, db1 = ops.bn_relu_bwd(sv["z1"], None, da1, False, None, *sv["s1"], b1.weight,
                                        dgamma_out=sink.get(b1.weight), dbeta_out=sink.get(b1.bias))
        del da1
    grads[c1.weight] = _wgrad(dz1, sv["x"], dc.in_ch, sink.get(c1.weight))
    grads[c1.bias] = sink[c1.bias] if c1.bias in sink else torch.zeros_like(c1.bias)
    grads[b1.weight], grads[b1.bias] = dg1, db1
    if not need_dx:
        return None
    return ops.conv3x3(dz1, dc._packed_dgrad(0), sv["x"].shape[4], tag="dgrad")["y"]


class _BiDateNetTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x_d1, x_d2, aug, *params):
        x5 = model.pack_pair(x_d1, x_d2, aug)
        _, b, h, w, _ = x5.shape
        dev = x5.device

        def cat(level, cs, cl):     # decoder input; the skip half (relu(d2*d1)) is written by the encoder's BN-apply kernel
            return ops._empty((1, b, h >> level, w >> level, cs + cl), dtype=torch.bfloat16, device=dev)
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
        ctx.fuse_head = FUSE_HEAD
        if FUSE_HEAD:
            u4, logits, sv["up4"] = _dc_forward(model.up4.conv, cat4, False, head=model.outc)   # :38-39
        else:
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
        dp = model.__dict__.get("_fb_dp")            # DataParallelStep that owns the gradients, if any
        sink = dp.sink if dp is not None else None
        dev_ = dlogits.device

        def done(name):          # every gradient of block `name` is written once the side stream's wgrads have joined
            if dp is not None and dp.wants_block(name):
                _join_side(dev_)
                dp.block_done(name)
        grads = {}
        dlogits = dlogits.contiguous().float()
        oc = model.outc.conv
        with torch.cuda.device(dlogits.device):
            s = sink or {}
            if ctx.fuse_head:
                # outconv's backward rides in BatchNorm-2's: dlogits -> dz2 directly, du4 is never materialised
                dcat4 = _dc_backward(model.up4.conv, sv["up4"], dlogits, False, None, True, grads, sink, head=model.outc)
            else:
                du4, grads[oc.weight], grads[oc.bias] = ops.outconv_bwd(dlogits, sv["up4"]["a2"], oc.weight,
                                                                        dw_out=s.get(oc.weight), db_out=s.get(oc.bias))
                dcat4 = _dc_backward(model.up4.conv, sv["up4"], du4, False, None, True, grads, sink)
            done("outc")
            done("up4")
            e1 = sv["inc"]["z2"]
            du3 = ops.up_input_bwd(dcat4, e1.shape[4], e1.shape[2] // 2, e1.shape[3] // 2)
            dcat3 = _dc_backward(model.up3.conv, sv["up3"], du3, False, None, True, grads, sink)
            done("up3")
            e2 = sv["down1"]["z2"]
            du2 = ops.up_input_bwd(dcat3, e2.shape[4], e2.shape[2] // 2, e2.shape[3] // 2)
            dcat2 = _dc_backward(model.up2.conv, sv["up2"], du2, False, None, True, grads, sink)
            done("up2")
            e3 = sv["down2"]["z2"]
            du1 = ops.up_input_bwd(dcat2, e3.shape[4], e3.shape[2] // 2, e3.shape[3] // 2)
            dcat1 = _dc_backward(model.up1.conv, sv["up1"], du1, False, None, True, grads, sink)
            done("up1")
            e4 = sv["down3"]["z2"]
            dp5 = ops.up_input_bwd(dcat1, e4.shape[4], e4.shape[2] // 2, e4.shape[3] // 2)   # d relu(x5_d2 * x5_d1)
            # encoder, bottom up: each level's output gets the product-fusion gradient (times the other date's
            # activation) plus the gradient flowing back through the max pool from the level below
            gp4 = _dc_backward(model.down4.mpconv[1], sv["down4"], dp5, True, None, True, grads, sink)
            done("down4")
            gp3 = _dc_backward(model.down3.mpconv[1], sv["down3"], dcat1, True, gp4, True, grads, sink)
            done("down3")
            gp2 = _dc_backward(model.down2.mpconv[1], sv["down2"], dcat2, True, gp3, True, grads, sink)
            done("down2")
            gp1 = _dc_backward(model.down1.mpconv[1], sv["down1"], dcat3, True, gp2, True, grads, sink)
            done("down1")
            _dc_backward(model.inc.conv, sv["inc"], dcat4, True, gp1, False, grads, sink)
            done("inc")
        _join_side(dev_)
        ctx.sv = None
        if sink is not None:
            # the kernels wrote into the bucket views that ARE p.grad: nothing for torch to accumulate
            return (None, None, None, None) + tuple(None if p in sink else grads.get(p) for p in params)
        return (None, None, None, None) + tuple(grads.get(p) for p in params)


def bidatenet_train_forward(model, x_d1, x_d2, aug=None):
    params = tuple(model.parameters())
    if not params:
        raise RuntimeError(
            "this BiDateNet has no parameters of its own: it is an nn.DataParallel replica (reference utils/helpers.py:335). "
            "fabric_b200 runs one process per GPU -- drop the nn.DataParallel wrapper and use "
            "fabric_b200.distributed.DataParallelStep under torchrun instead")
    with torch.cuda.device(x_d1.device):
        return _BiDateNetTrain.apply(model, x_d1.contiguous(), x_d2.contiguous(), aug, *params)


class _DoubleConvTrain(torch.autograd.Function):
    """One stand-alone ``double_conv`` in training mode with the reference's calling convention (NCHW fp32 in / out,
    models/unet_parts.py:21-23): the per-block entry point of ``double_conv`` / ``inconv`` / ``down`` / ``up`` ``.forward``."""

    @staticmethod
    def forward(ctx, dc, x, *params):
        with torch.cuda.device(x.device):
            x5 = ops.pack_input(x.contiguous(), c_pad=ops.cpad(dc.in_ch)).unsqueeze(0)
            a2, _, sv = _dc_forward(dc, x5, False)
            ctx.dc, ctx.sv, ctx.params, ctx.cin = dc, sv, params, x.shape[1]
            ctx.need_dx = x.requires_grad
            return ops.unpack_output(a2[0])

    @staticmethod
    def backward(ctx, dy):
        dc, sv, params = ctx.dc, ctx.sv, ctx.params
        grads = {}
        with torch.cuda.device(dy.device):
            ga = ops.pack_input(dy.contiguous().float(), c_pad=dc.out_ch).unsqueeze(0)
            need_dx = ctx.need_dx and dc.in_ch % 64 == 0       # the 13-band stem has no data gradient (its input is data)
            dx5 = _dc_backward(dc, sv, ga, False, None, need_dx, grads, None)
            _join_side(dy.device)
            dx = ops.unpack_output(dx5[0])[:, :ctx.cin].contiguous() if dx5 is not None else None
        ctx.sv = None
        return (None, dx) + tuple(grads.get(p) for p in params)


def double_conv_train_forward(dc, x):
    params = tuple(dc.parameters())
    if not params:
        raise RuntimeError("this block is an nn.DataParallel replica without parameters; see fabric_b200.distributed")
    return _DoubleConvTrain.apply(dc, x, *params)


class _OutconvTrain(torch.autograd.Function):
    """stand-alone ``outconv`` (models/unet_parts.py:88-90) with gradients"""

    @staticmethod
    def forward(ctx, oc, x, weight, bias):
        with torch.cuda.device(x.device):
            u5 = ops.pack_input(x.contiguous(), c_pad=x.shape[1]).unsqueeze(0)
            ctx.oc, ctx.u5 = oc, u5
            return ops.outconv(u5, weight, bias)

    @staticmethod
    def backward(ctx, dlogits):
        with torch.cuda.device(dlogits.device):
            du5, dw, db = ops.outconv_bwd(dlogits.contiguous().float(), ctx.u5, ctx.oc.conv.weight)
            return None, ops.unpack_output(du5[0]), dw, db


def outconv_train_forward(oc, x):
    return _OutconvTrain.apply(oc, x, oc.conv.weight, oc.conv.bias)
