"""Tensor-level wrappers over the C ABI.  torch is used for device memory and streams only.

Activation tensors ("NHWC5") are contiguous bf16 ``[G, B, H, W, C]``; G is the date group (2 while both dates
run through the weight-shared encoder as one launch, 1 in the decoder).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import Conv3x3Desc, ConvTuning, check


# bookkeeping for bench.py: how many of OUR kernels were launched, and (optionally) CUDA-event brackets around
# every conv launch on the launching stream so the roofline figure is measured live inside the timed region
LAUNCHES = 0
CONV_PROFILE = None   # set to a list to collect (tag, start_event, end_event, algorithmic_flops)


# FABRIC_B200_POISON=1: every output / workspace tensor allocated here starts as NaN (floats) or 0xFF bytes (integers)
# instead of whatever the caching allocator hands back, so a kernel that READS anything it (or its producer) did not write
# turns a test red instead of passing by luck.  The GPU test-suite is run once in this mode (tools/gpu_round2.sh poison):
# it is the initcheck that also covers tensors written by TMA stores, which compute-sanitizer's initcheck cannot see.
POISON = os.environ.get("FABRIC_B200_POISON", "0") == "1"


class ExactGlobal:
    """Exact-global data-parallel arithmetic: while ``ops.EXACT`` holds one of these, BatchNorm batch
    statistics (forward moments and the backward (sum dy, sum dy*xhat) pairs) and the loss sums are all-reduced across the
    process group between the kernels' reduce and finalize phases, so that N ranks x B pairs compute what one rank would
    compute on N*B pairs (SyncBN == the reference's single-device statistics; loss on the gathered batch, train.py:91-92).
    Latency-bound collectives on the critical path: off by default (``DataParallelStep(exact=True)`` turns it on)."""

    def __init__(self, process_group=None):
        import torch.distributed as dist
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.collectives = 0

    def all_reduce(self, t: torch.Tensor):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            self.collectives += 1
        return t


EXACT: Optional[ExactGlobal] = None


def _poison(t: torch.Tensor) -> torch.Tensor:
    if POISON and t.numel():
        if t.dtype.is_floating_point:
            t.fill_(float("nan"))
        else:
            t.view(torch.uint8).fill_(0xFF)
    return t


def _empty(*a, **kw) -> torch.Tensor:
    return _poison(torch.empty(*a, **kw))


def _empty_like(t: torch.Tensor) -> torch.Tensor:
    return _poison(torch.empty_like(t))


def _count(n: int = 1):
    global LAUNCHES
    LAUNCHES += n


# current device / raw stream handle through torch's C bindings: `torch.cuda.current_stream().cuda_stream` builds a Stream object
# and re-validates the device on every call (~4 us x ~175 launches = 0.7 ms of a 4.6 ms host-bound step at the reference's default
# 32 x 90 x 90 geometry, tools/host_overhead.py)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _device_index() -> int:
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _stream() -> int:
    if _raw_stream is not None:
        return _raw_stream(_device_index())
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.FabricB200Error("fabric_b200 ops need CUDA tensors on an sm_100 device; there is no CPU path")
        if not t.is_contiguous():
            raise _lib.FabricB200Error("fabric_b200 ops need contiguous tensors")
        # kernels, TMA descriptors and the stream all belong to the CURRENT device: a tensor living elsewhere would be
        # reached through peer access (or fault).  Fail loudly instead (use torch.cuda.set_device / torch.cuda.device).
        if cur is None:
            cur = _device_index()
        if t.device.index != cur:
            raise _lib.FabricB200Error(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}; wrap the call in "
                                       "`with torch.cuda.device(t.device):` (fabric_b200 launches on the current device's stream)")


def sm_count() -> int:
    return check(_lib.load().fabric_b200_sm_count(), "sm_count")


def cpad(c: int) -> int:
    """Channel padding of a conv input: 16 for the 13-band input, else the next multiple of 64."""
    return 16 if c <= 16 else (c + 63) // 64 * 64


def pack_input(x: torch.Tensor, out: Optional[torch.Tensor] = None, c_pad: Optional[int] = None) -> torch.Tensor:
    """NCHW fp32 [B,C,H,W] -> NHWC bf16 [B,H,W,Cpad] (zero padded channels)."""
    _need_cuda(x, out)
    if x.dtype != torch.float32:
        x = x.float()
    b, c, h, w = x.shape
    c_pad = c_pad or cpad(c)
    if out is None:
        out = _empty((b, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    assert out.shape == (b, h, w, c_pad) and out.dtype == torch.bfloat16
    check(_lib.load().fabric_b200_pack_nchw_f32_to_nhwc_bf16(_p(x), _p(out), b, c, c_pad, h, w, _stream()), "pack_input")
    _count()
    return out


def pack_input_raw(x: torch.Tensor, mean: torch.Tensor, inv_std: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Raw NCHW uint16 [B,C,H,W] (C <= 16) -> z-scored NHWC bf16 [B,H,W,16]; the normalisation of the reference's loader
    (utils/dataloaders.py:94-99) runs inside the pack kernel."""
    _need_cuda(x, mean, inv_std, out)
    if x.dtype != torch.uint16:
        raise _lib.FabricB200Error("pack_input_raw expects uint16 rasters")
    b, c, h, w = x.shape
    assert mean.dtype == torch.float32 and inv_std.dtype == torch.float32 and mean.numel() == c and inv_std.numel() == c
    if out is None:
        out = _empty((b, h, w, 16), dtype=torch.bfloat16, device=x.device)
    assert out.shape == (b, h, w, 16) and out.dtype == torch.bfloat16
    check(_lib.load().fabric_b200_pack_nchw_u16_to_nhwc_bf16(_p(x), _p(out), _p(mean), _p(inv_std), b, c, h, w, _stream()),
          "pack_input_raw")
    _count()
    return out


def pack_input_aug(x: torch.Tensor, aug: torch.Tensor, mean=None, inv_std=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NCHW [B,C,S,S] fp32 or uint16 -> NHWC bf16 [B,S,S,16] of the AUGMENTED patch; ``aug`` int32 [B,3] = (rot90 quarter
    turns, flip rows, flip columns) per sample (reference utils/dataloaders.py:152-163)."""
    _need_cuda(x, aug, mean, inv_std, out)
    b, c, s, s2 = x.shape
    assert s == s2 and aug.dtype == torch.int32 and aug.shape == (b, 3)
    dt = {torch.float32: 0, torch.uint16: 1}[x.dtype]
    if out is None:
        out = _empty((b, s, s, 16), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().fabric_b200_pack_nchw_aug(_p(x), dt, _p(out), _p(aug), _p(mean), _p(inv_std), b, c, s, _stream()),
          "pack_input_aug")
    _count()
    return out


def augment_labels(labels: torch.Tensor, aug: torch.Tensor) -> torch.Tensor:
    """labels int64 [B,S,S] -> the same rot90 / flips as ``pack_input_aug`` (out of place)."""
    _need_cuda(labels, aug)
    b, s, s2 = labels.shape
    assert s == s2 and labels.dtype == torch.int64 and aug.dtype == torch.int32 and aug.shape == (b, 3)
    out = _empty_like(labels)
    check(_lib.load().fabric_b200_augment_labels(_p(labels), _p(out), _p(aug), b, s, _stream()), "augment_labels")
    _count()
    return out


def unpack_output(x: torch.Tensor) -> torch.Tensor:
    """NHWC bf16 [B,H,W,C] -> NCHW fp32 [B,C,H,W]."""
    _need_cuda(x)
    b, h, w, c = x.shape
    out = _empty((b, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.load().fabric_b200_unpack_nhwc_bf16_to_nchw_f32(_p(x), _p(out), b, c, h, w, _stream()), "unpack_output")
    _count()
    return out


def pack_conv_weight(w: torch.Tensor, mode: int = 0, scale: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[Cout,Cin,3,3] fp32 -> bf16 [Cout,9,CinPad] (mode 0, forward) or [Cin,9,Cout] with flipped taps (mode 1, dgrad).
    ``scale`` [Cout] fp32 (mode 0 only) multiplies each output channel's filter before rounding: the eval-mode BatchNorm
    scale folded into the weights (see ``conv3x3(..., shift_in_acc=True)``)."""
    _need_cuda(w, scale)
    w = w.detach()
    if w.dtype != torch.float32:
        w = w.float()
    cout, cin = w.shape[0], w.shape[1]
    cp = cpad(cin) if mode == 0 else cin
    shape = (cout, 9, cp) if mode == 0 else (cin, 9, cout)
    if out is None:
        out = _empty(shape, dtype=torch.bfloat16, device=w.device)
    assert out.shape == shape and out.dtype == torch.bfloat16 and out.is_contiguous()
    check(_lib.load().fabric_b200_pack_conv3x3_weight_scaled(_p(w), _p(scale), _p(out), cout, cin, cp, mode, _stream()),
          "pack_conv_weight")
    _count()
    return out


def bn_fold_eval(bn: torch.nn.BatchNorm2d, conv_bias: Optional[torch.Tensor]):
    """Eval-mode BatchNorm + conv bias as per-channel (scale, shift) for the conv epilogue."""
    c = bn.num_features
    dev = bn.weight.device
    scale = _empty(c, dtype=torch.float32, device=dev)
    shift = _empty(c, dtype=torch.float32, device=dev)
    check(_lib.load().fabric_b200_bn_fold_eval(_p(bn.weight.detach()), _p(bn.bias.detach()), _p(bn.running_mean),
                                               _p(bn.running_var), _p(None if conv_bias is None else conv_bias.detach()),
                                               float(bn.eps), _p(scale), _p(shift), c, _stream()), "bn_fold_eval")
    _count()
    return scale, shift


def make_tuning(n_tile=0, halo=-1, a_stages=0, b_stages=0, b_resident=-1, grid=0, ctas=0, epi_warps=0, occupancy=0) -> ConvTuning:
    return ConvTuning(n_tile, halo, a_stages, b_stages, b_resident, grid, ctas, epi_warps, occupancy)


DEFAULT_TUNING = dict(n_tile=0, halo=-1, a_stages=0, b_stages=0, b_resident=-1, grid=0, ctas=0, epi_warps=0, occupancy=0)


def conv3x3(x5: torch.Tensor, w_packed: torch.Tensor, cout: int, scale: Optional[torch.Tensor] = None,
            shift: Optional[torch.Tensor] = None, relu: bool = False, pool: bool = False, stats: bool = False,
            head=None, store_main: bool = True, tune: Optional[dict] = None, out: Optional[torch.Tensor] = None,
            true_cin: Optional[int] = None, prod_out: Optional[torch.Tensor] = None, shift_in_acc: bool = False,
            tag: Optional[str] = None, bnbwd=None):
    """3x3 pad-1 convolution on tcgen05 (see include/fabric_b200.h: fabric_b200_conv3x3).

    Returns a dict with ``y`` [G,B,H,W,cout] bf16 and optionally ``pool`` [G,B,H/2,W/2,cout],
    ``stats`` (fp32 partial moments [grid, 2, n_tile, 2] plus ``n_tile``) and ``logits`` [G*B,2,H,W] fp32.
    """
    lib = _lib.load()
    _need_cuda(x5, w_packed, scale, shift, out)
    assert x5.dim() == 5 and x5.dtype == torch.bfloat16
    g, b, h, w, cin = x5.shape
    d = Conv3x3Desc()
    d.G, d.B, d.H, d.W, d.Cin, d.Cout = g, b, h, w, cin, cout
    d.relu, d.store_main, d.shift_in_acc = int(relu), int(store_main), int(shift_in_acc)
    t = dict(DEFAULT_TUNING)
    if tune:
        t.update(tune)
    d.tune = make_tuning(**t)
    res = {}
    y = None
    if store_main:
        y = out if out is not None else _empty((g, b, h, w, cout), dtype=torch.bfloat16, device=x5.device)
        assert y.shape == (g, b, h, w, cout)
    res["y"] = y
    d.x, d.w, d.y = _p(x5), _p(w_packed), _p(y)
    d.scale, d.shift = _p(scale), _p(shift)
    if pool:
        res["pool"] = _empty((g, b, h // 2, w // 2, cout), dtype=torch.bfloat16, device=x5.device)
        d.pool_out = _p(res["pool"])
    if head is not None:
        hw, hb = head
        res["logits"] = _empty((g * b, 2, h, w), dtype=torch.float32, device=x5.device)
        d.head_w, d.head_b, d.head_out = _p(hw), _p(hb), _p(res["logits"])
    if prod_out is not None:
        # fused relu(y[date 1] * y[date 0]) into channels [0, cout) of the decoder input [1,B,H,W,Ct]
        assert g == 2 and prod_out.shape[:4] == (1, b, h, w) and prod_out.dtype == torch.bfloat16
        d.prod_out, d.prod_channels = _p(prod_out), prod_out.shape[4]
    if bnbwd is not None:
        # data-gradient launch with the BatchNorm-backward reduce fused into its epilogue: bnbwd = (z5, BnCoef) of the
        # BatchNorm + ReLU this conv's output is the gradient of; y becomes dy (masked), res["stats"] the (sum dy, sum dy*xhat)
        # partials for bn_bwd_from_partials
        zb, coef = bnbwd
        _need_cuda(zb, coef.base)
        assert zb.shape == (g, b, h, w, cout) and zb.dtype == torch.bfloat16 and coef.base.shape == (4, g, cout)
        d.bnbwd_z, d.bnbwd_coef = _p(zb), _p(coef.base)
        stats = True
    if stats:
        # the workspace size depends on the grid the planner picks; the planner ignores the pointer value
        d.stats_ws = 1
        n = check(lib.fabric_b200_conv3x3_stats_ws_floats(C.byref(d)), "conv3x3 plan")
        ws = _empty(n, dtype=torch.float32, device=x5.device)
        d.stats_ws = _p(ws)
        grid = check(lib.fabric_b200_conv3x3_grid(C.byref(d)), "conv3x3 plan")
        res["stats"] = ws.view(grid, 2, -1, 2)
    prof = CONV_PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.fabric_b200_conv3x3(C.byref(d), _stream()), "conv3x3")
    _count()
    if prof is not None:
        e1.record()
        # algorithmic flops: true input channels (13, not the padded 16) -- SURVEY.md 8d
        cin_true = true_cin if true_cin is not None else cin
        prof.append(((tag + " " if tag else "") + f"{cin_true}->{cout}@{h}x{w}xG{g}", e0, e1,
                     2.0 * g * b * h * w * 9 * cin_true * cout))
    return res


def build_up_input(skip5: Optional[torch.Tensor], low5: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cat([relu(skip_d2*skip_d1), pad(bilinear_x2(low))], C) as one kernel.  skip5 [2,B,H,W,Cs];
    low5 [2,B,h,w,Cl] (product of both dates, up1) or [1,B,h,w,Cl].  Returns [1,B,H,W,Cs+Cl].
    With ``skip5=None`` and ``out`` given, only the upsampled channels [Cs, Cs+Cl) of ``out`` are written (the skip half
    was produced by the encoder conv's fused product epilogue)."""
    _need_cuda(skip5, low5, out)
    lg, b, h, w, cl = low5.shape
    if skip5 is not None:
        assert skip5.shape[0] == 2 and skip5.shape[1] == b
        _, _, h_, w_, cs = skip5.shape
        out = _empty((1, b, h_, w_, cs + cl), dtype=torch.bfloat16, device=low5.device)
    else:
        assert out is not None and out.shape[1] == b
        _, _, h_, w_, ct = out.shape
        cs = ct - cl
    check(_lib.load().fabric_b200_build_up_input(_p(skip5), _p(low5), _p(out), b, h_, w_, cs, h, w, cl, lg, _stream()),
          "build_up_input")
    _count()
    return out


def outconv(x5: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """1x1 head: [1,B,H,W,C] bf16 -> NCHW fp32 logits [B,2,H,W]."""
    _need_cuda(x5, weight, bias)
    g, b, h, w, c = x5.shape
    assert weight.shape[0] == 2, "the fused head is built for n_classes == 2"
    out = _empty((g * b, 2, h, w), dtype=torch.float32, device=x5.device)
    check(_lib.load().fabric_b200_outconv(_p(x5), _p(weight.detach().reshape(2, c).contiguous()), _p(bias.detach()), _p(out),
                                          g * b, h, w, c, _stream()), "outconv")
    _count()
    return out


# ------------------------------------------------------------------------------------------------ training ops
# 1: one N=3*64 MMA per K step (overlapping N atoms one pixel apart; validated bit-level against the 3-MMA form 0);
# 2: filter row through an 18-row Q halo tile; 3: per-shape choice between 1 and 2 (see fabric_b200/csrc/wgrad.cu)
WGRAD_WIDE = 3


class BnCoef(tuple):
    """(scale, shift, mean, invstd), each [G,C] fp32, as views of ONE [4,G,C] buffer (``.base``): the layout the fused
    BatchNorm-backward epilogue of the data-gradient conv reads (fb_conv3x3_desc.bnbwd_coef)."""
    base: torch.Tensor


def bn_finalize(stats: torch.Tensor, bn: torch.nn.BatchNorm2d, conv_bias, count_per_group: int, groups: int):
    """Finish train-mode BatchNorm from the conv epilogue's moment partials; updates the running statistics in
    place (momentum, unbiased variance, num_batches_tracked += groups).  Returns (scale, shift, mean, invstd) [G,C]."""
    grid, _, n_tile, _ = stats.shape
    c = bn.num_features
    dev = stats.device
    base = _empty((4, groups, c), dtype=torch.float32, device=dev)
    out = BnCoef(base[i] for i in range(4))
    out.base = base
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    if EXACT is not None and EXACT.world > 1:
        # every rank launched the same grid: the element-wise SUM of the per-CTA partial arrays is a valid partial array of
        # the global batch (finalize sums over CTAs), with world x the elements per group
        EXACT.all_reduce(stats)
        count_per_group = int(count_per_group) * EXACT.world
    check(_lib.load().fabric_b200_bn_finalize(_p(stats), grid, n_tile, c, groups, int(count_per_group),
                                              _p(None if conv_bias is None else conv_bias.detach()), _p(bn.weight.detach()),
                                              _p(bn.bias.detach()), _p(bn.running_mean), _p(bn.running_var),
                                              _p(bn.num_batches_tracked), mom, float(bn.eps), _p(out[0]), _p(out[1]),
                                              _p(out[2]), _p(out[3]), _stream()), "bn_finalize")
    _count()
    # the kernel wrote the running statistics behind torch's back: invalidate caches keyed on tensor versions
    bn.__dict__["_fb_stats_epoch"] = bn.__dict__.get("_fb_stats_epoch", 0) + 1
    return out


def bn_apply_relu(z5: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, pool: bool = False,
                  prod_out: Optional[torch.Tensor] = None, write_a: bool = True):
    """a = relu(z*scale[g]+shift[g]) (+ MaxPool2d(2) copy) (+ relu(a[1]*a[0]) into prod_out[..., :C], the skip half of the
    decoder input [1,B,H,W,Ct]).  ``write_a=False`` (only with pool and prod_out): the full-resolution activation is not
    stored at all -- its only consumers are the pooled copy and the product, and the backward pass recomputes it from z."""
    g, b, h, w, c = z5.shape
    a = _empty_like(z5) if write_a else None
    pl = _empty((g, b, h // 2, w // 2, c), dtype=torch.bfloat16, device=z5.device) if pool else None
    pc = 0
    if prod_out is not None:
        assert g == 2 and prod_out.shape[:4] == (1, b, h, w)
        pc = prod_out.shape[4]
    check(_lib.load().fabric_b200_bn_apply_relu(_p(z5), _p(scale), _p(shift), _p(a), _p(pl), _p(prod_out), pc, g, b, h, w, c,
                                                _stream()), "bn_apply_relu")
    _count()
    return a, pl


def bn_apply_relu_head(z5: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, head_w: torch.Tensor, head_b: torch.Tensor,
                       write_a: bool = True):
    """up4's last BatchNorm + ReLU and `outconv` in one pass: returns (a [1,B,H,W,64] bf16, logits [B,2,H,W] fp32).
    ``write_a=False``: the activation is not stored (returns None for it) -- the training step's backward recomputes it from
    z (``bn_head_bwd``), so the 0.54 GB tensor need not touch HBM."""
    _need_cuda(z5, scale, shift, head_w, head_b)
    g, b, h, w, c = z5.shape
    assert g == 1 and head_w.shape[0] == 2
    a = _empty_like(z5) if write_a else None
    logits = _empty((b, 2, h, w), dtype=torch.float32, device=z5.device)
    check(_lib.load().fabric_b200_bn_apply_relu_head(_p(z5), _p(scale), _p(shift), _p(a),
                                                     _p(head_w.detach().reshape(2, c).contiguous()), _p(head_b.detach()),
                                                     _p(logits), b, h, w, c, _stream()), "bn_apply_relu_head")
    _count()
    return a, logits


def bn_head_bwd(dlogits: torch.Tensor, z5: torch.Tensor, coef, gamma, head_w, dgamma_out=None, dbeta_out=None, dw_out=None,
                db_out=None):
    """backward of bn_apply_relu_head: (dz [1,B,H,W,64] bf16, dgamma, dbeta, d head weight [2,64,1,1], d head bias [2])"""
    lib = _lib.load()
    _need_cuda(dlogits, z5, dgamma_out, dbeta_out, dw_out, db_out)
    g, b, h, w, c = z5.shape
    dev = z5.device
    ws = _empty(check(lib.fabric_b200_bn_head_bwd_ws_floats(), "bn_head_bwd ws"), dtype=torch.float32, device=dev)
    dz = _empty_like(z5)
    dgamma = dgamma_out if dgamma_out is not None else _empty(c, dtype=torch.float32, device=dev)
    dbeta = dbeta_out if dbeta_out is not None else _empty(c, dtype=torch.float32, device=dev)
    dw = dw_out if dw_out is not None else _empty((2, c, 1, 1), dtype=torch.float32, device=dev)
    db = db_out if db_out is not None else _empty((2,), dtype=torch.float32, device=dev)

    def call(phase, cs, gs):
        check(lib.fabric_b200_bn_head_bwd(phase, _p(dlogits), _p(z5), _p(coef[0]), _p(coef[1]), _p(coef[2]), _p(coef[3]),
                                          _p(gamma.detach()), _p(head_w.detach().reshape(2, c).contiguous()), _p(dz), _p(dgamma),
                                          _p(dbeta), _p(dw), _p(db), _p(ws), b, h, w, c, cs, gs, _stream()), "bn_head_bwd")
    if EXACT is not None and EXACT.world > 1:
        call(1, 1.0, 1.0)
        EXACT.all_reduce(ws[:sm_count() * 2 * 258])
        call(2, float(EXACT.world), 1.0 / EXACT.world)
    else:
        call(3, 1.0, 1.0)
    _count(3)
    return dz, dgamma, dbeta, dw, db


LOSS_KINDS = {"tversky": 0, "dice": 1, "jaccard": 2, "focal": 3, "ce": 4, "bce": 4}


def seg_loss_fwd_bwd(kind: str, logits: torch.Tensor, labels: torch.Tensor, alpha=0.5, beta=0.5, gamma=0.0, eps=1e-7):
    """Fused loss value + dL/dlogits (utils/metrics.py).  logits [B,2,H,W] fp32, labels [B,H,W] or [B,1,H,W] int64."""
    _need_cuda(logits, labels)
    assert logits.dtype == torch.float32 and logits.shape[1] == 2, "the fused losses are built for n_classes == 2"
    if labels.dtype != torch.int64:
        labels = labels.long()
    b, _, h, w = logits.shape
    nd = labels.dim()
    if kind in ("focal", "ce", "bce"):
        nd = 3 if nd not in (3, 4) else nd
    assert labels.numel() == b * h * w
    lib = _lib.load()
    ws = _empty(check(lib.fabric_b200_seg_loss_ws_floats(b, h, w), "seg_loss ws"), dtype=torch.float32, device=logits.device)
    loss = _empty((), dtype=torch.float32, device=logits.device)
    dlogits = _empty_like(logits)
    if EXACT is not None and EXACT.world > 1:
        # loss on the global batch: sums all-reduced between the two phases (ratio losses), mean scaled by 1/world (focal / CE)
        args = (LOSS_KINDS[kind], float(alpha), float(beta), float(gamma), float(eps), _p(logits), _p(labels), nd, b, h, w,
                _p(loss), _p(dlogits), _p(ws), 1.0 / EXACT.world, _stream())
        check(lib.fabric_b200_seg_loss_phase(1, *args), "seg_loss sums")
        if LOSS_KINDS[kind] <= 2:
            off = check(lib.fabric_b200_seg_loss_sums_offset(b, h, w), "seg_loss ws")
            EXACT.all_reduce(ws[off:off + 6 * w])
        check(lib.fabric_b200_seg_loss_phase(2, *args), "seg_loss")
        if LOSS_KINDS[kind] > 2:
            EXACT.all_reduce(loss)
        _count(3)
        return loss, dlogits
    check(lib.fabric_b200_seg_loss_fwd_bwd(LOSS_KINDS[kind], float(alpha), float(beta), float(gamma), float(eps), _p(logits),
                                           _p(labels), nd, b, h, w, _p(loss), _p(dlogits), _p(ws), _stream()), "seg_loss")
    _count(3)
    return loss, dlogits


def outconv_bwd(dlogits: torch.Tensor, u5: torch.Tensor, weight: torch.Tensor, dw_out=None, db_out=None):
    """``dw_out`` [2,C,1,1] / ``db_out`` [2] fp32: write the parameter gradients there (e.g. views of the data-parallel
    gradient bucket) instead of fresh tensors."""
    g, b, h, w, c = u5.shape
    lib = _lib.load()
    _need_cuda(dlogits, u5, dw_out, db_out)
    ws = _empty(check(lib.fabric_b200_outconv_bwd_ws_floats(c), "outconv_bwd ws"), dtype=torch.float32, device=u5.device)
    du = _empty_like(u5)
    dw = dw_out if dw_out is not None else _empty((2, c), dtype=torch.float32, device=u5.device)
    db = db_out if db_out is not None else _empty((2,), dtype=torch.float32, device=u5.device)
    check(lib.fabric_b200_outconv_bwd(_p(dlogits), _p(u5), _p(weight.detach().reshape(2, c).contiguous()), _p(du), _p(dw),
                                      _p(db), _p(ws), g * b, h, w, c, _stream()), "outconv_bwd")
    _count(2)
    return du, dw.view(2, c, 1, 1), db


def bn_relu_bwd(z5, a5, ga, mul_other, gp, scale, shift, mean, invstd, gamma, dgamma_out=None, dbeta_out=None):
    """BatchNorm(train)+ReLU backward with fused product / max-pool adjoints (see include/fabric_b200.h).
    Returns (dz [G,B,H,W,C] bf16, dgamma [C], dbeta [C])."""
    g, b, h, w, c = z5.shape
    lib = _lib.load()
    _need_cuda(z5, a5, ga, gp, dgamma_out, dbeta_out)
    ws = _empty(check(lib.fabric_b200_bn_bwd_ws_floats(g, c), "bn_bwd ws"), dtype=torch.float32, device=z5.device)
    dz = _empty_like(z5)
    dgamma = dgamma_out if dgamma_out is not None else _empty(c, dtype=torch.float32, device=z5.device)
    dbeta = dbeta_out if dbeta_out is not None else _empty(c, dtype=torch.float32, device=z5.device)
    ga_groups, ga_ch = (ga.shape[0], ga.shape[4]) if ga is not None else (1, c)
    if EXACT is not None and EXACT.world > 1:
        args = (_p(z5), _p(a5), _p(ga), ga_groups, ga_ch, int(mul_other), _p(gp), _p(scale), _p(shift), _p(mean), _p(invstd),
                _p(gamma.detach()), _p(dz), _p(dgamma), _p(dbeta), _p(ws), g, b, h, w, c, float(EXACT.world), 1.0 / EXACT.world,
                _stream())
        check(lib.fabric_b200_bn_relu_bwd_phase(1, *args), "bn_relu_bwd reduce")
        EXACT.all_reduce(ws[:check(lib.fabric_b200_bn_bwd_partial_floats(g, c), "bn_bwd ws")])
        check(lib.fabric_b200_bn_relu_bwd_phase(2, *args), "bn_relu_bwd apply")
        _count(3)
        return dz, dgamma, dbeta
    check(lib.fabric_b200_bn_relu_bwd(_p(z5), _p(a5), _p(ga), ga_groups, ga_ch, int(mul_other), _p(gp), _p(scale), _p(shift),
                                      _p(mean), _p(invstd), _p(gamma.detach()), _p(dz), _p(dgamma), _p(dbeta), _p(ws),
                                      g, b, h, w, c, _stream()), "bn_relu_bwd")
    _count(3)
    return dz, dgamma, dbeta


def bn_bwd_from_partials(z5, dy5, partial, coef: BnCoef, gamma, dgamma_out=None, dbeta_out=None):
    """BatchNorm(train)+ReLU backward when the producing data-gradient conv already masked (dy5) and reduced (partial, the
    conv's ``res["stats"]`` view [grid,2,n_tile,2]): ONE pass over (dy, z).  Returns (dz, dgamma, dbeta)."""
    g, b, h, w, c = z5.shape
    lib = _lib.load()
    _need_cuda(z5, dy5, partial, dgamma_out, dbeta_out)
    grid, _, n_tile, _ = partial.shape
    dz = _empty_like(z5)
    dgamma = dgamma_out if dgamma_out is not None else _empty(c, dtype=torch.float32, device=z5.device)
    dbeta = dbeta_out if dbeta_out is not None else _empty(c, dtype=torch.float32, device=z5.device)
    ws = _empty(g * 3 * c, dtype=torch.float32, device=z5.device)
    count_scale = grad_scale = 1.0
    if EXACT is not None and EXACT.world > 1:
        EXACT.all_reduce(partial)
        count_scale, grad_scale = float(EXACT.world), 1.0 / EXACT.world
    check(lib.fabric_b200_bn_bwd_from_partials(_p(z5), _p(dy5), _p(partial), grid, n_tile, _p(coef[2]), _p(coef[3]),
                                               _p(gamma.detach()), _p(dz), _p(dgamma), _p(dbeta), _p(ws), g, b, h, w, c,
                                               count_scale, grad_scale, _stream()), "bn_bwd_from_partials")
    _count(2)
    return dz, dgamma, dbeta


def up_input_bwd(dcat5: torch.Tensor, cs: int, h: int, w: int) -> torch.Tensor:
    _, b, hh, ww, ct = dcat5.shape
    cl = ct - cs
    dlow = _empty((1, b, h, w, cl), dtype=torch.bfloat16, device=dcat5.device)
    check(_lib.load().fabric_b200_up_input_bwd(_p(dcat5), _p(dlow), b, hh, ww, cs, h, w, cl, _stream()), "up_input_bwd")
    _count()
    return dlow


# operand swap of the weight gradient for 64-channel dL/dz against a >= 128-channel input (up3.c1, up4.c1): the kernel's M
# dimension (128 MMA rows) then carries the INPUT's channels instead of 64 dL/dz channels x two of the three filter rows
# (75 % useful).  FABRIC_B200_WGRAD_SWAP=0 is the A/B switch.
WGRAD_SWAP = os.environ.get("FABRIC_B200_WGRAD_SWAP", "1") != "0"


def conv3x3_wgrad(dz5: torch.Tensor, x5: torch.Tensor, cin_true: int, splits: int = 0, wide=None,
                  out: Optional[torch.Tensor] = None, swap=None) -> torch.Tensor:
    """dW [Cout,Cin,3,3] fp32 = autograd weight gradient of conv3x3(x5, W) given dL/dz (tcgen05, split-K).  ``out``: write
    the gradient there (e.g. a view of the data-parallel gradient bucket).  ``swap``: run the kernel with the operand roles
    exchanged (see WGRAD_SWAP; None = automatic)."""
    lib = _lib.load()
    _need_cuda(dz5, x5, out)
    g, b, h, w, ca = dz5.shape
    cb = x5.shape[4]
    if swap is None:
        swap = WGRAD_SWAP and ca == 64 and cb >= 128 and cb % 128 == 0 and cin_true == cb
    elif swap and not (cb % 64 == 0 and cin_true == cb):
        raise ValueError("operand swap needs an unpadded input of a multiple of 64 channels")
    d = _lib.WgradDesc()
    d.G, d.B, d.H, d.W = g, b, h, w
    if swap:
        d.Ca, d.Cb, d.p, d.q = cb, ca, _p(x5), _p(dz5)
    else:
        d.Ca, d.Cb, d.p, d.q = ca, cb, _p(dz5), _p(x5)
    d.splits = splits
    d.wide = WGRAD_WIDE if wide is None else int(wide)
    d.ws = 16  # planner only checks alignment/non-null later
    n = check(lib.fabric_b200_conv3x3_wgrad_ws_floats(C.byref(d)), "wgrad plan")
    s = check(lib.fabric_b200_conv3x3_wgrad_splits(C.byref(d)), "wgrad plan")
    ws = _empty(n, dtype=torch.float32, device=dz5.device)
    d.ws = _p(ws)
    prof = CONV_PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.fabric_b200_conv3x3_wgrad(C.byref(d), _stream()), "conv3x3_wgrad")
    if prof is not None:
        e1.record()
        prof.append((f"wgrad {cin_true}->{ca}@{h}x{w}xG{g}", e0, e1, 2.0 * g * b * h * w * 9 * cin_true * ca))
    dw = out if out is not None else _empty((ca, cin_true, 3, 3), dtype=torch.float32, device=dz5.device)
    assert dw.shape == (ca, cin_true, 3, 3) and dw.dtype == torch.float32
    if swap:
        check(lib.fabric_b200_wgrad_reduce_swapped(_p(ws), s, ca, cin_true, _p(dw), _stream()), "wgrad_reduce")
    else:
        check(lib.fabric_b200_wgrad_reduce(_p(ws), s, ca, cin_true, cb, _p(dw), _stream()), "wgrad_reduce")
    _count(2)
    return dw


# ------------------------------------------------------------------------------------------------ scene ops
def gather_tiles(scene: torch.Tensor, origins: torch.Tensor, p: int, mean=None, inv_std=None, c_pad: Optional[int] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scene [C,H,W] fp32 / uint16 (device) -> tiles bf16 [N,p,p,Cpad] at origins int32 [N,2] (row, col)."""
    _need_cuda(scene, origins, mean, inv_std, out)
    c, h, w = scene.shape
    n = origins.shape[0]
    c_pad = c_pad or cpad(c)
    dt = {torch.float32: 0, torch.uint16: 1}.get(scene.dtype)
    if dt is None:
        raise _lib.FabricB200Error("scene must be float32 or uint16")
    assert origins.dtype == torch.int32
    if out is None:
        out = _empty((n, p, p, c_pad), dtype=torch.bfloat16, device=scene.device)
    check(_lib.load().fabric_b200_gather_tiles(_p(scene), dt, _p(origins), _p(out), _p(mean), _p(inv_std), n, c, c_pad, h, w, p,
                                               _stream()), "gather_tiles")
    _count()
    return out


def argmax_metrics(logits: torch.Tensor, labels: Optional[torch.Tensor] = None, want_mask: bool = True,
                   counts: Optional[torch.Tensor] = None, mask_out: Optional[torch.Tensor] = None):
    """torch.max(logits,1) indices as uint8 [B,H,W] and (with labels) confusion counts (TP, FP, FN, TN) accumulated
    into `counts` (uint64 [4], device).  ``mask_out``: write the mask there (contiguous uint8 [B,H,W])."""
    _need_cuda(logits, labels, counts, mask_out)
    b, _, h, w = logits.shape
    if mask_out is not None:
        assert mask_out.shape == (b, h, w) and mask_out.dtype == torch.uint8
        mask = mask_out
    else:
        mask = _empty((b, h, w), dtype=torch.uint8, device=logits.device) if want_mask else None
    if labels is not None:
        if counts is None:
            counts = torch.zeros(4, dtype=torch.int64, device=logits.device)
        if labels.dtype != torch.int64:
            labels = labels.long()
    check(_lib.load().fabric_b200_argmax_metrics(_p(logits), _p(labels), _p(mask), _p(counts), b, h, w, _stream()),
          "argmax_metrics")
    _count()
    return mask, counts


def scatter_tiles(masks: torch.Tensor, origins: torch.Tensor, canvas: torch.Tensor, first: int, count: int):
    _need_cuda(masks, origins, canvas)
    n, p, _ = masks.shape
    h, w = canvas.shape
    check(_lib.load().fabric_b200_scatter_tiles(_p(masks), _p(origins), _p(canvas), first, count, p, h, w, _stream()),
          "scatter_tiles")
    _count()
