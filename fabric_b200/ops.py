"""Tensor-level wrappers over the C ABI.  torch is used for device memory and streams only.

Activation tensors ("NHWC5") are contiguous bf16 ``[G, B, H, W, C]``; G is the date group (2 while both dates
run through the weight-shared encoder as one launch, 1 in the decoder).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import Conv3x3Desc, ConvTuning, check


# bookkeeping for bench.py: how many of OUR kernels were launched, and (optionally) CUDA-event brackets around
# every conv launch on the launching stream so the roofline figure is measured live inside the timed region
LAUNCHES = 0
CONV_PROFILE = None   # set to a list to collect (tag, start_event, end_event, algorithmic_flops)


def _count(n: int = 1):
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.FabricB200Error("fabric_b200 ops need CUDA tensors on an sm_100 device; there is no CPU path")
        if t is not None and not t.is_contiguous():
            raise _lib.FabricB200Error("fabric_b200 ops need contiguous tensors")


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
        out = torch.empty((b, h, w, c_pad), dtype=torch.bfloat16, device=x.device)
    assert out.shape == (b, h, w, c_pad) and out.dtype == torch.bfloat16
    check(_lib.load().fabric_b200_pack_nchw_f32_to_nhwc_bf16(_p(x), _p(out), b, c, c_pad, h, w, _stream()), "pack_input")
    _count()
    return out


def unpack_output(x: torch.Tensor) -> torch.Tensor:
    """NHWC bf16 [B,H,W,C] -> NCHW fp32 [B,C,H,W]."""
    _need_cuda(x)
    b, h, w, c = x.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.load().fabric_b200_unpack_nhwc_bf16_to_nchw_f32(_p(x), _p(out), b, c, h, w, _stream()), "unpack_output")
    _count()
    return out


def pack_conv_weight(w: torch.Tensor, mode: int = 0) -> torch.Tensor:
    """[Cout,Cin,3,3] fp32 -> bf16 [Cout,9,CinPad] (mode 0, forward) or [Cin,9,Cout] with flipped taps (mode 1, dgrad)."""
    _need_cuda(w)
    w = w.detach()
    if w.dtype != torch.float32:
        w = w.float()
    cout, cin = w.shape[0], w.shape[1]
    if mode == 0:
        cp = cpad(cin)
        out = torch.empty((cout, 9, cp), dtype=torch.bfloat16, device=w.device)
    else:
        cp = cin
        out = torch.empty((cin, 9, cout), dtype=torch.bfloat16, device=w.device)
    check(_lib.load().fabric_b200_pack_conv3x3_weight(_p(w), _p(out), cout, cin, cp, mode, _stream()), "pack_conv_weight")
    _count()
    return out


def bn_fold_eval(bn: torch.nn.BatchNorm2d, conv_bias: Optional[torch.Tensor]):
    """Eval-mode BatchNorm + conv bias as per-channel (scale, shift) for the conv epilogue."""
    c = bn.num_features
    dev = bn.weight.device
    scale = torch.empty(c, dtype=torch.float32, device=dev)
    shift = torch.empty(c, dtype=torch.float32, device=dev)
    check(_lib.load().fabric_b200_bn_fold_eval(_p(bn.weight.detach()), _p(bn.bias.detach()), _p(bn.running_mean),
                                               _p(bn.running_var), _p(None if conv_bias is None else conv_bias.detach()),
                                               float(bn.eps), _p(scale), _p(shift), c, _stream()), "bn_fold_eval")
    _count()
    return scale, shift


def make_tuning(n_tile=0, halo=-1, a_stages=0, b_stages=0, b_resident=-1, grid=0) -> ConvTuning:
    return ConvTuning(n_tile, halo, a_stages, b_stages, b_resident, grid)


DEFAULT_TUNING = dict(n_tile=0, halo=-1, a_stages=0, b_stages=0, b_resident=-1, grid=0)


def conv3x3(x5: torch.Tensor, w_packed: torch.Tensor, cout: int, scale: Optional[torch.Tensor] = None,
            shift: Optional[torch.Tensor] = None, relu: bool = False, pool: bool = False, stats: bool = False,
            head=None, store_main: bool = True, tune: Optional[dict] = None, out: Optional[torch.Tensor] = None,
            true_cin: Optional[int] = None):
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
    d.relu, d.store_main = int(relu), int(store_main)
    t = dict(DEFAULT_TUNING)
    if tune:
        t.update(tune)
    d.tune = make_tuning(**t)
    res = {}
    y = None
    if store_main:
        y = out if out is not None else torch.empty((g, b, h, w, cout), dtype=torch.bfloat16, device=x5.device)
        assert y.shape == (g, b, h, w, cout)
    res["y"] = y
    d.x, d.w, d.y = _p(x5), _p(w_packed), _p(y)
    d.scale, d.shift = _p(scale), _p(shift)
    if pool:
        res["pool"] = torch.empty((g, b, h // 2, w // 2, cout), dtype=torch.bfloat16, device=x5.device)
        d.pool_out = _p(res["pool"])
    if head is not None:
        hw, hb = head
        res["logits"] = torch.empty((g * b, 2, h, w), dtype=torch.float32, device=x5.device)
        d.head_w, d.head_b, d.head_out = _p(hw), _p(hb), _p(res["logits"])
    if stats:
        # the workspace size depends on the grid the planner picks; the planner ignores the pointer value
        d.stats_ws = 1
        n = check(lib.fabric_b200_conv3x3_stats_ws_floats(C.byref(d)), "conv3x3 plan")
        ws = torch.empty(n, dtype=torch.float32, device=x5.device)
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
        prof.append((f"{cin_true}->{cout}@{h}x{w}xG{g}", e0, e1, 2.0 * g * b * h * w * 9 * cin_true * cout))
    return res


def build_up_input(skip5: torch.Tensor, low5: torch.Tensor) -> torch.Tensor:
    """cat([relu(skip_d2*skip_d1), pad(bilinear_x2(low))], C) as one kernel.  skip5 [2,B,H,W,Cs];
    low5 [2,B,h,w,Cl] (product of both dates, up1) or [1,B,h,w,Cl].  Returns [1,B,H,W,Cs+Cl]."""
    _need_cuda(skip5, low5)
    assert skip5.shape[0] == 2
    _, b, h_, w_, cs = skip5.shape
    lg, b2, h, w, cl = low5.shape
    assert b2 == b
    out = torch.empty((1, b, h_, w_, cs + cl), dtype=torch.bfloat16, device=skip5.device)
    check(_lib.load().fabric_b200_build_up_input(_p(skip5), _p(low5), _p(out), b, h_, w_, cs, h, w, cl, lg, _stream()),
          "build_up_input")
    _count()
    return out


def outconv(x5: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """1x1 head: [1,B,H,W,C] bf16 -> NCHW fp32 logits [B,2,H,W]."""
    _need_cuda(x5, weight, bias)
    g, b, h, w, c = x5.shape
    assert weight.shape[0] == 2, "the fused head is built for n_classes == 2"
    out = torch.empty((g * b, 2, h, w), dtype=torch.float32, device=x5.device)
    check(_lib.load().fabric_b200_outconv(_p(x5), _p(weight.detach().reshape(2, c).contiguous()), _p(bias.detach()), _p(out),
                                          g * b, h, w, c, _stream()), "outconv")
    _count()
    return out
