"""GPU bring-up checks for the conv kernel and the whole eval forward (development tool, run under gpurun).

    python tools/gpu_check.py <group> [...]     groups: tap halo misc model bench

Each group runs in its own process (a trapping kernel poisons the CUDA context), results are printed as one
JSON line per case and appended to gpurun_out/gpu_check.jsonl.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from fabric_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def emit(rec):
    line = json.dumps(rec)
    print(line, flush=True)
    with open(os.path.join(OUT, "gpu_check.jsonl"), "a") as f:
        f.write(line + "\n")


def ref_conv(x5, w, scale, shift, relu):
    """fp32 reference on the bf16-rounded operands: x5 [G,B,H,W,Cin_pad] bf16, w [Cout,Cin,3,3] fp32."""
    g, b, h, wd, cp = x5.shape
    cin = w.shape[1]
    x = x5.reshape(g * b, h, wd, cp)[..., :cin].permute(0, 3, 1, 2).float()
    wq = w.bfloat16().float()
    y = F.conv2d(x, wq, None, padding=1)
    if scale is not None:
        y = y * scale[None, :, None, None] + shift[None, :, None, None]
    if relu:
        y = y.relu()
    return y  # [G*B, Cout, H, W] fp32


def conv_case(name, G, B, H, W, cin, cout, tune, relu=True, affine=True, pool=False, stats=False, head=False, seed=0):
    torch.manual_seed(seed)
    dev = "cuda"
    cp = ops.cpad(cin)
    x5 = torch.zeros(G, B, H, W, cp, device=dev, dtype=torch.bfloat16)
    x5[..., :cin] = torch.randn(G, B, H, W, cin, device=dev).bfloat16()
    w = torch.randn(cout, cin, 3, 3, device=dev) / (3.0 * cin ** 0.5)
    scale = (0.5 + torch.rand(cout, device=dev)) if affine else None
    shift = (0.3 * torch.randn(cout, device=dev)) if affine else None
    wp = ops.pack_conv_weight(w, 0)
    hd = None
    if head:
        hw = torch.randn(2, cout, device=dev) * 0.2
        hb = torch.randn(2, device=dev)
        hd = (hw, hb)
    rec = dict(case=name, G=G, B=B, H=H, W=W, cin=cin, cout=cout, tune=tune, pool=pool, stats=stats, head=head)
    try:
        res = ops.conv3x3(x5, wp, cout, scale, shift, relu=relu, pool=pool, stats=stats, head=hd, tune=tune)
        torch.cuda.synchronize()
        ref = ref_conv(x5, w, scale, shift, relu)
        y = res["y"].reshape(G * B, H, W, cout).permute(0, 3, 1, 2).float()
        err = (y - ref).abs().max().item()
        rel = ((y - ref).norm() / (ref.norm() + 1e-20)).item()
        rec.update(max_err=err, rel_l2=rel, ref_absmax=ref.abs().max().item())
        ok = rel < 5e-3
        refq = ref.bfloat16().float()
        if pool:
            pr = F.max_pool2d(refq, 2)
            pp = res["pool"].reshape(G * B, H // 2, W // 2, cout).permute(0, 3, 1, 2).float()
            perr = ((pp - pr).norm() / (pr.norm() + 1e-20)).item()
            rec.update(pool_rel=perr)
            ok = ok and perr < 5e-3
        if stats:
            st = res["stats"].double()  # [grid, 2, n_tile, 2]
            grid, _, nt, _ = st.shape
            ntiles = cout // nt
            tot = torch.zeros(2, cout, 2, dtype=torch.float64, device=dev)
            for c in range(grid):
                n_t = c % ntiles
                tot[:, n_t * nt:(n_t + 1) * nt] += st[c]
            rq = refq.reshape(G, B, cout, H, W).double()
            s1 = rq.sum(dim=(1, 3, 4))
            s2 = (rq * rq).sum(dim=(1, 3, 4))
            e1 = ((tot[:G, :, 0] - s1).abs().max() / (s1.abs().max() + 1e-9)).item()
            e2 = ((tot[:G, :, 1] - s2).abs().max() / (s2.abs().max() + 1e-9)).item()
            rec.update(stats_err1=e1, stats_err2=e2)
            ok = ok and e1 < 2e-2 and e2 < 2e-2
        if head:
            lr = torch.einsum("nchw,kc->nkhw", refq, hd[0]) + hd[1][None, :, None, None]
            herr = ((res["logits"] - lr).norm() / (lr.norm() + 1e-20)).item()
            rec.update(head_rel=herr)
            ok = ok and herr < 5e-3
        rec.update(ok=bool(ok))
    except Exception as e:  # noqa
        rec.update(ok=False, error=repr(e)[:400])
    emit(rec)
    return rec.get("ok", False)


def time_conv(name, G, B, H, W, cin, cout, tune, iters=10, **kw):
    dev = "cuda"
    cp = ops.cpad(cin)
    x5 = torch.randn(G, B, H, W, cp, device=dev).bfloat16()
    w = torch.randn(cout, cin, 3, 3, device=dev) / (3.0 * cin ** 0.5)
    wp = ops.pack_conv_weight(w, 0)
    scale = torch.ones(cout, device=dev)
    shift = torch.zeros(cout, device=dev)
    out = torch.empty(G, B, H, W, cout, device=dev, dtype=torch.bfloat16)
    rec = dict(case=name, G=G, B=B, H=H, W=W, cin=cin, cout=cout, tune=tune)
    try:
        for _ in range(3):
            ops.conv3x3(x5, wp, cout, scale, shift, relu=True, tune=tune, out=out, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.conv3x3(x5, wp, cout, scale, shift, relu=True, tune=tune, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * G * B * H * W * 9 * cin * cout
        rec.update(ms=ms, tflops=flops / ms / 1e9, ok=True)
    except Exception as e:  # noqa
        rec.update(ok=False, error=repr(e)[:400])
    emit(rec)


def group_tap():
    T = dict(halo=0, b_resident=0)
    conv_case("tap_single_tile", 1, 1, 16, 8, 64, 64, T)
    conv_case("tap_noaffine", 1, 1, 16, 8, 64, 64, T, relu=False, affine=False)
    conv_case("tap_multi_tile", 1, 2, 32, 24, 64, 64, T)
    conv_case("tap_edges", 2, 3, 20, 12, 64, 64, T)
    conv_case("tap_k128_n128", 1, 2, 32, 32, 128, 128, T)
    conv_case("tap_n256", 1, 2, 32, 32, 128, 256, dict(halo=0, b_resident=0, n_tile=256))
    conv_case("tap_cin16", 2, 2, 32, 32, 13, 64, T)
    conv_case("tap_small_h8", 2, 3, 8, 8, 64, 128, T)
    conv_case("tap_small_h4", 2, 5, 4, 4, 128, 128, T)
    conv_case("tap_small_h2", 2, 5, 2, 2, 128, 128, T)
    conv_case("tap_odd_45", 1, 2, 45, 45, 64, 128, T)
    conv_case("tap_odd_5", 2, 2, 5, 5, 64, 64, T)
    conv_case("tap_resident", 1, 4, 64, 64, 64, 64, dict(halo=0, b_resident=1, grid=8))
    conv_case("tap_many_tiles", 2, 8, 64, 64, 64, 128, dict(halo=0, b_resident=0))


def group_halo():
    T = dict(halo=1, b_resident=0)
    conv_case("halo_single_tile", 1, 1, 16, 8, 64, 64, T)
    conv_case("halo_multi_tile", 1, 2, 32, 24, 64, 64, T)
    conv_case("halo_edges", 2, 3, 20, 12, 64, 64, T)
    conv_case("halo_k128_n128", 1, 2, 32, 32, 128, 128, T)
    conv_case("halo_n256", 1, 2, 32, 32, 128, 256, dict(halo=1, b_resident=0, n_tile=256))
    conv_case("halo_odd_45", 1, 2, 45, 45, 64, 128, T)
    conv_case("halo_resident", 1, 4, 64, 64, 64, 64, dict(halo=1, b_resident=1, grid=8))
    conv_case("halo_many_tiles", 2, 8, 64, 64, 64, 128, T)


def group_misc():
    for halo in (0, 1):
        T = dict(halo=halo)
        conv_case(f"pool_h{halo}", 2, 2, 32, 32, 64, 64, T, pool=True)
        conv_case(f"pool_odd_h{halo}", 2, 2, 45, 45, 64, 128, T, pool=True)
        conv_case(f"stats_h{halo}", 2, 3, 32, 24, 64, 128, T, stats=True, relu=False, affine=False)
        conv_case(f"stats_odd_h{halo}", 2, 2, 45, 45, 64, 64, T, stats=True, relu=False, affine=False)
        conv_case(f"head_h{halo}", 1, 2, 32, 32, 64, 64, T, head=True)
    conv_case("pool_small", 2, 4, 4, 4, 128, 128, dict(halo=0), pool=True)
    conv_case("stats_small", 2, 5, 4, 4, 128, 128, dict(halo=0), stats=True, relu=False, affine=False)


def group_model():
    from fabric_b200 import BiDateNet
    from oracle import bidatenet_oracle as O
    sd = O.make_state_dict(seed=0)
    model = BiDateNet(13, 2)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    for (b, s, seed) in ((2, 32, 1), (1, 90, 2), (1, 256, 3)):
        for fuse in (True, False):
            rec = dict(case=f"model_eval_b{b}_s{s}_fuse{int(fuse)}")
            try:
                x1, x2, _ = O.make_inputs(b, s, seed=seed)
                t0 = time.time()
                ref = O.bidatenet_forward(x1, x2, sd, training=False)
                rec["oracle_s"] = time.time() - t0
                model.fuse_head = fuse
                with torch.no_grad():
                    out = model(x1.cuda(), x2.cuda()).cpu()
                rec.update(max_err=(out - ref).abs().max().item(), rel_l2=((out - ref).norm() / ref.norm()).item(),
                           argmax_agree=(out.argmax(1) == ref.argmax(1)).float().mean().item(),
                           ref_absmax=ref.abs().max().item())
                rec["ok"] = rec["rel_l2"] < 3e-2
            except Exception as e:  # noqa
                rec.update(ok=False, error=repr(e)[:400])
            emit(rec)


LAYERS = [  # name, G, H, cin, cout   (B = 64 pairs)
    ("inc.c1", 2, 256, 13, 64), ("inc.c2", 2, 256, 64, 64),
    ("down1.c1", 2, 128, 64, 128), ("down1.c2", 2, 128, 128, 128),
    ("down2.c1", 2, 64, 128, 256), ("down2.c2", 2, 64, 256, 256),
    ("down3.c1", 2, 32, 256, 512), ("down3.c2", 2, 32, 512, 512),
    ("down4.c1", 2, 16, 512, 512),
    ("up1.c1", 1, 32, 1024, 256), ("up1.c2", 1, 32, 256, 256),
    ("up2.c1", 1, 64, 512, 128), ("up3.c1", 1, 128, 256, 64),
    ("up4.c1", 1, 256, 128, 64),
]


def group_bench():
    B = int(os.environ.get("FB_BENCH_B", "64"))
    for name, G, H, cin, cout in LAYERS:
        tunes = [dict(halo=0, b_resident=0)]
        if cin >= 64:
            tunes.append(dict(halo=1, b_resident=0))
            tunes.append(dict(halo=1, b_resident=-1))
            if cout >= 256:
                tunes.append(dict(halo=1, n_tile=256))
                tunes.append(dict(halo=0, n_tile=256))
        for t in tunes:
            time_conv(name, G, B, H, H, cin, cout, t)


def group_fwd():
    from fabric_b200 import BiDateNet
    B = int(os.environ.get("FB_BENCH_B", "64"))
    model = BiDateNet(13, 2).cuda().eval()
    x1 = torch.randn(B, 13, 256, 256, device="cuda")
    x2 = torch.randn(B, 13, 256, 256, device="cuda")
    rec = dict(case=f"fwd_eval_b{B}")
    try:
        with torch.no_grad():
            for _ in range(3):
                model(x1, x2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                model(x1, x2)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rec.update(ms=ms, pairs_per_s=B / ms * 1e3, tflops=92.577e9 * B / ms / 1e9, ok=True)
    except Exception as e:  # noqa
        rec.update(ok=False, error=repr(e)[:400])
    emit(rec)


if __name__ == "__main__":
    for g in sys.argv[1:]:
        globals()["group_" + g]()
