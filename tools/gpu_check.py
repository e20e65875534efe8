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
        tunes = [dict(ctas=1), dict(ctas=2)]
        if cin >= 64:
            tunes.append(dict(ctas=2, b_resident=0))
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




# ================================================================================================ training path
def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _case(name, fn):
    rec = dict(case=name)
    try:
        rec.update(fn())
        rec["ok"] = bool(rec.pop("_ok"))
    except Exception as e:  # noqa
        import traceback
        rec.update(ok=False, error=repr(e)[:300], tb=traceback.format_exc()[-600:])
    emit(rec)


def group_train_units():
    dev = "cuda"
    torch.manual_seed(0)

    def bn_fwd():
        G, B, H, W, cin, C = 2, 3, 20, 12, 64, 128
        x5 = torch.randn(G, B, H, W, cin, device=dev).bfloat16()
        w = torch.randn(C, cin, 3, 3, device=dev) / 24
        bias = torch.randn(C, device=dev)
        bn = torch.nn.BatchNorm2d(C).to(dev)
        bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_()
        ref_bn = torch.nn.BatchNorm2d(C).to(dev); ref_bn.load_state_dict(bn.state_dict())
        r = ops.conv3x3(x5, ops.pack_conv_weight(w, 0), C, stats=True)
        s = ops.bn_finalize(r["stats"], bn, bias, B * H * W, G)
        a, pl = ops.bn_apply_relu(r["y"], s[0], s[1], pool=True)
        z = r["y"].float()  # stored conv output (no bias)
        outs = []
        for g in range(G):
            zg = z[g].permute(0, 3, 1, 2) + bias[None, :, None, None]
            outs.append(torch.relu(ref_bn(zg)))
        ref = torch.stack(outs).permute(0, 1, 3, 4, 2)
        e_a = _rel(a.float(), ref)
        e_p = _rel(pl.float(), F.max_pool2d(a.float().reshape(G * B, H, W, C).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).reshape(pl.shape))
        e_rm = _rel(bn.running_mean, ref_bn.running_mean)
        e_rv = _rel(bn.running_var, ref_bn.running_var)
        nbt = int(bn.num_batches_tracked)
        return dict(e_a=e_a, e_pool=e_p, e_rm=e_rm, e_rv=e_rv, nbt=nbt,
                    _ok=e_a < 5e-3 and e_p == 0 and e_rm < 1e-4 and e_rv < 1e-4 and nbt == 2)
    _case("bn_finalize_apply", bn_fwd)

    def losses():
        from oracle import bidatenet_oracle as O
        B, H, W = 3, 24, 40
        logits = torch.randn(B, 2, H, W) * 2
        labels = (torch.rand(B, H, W) < 0.2).long()
        out = {}
        ok = True
        for kind, fn in (("tversky", lambda l, t: O.tversky_loss(l, t, 0.1, 0.9)), ("dice", O.dice_loss),
                         ("jaccard", O.jaccard_loss), ("focal", lambda l, t: O.focal_loss(l, t, 2.0)),
                         ("ce", O.cross_entropy_loss)):
            for nd, lab in (("3d", labels), ("4d", labels[:, None])):
                l = logits.clone().requires_grad_(True)
                v = fn(l, lab); v.backward()
                loss, dl = ops.seg_loss_fwd_bwd(kind, logits.to(dev), lab.to(dev), 0.1, 0.9, 2.0, 1e-7)
                ev = abs(float(loss) - float(v)); eg = _rel(dl.cpu(), l.grad)
                out[f"{kind}_{nd}"] = (ev, eg)
                ok = ok and ev < 2e-6 and eg < 1e-4
        return dict(errs=out, _ok=ok)
    _case("losses", losses)

    def head_bwd():
        B, H, W, C = 2, 20, 28, 64
        u = torch.randn(1, B, H, W, C, device=dev).bfloat16()
        w = (torch.randn(2, C, 1, 1, device=dev) * 0.2).requires_grad_(True)
        b = torch.randn(2, device=dev).requires_grad_(True)
        uf = u[0].float().permute(0, 3, 1, 2).requires_grad_(True)
        dl = torch.randn(B, 2, H, W, device=dev)
        F.conv2d(uf, w, b).backward(dl)
        du, dw, db = ops.outconv_bwd(dl, u, w)
        e = (_rel(du[0].float(), uf.grad.permute(0, 2, 3, 1)), _rel(dw, w.grad), _rel(db, b.grad))
        return dict(errs=e, _ok=e[0] < 5e-3 and e[1] < 1e-4 and e[2] < 1e-4)
    _case("outconv_bwd", head_bwd)

    def bn_bwd(quad):
        G, B, H, W, C = 2, 2, 13, 10, 64
        z = torch.randn(G, B, H, W, C, device=dev).bfloat16()
        bn = torch.nn.BatchNorm2d(C).to(dev)
        bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.3)
        zf = z.float().requires_grad_(True)
        gam = bn.weight.detach().clone().requires_grad_(True)
        bet = bn.bias.detach().clone().requires_grad_(True)
        # reference forward per date group
        acts, stats = [], []
        for g in range(G):
            x = zf[g].permute(0, 3, 1, 2)
            m = x.mean((0, 2, 3)); v = x.var((0, 2, 3), unbiased=False)
            inv = torch.rsqrt(v + 1e-5)
            stats.append((m.detach(), inv.detach()))
            acts.append(torch.relu((x - m[None, :, None, None]) * inv[None, :, None, None] * gam[None, :, None, None]
                                   + bet[None, :, None, None]))
        mean = torch.stack([s[0] for s in stats]); invstd = torch.stack([s[1] for s in stats])
        scale = gam.detach()[None] * invstd
        shift = bet.detach()[None] - mean * scale
        a5, _ = ops.bn_apply_relu(z, scale.contiguous(), shift.contiguous())
        # the kernel's activation is bf16; make the reference consume exactly that tensor where it is an input
        a_ref = [a.permute(0, 2, 3, 1) for a in acts]
        if quad:
            # loss = sum(gcat[..., :C] * relu(a0*a1)) + sum(gp * maxpool(a_g))
            gcat = torch.randn(1, B, H, W, 2 * C, device=dev).bfloat16()
            gp = torch.randn(G, B, H // 2, W // 2, C, device=dev).bfloat16()
            prod = torch.relu(a_ref[0] * a_ref[1])
            loss = (gcat[0, ..., :C].float() * prod).sum()
            for g in range(G):
                loss = loss + (gp[g].float().permute(0, 3, 1, 2) * F.max_pool2d(acts[g], 2)).sum()
            loss.backward()
            dz, dg, db = ops.bn_relu_bwd(z, a5, gcat, True, gp, scale.contiguous(), shift.contiguous(), mean.contiguous(),
                                         invstd.contiguous(), gam)
        else:
            ga = torch.randn(G, B, H, W, C, device=dev).bfloat16()
            loss = sum((ga[g].float() * a_ref[g]).sum() for g in range(G))
            loss.backward()
            dz, dg, db = ops.bn_relu_bwd(z, None, ga, False, None, scale.contiguous(), shift.contiguous(), mean.contiguous(),
                                         invstd.contiguous(), gam)
        e = (_rel(dz.float(), zf.grad), _rel(dg, gam.grad), _rel(db, bet.grad))
        return dict(errs=e, _ok=e[0] < 1.5e-2 and e[1] < 1e-2 and e[2] < 1e-2)
    _case("bn_relu_bwd_plain", lambda: bn_bwd(False))
    _case("bn_relu_bwd_product_pool", lambda: bn_bwd(True))

    def up_bwd():
        out = {}
        ok = True
        for (H, W, h, w) in ((32, 32, 16, 16), (11, 11, 5, 5), (45, 45, 22, 22)):
            B, Cs, Cl = 2, 64, 128
            dcat = torch.randn(1, B, H, W, Cs + Cl, device=dev).bfloat16()
            low = torch.randn(B, Cl, h, w, device=dev, requires_grad=True)
            x1 = F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=True)
            dy, dx = H - x1.shape[2], W - x1.shape[3]
            x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
            (x1 * dcat[0, ..., Cs:].float().permute(0, 3, 1, 2)).sum().backward()
            dlow = ops.up_input_bwd(dcat, Cs, h, w)
            e = _rel(dlow[0].float(), low.grad.permute(0, 2, 3, 1))
            out[f"{H}"] = e
            ok = ok and e < 5e-3
        return dict(errs=out, _ok=ok)
    _case("up_input_bwd", up_bwd)

    def wgrad(G, B, H, W, cin, cout, wide):
        cp = ops.cpad(cin)
        x5 = torch.zeros(G, B, H, W, cp, device=dev, dtype=torch.bfloat16)
        x5[..., :cin] = torch.randn(G, B, H, W, cin, device=dev).bfloat16()
        dz = torch.randn(G, B, H, W, cout, device=dev).bfloat16()
        xf = x5[..., :cin].float().reshape(G * B, H, W, cin).permute(0, 3, 1, 2)
        wt = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
        (F.conv2d(xf, wt, padding=1) * dz.float().reshape(G * B, H, W, cout).permute(0, 3, 1, 2)).sum().backward()
        dw = ops.conv3x3_wgrad(dz, x5, cin, wide=wide)
        e = _rel(dw, wt.grad)
        return dict(err=e, _ok=e < 2e-3)
    for wide in (0, 1):
        _case(f"wgrad_64_64_w{wide}", lambda: wgrad(1, 2, 32, 24, 64, 64, wide))
        _case(f"wgrad_128_256_w{wide}", lambda: wgrad(2, 3, 20, 12, 128, 256, wide))
        _case(f"wgrad_stem13_w{wide}", lambda: wgrad(2, 2, 32, 32, 13, 64, wide))
        _case(f"wgrad_small4_w{wide}", lambda: wgrad(2, 5, 4, 4, 128, 128, wide))
        _case(f"wgrad_odd45_w{wide}", lambda: wgrad(1, 2, 45, 45, 64, 128, wide))

    def dgrad():
        G, B, H, W, cin, cout = 2, 2, 20, 12, 128, 64
        dz = torch.randn(G, B, H, W, cout, device=dev).bfloat16()
        w = torch.randn(cout, cin, 3, 3, device=dev) / 30
        x = torch.zeros(G * B, cin, H, W, device=dev, requires_grad=True)
        (F.conv2d(x, w.bfloat16().float(), padding=1) * dz.float().reshape(G * B, H, W, cout).permute(0, 3, 1, 2)).sum().backward()
        dx = ops.conv3x3(dz, ops.pack_conv_weight(w, 1), cin)["y"]
        e = _rel(dx.float().reshape(G * B, H, W, cin).permute(0, 3, 1, 2), x.grad)
        return dict(err=e, _ok=e < 5e-3)
    _case("dgrad", dgrad)


def group_train_model():
    from fabric_b200 import BiDateNet
    from fabric_b200.metrics import TverskyLoss
    from oracle import bidatenet_oracle as O
    golden = torch.load(os.path.join(ROOT, "tests", "golden", "bidatenet_golden.pt"))
    sd = O.make_state_dict(seed=0)

    def step():
        model = BiDateNet(13, 2)
        model.load_state_dict(sd)
        model = model.cuda().train()
        x1, x2, labels = golden["c1_x1"].cuda(), golden["c1_x2"].cuda(), golden["c1_labels"].cuda()
        logits = model(x1, x2)
        loss = TverskyLoss(alpha=0.1, beta=0.9)(logits, labels)
        loss.backward()
        out = dict(logits_rel=_rel(logits.detach().cpu(), golden["c1_logits_train"]),
                   loss=float(loss), loss_ref=float(golden["c1_loss_train"]))
        worst = ("", 0.0)
        errs = {}
        for k, p in model.named_parameters():
            ref_n = float(golden["c1_gradnorm/" + k])
            g = p.grad.detach().cpu()
            if k.endswith(".0.bias") or k.endswith(".3.bias"):
                continue
            en = abs(float(g.norm()) - ref_n) / (ref_n + 1e-30)
            if ("c1_grad/" + k) in golden:
                en = max(en, _rel(g, golden["c1_grad/" + k]))
            errs[k] = en
            if en > worst[1]:
                worst = (k, en)
        out["worst_grad"] = worst
        out["grad_errs"] = {k: round(v, 4) for k, v in errs.items()}
        st = model.state_dict()
        es = max(_rel(st[k[len("c1_newstat/"):]].float().cpu(), v.float()) for k, v in golden.items()
                 if k.startswith("c1_newstat/") and "num_batches" not in k)
        out["stats_rel"] = es
        out["nbt"] = (int(st["inc.conv.conv.1.num_batches_tracked"]), int(st["up1.conv.conv.1.num_batches_tracked"]))
        out["_ok"] = out["logits_rel"] < 3e-2 and abs(out["loss"] - out["loss_ref"]) < 5e-3 and worst[1] < 0.15 and es < 2e-2
        return out
    _case("train_step_c1_vs_reference_golden", step)


if __name__ == "__main__":
    for g in sys.argv[1:]:
        globals()["group_" + g]()
