"""GPU parity tests (run with ``-m gpu`` on a B200): the CUDA path, called through the C ABI, against
(1) the CPU oracle on seeded inputs, (2) the committed golden fixtures the unmodified reference produced, and
(3) size-independent properties at the BASELINE batch size.

Tolerances (stated): the kernels compute in bf16 with fp32 accumulation, the oracle in fp32.
  * one conv on bf16-rounded operands vs fp32 conv of the same operands: relative L2 <= 5e-3 (bf16 output rounding
    is 2^-9 = 2e-3 relative per element);
  * whole network (18 convs deep) logits vs fp32 oracle: relative L2 <= 1e-2, argmax agreement >= 99.9 %.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONV_TOL = 5e-3
NET_TOL = 1e-2


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from fabric_b200 import _lib
    _lib.load()                       # the extension must be there: no silent fallback
    return torch.device("cuda:0")


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def conv_ref_cpu(x5, w, scale, shift, relu):
    """oracle arithmetic (F.conv2d fp32 on CPU, reference models/unet_parts.py:13) on the bf16-rounded operands"""
    g, b, h, wd, cp = x5.shape
    cin = w.shape[1]
    x = x5.reshape(g * b, h, wd, cp)[..., :cin].permute(0, 3, 1, 2).float().cpu()
    y = F.conv2d(x, w.bfloat16().float().cpu(), None, padding=1)
    if scale is not None:
        y = y * scale.cpu()[None, :, None, None] + shift.cpu()[None, :, None, None]
    return y.relu() if relu else y


CONV_CASES = [
    # G, B, H, W, cin, cout, tune
    (1, 1, 16, 8, 64, 64, dict(halo=0)),
    (1, 1, 16, 8, 64, 64, dict(halo=1)),
    (2, 3, 20, 12, 64, 64, dict(halo=0)),            # ragged tiles: H, W not multiples of the 16x8 tile
    (2, 3, 20, 12, 64, 64, dict(halo=1)),
    (1, 2, 32, 32, 128, 128, dict(halo=1)),
    (1, 2, 32, 32, 128, 256, dict(halo=1, n_tile=256)),
    (1, 2, 32, 32, 128, 256, dict(halo=0, n_tile=128)),
    (2, 2, 32, 32, 13, 64, dict()),                    # 13-band stem, channels padded to 16 (halo box, 32B swizzle)
    (2, 2, 32, 32, 13, 64, dict(halo=0)),              # 13-band stem, nine tap boxes per stage
    (1, 3, 45, 45, 13, 64, dict(halo=1)),
    (2, 3, 8, 8, 64, 128, dict()),                     # small maps: tiles span several images
    (2, 5, 4, 4, 128, 128, dict()),
    (2, 5, 2, 2, 128, 128, dict()),
    (1, 2, 45, 45, 64, 128, dict(halo=1)),             # odd sizes (patch 90 path)
    (2, 2, 5, 5, 64, 64, dict()),
    (1, 4, 64, 64, 64, 64, dict(halo=1, b_resident=1, grid=8)),   # weights resident in smem, few persistent CTAs
    (1, 1, 1, 1, 64, 64, dict()),                      # degenerate 1x1 map
]


@pytest.mark.parametrize("G,B,H,W,cin,cout,tune", CONV_CASES)
def test_conv3x3_matches_oracle(cuda, G, B, H, W, cin, cout, tune):
    from fabric_b200 import ops
    torch.manual_seed(G * 1000 + H * 10 + cin)
    cp = ops.cpad(cin)
    x5 = torch.zeros(G, B, H, W, cp, device=cuda, dtype=torch.bfloat16)
    x5[..., :cin] = torch.randn(G, B, H, W, cin, device=cuda).bfloat16()
    w = torch.randn(cout, cin, 3, 3, device=cuda) / (3.0 * cin ** 0.5)
    scale = 0.5 + torch.rand(cout, device=cuda)
    shift = 0.3 * torch.randn(cout, device=cuda)
    res = ops.conv3x3(x5, ops.pack_conv_weight(w, 0), cout, scale, shift, relu=True, tune=tune)
    y = res["y"].reshape(G * B, H, W, cout).permute(0, 3, 1, 2).float().cpu()
    ref = conv_ref_cpu(x5, w, scale, shift, True)
    assert rel(y, ref) < CONV_TOL
    assert (y - ref).abs().max() <= 2 ** -7 * ref.abs().max() + 1e-6


@pytest.mark.parametrize("G,B,H,W,cin,cout,tune", CONV_CASES)
def test_conv3x3_folded_scale_and_shift_in_accumulator(cuda, G, B, H, W, cin, cout, tune):
    """Eval-mode form: BatchNorm scale folded into the packed weights, accumulator primed with the shift (TMEM preload by
    the epilogue warps): same result as the per-element affine epilogue up to the bf16 rounding of w * scale."""
    from fabric_b200 import ops
    torch.manual_seed(G * 1000 + H * 10 + cin + 1)
    cp = ops.cpad(cin)
    x5 = torch.zeros(G, B, H, W, cp, device=cuda, dtype=torch.bfloat16)
    x5[..., :cin] = torch.randn(G, B, H, W, cin, device=cuda).bfloat16()
    w = torch.randn(cout, cin, 3, 3, device=cuda) / (3.0 * cin ** 0.5)
    scale = 0.5 + torch.rand(cout, device=cuda)
    shift = 0.3 * torch.randn(cout, device=cuda)
    wf = ops.pack_conv_weight(w, 0, scale=scale)
    assert torch.equal(wf[..., :cin], (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(cout, 9, cin).bfloat16())
    kw = dict(pool=True) if (H % 2 == 0 and W % 2 == 0 and H > 1) else {}
    res = ops.conv3x3(x5, wf, cout, None, shift, relu=True, tune=tune, shift_in_acc=True, **kw)
    y = res["y"].reshape(G * B, H, W, cout).permute(0, 3, 1, 2).float().cpu()
    x = x5.reshape(G * B, H, W, cp)[..., :cin].permute(0, 3, 1, 2).float().cpu()
    ref = (F.conv2d(x, (w * scale[:, None, None, None]).bfloat16().float().cpu(), None, padding=1)
           + shift.cpu()[None, :, None, None]).relu()
    assert rel(y, ref) < CONV_TOL
    assert (y - ref).abs().max() <= 2 ** -7 * ref.abs().max() + 1e-6
    if kw:
        assert torch.equal(res["pool"].reshape(G * B, H // 2, W // 2, cout).permute(0, 3, 1, 2).float().cpu(),
                           F.max_pool2d(y, 2))
    # twice in a row on the same stream: the re-primed accumulators of the first launch must not leak into the second
    res2 = ops.conv3x3(x5, wf, cout, None, shift, relu=True, tune=tune, shift_in_acc=True)
    assert torch.equal(res2["y"], res["y"])


@pytest.mark.parametrize("halo", [0, 1])
def test_conv3x3_fused_pool_stats_head(cuda, halo):
    from fabric_b200 import ops
    torch.manual_seed(3)
    G, B, H, W, cin, cout = 2, 2, 45, 45, 64, 64
    x5 = torch.randn(G, B, H, W, cin, device=cuda).bfloat16()
    w = torch.randn(cout, cin, 3, 3, device=cuda) / (3.0 * cin ** 0.5)
    hw, hb = torch.randn(2, cout, device=cuda) * 0.2, torch.randn(2, device=cuda)
    # eval-style epilogue options (pool + 1x1 head) in one launch; the BatchNorm moments come from the training
    # instantiation (raw-accumulator epilogue, sums accumulated in registers) in a second launch of the same conv
    res = ops.conv3x3(x5, ops.pack_conv_weight(w, 0), cout, None, None, relu=False, pool=True,
                      head=(hw, hb), tune=dict(halo=halo))
    res_t = ops.conv3x3(x5, ops.pack_conv_weight(w, 0), cout, stats=True, tune=dict(halo=halo))
    assert torch.equal(res_t["y"], res["y"])
    res["stats"] = res_t["stats"]
    ref = conv_ref_cpu(x5, w, None, None, False)
    refq = ref.bfloat16().float()            # the stored activation is bf16; fused consumers see the stored value
    y = res["y"].reshape(G * B, H, W, cout).permute(0, 3, 1, 2).float().cpu()
    assert rel(y, ref) < CONV_TOL
    # nn.MaxPool2d(2) (floor) of the stored tensor: exact
    pool = res["pool"].reshape(G * B, H // 2, W // 2, cout).permute(0, 3, 1, 2).float().cpu()
    assert torch.equal(pool, F.max_pool2d(y, 2))
    # BatchNorm moments per date group of the stored tensor
    st = res["stats"].double().cpu()
    tot = st.sum(0)                          # one N tile -> every CTA holds the same channels
    yq = y.reshape(G, B, cout, H, W).double()
    assert torch.allclose(tot[:, :, 0], yq.sum(dim=(1, 3, 4)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[:, :, 1], (yq * yq).sum(dim=(1, 3, 4)), rtol=1e-4, atol=1e-2)
    # fused 1x1 head on the stored tensor
    lr = torch.einsum("nchw,kc->nkhw", y, hw.cpu()) + hb.cpu()[None, :, None, None]
    assert rel(res["logits"].cpu(), lr) < 1e-4
    del refq


def test_conv3x3_linearity_and_halo_equals_tap(cuda):
    """conv(2x) == 2 conv(x) exactly (power-of-two scaling commutes with bf16/fp32 rounding) and the two operand
    feeding modes give bit-identical results (same products, same accumulation order)."""
    from fabric_b200 import ops
    torch.manual_seed(5)
    x5 = torch.randn(1, 2, 48, 40, 128, device=cuda).bfloat16()
    wp = ops.pack_conv_weight(torch.randn(128, 128, 3, 3, device=cuda) / 30, 0)
    a = ops.conv3x3(x5, wp, 128, tune=dict(halo=1))["y"]
    b = ops.conv3x3(x5 * 2, wp, 128, tune=dict(halo=1))["y"]
    c = ops.conv3x3(x5, wp, 128, tune=dict(halo=0))["y"]
    assert torch.equal(a.float() * 2, b.float())
    assert torch.equal(a, c)


@pytest.mark.parametrize("H,W,cin,cout,tune", [(32, 32, 64, 64, dict()), (45, 45, 64, 128, dict()), (20, 12, 128, 256, dict()),
                                                (8, 8, 64, 64, dict()), (32, 24, 64, 64, dict(grid=4))])
def test_conv3x3_fused_date_product(cuda, H, W, cin, cout, tune):
    """relu(y[date 1] * y[date 0]) (reference bidate_model.py:35-38) fused into the conv epilogue == product of the
    stored outputs, bit-exact; the conv outputs themselves are unchanged by the pair scheduling."""
    from fabric_b200 import ops
    torch.manual_seed(11)
    B = 3
    x5 = torch.randn(2, B, H, W, cin, device=cuda).bfloat16()
    wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device=cuda) / (3.0 * cin ** 0.5), 0)
    cat = torch.full((1, B, H, W, cout + 64), 7.0, device=cuda, dtype=torch.bfloat16)
    plain = ops.conv3x3(x5, wp, cout, relu=True, tune=tune)["y"]
    fused = ops.conv3x3(x5, wp, cout, relu=True, tune=tune, prod_out=cat, pool=True)
    assert torch.equal(fused["y"], plain)
    assert torch.equal(cat[0, ..., :cout], torch.relu(plain[0].float() * plain[1].float()).bfloat16())
    assert bool((cat[0, ..., cout:] == 7.0).all())                 # the upsample half is left untouched
    if cout <= 128:
        # eval-mode variant: the full-resolution output is not written at all, only the pooled copy and the product
        cat2 = torch.full_like(cat, 7.0)
        lean = ops.conv3x3(x5, wp, cout, relu=True, tune=tune, prod_out=cat2, pool=True, store_main=False)
        assert lean["y"] is None
        assert torch.equal(cat2, cat)
        assert torch.equal(lean["pool"], fused["pool"])


def test_fused_and_unfused_decoder_inputs_agree(cuda):
    from oracle import bidatenet_oracle as O
    model = _model(cuda, O.make_state_dict(seed=0))
    for (b, s, seed) in ((2, 32, 1), (1, 90, 2), (2, 128, 3)):
        x1, x2, _ = O.make_inputs(b, s, seed=seed)
        with torch.no_grad():
            model.fuse_product = True
            a = model(x1.to(cuda), x2.to(cuda))
            model.fuse_product = False
            c = model(x1.to(cuda), x2.to(cuda))
        model.fuse_product = True
        assert torch.equal(a, c)


def test_pack_unpack_roundtrip(cuda):
    from fabric_b200 import ops
    x = torch.randn(3, 13, 37, 70, device=cuda)
    p = ops.pack_input(x)
    assert p.shape == (3, 37, 70, 16) and p.dtype == torch.bfloat16
    assert torch.equal(p[..., :13].permute(0, 3, 1, 2).float(), x.bfloat16().float())
    assert torch.count_nonzero(p[..., 13:]) == 0
    y = torch.randn(2, 19, 33, 64, device=cuda).bfloat16()
    assert torch.equal(ops.unpack_output(y), y.permute(0, 3, 1, 2).float())


@pytest.mark.parametrize("H,W,h,w,lg", [(32, 32, 16, 16, 2), (11, 11, 5, 5, 1), (45, 45, 22, 22, 1), (8, 8, 4, 4, 2)])
def test_build_up_input_matches_oracle(cuda, H, W, h, w, lg):
    """relu(d2*d1) skip + bilinear(align_corners=True) x2 + F.pad + cat -- reference bidate_model.py:35-38,
    unet_parts.py:56-58,65-78"""
    from fabric_b200 import ops
    torch.manual_seed(7)
    B, Cs, Cl = 2, 64, 128
    skip = torch.randn(2, B, H, W, Cs, device=cuda).relu().bfloat16()
    low = torch.randn(lg, B, h, w, Cl, device=cuda).relu().bfloat16()
    out = ops.build_up_input(skip, low)[0].float().cpu()
    s = skip.float().cpu()
    lo = low.float().cpu()
    lo = torch.relu(lo[0] * lo[1]) if lg == 2 else lo[0]
    x1 = F.interpolate(lo.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = H - x1.shape[2], W - x1.shape[3]
    x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
    ref = torch.cat([torch.relu(s[1] * s[0]).permute(0, 3, 1, 2), x1], 1).permute(0, 2, 3, 1)
    assert rel(out, ref) < 3e-3
    assert (out - ref).abs().max() <= 2 ** -7 * ref.abs().max()


def _model(cuda, sd):
    from fabric_b200 import BiDateNet
    m = BiDateNet(13, 2)
    m.load_state_dict(sd)
    return m.to(cuda).eval()


@pytest.mark.parametrize("key,fuse", [("c1", True), ("c1", False), ("p90", True), ("p90", False)])
def test_eval_forward_matches_reference_golden(cuda, golden, key, fuse):
    """config 1 (B=2, 13x32x32) and the reference's default patch 90 (F.pad branch), against the logits the
    unmodified reference produced."""
    from oracle import bidatenet_oracle as O
    model = _model(cuda, O.make_state_dict(seed=0))
    model.fuse_head = fuse
    with torch.no_grad():
        out = model(golden[f"{key}_x1"].to(cuda), golden[f"{key}_x2"].to(cuda)).cpu()
    ref = golden[f"{key}_logits_eval"]
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel(out, ref) < NET_TOL
    assert (out.argmax(1) == ref.argmax(1)).float().mean() >= 0.999


def test_eval_forward_256_matches_golden_and_oracle(cuda, golden):
    from oracle import bidatenet_oracle as O
    sd = O.make_state_dict(seed=0)
    x1, x2, _ = O.make_inputs(1, 256, seed=3)
    assert torch.allclose(x1.double().sum(), golden["p256_x1_sum"])
    model = _model(cuda, sd)
    with torch.no_grad():
        out = model(x1.to(cuda), x2.to(cuda)).cpu()
    ref = golden["p256_logits_eval"]
    assert rel(out, ref) < NET_TOL
    assert (out.argmax(1) == ref.argmax(1)).float().mean() >= 0.999


def test_encoder_blocks_match_oracle(cuda):
    """per-block parity (standalone NCHW fp32 entry points of the mirrored modules) vs the oracle"""
    from oracle import bidatenet_oracle as O
    sd = O.make_state_dict(seed=0)
    model = _model(cuda, sd)
    x1, _, _ = O.make_inputs(2, 32, seed=11)
    with torch.no_grad():
        a = model.inc(x1.to(cuda)).cpu()
        ref_a = O.inconv(x1, sd)
        assert rel(a, ref_a) < 5e-3
        b = model.down1(ref_a.to(cuda)).cpu()
        assert rel(b, O.down(ref_a, sd, "down1")) < 5e-3
        lo = O.down(ref_a, sd, "down1")
        u = model.up4  # up(128, 64): skip 64 ch @32, low 64 ch @16
        low = torch.rand(2, 64, 16, 16)
        skip = torch.rand(2, 64, 32, 32)
        c = u(low.to(cuda), skip.to(cuda)).cpu()
        assert rel(c, O.up(low, skip, sd, "up4")) < 5e-3
        lg = model.outc(skip.to(cuda)).cpu()
        assert rel(lg, O.outconv(skip, sd)) < 5e-3
    del lo


def test_full_batch_properties(cuda):
    """BASELINE size (64 pairs of 13x256x256): properties that need no oracle run.
    (a) a pair's logits do not depend on its batch neighbours (eval mode): bit-exact vs running it alone;
    (b) the network is symmetric in its two dates (shared encoder, product fusion): bit-exact swap."""
    from oracle import bidatenet_oracle as O
    model = _model(cuda, O.make_state_dict(seed=0))
    g = torch.Generator(device=cuda).manual_seed(5)
    x1 = torch.randn(64, 13, 256, 256, device=cuda, generator=g)
    x2 = torch.randn(64, 13, 256, 256, device=cuda, generator=g)
    with torch.no_grad():
        full = model(x1, x2)
        assert full.shape == (64, 2, 256, 256)
        assert torch.isfinite(full).all()
        for i in (0, 17, 63):
            assert torch.equal(model(x1[i:i + 1], x2[i:i + 1])[0], full[i])
        assert torch.equal(model(x2, x1), full)
        # and one pair of the big batch against the fp32 oracle
        ref = O.bidatenet_forward(x1[5:6].cpu(), x2[5:6].cpu(), O.make_state_dict(seed=0))
        assert rel(full[5:6].cpu(), ref) < NET_TOL


def test_state_dict_and_pickle_roundtrip_on_device(cuda, tmp_path):
    """train.py:222 pickles the whole model; reloading must give identical logits"""
    from oracle import bidatenet_oracle as O
    model = _model(cuda, O.make_state_dict(seed=0))
    x1, x2, _ = O.make_inputs(1, 32, seed=4)
    with torch.no_grad():
        a = model(x1.to(cuda), x2.to(cuda))
    p = tmp_path / "ckpt.pt"
    torch.save(model, p)
    m2 = torch.load(p, weights_only=False)
    with torch.no_grad():
        b = m2(x1.to(cuda), x2.to(cuda))
    assert torch.equal(a, b)


def test_host_pipeline_matches_direct_call(cuda):
    from fabric_b200.inference import predict_patches
    from oracle import bidatenet_oracle as O
    model = _model(cuda, O.make_state_dict(seed=0))
    x1, x2, _ = O.make_inputs(5, 64, seed=6)
    with torch.no_grad():
        direct = model(x1.to(cuda), x2.to(cuda)).cpu()
    logits = predict_patches(model, x1.numpy(), x2.numpy(), batch_size=2, return_logits=True)
    assert torch.equal(logits, direct)
    mask = predict_patches(model, x1.numpy(), x2.numpy(), batch_size=2, return_logits=False)
    assert torch.equal(mask, direct.argmax(1).to(torch.uint8))


# ------------------------------------------------------------------------------------------------ scene tiling (8f)
def test_gather_argmax_scatter_match_oracle(cuda):
    """device gather (+ fused z-score, fp32 and uint16 scenes) == oracle get_patches; argmax/metrics == torch;
    reassembly == oracle get_bands including the later-write-wins overlaps"""
    import numpy as np
    from fabric_b200 import ops
    from fabric_b200.scene import SceneTiler
    from oracle import bidatenet_oracle as O
    torch.manual_seed(13)
    h, w, p = 200, 150, 64
    raw = torch.randint(0, 4000, (13, h, w), dtype=torch.int32)
    mean, std = torch.rand(13) * 2000 + 500, torch.rand(13) * 900 + 100
    scene = ((raw.float() - mean[:, None, None]) / std[:, None, None])
    tiler = SceneTiler(h, w, p, cuda)
    ref_patches, hs, ws, lc, lr, _, _ = O.get_patches(scene.permute(1, 2, 0).numpy(), p)     # [N,p,p,13]
    refq = torch.from_numpy(ref_patches).bfloat16().float()
    t32 = tiler.gather(scene.to(cuda), 0, tiler.n)
    assert t32.shape == (tiler.n, p, p, 16)
    assert torch.equal(t32[..., :13].float().cpu(), refq) and torch.count_nonzero(t32[..., 13:]) == 0
    t16 = tiler.gather(raw.to(torch.uint16).to(cuda), 0, tiler.n, mean.to(cuda), (1.0 / std).to(cuda))
    assert ((t16[..., :13].float().cpu() - refq).abs() <= 2 ** -7 * refq.abs() + 1e-6).all()
    # argmax + confusion counts
    logits = torch.randn(tiler.n, 2, p, p, device=cuda)
    logits[0, :, 0, :8] = 0.0                                        # ties -> class 0 like torch.max
    labels = (torch.rand(tiler.n, p, p, device=cuda) < 0.3).long()
    mask, counts = ops.argmax_metrics(logits, labels)
    pred = torch.max(logits, 1)[1]
    assert torch.equal(mask.long(), pred)
    tp = int(((pred == 1) & (labels == 1)).sum()); fp = int(((pred == 1) & (labels == 0)).sum())
    fn = int(((pred == 0) & (labels == 1)).sum()); tn = int(((pred == 0) & (labels == 0)).sum())
    assert counts.tolist() == [tp, fp, fn, tn]
    # reassembly with overlaps
    canvas = tiler.reassemble(mask)
    ref_canvas = O.get_bands(mask.cpu().numpy().astype(np.float64), hs, ws, lc, lr, h, w, p)
    assert np.array_equal(canvas.cpu().numpy().astype(np.float64), ref_canvas)


def test_predict_scene_matches_reference_pipeline(cuda):
    """whole config-5 pipeline on a small scene vs oracle: get_patches -> BiDateNet (fp32 oracle) -> argmax -> get_bands"""
    import numpy as np
    from fabric_b200.scene import predict_scene
    from oracle import bidatenet_oracle as O
    sd = O.make_state_dict(seed=0)
    model = _model(cuda, sd)
    g = torch.Generator().manual_seed(17)
    h, w, p = 150, 100, 64
    d1, d2 = torch.randn(13, h, w, generator=g), torch.randn(13, h, w, generator=g)
    canvas, info = predict_scene(model, d1.to(cuda), d2.to(cuda), patch_size=p, batch_size=4)
    p1, hs, ws, lc, lr, _, _ = O.get_patches(d1.permute(1, 2, 0).numpy(), p)
    p2 = O.get_patches(d2.permute(1, 2, 0).numpy(), p)[0]
    assert info["tiles"] == p1.shape[0]
    with torch.no_grad():
        logits = O.bidatenet_forward(torch.from_numpy(p1).permute(0, 3, 1, 2), torch.from_numpy(p2).permute(0, 3, 1, 2), sd)
    ref = O.get_bands(torch.max(logits, 1)[1].numpy().astype(np.float64), hs, ws, lc, lr, h, w, p)
    agree = (canvas.cpu().numpy().astype(np.float64) == ref).mean()
    assert agree >= 0.999, agree


def test_batch_feeder_double_buffer(cuda):
    """Host batches staged on the side stream arrive intact and in order, also when the consumer is slow."""
    from fabric_b200.inference import BatchFeeder
    feeder = BatchFeeder(cuda)
    batches = [(torch.full((4, 13, 16, 16), float(i)).pin_memory(), torch.full((4, 16, 16), i, dtype=torch.int64).pin_memory())
               for i in range(5)]
    feeder.prefetch(batches[0])
    seen = []
    for i in range(5):
        x, lab = feeder.next()
        if i + 1 < 5:
            feeder.prefetch(batches[i + 1])
        y = (x * 2).sum() + lab.sum()          # "the step"
        torch.cuda._sleep(2_000_000)           # keep the stream busy while the next copy lands in the other slot
        seen.append(y)
        feeder.release()
    torch.cuda.synchronize()
    for i, y in enumerate(seen):
        assert float(y) == 2.0 * i * 4 * 13 * 256 + i * 4 * 256


def test_raw_uint16_input_matches_oracle_on_normalised_input(cuda):
    """Raw Sentinel-2 digital numbers in, z-score fused into the pack kernel (reference utils/dataloaders.py:94-99 with the
    band statistics of metadata.json:4-29) == the oracle fed with the host-normalised fp32 patches."""
    from fabric_b200 import ops
    from fabric_b200.inference import HostPipeline
    from oracle import bidatenet_oracle as O
    sd = O.make_state_dict(seed=0)
    model = _model(cuda, sd)
    mean = torch.tensor([1617.57, 1422.37, 1359.37, 1414.68, 1557.94, 1986.22, 2210.50, 2118.56, 2344.79, 711.84, 15.75,
                         2133.90, 1584.27])
    std = torch.tensor([319.12, 456.25, 590.13, 849.37, 811.31, 813.55, 891.85, 901.61, 954.77, 370.95, 9.23, 1116.59, 985.12])
    g = torch.Generator().manual_seed(5)
    raw = [(torch.randn(2, 13, 32, 32, generator=g) * std[None, :, None, None] + mean[None, :, None, None])
           .round().clamp(0, 65535).to(torch.int32).to(torch.uint16) for _ in range(2)]
    norm = [(r.to(torch.int32).float() - mean[None, :, None, None]) / std[None, :, None, None] for r in raw]   # the loader
    ref = O.bidatenet_forward(norm[0], norm[1], sd, training=False)
    model.set_input_normalisation(mean, std)
    assert "_fb_in_mean" not in model.state_dict()
    # pack kernel alone: bit-exact against the same arithmetic in torch
    packed = ops.pack_input_raw(raw[0].to(cuda), model._fb_in_mean, model._fb_in_inv_std)
    want = ((raw[0].to(torch.int32).float().to(cuda) - model._fb_in_mean[None, :, None, None])
            * model._fb_in_inv_std[None, :, None, None]).permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(packed[..., :13], want) and bool((packed[..., 13:] == 0).all())
    with torch.no_grad():
        out = model(raw[0].to(cuda), raw[1].to(cuda)).cpu()
    assert ((out - ref).norm() / ref.norm()).item() <= 1e-2
    # through the host pipeline (pinned uint16 in, logits out)
    host_out = torch.empty(2, 2, 32, 32).pin_memory()
    pipe = HostPipeline(model, chunk=2, n_channels=13, size=32, return_logits=True)
    h2d, _ = pipe.run(raw[0].pin_memory(), raw[1].pin_memory(), host_out)
    torch.cuda.synchronize()
    assert h2d == 2 * raw[0].numel() * 2 and torch.equal(host_out, out)


def test_fused_augmentation_matches_reference_loader(cuda):
    """rot90 / flips fused into the input pack (and applied to the labels) == the reference loader's numpy ops
    (utils/dataloaders.py:152-163) followed by the plain pack, bit-exact, fp32 and raw uint16 inputs."""
    import numpy as np
    from fabric_b200 import ops
    from oracle import bidatenet_oracle as O
    rng = np.random.default_rng(9)
    B, C, S = 8, 13, 24
    img = rng.standard_normal((B, 2, C, S, S)).astype(np.float32)
    lbl = rng.integers(0, 2, (B, S, S)).astype(np.int64)
    params = [(b % 4, (b // 2) % 2, (b // 4) % 2) for b in range(B)] 
    aug = torch.tensor(params, dtype=torch.int32, device=cuda)
    want_img = np.stack([O.augment_patch(img[b], lbl[b], *params[b])[0] for b in range(B)])
    want_lbl = np.stack([O.augment_patch(img[b], lbl[b], *params[b])[1] for b in range(B)])
    for d in range(2):
        got = ops.pack_input_aug(torch.from_numpy(img[:, d].copy()).to(cuda), aug)
        ref = ops.pack_input(torch.from_numpy(want_img[:, d].copy()).to(cuda))
        assert torch.equal(got, ref)
    assert torch.equal(ops.augment_labels(torch.from_numpy(lbl).to(cuda), aug).cpu(), torch.from_numpy(want_lbl.copy()))
    # raw uint16 + z-score + augmentation in one pass
    raw = rng.integers(0, 4000, (B, C, S, S)).astype(np.uint16)
    mean = torch.linspace(500, 2500, C, device=cuda)
    inv_std = 1.0 / torch.linspace(200, 900, C, device=cuda)
    want_raw = np.stack([O.augment_patch(np.stack([raw[b], raw[b]]), lbl[b], *params[b])[0][0] for b in range(B)])
    got = ops.pack_input_aug(torch.from_numpy(raw).to(cuda), aug, mean, inv_std)
    ref = ops.pack_input_raw(torch.from_numpy(want_raw.copy()).to(cuda), mean, inv_std)
    assert torch.equal(got, ref)
    # through the model entry point
    sd = O.make_state_dict(seed=0)
    model = _model(cuda, sd)
    x1, x2 = torch.from_numpy(img[:, 0].copy()).to(cuda), torch.from_numpy(img[:, 1].copy()).to(cuda)
    with torch.no_grad():
        a = model(x1, x2, aug=aug)
        b_ = model(torch.from_numpy(want_img[:, 0].copy()).to(cuda), torch.from_numpy(want_img[:, 1].copy()).to(cuda))
    assert torch.equal(a, b_)
