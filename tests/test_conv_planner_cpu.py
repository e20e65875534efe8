"""CPU: the conv launch planner (pure host arithmetic behind fabric_b200_conv3x3_plan) on every BiDateNet layer shape of
BASELINE configs 2-4 (64 pairs of 13x256x256) and on the ragged shapes of the GPU parity tests: shared memory fits the
B200 limit, persistent slots cover the work, CTA pairs / resident weights / epilogue warps are chosen as DESIGN.md says."""
import ctypes as C

import pytest

SMS, SMEM = 148, 232448          # B200: SM count, opt-in shared memory per block


@pytest.fixture(scope="module")
def lib(built_lib):
    return built_lib


def plan(lib, G, B, H, W, cin, cout, **kw):
    from fabric_b200 import _lib
    d = _lib.Conv3x3Desc()
    d.G, d.B, d.H, d.W, d.Cin, d.Cout = G, B, H, W, (16 if cin <= 16 else cin), cout
    d.relu, d.store_main = int(kw.get("relu", 1)), int(kw.get("store_main", 1))
    d.x = d.w = 0x1000
    d.y = 0x1000 if d.store_main else None
    for k in ("pool_out", "stats_ws", "prod_out", "head_out", "head_w", "head_b", "shift", "scale", "bnbwd_z", "bnbwd_coef"):
        if kw.get(k):
            setattr(d, k, 0x1000)
    d.prod_channels = kw.get("prod_channels", 0)
    d.shift_in_acc = int(kw.get("shift_in_acc", 0))
    d.tune = _lib.ConvTuning(kw.get("n_tile", 0), -1, 0, 0, -1, 0, kw.get("ctas", 0), kw.get("epi_warps", 0),
                             kw.get("occupancy", 0))
    out = _lib.ConvPlan()
    rc = lib.fabric_b200_conv3x3_plan(C.byref(d), SMS, SMEM, C.byref(out))
    return rc, out


# name, G, H, cin, cout  (B = 64 pairs)
LAYERS = [("inc.c1", 2, 256, 13, 64), ("inc.c2", 2, 256, 64, 64), ("down1.c1", 2, 128, 64, 128), ("down1.c2", 2, 128, 128, 128),
          ("down2.c1", 2, 64, 128, 256), ("down2.c2", 2, 64, 256, 256), ("down3.c1", 2, 32, 256, 512),
          ("down3.c2", 2, 32, 512, 512), ("down4.c1", 2, 16, 512, 512), ("up1.c1", 1, 32, 1024, 256), ("up1.c2", 1, 32, 256, 256),
          ("up2.c1", 1, 64, 512, 128), ("up2.c2", 1, 64, 128, 128), ("up3.c1", 1, 128, 256, 64), ("up3.c2", 1, 128, 64, 64),
          ("up4.c1", 1, 256, 128, 64), ("up4.c2", 1, 256, 64, 64)]


@pytest.mark.parametrize("name,G,H,cin,cout", LAYERS)
@pytest.mark.parametrize("mode", ["eval", "train_fwd", "dgrad", "dgrad_bnbwd"])
def test_every_layer_has_a_valid_pair_plan(lib, name, G, H, cin, cout, mode):
    kw = {}
    if mode == "eval":
        kw = dict(shift=1, shift_in_acc=1)
    elif mode == "train_fwd":
        kw = dict(stats_ws=1, relu=0)           # raw accumulator + BatchNorm moment partials
    else:
        kw = dict(relu=0)
        if mode == "dgrad_bnbwd":               # data gradient of c2 with the BatchNorm-1 backward reduce in its epilogue
            if not name.endswith(".c2"):
                pytest.skip("only the second conv's data gradient feeds a BatchNorm backward")
            kw.update(stats_ws=1, bnbwd_z=1, bnbwd_coef=1)
        if cin == 13:
            pytest.skip("the stem needs no data gradient")
        cin, cout = cout, cin           # the data gradient is the same kernel with the channel roles swapped
    rc, p = plan(lib, G, 64, H, H, cin, cout, **kw)
    assert rc == 0, lib.fabric_b200_last_error()
    assert p.smem_bytes <= SMEM
    # per-channel sums of the 64- and 128-wide tiles accumulate in registers (training instantiation)
    assert p.reg_stats == (1 if "stats_ws" in kw and p.n_tile <= 128 else 0)
    assert p.ctas == 2 and p.grid % 2 == 0                            # CTA pairs on every real layer
    assert p.n_tile == (64 if cout == 64 else 256 if cout % 256 == 0 else 128)
    stem = cin <= 16
    assert p.epi_warps == (4 if stem else 8 if p.n_tile <= 128 else 4)
    assert p.ctas_per_sm == (2 if stem else 1)                         # the stem runs two CTAs per SM
    assert p.grid <= SMS * p.ctas_per_sm
    assert p.a_stages >= 2 and p.b_stages >= 1
    n_tiles = cout // p.n_tile
    assert (p.grid // 2) % n_tiles == 0                                # each CTA keeps one N tile
    m_tiles = G * 64 * (H // 16 if H >= 16 else 1) * (H // 8)
    assert p.total_units == m_tiles // 2 * n_tiles
    # resident weights: always for the small slabs, and then the slab is exactly 9 * Cin / CK stages
    cin_p = 16 if cin <= 16 else cin
    if cin_p * p.n_tile <= 128 * 64:
        assert p.b_resident == 1
    if p.b_resident:
        assert p.b_stages == 9 * (cin_p // p.ck)
        a_stage = 23552 if p.ck == 64 else 6144                                # halo tile: 18 x 10 pixels x CK channels
        assert 9 * cin_p * p.n_tile + p.a_stages * a_stage <= p.smem_bytes     # half the N tile per CTA of the pair
    if name == "down1.c2" and mode == "eval":
        assert p.b_resident == 1                                               # the 147 KB slab fits next to two input stages


def test_lean_eval_encoder_levels(lib):
    # inc.c2 / down1.c2 in eval: pooled copy + product, no full-resolution output, two staging buffers, both through TMA
    rc, p = plan(lib, 2, 64, 256, 256, 64, 64, pool_out=1, prod_out=1, prod_channels=128, store_main=0, shift=1, shift_in_acc=1)
    assert rc == 0 and p.out_bufs == 2 and p.pool_tma == 1 and p.prod_tma == 1 and p.b_resident == 1 and p.smem_bytes <= SMEM
    rc, p = plan(lib, 2, 64, 128, 128, 128, 128, pool_out=1, prod_out=1, prod_channels=256, store_main=0, shift=1, shift_in_acc=1)
    assert rc == 0 and p.n_tile == 64 and p.b_resident == 1 and p.out_bufs == 2 and p.smem_bytes <= SMEM
    # 256 wide: single buffer, product through L2, main output required
    rc, p = plan(lib, 2, 64, 64, 64, 256, 256, pool_out=1, prod_out=1, prod_channels=512, store_main=1)
    assert rc == 0 and p.out_bufs == 1 and p.prod_tma == 0 and p.pool_tma == 1
    rc, _ = plan(lib, 2, 64, 64, 64, 256, 256, pool_out=1, prod_out=1, prod_channels=512, store_main=0)
    assert rc != 0


@pytest.mark.parametrize("G,B,H,W,cin,cout", [(1, 1, 16, 8, 64, 64), (2, 3, 20, 12, 64, 64), (1, 3, 45, 45, 13, 64),
                                              (2, 5, 2, 2, 128, 128), (1, 1, 1, 1, 64, 64), (2, 2, 5, 5, 64, 64)])
def test_ragged_shapes_plan(lib, G, B, H, W, cin, cout):
    rc, p = plan(lib, G, B, H, W, cin, cout)
    assert rc == 0, lib.fabric_b200_last_error()
    assert p.smem_bytes <= SMEM and p.grid >= 1 and p.total_units >= 1
    bh = 16 if H > 8 else 8 if H > 4 else 4 if H > 2 else 2
    m_tiles = G * ((B + 16 // bh - 1) // (16 // bh)) * ((H + bh - 1) // bh) * ((W + 7) // 8)
    assert p.ctas == (2 if m_tiles % 2 == 0 else 1)                    # an odd number of pixel tiles cannot be paired
    assert p.total_units == m_tiles // p.ctas * (cout // p.n_tile)


def test_bad_requests_are_refused(lib):
    assert plan(lib, 1, 1, 16, 16, 48, 64)[0] != 0                      # Cin not 16 / k*64
    assert plan(lib, 1, 1, 16, 16, 64, 96)[0] != 0                      # Cout not a multiple of 64
    assert plan(lib, 1, 2, 16, 16, 64, 64, ctas=2)[0] == 0
    assert plan(lib, 1, 1, 16, 8, 64, 64, ctas=2)[0] != 0               # a single pixel tile cannot form a pair
    assert plan(lib, 1, 2, 16, 16, 64, 256, epi_warps=8)[0] != 0        # 8 epilogue warps only up to 128-wide tiles
    assert plan(lib, 1, 2, 16, 16, 64, 64, shift_in_acc=1)[0] != 0      # needs the shift vector


# ---------------------------------------------------------------------------------------------- weight gradient
def wgrad_plan(lib, G, B, H, W, ca, cb, wide=3, splits=0):
    from fabric_b200 import _lib
    d = _lib.WgradDesc()
    d.G, d.B, d.H, d.W, d.Ca, d.Cb = G, B, H, W, ca, (16 if cb <= 16 else cb)
    d.p = d.q = d.ws = 0x1000
    d.splits, d.wide = splits, wide
    out = _lib.WgradPlan()
    rc = lib.fabric_b200_conv3x3_wgrad_plan(C.byref(d), SMS, SMEM, C.byref(out))
    return rc, out


@pytest.mark.parametrize("name,G,H,cin,cout", LAYERS)
def test_wgrad_grid_fills_whole_waves(lib, name, G, H, cin, cout):
    """The weight-gradient grid is not persistent: the planner must not leave a mostly empty last wave (profiles/
    r01_ncu_full_wgrad.md), and it picks the second form exactly where it measured faster."""
    rc, p = wgrad_plan(lib, G, 64, H, H, cout, cin)          # P = dL/dz has the layer's OUTPUT channels
    assert rc == 0, lib.fabric_b200_last_error()
    assert p.smem_bytes <= SMEM and p.stages >= 2
    # second form where it measured faster; halo-P form for 64 dL/dz channels from 16384 pixel tiles up (the 256 x 256 layers)
    want = 2 if (cout >= 128 and 64 <= cin <= 128) else (3 if (cout == 64 and p.tiles_total >= 16384) else 1)
    assert p.form == want, (p.form, want, p.tiles_total)
    if p.form == 3:
        assert p.items == max(1, (16 if cin <= 16 else cin) // 64)      # ONE item per Q chunk covers all three filter rows
    assert p.grid == p.items * p.splits and p.splits * 4 <= p.tiles_total
    waves = -(-p.grid // SMS)
    assert p.grid / (waves * SMS) >= 0.85, (p.grid, waves)    # every wave at least 85 % full
    # no other split count (within the planner's search range) does better by more than rounding
    best = min(-(-p.items * s // SMS) / s for s in range(1, min(4 * SMS // p.items + 1, p.tiles_total // 4) + 1))
    assert waves / p.splits <= best * 1.001


def test_wgrad_forms_and_fallbacks(lib):
    assert wgrad_plan(lib, 2, 64, 128, 128, 128, 128, wide=1)[1].form == 1
    assert wgrad_plan(lib, 2, 64, 128, 128, 128, 128, wide=2)[1].form == 2
    assert wgrad_plan(lib, 2, 64, 256, 256, 64, 13, wide=2)[1].form == 3      # 13-band stem: never the second form (first form, halo-P at this size)
    assert wgrad_plan(lib, 2, 5, 4, 4, 128, 128, wide=2)[1].form == 1         # maps of 8 rows or fewer
    assert wgrad_plan(lib, 1, 1, 16, 16, 96, 64)[0] != 0                      # Ca must be a multiple of 64
    # halo-P: forced by wide = 4 on small maps, never for Ca != 64 or maps of 8 rows or fewer, not with the three-MMA form
    assert wgrad_plan(lib, 1, 2, 32, 24, 64, 64, wide=4)[1].form == 3 and wgrad_plan(lib, 1, 2, 32, 24, 64, 64, wide=1)[1].form == 1
    assert wgrad_plan(lib, 1, 2, 32, 24, 128, 64, wide=4)[1].form == 1 and wgrad_plan(lib, 1, 8, 8, 8, 64, 64, wide=4)[1].form == 1
    assert wgrad_plan(lib, 2, 64, 256, 256, 64, 13, wide=0)[1].form == 1
    # the operand-swapped call of ops.conv3x3_wgrad for up4.c1 (p = the 128-channel input, q = 64-channel dL/dz): second form
    assert wgrad_plan(lib, 1, 64, 256, 256, 128, 64)[1].form == 2
