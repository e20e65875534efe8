"""CPU, world_size 2, gloo: the data-parallel step logic (one flat all-reduce averaging gradients + BN running
statistics, parameter broadcast) -- the N > 1 path of bench.py / SURVEY.md 8e."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fabric_b200.distributed import DataParallelStep
    torch.manual_seed(100 + rank)                      # different initial weights per rank on purpose
    model = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1), nn.BatchNorm2d(4), nn.ReLU(), nn.Conv2d(4, 2, 1))
    dp = DataParallelStep(model)
    dp.broadcast_parameters(0)
    w0 = [p.detach().clone() for p in model.parameters()]
    torch.manual_seed(7 + rank)                        # per-rank shard of the batch
    x = torch.randn(4, 3, 8, 8)
    model.train()
    model(x).square().mean().backward()
    local = [p.grad.clone() for p in model.parameters()]
    local_rm = model[1].running_mean.clone()
    dp.sync()
    gathered = [torch.zeros_like(torch.cat([g.flatten() for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.flatten() for g in local]))
    rms = [torch.zeros_like(local_rm) for _ in range(world)]
    dist.all_gather(rms, local_rm)
    ok = torch.allclose(torch.cat([p.grad.flatten() for p in model.parameters()]), sum(gathered) / world, atol=1e-7)
    ok = ok and torch.allclose(model[1].running_mean, sum(rms) / world, atol=1e-7)
    wall = [torch.zeros_like(torch.cat([w.flatten() for w in w0])) for _ in range(world)]
    dist.all_gather(wall, torch.cat([w.flatten() for w in w0]))
    ok = ok and all(torch.equal(wall[0], w) for w in wall)          # broadcast made the replicas identical
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_data_parallel_step_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_scene_plan_bands_partition_rows_and_reproduce_reference_reassembly():
    """ScenePlan (config 5 row-band sharding): bands are disjoint and cover every row; composing each band from its own
    tiles in the reference's overwrite order and concatenating equals the reference `_get_bands` on the full tile list."""
    import numpy as np
    from fabric_b200.scene import ScenePlan, tile_origins
    from oracle import bidatenet_oracle as O
    rng = np.random.RandomState(0)
    for (h, w, p) in ((300, 200, 64), (70, 97, 32), (64, 64, 64), (129, 65, 32)):
        bands13 = rng.randn(h, w, 13).astype(np.float32)
        patches, hs, ws, lc, lr, _, _ = O.get_patches(bands13, p)
        origins = tile_origins(h, w, p)[0]
        tile_id = {o: i for i, o in enumerate(origins)}            # later duplicates (same origin) keep the LAST index
        masks = np.stack([np.full((p, p), float(i % 251)) + rng.rand(p, p) for i in range(len(origins))])
        want = O.get_bands(masks, hs, ws, lc, lr, h, w, patch_size=p)
        for world in (1, 2, 3, 8):
            rows, n = [], 0
            for r in range(world):
                plan = ScenePlan(h, w, p, r, world)
                canvas = np.zeros((plan.band_rows, w))
                first = 0
                for cls_i, (f, cnt) in enumerate(plan.classes):
                    for (y, x) in plan.origins_list[f:f + cnt]:
                        gy = y + plan.row0
                        # global tile index of this origin within its class (grid / last column / last row / corner)
                        base = [0, hs * ws, hs * ws + lc, hs * ws + lc + lr][cls_i]
                        k = {0: (gy // p) * ws + x // p, 1: gy // p, 2: x // p, 3: 0}[cls_i] + base
                        assert origins[k] == (gy, x)
                        canvas[y:y + p, x:x + p] = masks[k]
                    n += cnt
                rows.append(canvas)
            assert n == len(origins)
            assert np.array_equal(np.concatenate(rows, 0), want), (h, w, p, world)


def test_reference_pickle_loads_through_the_shim(tmp_path):
    """torch.save(model) as the reference does it (train.py:222), written with the REFERENCE's own classes registered
    under their real module paths, resolves through this repo's `models/` shim to fabric_b200 classes whose extra
    attributes are derived in __setstate__."""
    import sys
    import pytest
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference sources not reachable")
    RefNet, _, parts, _ = ref_loader.load()
    from oracle import bidatenet_oracle as O
    ref = RefNet(13, 2)
    ref.load_state_dict(O.make_state_dict(seed=0))
    net_mod = sys.modules[RefNet.__module__]
    saved = {k: sys.modules.get(k) for k in ("models", "models.bidate_model", "models.unet_parts")}
    classes = [RefNet] + [getattr(parts, n) for n in ("double_conv", "inconv", "down", "up", "outconv")]
    old_names = [c.__module__ for c in classes]
    try:                                   # pickle the reference model exactly as train.py:222 would name its classes
        import types
        pkg = types.ModuleType("models"); pkg.__path__ = []
        sys.modules.update({"models": pkg, "models.bidate_model": net_mod, "models.unet_parts": parts})
        RefNet.__module__ = "models.bidate_model"
        for c in classes[1:]:
            c.__module__ = "models.unet_parts"
        path = tmp_path / "reference_model.pt"
        torch.save(ref, path)
    finally:
        for c, n in zip(classes, old_names):
            c.__module__ = n
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    import models.bidate_model as shim                   # this repo's shim again
    loaded = torch.load(path, weights_only=False)
    from fabric_b200 import BiDateNet
    from fabric_b200.unet_parts import double_conv
    assert type(loaded) is BiDateNet is shim.BiDateNet
    dcs = [m for m in loaded.modules() if isinstance(m, double_conv)]
    assert len(dcs) == 9 and dcs[0].in_ch == 13 and dcs[0].out_ch == 64 and dcs[5].in_ch == 1024 and dcs[5].fold_bn
    assert loaded.fuse_head and loaded.fuse_product
    for k, v in ref.state_dict().items():
        assert torch.equal(loaded.state_dict()[k], v), k
