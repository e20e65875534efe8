"""Pin the oracle to the reference and write the golden fixtures.

Runs ONLY in the build container, where the unmodified reference is mounted at
/root/reference.  It (1) imports the reference's own modules
(models/bidate_model.py, utils/metrics.py), (2) asserts that every function in
``oracle/bidatenet_oracle.py`` reproduces them bit-for-bit / to fp32 round-off on
seeded inputs, and (3) writes what the reference produced to ``tests/golden/`` so
that the CPU and GPU test-suites (which cannot see /root/reference) can re-check
the oracle and the CUDA path against the reference's numbers.

    python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("FABRIC_REFERENCE", "/root/reference")

from oracle import bidatenet_oracle as O  # noqa: E402


def load_reference():
    """the reference's own modules, loaded by file path under private names (oracle/ref_loader.py): this repo's
    ``models/`` pickle shim is a regular package and would shadow the reference's namespace package on sys.path"""
    from oracle import ref_loader
    BiDateNet, ref_metrics, _, root = ref_loader.load()
    return BiDateNet, ref_metrics


def load_reference_host_side():
    """utils/dataloaders.py and utils/inference.py of the reference, loaded by file path with their unavailable
    imports stubbed (rasterio, cv2, utils.helpers: IO / plotting only) and the removed sklearn
    ``image.extract_patches`` mapped to its surviving private twin ``_extract_patches`` (same function, renamed in
    scikit-learn 0.24).  Only the pure-numpy functions are used: ``onera_siamese_loader``, ``_get_patches``,
    ``_get_bands``."""
    import importlib.util
    import types
    from sklearn.feature_extraction import image as sk_image
    if not hasattr(sk_image, "extract_patches"):
        sk_image.extract_patches = sk_image._extract_patches
    saved = {k: sys.modules.get(k) for k in ("rasterio", "cv2", "utils", "utils.dataloaders", "utils.helpers")}
    try:
        for name in ("rasterio", "cv2"):
            if saved[name] is None:
                try:
                    __import__(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        pkg = types.ModuleType("utils")
        pkg.__path__ = []
        sys.modules["utils"] = pkg
        helpers = types.ModuleType("utils.helpers")
        helpers.log_figure = helpers.scale = None
        sys.modules["utils.helpers"] = helpers

        def load(name, rel):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            return mod
        dl = load("utils.dataloaders", "utils/dataloaders.py")
        inf = load("utils.inference", "utils/inference.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return dl, inf


def check_host_side():
    """oracle tiler / reassembler / augmentation == the reference's numpy code (inference.py:134-236,
    dataloaders.py:148-165), bit-exact, including ragged scene sizes and every (rot, flip, flip) combination."""
    import random
    dl, inf = load_reference_host_side()
    rng = np.random.RandomState(7)
    for (h, w, p) in ((64, 64, 32), (70, 97, 32), (96, 64, 32), (33, 47, 16), (48, 48, 16)):
        bands = rng.randn(h, w, 13).astype(np.float32)
        ref = inf._get_patches(bands, p)
        ora = O.get_patches(bands, p)
        assert np.array_equal(ref[0], ora[0]) and tuple(ref[1:]) == tuple(ora[1:]), (h, w, p)
        masks = rng.randint(0, 2, size=(ref[0].shape[0], p, p)).astype(np.float64)
        a = inf._get_bands(masks, *ref[1:], patch_size=p)
        b = O.get_bands(masks, *ora[1:], patch_size=p)
        assert np.array_equal(a, b), (h, w, p)
    # augmentation: replay the loader's three random draws (randint(0,3), random(), random()) from the same seed
    size = 12
    img = rng.randn(2, 13, 40, 40).astype(np.float32)
    lbl = rng.randint(0, 2, size=(40, 40)).astype(np.uint8)
    dataset = {"c": {"images": img, "labels": lbl}}
    seen = set()
    for seed in range(64):
        random.seed(seed)
        rot = random.randint(0, 3)
        f0 = random.random() > 0.5
        f1 = random.random() > 0.5
        random.seed(seed)
        d1, d2, lab = dl.onera_siamese_loader(dataset, "c", 5, 9, size, True)
        oi, ol = O.augment_patch(img[:, :, 5:5 + size, 9:9 + size], lbl[5:5 + size, 9:9 + size], rot, f0, f1)
        assert np.array_equal(d1, oi[0]) and np.array_equal(d2, oi[1]) and np.array_equal(lab, ol), seed
        # the gather form used by the device kernel
        src = img[0, 0, 5:5 + size, 9:9 + size]
        for (i, j) in ((0, 0), (3, 7), (size - 1, 2)):
            si, sj = O.augment_source_index(i, j, size, rot, f0, f1)
            assert d1[0, i, j] == src[si, sj], (seed, i, j)
        seen.add((rot, f0, f1))
    assert len(seen) == 16, "not every augmentation combination was exercised"
    d1, d2, lab = dl.onera_siamese_loader(dataset, "c", 5, 9, size, False)
    oi, ol = O.augment_patch(img[:, :, 5:5 + size, 9:9 + size], lbl[5:5 + size, 9:9 + size], 0, False, False)
    assert np.array_equal(d1, oi[0]) and np.array_equal(lab, ol)
    return len(seen)


def ref_model(BiDateNet, sd):
    m = BiDateNet(13, 2)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def build_golden():
    """Assert oracle == reference on every check and return what the REFERENCE produced (the golden dict)."""
    BiDateNet, RM = load_reference()

    # state_dict spec must equal the reference's, key for key
    ref_sd = BiDateNet(13, 2).state_dict()
    spec = O.state_dict_spec()
    assert [k for k, _, _ in spec] == list(ref_sd.keys())
    for k, shape, dt in spec:
        assert tuple(ref_sd[k].shape) == tuple(shape) and ref_sd[k].dtype == dt, k

    sd = O.make_state_dict(seed=0)
    golden = {}

    # ---- config 1: B=2, 13x32x32 -------------------------------------------------------
    x1, x2, labels = O.make_inputs(2, 32, seed=1)
    model = ref_model(BiDateNet, sd)

    model.eval()
    with torch.no_grad():
        ref_eval = model(x1, x2)
    ora_eval = O.bidatenet_forward(x1, x2, sd, training=False)
    assert torch.equal(ref_eval, ora_eval), (ref_eval - ora_eval).abs().max()
    golden.update(c1_x1=x1, c1_x2=x2, c1_labels=labels, c1_logits_eval=ref_eval)

    # losses on the eval logits, 3-D and 4-D labels (dims quirk)
    crit_ref = {
        "dice": RM.dice_loss, "jaccard": RM.jaccard_loss,
        "tversky": RM.TverskyLoss(alpha=0.1, beta=0.9),
        "focal": RM.FocalLoss(gamma=2.0),
    }
    crit_ora = {
        "dice": O.dice_loss, "jaccard": O.jaccard_loss,
        "tversky": lambda l, t: O.tversky_loss(l, t, 0.1, 0.9),
        "focal": lambda l, t: O.focal_loss(l, t, 2.0),
    }
    import warnings
    warnings.simplefilter("ignore")
    for name in crit_ref:
        for nd, lab in (("3d", labels), ("4d", labels[:, None])):
            if name == "focal" and nd == "4d":
                lab_ = lab
            else:
                lab_ = lab
            lr = ref_eval.clone().requires_grad_(True)
            lo = ref_eval.clone().requires_grad_(True)
            a = crit_ref[name](lr, lab_)
            b = crit_ora[name](lo, lab_)
            a.backward()
            b.backward()
            assert torch.equal(a.detach(), b.detach()), (name, nd, a, b)
            assert torch.equal(lr.grad, lo.grad), (name, nd)
            golden[f"c1_loss_{name}_{nd}"] = a.detach()
            golden[f"c1_dlogits_{name}_{nd}"] = lr.grad.clone()

    # training step: fwd (batch stats) + tversky + bwd, BN running-stat updates
    model = ref_model(BiDateNet, sd)
    model.train()
    logits_t = model(x1, x2)
    loss = RM.TverskyLoss(alpha=0.1, beta=0.9)(logits_t, labels)
    loss.backward()
    ref_grads = {k: p.grad for k, p in model.named_parameters()}
    ref_new = model.state_dict()
    l_o, logits_o, grads_o, new_o = O.train_step(x1, x2, labels, sd, lambda l, t: O.tversky_loss(l, t, 0.1, 0.9))
    assert torch.allclose(logits_t.detach(), logits_o, rtol=0, atol=2e-5), (logits_t.detach() - logits_o).abs().max()
    assert abs(float(loss) - float(l_o)) < 1e-6
    for k, g in ref_grads.items():
        d = (g - grads_o[k]).norm() / (g.norm() + 1e-12)
        # conv biases feeding a train-mode BN have a true gradient of 0: only round-off remains
        if k.endswith(".0.bias") or k.endswith(".3.bias"):
            assert g.abs().max() < 1e-5 and grads_o[k].abs().max() < 1e-5, k
        else:
            assert d < 2e-4, (k, float(d))
    for k, v in new_o.items():
        assert torch.allclose(ref_new[k].float(), v.float(), rtol=1e-5, atol=1e-6), k
    nbt = {k: int(v) for k, v in ref_new.items() if k.endswith("num_batches_tracked")}
    assert nbt["inc.conv.conv.1.num_batches_tracked"] == 2 and nbt["up1.conv.conv.1.num_batches_tracked"] == 1
    golden.update(c1_logits_train=logits_t.detach(), c1_loss_train=loss.detach())
    for k, g in ref_grads.items():
        golden["c1_gradnorm/" + k] = g.norm()
        golden["c1_gradhead/" + k] = g.flatten()[:64].clone()
        if g.dim() == 1:
            golden["c1_grad/" + k] = g.clone()
    golden["c1_grad/outc.conv.weight"] = ref_grads["outc.conv.weight"].clone()
    for k, v in ref_new.items():
        if "running_" in k or "num_batches" in k:
            golden["c1_newstat/" + k] = v.clone()

    # ---- patch 90 (reference default, metadata.json:32): F.pad branch in `up` ----------
    x1p, x2p, _ = O.make_inputs(1, 90, seed=2)
    model = ref_model(BiDateNet, sd)
    model.eval()
    with torch.no_grad():
        ref90 = model(x1p, x2p)
    assert torch.equal(ref90, O.bidatenet_forward(x1p, x2p, sd))
    golden.update(p90_x1=x1p, p90_x2=x2p, p90_logits_eval=ref90)

    # ---- one 256x256 pair (the benchmark patch size), eval ------------------------------
    x1b, x2b, _ = O.make_inputs(1, 256, seed=3)
    with torch.no_grad():
        ref256 = model(x1b, x2b)
    assert torch.equal(ref256, O.bidatenet_forward(x1b, x2b, sd))
    # inputs are regenerated from the seed (2 x 3.4 MB); keep the logits + an input checksum
    golden.update(p256_logits_eval=ref256, p256_x1_sum=x1b.double().sum(), p256_x2_sum=x2b.double().sum())

    # weights checksum so a different RNG stream on another machine is detected
    for k in ("inc.conv.conv.0.weight", "down4.mpconv.1.conv.3.weight", "up1.conv.conv.0.weight", "outc.conv.weight"):
        golden["sd_sum/" + k] = sd[k].double().sum()
        golden["sd_abssum/" + k] = sd[k].double().abs().sum()

    return golden


def main():
    torch.set_num_threads(os.cpu_count())
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    golden = build_golden()
    n_aug = check_host_side()
    path = os.path.join(out_dir, "bidatenet_golden.pt")
    if os.path.exists(path):
        old = torch.load(path)
        same = set(old) == set(golden) and all(torch.equal(torch.as_tensor(old[k]), torch.as_tensor(golden[k])) for k in golden)
        print(f"committed fixture {'is reproduced bit-for-bit' if same else 'DIFFERS from this run'} ({len(old)} tensors)")
    torch.save(golden, path)
    sz = os.path.getsize(path)
    print(f"oracle == reference on all checks (network, losses, training step, tiler, {n_aug} augmentations); "
          f"wrote {len(golden)} tensors, {sz/1e6:.2f} MB")


if __name__ == "__main__":
    main()
