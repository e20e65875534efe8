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
    # the reference's package is called `models` / `utils`: import it under its own names
    # from its own directory (this repo's look-alike `models` package must not shadow it).
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")
             or k == "utils" or k.startswith("utils.")}
    sys.path.insert(0, REF)
    try:
        from models.bidate_model import BiDateNet            # noqa
        from utils import metrics as ref_metrics              # noqa
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return BiDateNet, ref_metrics


def ref_model(BiDateNet, sd):
    m = BiDateNet(13, 2)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def main():
    torch.set_num_threads(os.cpu_count())
    BiDateNet, RM = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

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

    torch.save(golden, os.path.join(out_dir, "bidatenet_golden.pt"))
    sz = os.path.getsize(os.path.join(out_dir, "bidatenet_golden.pt"))
    print(f"oracle == reference on all checks; wrote {len(golden)} tensors, {sz/1e6:.2f} MB")


if __name__ == "__main__":
    main()
