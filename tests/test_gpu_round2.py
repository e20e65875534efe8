"""GPU tests added in round 2 (run with -m gpu on a B200):

* the training step at the BENCHMARK shape (64 pairs of 13x256x256): determinism, finiteness, exact linearity of the
  backward in the upstream gradient, date-swap symmetry, teacher-forced gradients of one 256x256 pair vs fp32 autograd;
* training at the reference's default geometry (patch 90, batch 32: metadata.json:32-33,40 -- the F.pad branch of `up`);
* a short SGD trajectory against the fp32 oracle;
* the data-parallel step: gradients written straight into the all-reduce bucket, the fused update kernel (SGD + running
  statistics + packed-weight refresh) bit-equal to torch.optim.SGD + the pack kernels, 2-rank NCCL equivalence;
* per-block train-mode entry points (double_conv / down / up / outconv .forward with gradients) vs torch;
* cache invalidation after writes that do not bump tensor versions; scene row-band sharding; device guard.
"""
import os
import socket

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from fabric_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _model(cuda, seed=0, train=True):
    from fabric_b200 import BiDateNet
    from oracle import bidatenet_oracle as O
    m = BiDateNet(13, 2)
    m.load_state_dict(O.make_state_dict(seed=seed))
    m = m.to(cuda)
    return m.train() if train else m.eval()


def _grads(model):
    return {k: p.grad.detach().clone() for k, p in model.named_parameters()}


def _step(model, x1, x2, labels, scale=1.0):
    from fabric_b200.metrics import TverskyLoss
    for p in model.parameters():
        p.grad = None
    loss = TverskyLoss(alpha=0.1, beta=0.9)(model(x1, x2), labels)
    (loss * scale).backward()
    return loss.detach().clone()


# ------------------------------------------------------------------------------------------------ benchmark shape
def test_train_step_full_batch_properties(cuda):
    """BASELINE configs[2] shape (64 x 13x256x256), properties that need no oracle at this size:
    finite, deterministic (two runs bit-equal), backward exactly linear in the upstream gradient (x2 is exact in bf16 and
    fp32), and date-swap symmetry (the network is symmetric in its two inputs: shared encoder, commutative fusion)."""
    g = torch.Generator(device=cuda).manual_seed(11)
    x1 = torch.randn(64, 13, 256, 256, device=cuda, generator=g)
    x2 = torch.randn(64, 13, 256, 256, device=cuda, generator=g)
    labels = (torch.rand(64, 256, 256, device=cuda, generator=g) < 0.1).long()
    runs = []
    for scale, swap in ((1.0, False), (1.0, False), (2.0, False), (1.0, True)):
        model = _model(cuda)
        loss = _step(model, x2 if swap else x1, x1 if swap else x2, labels, scale)
        runs.append((loss, _grads(model), {k: v.clone() for k, v in model.state_dict().items() if "running" in k}))
        del model
    (l0, g0, s0), (l1, g1, s1), (l2, g2, _), (l3, g3, _) = runs
    assert torch.isfinite(l0) and all(torch.isfinite(v).all() for v in g0.values())
    assert 0.0 < float(l0) < 1.0
    assert torch.equal(l0, l1)
    for k in g0:
        assert torch.equal(g0[k], g1[k]), f"not deterministic: {k}"
        assert torch.equal(g2[k], 2 * g0[k]), f"backward not linear in the upstream gradient: {k}"
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k
    # swapped dates: same logits / loss; gradients equal up to the fp32 summation order of the two date groups
    assert abs(float(l3) - float(l0)) <= 1e-6
    for k in g0:
        if k.endswith(".0.bias") or k.endswith(".3.bias"):
            continue
        assert rel(g3[k], g0[k]) <= 2e-3, (k, rel(g3[k], g0[k]))
    # conv biases in front of a train-mode BatchNorm: exactly zero gradient; every other gradient is non-trivial
    for k, v in g0.items():
        if k.endswith(".0.bias") or k.endswith(".3.bias"):
            assert float(v.abs().max()) == 0.0, k
        else:
            assert float(v.abs().max()) > 0.0, k


@pytest.mark.parametrize("batch,size,seed", [(1, 256, 7), (3, 90, 11)])
def test_backward_teacher_forced_at_benchmark_patch_and_reference_default_patch(cuda, batch, size, seed):
    """fp32 torch autograd of the oracle graph evaluated AT the tensors the CUDA forward stored reproduces the CUDA
    gradients at the benchmark patch size (256, one pair) and at the reference's default patch (90: odd sizes, the F.pad
    branch of `up`, unet_parts.py:68-72).  Tolerance: rel-L2 <= 3e-2 per parameter tensor (bf16 activation gradients)."""
    from fabric_b200 import autograd
    from fabric_b200.metrics import TverskyLoss
    from oracle import bidatenet_oracle as O
    from oracle import bidatenet_oracle_bf16 as Q
    from tests.test_gpu_train import _saved_as_nchw
    sd = O.make_state_dict(seed=0)
    model = _model(cuda)
    x1, x2, labels = O.make_inputs(batch, size, seed=seed)
    autograd.KEEP_SAVED = True
    try:
        logits = model(x1.to(cuda), x2.to(cuda))
        saved = _saved_as_nchw(autograd.LAST_SAVED)
    finally:
        autograd.KEEP_SAVED, autograd.LAST_SAVED = False, None
    loss = TverskyLoss(alpha=0.1, beta=0.9)(logits, labels.to(cuda))
    loss.backward()
    for g in (0, 1):
        saved[f"inc.x.{g}"] = saved[f"inc.x.{g}"][:, :13].contiguous()
    loss_f, logits_f, grads_f = Q.train_step_forced(x1, x2, labels, sd, lambda l, t: O.tversky_loss(l, t, 0.1, 0.9), saved)
    assert rel(logits.detach().cpu(), logits_f) <= 1e-4
    assert abs(loss.item() - float(loss_f)) <= 1e-5
    worst = (0.0, "")
    for k, p in model.named_parameters():
        if k.endswith(".0.bias") or k.endswith(".3.bias"):
            continue
        worst = max(worst, (rel(p.grad.detach().cpu(), grads_f[k]), k))
    assert worst[0] <= 3e-2, worst


def test_training_step_reference_default_geometry(cuda):
    """patch 90, batch 32 (metadata.json:32-33,40; odd sizes: the F.pad branch of `up`): one full step through the
    reference-facing API, deterministic, against (a) the free-running fp32 oracle -- loss, train-mode logits, BatchNorm
    running statistics, conv-gradient norms and directions -- and (b) the bf16 precision-model oracle (same algorithm,
    rounded where the kernels store bf16), which the CUDA path must follow much more tightly."""
    from fabric_b200.metrics import TverskyLoss
    from oracle import bidatenet_oracle as O
    from oracle import bidatenet_oracle_bf16 as Q
    sd = O.make_state_dict(seed=0)
    x1, x2, labels = O.make_inputs(32, 90, seed=21)
    torch.set_num_threads(os.cpu_count())
    crit_o = lambda l, t: O.tversky_loss(l, t, 0.1, 0.9)   # noqa: E731
    l_o, logits_o, grads_o, new_o = O.train_step(x1, x2, labels, sd, crit_o)
    l_q, logits_q, grads_q = Q.train_step(x1, x2, labels, sd, crit_o)
    outs = []
    for _ in range(2):
        model = _model(cuda)
        logits = model(x1.to(cuda), x2.to(cuda))
        loss = TverskyLoss(alpha=0.1, beta=0.9)(logits, labels.to(cuda))
        loss.backward()
        outs.append((logits.detach().clone(), loss.detach().clone(), _grads(model), model.state_dict()))
    logits, loss, grads, st = outs[0]
    assert torch.equal(logits, outs[1][0]) and torch.equal(loss, outs[1][1])
    assert all(torch.equal(grads[k], outs[1][2][k]) for k in grads)
    m = dict(logits_vs_fp32=rel(logits.cpu(), logits_o), logits_vs_model=rel(logits.cpu(), logits_q),
             model_vs_fp32=rel(logits_q, logits_o), loss=float(loss), loss_fp32=float(l_o), loss_model=float(l_q))
    worst32, worstq, worst_cos = (0.0, ""), (0.0, ""), (1.0, "")
    for k, g in grads.items():
        if g.dim() == 4:
            worst32 = max(worst32, (rel(g.cpu(), grads_o[k]), k))
            worstq = max(worstq, (rel(g.cpu(), grads_q[k]), k))
            a, b = g.cpu().double().flatten(), grads_o[k].double().flatten()
            worst_cos = min(worst_cos, (float(a @ b / (a.norm() * b.norm())), k))
    m.update(worst_wgrad_vs_fp32=worst32, worst_wgrad_vs_model=worstq, worst_cos_vs_fp32=worst_cos,
             model_wgrad_vs_fp32=max((rel(grads_q[k], grads_o[k]), k) for k in grads if grads[k].dim() == 4))
    print("default-geometry step:", m)
    assert m["logits_vs_fp32"] <= 6e-2 and abs(float(loss) - float(l_o)) <= 3e-3, m
    # the CUDA path sits as close to the fp32 reference as the precision model itself does (factor 1.5)
    assert m["logits_vs_fp32"] <= 1.5 * m["model_vs_fp32"] + 5e-3, m
    assert worst32[0] <= 1.5 * m["model_wgrad_vs_fp32"][0] + 2e-2, m
    assert worst_cos[0] >= 0.8, m          # (the precision model itself is 0.53 rel-L2 off fp32 on the stem's gradient here)
    for k, v in new_o.items():
        if "running" in k:
            assert rel(st[k].cpu().float(), v.float()) <= 2e-2, k
        if "num_batches" in k:
            assert int(st[k]) == int(v), k


def test_sgd_trajectory_follows_fp32_oracle(cuda):
    """12 SGD steps (lr 0.05, Tversky) on a fixed batch of 8 pairs 13x64x64: the loss curve of the CUDA path stays within
    1.5e-2 of the fp32 oracle's at every step and both decrease -- a systematic bias in the bf16 BatchNorm-backward or
    weight-gradient path would show up as a diverging curve."""
    from fabric_b200.distributed import DataParallelStep
    from fabric_b200.metrics import TverskyLoss
    from oracle import bidatenet_oracle as O
    torch.set_num_threads(os.cpu_count())
    lr, steps = 0.05, 12
    sd = O.make_state_dict(seed=0)
    x1, x2, labels = O.make_inputs(8, 64, seed=31)
    crit_o = lambda l, t: O.tversky_loss(l, t, 0.1, 0.9)   # noqa: E731
    cur = {k: v.clone() for k, v in sd.items()}
    curve_o = []
    for _ in range(steps):
        l_o, _, grads_o, new_o = O.train_step(x1, x2, labels, cur, crit_o)
        curve_o.append(float(l_o))
        for k, g in grads_o.items():
            cur[k] = cur[k] - lr * g
        cur.update(new_o)
    model = _model(cuda)
    dp = DataParallelStep(model)
    crit = TverskyLoss(alpha=0.1, beta=0.9)
    a, b, lab = x1.to(cuda), x2.to(cuda), labels.to(cuda)
    curve = []
    for _ in range(steps):
        dp.zero_grad()
        loss = crit(model(a, b), lab)
        loss.backward()
        dp.sync_and_step(lr)
        curve.append(float(loss.detach()))
    print("oracle", [round(v, 4) for v in curve_o])
    print("cuda  ", [round(v, 4) for v in curve])
    assert curve_o[-1] < curve_o[0] - 0.02 and curve[-1] < curve[0] - 0.02
    assert max(abs(u - v) for u, v in zip(curve, curve_o)) <= 1.5e-2, list(zip(curve, curve_o))


# ------------------------------------------------------------------------------------------------ data-parallel step
def test_gradients_land_in_the_bucket_and_fused_update_matches_torch(cuda):
    """DataParallelStep: p.grad is a view of the flat bucket and the backward kernels write there (no copies); the ONE
    update kernel == torch.optim.SGD + pack_conv_weight for both packed layouts, bit for bit; a second step then uses the
    refreshed packed copies (no pack launches) and equals a model that repacks from scratch."""
    from fabric_b200 import ops
    from fabric_b200.distributed import DataParallelStep
    from oracle import bidatenet_oracle as O
    x1, x2, labels = O.make_inputs(2, 32, seed=3)
    x1, x2, labels = x1.to(cuda), x2.to(cuda), labels.to(cuda)
    ref = _model(cuda)                                   # plain autograd Function + torch SGD
    opt = torch.optim.SGD(ref.parameters(), lr=0.5)
    l_ref = _step(ref, x1, x2, labels)
    g_ref = _grads(ref)
    opt.step()
    model = _model(cuda)
    dp = DataParallelStep(model)
    for p in model.parameters():
        assert p.grad.data_ptr() >= dp.bucket.data_ptr() and p.grad.data_ptr() < dp.bucket.data_ptr() + dp.bucket.numel() * 4
    dp.zero_grad()
    from fabric_b200.metrics import TverskyLoss
    n0 = ops.LAUNCHES
    loss = TverskyLoss(alpha=0.1, beta=0.9)(model(x1, x2), labels)
    loss.backward()
    assert torch.equal(loss.detach(), l_ref)
    for k, p in model.named_parameters():
        assert p.grad is dp.sink[p] and torch.equal(p.grad, g_ref[k]), k
    dp.sync_and_step(0.5)
    launches_step1 = ops.LAUNCHES - n0
    for (k, p), q in zip(model.named_parameters(), ref.parameters()):
        assert torch.equal(p.detach(), q.detach()), k
    for k, v in ref.state_dict().items():
        assert torch.equal(model.state_dict()[k], v), k
    # packed copies maintained by the update kernel == the pack kernels run on the new weights
    from fabric_b200.unet_parts import double_conv
    for m in model.modules():
        if isinstance(m, double_conv):
            for idx in (0, 3):
                w = m.conv[idx].weight
                assert torch.equal(m._packed(idx, training=True), ops.pack_conv_weight(w, 0)), idx
                assert torch.equal(m._packed_dgrad(idx), ops.pack_conv_weight(w, 1)), idx
    # second step: no pack kernels are launched and the result equals the reference model's second step
    l_ref2 = _step(ref, x1, x2, labels)
    n1 = ops.LAUNCHES
    loss2 = TverskyLoss(alpha=0.1, beta=0.9)(model(x1, x2), labels)
    loss2.backward()
    dp.sync_and_step(0.5)
    assert torch.equal(loss2.detach(), l_ref2)
    assert ops.LAUNCHES - n1 == launches_step1
    g_ref2 = _grads(ref)
    for k, p in model.named_parameters():
        assert torch.equal(p.grad, g_ref2[k]), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out, exact=False, steps=2, loss_kind="tversky", size=32):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fabric_b200 import BiDateNet
    from fabric_b200.distributed import DataParallelStep
    from fabric_b200.metrics import TverskyLoss
    from oracle import bidatenet_oracle as O
    model = BiDateNet(13, 2)
    model.load_state_dict(O.make_state_dict(seed=rank))        # different weights per rank on purpose
    model = model.to(dev).train()
    dp = DataParallelStep(model, exact=exact)
    dp.broadcast_parameters(0)
    x1, x2, labels = O.make_inputs(2 * world, size, seed=5)    # the global batch; rank r takes pairs [2r, 2r+2)
    sl = slice(2 * rank, 2 * rank + 2)
    from fabric_b200 import metrics, ops
    crit = {"tversky": TverskyLoss(alpha=0.1, beta=0.9), "focal": metrics.FocalLoss(2.0), "dice": metrics.dice_loss}[loss_kind]
    losses = []
    for _ in range(steps):
        dp.zero_grad()
        loss = crit(model(x1[sl].to(dev), x2[sl].to(dev)), labels[sl].to(dev))
        loss.backward()
        dp.sync_and_step(0.25)
        losses.append(float(loss.detach()))
    torch.cuda.synchronize()
    res = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    res["_losses"] = torch.tensor(losses)
    res["_collectives"] = torch.tensor(ops.EXACT.collectives if exact else 0)
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_data_parallel_step_nccl_two_ranks_equals_manual_average(cuda):
    """Two NCCL ranks x 2 pairs, two steps of DataParallelStep (segmented async all-reduce + fused update) == one process
    that runs both shards through the same kernels and applies p -= lr * mean_r(g_r) and mean_r(running stats)."""
    import torch.multiprocessing as mp
    from fabric_b200.metrics import TverskyLoss
    from oracle import bidatenet_oracle as O
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = [out[r] for r in range(world)]
    for k in got[0]:
        if not k.startswith("_"):
            assert torch.equal(got[0][k], got[1][k]), f"replicas diverged: {k}"
    # manual emulation on one GPU
    x1, x2, labels = O.make_inputs(2 * world, 32, seed=5)
    crit = TverskyLoss(alpha=0.1, beta=0.9)
    sd = O.make_state_dict(seed=0)
    for _ in range(2):
        grads, stats = [], []
        for r in range(world):
            m = _model(cuda)
            m.load_state_dict(sd)
            m.train()
            sl = slice(2 * r, 2 * r + 2)
            crit(m(x1[sl].to(cuda), x2[sl].to(cuda)), labels[sl].to(cuda)).backward()
            grads.append({k: p.grad.detach().cpu() for k, p in m.named_parameters()})
            stats.append({k: v.detach().cpu() for k, v in m.state_dict().items()})
        new = {}
        for k, v in sd.items():
            if k in grads[0]:
                new[k] = v - 0.25 * 0.5 * (grads[0][k] + grads[1][k])
            elif "running" in k:
                new[k] = 0.5 * (stats[0][k] + stats[1][k])
            else:
                new[k] = stats[0][k]
        sd = new
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            assert torch.allclose(got[0][k], v, rtol=1e-5, atol=1e-7), k
        else:
            assert torch.equal(got[0][k], v), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("loss_kind", ["tversky", "focal"])
def test_exact_global_mode_equals_single_rank_on_the_whole_batch(cuda, loss_kind):
    """SURVEY 8e equivalence: N ranks x B pairs with DataParallelStep(exact=True) (SyncBN: batch statistics and their
    backward sums all-reduced; loss on the global batch: train.py:91-92) == 1 rank x N*B pairs.  Same loss to fp32 round-off;
    parameters after one step equal up to the bf16 re-rounding that a different fp32 summation order of the statistics
    causes (rel-L2 <= 2e-2 per tensor, printed)."""
    import torch.multiprocessing as mp
    from fabric_b200 import metrics
    from fabric_b200.distributed import DataParallelStep
    from oracle import bidatenet_oracle as O
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    size = 64
    mp.spawn(_nccl_worker, args=(world, _free_port(), out, True, 1, loss_kind, size), nprocs=world, join=True)
    got = [out[r] for r in range(world)]
    for k in got[0]:
        if not k.startswith("_"):
            assert torch.equal(got[0][k], got[1][k]), f"replicas diverged: {k}"
    assert int(got[0]["_collectives"]) >= 18 + 18 + 1          # 18 BatchNorms forward + backward, the loss
    # one rank, the whole batch of 4 pairs, same weights
    x1, x2, labels = O.make_inputs(2 * world, size, seed=5)
    model = _model(cuda)
    dp = DataParallelStep(model)
    crit = {"tversky": metrics.TverskyLoss(alpha=0.1, beta=0.9), "focal": metrics.FocalLoss(2.0)}[loss_kind]
    loss = crit(model(x1.to(cuda), x2.to(cuda)), labels.to(cuda))
    loss.backward()
    dp.sync_and_step(0.25)
    dp.close()
    print("exact-global losses:", float(got[0]["_losses"][0]), float(got[1]["_losses"][0]), "single rank:", float(loss.detach()))
    assert abs(float(got[0]["_losses"][0]) - float(loss.detach())) <= 2e-5 and got[0]["_losses"][0] == got[1]["_losses"][0]
    want = model.state_dict()
    sd0 = O.make_state_dict(seed=0)
    worst = {"conv": (0.0, ""), "bn": (0.0, ""), "head": (0.0, "")}
    for k, v in want.items():
        if not v.dtype.is_floating_point:
            assert torch.equal(got[0][k], v.cpu()), k
            continue
        if "running" in k:
            assert rel(got[0][k], v.cpu()) <= 1e-4, (k, rel(got[0][k], v.cpu()))
            continue
        # compare the UPDATE (p_new - p_old), not the parameter: the step is small against the weights
        d_got, d_want = got[0][k] - sd0[k], v.cpu() - sd0[k]
        if float(d_want.abs().max()) == 0.0:
            assert float(d_got.abs().max()) == 0.0, k
            continue
        kind = "head" if k.startswith("outc") else ("conv" if v.dim() == 4 else "bn")
        worst[kind] = max(worst[kind], (rel(d_got, d_want), k))
    print("exact-global worst update mismatch:", worst)
    # The two runs differ only in the fp32 summation order of the BatchNorm sums (all-reduce of per-rank partials vs one
    # rank's partials): ~1e-7 on scale / shift, which re-rounds a few bf16 activations and flips a few ReLU masks per layer;
    # through 18 BatchNorms on a 4-pair batch that noise is amplified like any other perturbation of this network at random
    # init (rounding only the INPUTS to bf16 moves the fp32 reference's gradients by 21-25 %).  Measured at
    # 32x32: head 7e-4, conv weights 7e-2, BatchNorm gamma / beta 8e-2; loss 7e-6.  A wrong collective (a missing 1/world,
    # statistics of one rank only, a local loss) shows as O(1) in the updates and as 1e-2 in the loss.
    assert worst["head"][0] <= 2e-2 and worst["conv"][0] <= 0.12 and worst["bn"][0] <= 0.15, worst


# ------------------------------------------------------------------------------------------------ per-block entry points
def _ref_double_conv(cin, cout, cuda):
    m = torch.nn.Sequential(torch.nn.Conv2d(cin, cout, 3, padding=1), torch.nn.BatchNorm2d(cout), torch.nn.ReLU(),
                            torch.nn.Conv2d(cout, cout, 3, padding=1), torch.nn.BatchNorm2d(cout), torch.nn.ReLU()).to(cuda)
    with torch.no_grad():
        for i in (1, 4):
            m[i].weight.uniform_(0.5, 1.5)
            m[i].bias.normal_(0, 0.2)
    return m


def _copy_dc(dst, src):
    with torch.no_grad():
        for i in (0, 1, 3, 4):
            dst.conv[i].load_state_dict(src[i].state_dict())


@pytest.mark.parametrize("block", ["double_conv", "down", "up", "inconv"])
def test_block_train_mode_forward_backward_matches_torch(cuda, block):
    """`double_conv` / `inconv` / `down` / `up` `.forward` in .train() mode (reference unet_parts.py:21-23,31-33,44-46,64-80):
    output (batch statistics), running-stat updates and gradients for inputs and parameters.  Yardsticks: the torch
    modules in fp32 (output, running statistics, gradient direction) and the same torch graph with bf16 rounding at the
    kernels' storage points (oracle/bidatenet_oracle_bf16.py: `double_conv`), which the gradients must follow closely."""
    from fabric_b200 import unet_parts as P
    from oracle import bidatenet_oracle_bf16 as Q
    torch.manual_seed(4)
    B, H, W = 3, 24, 40
    glue = lambda *a: a[0]                                             # noqa: E731
    if block == "double_conv":
        mine, cin, cout = P.double_conv(64, 128), 64, 128
        dc = mine
        args = (torch.randn(B, cin, H, W, device=cuda),)
    elif block == "inconv":
        mine, cin, cout = P.inconv(13, 64), 13, 64
        dc = mine.conv
        args = (torch.randn(B, cin, H, W, device=cuda),)
    elif block == "down":
        mine, cin, cout = P.down(64, 128), 64, 128
        dc = mine.mpconv[1]
        args = (torch.randn(B, cin, 2 * H, 2 * W, device=cuda),)
        glue = lambda x: F.max_pool2d(x, 2)                            # noqa: E731
    else:
        mine, cin, cout = P.up(128, 64), 128, 64
        dc = mine.conv
        args = (torch.randn(B, 64, H // 2, W // 2 - 1, device=cuda), torch.randn(B, 64, H, W, device=cuda).relu())

        def glue(x1, x2):
            x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
            dy, dx = x2.size(2) - x1.size(2), x2.size(3) - x1.size(3)
            x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
            return torch.cat([x2, x1], dim=1)
    ref = _ref_double_conv(cin, cout, cuda)
    _copy_dc(dc, ref)
    mine = mine.to(cuda).train()
    ref.train()
    need_dx = block != "inconv"
    args = tuple(a.bfloat16().float().requires_grad_(need_dx) for a in args)        # bf16-representable inputs
    rargs = tuple(a.detach().clone().requires_grad_(need_dx) for a in args)
    qargs = tuple(a.detach().clone().requires_grad_(need_dx) for a in args)
    y = mine(*args)
    yr = ref(glue(*rargs))
    Pq = {f"p.{i}.{n}": t.detach().clone().requires_grad_(True) for i in (0, 1, 3, 4) for n, t in ref[i].named_parameters()}
    yq = Q.double_conv(Q.r(glue(*qargs)), Pq, "p", training=True)
    gy = torch.randn_like(yr).bfloat16().float()
    y.backward(gy)
    yr.backward(gy)
    yq.backward(gy)
    m = dict(y_vs_fp32=rel(y, yr), y_vs_model=rel(y, yq))
    for i in (0, 3):
        m[f"w{i}_vs_fp32"] = rel(dc.conv[i].weight.grad, ref[i].weight.grad)
        m[f"w{i}_vs_model"] = rel(dc.conv[i].weight.grad, Pq[f"p.{i}.weight"].grad)
        m[f"model_w{i}_vs_fp32"] = rel(Pq[f"p.{i}.weight"].grad, ref[i].weight.grad)
        for n in ("weight", "bias"):
            m[f"bn{i + 1}.{n}_vs_model"] = rel(getattr(dc.conv[i + 1], n).grad, Pq[f"p.{i + 1}.{n}"].grad)
            m[f"bn{i + 1}.{n}_vs_fp32"] = rel(getattr(dc.conv[i + 1], n).grad, getattr(ref[i + 1], n).grad)
        m[f"rm{i + 1}"] = rel(dc.conv[i + 1].running_mean, ref[i + 1].running_mean)
        m[f"rv{i + 1}"] = rel(dc.conv[i + 1].running_var, ref[i + 1].running_var)
    if need_dx:
        for j, (a, r_, q) in enumerate(zip(args, rargs, qargs)):
            m[f"dx{j}_vs_fp32"] = rel(a.grad, r_.grad)
            m[f"dx{j}_vs_model"] = rel(a.grad, q.grad)
    print(block, {k: round(v, 5) for k, v in m.items()})
    assert y.shape == yr.shape and m["y_vs_fp32"] <= 1e-2 and m["y_vs_model"] <= 1e-2, m
    for i in (0, 3):
        assert m[f"rm{i + 1}"] <= 1e-2 and m[f"rv{i + 1}"] <= 1e-2, m
        # gradients: as close to fp32 as the precision model is (x1.5), and close to the precision model itself
        assert m[f"w{i}_vs_fp32"] <= 1.5 * m[f"model_w{i}_vs_fp32"] + 1e-2, m
        assert m[f"w{i}_vs_model"] <= 5e-2, m
        assert m[f"bn{i + 1}.weight_vs_model"] <= 8e-2 and m[f"bn{i + 1}.bias_vs_model"] <= 8e-2, m
    if need_dx:
        assert all(m[f"dx{j}_vs_model"] <= 5e-2 for j in range(len(args))), m


def test_outconv_train_mode_has_gradients(cuda):
    from fabric_b200 import unet_parts as P
    torch.manual_seed(5)
    mine = P.outconv(64, 2).to(cuda).train()
    ref = torch.nn.Conv2d(64, 2, 1).to(cuda)
    ref.load_state_dict(mine.conv.state_dict())
    x = torch.randn(2, 64, 20, 12, device=cuda).bfloat16().float().requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = mine(x), ref(xr)
    assert rel(y, yr) <= 1e-5
    gy = torch.randn_like(yr)
    y.backward(gy)
    yr.backward(gy)
    assert rel(mine.conv.weight.grad, ref.weight.grad) <= 1e-4 and rel(mine.conv.bias.grad, ref.bias.grad) <= 1e-5
    assert rel(x.grad, xr.grad) <= 5e-3        # stored as bf16


# ------------------------------------------------------------------------------------------------ caches / pickles / guard
def test_caches_follow_writes_that_do_not_bump_versions(cuda):
    """`.data` writes keep `_version`; the packed-weight / folded-BN caches must not serve stale weights after
    load_state_dict, a train()/eval() round trip, `.to()`, or an explicit invalidate_caches()."""
    from oracle import bidatenet_oracle as O
    model = _model(cuda, train=False)
    x1, x2, _ = O.make_inputs(2, 32, seed=9)
    x1, x2 = x1.to(cuda), x2.to(cuda)
    with torch.no_grad():
        y0 = model(x1, x2).clone()
        w = model.down2.mpconv[1].conv[0].weight
        v0 = w._version
        w.data.mul_(1.5)                                  # no version bump
        assert w._version == v0
        model.invalidate_caches()
        y1 = model(x1, x2).clone()
        assert not torch.equal(y0, y1)
        model.up1.conv.conv[4].running_var.data.mul_(2.0)
        model.train(); model.eval()                        # mode round trip drops the caches too
        y2 = model(x1, x2).clone()
        assert not torch.equal(y1, y2)
        model.load_state_dict(O.make_state_dict(seed=0))   # same tensors, new values, in place
        assert torch.equal(model(x1, x2), y0)
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ref = O.bidatenet_forward(x1.cpu(), x2.cpu(), sd, training=False)
    assert rel(y0.cpu(), ref) <= 1e-2


def test_reference_style_pickle_runs(cuda, tmp_path):
    """a whole-model pickle carrying only the REFERENCE's attributes (train.py:222) loads through the models/ shim and runs
    in eval and train mode (the extra attributes of this repo's classes are derived in __setstate__)"""
    from fabric_b200 import BiDateNet
    from fabric_b200.unet_parts import double_conv
    from oracle import bidatenet_oracle as O
    m = BiDateNet(13, 2)
    m.load_state_dict(O.make_state_dict(seed=0))
    for mod in m.modules():                                # strip everything the reference's classes do not have
        for name in ("in_ch", "out_ch", "tune1", "tune2", "fold_bn", "fuse_head", "fuse_product"):
            mod.__dict__.pop(name, None)
    path = tmp_path / "ref_style.pt"
    torch.save(m, path)
    m2 = torch.load(path, weights_only=False).to(cuda).eval()
    assert all(hasattr(d, "in_ch") for d in m2.modules() if isinstance(d, double_conv))
    x1, x2, labels = O.make_inputs(2, 32, seed=1)
    with torch.no_grad():
        y = m2(x1.to(cuda), x2.to(cuda))
    ref = O.bidatenet_forward(x1, x2, O.make_state_dict(seed=0), training=False)
    assert rel(y.cpu(), ref) <= 1e-2
    m2.train()
    from fabric_b200.metrics import dice_loss
    dice_loss(m2(x1.to(cuda), x2.to(cuda)), labels.to(cuda)).backward()
    assert all(p.grad is not None for p in m2.parameters())


def test_dataparallel_replica_is_rejected_with_a_pointer(cuda):
    """nn.DataParallel replicas have no parameters of their own (reference helpers.py:335 wraps the model): the forward must
    say what to use instead, not fail deep inside autograd"""
    model = _model(cuda)
    replica = model._replicate_for_data_parallel()
    with pytest.raises(RuntimeError, match="DataParallelStep"):
        replica(torch.zeros(1, 13, 32, 32, device=cuda), torch.zeros(1, 13, 32, 32, device=cuda))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tensor_on_another_device_is_rejected(cuda):
    from fabric_b200 import _lib, ops
    x = torch.zeros(1, 13, 32, 32, device="cuda:1")
    with pytest.raises(_lib.FabricB200Error, match="current device"):
        ops.pack_input(x)
    with torch.cuda.device(1):
        assert ops.pack_input(x).device.index == 1


# ------------------------------------------------------------------------------------------------ scene bands
def test_scene_row_bands_concatenate_to_the_single_rank_mask(cuda):
    """config 5 sharding: ranks own disjoint row bands; each band computed from ONLY its scene rows; the concatenation of
    the bands equals the single-rank mask, which equals the reference pipeline's (existing test)."""
    from fabric_b200.scene import ScenePlan, predict_scene, predict_scene_band
    model = _model(cuda, train=False)
    g = torch.Generator(device=cuda).manual_seed(3)
    H, W, p = 300, 200, 64
    d1 = torch.randn(13, H, W, device=cuda, generator=g)
    d2 = torch.randn(13, H, W, device=cuda, generator=g)
    full, info = predict_scene(model, d1, d2, patch_size=p, batch_size=7)
    assert info["tiles"] == 4 * 3 + 4 + 3 + 1
    for world in (2, 3, 5):
        bands = []
        for r in range(world):
            plan = ScenePlan(H, W, p, r, world)
            bands.append(predict_scene_band(model, d1[:, plan.row0:plan.row1].contiguous(), d2[:, plan.row0:plan.row1].contiguous(),
                                            plan, batch_size=5))
        assert torch.equal(torch.cat(bands, 0), full), world


def test_step_update_kernel_modes(cuda):
    """fabric_b200_train_step_update: mode 0 (SGD), mode 1 (scale), mode 2 (SGD + packed refresh incl. the 13 -> 16 padded stem),
    mode 3 (the same per 16 x 128-channel tile through shared memory: what DataParallelStep emits)"""
    import struct
    from fabric_b200 import _lib, ops
    torch.manual_seed(6)
    w = torch.randn(64, 13, 3, 3, device=cuda)
    gw = torch.randn_like(w)
    b = torch.randn(100000, device=cuda)
    gb = torch.randn_like(b)
    st = torch.randn(777, device=cuda)
    wf, wd = ops.pack_conv_weight(w, 0), ops.pack_conv_weight(w, 1)
    w0, b0, st0 = w.clone(), b.clone(), st.clone()
    recs, n = bytearray(), 0

    def emit(p, g, cnt, mode, f=0, d=0, cout=0, cin=0, cp=0):
        nonlocal recs, n
        off = 0
        while off < cnt:
            k = min(65536, cnt - off)
            recs += struct.pack("<QQQQiiiiiiii", p + 4 * off, g + 4 * off, f, d, k, mode, off, cout, cin, cp, 0, 0)
            off += k
            n += 1
    emit(w.data_ptr(), gw.data_ptr(), w.numel(), 2, wf.data_ptr(), wd.data_ptr(), 64, 13, 16)
    emit(b.data_ptr(), gb.data_ptr(), b.numel(), 0)
    emit(st.data_ptr(), st.data_ptr(), st.numel(), 1)
    # mode 3 on a weight with ragged tiles in both directions (40 = 16 + 16 + 8 output, 200 = 128 + 72 input channels)
    w3 = torch.randn(40, 200, 3, 3, device=cuda)
    g3 = torch.randn_like(w3)
    w30 = w3.clone()
    cp3 = ops.cpad(200)
    wf3 = torch.zeros(40, 9, cp3, dtype=torch.bfloat16, device=cuda)
    wd3 = torch.zeros(200, 9, 40, dtype=torch.bfloat16, device=cuda)
    for co0 in range(0, 40, 16):
        for ci0 in range(0, 200, 128):
            recs += struct.pack("<QQQQiiiiiiii", w3.data_ptr(), g3.data_ptr(), wf3.data_ptr(), wd3.data_ptr(), min(16, 40 - co0), 3,
                                co0, 40, 200, cp3, ci0, min(128, 200 - ci0))
            n += 1
    table = torch.frombuffer(recs, dtype=torch.uint8).clone().to(cuda)
    _lib.check(_lib.load().fabric_b200_train_step_update(table.data_ptr(), n, 0.3, 0.5, 0.25, torch.cuda.current_stream().cuda_stream))
    lr = torch.tensor(0.3, dtype=torch.float32) * torch.tensor(0.5, dtype=torch.float32)
    assert torch.allclose(w, w0 - float(lr) * gw, rtol=0, atol=1e-6) and torch.allclose(b, b0 - float(lr) * gb, rtol=0, atol=1e-6)
    assert torch.equal(st, st0 * 0.25)
    assert torch.equal(wf, ops.pack_conv_weight(w, 0)) and torch.equal(wd, ops.pack_conv_weight(w, 1))
    assert torch.allclose(w3, w30 - float(lr) * g3, rtol=0, atol=1e-6)
    assert torch.equal(wf3, ops.pack_conv_weight(w3, 0)) and torch.equal(wd3, ops.pack_conv_weight(w3, 1))


# ------------------------------------------------------------------------------------------------ CUDA graph step
@pytest.mark.parametrize("batch,size", [(2, 32), (4, 90)])
def test_graphed_train_step_equals_eager_steps(cuda, batch, size):
    """fabric_b200.graph.GraphedTrainStep: forward + loss + backward + fused update captured once, replayed per step.
    Three replays on three different batches leave exactly the parameters, running statistics and losses that three eager
    steps on the same batches leave (bit for bit: same kernels, same launch configurations)."""
    from fabric_b200.distributed import DataParallelStep
    from fabric_b200.graph import GraphedTrainStep
    from fabric_b200.metrics import TverskyLoss
    from oracle import bidatenet_oracle as O
    crit = TverskyLoss(alpha=0.1, beta=0.9)
    batches = [tuple(t.to(cuda) for t in O.make_inputs(batch, size, seed=40 + i)) for i in range(3)]
    eager = _model(cuda)
    dp_e = DataParallelStep(eager)
    losses_e = []
    for b in batches:
        dp_e.zero_grad()
        loss = crit(eager(b[0], b[1]), b[2])
        loss.backward()
        dp_e.sync_and_step(0.1)
        losses_e.append(float(loss.detach()))
    dp_e.close()
    graphed = _model(cuda)
    dp_g = DataParallelStep(graphed)
    step = GraphedTrainStep(graphed, crit, dp_g, 0.1, batches[0])
    # capture (and its warm-up steps) must leave the model where it was
    for (k, v), w in zip(graphed.state_dict().items(), _model(cuda).state_dict().values()):
        assert torch.equal(v, w), k
    losses_g = [float(step(*b)) for b in batches]
    assert losses_g == losses_e, (losses_g, losses_e)
    for (k, v), w in zip(graphed.state_dict().items(), eager.state_dict().values()):
        assert torch.equal(v, w), k
    # and the model still evaluates with the updated weights (version-keyed caches were invalidated)
    graphed.eval(); eager.eval()
    with torch.no_grad():
        assert torch.equal(graphed(batches[0][0], batches[0][1]), eager(batches[0][0], batches[0][1]))
    dp_g.close()


# ------------------------------------------------------------------------------------------------ fused BatchNorm backward
@pytest.mark.parametrize("G,B,H,W,cin,cout", [(2, 3, 20, 12, 128, 64), (1, 2, 32, 24, 64, 128), (2, 2, 16, 16, 256, 256),
                                              (1, 5, 7, 9, 64, 64)])
def test_dgrad_epilogue_bn_backward_reduce_matches_standalone_kernels(cuda, G, B, H, W, cin, cout):
    """fabric_b200_conv3x3 with bnbwd_z: the epilogue's masked output equals relu'(bn(z)) * (plain conv output) bit for bit,
    its partials sum to (sum dy, sum dy*xhat), and fabric_b200_bn_bwd_from_partials gives the dz / dgamma / dbeta of the
    stand-alone reduce + apply kernels (fp32 summation order aside) -- all three tile widths (64 / 128 in the register-
    accumulating instantiation, 256 through the shuffle tree) and a ragged map."""
    from fabric_b200 import ops
    torch.manual_seed(12)
    x5 = torch.randn(G, B, H, W, cin, device=cuda).bfloat16()             # plays dL/dz2
    wq = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device=cuda) / (3 * cin ** 0.5), 0)
    z5 = torch.randn(G, B, H, W, cout, device=cuda).bfloat16()            # pre-activation of the BatchNorm being differentiated
    bn = torch.nn.BatchNorm2d(cout).to(cuda)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_(0, 0.3)
    zf = z5.float()
    mean = zf.mean((1, 2, 3))
    var = zf.var((1, 2, 3), unbiased=False)
    invstd = torch.rsqrt(var + 1e-5)
    scale = bn.weight[None] * invstd
    shift = bn.bias[None] - mean * scale
    base = torch.stack([scale, shift, mean, invstd]).contiguous()
    coef = ops.BnCoef(base[i] for i in range(4))
    coef.base = base
    plain = ops.conv3x3(x5, wq, cout)["y"]
    fused = ops.conv3x3(x5, wq, cout, bnbwd=(z5, coef))
    mask = (zf * scale[:, None, None, None] + shift[:, None, None, None]) > 0
    want_dy = torch.where(mask, plain.float(), torch.zeros((), device=cuda)).bfloat16()
    assert torch.equal(fused["y"], want_dy)
    part = fused["stats"]                                                 # [grid, 2, n_tile, 2], CTA i holds N tile i % ntiles
    grid, _, n_tile, _ = part.shape
    nt = cout // n_tile
    sums = part.view(grid // nt, nt, 2, n_tile, 2).double().sum(0).permute(1, 0, 2, 3).reshape(2, cout, 2)[:G]
    dyf = want_dy.double()
    xhat = (zf.double() - mean.double()[:, None, None, None]) * invstd.double()[:, None, None, None]
    # (the epilogue accumulates the raw second sum, sum dy * z; the finalize kernel maps it to sum dy * xhat)
    assert rel(sums[..., 0], dyf.sum((1, 2, 3))) <= 1e-4 and rel(sums[..., 1], (dyf * zf.double()).sum((1, 2, 3))) <= 1e-4
    del xhat
    dz_f, dg_f, db_f = ops.bn_bwd_from_partials(z5, fused["y"], part, coef, bn.weight)
    dz_s, dg_s, db_s = ops.bn_relu_bwd(z5, None, plain, False, None, scale, shift, mean, invstd, bn.weight)
    assert rel(dg_f, dg_s) <= 1e-4 and rel(db_f, db_s) <= 1e-4
    assert rel(dz_f.float(), dz_s.float()) <= 2e-3                        # bf16 re-rounding of a few elements
    assert (dz_f != dz_s).float().mean().item() <= 0.02


def test_training_step_with_and_without_fused_bn_backward_reduce_agree(cuda):
    """whole step (config 1 and an odd-size batch): FUSE_BN_BWD_REDUCE on / off give the same loss and gradients up to the
    fp32 summation order of the (sum dy, sum dy*xhat) pairs"""
    from fabric_b200 import autograd
    from oracle import bidatenet_oracle as O
    for (b, s_, seed) in ((2, 32, 1), (3, 90, 2)):
        x1, x2, labels = (t.to(cuda) for t in O.make_inputs(b, s_, seed=seed))
        res = []
        for fuse in (True, False):
            autograd.FUSE_BN_BWD_REDUCE = fuse
            try:
                model = _model(cuda)
                loss = _step(model, x1, x2, labels)
                res.append((loss, _grads(model)))
            finally:
                autograd.FUSE_BN_BWD_REDUCE = True
        assert torch.equal(res[0][0], res[1][0])
        worst = max((rel(res[0][1][k], res[1][1][k]), k) for k in res[0][1] if float(res[1][1][k].abs().max()) > 0)
        print("fused vs stand-alone BN backward reduce:", (b, s_), worst)
        assert worst[0] <= 2e-2, worst          # (worst: a cancellation-dominated BatchNorm beta gradient, 1.05e-2 measured)


@pytest.mark.parametrize("H,W,C,gp", [(32, 32, 64, True), (45, 45, 128, True), (6, 6, 256, True), (16, 16, 512, False)])
def test_bn_backward_recomputing_the_activation_matches_reading_it(cuda, H, W, C, gp):
    """product-fused encoder levels: BatchNorm-2's backward with a == NULL (activation recomputed from z exactly as
    bn_apply stored it) == the same kernels reading the stored activation, bit for bit -- incl. odd sizes, where the last
    row / column has no pooling window, and the arg-max routing of the max pool."""
    from fabric_b200 import ops
    torch.manual_seed(21)
    G, B = 2, 3
    z5 = torch.randn(G, B, H, W, C, device=cuda).bfloat16()
    bn = torch.nn.BatchNorm2d(C).to(cuda)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_(0, 0.3)
    zf = z5.float()
    mean, var = zf.mean((1, 2, 3)), zf.var((1, 2, 3), unbiased=False)
    invstd = torch.rsqrt(var + 1e-5)
    scale = (bn.weight[None] * invstd).contiguous()
    shift = (bn.bias[None] - mean * scale).contiguous()
    cat = torch.zeros(1, B, H, W, C + 64, device=cuda, dtype=torch.bfloat16)
    a5, pooled = ops.bn_apply_relu(z5, scale, shift, pool=True, prod_out=cat)
    a_none, pooled2 = ops.bn_apply_relu(z5, scale, shift, pool=True, prod_out=torch.zeros_like(cat), write_a=False)
    assert a_none is None and torch.equal(pooled, pooled2)
    ga = torch.randn(1, B, H, W, C + 64, device=cuda).bfloat16()
    gpt = torch.randn(G, B, H // 2, W // 2, C, device=cuda).bfloat16() if gp else None
    args = (ga, True, gpt, scale, shift, mean.contiguous(), invstd.contiguous(), bn.weight)
    dz_a, dg_a, db_a = ops.bn_relu_bwd(z5, a5, *args)
    dz_r, dg_r, db_r = ops.bn_relu_bwd(z5, None, *args)
    # dy is the same bit for bit; the two variants reduce it over a different number of blocks, so the fp32 sums -- and a
    # handful of bf16 roundings of dz -- may differ
    assert rel(dg_a, dg_r) <= 1e-5 and rel(db_a, db_r) <= 1e-5
    assert rel(dz_a.float(), dz_r.float()) <= 1e-4 and (dz_a != dz_r).float().mean().item() <= 1e-3


def test_training_step_with_and_without_stored_encoder_activations_agree(cuda):
    from fabric_b200 import autograd
    from oracle import bidatenet_oracle as O
    for (b, s_, seed) in ((2, 32, 1), (2, 90, 2)):
        x1, x2, labels = (t.to(cuda) for t in O.make_inputs(b, s_, seed=seed))
        res = []
        for rec in (True, False):
            autograd.RECOMPUTE_ENCODER_ACT = rec
            try:
                model = _model(cuda)
                loss = _step(model, x1, x2, labels)
                res.append((loss, _grads(model)))
            finally:
                autograd.RECOMPUTE_ENCODER_ACT = True
        assert torch.equal(res[0][0], res[1][0])
        for k in res[0][1]:
            # the same dz bit for bit; BatchNorm sums are reduced over a different block count (fp32 order)
            assert rel(res[0][1][k], res[1][1][k]) <= 1e-2 or float(res[1][1][k].abs().max()) == 0.0, k


@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (2, 45, 45), (1, 7, 5), (6, 256, 256)])   # (the last: unrolled main loops)
def test_fused_bn_head_forward_and_backward_match_the_separate_kernels(cuda, B, H, W):
    """bn_apply_relu_head == bn_apply_relu + outconv (activation bit-equal, logits to fp32 summation order); bn_head_bwd ==
    outconv_bwd + bn_relu_bwd up to the bf16 rounding of the du tensor the fused path never materialises; incl. pixel
    counts that are not a multiple of the 4 pixels a warp handles."""
    from fabric_b200 import ops
    torch.manual_seed(31)
    C = 64
    z5 = torch.randn(1, B, H, W, C, device=cuda).bfloat16()
    bn = torch.nn.BatchNorm2d(C).to(cuda)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_(0, 0.3)
    zf = z5.float()
    mean, var = zf.mean((1, 2, 3)), zf.var((1, 2, 3), unbiased=False)
    invstd = torch.rsqrt(var + 1e-5)
    scale = (bn.weight[None] * invstd).contiguous()
    shift = (bn.bias[None] - mean * scale).contiguous()
    coef = (scale, shift, mean.contiguous(), invstd.contiguous())
    hw = (torch.randn(2, C, 1, 1, device=cuda) * 0.2).contiguous()
    hb = torch.randn(2, device=cuda)
    a_f, logits_f = ops.bn_apply_relu_head(z5, scale, shift, hw, hb)
    a_s, _ = ops.bn_apply_relu(z5, scale, shift)
    logits_s = ops.outconv(a_s, hw, hb)
    assert torch.equal(a_f, a_s)
    assert rel(logits_f, logits_s) <= 1e-6
    dlogits = torch.randn(B, 2, H, W, device=cuda)
    dz_f, dg_f, db_f, dw_f, dhb_f = ops.bn_head_bwd(dlogits, z5, coef, bn.weight, hw)
    du, dw_s, dhb_s = ops.outconv_bwd(dlogits, a_s, hw)
    dz_s, dg_s, db_s = ops.bn_relu_bwd(z5, None, du, False, None, *coef, bn.weight)
    assert rel(dw_f, dw_s) <= 1e-5 and rel(dhb_f, dhb_s) <= 1e-5
    assert rel(dg_f, dg_s) <= 5e-3 and rel(db_f, db_s) <= 5e-3       # du rounded to bf16 on the separate path only
    assert rel(dz_f.float(), dz_s.float()) <= 6e-3
    # against fp64 torch autograd of the same graph
    zt = zf.double().requires_grad_(True)
    g_, b_ = bn.weight.double().detach().requires_grad_(True), bn.bias.double().detach().requires_grad_(True)
    hwt = hw.double().reshape(2, C).detach().requires_grad_(True)
    m_, v_ = zt.mean((1, 2, 3)), zt.var((1, 2, 3), unbiased=False)
    at = torch.relu((zt - m_) * torch.rsqrt(v_ + 1e-5) * g_ + b_)
    lt = torch.einsum("gbhwc,kc->bkhw", at, hwt)
    lt.backward(dlogits.double())
    assert rel(dz_f.float(), zt.grad) <= 6e-3 and rel(dg_f, g_.grad) <= 2e-3 and rel(db_f, b_.grad) <= 2e-3
    assert rel(dw_f.reshape(2, C), hwt.grad) <= 5e-3


def test_training_step_with_and_without_fused_head_agree(cuda):
    from fabric_b200 import autograd
    from oracle import bidatenet_oracle as O
    x1, x2, labels = (t.to(cuda) for t in O.make_inputs(2, 48, seed=4))
    res = []
    for fuse in (True, False):
        autograd.FUSE_HEAD = fuse
        try:
            model = _model(cuda)
            loss = _step(model, x1, x2, labels)
            res.append((loss, _grads(model)))
        finally:
            autograd.FUSE_HEAD = True
    assert abs(float(res[0][0]) - float(res[1][0])) <= 1e-6
    worst = max((rel(res[0][1][k], res[1][1][k]), k) for k in res[0][1] if float(res[1][1][k].abs().max()) > 0)
    print("fused vs separate head:", worst)
    assert worst[0] <= 3e-2, worst
