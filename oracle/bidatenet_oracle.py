"""CPU oracle for the BiDateNet hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a plain fp32 CPU restatement of the reference's algorithm for the
one path this repo accelerates (granularai/fabric, BiDateNet forward / loss /
backward).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; nothing under
``fabric_b200/`` does, and the product path raises if its CUDA library is missing.

Parity pin: the reference has no tests or golden vectors for this path
(SURVEY.md section 4 / 8c), so the pin is the reference itself -- its modules
are imported unmodified in the build container by ``oracle/make_golden.py``,
which (a) checks every function below against them on seeded inputs and (b)
writes the fixtures under ``tests/golden/``.  The arithmetic of the individual
ops lives in the reference's third-party dependency PyTorch (Dockerfile:1 pins
pytorch 1.0.1; the semantics of conv2d / batch_norm / max_pool2d / bilinear
``align_corners=True`` interpolate / softmax are unchanged in the torch 2.11 of
this image), so this restatement spells the network out over ``torch.nn.functional``
on explicit parameter dictionaries instead of ``nn.Module`` objects.

Every function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default, models/unet_parts.py:14,17
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default

# (prefix, in_ch, out_ch) of every double_conv, in registration order
# models/bidate_model.py:10-19
DOUBLE_CONVS: List[Tuple[str, int, int]] = [
    ("inc.conv.conv", 13, 64),
    ("down1.mpconv.1.conv", 64, 128),
    ("down2.mpconv.1.conv", 128, 256),
    ("down3.mpconv.1.conv", 256, 512),
    ("down4.mpconv.1.conv", 512, 512),
    ("up1.conv.conv", 1024, 256),
    ("up2.conv.conv", 512, 128),
    ("up3.conv.conv", 256, 64),
    ("up4.conv.conv", 128, 64),
]


def state_dict_spec(n_channels: int = 13, n_classes: int = 2):
    """Keys / shapes / dtypes of ``BiDateNet(n_channels, n_classes).state_dict()``
    in registration order (models/bidate_model.py:10-20, models/unet_parts.py:12-19,86)."""
    spec = []
    for prefix, cin, cout in DOUBLE_CONVS:
        if prefix.startswith("inc"):
            cin = n_channels
        for conv_i, bn_i, ci in ((0, 1, cin), (3, 4, cout)):
            spec.append((f"{prefix}.{conv_i}.weight", (cout, ci, 3, 3), torch.float32))
            spec.append((f"{prefix}.{conv_i}.bias", (cout,), torch.float32))
            spec.append((f"{prefix}.{bn_i}.weight", (cout,), torch.float32))
            spec.append((f"{prefix}.{bn_i}.bias", (cout,), torch.float32))
            spec.append((f"{prefix}.{bn_i}.running_mean", (cout,), torch.float32))
            spec.append((f"{prefix}.{bn_i}.running_var", (cout,), torch.float32))
            spec.append((f"{prefix}.{bn_i}.num_batches_tracked", (), torch.int64))
    spec.append(("outc.conv.weight", (n_classes, 64, 1, 1), torch.float32))
    spec.append(("outc.conv.bias", (n_classes,), torch.float32))
    return spec


def make_state_dict(seed: int = 0, n_channels: int = 13, n_classes: int = 2,
                    randomize_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights, independent of module construction order.

    Conv weights/biases follow the distribution of ``nn.Conv2d``'s default init
    (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both).
    With ``randomize_bn`` the BN affine parameters and running statistics are
    perturbed away from (1, 0, 0, 1) so that eval-mode parity exercises them.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, dtype in state_dict_spec(n_channels, n_classes):
        leaf = key.rsplit(".", 1)[1]
        is_bn = key.split(".")[-2] in ("1", "4") and not key.startswith("outc")
        if dtype == torch.int64:
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif not is_bn:
            if leaf == "weight":
                fan_in = shape[1] * shape[2] * shape[3]
                sd["_fan_in"] = fan_in
            bound = 1.0 / math.sqrt(sd["_fan_in"])
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            if not randomize_bn:
                sd[key] = torch.ones(shape) if leaf in ("weight", "running_var") else torch.zeros(shape)
            elif leaf == "weight":
                sd[key] = 0.75 + 0.5 * torch.rand(shape, generator=g)
            elif leaf == "bias":
                sd[key] = 0.2 * torch.randn(shape, generator=g)
            elif leaf == "running_mean":
                sd[key] = 0.1 * torch.randn(shape, generator=g)
            else:  # running_var
                sd[key] = 0.5 + torch.rand(shape, generator=g)
    sd.pop("_fan_in")
    return sd


def make_inputs(batch: int, size: int, seed: int = 1, n_channels: int = 13, p_change: float = 0.1):
    """Synthetic OSCD-shaped batch: z-scored bands ~ N(0,1) (utils/dataloaders.py:94-99),
    sparse binary change labels (utils/dataloaders.py:149-165)."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(batch, n_channels, size, size, generator=g)
    x2 = torch.randn(batch, n_channels, size, size, generator=g)
    labels = (torch.rand(batch, size, size, generator=g) < p_change).long()
    return x1, x2, labels


# ----------------------------------------------------------------------------------------
# network
# ----------------------------------------------------------------------------------------

def _bn(x, sd, prefix, training, new_stats):
    """nn.BatchNorm2d (models/unet_parts.py:14,17).  In training mode the batch
    statistics of THIS call normalise, and the running statistics move by
    momentum 0.1 towards (mean, unbiased var); num_batches_tracked += 1."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if not training:
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b,
                            False, BN_MOMENTUM, BN_EPS)
    rm = new_stats.setdefault(prefix + ".running_mean", sd[prefix + ".running_mean"].clone())
    rv = new_stats.setdefault(prefix + ".running_var", sd[prefix + ".running_var"].clone())
    nbt = new_stats.setdefault(prefix + ".num_batches_tracked", sd[prefix + ".num_batches_tracked"].clone())
    dims = (0, 2, 3)
    mean = x.mean(dims)
    var_b = x.var(dims, unbiased=False)
    n = x.numel() // x.shape[1]
    with torch.no_grad():
        rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
        rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var_b.detach() * (n / max(n - 1, 1)))
        nbt.add_(1)
    xh = (x - mean[None, :, None, None]) * torch.rsqrt(var_b + BN_EPS)[None, :, None, None]
    return xh * w[None, :, None, None] + b[None, :, None, None]


def double_conv(x, sd, prefix, training=False, new_stats=None):
    """(conv3x3 pad1 => BN => ReLU) * 2 -- models/unet_parts.py:8-23."""
    x = F.conv2d(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], padding=1)   # :13
    x = F.relu(_bn(x, sd, prefix + ".1", training, new_stats))                       # :14-15
    x = F.conv2d(x, sd[prefix + ".3.weight"], sd[prefix + ".3.bias"], padding=1)   # :16
    x = F.relu(_bn(x, sd, prefix + ".4", training, new_stats))                       # :17-18
    return x


def inconv(x, sd, training=False, new_stats=None):
    """models/unet_parts.py:26-33."""
    return double_conv(x, sd, "inc.conv.conv", training, new_stats)


def down(x, sd, name, training=False, new_stats=None):
    """MaxPool2d(2) then double_conv -- models/unet_parts.py:36-46."""
    return double_conv(F.max_pool2d(x, 2), sd, f"{name}.mpconv.1.conv", training, new_stats)


def up(x1, x2, sd, name, training=False, new_stats=None):
    """Bilinear x2 (align_corners=True), zero-pad to the skip size, cat([skip, up]),
    double_conv -- models/unet_parts.py:49-80."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)       # :56-58,65
    dy = x2.shape[2] - x1.shape[2]
    dx = x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))                      # :71-72
    x = torch.cat([x2, x1], dim=1)                                                      # :78
    return double_conv(x, sd, f"{name}.conv.conv", training, new_stats)


def outconv(x, sd):
    """1x1 conv head -- models/unet_parts.py:83-90."""
    return F.conv2d(x, sd["outc.conv.weight"], sd["outc.conv.bias"])


def encoder(x, sd, training=False, new_stats=None):
    """One date through the weight-shared encoder -- models/bidate_model.py:23-27 / :29-33."""
    x1 = inconv(x, sd, training, new_stats)
    x2 = down(x1, sd, "down1", training, new_stats)
    x3 = down(x2, sd, "down2", training, new_stats)
    x4 = down(x3, sd, "down3", training, new_stats)
    x5 = down(x4, sd, "down4", training, new_stats)
    return x1, x2, x3, x4, x5


def bidatenet_forward(x_d1, x_d2, sd, training=False, return_new_stats=False, return_intermediates=False):
    """BiDateNet.forward -- models/bidate_model.py:22-40.

    The encoder runs on date 1 and then on date 2 with the same weights; in training
    mode every encoder BatchNorm therefore normalises each date with that date's own
    batch statistics and updates its running statistics twice (date 1 first).
    The skips are relu(d2 * d1) at five scales (an elementwise PRODUCT, :35-38).
    """
    new_stats: Dict[str, torch.Tensor] = {}
    e1 = encoder(x_d1, sd, training, new_stats)
    e2 = encoder(x_d2, sd, training, new_stats)
    f = [torch.relu(b * a) for a, b in zip(e1, e2)]
    x = up(f[4], f[3], sd, "up1", training, new_stats)   # :35
    u1 = x
    x = up(x, f[2], sd, "up2", training, new_stats)      # :36
    u2 = x
    x = up(x, f[1], sd, "up3", training, new_stats)      # :37
    u3 = x
    x = up(x, f[0], sd, "up4", training, new_stats)      # :38
    u4 = x
    logits = outconv(x, sd)                              # :39
    out = (logits,)
    if return_new_stats:
        out = out + (new_stats,)
    if return_intermediates:
        out = out + ({"enc_d1": e1, "enc_d2": e2, "fuse": f, "up": (u1, u2, u3, u4)},)
    return out if len(out) > 1 else logits


# ----------------------------------------------------------------------------------------
# losses (utils/metrics.py); C >= 2 branches only -- the network has n_classes = 2
# ----------------------------------------------------------------------------------------

def _soft_sums(logits, true):
    """Shared front end of dice / jaccard / tversky -- utils/metrics.py:75-80, 110-115, 158-164.

    ``true`` may be [B,H,W] (what train.py:85,92 passes) or [B,1,H,W] (what the
    docstrings and notebooks/losses.ipynb use).  ``dims`` is derived from
    ``true.ndimension()`` (:79), so 3-D labels reduce over (batch, H) only and leave a
    [C, W] map, while 4-D labels reduce over (batch, H, W) and leave [C].
    """
    c = logits.shape[1]
    t = true.squeeze(1) if true.dim() == 4 else true
    one_hot = torch.eye(c)[t].permute(0, 3, 1, 2).float().type(logits.type())
    probas = F.softmax(logits, dim=1)
    dims = (0,) + tuple(range(2, true.dim()))
    return probas, one_hot, dims


def dice_loss(logits, true, eps=1e-7):
    """utils/metrics.py:51-83."""
    p, t, dims = _soft_sums(logits, true)
    inter = torch.sum(p * t, dims)
    card = torch.sum(p + t, dims)
    return 1 - (2.0 * inter / (card + eps)).mean()


def jaccard_loss(logits, true, eps=1e-7):
    """utils/metrics.py:86-119."""
    p, t, dims = _soft_sums(logits, true)
    inter = torch.sum(p * t, dims)
    card = torch.sum(p + t, dims)
    return 1 - (inter / (card - inter + eps)).mean()


def tversky_loss(logits, true, alpha=0.5, beta=0.5, eps=1e-7):
    """TverskyLoss.forward -- utils/metrics.py:130-171 (default loss, metadata.json:42-44)."""
    p, t, dims = _soft_sums(logits, true)
    inter = torch.sum(p * t, dims)
    fps = torch.sum(p * (1 - t), dims)
    fns = torch.sum((1 - p) * t, dims)
    return 1 - (inter / (inter + alpha * fps + beta * fns + eps)).mean()


def focal_loss(logits, true, gamma=0.0):
    """FocalLoss.forward with alpha=None, size_average=True -- utils/metrics.py:19-48.
    ``pt`` is detached (:35, ``Variable(logpt.data.exp())``)."""
    n, c = logits.shape[0], logits.shape[1]
    x = logits.view(n, c, -1).transpose(1, 2).contiguous().view(-1, c)   # :20-28
    tgt = true.reshape(-1, 1)
    logpt = F.log_softmax(x, dim=1).gather(1, tgt).view(-1)              # :32-34
    pt = logpt.detach().exp()                                            # :35
    return (-1 * (1 - pt) ** gamma * logpt).mean()                       # :43-46


def cross_entropy_loss(logits, true):
    """Deliberate extension (SURVEY.md 8a 'bce' row): the reference's ``bce`` choice
    (utils/helpers.py:303-304) raises on 2-channel logits, so the working stand-in is the
    2-class softmax cross entropy, which equals BCE on the logit difference."""
    t = true.squeeze(1) if true.dim() == 4 else true
    return F.cross_entropy(logits, t)


def get_criterion(loss_function: str, tversky_alpha=0.1, tversky_beta=0.9, focal_gamma=2.0):
    """utils/helpers.py:288-314."""
    if loss_function == "bce":
        return cross_entropy_loss
    if loss_function == "focal":
        return lambda l, t: focal_loss(l, t, focal_gamma)
    if loss_function == "dice":
        return dice_loss
    if loss_function == "jaccard":
        return jaccard_loss
    if loss_function == "tversky":
        return lambda l, t: tversky_loss(l, t, tversky_alpha, tversky_beta)
    raise ValueError(loss_function)


# ----------------------------------------------------------------------------------------
# one training step (train.py:88-95): returns loss, grads, updated BN statistics
# ----------------------------------------------------------------------------------------

def train_step(x_d1, x_d2, labels, sd, criterion):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype == torch.float32 and "running_" not in k}
    full = dict(sd)
    full.update(params)
    logits, new_stats = bidatenet_forward(x_d1, x_d2, full, training=True, return_new_stats=True)
    loss = criterion(logits, labels)
    grads = torch.autograd.grad(loss, list(params.values()))
    return loss.detach(), logits.detach(), dict(zip(params.keys(), grads)), new_stats


# ----------------------------------------------------------------------------------------
# scene tiler / reassembler (utils/inference.py:134-236), numpy
# ----------------------------------------------------------------------------------------

def _extract_patches(arr: np.ndarray, p: int) -> np.ndarray:
    """sklearn ``image.extract_patches(arr, (p, p, C), p)`` restated: non-overlapping
    p x p windows, floor(H/p) x floor(W/p) of them (utils/inference.py:152-154)."""
    h, w, c = arr.shape
    hs, ws = h // p, w // p
    v = arr[:hs * p, :ws * p].reshape(hs, p, ws, p, c).transpose(0, 2, 1, 3, 4)
    return v.reshape(hs, ws, 1, p, p, c)


def get_patches(bands: np.ndarray, patch_dim: int = 64):
    """``_get_patches`` -- utils/inference.py:134-181.  bands is [H, W, 13]."""
    patches = _extract_patches(bands, patch_dim)
    hs, ws = patches.shape[0], patches.shape[1]
    patches = patches.reshape(-1, patch_dim, patch_dim, bands.shape[2])
    last_row = bands[bands.shape[0] - patch_dim:, :, :]
    last_column = bands[:, bands.shape[1] - patch_dim:, :]
    corner = np.asarray([bands[bands.shape[0] - patch_dim:, bands.shape[1] - patch_dim:, :]])
    last_column = _extract_patches(last_column, patch_dim).reshape(-1, patch_dim, patch_dim, bands.shape[2])
    last_row = _extract_patches(last_row, patch_dim).reshape(-1, patch_dim, patch_dim, bands.shape[2])
    lc, lr = last_column.shape[0], last_row.shape[0]
    patches = np.vstack((patches, last_column, last_row, corner))
    return patches, hs, ws, lc, lr, bands.shape[0], bands.shape[1]


def get_bands(patches: np.ndarray, hs, ws, lc, lr, h, w, patch_size: int = 64) -> np.ndarray:
    """``_get_bands`` -- utils/inference.py:184-236.  Later writes overwrite earlier ones:
    grid, then last column, then last row, then corner."""
    corner = patches[-1]
    last_row = patches[-lr - 1:-1]
    last_column = patches[-lc - lr - 1:-lr - 1]
    grid = patches[:-lc - lr - 1]
    img = np.zeros((h, w))
    k = 0
    for i in range(hs):
        for j in range(ws):
            img[i * patch_size:(i + 1) * patch_size, j * patch_size:(j + 1) * patch_size] = grid[k]
            k += 1
    for i in range(lc):
        img[i * patch_size:(i + 1) * patch_size, w - patch_size:] = last_column[i]
    for i in range(lr):
        img[h - patch_size:, i * patch_size:(i + 1) * patch_size] = last_row[i]
    img[h - patch_size:, w - patch_size:] = corner
    return img


# ------------------------------------------------------------------------------------------------ augmentation
def augment_patch(img: np.ndarray, lbl: np.ndarray, rot_deg: int, flip0: bool, flip1: bool):
    """Reference utils/dataloaders.py:152-163 (`onera_siamese_loader`, aug branch) with the three random draws made
    explicit: ``img`` [2, C, S, S] (both dates), ``lbl`` [S, S]; rot90 by ``rot_deg`` quarter turns, then the two flips."""
    out_img = np.rot90(img, rot_deg, [2, 3]).copy()              # :153
    out_lbl = np.rot90(lbl, rot_deg, [0, 1]).copy()              # :154
    if flip0:                                                    # :156-158
        out_img = np.flip(out_img, axis=2).copy()
        out_lbl = np.flip(out_lbl, axis=0).copy()
    if flip1:                                                    # :160-162
        out_img = np.flip(out_img, axis=3).copy()
        out_lbl = np.flip(out_lbl, axis=1).copy()
    return out_img, out_lbl


def augment_source_index(i: int, j: int, size: int, rot_deg: int, flip0: bool, flip1: bool):
    """(row, col) in the UN-augmented patch that lands at (i, j) after `augment_patch` -- the gather form the device
    kernel uses (fabric_b200/csrc/capi.cu: pack_nchw16_kernel with `aug`)."""
    s1 = size - 1
    if flip1:
        j = s1 - j
    if flip0:
        i = s1 - i
    rot_deg %= 4
    if rot_deg == 1:
        return j, s1 - i
    if rot_deg == 2:
        return s1 - i, s1 - j
    if rot_deg == 3:
        return s1 - j, i
    return i, j
