"""Full-scene tiled inference on the device: the B200 version of reference ``utils/inference.py`` (`_get_patches`
:134-181, `_get_bands` :184-236) and the prediction loop at ``train.py:182-205``.

The reference extracts every patch on the host (``np.vstack`` of N x 13 x p x p fp32 copies), ships each batch to the
GPU, pulls each argmax mask back and reassembles the (h, w) canvas with a Python double loop.  Here the bi-date
scene stays resident in HBM (13 x 10000 x 10000 fp32 = 5.2 GB per date), tiles are gathered straight into packed
NHWC bf16 batches (normalisation fused), masks are written into a device canvas, and tiles shard over ranks.

Tile order and overwrite rule are the reference's: the hs x ws grid of non-overlapping tiles, then the last-column
tiles, the last-row tiles and the corner tile; later writes win.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


def tile_origins(h: int, w: int, p: int):
    """(origins [(row, col)...], hs, ws, lc, lr) in the order of reference `_get_patches` (inference.py:152-181)."""
    if h < p or w < p:
        raise ValueError("scene smaller than the patch size")
    hs, ws = h // p, w // p
    origins = [(i * p, j * p) for i in range(hs) for j in range(ws)]          # extract_patches grid
    lc, lr = h // p, w // p
    origins += [(i * p, w - p) for i in range(lc)]                             # last column
    origins += [(h - p, j * p) for j in range(lr)]                             # last row
    origins += [(h - p, w - p)]                                                # corner
    return origins, hs, ws, lc, lr


class SceneTiler:
    """Device-side `_get_patches` / `_get_bands` for one scene geometry."""

    def __init__(self, h: int, w: int, patch_size: int, device):
        self.h, self.w, self.p = h, w, patch_size
        org, self.hs, self.ws, self.lc, self.lr = tile_origins(h, w, patch_size)
        self.n = len(org)
        self.origins = torch.tensor(org, dtype=torch.int32, device=device)
        # tile classes whose members never overlap each other: (first, count), in overwrite order
        g = self.hs * self.ws
        self.classes = [(0, g), (g, self.lc), (g + self.lc, self.lr), (g + self.lc + self.lr, 1)]

    def gather(self, scene: torch.Tensor, first: int, count: int, mean=None, inv_std=None, out=None) -> torch.Tensor:
        """tiles [first, first+count) of `scene` [C,H,W] as packed NHWC bf16 [count,p,p,16]"""
        return ops.gather_tiles(scene, self.origins[first:first + count], self.p, mean, inv_std, out=out)

    def reassemble(self, masks: torch.Tensor) -> torch.Tensor:
        """masks uint8 [N,p,p] (all tiles, reference order) -> canvas uint8 [h,w]  (`_get_bands`)"""
        canvas = torch.zeros((self.h, self.w), dtype=torch.uint8, device=masks.device)
        for first, count in self.classes:
            if count:
                ops.scatter_tiles(masks, self.origins, canvas, first, count)
        return canvas


@torch.no_grad()
def predict_scene(model, scene_d1: torch.Tensor, scene_d2: torch.Tensor, patch_size: int = 256, batch_size: int = 64,
                  mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None, rank: int = 0, world: int = 1,
                  process_group=None) -> Tuple[Optional[torch.Tensor], dict]:
    """Change mask of a whole bi-date scene.  scene_d*: [13,H,W] fp32 (already z-scored) or uint16 raw bands with
    `mean`/`std` [13] (dataloaders.py:94-99), on the model's device.  Tiles are processed `batch_size` at a time; with
    `world` > 1 rank r takes every world-th batch and the masks are gathered on rank 0 (no other collective).
    Returns (canvas uint8 [H,W] on rank 0 else None, info)."""
    dev = scene_d1.device
    c, h, w = scene_d1.shape
    tiler = SceneTiler(h, w, patch_size, dev)
    inv_std = (1.0 / std).float().contiguous() if std is not None else None
    mean = mean.float().contiguous() if mean is not None else None
    n, p = tiler.n, patch_size
    masks = torch.zeros((n, p, p), dtype=torch.uint8, device=dev)
    starts = list(range(0, n, batch_size))
    mine = starts[rank::world]
    x5 = torch.empty((2, batch_size, p, p, ops.cpad(c)), dtype=torch.bfloat16, device=dev)
    was_training = model.training
    model.eval()
    for s in mine:
        k = min(batch_size, n - s)
        buf = x5[:, :k]
        if k < batch_size:
            buf = torch.empty((2, k, p, p, ops.cpad(c)), dtype=torch.bfloat16, device=dev)
        tiler.gather(scene_d1, s, k, mean, inv_std, out=buf[0])
        tiler.gather(scene_d2, s, k, mean, inv_std, out=buf[1])
        logits = model.forward_packed(buf)
        m, _ = ops.argmax_metrics(logits)
        masks[s:s + k] = m
    model.train(was_training)
    if world > 1:
        import torch.distributed as dist
        # every tile was written by exactly one rank (zeros elsewhere): a sum-reduce onto rank 0 is the gather
        dist.reduce(masks, dst=0, op=dist.ReduceOp.SUM, group=process_group)
        if rank != 0:
            return None, dict(tiles=n, tiles_this_rank=sum(min(batch_size, n - s) for s in mine))
    canvas = tiler.reassemble(masks)
    return canvas, dict(tiles=n, tiles_this_rank=sum(min(batch_size, n - s) for s in mine), hs=tiler.hs, ws=tiler.ws,
                        lc=tiler.lc, lr=tiler.lr)
