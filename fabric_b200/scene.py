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


class ScenePlan:
    """Row-band sharding of one scene geometry over `world` ranks (BASELINE configs[4]).

    The reference's tile list (`_get_patches`) is cut into TILE ROWS: row i of the hs x ws grid together with its
    last-column tile; the bottom strip (the `lr` last-row tiles + the corner, rows [h-p, h)) belongs to the unit of the
    last grid row, whose rows it partly overwrites.  Rank r owns a contiguous run of tile rows, needs ONLY the scene rows
    [row0, row1) those tiles cover, and produces exactly rows [row0, row1) of the mask -- bands are disjoint, so the
    full mask is the concatenation of the bands (no reduction), and the reference's overwrite order (grid, last column,
    last row, corner; inference.py:219-234) is kept inside each band."""

    def __init__(self, h: int, w: int, patch_size: int, rank: int = 0, world: int = 1):
        p = patch_size
        if h < p or w < p:
            raise ValueError("scene smaller than the patch size")
        self.h, self.w, self.p, self.rank, self.world = h, w, p, rank, world
        hs, ws = h // p, w // p
        self.hs, self.ws, self.lc, self.lr = hs, ws, hs, ws
        self.n_tiles = hs * ws + hs + ws + 1
        # tiles per tile-row unit: ws grid tiles + 1 last-column tile (+ the bottom strip for the last unit)
        cost = [ws + 1] * hs
        cost[-1] += ws + 1
        total = sum(cost)
        # contiguous split balanced by tile count: unit u goes to the rank its cumulative midpoint falls into
        bounds, acc = [0], 0
        for r in range(world):
            target = total * (r + 1) / world
            u = bounds[-1]
            while u < hs and acc + cost[u] / 2 <= target:
                acc += cost[u]
                u += 1
            bounds.append(u)
        bounds[-1] = hs
        self.u0, self.u1 = bounds[rank], bounds[rank + 1]
        last = self.u1 == hs and self.u1 > self.u0
        self.row0 = self.u0 * p
        self.row1 = h if last else self.u1 * p
        if self.u1 == self.u0:
            self.row1 = self.row0
        self.band_rows = self.row1 - self.row0       # scene rows this rank holds == mask rows it produces
        self.out_rows = self.band_rows
        # this rank's tiles as (row, col) origins RELATIVE to its band, grouped by overwrite class
        grid = [((i * p) - self.row0, j * p) for i in range(self.u0, self.u1) for j in range(ws)]
        col = [((i * p) - self.row0, w - p) for i in range(self.u0, self.u1)]
        strip = [((h - p) - self.row0, j * p) for j in range(ws)] if last else []
        corner = [((h - p) - self.row0, w - p)] if last else []
        self.origins_list = grid + col + strip + corner
        self.n_mine = len(self.origins_list)
        self.classes, first = [], 0
        for cls in (grid, col, strip, corner):
            self.classes.append((first, len(cls)))
            first += len(cls)
        self.all_bands = [(bounds[r] * p, (h if bounds[r + 1] == hs and bounds[r + 1] > bounds[r] else bounds[r + 1] * p)
                           if bounds[r + 1] > bounds[r] else bounds[r] * p) for r in range(world)]


@torch.no_grad()
def predict_scene_band(model, band_d1: torch.Tensor, band_d2: torch.Tensor, plan: ScenePlan, batch_size: int = 64,
                       mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """This rank's band of the change mask.  band_d*: [13, plan.band_rows, W] (rows [plan.row0, plan.row1) of the scene;
    fp32 already z-scored, or raw uint16 with `mean` / `std`), on the model's device.  Returns uint8 [band_rows, W]."""
    dev = band_d1.device
    c, rows, w = band_d1.shape
    assert rows == plan.band_rows and w == plan.w, "band does not match the plan"
    p, n = plan.p, plan.n_mine
    canvas = torch.zeros((rows, w), dtype=torch.uint8, device=dev)
    if n == 0:
        return canvas
    with torch.cuda.device(dev):
        origins = torch.tensor(plan.origins_list, dtype=torch.int32, device=dev)
        inv_std = (1.0 / std).float().contiguous() if std is not None else None
        mean = mean.float().contiguous() if mean is not None else None
        masks = ops._empty((n, p, p), dtype=torch.uint8, device=dev)
        x5 = ops._empty((2, min(batch_size, n), p, p, ops.cpad(c)), dtype=torch.bfloat16, device=dev)
        was_training = model.training
        model.eval()
        for s in range(0, n, batch_size):
            k = min(batch_size, n - s)
            buf = x5[:, :k] if k == x5.shape[1] else ops._empty((2, k, p, p, ops.cpad(c)), dtype=torch.bfloat16, device=dev)
            ops.gather_tiles(band_d1, origins[s:s + k], p, mean, inv_std, out=buf[0])
            ops.gather_tiles(band_d2, origins[s:s + k], p, mean, inv_std, out=buf[1])
            logits = model.forward_packed(buf)
            ops.argmax_metrics(logits, mask_out=masks[s:s + k])
        model.train(was_training)
        for first, count in plan.classes:        # overwrite order of the reference: grid, last column, last row, corner
            if count:
                ops.scatter_tiles(masks, origins, canvas, first, count)
    return canvas


def gather_bands(band: torch.Tensor, plan: ScenePlan, process_group=None, dst: int = 0) -> Optional[torch.Tensor]:
    """Concatenate the ranks' mask bands on rank `dst` (uint8 [H, W]); other ranks get None.  One `gather` of equal-size
    (padded) bands -- the only collective of scene inference."""
    if plan.world == 1:
        return band
    import torch.distributed as dist
    max_rows = max(b1 - b0 for b0, b1 in plan.all_bands)
    padded = torch.zeros((max_rows, plan.w), dtype=torch.uint8, device=band.device)
    padded[:band.shape[0]] = band
    outs = [torch.empty_like(padded) for _ in range(plan.world)] if plan.rank == dst else None
    dist.gather(padded, outs, dst=dst, group=process_group)
    if plan.rank != dst:
        return None
    return torch.cat([o[:b1 - b0] for o, (b0, b1) in zip(outs, plan.all_bands)], dim=0)


@torch.no_grad()
def predict_scene(model, scene_d1: torch.Tensor, scene_d2: torch.Tensor, patch_size: int = 256, batch_size: int = 64,
                  mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None, rank: int = 0, world: int = 1,
                  process_group=None) -> Tuple[Optional[torch.Tensor], dict]:
    """Change mask of a whole bi-date scene (the loop at reference train.py:182-205 + utils/inference.py:134-236).
    scene_d*: [13,H,W] fp32 (already z-scored) or uint16 raw bands with `mean`/`std` [13] (dataloaders.py:94-99), on the
    model's device.  With `world` > 1 the scene is sharded by row bands (``ScenePlan``): this convenience wrapper slices
    the band out of the full scene it is given; callers that load the scene themselves should upload only rows
    [plan.row0, plan.row1) and call ``predict_scene_band``.  Returns (canvas uint8 [H,W] on rank 0 else None, info)."""
    c, h, w = scene_d1.shape
    plan = ScenePlan(h, w, patch_size, rank, world)
    band = predict_scene_band(model, scene_d1[:, plan.row0:plan.row1].contiguous() if world > 1 else scene_d1,
                              scene_d2[:, plan.row0:plan.row1].contiguous() if world > 1 else scene_d2, plan, batch_size,
                              mean, std)
    canvas = gather_bands(band, plan, process_group)
    info = dict(tiles=plan.n_tiles, tiles_this_rank=plan.n_mine, hs=plan.hs, ws=plan.ws, lc=plan.lc, lr=plan.lr,
                rows=(plan.row0, plan.row1))
    return canvas, info
