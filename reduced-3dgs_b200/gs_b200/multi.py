"""Replicate-scene / shard-views data parallelism on torch.distributed.

The hot path shards naturally over views: each view's forward+backward is independent given the scene.  So:
  * the scene is broadcast once from rank 0 (`broadcast_scene`) — the only transfer of model state;
  * rank g of N renders views {g, g+N, ...} (`shard_views`) with NO data-path collective;
  * when a view batch is used for one optimisation step, every rank accumulates per-view gradients locally
    (the kernels add into the buffers, `GsbGrads.accumulate`), and ONE all-reduce(sum) over the per-Gaussian
    gradient buffers closes the batch (`GradAccumulator.all_reduce`).  The per-view densification statistics are
    non-linear per view (||dL_dmean2D|| per view, reference gaussian_model.py:693-695) and are therefore
    accumulated per view BEFORE the reduction.
Works with backend "nccl" (one process per GPU, NVLink/NVSwitch) and "gloo" (CPU tensors, used by the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
GRAD_WIDTH = {"dL_dmeans2D": 3, "dL_dcolors": 3, "dL_dopacity": 1, "dL_dmeans3D": 3, "dL_dcov3D": 6, "dL_dscales": 3,
              "dL_drotations": 4}


def world() -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_views(n_views: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Indices of the views rank `rank` renders: {rank, rank + N, ...}."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, n_views, world_size))


def broadcast_scene(tensors: Sequence[torch.Tensor], src: int = 0) -> None:
    """In-place broadcast of the scene tensors (fp32 attributes or u8 id planes + codebooks) from `src`."""
    _, w = world()
    if w == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)


class GradAccumulator:
    """Per-Gaussian gradient buffers that live across the views of a batch.

    `buffers()` returns the 8-tuple in the order of rasterize_gaussians_backward (pass it as `accumulate_into`, and
    `view_means2D` as the keyword of the same name); `observe_view()` folds the per-view, non-linear densification
    statistics from THIS view's screen-space gradient; `all_reduce()` closes the batch:

      * one SUM over `small`: the 14 non-SH floats per Gaussian the optimiser consumes (dL_dmeans2D 3, dL_dopacity 1,
        dL_dmeans3D 3, dL_dscales 3, dL_drotations 4) followed by the two additive statistics (xyz_gradient_accum, denom);
      * the SUM over dL_dsh (3M floats per Gaussian; 14 + 3M = 62 at M = 16).  With `band_counts` (the model's per-degree group
        sizes; Gaussians ordered by degree, as reduced-3dgs models with variable SH bands are — gaussian_model.py:281-308) only the
        ACTIVE coefficients travel: group d sends its [N_d, (d+1)^2, 3] block (groups below the top degree are packed into one
        contiguous buffer, the top group is contiguous as it is) — the rest of the plane is zero on every rank by construction;
      * one MAX over max_radii2D.
    dL_dcolors / dL_dcov3D are intermediates nobody trains on when SH and scale/rotation are the parameters, so they stay local
    (`local`, 9 floats per Gaussian, not communicated).  `all_reduce(async_op=True)` returns after launching the collectives
    (on NCCL's communication stream) — `wait()` joins them and unpacks — so a caller can overlap the reduction with work that does
    not touch the buffers."""

    SMALL = ("dL_dmeans2D", "dL_dopacity", "dL_dmeans3D", "dL_dscales", "dL_drotations")

    def __init__(self, P: int, M: int, device, band_counts: Optional[Sequence[int]] = None):
        self.P, self.M = P, M
        n_small = sum(GRAD_WIDTH[n] for n in self.SMALL)           # 14
        self.floats_per_gaussian = n_small + 3 * M                  # 62 at M = 16: what the optimiser consumes
        self.small = torch.zeros(P * (n_small + 2), dtype=torch.float32, device=device)
        self.sh = torch.zeros(P, M, 3, dtype=torch.float32, device=device)
        self.local = torch.zeros(P * (GRAD_WIDTH["dL_dcolors"] + GRAD_WIDTH["dL_dcov3D"]), dtype=torch.float32, device=device)
        views, off = {"dL_dsh": self.sh}, 0
        for name in self.SMALL:
            w = GRAD_WIDTH[name]
            views[name] = self.small[off:off + P * w].view(P, w)
            off += P * w
        # densification statistics (reference gaussian_model.py:693-695 add_densification_stats, train.py:134-139): the additive ones
        # ride at the end of `small`
        self.xyz_gradient_accum = self.small[off:off + P].view(P, 1)
        self.denom = self.small[off + P:off + 2 * P].view(P, 1)
        views["dL_dcolors"] = self.local[:3 * P].view(P, 3)
        views["dL_dcov3D"] = self.local[3 * P:].view(P, 6)
        self._views = [views[n] for n in GRAD_NAMES]
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=device)
        self.view_means2D = torch.zeros(P, 3, dtype=torch.float32, device=device)   # written (not added to) by every backward
        self.n_views = 0
        self._pending = []
        # degree-banded SH payload
        self._groups = None
        if band_counts is not None:
            counts = [int(c) for c in band_counts]
            if sum(counts) != P or any((d + 1) ** 2 > M for d, c in enumerate(counts) if c):
                raise ValueError("band_counts must sum to P and fit the SH layout")
            self._groups, start = [], 0
            for d, c in enumerate(counts):
                if c:
                    self._groups.append((start, start + c, (d + 1) ** 2))
                start += c
            n_packed = sum((e - b) * k * 3 for b, e, k in self._groups if k < M)
            self._sh_packed = torch.zeros(n_packed, dtype=torch.float32, device=device)

    @property
    def payload_floats(self) -> int:
        """floats that cross the wire per all_reduce() (SUM collectives)."""
        if self._groups is None:
            return self.small.numel() + self.sh.numel()
        return self.small.numel() + self._sh_packed.numel() + sum((e - b) * k * 3 for b, e, k in self._groups if k == self.M)

    def buffers(self):
        return tuple(self._views)

    def zero_(self):
        self.small.zero_()
        self.sh.zero_()
        self.local.zero_()
        self.max_radii2D.zero_()
        self.n_views = 0

    def observe_view(self, radii: torch.Tensor, dL_dmeans2D_view: Optional[torch.Tensor] = None):
        """Per-view statistics must be taken from THIS view's screen-space gradient (`view_means2D`, which the backward wrote
        for the view just rendered), before it is summed with other views."""
        g = self.view_means2D if dL_dmeans2D_view is None else dL_dmeans2D_view
        vis = (radii > 0).view(-1, 1)
        self.xyz_gradient_accum += torch.where(vis, torch.norm(g[:, :2], dim=-1, keepdim=True), torch.zeros_like(self.xyz_gradient_accum))
        self.denom += vis.to(torch.float32)
        self.max_radii2D = torch.where(vis.view(-1), torch.max(self.max_radii2D, radii.to(torch.float32)), self.max_radii2D)
        self.n_views += 1

    def inactive_sh_is_zero(self) -> bool:
        """True when no gradient sits outside the active coefficients of any degree group (what the banded payload relies on; the
        backward kernels never write beyond a Gaussian's degree).  Synchronises: for tests and warm-up, not the timed path."""
        if self._groups is None:
            return True
        return all(float(self.sh[b:e, k:, :].abs().sum()) == 0.0 for b, e, k in self._groups if k < self.M)

    def _packed_views(self):
        off = 0
        for b, e, k in self._groups:
            if k < self.M:
                n = (e - b) * k * 3
                yield self.sh[b:e, :k, :], self._sh_packed[off:off + n].view(e - b, k, 3)
                off += n

    def all_reduce(self, async_op: bool = False):
        _, w = world()
        if w == 1:
            return
        self._pending = [dist.all_reduce(self.small, op=dist.ReduceOp.SUM, async_op=True)]
        if self._groups is None:
            self._pending.append(dist.all_reduce(self.sh, op=dist.ReduceOp.SUM, async_op=True))
        else:
            for src, dst in self._packed_views():
                dst.copy_(src)                                   # active coefficients of the lower-degree groups -> one contiguous buffer
            if self._sh_packed.numel():
                self._pending.append(dist.all_reduce(self._sh_packed, op=dist.ReduceOp.SUM, async_op=True))
            for b, e, k in self._groups:
                if k == self.M:                                  # the top-degree group's rows are contiguous and fully active
                    self._pending.append(dist.all_reduce(self.sh[b:e], op=dist.ReduceOp.SUM, async_op=True))
        self._pending.append(dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX, async_op=True))
        if not async_op:
            self.wait()

    def wait(self):
        for h in self._pending:
            h.wait()
        if self._pending and self._groups is not None:
            for src, dst in self._packed_views():
                src.copy_(dst)
        self._pending = []


def gather_images(image: torch.Tensor, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Optional: collect the rendered images of all ranks on `dst` (forward-only serving)."""
    r, w = world()
    if w == 1:
        return [image]
    out = [torch.empty_like(image) for _ in range(w)] if r == dst else None
    dist.gather(image, out, dst=dst)
    return out
