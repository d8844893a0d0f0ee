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

    `buffers()` returns the 8-tuple in the order of rasterize_gaussians_backward (pass it as `accumulate_into`);
    `observe_view()` folds the per-view, non-linear densification statistics; `all_reduce()` sums the buffers
    over ranks in one flat collective per dtype."""

    def __init__(self, P: int, M: int, device):
        self.P, self.M = P, M
        widths = [3, 3, 1, 3, 6, 3 * M, 3, 4]
        self.flat = torch.zeros(P * sum(widths), dtype=torch.float32, device=device)
        self._views, off = [], 0
        for name, w in zip(GRAD_NAMES, widths):
            v = self.flat[off:off + P * w]
            off += P * w
            self._views.append(v.view(P, M, 3) if name == "dL_dsh" else v.view(P, w))
        # densification statistics (reference gaussian_model.py:693-695 add_densification_stats, train.py:134-139)
        self.xyz_gradient_accum = torch.zeros(P, 1, dtype=torch.float32, device=device)
        self.denom = torch.zeros(P, 1, dtype=torch.float32, device=device)
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=device)
        self.n_views = 0

    def buffers(self):
        return tuple(self._views)

    def zero_(self):
        self.flat.zero_()
        self.n_views = 0

    def observe_view(self, dL_dmeans2D_view: torch.Tensor, radii: torch.Tensor):
        """Per-view statistics must be taken from THIS view's screen-space gradient, before it is summed with others."""
        vis = radii > 0
        self.xyz_gradient_accum[vis] += torch.norm(dL_dmeans2D_view[vis, :2], dim=-1, keepdim=True)
        self.denom[vis] += 1
        self.max_radii2D[vis] = torch.max(self.max_radii2D[vis], radii[vis].to(torch.float32))
        self.n_views += 1

    def all_reduce(self):
        _, w = world()
        if w == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        dist.all_reduce(self.xyz_gradient_accum, op=dist.ReduceOp.SUM)
        dist.all_reduce(self.denom, op=dist.ReduceOp.SUM)
        dist.all_reduce(self.max_radii2D, op=dist.ReduceOp.MAX)


def gather_images(image: torch.Tensor, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Optional: collect the rendered images of all ranks on `dst` (forward-only serving)."""
    r, w = world()
    if w == 1:
        return [image]
    out = [torch.empty_like(image) for _ in range(w)] if r == dst else None
    dist.gather(image, out, dst=dst)
    return out
