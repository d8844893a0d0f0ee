"""A minimal stand-in for the reference's `GaussianModel` (scene/gaussian_model.py) — exactly the attributes
`gaussian_renderer.render()` reads (GR:19-148) — on top of either fp32 tensors or the quantised device layout.

    view = GaussianModelView.from_ply("point_cloud_quantised.ply", quantised=True, device="cuda")
    image = gaussian_renderer.render(camera, view, pipe, background)["render"]

With a quantised model the u8 id planes and the 20x256 centre table stay as they are on the device (`view.quant`); render()
hands them to the fused de-quantising preprocess, so the 248 B/Gaussian fp32 expansion of the reference's `load_ply`
(GM:371-387) never exists in memory.  The fp32 tensors are then only placeholders of the right shape for the autograd graph
(xyz is the one attribute that is not quantised, GM:285).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ply, synth


class GaussianModelView:
    def __init__(self, scene: synth.Scene, device, quant: Optional[synth.QuantScene] = None, prune_mask: Optional[torch.Tensor] = None,
                 requires_grad: bool = True, variable_sh_bands: bool = False):
        """`variable_sh_bands` (inference, reference GaussianModel(sh_degree, variable_sh_bands=True)): `get_features` is then the
        reference's LIST of per-degree tensors [N_d, (d+1)^2, 3] (scene/gaussian_model.py:153-163) that
        `render(..., variable_sh_bands=True)` flattens into the packed layout of rasterize_gaussians_variableSH_bands; the
        Gaussians must be ordered by degree, as the reference's per-degree PLY groups are."""
        g = requires_grad
        self.variable_sh_bands = variable_sh_bands
        if variable_sh_bands:
            deg = scene.degrees.view(-1)
            if not bool((deg[1:] >= deg[:-1]).all()):
                raise ValueError("variable_sh_bands needs Gaussians ordered by SH degree (the reduced-3dgs PLY groups)")
            if quant is not None:
                raise ValueError("variable_sh_bands is the fp32 packed-SH inference layout; a quantised model uses the fused id planes")
        self._xyz = scene.means3D.to(device).requires_grad_(g)
        self._opacity = scene.opacity.to(device).requires_grad_(g and quant is None)          # raw logits (GM:149-150 activates later)
        self._scaling = scene.scales.to(device).requires_grad_(g and quant is None)           # already exp-activated
        self._rotation = scene.rotations.to(device).requires_grad_(g and quant is None)       # already normalised
        self._features = scene.sh.to(device).requires_grad_(g and quant is None)              # [P,16,3]
        self._degrees = scene.degrees.to(device)
        self.active_sh_degree = self.max_sh_degree = 3
        self.quant = None if quant is None else quant.to(device)
        self.prune_mask = None if prune_mask is None else prune_mask.to(device)
        self.per_band_count = [int((scene.degrees == d).sum()) for d in range(4)]

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: s._scaling)
    get_rotation = property(lambda s: s._rotation)
    @property
    def get_features(self):
        if not self.variable_sh_bands:
            return self._features
        out, start = [], 0                                   # gaussian_model.py:153-163: cat(features_dc[group], features_rest[group]) per degree
        for d, n in enumerate(self.per_band_count):
            out.append(self._features[start:start + n, :(d + 1) ** 2, :])
            start += n
        return out

    num_primitives = property(lambda s: int(s._xyz.shape[0]))

    def params(self):
        return [self._xyz, self._opacity, self._scaling, self._rotation, self._features]

    @classmethod
    def from_ply(cls, path: str, quantised: bool = True, half_float: bool = False, device="cuda", requires_grad: bool = False):
        """Reduced-3DGS PLY (GM:239-311) -> view.  Quantised files keep their id planes; the fp32 members are de-quantised once
        here only because render() wants tensors of the right shape to hang the graph on (they are not read by the kernels)."""
        m = ply.load_reduced_ply(path, half_float=half_float, quantised=quantised, device=device)
        if quantised:
            return cls(m.dequantise(), device, quant=m, requires_grad=requires_grad)
        return cls(m, device, requires_grad=requires_grad)
