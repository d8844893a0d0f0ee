"""Run OUR CUDA path through the reference-facing `_C` API and return numpy dicts shaped like oracle outputs
(shared by the GPU parity tests, __graft_entry__.smoke and tools/)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reduced-3dgs_b200"))

from diff_gaussian_rasterization import _C  # noqa: E402

EMPTY = torch.Tensor([])


def forward_args(scene, cam, bg, extra=None, dev="cuda"):
    extra = extra or {}
    cov, col = extra.get("cov3D_precomp"), extra.get("colors_precomp")
    return (bg.to(dev), scene.means3D.to(dev), EMPTY if col is None else col.to(dev), scene.opacity.to(dev),
            EMPTY if cov is not None else scene.scales.to(dev), EMPTY if cov is not None else scene.rotations.to(dev), 1.0,
            EMPTY if cov is None else cov.to(dev), cam.world_view_transform.to(dev), cam.full_proj_transform.to(dev),
            math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), cam.image_height, cam.image_width,
            EMPTY if col is not None else scene.sh.to(dev), scene.degrees.to(dev), cam.camera_center.to(dev), False, False)


def run_forward(scene, cam, bg, extra=None, prune_mask=None, quant=None, dev="cuda"):
    """-> (args, raw outputs, dict of numpy intermediates in reference layouts)."""
    args = forward_args(scene, cam, bg, extra, dev)
    dbg = {}
    out = _C.rasterize_gaussians(*args, prune_mask=None if prune_mask is None else prune_mask.to(dev),
                                 quant=None if quant is None else quant.to(dev), debug_out=dbg)
    R, color, radii, geomB, binB, imgB = out
    st = _C.export_state(geomB, binB, imgB, R, cam.image_width, cam.image_height, P=scene.means3D.shape[0])
    torch.cuda.synchronize()
    res = dict(num_rendered=R, color=color.cpu().numpy(), radii=radii.cpu().numpy(),
               depths=dbg["depths"].cpu().numpy(), means2D=dbg["means2D"].cpu().numpy(), cov3D=dbg["cov3D"].cpu().numpy(),
               conic_opacity=dbg["conic_opacity"].cpu().numpy(), rgb=dbg["rgb"].cpu().numpy(),
               tiles_touched=dbg["tiles_touched"].cpu().numpy().astype(np.uint32), clamped=dbg["clamped"].cpu().numpy(),
               keys=st["keys"].cpu().numpy().astype(np.uint64), point_list=st["point_list"].cpu().numpy().astype(np.uint32),
               ranges=st["ranges"].cpu().numpy().astype(np.uint32), final_T=st["final_T"].cpu().numpy(),
               n_contrib=st["n_contrib"].cpu().numpy().astype(np.uint32))
    return args, out, res


GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]


def run_backward(args, out, dL, lam=0.0, prune_mask=None, quant=None):
    (bg, means3D, colors, opacity, scales, rotations, mod, cov, view, proj, tx, ty, H, W, sh, degrees, campos, _, _) = args
    R, color, radii, geom, binning, img = out
    dev = means3D.device
    grads = _C.rasterize_gaussians_backward(bg, means3D, radii, colors, scales, rotations, mod, cov, view, proj, tx, ty,
                                            dL.to(dev), sh, degrees, campos, geom, R, binning, img, lam, False,
                                            prune_mask=None if prune_mask is None else prune_mask.to(dev),
                                            quant=None if quant is None else quant.to(dev), want_conic=True)
    torch.cuda.synchronize()
    res = {n: g.cpu().numpy() for n, g in zip(GRAD_NAMES + ["dL_dconic"], grads)}
    return res
