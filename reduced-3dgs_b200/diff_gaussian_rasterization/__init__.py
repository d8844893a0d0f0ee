"""Drop-in replacement for the reference Python package `diff_gaussian_rasterization`
(submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py): same
GaussianRasterizationSettings (:169-181), GaussianRasterizer (:183-234), rasterize_gaussians (:21-46) and
_RasterizeGaussians autograd op (:48-167), backed by the B200-native kernels through `_C`.

Differences, all additive: GaussianRasterizer.forward takes keyword-only `prune_mask` and `quant`
(fused resolution-aware prune mask / codebook de-quantisation); the forward no longer forces
debug=True (reference :85 hard-wires a device sync after every stage); gradients are allocated uninitialised
because the kernels write every element.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C


def cpu_deep_copy_tuple(input_tuple):
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(means3D, means2D, sh, degrees, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, lambda_sh_sparsity, prune_mask=None, quant=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, degrees, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, lambda_sh_sparsity, prune_mask, quant)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, degrees, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, lambda_sh_sparsity, prune_mask=None, quant=None):
        args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations, raster_settings.scale_modifier,
                cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width, sh, degrees,
                raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)   # Copy them before they can be corrupted (reference :90-97)
            try:
                num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(
                    *args, prune_mask=prune_mask, quant=quant)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(
                *args, prune_mask=prune_mask, quant=quant)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.lambda_sh_sparsity = lambda_sh_sparsity
        ctx.prune_mask = prune_mask
        ctx.quant = quant
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer, degrees)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer,
         degrees) = ctx.saved_tensors
        args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations, raster_settings.scale_modifier,
                cov3Ds_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                raster_settings.tanfovy, grad_out_color, sh, degrees, raster_settings.campos, geomBuffer, num_rendered,
                binningBuffer, imgBuffer, ctx.lambda_sh_sparsity, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                grads8 = _C.rasterize_gaussians_backward(*args, prune_mask=ctx.prune_mask, quant=ctx.quant)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads8 = _C.rasterize_gaussians_backward(*args, prune_mask=ctx.prune_mask, quant=ctx.quant)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = grads8
        if ctx.quant is not None:
            # inputs were id planes: the per-Gaussian attribute gradients have no autograd destination; expose them with the
            # semantics of `.grad`: they accumulate over backward calls until the caller resets `quant.grads = None`
            new = dict(sh=grad_sh, opacity=grad_opacities, scales=grad_scales, rotations=grad_rotations)
            old = getattr(ctx.quant, "grads", None)
            if old:
                for k, g in new.items():
                    old[k].add_(g)
            else:
                ctx.quant.grads = new
        need = ctx.needs_input_grad
        grads = (grad_means3D, grad_means2D, grad_sh if need[2] else None, None,
                 grad_colors_precomp if need[4] else None, grad_opacities if need[5] else None,
                 grad_scales if need[6] else None, grad_rotations if need[7] else None,
                 grad_cov3Ds_precomp if need[8] else None, None, None, None, None)
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, degrees=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, lambda_sh_sparsity=0., *, prune_mask=None, quant=None):
        raster_settings = self.raster_settings
        if quant is None:
            if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
                raise Exception('Please provide excatly one of either SHs or precomputed colors!')
            if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                    ((scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        if shs is None:
            shs = empty
        if colors_precomp is None:
            colors_precomp = empty
        if scales is None:
            scales = empty
        if rotations is None:
            rotations = empty
        if cov3D_precomp is None:
            cov3D_precomp = empty
        if opacities is None:
            opacities = empty
        return rasterize_gaussians(means3D, means2D, shs, degrees, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings, lambda_sh_sparsity, prune_mask, quant)
