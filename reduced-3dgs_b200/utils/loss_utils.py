"""Drop-in for the reference's `utils/loss_utils.py` (l1_loss, l2_loss, ssim) on the fused CUDA loss kernels of
include/gs_b200.h (gsb_l1_ssim_forward / _backward), plus `l1_ssim_loss` = the combination train.py:110-115 builds
((1 - lambda_dssim) * L1 + lambda_dssim * (1 - SSIM)) in ONE forward and ONE backward launch.

The kernels implement the training configuration of the reference: window_size 11 (sigma 1.5), mean over the whole image,
gradient w.r.t. the first image only (the ground truth has none).  Other settings raise instead of silently running a
different code path (there is no PyTorch fallback in this package).
"""
from __future__ import annotations

import torch

from gs_b200 import lib as _lib
from gs_b200.lib import f32, ptr


class _L1SSIM(torch.autograd.Function):
    """returns (l1, ssim) as 0-dim tensors; backward takes dL/dl1 and dL/dssim."""

    @staticmethod
    def forward(ctx, image, gt):
        if not image.is_cuda:
            raise RuntimeError("gs_b200: loss kernels need CUDA tensors (no CPU path exists)")
        dev = image.device
        x, y = f32(image, dev), f32(gt, dev)
        if x.dim() == 4 and x.shape[0] == 1:
            x, y = x[0], y[0]
        if x.dim() != 3 or x.shape != y.shape:
            raise RuntimeError("l1 / ssim: expected two [C,H,W] images of the same shape")
        C, H, W = (int(v) for v in x.shape)
        L = _lib.lib()
        maps = torch.empty((3, C, H, W), dtype=torch.float32, device=dev)
        partial = torch.empty((int(L.gsb_l1_ssim_blocks(C, H, W)), 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.gsb_l1_ssim_forward(ptr(x), ptr(y), C, H, W, ptr(maps), ptr(partial), _lib.current_stream(dev)))
        sums = partial.sum(dim=0, dtype=torch.float64) / float(C * H * W)
        ctx.save_for_backward(x, y, maps)
        ctx.shape = tuple(image.shape)
        return sums[1].float(), sums[0].float()

    @staticmethod
    def backward(ctx, g_l1, g_ssim):
        x, y, maps = ctx.saved_tensors
        C, H, W = (int(v) for v in x.shape)
        dev = x.device
        L = _lib.lib()
        # the incoming gradients stay on the device; an output that was not used contributes with coefficient 0
        u1 = g_l1.to(torch.float32).reshape(1).contiguous() if g_l1 is not None else None
        u2 = g_ssim.to(torch.float32).reshape(1).contiguous() if g_ssim is not None else None
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(L.gsb_l1_ssim_backward(ptr(x), ptr(y), C, H, W, ptr(maps), 1.0 if u1 is not None else 0.0, ptr(u1),
                                              1.0 if u2 is not None else 0.0, ptr(u2), ptr(out), _lib.current_stream(dev)))
        return out.view(ctx.shape), None


class _L1SSIMLoss(torch.autograd.Function):
    """(1 - lambda) * L1 + lambda * (1 - SSIM) with a sync-free backward (lambda is a Python float)."""

    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        l1, ssim_v = _L1SSIM.forward(ctx, image, gt)
        ctx.lam = float(lambda_dssim)
        return (1.0 - ctx.lam) * l1 + ctx.lam * (1.0 - ssim_v)

    @staticmethod
    def backward(ctx, g):
        x, y, maps = ctx.saved_tensors
        C, H, W = (int(v) for v in x.shape)
        dev = x.device
        L = _lib.lib()
        up = g.to(torch.float32).reshape(1).contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(L.gsb_l1_ssim_backward(ptr(x), ptr(y), C, H, W, ptr(maps), 1.0 - ctx.lam, ptr(up), -ctx.lam, ptr(up), ptr(out),
                                              _lib.current_stream(dev)))
        return out.view(ctx.shape), None, None


def l1_ssim_loss(image, gt, lambda_dssim: float = 0.2):
    """train.py:110-115: (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt)), fused."""
    return _L1SSIMLoss.apply(image, gt, lambda_dssim)


def l1_loss(network_output, gt):
    """loss_utils.py:17-18."""
    return _L1SSIM.apply(network_output, gt)[0]


def l2_loss(network_output, gt):
    """loss_utils.py:20-21 (not on the training path; plain tensor arithmetic as in the reference)."""
    return ((network_output - gt) ** 2).mean()


def ssim(img1, img2, window_size=11, size_average=True, aggregate=True):
    """loss_utils.py:33-50, training configuration only."""
    if window_size != 11 or not size_average or not aggregate:
        raise NotImplementedError("gs_b200 ssim: only window_size=11, size_average=True, aggregate=True (the training configuration, "
                                  "train.py:111) is implemented")
    if img2.requires_grad:
        raise NotImplementedError("gs_b200 ssim: the gradient w.r.t. the second image is not implemented (it is the ground truth)")
    return _L1SSIM.apply(img1, img2)[1]
