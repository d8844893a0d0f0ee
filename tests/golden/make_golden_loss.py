"""Golden vectors for the loss row from the reference's OWN Python (utils/loss_utils.py), imported from /root/reference and
run on CPU in the build container (the GPU box has no /root/reference).  Writes tests/golden/loss1.npz.

    python tests/golden/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GS_REFERENCE_ROOT", "/root/reference")


def inputs(seed=51, C=3, H=70, W=90):
    """A smooth image pair in [0,1] plus noise (sizes deliberately not multiples of 16)."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(6 * xx + c) * torch.cos(4 * yy - c) for c in range(C)])
    gt = (base + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1).contiguous()
    img = (base * 0.9 + 0.05 + 0.08 * torch.randn(C, H, W, generator=g)).clamp(0, 1).contiguous()
    return img, gt


def main():
    sys.path.insert(0, REF)
    from utils.loss_utils import l1_loss, ssim          # the reference's functions, unmodified
    img, gt = inputs()
    lam = 0.2
    x = img.clone().requires_grad_(True)
    Ll1 = l1_loss(x, gt)
    s = ssim(x, gt)
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - s)          # train.py:110-115
    loss.backward()
    x2 = img.clone().requires_grad_(True)
    ssim(x2, gt).backward()
    out = dict(l1=np.float32(Ll1.item()), ssim=np.float32(s.item()), loss=np.float32(loss.item()), grad=x.grad.numpy(),
               grad_ssim_only=x2.grad.numpy(), lambda_dssim=np.float32(lam))
    np.savez_compressed(os.path.join(HERE, "loss1.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
