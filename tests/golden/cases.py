"""Definitions of the golden-vector cases (inputs are regenerated from seeds, only outputs are stored).

`c1` is BASELINE.json configs[0] (10k random Gaussians, SH degree 0, 256x256, forward); the others are small
cases that cover the branches the reference has: uniform degree-3 SH with backward and a non-black background on a
non-multiple-of-16 image, mixed per-Gaussian degrees with the SH-sparsity term and the packed variable-SH inference
entry point, and precomputed covariance / colour inputs.
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "reduced-3dgs_b200"))
from gs_b200 import synth  # noqa: E402

CASES = {
    "c1": dict(P=10_000, seed=1, W=256, H=256, sh_degree=0, mixed=False, log_scale=math.log(0.02), bg=(0.0, 0.0, 0.0),
               backward=True, lam=0.0, packed=False, precomp=False),
    "g2": dict(P=4_000, seed=12, W=200, H=120, sh_degree=3, mixed=False, log_scale=math.log(0.05), bg=(0.1, 0.5, 0.9),
               backward=True, lam=0.0, packed=False, precomp=False),
    "g3": dict(P=4_000, seed=13, W=176, H=144, sh_degree=3, mixed=True, log_scale=math.log(0.045), bg=(0.3, 0.2, 0.1),
               backward=True, lam=0.1, packed=True, precomp=False),
    "g4": dict(P=2_000, seed=14, W=128, H=96, sh_degree=1, mixed=False, log_scale=math.log(0.07), bg=(1.0, 1.0, 1.0),
               backward=True, lam=0.0, packed=False, precomp=True),
}


def build_inputs(name):
    c = CASES[name]
    W, H = c["W"], c["H"]
    scene = synth.make_scene(c["P"], c["seed"], sh_degree=c["sh_degree"], mixed_degrees=c["mixed"],
                             box=(1.9 * W / H, 1.9, 1.0), log_scale_mean=c["log_scale"],
                             M=16 if c["mixed"] else None, near_frac=0.01)
    cam = synth.make_camera(W, H)
    bg = torch.tensor(c["bg"], dtype=torch.float32)
    dL = synth.grad_image(W, H, c["seed"] + 1)
    extra = {}
    if c["precomp"]:
        # precomputed covariance (reference pipe.compute_cov3D_python, gaussian_model.py:50-54) and colours
        g = torch.Generator().manual_seed(c["seed"] + 7)
        q, s = scene.rotations, scene.scales
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        Rm = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                          2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                          2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
        Lm = Rm * s.view(-1, 1, 3)
        cov = Lm @ Lm.transpose(1, 2)
        extra["cov3D_precomp"] = torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2],
                                              cov[:, 2, 2]], dim=1).contiguous()
        extra["colors_precomp"] = torch.rand(c["P"], 3, generator=g).contiguous()
    return c, scene, cam, bg, dL, extra
