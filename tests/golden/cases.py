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


# ---- reduced-3dgs tools: SH-culling statistics, redundancy score ------------------------------
TOOLS_CASES = {
    "t1": dict(P=3_000, seed=21, views=[(200, 120, -8.0), (160, 160, 5.0), (96, 128, 14.0)], log_scale=math.log(0.05), knn=6,
               radius_scale=6.0),
}


def build_tools_inputs(name, P=None, views=None):
    """Scene with mixed SH degrees, a few cameras of different sizes (yaw in degrees about y), brute-force k nearest
    neighbours (the reference takes them from simple-knn, scene/__init__.py:159-160)."""
    import numpy as np
    c = dict(TOOLS_CASES[name])
    if P is not None:
        c["P"] = P
    if views is not None:
        c["views"] = views
    scene = synth.make_scene(c["P"], c["seed"], sh_degree=3, mixed_degrees=True, box=(2.4, 1.9, 1.0), log_scale_mean=c["log_scale"],
                             M=16, near_frac=0.01)
    cams = []
    for (W, H, yaw) in c["views"]:
        th = math.radians(yaw)
        Rc2w = np.array([[math.cos(th), 0, math.sin(th)], [0, 1, 0], [-math.sin(th), 0, math.cos(th)]])
        C = Rc2w @ np.array([0.0, 0.0, -4.0])
        cams.append(synth.make_camera(W, H, Rc2w, -Rc2w.T @ C))
    xyz = scene.means3D
    knn = c["knn"]
    nb = torch.empty((c["P"], knn), dtype=torch.int32)
    for a in range(0, c["P"], 4096):
        d = torch.cdist(xyz[a:a + 4096], xyz)
        d[torch.arange(d.shape[0]), torch.arange(a, a + d.shape[0])] = float("inf")
        nb[a:a + 4096] = d.topk(knn, dim=1, largest=False).indices.to(torch.int32)
    return c, scene, cams, nb


def tools_camera_tensors(cams):
    """The stacked per-camera tensors exactly as gaussian_model.py:727-733 / scene/__init__.py:145-151 build them."""
    return dict(
        positions=torch.stack([c.camera_center for c in cams]),
        views=torch.stack([c.world_view_transform for c in cams]),
        projs=torch.stack([c.full_proj_transform for c in cams]),
        inv_projs=torch.stack([c.full_proj_transform.inverse() for c in cams]),
        tanx=torch.tensor([math.tan(c.FoVx * 0.5) for c in cams], dtype=torch.float32),
        tany=torch.tensor([math.tan(c.FoVy * 0.5) for c in cams], dtype=torch.float32),
        H=torch.tensor([c.image_height for c in cams], dtype=torch.int32),
        W=torch.tensor([c.image_width for c in cams], dtype=torch.int32))


# ---- codebook k-means --------------------------------------------------------------------------
# n is a multiple of 256 on purpose: in the reference's updateIdsCUDA the threads of the last, partial block return BEFORE they
# load their share of the centres into shared memory (kmeans.cu:81-91), so for n % 256 != 0 the last n % 256 ids are computed
# against partly uninitialised shared memory (observed on B200: 41 of the last 64 ids wrong for n = 40 000) — undefined
# behaviour that is not reproduced here; the tests compare those trailing values only against the oracle.
KMEANS_CASES = {"k1": dict(n=40_960, zeros=4_096, k=256, seed=31, tol=1e-4, max_iterations=500)}


def build_kmeans_inputs(name, n=None):
    """values ~ N(0,1) with a block of exact zeros and repeated values (culled SH coefficients are exact zeros in the reference),
    initial centres sampled from the values as generate_codebook does (gaussian_model.py:36-39)."""
    c = dict(KMEANS_CASES[name])
    if n is not None:
        c["n"], c["zeros"] = n, n // 10
    g = torch.Generator().manual_seed(c["seed"])
    v = torch.randn(c["n"], generator=g)
    v[torch.randperm(c["n"], generator=g)[: c["zeros"]]] = 0.0
    v[::7] = v[1::7][: v[::7].shape[0]]                       # exact duplicates
    v = v.view(-1, 1).contiguous()
    centers = v[torch.randint(c["n"], (c["k"],), generator=g)].view(-1).contiguous()
    return c, v, centers
