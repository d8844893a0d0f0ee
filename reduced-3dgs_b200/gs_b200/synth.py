"""Seeded synthetic Gaussian clouds and cameras for parity tests and bench.py.

Everything is generated on the CPU with an explicit ``torch.Generator`` so that
this container, the GPU box and the CPU oracle all see identical bits
(SURVEY.md §8(d)).  Matrix conventions reproduce the reference's camera code:
``world_view_transform`` / ``full_proj_transform`` are stored TRANSPOSED
(row-vector convention), reference scene/cameras.py:54-58 and
utils/graphics_utils.py:38-71.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

# Order of the 20 reduced-3dgs codebooks (reference README.md:132-150,
# scene/gaussian_model.py:252-272): features_dc, features_rest_0..14, opacity,
# scaling, rotation_re, rotation_im.
CB_FEATURES_DC = 0
CB_FEATURES_REST0 = 1
CB_OPACITY = 16
CB_SCALING = 17
CB_ROT_RE = 18
CB_ROT_IM = 19
NUM_CODEBOOKS = 20


@dataclass
class Camera:
    """Mirror of the reference MiniCam (scene/cameras.py:60-72)."""
    image_width: int
    image_height: int
    FoVx: float
    FoVy: float
    znear: float
    zfar: float
    world_view_transform: torch.Tensor   # [4,4] transposed view matrix
    full_proj_transform: torch.Tensor    # [4,4] transposed (view @ proj)
    camera_center: torch.Tensor          # [3]

    def to(self, device):
        return Camera(self.image_width, self.image_height, self.FoVx, self.FoVy, self.znear, self.zfar,
                      self.world_view_transform.to(device), self.full_proj_transform.to(device),
                      self.camera_center.to(device))


def _world2view2(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    # utils/graphics_utils.py:38-49 with translate=0, scale=1 (two inversions kept:
    # they are part of how the reference rounds the matrix to float32).
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def _projection(znear: float, zfar: float, fovX: float, fovY: float) -> torch.Tensor:
    # utils/graphics_utils.py:51-71
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = -top
    right = tanHalfFovX * znear
    left = -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(width: int, height: int, R: Optional[np.ndarray] = None, T: Optional[np.ndarray] = None,
                fovy_deg: float = 50.0, znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """Pinhole camera; R is camera-to-world rotation, T the COLMAP translation
    (scene/cameras.py:54-58).  Default: R=I, T=(0,0,4) -> camera at z=-4 looking +z."""
    if R is None:
        R = np.eye(3)
    if T is None:
        T = np.array([0.0, 0.0, 4.0])
    fovy = math.radians(fovy_deg)
    fovx = 2.0 * math.atan(math.tan(fovy / 2) * width / height)
    wvt = torch.tensor(_world2view2(R, T)).transpose(0, 1).contiguous()
    proj = _projection(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return Camera(width, height, fovx, fovy, znear, zfar, wvt, full, center)


def orbit_cameras(n: int, width: int, height: int, radius: float = 6.0, fovy_deg: float = 50.0) -> List[Camera]:
    """n cameras on a circle in the xz-plane looking at the origin, +y up."""
    cams = []
    for i in range(n):
        th = 2.0 * math.pi * i / n
        C = np.array([radius * math.sin(th), 0.0, -radius * math.cos(th)])
        # camera-to-world rotation: columns are camera x,y,z axes in world coords
        zc = -C / np.linalg.norm(C)
        yc = np.array([0.0, 1.0, 0.0])
        xc = np.cross(yc, zc)
        xc /= np.linalg.norm(xc)
        Rc2w = np.stack([xc, yc, zc], axis=1)
        T = -Rc2w.T @ C
        cams.append(make_camera(width, height, Rc2w, T, fovy_deg))
    return cams


@dataclass
class Scene:
    """Reference tensor contracts: all fp32 contiguous."""
    means3D: torch.Tensor     # [P,3]
    opacity: torch.Tensor     # [P,1] RAW logits (sigmoid is in-kernel)
    scales: torch.Tensor      # [P,3] exp-activated
    rotations: torch.Tensor   # [P,4] (r,x,y,z) normalised
    sh: torch.Tensor          # [P,M,3]
    degrees: torch.Tensor     # [P,1] int32
    extras: Dict[str, torch.Tensor] = field(default_factory=dict)

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])

    def to(self, device):
        return Scene(self.means3D.to(device), self.opacity.to(device), self.scales.to(device),
                     self.rotations.to(device), self.sh.to(device), self.degrees.to(device),
                     {k: v.to(device) for k, v in self.extras.items()})

    def compact(self, keep: torch.Tensor) -> "Scene":
        """Physically delete rows (reference prune_points semantics, gaussian_model.py:553-563)."""
        return Scene(self.means3D[keep].contiguous(), self.opacity[keep].contiguous(),
                     self.scales[keep].contiguous(), self.rotations[keep].contiguous(),
                     self.sh[keep].contiguous(), self.degrees[keep].contiguous())

    def packed_sh(self):
        """Packed per-degree SH layout of the variable-SH inference path (gaussian_renderer/__init__.py:85,90-92).
        Requires Gaussians ordered by degree. Returns (flat, per_band_count[4], cumsum[4], coeffs_num[4])."""
        deg = self.degrees.view(-1)
        assert bool((deg[1:] >= deg[:-1]).all()), "packed SH needs degree-sorted Gaussians"
        counts = [int((deg == d).sum()) for d in range(4)]
        chunks, start = [], 0
        for d in range(4):
            k = (d + 1) ** 2
            chunks.append(self.sh[start:start + counts[d], :k, :].reshape(-1))
            start += counts[d]
        flat = torch.cat(chunks).contiguous()
        pbc = torch.tensor(counts, dtype=torch.int32)
        return flat, pbc, torch.cumsum(pbc, 0).to(torch.int32), torch.tensor([1, 4, 9, 16], dtype=torch.int32)


def make_scene(P: int, seed: int, sh_degree: int = 3, mixed_degrees: bool = False,
               box=(3.6, 2.0, 1.0), log_scale_mean: float = math.log(0.006), M: Optional[int] = None,
               near_frac: float = 0.001) -> Scene:
    """Synthetic cloud of SURVEY §8(d). ``mixed_degrees`` draws degrees with
    P(0,1,2,3)=(0.50,0.20,0.15,0.15) and sorts Gaussians by degree (reduced-3dgs banding)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(P, 3, generator=g)
    xyz = (u * 2.0 - 1.0) * torch.tensor(box)
    n_near = int(round(P * near_frac))
    if n_near > 0:
        idx = torch.randperm(P, generator=g)[:n_near]
        xyz[idx, 2] = -4.5 + 0.7 * torch.rand(n_near, generator=g)
    scales = torch.exp(log_scale_mean + 0.6 * torch.randn(P, 3, generator=g))
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    opacity = 2.0 * torch.randn(P, 1, generator=g)
    if M is None:
        M = (sh_degree + 1) ** 2
    sh = torch.empty(P, M, 3)
    sh[:, 0, :] = torch.randn(P, 3, generator=g)
    if M > 1:
        sh[:, 1:, :] = 0.15 * torch.randn(P, M - 1, 3, generator=g)
    if mixed_degrees:
        r = torch.rand(P, generator=g)
        deg = (r >= 0.5).int() + (r >= 0.7).int() + (r >= 0.85).int()
        deg, _ = torch.sort(deg)
        ncoef = (deg + 1) ** 2
        mask = torch.arange(M).view(1, M) >= ncoef.view(P, 1)
        sh[mask.unsqueeze(-1).expand(P, M, 3)] = 0.0      # unused coefficients zero (gaussian_model.py:726)
    else:
        deg = torch.full((P,), sh_degree, dtype=torch.int32)
    return Scene(xyz.contiguous(), opacity.contiguous(), scales.contiguous(), q.contiguous(), sh.contiguous(),
                 deg.to(torch.int32).view(P, 1).contiguous())


# ---- the named benchmark configurations (BASELINE.json `configs`) ----------------------------

def config_scene(name: str, P_override: Optional[int] = None) -> Scene:
    if name == "C1":
        return make_scene(P_override or 10_000, 1, sh_degree=0, box=(1.9, 1.9, 1.0), log_scale_mean=math.log(0.02))
    if name == "C2":
        return make_scene(P_override or 500_000, 2, sh_degree=3)
    if name in ("C3", "C4"):
        return make_scene(P_override or 3_000_000, 3, sh_degree=3, mixed_degrees=True)
    if name == "C5":
        return make_scene(P_override or 6_000_000, 5, sh_degree=3, mixed_degrees=True)
    raise ValueError(name)


def config_image(name: str):
    return {"C1": (256, 256), "C2": (1920, 1080), "C3": (1920, 1080), "C4": (1920, 1080), "C5": (3840, 2160)}[name]


def grad_image(W: int, H: int, seed: int) -> torch.Tensor:
    """dL/dcolor for backward: N(0,1) [3,H,W]."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(3, H, W, generator=g)


# ---- codebook quantisation (reduced-3dgs layout) ---------------------------------------------

@dataclass
class QuantScene:
    """u8 id planes + 20x256 codebook table; xyz/degrees stay fp32/int32 (position is not quantised,
    gaussian_model.py:285)."""
    means3D: torch.Tensor      # [P,3] f32
    degrees: torch.Tensor      # [P,1] i32
    ids_dc: torch.Tensor       # [P,3]  u8  (codebook 0)
    ids_rest: torch.Tensor     # [P,15,3] u8 (codebook 1+k for coefficient k, shared by RGB)
    ids_opacity: torch.Tensor  # [P]    u8  (centres are pre-activation logits)
    ids_scaling: torch.Tensor  # [P,3]  u8  (centres are log-scales)
    ids_rot: torch.Tensor      # [P,4]  u8  (col 0 -> rotation_re, cols 1-3 -> rotation_im)
    centers: torch.Tensor      # [20,256] f32

    def to(self, device):
        return QuantScene(*[getattr(self, f).to(device) for f in
                            ("means3D", "degrees", "ids_dc", "ids_rest", "ids_opacity", "ids_scaling", "ids_rot", "centers")])

    def dequantise(self) -> Scene:
        """Reference-equivalent fp32 expansion: centers[ids] exactly as gaussian_model.py:371-387, then
        exp / normalize as get_scaling / get_rotation (gaussian_model.py:141-146)."""
        c = self.centers
        P = self.means3D.shape[0]
        f_dc = c[CB_FEATURES_DC][self.ids_dc.long()].view(P, 1, 3)
        k = torch.arange(15, device=c.device).view(1, 15, 1).expand(P, 15, 3)
        f_rest = c[CB_FEATURES_REST0:CB_FEATURES_REST0 + 15][k, self.ids_rest.long()]
        ncoef = (self.degrees.view(P, 1).long() + 1) ** 2
        inactive = torch.arange(1, 16, device=c.device).view(1, 15) >= ncoef
        f_rest = torch.where(inactive.unsqueeze(-1), torch.zeros_like(f_rest), f_rest)
        sh = torch.cat([f_dc, f_rest], dim=1).contiguous()
        opacity = c[CB_OPACITY][self.ids_opacity.long()].view(P, 1)
        scales = torch.exp(c[CB_SCALING][self.ids_scaling.long()])
        rot = torch.cat([c[CB_ROT_RE][self.ids_rot[:, 0:1].long()], c[CB_ROT_IM][self.ids_rot[:, 1:].long()]], dim=1)
        rot = torch.nn.functional.normalize(rot)
        return Scene(self.means3D, opacity.contiguous(), scales.contiguous(), rot.contiguous(), sh, self.degrees)


def _nearest_ids(values: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    mids = (centers[1:] + centers[:-1]) * 0.5
    return torch.bucketize(values.contiguous(), mids).to(torch.uint8)


def _quantile_centers(values: torch.Tensor, g: torch.Generator) -> torch.Tensor:
    v = values.reshape(-1)
    if v.numel() > 1_000_000:
        v = v[torch.randint(v.numel(), (1_000_000,), generator=g)]
    v, _ = torch.sort(v)
    q = ((torch.arange(256, dtype=torch.float64) + 0.5) / 256 * (v.numel() - 1)).long()
    c = v[q].clone()
    # strictly increasing centres so bucketize is a nearest-centre search
    for i in range(1, 256):
        if c[i] <= c[i - 1]:
            c[i] = torch.nextafter(c[i - 1], torch.tensor(float("inf")))
    return c


def quantise_scene(scene: Scene, seed: int = 0) -> QuantScene:
    """256 quantile centres per attribute, nearest-centre ids."""
    assert scene.sh.shape[1] == 16
    g = torch.Generator().manual_seed(seed)
    P = scene.P
    centers = torch.zeros(NUM_CODEBOOKS, 256)
    f_dc = scene.sh[:, 0, :]
    centers[CB_FEATURES_DC] = _quantile_centers(f_dc, g)
    ids_dc = _nearest_ids(f_dc, centers[CB_FEATURES_DC])
    ids_rest = torch.zeros(P, 15, 3, dtype=torch.uint8)
    ncoef = (scene.degrees.view(-1).long() + 1) ** 2
    for k in range(15):
        active = ncoef > (k + 1)
        vals = scene.sh[:, k + 1, :]
        src = vals[active] if bool(active.any()) else vals
        centers[CB_FEATURES_REST0 + k] = _quantile_centers(src, g)
        ids_rest[:, k, :] = _nearest_ids(vals, centers[CB_FEATURES_REST0 + k])
    centers[CB_OPACITY] = _quantile_centers(scene.opacity, g)
    ids_op = _nearest_ids(scene.opacity.view(-1), centers[CB_OPACITY])
    logs = torch.log(scene.scales)
    centers[CB_SCALING] = _quantile_centers(logs, g)
    ids_sc = _nearest_ids(logs, centers[CB_SCALING])
    centers[CB_ROT_RE] = _quantile_centers(scene.rotations[:, 0], g)
    centers[CB_ROT_IM] = _quantile_centers(scene.rotations[:, 1:], g)
    ids_rot = torch.cat([_nearest_ids(scene.rotations[:, 0:1], centers[CB_ROT_RE]),
                         _nearest_ids(scene.rotations[:, 1:], centers[CB_ROT_IM])], dim=1)
    return QuantScene(scene.means3D, scene.degrees, ids_dc.contiguous(), ids_rest.contiguous(), ids_op.contiguous(),
                      ids_sc.contiguous(), ids_rot.contiguous(), centers.contiguous())


def prune_mask(P: int, seed: int, frac: float = 0.5) -> torch.Tensor:
    """Bernoulli(frac) prune mask, 1 = pruned."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(P, generator=g) < frac).to(torch.uint8)
