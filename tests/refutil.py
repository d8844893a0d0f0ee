"""Helpers for talking to the compiled reference (oracle/_ref/_refC.so) in GPU tests — test infrastructure.

decode_* follow the reference's private blob layouts (rasterizer_impl.cu:163-202 fromChunk / rasterizer_impl.h:21-74):
every array is bump-allocated at the next 128-byte boundary.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def load_ref():
    import build_ref
    return build_ref.load()


def _carve(buf: np.ndarray, off: int, dtype, count: int):
    off = (off + 127) & ~127
    nbytes = np.dtype(dtype).itemsize * count
    return np.frombuffer(buf[off:off + nbytes].tobytes(), dtype=dtype), off + nbytes


def decode_geom(geomBuffer: torch.Tensor, P: int):
    b = geomBuffer.detach().cpu().numpy()
    assert geomBuffer.data_ptr() % 128 == 0
    out, off = {}, 0
    out["depths"], off = _carve(b, off, np.float32, P)
    cl, off = _carve(b, off, np.uint8, 3 * P)
    out["clamped"] = cl.reshape(P, 3)
    out["internal_radii"], off = _carve(b, off, np.int32, P)
    m, off = _carve(b, off, np.float32, 2 * P)
    out["means2D"] = m.reshape(P, 2)
    c, off = _carve(b, off, np.float32, 6 * P)
    out["cov3D"] = c.reshape(P, 6)
    co, off = _carve(b, off, np.float32, 4 * P)
    out["conic_opacity"] = co.reshape(P, 4)
    r, off = _carve(b, off, np.float32, 3 * P)
    out["rgb"] = r.reshape(P, 3)
    out["tiles_touched"], off = _carve(b, off, np.uint32, P)
    return out


def decode_binning(binningBuffer: torch.Tensor, R: int):
    b = binningBuffer.detach().cpu().numpy()
    out, off = {}, 0
    out["point_list"], off = _carve(b, off, np.uint32, R)
    out["point_list_unsorted"], off = _carve(b, off, np.uint32, R)
    out["keys"], off = _carve(b, off, np.uint64, R)
    out["keys_unsorted"], off = _carve(b, off, np.uint64, R)
    return out


def decode_image(imgBuffer: torch.Tensor, W: int, H: int):
    b = imgBuffer.detach().cpu().numpy()
    N = W * H
    out, off = {}, 0
    a, off = _carve(b, off, np.float32, N)
    out["final_T"] = a.reshape(H, W)
    n, off = _carve(b, off, np.uint32, N)
    out["n_contrib"] = n.reshape(H, W)
    r, off = _carve(b, off, np.uint32, 2 * N)
    ntiles = ((W + 15) // 16) * ((H + 15) // 16)
    out["ranges"] = r.reshape(N, 2)[:ntiles].copy()
    return out


EMPTY = torch.Tensor([])


def ref_forward(refC, scene, cam, bg, extra=None, debug=False, dev="cuda"):
    """_C.rasterize_gaussians with the reference's positional signature (rasterize_points.h:43-63)."""
    import math
    extra = extra or {}
    cov = extra.get("cov3D_precomp")
    col = extra.get("colors_precomp")
    args = (bg.to(dev), scene.means3D.to(dev), EMPTY if col is None else col.to(dev), scene.opacity.to(dev),
            EMPTY if cov is not None else scene.scales.to(dev), EMPTY if cov is not None else scene.rotations.to(dev), 1.0,
            EMPTY if cov is None else cov.to(dev), cam.world_view_transform.to(dev), cam.full_proj_transform.to(dev),
            math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), cam.image_height, cam.image_width,
            EMPTY if col is not None else scene.sh.to(dev), scene.degrees.to(dev), cam.camera_center.to(dev), False, debug)
    return args, refC.rasterize_gaussians(*args)


def ref_backward(refC, fwd_args, fwd_out, dL, lam=0.0, debug=False):
    """_C.rasterize_gaussians_backward (rasterize_points.h:65-88)."""
    (bg, means3D, colors, opacity, scales, rotations, mod, cov, view, proj, tx, ty, H, W, sh, degrees, campos, _, _) = fwd_args
    R, color, radii, geom, binning, img = fwd_out
    return refC.rasterize_gaussians_backward(bg, means3D, radii, colors, scales, rotations, mod, cov, view, proj, tx, ty,
                                             dL.to(means3D.device), sh, degrees, campos, geom, R, binning, img, lam, debug)
