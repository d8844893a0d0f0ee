"""Reduced-3DGS PLY  <->  device layout of the fused rasterizer.

The reference stores a model as a binary little-endian PLY with one `vertex_<d>` element per SH degree d = 0..3 (Gaussians are
grouped by degree; a group carries only its 3*((d+1)^2 - 1) `f_rest_*` coefficients, channel-major: rrr ggg bbb) and, when
quantised, a final `codebook_centers` element of 256 rows x 20 codebooks (scene/gaussian_model.py:239-311 save_ply, :398-483
load_ply, README.md:76-165).  Attributes are `u1` codebook ids when quantised; floats are `f4`, or IEEE halves stored in `int16`
properties when half_float (plyfile has no float16).

The reference loader expands everything to fp32 tensors (`centers[ids]`, gather) before the rasterizer sees it.  Here the u8 id
planes, the 20x256 centre table and xyz go to the device as they are (35 B per Gaussian instead of 248 B) and the
fused preprocess kernel de-quantises on the fly (`quant=` of diff_gaussian_rasterization); `QuantScene.dequantise()` gives the
reference-equivalent fp32 tensors when they are needed.

`plyfile` is not a dependency: the small reader / writer below handles exactly what the format uses (binary_little_endian,
scalar properties).  Parity note: the reference's own reader cannot run where plyfile is absent, so this file is pinned to the
format as written in gaussian_model.py and to round trips, not to bytes produced by the reference.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple, Union

import numpy as np
import torch

from . import synth

# PLY scalar type names (as plyfile writes them) <-> numpy codes; the aliases on the right are accepted when reading
_PLY_OF_NP = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}
_NP_OF_PLY = {v: k for k, v in _PLY_OF_NP.items()}
_NP_OF_PLY.update({"int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"})

CODEBOOK_NAMES = ["features_dc"] + [f"features_rest_{i}" for i in range(15)] + ["opacity", "scaling", "rotation_re", "rotation_im"]


def write_ply(path: str, elements: List[Tuple[str, np.ndarray]]) -> None:
    """elements: (name, structured array with scalar fields). Binary little-endian, header as plyfile writes it."""
    head = ["ply", "format binary_little_endian 1.0"]
    for name, arr in elements:
        head.append(f"element {name} {arr.shape[0]}")
        for field in arr.dtype.names:
            dt = arr.dtype.fields[field][0]
            head.append(f"property {_PLY_OF_NP[dt.kind + str(dt.itemsize)]} {field}")
    head.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode("ascii"))
        for _, arr in elements:
            f.write(np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("<"), copy=False)).tobytes())


def read_ply(path: str) -> "OrderedDict[str, np.ndarray]":
    """-> {element name: structured array}, in file order."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.find(b"end_header")
    if not data.startswith(b"ply") or end < 0:
        raise ValueError(f"{path}: not a PLY file")
    body = data.find(b"\n", end) + 1
    lines = data[:end].decode("ascii").split("\n")
    fmt, layout = None, []
    for ln in lines[1:]:
        tok = ln.split()
        if not tok or tok[0] == "comment" or tok[0] == "obj_info":
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            layout.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                raise ValueError(f"{path}: list properties are not part of the reduced-3dgs format")
            layout[-1][2].append((tok[2], "<" + _NP_OF_PLY[tok[1]]))
        else:
            raise ValueError(f"{path}: unexpected header line {ln!r}")
    if fmt != "binary_little_endian":
        raise ValueError(f"{path}: only binary_little_endian is supported (got {fmt})")
    out = OrderedDict()
    off = body
    for name, count, props in layout:
        dt = np.dtype(props)
        out[name] = np.frombuffer(data, dtype=dt, count=count, offset=off)
        off += dt.itemsize * count
    if off != len(data):
        raise ValueError(f"{path}: {len(data) - off} trailing bytes")
    return out


def _attributes(rest_coeffs: int) -> List[str]:
    """gaussian_model.py:231-237 construct_list_of_attributes."""
    return (["x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(rest_coeffs)] +
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])


def _f16_bits(a: np.ndarray) -> np.ndarray:
    return a.astype(np.float16).view(np.int16)


def save_reduced_ply(path: str, model: Union[synth.QuantScene, synth.Scene], half_float: bool = False) -> None:
    """gaussian_model.py:239-311 save_ply(path, quantised=isinstance(model, QuantScene), half_float).
    For a fp32 `Scene` the stored attributes are the reference's raw parameters: opacity logits, LOG scales, rotations as given,
    f_dc / f_rest split from the [P,16,3] SH tensor."""
    quantised = isinstance(model, synth.QuantScene)
    ftype = "<i2" if half_float else "<f4"
    atype = "u1" if quantised else ftype
    conv = (lambda a: _f16_bits(a)) if half_float else (lambda a: a.astype(np.float32))
    xyz = model.means3D.detach().cpu().numpy()
    deg = model.degrees.detach().cpu().numpy().reshape(-1)
    if quantised:
        f_dc = model.ids_dc.cpu().numpy()                                  # [P,3]
        f_rest = model.ids_rest.cpu().numpy()                              # [P,15,3]
        opacity = model.ids_opacity.cpu().numpy().reshape(-1, 1)
        scale = model.ids_scaling.cpu().numpy()
        rot = model.ids_rot.cpu().numpy()
    else:
        sh = model.sh.detach().cpu().numpy()
        f_dc, f_rest = conv(sh[:, 0, :]), conv(sh[:, 1:, :])
        opacity = conv(model.opacity.detach().cpu().numpy().reshape(-1, 1))
        scale = conv(np.log(model.scales.detach().cpu().numpy()))
        rot = conv(model.rotations.detach().cpu().numpy())
    elements = []
    for d in range(4):
        coeffs = (d + 1) ** 2 - 1
        m = deg == d
        names = _attributes(coeffs * 3)
        dt = np.dtype([(n, ftype if n in ("x", "y", "z") else atype) for n in names])
        el = np.empty(int(m.sum()), dtype=dt)
        # rest features are saved rrr ggg bbb (gaussian_model.py:292-296: transpose(1, 2).flatten)
        rest = np.transpose(f_rest[m][:, :coeffs, :], (0, 2, 1)).reshape(int(m.sum()), coeffs * 3)
        cols = np.concatenate([np.zeros((int(m.sum()), 3), f_dc.dtype), f_dc[m], rest, opacity[m], scale[m], rot[m]], axis=1)
        for i, n in enumerate(names):
            el[n] = conv(xyz[m][:, i]) if i < 3 else cols[:, i]
        elements.append((f"vertex_{d}", el))
    if quantised:
        centers = model.centers.detach().cpu().numpy()                     # [20,256]
        dt = np.dtype([(n, ftype) for n in CODEBOOK_NAMES])
        cb = np.empty(256, dtype=dt)
        for k, n in enumerate(CODEBOOK_NAMES):
            cb[n] = conv(centers[k])
        elements.append(("codebook_centers", cb))
    write_ply(path, elements)


def _floats(col: np.ndarray, half_float: bool) -> np.ndarray:
    """pcast_i16_to_f32 of the reference: reinterpret the int16 payload as IEEE half, widen to fp32."""
    return col.view(np.float16).astype(np.float32) if half_float else col.astype(np.float32)


def load_reduced_ply(path: str, half_float: bool = False, quantised: bool = True, device="cpu") -> Union[synth.QuantScene, synth.Scene]:
    """gaussian_model.py:398-483 load_ply.  quantised -> QuantScene with the u8 id planes exactly as stored (no fp32 expansion;
    the ids of coefficients beyond a group's degree are 0, as the reference pads them, :352-356); otherwise a fp32 Scene with the
    reference's activations applied (sigmoid stays in the rasterizer, scales = exp, rotations normalised, get_* of :141-158)."""
    el = read_ply(path)
    groups = [el[f"vertex_{d}"] for d in range(4)]
    counts = [g.shape[0] for g in groups]
    P = int(sum(counts))
    xyz = np.concatenate([np.stack([_floats(g["x"], half_float), _floats(g["y"], half_float), _floats(g["z"], half_float)], axis=1)
                          for g in groups], axis=0) if P else np.zeros((0, 3), np.float32)
    degrees = np.concatenate([np.full((c, 1), d, np.int32) for d, c in enumerate(counts)], axis=0)

    def stack(g, name, n):
        return np.stack([g[f"{name}_{i}"] for i in range(n)], axis=1) if g.shape[0] else np.zeros((0, n), g.dtype[f"{name}_0"])

    f_dc, f_rest, opac, scale, rot = [], [], [], [], []
    for d, g in enumerate(groups):
        coeffs = (d + 1) ** 2 - 1
        n = g.shape[0]
        f_dc.append(stack(g, "f_dc", 3))
        if coeffs:
            r = stack(g, "f_rest", coeffs * 3).reshape(n, 3, coeffs)                      # channel-major in the file
        else:
            r = np.zeros((n, 3, 0), g.dtype["f_dc_0"])
        r = np.concatenate([r, np.zeros((n, 3, 15 - coeffs), r.dtype)], axis=2)          # padded with zeros (:352-356)
        f_rest.append(np.transpose(r, (0, 2, 1)))                                        # -> [n,15,3] coefficient-major
        opac.append(g["opacity"].reshape(n, 1))
        scale.append(stack(g, "scale", 3))
        rot.append(stack(g, "rot", 4))
    cat = lambda xs: np.ascontiguousarray(np.concatenate(xs, axis=0))
    f_dc, f_rest, opac, scale, rot = cat(f_dc), cat(f_rest), cat(opac), cat(scale), cat(rot)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    if quantised:
        cb = el["codebook_centers"]
        centers = np.stack([_floats(cb[n], half_float) for n in CODEBOOK_NAMES], axis=0)   # [20,256]
        return synth.QuantScene(t(xyz), t(degrees), t(f_dc), t(f_rest), t(opac.reshape(-1)), t(scale), t(rot), t(centers))
    fl = lambda a: _floats(a, half_float)
    sh = np.concatenate([fl(f_dc).reshape(P, 1, 3), fl(f_rest)], axis=1)
    rotn = torch.nn.functional.normalize(torch.from_numpy(fl(rot)))
    return synth.Scene(t(xyz), t(fl(opac)), torch.exp(torch.from_numpy(fl(scale))).to(device), rotn.to(device), t(sh), t(degrees))
