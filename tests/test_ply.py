"""CPU: reduced-3dgs PLY <-> device layout (gs_b200/ply.py; reference scene/gaussian_model.py:239-311 save_ply, :398-483 load_ply).
PINNED to the reference: tests/golden/ref_{quant,quant_half,fp32}.ply were written by the reference's own
`GaussianModel.save_ply` and ref_ply_expected.npz holds what its own `load_ply` / `_parse_vertex_group` read back
(tests/golden/make_golden_ply.py imports /root/reference/scene/gaussian_model.py; only the `plyfile` container writer, absent in
this image, is a stand-in).  Our reader must reproduce those tensors and our writer those bytes.  Further checks: header layout,
lossless (or exactly half-rounded) round trips, and the reference's de-quantisation restated with its own torch ops."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "reduced-3dgs_b200"))
from gs_b200 import ply, synth  # noqa: E402


def _model(P=500, seed=5):
    scene = synth.make_scene(P, seed, sh_degree=3, mixed_degrees=True, M=16)
    order = torch.argsort(scene.degrees.view(-1), stable=True)                 # the file groups Gaussians by degree
    scene = synth.Scene(*[getattr(scene, f)[order].contiguous() for f in ("means3D", "opacity", "scales", "rotations", "sh", "degrees")])
    return scene, synth.quantise_scene(scene, seed=0)


def test_header_and_layout(tmp_path):
    scene, q = _model(40)
    path = str(tmp_path / "point_cloud_quantised.ply")
    ply.save_reduced_ply(path, q)
    raw = open(path, "rb").read()
    head = raw[: raw.index(b"end_header\n") + 11].decode("ascii").split("\n")
    assert head[0] == "ply" and head[1] == "format binary_little_endian 1.0"
    counts = [int((q.degrees.view(-1) == d).sum()) for d in range(4)]
    want = []
    for d in range(4):
        want.append(f"element vertex_{d} {counts[d]}")
        want += ["property float x", "property float y", "property float z"]
        want += [f"property uchar f_dc_{i}" for i in range(3)] + [f"property uchar f_rest_{i}" for i in range(3 * ((d + 1) ** 2 - 1))]
        want += ["property uchar opacity"] + [f"property uchar scale_{i}" for i in range(3)] + [f"property uchar rot_{i}" for i in range(4)]
    want.append("element codebook_centers 256")
    want += [f"property float {n}" for n in ply.CODEBOOK_NAMES]
    assert head[2:-2] == want and head[-2] == "end_header"
    body = sum(c * (12 + 3 + 3 * ((d + 1) ** 2 - 1) + 1 + 3 + 4) for d, c in enumerate(counts)) + 256 * 20 * 4
    assert len(raw) == len("\n".join(head)) + body
    # bytes per Gaussian on disk / on the device: 35 at degree 3 ids + 12 xyz vs 59 floats = 236 B in the fp32 file
    el = ply.read_ply(path)
    assert list(el) == ["vertex_0", "vertex_1", "vertex_2", "vertex_3", "codebook_centers"]
    assert el["vertex_3"].dtype.itemsize == 12 + 3 + 45 + 1 + 3 + 4


@pytest.mark.parametrize("half", [False, True])
def test_quantised_round_trip_and_reference_dequant(tmp_path, half):
    scene, q = _model()
    path = str(tmp_path / "q.ply")
    ply.save_reduced_ply(path, q, half_float=half)
    r = ply.load_reduced_ply(path, half_float=half, quantised=True)
    assert torch.equal(r.degrees, q.degrees) and r.ids_dc.dtype == torch.uint8
    for f in ("ids_dc", "ids_opacity", "ids_scaling", "ids_rot"):
        assert torch.equal(getattr(r, f), getattr(q, f)), f
    ncoef = (q.degrees.view(-1, 1).long() + 1) ** 2 - 1
    active = (torch.arange(15).view(1, 15) < ncoef).unsqueeze(-1).expand(-1, -1, 3)
    assert torch.equal(r.ids_rest[active], q.ids_rest[active]) and int(r.ids_rest[~active].sum()) == 0     # padded with zero ids
    if half:
        assert torch.equal(r.centers, q.centers.half().float()) and torch.equal(r.means3D, q.means3D.half().float())
    else:
        assert torch.equal(r.centers, q.centers) and torch.equal(r.means3D, q.means3D)
    # the reference's de-quantisation (gaussian_model.py:371-387), written with its own ops on the loaded planes
    c = r.centers
    P = r.means3D.shape[0]
    features_dc = c[0][r.ids_dc.view(-1).long()].view(-1, 1, 3)
    rest_table = c[1:16].t().contiguous()                                      # [256,15] like codebook_centers_torch['features_rest']
    ids_cm = r.ids_rest.transpose(1, 2).reshape(P * 3, 15)                     # the reference keeps [P,3,15] channel-major
    features_rest = rest_table.gather(0, ids_cm.long()).view(P, 3, 15).transpose(1, 2)
    opacity = c[16][r.ids_opacity.long()].view(P, 1)
    scaling = c[17][r.ids_scaling.reshape(P * 3).long()].view(P, 3)
    rotation = torch.cat((c[18][r.ids_rot[:, 0:1].long()], c[19][r.ids_rot[:, 1:].reshape(P * 3).long()].view(P, 3)), dim=1)
    d = r.dequantise()
    assert torch.equal(d.sh[:, :1], features_dc) and torch.equal(d.opacity, opacity)
    assert torch.equal(d.sh[:, 1:][active], features_rest[active]) and float(d.sh[:, 1:][~active].abs().sum()) == 0.0
    assert torch.equal(d.scales, torch.exp(scaling)) and torch.equal(d.rotations, torch.nn.functional.normalize(rotation))


def test_fp32_round_trip(tmp_path):
    scene, _ = _model(300)
    path = str(tmp_path / "f.ply")
    ply.save_reduced_ply(path, scene)
    r = ply.load_reduced_ply(path, quantised=False)
    assert torch.equal(r.means3D, scene.means3D) and torch.equal(r.degrees, scene.degrees) and torch.equal(r.opacity, scene.opacity)
    ncoef = (scene.degrees.view(-1, 1).long() + 1) ** 2
    active = (torch.arange(16).view(1, 16) < ncoef).unsqueeze(-1).expand(-1, -1, 3)
    assert torch.equal(r.sh[active], scene.sh[active]) and float(r.sh[~active].abs().sum()) == 0.0
    assert torch.allclose(r.scales, scene.scales, rtol=3e-7, atol=0) and torch.allclose(r.rotations, scene.rotations, atol=1e-7)


def test_reader_rejects_what_the_format_does_not_use(tmp_path):
    p = tmp_path / "bad.ply"
    p.write_bytes(b"ply\nformat ascii 1.0\nelement vertex_0 0\nend_header\n")
    with pytest.raises(ValueError):
        ply.read_ply(str(p))
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n\x00")
    with pytest.raises(ValueError):
        ply.read_ply(str(p))
    p.write_bytes(b"plx\n")
    with pytest.raises(ValueError):
        ply.read_ply(str(p))


def test_model_view_from_ply(tmp_path):
    """gs_b200.model.GaussianModelView: the attributes gaussian_renderer.render() reads, built straight from a quantised PLY."""
    from gs_b200.model import GaussianModelView
    scene, q = _model(200)
    path = str(tmp_path / "q.ply")
    ply.save_reduced_ply(path, q)
    v = GaussianModelView.from_ply(path, quantised=True, device="cpu")
    assert v.quant is not None and v.quant.ids_rest.dtype == torch.uint8 and v.num_primitives == 200
    assert v.per_band_count == [int((scene.degrees == d).sum()) for d in range(4)] and sum(v.per_band_count) == 200
    assert tuple(v.get_features.shape) == (200, 16, 3) and tuple(v._opacity.shape) == (200, 1) and v._degrees.dtype == torch.int32
    assert torch.equal(v.get_xyz, q.means3D) and not v.get_xyz.requires_grad
    d = q.dequantise()
    assert torch.equal(v.get_scaling, d.scales) and torch.equal(v.get_rotation, d.rotations)
    t = GaussianModelView(scene, "cpu")                       # trainable fp32 view
    assert all(p.requires_grad for p in t.params()) and t.quant is None


GOLD = os.path.join(HERE, "golden")


@pytest.mark.parametrize("tag,quantised,half", [("quant", True, False), ("quant_half", True, True), ("fp32", False, False)])
def test_files_written_by_the_reference(tmp_path, tag, quantised, half):
    """Bytes written by the reference's save_ply -> our loader == what the reference's load_ply produced from the same bytes;
    and for the quantised layouts our writer reproduces the reference's file byte for byte."""
    path = os.path.join(GOLD, f"ref_{tag}.ply")
    exp = {k[len(tag) + 1:]: v for k, v in np.load(os.path.join(GOLD, "ref_ply_expected.npz")).items() if k.startswith(tag + "_")}
    m = ply.load_reduced_ply(path, half_float=half, quantised=quantised)
    P = exp["xyz"].shape[0]
    deg = torch.from_numpy(exp["degrees"]).view(-1)
    assert torch.equal(m.means3D, torch.from_numpy(exp["xyz"])) and torch.equal(m.degrees.view(-1), deg)
    ncoef = (deg.view(-1, 1).long() + 1) ** 2 - 1
    active = (torch.arange(15).view(1, 15) < ncoef).unsqueeze(-1).expand(-1, -1, 3)
    if quantised:
        assert isinstance(m, synth.QuantScene) and m.ids_rest.dtype == torch.uint8
        d = m.dequantise()                                      # scales exp-activated, rotations normalised, inactive bands zeroed
        # raw (pre-activation) values exactly as _parse_vertex_group de-quantises them (gaussian_model.py:371-387)
        c = m.centers
        assert torch.equal(c[17][m.ids_scaling.long()], torch.from_numpy(exp["scaling"]))
        rot_raw = torch.cat((c[18][m.ids_rot[:, 0:1].long()], c[19][m.ids_rot[:, 1:].long()]), dim=1)
        assert torch.equal(rot_raw, torch.from_numpy(exp["rotation"]))
        # the reference pads the ids of coefficients beyond a group's degree with 0 (:352-356), i.e. its tensor holds centre 0 there;
        # the kernels never read those coefficients and dequantise() zeroes them (GM:726 semantics)
        assert int(m.ids_rest[~active].sum()) == 0
    else:
        d = m
    assert torch.equal(d.sh[:, :1], torch.from_numpy(exp["features_dc"]))
    assert torch.equal(d.sh[:, 1:][active], torch.from_numpy(exp["features_rest"])[active])
    assert torch.equal(d.opacity, torch.from_numpy(exp["opacity"]).view(P, 1))
    assert torch.equal(d.scales, torch.from_numpy(exp["get_scaling"]))
    assert torch.equal(d.rotations, torch.from_numpy(exp["get_rotation"]))
    if quantised:
        out = str(tmp_path / "rewritten.ply")
        ply.save_reduced_ply(out, m, half_float=half)
        assert open(out, "rb").read() == open(path, "rb").read(), "our writer must reproduce the reference-written file byte for byte"
