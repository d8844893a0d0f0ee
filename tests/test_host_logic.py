"""CPU: host-side logic — synthetic inputs, matrix conventions, codebook layout, API argument contracts."""
import math

import numpy as np
import pytest
import torch

from gs_b200 import synth


def test_scene_is_deterministic_and_well_formed():
    a, b = synth.make_scene(5000, 7, mixed_degrees=True), synth.make_scene(5000, 7, mixed_degrees=True)
    for f in ("means3D", "opacity", "scales", "rotations", "sh", "degrees"):
        assert torch.equal(getattr(a, f), getattr(b, f))
    assert torch.allclose(a.rotations.norm(dim=1), torch.ones(5000), atol=1e-6)
    d = a.degrees.view(-1)
    assert bool((d[1:] >= d[:-1]).all()), "reduced-3dgs banding: Gaussians grouped by degree"
    k = (d.long() + 1) ** 2
    for i in (0, 2500, 4999):
        assert not a.sh[i, int(k[i]):].any(), "unused coefficients are zero (gaussian_model.py:726)"
    assert a.means3D[:, 2].min() < -3.8, "near-plane cull branch is exercised"


def test_camera_conventions_match_reference():
    cam = synth.make_camera(1920, 1080)
    # camera at z=-4 looking +z; matrices stored transposed (scene/cameras.py:54-58)
    assert torch.allclose(cam.camera_center, torch.tensor([0.0, 0.0, -4.0]), atol=1e-6)
    v = cam.world_view_transform
    p = torch.tensor([0.5, -0.25, 1.0, 1.0]) @ v            # row-vector convention
    assert torch.allclose(p[:3], torch.tensor([0.5, -0.25, 5.0]), atol=1e-6)
    assert math.isclose(math.tan(cam.FoVx / 2) / math.tan(cam.FoVy / 2), 1920 / 1080, rel_tol=1e-6)
    h = torch.tensor([0.0, 0.0, 1.0, 1.0]) @ cam.full_proj_transform
    assert abs(float(h[0])) < 1e-6 and abs(float(h[3]) - 5.0) < 1e-5      # w = view-space depth


def test_orbit_cameras_look_at_origin():
    for cam in synth.orbit_cameras(8, 640, 360):
        o = torch.tensor([0.0, 0.0, 0.0, 1.0]) @ cam.world_view_transform
        assert abs(float(o[0])) < 1e-5 and abs(float(o[1])) < 1e-5 and abs(float(o[2]) - 6.0) < 1e-4


def test_quantise_layout_and_dequantise():
    s = synth.make_scene(3000, 3, mixed_degrees=True)
    q = synth.quantise_scene(s)
    assert q.centers.shape == (20, 256) and q.ids_rest.shape == (3000, 15, 3) and q.ids_rot.shape == (3000, 4)
    assert q.ids_dc.dtype == torch.uint8
    d = q.dequantise()
    # reference-equivalent gather (gaussian_model.py:371-387)
    assert torch.equal(d.opacity.view(-1), q.centers[16][q.ids_opacity.long()])
    assert torch.equal(d.sh[:, 0, :], q.centers[0][q.ids_dc.long()])
    assert torch.allclose(d.rotations.norm(dim=1), torch.ones(3000), atol=1e-6)
    assert float(((d.scales - s.scales).abs() / s.scales).median()) < 0.02      # 256 quantile centres: ~1% typical error
    k = (s.degrees.view(-1).long() + 1) ** 2
    assert not d.sh[0, int(k[0]):].any()


def test_packed_sh_layout_matches_getSHOffset():
    s = synth.make_scene(2000, 5, mixed_degrees=True)
    flat, pbc, cs, cn = s.packed_sh()
    assert cn.tolist() == [1, 4, 9, 16] and int(pbc.sum()) == 2000
    # forward.cu:19-36
    def offset(idx):
        off = 0
        for d in range(4):
            if idx < int(cs[d]):
                first = 0 if d == 0 else int(cs[d - 1])
                return off + (idx - first) * int(cn[d]), d
            off += int(pbc[d]) * int(cn[d])
    for idx in (0, int(cs[0]), int(cs[1]) + 3, 1999):
        o, d = offset(idx)
        assert d == int(s.degrees[idx])
        assert torch.equal(flat[3 * o:3 * (o + (d + 1) ** 2)].view(-1, 3), s.sh[idx, :(d + 1) ** 2])


def test_python_api_argument_contracts():
    """Same exceptions as the reference wrapper (diff_gaussian_rasterization/__init__.py:203-207)."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _C
    cam = synth.make_camera(64, 64)
    st = GaussianRasterizationSettings(64, 64, 0.5, 0.5, torch.zeros(3), 1.0, cam.world_view_transform, cam.full_proj_transform, 3,
                                       cam.camera_center, False, False)
    r = GaussianRasterizer(st)
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), scales=torch.ones(4, 3))
    # no CPU fallback: CPU tensors are refused loudly
    with pytest.raises(RuntimeError, match="CUDA"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), degrees=torch.zeros(4, 1, dtype=torch.int32),
          scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        _C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 2), *([torch.Tensor([])] * 4), 1.0, torch.Tensor([]),
                               cam.world_view_transform, cam.full_proj_transform, 0.5, 0.5, 64, 64, torch.Tensor([]),
                               torch.Tensor([]), cam.camera_center, False, False)


def test_settings_tuple_fields():
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    assert GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                                     "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug")


def test_bench_algorithmic_bytes_worked_example():
    """SURVEY.md §8(d) worked example (C2 with V = 0.85 P, R = 4 P)."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    P = 500_000
    V, R = int(0.85 * P), 4 * P
    B = bench.algorithmic_bytes(P, V, R, 16 * V, 1920 * 1080, 8160, 1.0, 1.0, False, False)
    assert abs(B["preprocess"] / 1e6 - 126.6) < 1.0
    assert abs(B["binning"] / 1e6 - 92.9) < 1.0
    assert abs(B["render_forward"] / 1e6 - 121.5) < 1.0
    assert abs(B["render_backward"] / 1e6 - 152.1) < 1.0
    assert abs(B["preprocess_backward"] / 1e6 - 223.2) < 1.0


def test_bench_numa_binding_is_best_effort():
    """bench.py binds the ranks of a multi-GPU run to their GPU's local CPUs; without NVML / a GPU it must leave the affinity alone."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    got = bench.bind_to_gpu_numa_node(0)
    after = os.sched_getaffinity(0)
    assert got is None or (isinstance(got, int) and got == len(after))
    assert after <= before and len(after) >= min(4, len(before))
    os.sched_setaffinity(0, before)
