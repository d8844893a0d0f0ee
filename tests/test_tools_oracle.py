"""CPU: the oracle's restatement of the reduced-3dgs tools (oracle/gs_oracle.py: colours_variance, min_projected_pixel_size,
sphere_ellipsoid_intersection, min_redundancy_value) against the golden outputs of the reference itself (tests/golden/t1.npz,
generated on a B200 by tests/golden/make_golden.py from oracle/_ref/_refC.so)."""
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import cases  # noqa: E402
import gs_oracle as O  # noqa: E402

GOLD = os.path.join(HERE, "golden")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = ~(np.isnan(a) | np.isnan(b))
    return np.abs(a[m] - b[m]).max() / (np.abs(b[m]).max() + 1e-30)


@pytest.mark.parametrize("name", list(cases.TOOLS_CASES))
def test_tools_oracle_against_reference_goldens(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.isfile(path):
        pytest.skip("golden not generated yet")
    ref = dict(np.load(path))
    c, scene, cams, nb = cases.build_tools_inputs(name)
    ct = cases.tools_camera_tensors(cams)
    d, v, m, per = O.colours_variance(ct["positions"], scene.means3D, scene.opacity, scene.scales, scene.rotations, ct["views"], ct["projs"],
                                      ct["tanx"], ct["tany"], ct["H"], ct["W"], scene.sh, scene.degrees)
    border = any(p["borderline"].any() for p in per)
    tol = 2e-5 if not border else 2e-3
    for k, a in (("cv_distance", d), ("cv_variance", v), ("cv_mean", m)):
        assert np.array_equal(np.isnan(a), np.isnan(ref[k])), k
        assert _rel(a, ref[k]) <= tol, (k, _rel(a, ref[k]))
    px = O.min_projected_pixel_size(ct["projs"], ct["inv_projs"], scene.means3D, ct["H"], ct["W"])
    assert np.array_equal(px, ref["pixel_size"]), "same operation order, IEEE div/sqrt: expected bit-identical"
    red, mask, bl = O.sphere_ellipsoid_intersection(scene.means3D, scene.scales, scene.rotations, nb, ref["half_diagonal"], c["knn"])
    ok = ~bl
    assert np.array_equal(mask[ok], ref["intersection_mask"][ok]) and np.array_equal(red[ok], ref["redundancy"][ok])
    assert (mask != ref["intersection_mask"]).sum() <= bl.sum()
    P = scene.P
    idx = np.concatenate([np.arange(P, dtype=np.int32).reshape(-1, 1), nb.numpy()], 1)
    mk = np.concatenate([np.ones((P, 1), bool), ref["intersection_mask"]], 1)
    mn = O.min_redundancy_value(ref["redundancy"] + 1, idx, mk, c["knn"] + 1)
    assert np.array_equal(mn, ref["min_redundancy"])


def test_tools_oracle_self_consistency():
    """Independent checks that do not need the golden: pixel size against the reference's own PyTorch formulation
    (scene/__init__.py:104-142, restated with numpy in double), redundancy counts against a direct evaluation."""
    c, scene, cams, nb = cases.build_tools_inputs("t1", P=500)
    ct = cases.tools_camera_tensors(cams)
    xyz = scene.means3D.numpy().astype(np.float64)
    px = O.min_projected_pixel_size(ct["projs"], ct["inv_projs"], scene.means3D, ct["H"], ct["W"]).reshape(-1)
    want = np.full(xyz.shape[0], 10000.0)
    for i in range(len(cams)):
        M, Mi = ct["projs"][i].numpy().astype(np.float64), ct["inv_projs"][i].numpy().astype(np.float64)
        W, H = int(ct["W"][i]), int(ct["H"][i])
        h = np.concatenate([xyz, np.ones((xyz.shape[0], 1))], 1) @ M
        h = h / h[:, 3:]
        inside = (np.abs(h[:, 0]) <= 1) & (np.abs(h[:, 1]) <= 1) & (h[:, 2] <= 1) & (h[:, 2] >= 0)
        p1 = np.zeros_like(h); p0 = np.zeros_like(h)
        p1[:, 0 if W > H else 1] = min(2 / W, 2 / H)
        p1[:, 2] = h[:, 2]; p1[:, 3] = 1; p0[:, 2] = h[:, 2]; p0[:, 3] = 1
        q1 = p1 @ Mi; q1 = q1 / q1[:, 3:]
        q0 = p0 @ Mi; q0 = q0 / q0[:, 3:]
        dist = np.linalg.norm((q1 - q0)[:, :3], axis=1)
        want[inside] = np.minimum(want[inside], dist[inside])
    seen = want < 10000
    assert np.array_equal(seen, px < 10000)
    assert np.abs(px[seen] / want[seen] - 1).max() < 2e-3          # fp32 through an ill-conditioned inverse projection
    rad = (px * 6.0 * math.sqrt(3) / 2).astype(np.float32)
    red, mask, bl = O.sphere_ellipsoid_intersection(scene.means3D, scene.scales, scene.rotations, nb, rad, c["knn"])
    assert np.array_equal(red.reshape(-1), mask.sum(1)) and 0 < mask.mean() < 1
    mn = O.min_redundancy_value(red, nb, mask, c["knn"])
    assert np.all(mn <= xyz.shape[0]) and np.all(mn >= 0)


@pytest.mark.parametrize("name", list(cases.KMEANS_CASES))
def test_kmeans_oracle_against_reference_goldens(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.isfile(path):
        pytest.skip("golden not generated yet")
    ref = dict(np.load(path))
    c, v, centers = cases.build_kmeans_inputs(name)
    assert np.array_equal(O.kmeans_update_ids(v.numpy(), centers.numpy()).reshape(-1, 1), ref["ids_iter0"])
    ids1, c1, _ = O.kmeans(v.numpy(), centers.numpy(), 0.0, 1)
    assert np.abs(c1 - ref["centers_iter1"]).max() <= 2e-6 * np.abs(ref["centers_iter1"]).max() + 1e-7
    idsf, cf, it = O.kmeans(v.numpy(), centers.numpy(), c["tol"], c["max_iterations"])
    cost = np.abs(v.numpy().reshape(-1).astype(np.float64) - cf.astype(np.float64)[idsf.reshape(-1)]).mean()
    assert abs(cost - float(ref["cost_final"])) <= max(2e-3 * float(ref["cost_final"]), 4 * float(ref["noise_cost_final"]))


def test_kmeans_oracle_properties():
    c, v, centers = cases.build_kmeans_inputs("k1", n=5000)
    ids, cc, it = O.kmeans(v.numpy(), centers.numpy(), 1e-4, 100)
    assert ids.shape == (5000, 1) and cc.shape == (256,) and 1 <= it <= 100
    d = np.abs(v.numpy().reshape(-1, 1) - cc.reshape(1, -1))
    assert np.allclose(d[np.arange(5000), ids.reshape(-1)], d.min(1), atol=1e-7)


def test_loss_oracle_against_reference_golden():
    """oracle.l1_ssim (float64, analytic gradient) vs the reference's own utils/loss_utils.py + autograd (tests/golden/loss1.npz,
    written by tests/golden/make_golden_loss.py by importing /root/reference on CPU)."""
    import make_golden_loss as M
    ref = dict(np.load(os.path.join(GOLD, "loss1.npz")))
    img, gt = M.inputs()
    l1, ss, loss, grad = O.l1_ssim(img, gt, float(ref["lambda_dssim"]))
    assert abs(l1 - float(ref["l1"])) < 1e-7 and abs(ss - float(ref["ssim"])) < 2e-5 and abs(loss - float(ref["loss"])) < 1e-5
    assert np.abs(grad - ref["grad"]).max() <= 2e-4 * np.abs(ref["grad"]).max()
    _, _, _, g_ssim = O.l1_ssim(img, gt, 1.0)                      # loss = 1 - ssim  ->  d ssim = -grad
    assert np.abs(-g_ssim - ref["grad_ssim_only"]).max() <= 2e-4 * np.abs(ref["grad_ssim_only"]).max()
