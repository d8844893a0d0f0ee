"""Generate tests/golden/<case>.npz from the REFERENCE ITSELF (oracle/_ref/_refC.so, the unmodified reference
rasterizer compiled by oracle/build_ref.py) on a B200, and print a first oracle-vs-reference comparison.

Run on the GPU box:   python tests/golden/make_golden.py            (writes gpurun_out/golden/*.npz)
then copy the files into tests/golden/ and commit them.  The inputs are not stored: tests/golden/cases.py
regenerates them from seeds.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases  # noqa: E402
import refutil  # noqa: E402
import gs_oracle  # noqa: E402

GRAD_NAMES = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]


def run_reference(refC, name):
    c, scene, cam, bg, dL, extra = cases.build_inputs(name)
    W, H, P = c["W"], c["H"], c["P"]
    args, out = refutil.ref_forward(refC, scene, cam, bg, extra)
    R, color, radii, geomB, binB, imgB = out
    torch.cuda.synchronize()
    res = dict(num_rendered=np.int64(R), color=color.cpu().numpy(), radii=radii.cpu().numpy())
    res.update(refutil.decode_geom(geomB, P))
    res.update(refutil.decode_binning(binB, R))
    res.update(refutil.decode_image(imgB, W, H))
    res["mark_visible"] = refC.mark_visible(scene.means3D.cuda(), cam.world_view_transform.cuda(),
                                            cam.full_proj_transform.cuda()).cpu().numpy()
    if c["backward"]:
        grads = refutil.ref_backward(refC, args, out, dL, c["lam"])
        torch.cuda.synchronize()
        for n, g in zip(GRAD_NAMES, grads):
            res[n] = g.cpu().numpy()
        # a second run shows the reference's own atomic-order noise
        grads2 = refutil.ref_backward(refC, args, out, dL, c["lam"])
        for n, g, g2 in zip(GRAD_NAMES, grads, grads2):
            res["noise_" + n] = np.float32((g - g2).abs().max().item() if g.numel() else 0.0)
    if c["packed"]:
        flat, pbc, cs, cn = scene.packed_sh()
        E = refutil.EMPTY
        o2 = refC.rasterize_gaussians_variableSH_bands(
            bg.cuda(), scene.means3D.cuda(), E, scene.opacity.cuda(), scene.scales.cuda(), scene.rotations.cuda(), 1.0, E,
            cam.world_view_transform.cuda(), cam.full_proj_transform.cuda(), math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5),
            H, W, flat.cuda(), pbc.cuda(), cs.cuda(), cn.cuda(), scene.degrees.cuda(), cam.camera_center.cuda(), False, False)
        res["packed_color"] = o2[1].cpu().numpy()
        res["packed_radii"] = o2[2].cpu().numpy()
        res["packed_num_rendered"] = np.int64(o2[0])
    return res


def run_reference_tools(refC, name):
    """Outputs of the reference's own reduced_3dgs entry points (ext.cpp:21-24) on a tools case."""
    c, scene, cams, nb = cases.build_tools_inputs(name)
    ct = {k: v.cuda() for k, v in cases.tools_camera_tensors(cams).items()}
    sc = scene.to("cuda")
    P, knn = c["P"], c["knn"]

    def cv():
        return refC.calculate_colours_variance(ct["positions"], sc.means3D, sc.opacity, sc.scales, sc.rotations, ct["views"], ct["projs"],
                                               ct["tanx"], ct["tany"], ct["H"], ct["W"], sc.sh, sc.degrees, 3)
    d, v, m = cv()
    d2, v2, m2 = cv()
    res = dict(cv_distance=d.cpu().numpy(), cv_variance=v.cpu().numpy(), cv_mean=m.cpu().numpy())
    for n, a, b in (("cv_distance", d, d2), ("cv_variance", v, v2), ("cv_mean", m, m2)):
        res["noise_" + n] = np.float32(torch.nan_to_num(a - b).abs().max().item())
    px = refC.find_minimum_projected_pixel_size(ct["projs"], ct["inv_projs"], sc.means3D, ct["H"], ct["W"])
    res["pixel_size"] = px.cpu().numpy()
    half_diag = px * c["radius_scale"] * torch.sqrt(torch.tensor([3.0], device="cuda")) / 2          # scene/__init__.py:154-155
    red, mask = refC.sphere_ellipsoid_intersection(sc.means3D, sc.scales, sc.rotations, nb.cuda(), half_diag, knn)
    res["half_diagonal"] = half_diag.cpu().numpy()
    res["redundancy"] = red.cpu().numpy()
    res["intersection_mask"] = mask.cpu().numpy()
    red1 = red + 1                                                                                    # scene/__init__.py:170-176
    idx = torch.cat((torch.arange(P, device="cuda", dtype=torch.int).view(-1, 1), nb.cuda()), dim=1)
    mk = torch.cat((torch.ones((P, 1), device="cuda", dtype=torch.bool), mask), dim=1)
    res["min_redundancy"] = refC.allocate_minimum_redundancy_value(red1, idx, mk, knn + 1)[0].cpu().numpy()
    frac = float(mask.float().mean())
    assert 0.05 < frac < 0.95, frac
    return res


def run_reference_kmeans(refC, name):
    c, v, centers = cases.build_kmeans_inputs(name)
    v, centers = v.cuda(), centers.cuda()
    ids0, c0 = refC.kmeans_cuda(v, centers, c["tol"], 0)
    ids1, c1 = refC.kmeans_cuda(v, centers, 0.0, 1)
    idsf, cf = refC.kmeans_cuda(v, centers, c["tol"], c["max_iterations"])
    idsf2, cf2 = refC.kmeans_cuda(v, centers, c["tol"], c["max_iterations"])
    cost = lambda ids, cc: float((v.view(-1) - cc[ids.view(-1).long()]).abs().double().mean())
    return dict(ids_iter0=ids0.cpu().numpy(), centers_iter0=c0.cpu().numpy(), ids_iter1=ids1.cpu().numpy(), centers_iter1=c1.cpu().numpy(),
                ids_final=idsf.cpu().numpy(), centers_final=cf.cpu().numpy(), cost_final=np.float64(cost(idsf, cf)),
                noise_centers_final=np.float32((cf.sort().values - cf2.sort().values).abs().max().item()),
                noise_cost_final=np.float64(abs(cost(idsf, cf) - cost(idsf2, cf2))))


def oracle_run(name):
    c, scene, cam, bg, dL, extra = cases.build_inputs(name)
    kw = dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
              W=c["W"], H=c["H"], tan_fovx=math.tan(cam.FoVx * 0.5), tan_fovy=math.tan(cam.FoVy * 0.5))
    cov, col = extra.get("cov3D_precomp"), extra.get("colors_precomp")
    fwd = gs_oracle.forward(scene.means3D, scene.opacity, None if cov is not None else scene.scales,
                            None if cov is not None else scene.rotations, None if col is not None else scene.sh,
                            scene.degrees, col, cov, bg=bg, **kw)
    bwd = None
    if c["backward"]:
        bwd = gs_oracle.backward(fwd, dL, scene.means3D, None if cov is not None else scene.scales,
                                 None if cov is not None else scene.rotations, None if col is not None else scene.sh,
                                 scene.degrees, bg=bg, lambda_sh_sparsity=c["lam"], **kw)
    return fwd, bwd


def compare(name, ref, fwd, bwd, log=print):
    """Mismatch report oracle vs reference. Returns dict of counts."""
    vis = ref["radii"] > 0
    rep = {}

    def cnt(key, a, b, mask=None):
        a, b = np.asarray(a), np.asarray(b)
        if mask is not None:
            a, b = a[mask], b[mask]
        n = int((a != b).sum())
        rep[key] = n
        extra = ""
        if n and a.dtype.kind == "f":
            extra = " max|d|=%.3g" % float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
        log(f"  [{name}] {key:22s} mismatches {n:8d} / {a.size}{extra}")

    cnt("radii", ref["radii"], fwd["radii"])
    cnt("tiles_touched", ref["tiles_touched"], fwd["tiles_touched"])
    cnt("depths", ref["depths"], fwd["depths"], vis)
    cnt("means2D", ref["means2D"], fwd["means2D"], vis)
    cnt("cov3D", ref["cov3D"], fwd["cov3D"], vis)
    cnt("conic", ref["conic_opacity"][:, :3], fwd["conic_opacity"][:, :3], vis)
    cnt("opacity", ref["conic_opacity"][:, 3], fwd["conic_opacity"][:, 3], vis)
    cnt("rgb", ref["rgb"], fwd["rgb"], vis)
    cnt("clamped", ref["clamped"], fwd["clamped"], vis)
    log(f"  [{name}] num_rendered ref {int(ref['num_rendered'])} oracle {fwd['num_rendered']}")
    if int(ref["num_rendered"]) == fwd["num_rendered"]:
        cnt("keys", ref["keys"], fwd["keys"])
        cnt("point_list", ref["point_list"], fwd["point_list"])
        cnt("ranges", ref["ranges"], fwd["ranges"])
        nb = ~fwd["borderline"]
        cnt("n_contrib", ref["n_contrib"], fwd["n_contrib"])
        cnt("n_contrib(nonborder)", ref["n_contrib"], fwd["n_contrib"], nb)
        cnt("final_T", ref["final_T"], fwd["final_T"])
        d = np.abs(ref["color"] - fwd["color"])
        rep["color_max"] = float(d.max())
        rep["color_max_nonborder"] = float(d[:, nb].max())
        log(f"  [{name}] color max|d| {d.max():.3g}  (non-borderline {d[:, nb].max():.3g}; borderline px {int((~nb).sum())})")
    if bwd is not None:
        for n in GRAD_NAMES:
            a, b = ref[n].astype(np.float64), bwd[n].astype(np.float64).reshape(ref[n].shape)
            if a.size == 0:
                continue
            err = np.abs(a - b).max()
            scale = np.abs(a).max()
            rep["grad_" + n] = float(err / (scale + 1e-30))
            log(f"  [{name}] {n:14s} max|d| {err:.3g}  max|ref| {scale:.3g}  ref-noise {float(ref.get('noise_' + n, 0)):.3g}")
    return rep


def main():
    refC = refutil.load_ref()
    assert refC is not None, "oracle/_ref/_refC.so missing (run oracle/build_ref.py where /root/reference exists)"
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in cases.CASES:
        ref = run_reference(refC, name)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **ref)
        fwd, bwd = oracle_run(name)
        compare(name, ref, fwd, bwd)
    for name in cases.TOOLS_CASES:
        ref = run_reference_tools(refC, name)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **ref)
        print(f"  [{name}] tools golden:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in ref.items()})
    for name in cases.KMEANS_CASES:
        ref = run_reference_kmeans(refC, name)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **ref)
        print(f"  [{name}] kmeans golden: cost {float(ref['cost_final']):.6g}, reference run-to-run noise: centres "
              f"{float(ref['noise_centers_final']):.3g}, cost {float(ref['noise_cost_final']):.3g}")
    print("golden written to", out_dir, {f: os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)})


if __name__ == "__main__":
    main()
