"""Drop-in for the reference's pybind module `diff_gaussian_rasterization._C` (ext.cpp:17-20), re-hosted on the
C ABI of include/gs_b200.h through ctypes.  Same function names, positional signatures, return tuples and error
behaviour as rasterize_points.h:18-93; tensors in, torch.Tensors out.

Build-defined extensions (keyword-only): `prune_mask` (u8/bool [P], 1 = pruned), `quant`
(a gs_b200.synth.QuantScene-like object with u8 id planes + [20,256] centres) and `debug_out` (dict that
receives the forward intermediates in the reference's GeometryState layouts).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from gs_b200 import lib as _lib
from gs_b200.lib import GsbCamera, GsbDebug, GsbGrads, GsbQuant, GsbScene, BlobAllocator, f32, on_device, ptr


def _carve_f32(device, shapes):
    """The gradient outputs as views of ONE allocation (each starting on a 256-byte boundary).  Eight separate mid-size
    tensors per backward (1-100 MB, different sizes) fragment the caching allocator's split blocks and provoke a cudaMalloc
    inside the training loop every now and then (measured: 1-90 ms); one request of a constant size is always reused."""
    sizes = [int(math.prod(sh)) for sh in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 63) // 64 * 64
    flat = torch.empty(max(total, 1), dtype=torch.float32, device=device)
    return [flat[o:o + n].view(sh) for o, n, sh in zip(offs, sizes, shapes)]


def _device_of(means3D: torch.Tensor) -> torch.device:
    if not means3D.is_cuda:
        raise RuntimeError("gs_b200: means3D must live on a CUDA device (no CPU path exists)")
    d = means3D.device
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


def _camera(device, bg, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, H, W, prefiltered, keep):
    bg, viewmatrix, projmatrix, campos = (f32(bg, device), f32(viewmatrix, device), f32(projmatrix, device),
                                           f32(campos, device))
    keep += [bg, viewmatrix, projmatrix, campos]
    return GsbCamera(int(W), int(H), float(tan_fovx), float(tan_fovy), ptr(viewmatrix), ptr(projmatrix), ptr(campos),
                     ptr(bg), int(bool(prefiltered)))


def _quant_struct(quant, device, keep):
    def u8(t):
        t = t.to(device=device, dtype=torch.uint8).contiguous()
        keep.append(t)
        return t.data_ptr()
    centers = f32(quant.centers, device)
    keep.append(centers)
    q = GsbQuant(u8(quant.ids_dc), u8(quant.ids_rest), u8(quant.ids_opacity), u8(quant.ids_scaling), u8(quant.ids_rot),
                 centers.data_ptr())
    keep.append(q)
    return C.pointer(q)


def _scene(device, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, sh, degrees, keep,
           packed_counts=None, prune_mask=None, quant=None):
    means3D = f32(means3D, device)
    P = int(means3D.shape[0]) if means3D is not None else 0
    colors, opacity, scales, rotations = f32(colors, device), f32(opacity, device), f32(scales, device), f32(rotations, device)
    cov3D_precomp, sh = f32(cov3D_precomp, device), f32(sh, device)
    if degrees is not None and degrees.numel() > 0:
        degrees = degrees.to(device=device, dtype=torch.int32).contiguous()
    else:
        degrees = None
    if prune_mask is not None:
        prune_mask = prune_mask.to(device=device, dtype=torch.uint8).contiguous()
    keep += [means3D, colors, opacity, scales, rotations, cov3D_precomp, sh, degrees, prune_mask]
    M = 0
    if quant is not None:
        M = 16
    elif sh is not None and packed_counts is None:
        M = int(sh.shape[1])                                       # rasterize_points.cu:187-191
    s = GsbScene()
    s.P, s.M = P, M
    s.means3D, s.opacities, s.scales, s.rotations = ptr(means3D), ptr(opacity), ptr(scales), ptr(rotations)
    s.cov3D_precomp, s.shs, s.colors_precomp, s.degrees = ptr(cov3D_precomp), ptr(sh), ptr(colors), ptr(degrees)
    s.scale_modifier = float(scale_modifier)
    s.sh_packed = 0
    if packed_counts is not None:
        s.sh_packed = 1
        for d in range(4):
            s.band_count[d] = int(packed_counts[d]) if d < len(packed_counts) else 0
    s.prune_mask = ptr(prune_mask)
    s.quant = _quant_struct(quant, device, keep) if quant is not None else None
    return s, P, M


def _forward(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
             tan_fovx, tan_fovy, image_height, image_width, sh, degrees, campos, prefiltered, debug, packed_counts=None,
             prune_mask=None, quant=None, debug_out=None, statistics=None):
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")          # rasterize_points.cu:158-161
    device = _device_of(means3D)
    L = _lib.lib()
    keep = []
    H, W = int(image_height), int(image_width)
    with on_device(device):
        scene, P, M = _scene(device, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, sh, degrees,
                             keep, packed_counts, prune_mask, quant)
        cam = _camera(device, background, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, H, W, prefiltered, keep)
        out_color = torch.empty((3, H, W), dtype=torch.float32, device=device)
        radii = torch.empty((P,), dtype=torch.int32, device=device)
        blobs = BlobAllocator.for_device(device)
        cbs = blobs.cb
        dbg_ptr = None
        if debug_out is not None:
            d = dict(depths=torch.zeros(P, device=device), means2D=torch.zeros(P, 2, device=device),
                     cov3D=torch.zeros(P, 6, device=device), conic_opacity=torch.zeros(P, 4, device=device),
                     rgb=torch.zeros(P, 3, device=device), tiles_touched=torch.zeros(P, dtype=torch.int32, device=device),
                     clamped=torch.zeros(P, 3, dtype=torch.uint8, device=device))
            debug_out.update(d)
            dbg = GsbDebug(*[ptr(d[k]) for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "tiles_touched", "clamped")])
            dbg_ptr = C.pointer(dbg)
        R = C.c_int64(0)
        if statistics is not None:                               # (touched_pixels int32 [P,1], transmittance_sum f32 [P,1]) to fill
            st = L.gsb_forward_statistics(C.byref(scene), C.byref(cam), cbs["geom"], None, cbs["binning"], None, cbs["image"], None,
                                          out_color.data_ptr(), ptr(radii), C.byref(R), ptr(statistics[0]), ptr(statistics[1]),
                                          _lib.current_stream(device))
        else:
            st = L.gsb_forward(C.byref(scene), C.byref(cam), cbs["geom"], None, cbs["binning"], None, cbs["image"], None,
                               out_color.data_ptr(), ptr(radii), C.byref(R), dbg_ptr, _lib.current_stream(device))
        geomB, binB, imgB = blobs.take("geom"), blobs.take("binning"), blobs.take("image")
        _lib.check(st)
        if debug:
            torch.cuda.synchronize(device)                      # reference CHECK_CUDA(debug) semantics, auxiliary.h:161-168
    return int(R.value), out_color, radii, geomB, binB, imgB


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                        projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degrees, campos, prefiltered, debug,
                        *, prune_mask=None, quant=None, debug_out=None):
    """rasterize_points.h:43-63 RasterizeGaussiansCUDA -> (R, color, radii, geomBuffer, binningBuffer, imgBuffer)."""
    return _forward(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                    projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degrees, campos, prefiltered, debug,
                    None, prune_mask, quant, debug_out)


def rasterize_gaussians_variableSH_bands(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                                         viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh,
                                         perBandPrimitiveCount, cumSumPrimitiveCount, coeffsNum, degrees, campos, prefiltered,
                                         debug, *, prune_mask=None, debug_out=None):
    """rasterize_points.h:18-41 RasterizeGaussiansVariableSHBandsCUDA (inference, packed per-degree SH groups).
    cumSumPrimitiveCount / coeffsNum are implied by perBandPrimitiveCount ([1,4,9,16] per gaussian_renderer:90-92)."""
    counts = [int(v) for v in perBandPrimitiveCount.detach().cpu().tolist()]
    return _forward(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                    projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degrees, campos, prefiltered, debug,
                    counts, prune_mask, None, debug_out)


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                                 projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degrees, campos, geomBuffer, R,
                                 binningBuffer, imageBuffer, lambda_sh_sparsity, debug, *, prune_mask=None, quant=None,
                                 accumulate_into=None, want_conic=False, view_means2D=None):
    """rasterize_points.h:65-88 RasterizeGaussiansBackwardCUDA ->
    (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations).
    `accumulate_into`: the same 8-tuple from a previous call; gradients are added in place (view-batch accumulation);
    `view_means2D` ([P,3], accumulate mode): receives THIS view's dL_dmeans2D on its own (per-view densification statistics)."""
    device = _device_of(means3D)
    L = _lib.lib()
    keep = []
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    with on_device(device):
        scene, P, M = _scene(device, means3D, colors, None, scales, rotations, scale_modifier, cov3D_precomp, sh, degrees, keep,
                             None, prune_mask, quant)
        if quant is None:
            scene.opacities = means3D.data_ptr() if P > 0 else None      # not read by the backward; keeps check_scene satisfied
        cam = _camera(device, background, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, H, W, False, keep)
        dL = f32(dL_dout_color, device)
        if accumulate_into is not None:
            outs = list(accumulate_into)
        else:
            outs = _carve_f32(device, [(P, 3), (P, 3), (P, 1), (P, 3), (P, 6), (P, M, 3), (P, 3), (P, 4)])
        conic = torch.empty((P, 4), dtype=torch.float32, device=device) if want_conic else None
        if view_means2D is not None and (accumulate_into is None or tuple(view_means2D.shape) != (P, 3) or
                                         view_means2D.dtype != torch.float32 or not view_means2D.is_contiguous()):
            raise RuntimeError("view_means2D needs accumulate_into and a contiguous fp32 [P,3] tensor")
        g = GsbGrads(ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), ptr(outs[4]), ptr(outs[5]), ptr(outs[6]), ptr(outs[7]),
                     ptr(conic), 1 if accumulate_into is not None else 0, ptr(view_means2D))
        radii = radii.to(device=device, dtype=torch.int32).contiguous()
        st = L.gsb_backward(C.byref(scene), C.byref(cam), int(R), ptr(radii), ptr(geomBuffer), ptr(binningBuffer),
                            ptr(imageBuffer), ptr(dL), C.byref(g), float(lambda_sh_sparsity), _lib.current_stream(device))
        _lib.check(st)
        if debug:
            torch.cuda.synchronize(device)
    if want_conic:
        return tuple(outs) + (conic,)
    return tuple(outs)


def calculate_colours_variance(cam_positions, means3D, opacity, scales, rotations, cam_viewmatrices, cam_projmatrices, tan_fovxs,
                               tan_fovys, image_height, image_width, sh, degrees, max_sh_deg):
    """reduced_3dgs.h:28-43 Reduced3DGS::calculateColourVariance (reduced_3dgs.cu:41-203) ->
    (average colour distance to each lower SH truncation [P, max_sh_deg], weighted colour variance [P,1,3], weighted mean colour [P,1,3]).
    Per camera: one forward with the visibility statistics on (gsb_forward_statistics) + one fused statistics kernel
    (gsb_sh_statistics_update) in place of the reference's ~30 ATen ops; the camera parameters are read back once, not per camera."""
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")          # reduced_3dgs.cu:57-60
    device = _device_of(means3D)
    if int(max_sh_deg) != 3:
        raise RuntimeError("calculate_colours_variance: the reference's colour table has 4 slots per Gaussian "
                           "(reduced_3dgs/sh_culling.cu:21), i.e. it is only meaningful for max_sh_deg == 3")
    L = _lib.lib()
    P = int(means3D.size(0))
    n_cams = int(cam_positions.size(0))
    M = int(sh.size(1)) if (P != 0 and sh.size(0) != 0) else 0
    f = lambda t: f32(t, device)
    zeros = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)
    wsum, wsumsq, dist, mean, var = zeros(P, 1), zeros(P, 1), zeros(P, 3), zeros(P, 1, 3), zeros(P, 1, 3)
    if P == 0 or n_cams == 0:
        return dist / wsum, var / wsum.view(-1, 1, 1), mean
    means3D, sh, cam_positions = f(means3D), f(sh), f(cam_positions)
    views, projs = f(cam_viewmatrices), f(cam_projmatrices)
    deg = degrees.to(device=device, dtype=torch.int32).contiguous()
    Hs, Ws = [int(v) for v in image_height.cpu().tolist()], [int(v) for v in image_width.cpu().tolist()]
    txs, tys = [float(v) for v in tan_fovxs.cpu().tolist()], [float(v) for v in tan_fovys.cpu().tolist()]
    bg = zeros(3)                                             # reduced_3dgs.cu:112 background is irrelevant here
    empty = torch.empty(0)
    touched = torch.empty((P, 1), dtype=torch.int32, device=device)
    tsum = torch.empty((P, 1), dtype=torch.float32, device=device)
    stream = _lib.current_stream(device)
    for i in range(n_cams):
        _, _, radii, _, _, _ = _forward(bg, means3D, empty, opacity, scales, rotations, 1.0, empty, views[i], projs[i], txs[i], tys[i],
                                        Hs[i], Ws[i], sh, deg, cam_positions[i], False, False, statistics=(touched, tsum))
        with on_device(device):
            _lib.check(L.gsb_sh_statistics_update(P, M, ptr(deg), ptr(means3D), cam_positions[i].data_ptr(), ptr(sh), ptr(radii), ptr(touched),
                                                  ptr(tsum), ptr(wsum), ptr(wsumsq), ptr(dist), ptr(mean), ptr(var), stream))
    return dist / wsum, var / wsum.view(-1, 1, 1), mean


def find_minimum_projected_pixel_size(w2ndc_transforms, w2ndc_transforms_inverse, means3D, image_height, image_width):
    """reduced_3dgs.h:61-66 Reduced3DGS::calculatePixelSize (reduced_3dgs.cu:246-268) -> float [P,1]."""
    device = _device_of(means3D)
    L = _lib.lib()
    P, n = int(means3D.size(0)), int(w2ndc_transforms.size(0))
    out = torch.empty((P, 1), dtype=torch.float32, device=device)
    if P == 0:
        return out
    i32 = lambda t: t.to(device=device, dtype=torch.int32).contiguous()
    m, mi, xyz, hs, ws = f32(w2ndc_transforms, device), f32(w2ndc_transforms_inverse, device), f32(means3D, device), i32(image_height), i32(image_width)
    with on_device(device):
        _lib.check(L.gsb_min_projected_pixel_size(P, ptr(xyz), n, ptr(m), ptr(mi), ptr(hs), ptr(ws), ptr(out), _lib.current_stream(device)))
    return out


def sphere_ellipsoid_intersection(means3D, scales, rotations, neighbours_indices, sphere_radius, knn):
    """reduced_3dgs.h:45-51 Reduced3DGS::intersectionTest (reduced_3dgs.cu:205-243) -> (redundancy_values int32 [P,1], intersection_mask bool [P,knn])."""
    device = _device_of(means3D)
    L = _lib.lib()
    P, knn = int(means3D.size(0)), int(knn)
    red = torch.empty((P, 1), dtype=torch.int32, device=device)
    mask = torch.empty((P, knn), dtype=torch.bool, device=device)
    if P == 0:
        return red, mask
    nb = neighbours_indices.to(device=device, dtype=torch.int32).contiguous()
    xyz, sc, rot, rad = f32(means3D, device), f32(scales, device), f32(rotations, device), f32(sphere_radius, device)
    with on_device(device):
        _lib.check(L.gsb_sphere_ellipsoid_intersection(P, ptr(xyz), ptr(sc), ptr(rot), ptr(nb), ptr(rad), knn, ptr(red), ptr(mask),
                                                       _lib.current_stream(device)))
    return red, mask


def allocate_minimum_redundancy_value(redundancy_values, neighbours_indices, intersection_mask, knn):
    """reduced_3dgs.h:53-59 Reduced3DGS::assignFinalRedundancyValue (reduced_3dgs.cu:270-287) -> 1-tuple (int32 [P,1],)."""
    device = _device_of(redundancy_values)
    L = _lib.lib()
    P, knn = int(redundancy_values.size(0)), int(knn)
    out = torch.empty((P, 1), dtype=torch.int32, device=device)
    if P == 0:
        return (out,)
    red = redundancy_values.to(device=device, dtype=torch.int32).contiguous()
    nb = neighbours_indices.to(device=device, dtype=torch.int32).contiguous()
    mask = intersection_mask.to(device=device, dtype=torch.bool).contiguous()
    with on_device(device):
        _lib.check(L.gsb_min_redundancy_value(P, ptr(red), ptr(nb), ptr(mask), knn, ptr(out), _lib.current_stream(device)))
    return (out,)


def kmeans_cuda(values, centers, tol, max_iterations):
    """reduced_3dgs.h:21-26 Reduced3DGS::kmeans (reduced_3dgs.cu:289-338) -> (ids int32 [n,1], centers float32 [k]).
    `values` is the [n,1] column of one attribute, `centers` the [k] initial centres (gaussian_model.py:36-41)."""
    device = _device_of(values)
    L = _lib.lib()
    v = f32(values, device)
    c = f32(centers, device)
    n, k = int(values.size(0)), int(centers.size(0))
    ids = torch.zeros((n, 1), dtype=torch.int32, device=device)
    out = torch.empty((k,), dtype=torch.float32, device=device)
    with on_device(device):
        ws = torch.empty(int(L.gsb_kmeans_workspace_bytes(n, k)), dtype=torch.uint8, device=device)
        _lib.check(L.gsb_kmeans(ptr(v.reshape(-1)) if n else None, n, ptr(c.reshape(-1)), k, float(tol), int(max_iterations),
                                ptr(ids), out.data_ptr(), ws.data_ptr(), _lib.current_stream(device)))
    return ids, out


def mark_visible(means3D, viewmatrix, projmatrix):
    """rasterize_points.h:90-93 markVisible -> bool[P]."""
    device = _device_of(means3D)
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=device)
    if P:
        m, v, p = f32(means3D, device), f32(viewmatrix, device), f32(projmatrix, device)
        with on_device(device):
            _lib.check(_lib.lib().gsb_mark_visible(P, ptr(m), ptr(v), ptr(p), present.data_ptr(), _lib.current_stream(device)))
    return present


def export_state(geomBuffer, binningBuffer, imageBuffer, R, W, H, P=0):
    """Decode the private blobs into reference-layout arrays (tests / tooling)."""
    device = imageBuffer.device
    L = _lib.lib()
    out = {}
    with on_device(device):
        keys = torch.zeros(max(R, 0), dtype=torch.int64, device=device)
        pl = torch.zeros(max(R, 0), dtype=torch.int32, device=device)
        if R > 0:
            _lib.check(L.gsb_export_binning(ptr(geomBuffer), int(P), ptr(binningBuffer), int(R), ptr(imageBuffer), W, H,
                                            ptr(keys), ptr(pl), _lib.current_stream(device)))
        T = ((W + 15) // 16) * ((H + 15) // 16)
        final_T = torch.zeros(H, W, device=device)
        n_contrib = torch.zeros(H, W, dtype=torch.int32, device=device)
        ranges = torch.zeros(T, 2, dtype=torch.int32, device=device)
        _lib.check(L.gsb_export_image(ptr(imageBuffer), W, H, ptr(final_T), ptr(n_contrib), ptr(ranges), _lib.current_stream(device)))
    out.update(keys=keys, point_list=pl, final_T=final_T, n_contrib=n_contrib, ranges=ranges)
    return out


def debug_dequant(quant):
    """Fused de-quantisation on its own (test helper): -> (scales [P,3], rotations [P,4]) as the kernels compute them."""
    device = quant.means3D.device
    P = int(quant.means3D.shape[0])
    keep = []
    q = _quant_struct(quant, device, keep)
    scales = torch.empty(P, 3, device=device)
    rots = torch.empty(P, 4, device=device)
    with on_device(device):
        _lib.check(_lib.lib().gsb_debug_dequant(q, P, scales.data_ptr(), rots.data_ptr(), _lib.current_stream(device)))
    return scales, rots
