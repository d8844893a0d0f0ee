"""GPU: our CUDA path against the UNMODIFIED reference compiled for this GPU (oracle/_ref/_refC.so), live, at full size.
Skipped when the reference module is unavailable (it is built where /root/reference exists and shipped by gpurun)."""
import math

import numpy as np
import pytest
import torch

import refutil
from gs_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", ["C2", "C3small"])
def test_full_size_against_reference(refC, cfg):
    if refC is None:
        pytest.skip("oracle/_ref/_refC.so not available")
    import ours
    if cfg == "C2":
        scene = synth.config_scene("C2")
        quant = None
    else:
        scene0 = synth.config_scene("C3", 400_000)
        quant = synth.quantise_scene(scene0)
        d = quant.to("cuda").dequantise()
        scene = synth.Scene(*[getattr(d, f).cpu() for f in ("means3D", "opacity", "scales", "rotations", "sh", "degrees")])
    W, H = 1920, 1080
    cam = synth.make_camera(W, H)
    bg = torch.tensor([0.05, 0.1, 0.2])
    dL = synth.grad_image(W, H, 77)
    rargs, rout = refutil.ref_forward(refC, scene, cam, bg)
    R, color, radii, geomB, binB, imgB = rout
    torch.cuda.synchronize()
    g = refutil.decode_geom(geomB, scene.P)
    b = refutil.decode_binning(binB, R)
    im = refutil.decode_image(imgB, W, H)
    args, out, fwd = ours.run_forward(scene, cam, bg, quant=quant)
    assert fwd["num_rendered"] == R
    assert np.array_equal(fwd["radii"], radii.cpu().numpy())
    vis = fwd["radii"] > 0
    assert np.array_equal(fwd["depths"][vis].view(np.uint32), g["depths"][vis].view(np.uint32))
    assert np.array_equal(fwd["keys"], b["keys"]) and np.array_equal(fwd["point_list"], b["point_list"])
    assert np.array_equal(fwd["ranges"], im["ranges"]) and np.array_equal(fwd["n_contrib"], im["n_contrib"])
    assert np.abs(fwd["color"] - color.cpu().numpy()).max() <= 1e-4
    if quant is None:
        assert np.array_equal(fwd["color"], color.cpu().numpy()), "fp32 path is expected to be bit-identical to the reference"
        for k in ("means2D", "conic_opacity", "rgb"):
            assert np.array_equal(fwd[k][vis], g[k][vis]), k
    rg = refutil.ref_backward(refC, rargs, rout, dL, 0.0)
    rg2 = refutil.ref_backward(refC, rargs, rout, dL, 0.0)
    og = ours.run_backward(args, out, dL, 0.0, quant=quant)
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"]
    for n, a, a2 in zip(names, rg, rg2):
        a = a.cpu().numpy().astype(np.float64)
        scale = np.abs(a).max() + 1e-30
        noise = np.abs(a - a2.cpu().numpy()).max() / scale
        err = np.abs(a - og[n].reshape(a.shape)).max() / scale
        assert err < max(2e-4, 4 * noise), (n, err, noise)
