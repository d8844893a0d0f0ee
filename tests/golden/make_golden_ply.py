"""Golden PLY fixtures written by the reference's OWN `GaussianModel.save_ply` (scene/gaussian_model.py:239-311) and read back
by its own `load_ply` / `_parse_vertex_group` (:318-483), imported from /root/reference and run on the CPU of the build
container.  Writes tests/golden/ref_{quant,quant_half,fp32}.ply and ref_ply_expected.npz.

    python tests/golden/make_golden_ply.py

What is NOT the reference here, and why: `plyfile` is not installed, so the container writer is tests/golden/plyfile_standin
(the header / record layout plyfile produces for these calls); `simple_knn._C` and `diff_gaussian_rasterization._C` are imported
by the module but not used by save_ply / load_ply and are stubbed; `.cuda()` is the identity (no GPU here); `np.cast`, removed in
NumPy 2, is restored as `astype` for the half_float branch (:269).  The model content (xyz, SH, logits, log-scales, rotations,
degrees, 20 codebooks with u8 ids) is synthetic; every byte layout decision is the reference's code.
"""
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GS_REFERENCE_ROOT", "/root/reference")
P_MODEL, SEED = 300, 5


def synthetic_model():
    sys.path.insert(0, os.path.join(ROOT, "reduced-3dgs_b200"))
    from gs_b200 import synth
    scene = synth.make_scene(P_MODEL, SEED, sh_degree=3, mixed_degrees=True, M=16)
    order = torch.argsort(scene.degrees.view(-1), stable=True)
    scene = synth.Scene(*[getattr(scene, f)[order].contiguous() for f in ("means3D", "opacity", "scales", "rotations", "sh", "degrees")])
    return scene, synth.quantise_scene(scene, seed=0)


def import_reference_model():
    sys.path.insert(0, os.path.join(HERE, "plyfile_standin"))
    sys.path.insert(0, REF)
    for name, attrs in (("simple_knn", {}), ("simple_knn._C", {"distCUDA2": None}),
                        ("diff_gaussian_rasterization", {}), ("diff_gaussian_rasterization._C", {"calculate_colours_variance": None, "kmeans_cuda": None})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]                                  # the reference's own utils package, not the drop-in namespace portion
    torch.Tensor.cuda = lambda self, *a, **k: self
    if not hasattr(np, "cast"):
        class _Cast:
            def __getitem__(self, t):
                return lambda a: np.asarray(a).astype(t)
        np.cast = _Cast()
    import importlib.util                                   # the file itself, not the `scene` package (its __init__ pulls in the dataset readers)
    spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
    gm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gm)
    return gm


def fill(gm, scene, quant):
    m = gm.GaussianModel(3)
    m._xyz = scene.means3D.clone()
    m._features_dc = scene.sh[:, 0:1, :].clone().contiguous()
    m._features_rest = scene.sh[:, 1:, :].clone().contiguous()
    m._opacity = scene.opacity.clone()                      # raw logits
    m._scaling = torch.log(scene.scales)                    # the model stores log-scales
    m._rotation = scene.rotations.clone()
    m._degrees = scene.degrees.clone()
    m.active_sh_degree = 3
    c = quant.centers
    cb = OrderedDict()
    cb["features_dc"] = gm.Codebook(quant.ids_dc.clone(), c[0].view(-1, 1))
    for i in range(15):
        cb[f"features_rest_{i}"] = gm.Codebook(quant.ids_rest[:, i, :].clone(), c[1 + i].view(-1, 1))
    cb["opacity"] = gm.Codebook(quant.ids_opacity.view(-1, 1).clone(), c[16].view(-1, 1))
    cb["scaling"] = gm.Codebook(quant.ids_scaling.clone(), c[17].view(-1, 1))
    cb["rotation_re"] = gm.Codebook(quant.ids_rot[:, 0:1].clone(), c[18].view(-1, 1))
    cb["rotation_im"] = gm.Codebook(quant.ids_rot[:, 1:].clone(), c[19].view(-1, 1))
    m._codebook_dict = cb
    return m


def main():
    scene, quant = synthetic_model()
    gm = import_reference_model()
    model = fill(gm, scene, quant)
    expected = {}
    for tag, quantised, half in (("quant", True, False), ("quant_half", True, True), ("fp32", False, False)):
        path = os.path.join(HERE, f"ref_{tag}.ply")
        model.save_ply(path, quantised=quantised, half_float=half)          # the reference's writer
        back = gm.GaussianModel(3)
        back.load_ply(path, half_float=half, quantised=quantised)           # the reference's reader + de-quantisation (:371-387)
        for k, v in (("xyz", back._xyz), ("features_dc", back._features_dc), ("features_rest", back._features_rest), ("opacity", back._opacity),
                     ("scaling", back._scaling), ("rotation", back._rotation), ("degrees", back._degrees)):
            expected[f"{tag}_{k}"] = v.detach().cpu().numpy()
        # the activations render() reads (gaussian_model.py:141-146)
        expected[f"{tag}_get_scaling"] = back.get_scaling.detach().numpy()
        expected[f"{tag}_get_rotation"] = back.get_rotation.detach().numpy()
        print(tag, os.path.getsize(path), "bytes")
    np.savez_compressed(os.path.join(HERE, "ref_ply_expected.npz"), **expected)


if __name__ == "__main__":
    main()
