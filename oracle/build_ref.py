"""Build the UNMODIFIED reference rasterizer into oracle/_ref/ (test infrastructure only).

The reference (graphdeco-inria/reduced-3dgs, submodules/diff-gaussian-rasterization) is CUDA and
needs GLM, whose submodule directory is empty in /root/reference.  We compile its translation units
(rasterizer + the reduced_3dgs tools) where they lie (no sources are copied), against the minimal GLM stand-in in
oracle/glm_shim/, with the flags torch's BuildExtension would pass for the reference's setup.py
(setup.py:23-27: only the GLM include path and --disable-warnings; arch = the GPU's: sm_100).

Outputs (git-ignored, but shipped to the GPU box by gpurun):
    oracle/_ref/_refC.so      pybind module with the reference's nine entry points (ext.cpp:17-25):
                              rasterize_gaussians, rasterize_gaussians_backward,
                              rasterize_gaussians_variableSH_bands, mark_visible, calculate_colours_variance,
                              sphere_ellipsoid_intersection, allocate_minimum_redundancy_value,
                              find_minimum_projected_pixel_size, kmeans_cuda

Uses: GPU-vs-GPU parity tests (tests/test_gpu_vs_reference.py), golden fixture generation
(tests/golden/make_golden.py) and the `bench.py --impl reference` arm.  Never used by the product.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GS_REFERENCE_ROOT", "/root/reference")
DGR = os.path.join(REF, "submodules", "diff-gaussian-rasterization")
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "_refC.so")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(DGR, "cuda_rasterizer", "forward.cu"))


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Returns the path of the built module, or None when /root/reference is absent (GPU box)."""
    if not reference_available():
        return SO if os.path.isfile(SO) else None
    srcs = [os.path.join(DGR, "cuda_rasterizer", f) for f in ("rasterizer_impl.cu", "forward.cu", "backward.cu")]
    srcs += [os.path.join(DGR, "rasterize_points.cu"), os.path.join(DGR, "reduced_3dgs.cu")]
    srcs += [os.path.join(DGR, "reduced_3dgs", f) for f in ("kmeans.cu", "redundancy_score.cu", "sh_culling.cu")]
    srcs += [os.path.join(HERE, "ref_binding.cpp")]
    deps = srcs + [os.path.join(HERE, "glm_shim", "glm", "glm.hpp"), __file__]
    if not force and os.path.isfile(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    from torch.utils import cpp_extension as ce
    import torch
    inc = []
    for p in ce.include_paths("cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(True):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"], "-I", os.path.join(HERE, "glm_shim"), "-I", DGR,
            "-I", os.path.join(DGR, "cuda_rasterizer")]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-std=c++17", "-DTORCH_EXTENSION_NAME=_refC", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-include", "cstdint"]
    nvcc_flags = ["-gencode", "arch=compute_100,code=sm_100", "--disable-warnings", "--expt-relaxed-constexpr",
                  "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                  "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                  "--compiler-options", "-fPIC"]
    objs, cmds = [], []
    for s in srcs:
        o = os.path.join(OUT, "obj", os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmds.append(["nvcc", "-c", s, "-o", o] + common + nvcc_flags + inc)
        else:
            cmds.append(["g++", "-c", s, "-o", o, "-O2", "-fPIC"] + common + inc)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            print(" ".join(cmd[:6]), "... ok", flush=True)

    with ThreadPoolExecutor(max_workers=len(cmds)) as ex:
        list(ex.map(run, cmds))
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["g++", "-shared", "-o", SO] + objs + ["-L", libdir, "-L", "/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu",
                                                   "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
                                                   f"-Wl,-rpath,{libdir}"]
    run(link)
    return SO


def load():
    """Import oracle/_ref/_refC.so (build first if the reference tree is present). None if unavailable."""
    path = build()
    if path is None or not os.path.isfile(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location("_refC", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("reference module:", p)
