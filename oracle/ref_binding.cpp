// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// pybind module that exposes the four hot-path entry points of the UNMODIFIED
// reference rasterizer (declared in the reference's rasterize_points.h:18-93 and
// bound by the reference's ext.cpp:17-20).  The reference's own ext.cpp cannot be
// used because it also binds the reduced_3dgs training-time tools
// (ext.cpp:21-25), which are out of scope and need far more of GLM than the
// shim in oracle/glm_shim provides.  This file is ours; it includes the
// reference header where it lies under /root/reference at build time
// (oracle/build_ref.py passes -I), no reference source is copied.
#include <torch/extension.h>
#include "rasterize_points.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
	m.def("rasterize_gaussians_variableSH_bands", &RasterizeGaussiansVariableSHBandsCUDA);
	m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
	m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
	m.def("mark_visible", &markVisible);
}
