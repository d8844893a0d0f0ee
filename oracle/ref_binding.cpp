// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Python module around the UNMODIFIED reference extension: the nine entry points the reference declares in
// rasterize_points.h:18-93 and reduced_3dgs.h:19-67 (its own ext.cpp:17-25 exports them under the same Python names).
// This file is ours; the two reference headers are included where they lie under /root/reference at build time
// (oracle/build_ref.py passes -I), no reference source is copied.
#include <torch/extension.h>
#include "rasterize_points.h"
#include "reduced_3dgs.h"

namespace {
template <class Fn>
void expose(pybind11::module_& mod, const char* python_name, Fn* entry)
{
	mod.def(python_name, entry);
}
} // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, mod)
{
	// rasterizer (hot path): forward, forward with packed per-degree SH, backward, frustum test
	expose(mod, "rasterize_gaussians", &RasterizeGaussiansCUDA);
	expose(mod, "rasterize_gaussians_variableSH_bands", &RasterizeGaussiansVariableSHBandsCUDA);
	expose(mod, "rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
	expose(mod, "mark_visible", &markVisible);
	// reduced-3dgs tools: SH-culling statistics, redundancy score (three steps), codebook k-means
	expose(mod, "calculate_colours_variance", &Reduced3DGS::calculateColourVariance);
	expose(mod, "find_minimum_projected_pixel_size", &Reduced3DGS::calculatePixelSize);
	expose(mod, "sphere_ellipsoid_intersection", &Reduced3DGS::intersectionTest);
	expose(mod, "allocate_minimum_redundancy_value", &Reduced3DGS::assignFinalRedundancyValue);
	expose(mod, "kmeans_cuda", &Reduced3DGS::kmeans);
}
