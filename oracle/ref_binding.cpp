// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// pybind module that exposes the entry points of the UNMODIFIED reference extension
// (declared in the reference's rasterize_points.h:18-93 and reduced_3dgs.h:19-67, bound by the
// reference's ext.cpp:17-25 under the same names).  This file is ours; it includes the
// reference headers where they lie under /root/reference at build time
// (oracle/build_ref.py passes -I), no reference source is copied.
#include <torch/extension.h>
#include "rasterize_points.h"
#include "reduced_3dgs.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
	m.def("rasterize_gaussians_variableSH_bands", &RasterizeGaussiansVariableSHBandsCUDA);
	m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
	m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
	m.def("mark_visible", &markVisible);
	m.def("calculate_colours_variance", &Reduced3DGS::calculateColourVariance);
	m.def("sphere_ellipsoid_intersection", &Reduced3DGS::intersectionTest);
	m.def("allocate_minimum_redundancy_value", &Reduced3DGS::assignFinalRedundancyValue);
	m.def("find_minimum_projected_pixel_size", &Reduced3DGS::calculatePixelSize);
	m.def("kmeans_cuda", &Reduced3DGS::kmeans);
}
