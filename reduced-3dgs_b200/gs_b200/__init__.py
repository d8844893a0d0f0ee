"""gs_b200 — host-side support package of the B200-native splat rasterizer (ctypes binding, synthetic scenes,
view-sharded multi-GPU helper).  The reference-facing API lives in the sibling drop-in packages
`diff_gaussian_rasterization` and `gaussian_renderer`."""
