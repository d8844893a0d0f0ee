// gsb_tools.cu — the reduced-3dgs tools that sit on either side of the rasterizer:
//   * SH-culling statistics: per-camera update of the transmittance-weighted colour statistics that decide each
//     Gaussian's SH degree (reference reduced_3dgs.cu:41-203 calculateColourVariance + reduced_3dgs/sh_culling.cu)
//   * resolution-aware redundancy score: pixel footprint of a Gaussian centre over all cameras, sphere / ellipsoid
//     intersection count against the k nearest neighbours, minimum score over intersecting neighbours
//     (reference reduced_3dgs/redundancy_score.cu, reduced_3dgs.cu:205-287)
//
// The reference runs these as ~30 ATen element-wise ops per camera (colour statistics) and one kernel launch + one host
// synchronisation per camera (pixel size); here each is ONE fused pass over the Gaussians, bandwidth-bound.
#include "gsb_common.cuh"

namespace gsb {

// auxiliary.h:22-38 (the same table the rasterizer uses)
__device__ __constant__ const float kT_C2[5] = { 1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f };
__device__ __constant__ const float kT_C3[7] = { -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
	-0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f };

// ------------------------------------------------------------------------------------------------
// One camera's update of the colour statistics (reduced_3dgs.cu:150-201), one thread per Gaussian.
//   t        = transmittance_sum / max(touched_pixels, 1)                                     :154
//   wSum    += t;  wSumSq += t^2                                                              :155-156
//   colours  = SH colour truncated after band 0, 1, 2, 3 (slot k exists only if k <= degree;
//              slots above the Gaussian's degree stay 0; invisible Gaussians are all 0)       :158-165, sh_culling.cu:6-57
//   dist[d] += t * || colours[3] - colours[d] ||_2   (NaN -> 0)            d = 0, 1, 2         :167-181
//   visible Gaussians: weighted running mean of colours[3]; variance += t * (colour - new mean)^2  :183-200
// The colour table has 4 slots per Gaussian (sh_culling.cu:21 hard-codes the stride), i.e. max_sh_degree = 3.
__global__ void __launch_bounds__(256) sh_stats_update_kernel(int P, int M, const int* __restrict__ degrees,
	const float* __restrict__ means3D, const float* __restrict__ campos, const float* __restrict__ shs,
	const int* __restrict__ radii, const int* __restrict__ touched, const float* __restrict__ tsum,
	float* __restrict__ wSum, float* __restrict__ wSumSq, float* __restrict__ dist_accum,
	float* __restrict__ mean, float* __restrict__ variance)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	const int tp = touched[idx];
	const float t = tsum[idx] / (float)max(tp, 1);
	const float ws = wSum[idx] + t;
	wSum[idx] = ws;
	wSumSq[idx] += t * t;
	const bool present = radii[idx] > 0;

	float col[4][3];
#pragma unroll
	for (int k = 0; k < 4; k++) col[k][0] = col[k][1] = col[k][2] = 0.0f;
	if (present)
	{
		const float px = means3D[3 * idx], py = means3D[3 * idx + 1], pz = means3D[3 * idx + 2];
		float dx = px - campos[0], dy = py - campos[1], dz = pz - campos[2];
		const float len = sqrtf(dx * dx + dy * dy + dz * dz);
		dx = dx / len; dy = dy / len; dz = dz / len;
		const float* sh = shs + (size_t)idx * M * 3;
		const int deg = degrees[idx];
		float res[3];
#pragma unroll
		for (int c = 0; c < 3; c++)
		{
			res[c] = kSH_C0 * sh[c] + 0.5f;
			col[0][c] = fmaxf(res[c], 0.0f);
		}
		if (deg > 0)
		{
			const float x = dx, y = dy, z = dz;
#pragma unroll
			for (int c = 0; c < 3; c++)
			{
				res[c] = res[c] - kSH_C1 * y * sh[3 + c] + kSH_C1 * z * sh[6 + c] - kSH_C1 * x * sh[9 + c];
				col[1][c] = fmaxf(res[c], 0.0f);
			}
			if (deg > 1)
			{
				const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
				for (int c = 0; c < 3; c++)
				{
					res[c] = res[c] + kT_C2[0] * xy * sh[12 + c] + kT_C2[1] * yz * sh[15 + c] + kT_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] +
						kT_C2[3] * xz * sh[21 + c] + kT_C2[4] * (xx - yy) * sh[24 + c];
					col[2][c] = fmaxf(res[c], 0.0f);
				}
				if (deg > 2)
				{
#pragma unroll
					for (int c = 0; c < 3; c++)
					{
						res[c] = res[c] + kT_C3[0] * y * (3.0f * xx - yy) * sh[27 + c] + kT_C3[1] * xy * z * sh[30 + c] +
							kT_C3[2] * y * (4.0f * zz - xx - yy) * sh[33 + c] + kT_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
							kT_C3[4] * x * (4.0f * zz - xx - yy) * sh[39 + c] + kT_C3[5] * z * (xx - yy) * sh[42 + c] +
							kT_C3[6] * x * (xx - 3.0f * yy) * sh[45 + c];
						col[3][c] = fmaxf(res[c], 0.0f);
					}
				}
			}
		}
	}
#pragma unroll
	for (int d = 0; d < 3; d++)
	{
		const float a = col[3][0] - col[d][0], b = col[3][1] - col[d][1], c = col[3][2] - col[d][2];
		float dist = sqrtf(a * a + b * b + c * c);
		if (isnan(dist)) dist = 0.0f;
		dist_accum[3 * idx + d] += t * dist;
	}
	if (present)
	{
		float coef = t / ws;
		if (isnan(coef)) coef = 0.0f;
#pragma unroll
		for (int c = 0; c < 3; c++)
		{
			const float m_old = mean[3 * idx + c];
			const float m_new = m_old + coef * (col[3][c] - m_old);
			mean[3 * idx + c] = m_new;
			// reduced_3dgs.cu:185 `auto mean_old = mean;` is a handle to the SAME tensor, so after the in-place update of `mean`
			// both factors of the variance term (:196-200) see the new mean
			variance[3 * idx + c] += t * (col[3][c] - m_new) * (col[3][c] - m_new);
		}
	}
}

int launch_sh_stats_update(int P, int M, const int* degrees, const float* means3D, const float* campos, const float* shs,
	const int* radii, const int* touched, const float* tsum, float* wSum, float* wSumSq, float* dist_accum, float* mean,
	float* variance, cudaStream_t stream)
{
	if (P <= 0) return GSB_OK;
	ProfScope prof(K_TOOLS, stream);
	sh_stats_update_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, M, degrees, means3D, campos, shs, radii, touched, tsum, wSum, wSumSq,
		dist_accum, mean, variance);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Minimum world-space size of one pixel at each Gaussian centre over all cameras (redundancy_score.cu:45-101,
// reduced_3dgs.cu:246-268).  The reference launches one kernel per camera (reading H, W back to the host each time) that
// min-updates pixel_sizes in global memory; here one thread walks all cameras with the running minimum in a register.
// Matrices are the reference's flat 4x4 tensors reinterpreted as column-major glm::mat4 (m[c][r] = flat[4c + r]);
// M * v is evaluated as (M[0] v.x + M[1] v.y) + (M[2] v.z + M[3] v.w) like GLM.
// The float operation order is the reference build's (read off its SASS): per row fma(v.x, m0, v.y*m1) + fma(v.z, m2, m3*v.w).
__device__ __forceinline__ void mat4_mul(const float* __restrict__ m, float vx, float vy, float vz, float (&o)[4])
{
#pragma unroll
	for (int r = 0; r < 4; r++)
		o[r] = __fadd_rn(__fmaf_rn(vx, m[r], __fmul_rn(vy, m[4 + r])), __fmaf_rn(vz, m[8 + r], m[12 + r]));     // v.w == 1
}

__global__ void __launch_bounds__(256) pixel_size_kernel(int P, const float* __restrict__ means3D, int n_cams,
	const float* __restrict__ w2ndc, const float* __restrict__ w2ndc_inv, const int* __restrict__ heights, const int* __restrict__ widths,
	float* __restrict__ pixel_sizes)
{
	extern __shared__ float s_mat[];                       // [n_cams][32]: forward | inverse
	for (int i = threadIdx.x; i < n_cams * 32; i += blockDim.x)
	{
		const int cam = i >> 5, k = i & 31;
		s_mat[i] = k < 16 ? w2ndc[16 * cam + k] : w2ndc_inv[16 * cam + k - 16];
	}
	__syncthreads();
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	const float cx = means3D[3 * idx], cy = means3D[3 * idx + 1], cz = means3D[3 * idx + 2];
	float best = 10000.0f;                                  // reduced_3dgs.cu:256 initial value
	for (int cam = 0; cam < n_cams; cam++)
	{
		const float* pm = s_mat + 32 * cam;
		const float* im = pm + 16;
		float ph[4];
		mat4_mul(pm, cx, cy, cz, ph);
		float pw = __fdiv_rn(1.0f, __fadd_rn(ph[3], 0.0000001f));
		const float qx = __fmul_rn(ph[0], pw), qy = __fmul_rn(ph[1], pw), qz = __fmul_rn(ph[2], pw);
		const bool inside = qx <= 1.0f && qy <= 1.0f && qz <= 1.0f && qx >= -1.0f && qy >= -1.0f && qz >= 0.0f;
		if (!inside) continue;
		const int W = widths[cam], H = heights[cam];
		float ex = 0.0f, ey = 0.0f;
		if (W > H) ex = __fdiv_rn(2.0f, (float)W); else ey = __fdiv_rn(2.0f, (float)H);
		float e[4], s[4];
		mat4_mul(im, ex, ey, qz, e);
		pw = __fdiv_rn(1.0f, __fadd_rn(e[3], 0.0000001f));
		const float enx = __fmul_rn(e[0], pw), eny = __fmul_rn(e[1], pw), enz = __fmul_rn(e[2], pw);
		mat4_mul(im, 0.0f, 0.0f, qz, s);
		pw = __fdiv_rn(1.0f, __fadd_rn(s[3], 0.0000001f));
		const float dx = __fmaf_rn(-s[0], pw, enx), dy = __fmaf_rn(-s[1], pw, eny), dz = __fmaf_rn(-s[2], pw, enz);
		const float len = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
		best = fminf(best, len);
	}
	pixel_sizes[idx] = best;
}

int launch_pixel_size(int P, const float* means3D, int n_cams, const float* w2ndc, const float* w2ndc_inv, const int* heights,
	const int* widths, float* pixel_sizes, cudaStream_t stream)
{
	if (P <= 0) return GSB_OK;
	if (n_cams > 1024) { set_error("find_minimum_projected_pixel_size: more than 1024 cameras per call"); return GSB_EINVAL; }
	ProfScope prof(K_TOOLS, stream);
	if (int e = ensure_dyn_smem((const void*)pixel_size_kernel, 1024 * 32 * 4)) return e;
	pixel_size_kernel<<<(P + 255) / 256, 256, (size_t)n_cams * 32 * sizeof(float), stream>>>(P, means3D, n_cams, w2ndc, w2ndc_inv, heights, widths,
		pixel_sizes);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Sphere / ellipsoid intersection against the k nearest neighbours (redundancy_score.cu:119-160 + buildRotationMatrixCUDA
// :186-205, fused: the 3x3 rotation is rebuilt from the quaternion in registers instead of a [P,3,3] tensor round trip).
// Reference quirk kept: the rotation used for neighbour i is the CURRENT Gaussian's (`R[idx]`, :143), not the neighbour's.
// One warp handles 32 Gaussians; the neighbour list rows are read with coalesced loads (lane = neighbour slot).
__global__ void __launch_bounds__(256) sphere_ellipsoid_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ scales,
	const float* __restrict__ rotations, const int* __restrict__ neighbours, const float* __restrict__ sphere_radius, int knn,
	int* __restrict__ redundancy_values, uint8_t* __restrict__ intersection_mask)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= P) return;
	const float cx = means3D[3 * idx], cy = means3D[3 * idx + 1], cz = means3D[3 * idx + 2];
	const float rad = sphere_radius[idx];
	const float4 q = reinterpret_cast<const float4*>(rotations)[idx];
	const float r = q.x, x = q.y, y = q.z, z = q.w;
	// column-major mat3(c0 | c1 | c2) of redundancy_score.cu:201-204, in the operation order of the reference build (its SASS)
	const float rz = __fmul_rn(r, z), ry = __fmul_rn(r, y), yz = __fmul_rn(y, z), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
	const float m00 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(yy, zz)));
	const float m01 = __fmul_rn(2.f, __fmaf_rn(x, y, rz)), m02 = __fmul_rn(2.f, __fmaf_rn(x, z, -ry));
	const float m10 = __fmul_rn(2.f, __fmaf_rn(x, y, -rz)), m11 = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, zz)));
	const float m12 = __fmul_rn(2.f, __fmaf_rn(r, x, yz));
	const float m20 = __fmul_rn(2.f, __fmaf_rn(x, z, ry)), m21 = __fmul_rn(2.f, __fmaf_rn(-r, x, yz));
	const float m22 = __fsub_rn(1.f, __fmul_rn(2.f, __fmaf_rn(x, x, yy)));
	const int* nb = neighbours + (size_t)idx * knn;
	uint8_t* mk = intersection_mask + (size_t)idx * knn;
	int count = 0;
	for (int i = 0; i < knn; i++)
	{
		const int n = nb[i];
		const float dx = __fsub_rn(cx, means3D[3 * n]), dy = __fsub_rn(cy, means3D[3 * n + 1]), dz = __fsub_rn(cz, means3D[3 * n + 2]);
		const float sx = __fadd_rn(scales[3 * n], rad), sy = __fadd_rn(scales[3 * n + 1], rad), sz = __fadd_rn(scales[3 * n + 2], rad);
		// row vector times matrix: component c = dot(column c, d), contracted as fma(d.z, m2, fma(d.x, m0, d.y*m1))
		const float lx = __fmaf_rn(dz, m02, __fmaf_rn(dx, m00, __fmul_rn(dy, m01)));
		const float ly = __fmaf_rn(dz, m12, __fmaf_rn(dx, m10, __fmul_rn(dy, m11)));
		const float lz = __fmaf_rn(dz, m22, __fmaf_rn(dx, m20, __fmul_rn(dy, m21)));
		// glm::pow(v, vec3(2)) is the full powf (the reference build does not reduce it to a product)
		const float bx = __frcp_rn(powf(sx, 2.0f)), by = __frcp_rn(powf(sy, 2.0f)), bz = __frcp_rn(powf(sz, 2.0f));
		const float dot = __fmaf_rn(bz, powf(lz, 2.0f), __fmaf_rn(bx, powf(lx, 2.0f), __fmul_rn(by, powf(ly, 2.0f))));
		const bool hit = dot < 1.0f;
		mk[i] = hit ? 1 : 0;
		count += hit ? 1 : 0;
	}
	redundancy_values[idx] = count;
}

int launch_sphere_ellipsoid(int P, const float* means3D, const float* scales, const float* rotations, const int* neighbours,
	const float* sphere_radius, int knn, int* redundancy_values, uint8_t* intersection_mask, cudaStream_t stream)
{
	if (P <= 0) return GSB_OK;
	ProfScope prof(K_TOOLS, stream);
	sphere_ellipsoid_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, scales, rotations, neighbours, sphere_radius, knn,
		redundancy_values, intersection_mask);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

// ------------------------------------------------------------------------------------------------
// minimum_redundancy[n] = min over Gaussians i that intersect neighbour n of redundancy[i] (redundancy_score.cu:6-27);
// initial value P (reduced_3dgs.cu:279).  Integer atomicMin: the result does not depend on the order.
__global__ void __launch_bounds__(256) fill_int_kernel(int n, int v, int* __restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = v;
}
__global__ void __launch_bounds__(256) min_redundancy_kernel(int P, const int* __restrict__ redundancy_values, const int* __restrict__ neighbours,
	const uint8_t* __restrict__ intersection_mask, int knn, int* __restrict__ minimum)
{
	const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (Gaussian, neighbour slot): coalesced
	if (e >= (long long)P * knn) return;
	if (intersection_mask[e]) atomicMin(&minimum[neighbours[e]], redundancy_values[e / knn]);
}

int launch_min_redundancy(int P, const int* redundancy_values, const int* neighbours, const uint8_t* intersection_mask, int knn,
	int* minimum, cudaStream_t stream)
{
	if (P <= 0) return GSB_OK;
	ProfScope prof(K_TOOLS, stream);
	fill_int_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, P, minimum);
	GSB_LAUNCHED();
	const long long E = (long long)P * knn;
	if (E > 0)
	{
		min_redundancy_kernel<<<(unsigned)((E + 255) / 256), 256, 0, stream>>>(P, redundancy_values, neighbours, intersection_mask, knn, minimum);
		GSB_LAUNCHED();
	}
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

} // namespace gsb
