// gsb_render.cu — per-tile alpha compositing, forward and backward (sm_100a).
//
// Replaces reference forward.cu:462-582 renderCUDA and backward.cu:438-595 renderCUDA.
//
// Layout of work: one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel sub-rectangle.  The tile's
// depth-sorted instance list is staged through shared memory in batches of 256 records of 48 bytes (three
// 128-bit gathers per instance).  For every 32 staged Gaussians each lane tests ONE Gaussian against the warp's
// 8x4 rectangle (closed-form bound on the Gaussian's maximum over the rectangle) and a ballot turns the
// results into a work mask, so the per-pixel loop only visits Gaussians that can reach alpha >= 1/255 somewhere
// in the warp.  Skipped (pixel, Gaussian) pairs are exactly pairs the reference `continue`s on, so n_contrib,
// final_T and colours are unchanged.  The per-pixel arithmetic is the reference's, operation for operation.
//
// Backward: instead of 9 global atomicAdds per contributing (pixel, Gaussian) pair (backward.cu:561-592) the 9
// partial gradients are reduced across the warp with a transposed butterfly (14 shuffles), combined across the
// 8 warps in shared memory and flushed once per (tile, Gaussian) with two vector reductions + one scalar.
#include "gsb_common.cuh"

namespace gsb {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Shared-memory staging record: 48 bytes per instance, r0 | r1 | (g, b, -, -); read with one address + immediate offsets.
#define SREC_BYTES 48
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ float2 lds64(uint32_t addr)
{
	float2 v;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
	return v;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) render_forward_kernel(const uint2* __restrict__ ranges,
	const uint32_t* __restrict__ point_list,
	int W, int H, const float4* __restrict__ rec, const float* __restrict__ bg,
	float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, uint32_t* __restrict__ tile_max)
{
	__shared__ __align__(16) float4 s_rec[256 * 3];
	__shared__ uint32_t s_max;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile = blockIdx.y * gridDim.x + blockIdx.x;
	const int wx0 = blockIdx.x * GSB_TILE_X + (warp & 1) * 8, wy0 = blockIdx.y * GSB_TILE_Y + (warp >> 1) * 4;
	const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);
	const uint2 range = ranges[tile];
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_rec);
	if (tid == 0) s_max = 0;

	bool done = !inside;
	float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
	uint32_t last = 0;
	for (uint32_t b = range.x; b < range.y; b += 256)
	{
		if (__syncthreads_count(done) == 256) break;
		const int n = min(256u, range.y - b);
		if (tid < n)
		{
			const uint32_t id = point_list[b + tid];
			const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1], r2 = rec[3 * (size_t)id + 2];
			s_rec[3 * tid] = r0; s_rec[3 * tid + 1] = r1; s_rec[3 * tid + 2] = r2;
		}
		__syncthreads();
		bool warp_done = __all_sync(0xffffffffu, done);
		for (int c0 = 0; c0 < n && !warp_done; c0 += 32)
		{
			const int j = c0 + lane;
			bool keep = false;
			if (j < n)
			{
				const float4 r0 = lds128(sbase + j * SREC_BYTES), r1 = lds128(sbase + j * SREC_BYTES + 16);
				keep = rect_may_contribute(r1.x, r1.y, r0.x, r0.y, r0.z, r0.w, rx0, rx1, ry0, ry1);
			}
			unsigned mask = __ballot_sync(0xffffffffu, keep);
			const uint32_t cbase = sbase + c0 * SREC_BYTES;
			const uint32_t nbase = (b - range.x) + c0 + 1;          // 1-based list position of the chunk's first entry
			// Branch-free per-pixel body: in a surviving warp some lane nearly always takes every path of the reference's
			// if/continue chain, so predicating costs nothing and removes the divergence bookkeeping.  The arithmetic and
			// the order of the tests are the reference's (forward.cu:535-569); a masked lane changes no state.
			while (mask)
			{
				const int bit = __ffs(mask) - 1; mask &= mask - 1;
				const uint32_t addr = cbase + bit * SREC_BYTES;
				const float4 r0 = lds128(addr), r1 = lds128(addr + 16);
				const float2 gb = lds64(addr + 32);
				const float dx = __fsub_rn(r1.x, pxf), dy = __fsub_rn(r1.y, pyf);
				const float power = pair_power(r0.x, r0.y, r0.z, dx, dy);
				// power > 0: reference `continue`; power < pth: alpha = opacity*exp(power) is provably < 1/255
				bool v = !done && !(power > 0.0f) && !(power < r0.w);
				const float alpha = fminf(0.99f, __fmul_rn(r1.z, exp_ref(power)));
				v = v && !(alpha < 1.0f / 255.0f);
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				const bool stop = v && (test_T < 0.0001f);
				done = done || stop;
				v = v && !stop;
				C0 = v ? __fmaf_rn(T, __fmul_rn(r1.w, alpha), C0) : C0;
				C1 = v ? __fmaf_rn(T, __fmul_rn(gb.x, alpha), C1) : C1;
				C2 = v ? __fmaf_rn(T, __fmul_rn(gb.y, alpha), C2) : C2;
				T = v ? test_T : T;
				last = v ? nbase + bit : last;
			}
			warp_done = __all_sync(0xffffffffu, done);
		}
	}
	if (inside)
	{
		const size_t pid = (size_t)W * py + px, N = (size_t)W * H;
		final_T[pid] = T;
		n_contrib[pid] = last;
		out_color[pid] = __fmaf_rn(bg[0], T, C0);
		out_color[N + pid] = __fmaf_rn(bg[1], T, C1);
		out_color[2 * N + pid] = __fmaf_rn(bg[2], T, C2);
	}
	// where the backward pass has to start for this tile
	uint32_t m = inside ? last : 0u;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
	__syncthreads();
	if (lane == 0 && m) atomicMax(&s_max, m);
	__syncthreads();
	if (tid == 0) tile_max[tile] = s_max;
}

// ------------------------------------------------------------------------------------------------
// Sum 8 per-lane values over the warp with a transposed butterfly: afterwards lane l holds the total of value
// (l >> 2) & 7.  4+2+1 exchange shuffles + 2 plain ones instead of 8 x 5.
__device__ __forceinline__ float warp_reduce8(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7, int lane)
{
	const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
	float a0 = (b4 ? v4 : v0) + __shfl_xor_sync(0xffffffffu, b4 ? v0 : v4, 16);
	float a1 = (b4 ? v5 : v1) + __shfl_xor_sync(0xffffffffu, b4 ? v1 : v5, 16);
	float a2 = (b4 ? v6 : v2) + __shfl_xor_sync(0xffffffffu, b4 ? v2 : v6, 16);
	float a3 = (b4 ? v7 : v3) + __shfl_xor_sync(0xffffffffu, b4 ? v3 : v7, 16);
	float c0 = (b3 ? a2 : a0) + __shfl_xor_sync(0xffffffffu, b3 ? a0 : a2, 8);
	float c1 = (b3 ? a3 : a1) + __shfl_xor_sync(0xffffffffu, b3 ? a1 : a3, 8);
	float d = (b2 ? c1 : c0) + __shfl_xor_sync(0xffffffffu, b2 ? c0 : c1, 4);
	d += __shfl_xor_sync(0xffffffffu, d, 2);
	d += __shfl_xor_sync(0xffffffffu, d, 1);
	return d;   // value index = 4*b4 + 2*b3 + b2
}

#define ACC_STRIDE 9
// acc record (12 floats per Gaussian, 48 B): [dcol.r dcol.g dcol.b dop | sx sy cxx cxy | cyy - - -] where
// sx = sum dL_dG*dG_ddelx, sy = sum dL_dG*dG_ddely, cxx = sum gdx*dx*dL_dG, cxy = sum gdx*dy*dL_dG, cyy = sum gdy*dy*dL_dG;
// the constant factors (0.5*W, 0.5*H, -0.5) of backward.cu:583-589 are applied once per Gaussian by the consumer.
__global__ void __launch_bounds__(256) render_backward_kernel(const uint2* __restrict__ ranges,
	const uint32_t* __restrict__ point_list,
	int W, int H, const float4* __restrict__ rec, const float* __restrict__ bg,
	const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ tile_max,
	const float* __restrict__ dL_dpixels, float* __restrict__ acc)
{
	__shared__ __align__(16) float4 s_rec[256 * 3];
	__shared__ uint32_t s_id[256];
	__shared__ float s_acc[256 * ACC_STRIDE];
	__shared__ uint32_t s_touched[256];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile = blockIdx.y * gridDim.x + blockIdx.x;
	const uint32_t hi = tile_max[tile];
	if (hi == 0) return;
	const int wx0 = blockIdx.x * GSB_TILE_X + (warp & 1) * 8, wy0 = blockIdx.y * GSB_TILE_Y + (warp >> 1) * 4;
	const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);
	const uint2 range = ranges[tile];
	const size_t pid = (size_t)W * py + px, N = (size_t)W * H;
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_rec);

	const float T_final = inside ? final_Ts[pid] : 0.0f;
	float T = T_final;
	const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;
	float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
	if (inside) { dLp0 = dL_dpixels[pid]; dLp1 = dL_dpixels[N + pid]; dLp2 = dL_dpixels[2 * N + pid]; }
	const float bg_dot_dpixel = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
	float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
	uint32_t wmax = last_contributor;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));

	for (uint32_t b = 0; b < hi; b += 256)
	{
		const int n = min(256u, hi - b);
		__syncthreads();
		if (tid < n)
		{
			const uint32_t id = point_list[range.x + (hi - 1 - b - tid)];
			const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1], r2 = rec[3 * (size_t)id + 2];
			s_rec[3 * tid] = r0; s_rec[3 * tid + 1] = r1; s_rec[3 * tid + 2] = r2; s_id[tid] = id;
		}
#pragma unroll
		for (int k = 0; k < ACC_STRIDE; k++) s_acc[k * 256 + tid] = 0.0f;   // plain zero fill of the [256][9] array
		s_touched[tid] = 0;
		__syncthreads();
		for (int c0 = 0; c0 < n; c0 += 32)
		{
			const int j = c0 + lane;
			bool keep = false;
			if (j < n && (hi - 1 - b - j) < wmax)
			{
				const float4 r0 = lds128(sbase + j * SREC_BYTES), r1 = lds128(sbase + j * SREC_BYTES + 16);
				keep = rect_may_contribute(r1.x, r1.y, r0.x, r0.y, r0.z, r0.w, rx0, rx1, ry0, ry1);
			}
			unsigned mask = __ballot_sync(0xffffffffu, keep);
			while (mask)
			{
				const int jj = c0 + __ffs(mask) - 1; mask &= mask - 1;
				const uint32_t pos = hi - 1 - b - jj;
				const uint32_t addr = sbase + jj * SREC_BYTES;
				const float4 r0 = lds128(addr), r1 = lds128(addr + 16);
				const float dx = __fsub_rn(r1.x, pxf), dy = __fsub_rn(r1.y, pyf);
				bool active = pos < last_contributor;                                   // backward.cu:524-526
				float G = 0.f, alpha = 0.f;
				if (active)
				{
					const float power = pair_power(r0.x, r0.y, r0.z, dx, dy);
					active = !(power > 0.0f || power < r0.w);
					if (active)
					{
						G = exp_ref(power);
						alpha = fminf(0.99f, __fmul_rn(r1.z, G));
						active = !(alpha < 1.0f / 255.0f);
					}
				}
				if (!__any_sync(0xffffffffu, active)) continue;
				float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
				if (active)
				{
					const float one_m_alpha = 1.0f - alpha;
					const float inv = __frcp_rn(one_m_alpha);
					T = T * inv;                                                        // backward.cu:541
					const float dchannel_dcolor = alpha * T;
					const float2 gb = lds64(addr + 32);
					const float cr = r1.w, cg = gb.x, cb = gb.y;
					const float oml = 1.0f - last_alpha;
					ar0 = last_alpha * lc0 + oml * ar0; lc0 = cr;
					ar1 = last_alpha * lc1 + oml * ar1; lc1 = cg;
					ar2 = last_alpha * lc2 + oml * ar2; lc2 = cb;
					float dL_dalpha = (cr - ar0) * dLp0 + (cg - ar1) * dLp1 + (cb - ar2) * dLp2;
					v0 = dchannel_dcolor * dLp0; v1 = dchannel_dcolor * dLp1; v2 = dchannel_dcolor * dLp2;
					dL_dalpha *= T;
					last_alpha = alpha;
					dL_dalpha += (-T_final * inv) * bg_dot_dpixel;                      // backward.cu:569-572
					const float dL_dG = r1.z * dL_dalpha;
					const float gdx = G * dx, gdy = G * dy;
					v3 = G * dL_dalpha;                                                 // dL_dopacity
					v4 = dL_dG * (-gdx * r0.x - gdy * r0.y);                            // dL_dG * dG_ddelx
					v5 = dL_dG * (-gdy * r0.z - gdx * r0.y);                            // dL_dG * dG_ddely
					v6 = gdx * dx * dL_dG; v7 = gdx * dy * dL_dG; v8 = gdy * dy * dL_dG;
				}
				const float r8 = warp_reduce8(v0, v1, v2, v3, v4, v5, v6, v7, lane);
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) v8 += __shfl_xor_sync(0xffffffffu, v8, o);
				if ((lane & 3) == 0) atomicAdd(&s_acc[jj * ACC_STRIDE + (lane >> 2)], r8);
				if (lane == 1) { atomicAdd(&s_acc[jj * ACC_STRIDE + 8], v8); s_touched[jj] = 1; }
			}
		}
		__syncthreads();
		if (tid < n && s_touched[tid])
		{
			float* a = acc + 12 * (size_t)s_id[tid];
			const float* sa = s_acc + tid * ACC_STRIDE;
			red_add_v4(a, sa[0], sa[1], sa[2], sa[3]);
			red_add_v4(a + 4, sa[4], sa[5], sa[6], sa[7]);
			atomicAdd(a + 8, sa[8]);
		}
	}
}

// ------------------------------------------------------------------------------------------------
int launch_render_forward(const ImageState& img, const BinningState& b, const GeomState& g, int W, int H, const float* bg,
	float* out_color, cudaStream_t stream)
{
	const dim3 grid((W + GSB_TILE_X - 1) / GSB_TILE_X, (H + GSB_TILE_Y - 1) / GSB_TILE_Y);
	ProfScope prof(K_RENDER_FWD, stream);
	render_forward_kernel<<<grid, 256, 0, stream>>>(img.ranges, b.point_list, W, H, g.rec, bg,
		img.final_T, img.n_contrib, out_color, img.tile_max_contrib);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

int launch_render_backward(const ImageState& img, const BinningState& b, const GeomState& g, int W, int H, const float* bg,
	const float* dL_dpix, float* acc, cudaStream_t stream)
{
	const dim3 grid((W + GSB_TILE_X - 1) / GSB_TILE_X, (H + GSB_TILE_Y - 1) / GSB_TILE_Y);
	ProfScope prof(K_RENDER_BWD, stream);
	render_backward_kernel<<<grid, 256, 0, stream>>>(img.ranges, b.point_list, W, H, g.rec, bg,
		img.final_T, img.n_contrib, img.tile_max_contrib, dL_dpix, acc);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

} // namespace gsb
