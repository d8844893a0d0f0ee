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
// per-Gaussian sums over a warp's 32 pixels are computed as a small matrix product on the tensor cores (3xTF32
// mma.sync, see below) and added to the per-Gaussian accumulator with three vector reductions per (warp, Gaussian).
// The tile's list is staged through a ring of shared-memory buffers filled by TMA bulk copies (mbarrier-tracked); the survivors
// of each staged batch are compacted into a per-warp work queue before the per-pixel loop.
#include <cstddef>
#include <cstdlib>
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
// STATS = true additionally accumulates, per Gaussian, the number of pixels it contributed to and the sum of the
// transmittance T in front of it at those pixels (forward.cu:560-564, `calculate_mean_transmittance`): one pair of
// atomics per (warp, Gaussian) after a ballot / shuffle reduction instead of two per (pixel, Gaussian).
// Instruction budget (the kernel is issue-bound: ~90 % of the issue slots are busy, ncu).  "This pixel is finished" is the SIGN
// BIT of T: a finished pixel has T < 0, so its test_T = T * (1 - alpha) <= 0 < 1e-4 takes the reference's stop branch again and
// changes nothing — no separate flag to carry, turn into a predicate and back on every iteration.  T is carried as its bit pattern
// (carried as a float, nvcc 12.9 folds the sign-setting arm of the update away — reproduced in isolation — and finished pixels
// resume).  The two constants of expf's range reduction come from the constant bank (exp_loop).  The per-pair arithmetic is
// untouched: colours, final_T and n_contrib stay bit-identical.
// What bounds this kernel (ncu, 3 M Gaussians / 1080p): the FP32 pipe.  ~28 FADD/FMUL/FFMA-class warp instructions per (warp,
// entry) x 10.5 M iterations + the cull arithmetic = ~350 M fp32 warp instructions; the pipe takes one every other cycle per
// sub-partition (`sm__pipe_fma_cycles_active` sits at 50 % in every variant) = 0.60 ms of the measured 0.655 ms.  Three variants
// confirmed it (all bit-identical, all within 4 % of each other): this one (48 instructions per entry instead of 52: same time,
// issue slots 89 -> 84 %); a per-warp work queue as in the backward (45 per entry, but its cull-phase bookkeeping cost as much);
// and packed FFMA2 over two consecutive entries (34 per entry, 36 % fewer fp32 instructions — but an FFMA2 holds the pipe for
// two passes: pipe cycles unchanged, 0.68 ms).  The kernel is at the fp32 roofline of the reference's per-pair arithmetic.
// Staging: the batch is gathered with three 128-bit loads per thread and stored to shared memory.  The TMA path of the backward
// (one cp.async.bulk per 48-byte record, completion on an mbarrier) was measured here as well: 0.73 ms single-buffered, 0.75 ms with
// a two-deep ring (vs 0.66 ms) — 256 small copies per batch serialise in the copy engine, while the load/store path issues them
// from 256 threads at once and this kernel has no other use for the time a ring would hide.
template <bool STATS>
__global__ void __launch_bounds__(256, 6) render_forward_kernel(const uint2* __restrict__ ranges,
	const uint32_t* __restrict__ point_list,
	int W, int H, const float4* __restrict__ rec, const float* __restrict__ bg,
	float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, uint32_t* __restrict__ tile_max,
	int32_t* __restrict__ touched_pixels, float* __restrict__ transmittance)
{
	__shared__ __align__(16) float4 s_rec[256 * 3];
	__shared__ uint32_t s_id[STATS ? 256 : 1];
	__shared__ uint32_t s_max;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile = blockIdx.y * gridDim.x + blockIdx.x;
	const int wx0 = blockIdx.x * GSB_TILE_X + (warp & 1) * 8, wy0 = blockIdx.y * GSB_TILE_Y + (warp >> 1) * 4;
	const int px = wx0 + (lane & 7), py = wy0 + (lane >> 3);
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float rx0 = (float)wx0, rx1 = (float)(wx0 + 7), ry0 = (float)wy0, ry1 = (float)(wy0 + 3);
	const uint2 range = ranges[tile];
	uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_rec);
	asm volatile("" : "+r"(sbase));                       // keep the shared-window address in a register (otherwise re-derived from SR_CgaCtaId every iteration)
	if (tid == 0) s_max = 0;

	uint32_t Tb = inside ? 0x3f800000u : 0xbf800000u;      // bits of T; sign set = finished
	float C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
	uint32_t last = 0;
	for (uint32_t b = range.x; b < range.y; b += 256)
	{
		if (__syncthreads_count((int)Tb < 0) == 256) break;
		const int n = min(256u, range.y - b);
		if (tid < n)
		{
			const uint32_t id = point_list[b + tid];
			const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1], r2 = rec[3 * (size_t)id + 2];
			s_rec[3 * tid] = r0; s_rec[3 * tid + 1] = r1; s_rec[3 * tid + 2] = r2;
			if (STATS) s_id[tid] = id;
		}
		__syncthreads();
		bool warp_done = __all_sync(0xffffffffu, (int)Tb < 0);
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
			uint32_t nbase = (b - range.x) + c0 + 1;                // 1-based list position of the chunk's first entry
			asm volatile("" : "+r"(nbase));                         // (kept in a vector register: `last = nbase + bit` is then one predicated add)
			// Branch-free per-pixel body: in a surviving warp some lane nearly always takes every path of the reference's
			// if/continue chain, so predicating costs nothing and removes the divergence bookkeeping.  The arithmetic and
			// the order of the tests are the reference's (forward.cu:535-569); a masked lane changes no state.
			while (mask)
			{
				const int bit = __ffs(mask) - 1; mask &= mask - 1;
				const uint32_t addr = cbase + bit * SREC_BYTES;
				const float4 r0 = lds128(addr), r1 = lds128(addr + 16);
				const float2 gb = lds64(addr + 32);
				const float T = __uint_as_float(Tb);
				const float dx = __fsub_rn(r1.x, pxf), dy = __fsub_rn(r1.y, pyf);
				const float power = pair_power(r0.x, r0.y, r0.z, dx, dy);
				// power > 0: reference `continue`; power < pth: alpha = opacity*exp(power) is provably < 1/255
				const bool inwin = !(power > 0.0f) && !(power < r0.w);
				const float alpha = fminf(0.99f, __fmul_rn(r1.z, exp_loop(power)));
				const bool cand = inwin && !(alpha < 1.0f / 255.0f);
				const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
				const bool stop = cand && (test_T < 0.0001f);           // also true for every finished pixel (T < 0)
				const bool v = cand && !(test_T < 0.0001f);
				if (STATS)
				{
					const unsigned cm = __ballot_sync(0xffffffffu, v);
					if (cm)
					{
						float ts = v ? T : 0.0f;
#pragma unroll
						for (int o = 16; o > 0; o >>= 1) ts += __shfl_xor_sync(0xffffffffu, ts, o);
						if (lane == 0)
						{
							const uint32_t gid = s_id[c0 + bit];
							atomicAdd(&touched_pixels[gid], (int)__popc(cm));
							atomicAdd(&transmittance[gid], ts);
						}
					}
				}
				if (v)
				{
					C0 = __fmaf_rn(T, __fmul_rn(r1.w, alpha), C0);
					C1 = __fmaf_rn(T, __fmul_rn(gb.x, alpha), C1);
					C2 = __fmaf_rn(T, __fmul_rn(gb.y, alpha), C2);
					Tb = __float_as_uint(test_T);
					last = nbase + bit;
				}
				if (stop) Tb |= 0x80000000u;
			}
			warp_done = __all_sync(0xffffffffu, (int)Tb < 0);
		}
	}
	const float T = __uint_as_float(Tb & 0x7fffffffu);
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
// Backward.  For one Gaussian and the 32 pixels of a warp the reference accumulates 9 sums (backward.cu:561-592).  With
//     u_p = alpha_p * T_p                 (dL/dcolour weight)        w_p = G_p * dL/dalpha_p
// and d = (X - px, Y - py) (X, Y = Gaussian centre, px, py = pixel, all tile-local) they are
//     dL/dcolour_c = sum_p u_p * dLdpix_c(p)            dL/dopacity = sum_p w_p =: M0
//     sum_p w_p dx   = X M0 - Mx        sum_p w_p dx^2  = X^2 M0 - 2X Mx + Mxx      (Mx = sum_p w_p px, ... the moments of w)
// i.e. [Gaussians x pixels] . [pixels x 9 per-pixel constants]: a tiny matrix product.  The kernel therefore stashes
// (w, u) for up to 16 surviving Gaussians per warp and lets the tensor cores do the pixel sums with 3xTF32-split
// m16n8k8 MMAs (fp32-level accuracy: operands are split hi/lo, the per-pixel weights 1, px, py, px^2, px*py, py^2 are exact in
// TF32), instead of a 14-shuffle butterfly + 9 products per (warp, Gaussian).  tcgen05 is not applicable to a per-warp
// 16x32x8 product (it needs 64/128-row tiles from shared-memory descriptors and TMEM); this is the warp-level MMA path.
#define BWD_BATCH 64
#define STASH_LD 36            // row stride of the (w, u) stash: conflict-free A-fragment loads
#define ACC_STRIDE 9           // per staged Gaussian: [dcol0 dcol1 dcol2 M0 Mx My Mxx Mxy Myy]

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
	asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
		: "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo)
{
	hi = __float_as_uint(v) & 0xffffe000u;                    // the MMA reads the top 19 bits: truncation is a valid TF32
	lo = __float_as_uint(v - __uint_as_float(hi));
}

// acc record written for the preprocess backward (12 floats per Gaussian, 48 B): [dcol.r dcol.g dcol.b dop | sx sy cxx cxy | cyy - - -]
// sx = sum dL_dG*dG_ddelx, sy = sum dL_dG*dG_ddely, cxx = sum gdx*dx*dL_dG, cxy = sum gdx*dy*dL_dG, cyy = sum gdy*dy*dL_dG;
// the constant factors (0.5*W, 0.5*H, -0.5) of backward.cu:583-589 are applied once per Gaussian by the consumer.
__device__ __forceinline__ float rcp_approx(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#define BWD_STAGES 4
// Everything one warp owns privately sits in ONE block: a single base address in a register, every member an immediate offset
// (with separate per-member arrays ptxas re-derived five base addresses inside the hot loop once registers ran out).
struct BwdWarp {
	float w[16 * STASH_LD];                // stash of up to 16 surviving Gaussians: w = G * dL/dalpha per pixel lane
	float u[16 * STASH_LD];                //                                         u = alpha * T per pixel lane
	float dlp[3 * 32];                     // dL/dpixel of the warp's 32 pixels, channel-major: B operand of the colour product
	uint32_t rowid[16];                    // Gaussian id of each stashed row (rows outlive their staging buffer; the flush re-reads the record)
	uint8_t queue[BWD_BATCH + 16];         // work queue: batch indices of the entries that survived the warp's cull (+ slack: the loop reads one ahead)
};
#define BW_OFF_U (16 * STASH_LD * 4)
#define BW_OFF_ROWID (2 * 16 * STASH_LD * 4 + 3 * 32 * 4)
#define BW_OFF_QUEUE (BW_OFF_ROWID + 16 * 4)
struct BwdSmem {
	float4 rec[BWD_STAGES][BWD_BATCH * 3]; // ring of staged batches of the tile's list: one 48-byte TMA bulk copy per instance (the record carries its Gaussian id in r2.w)
	uint64_t full[BWD_STAGES];             // mbarrier: the stage's copies have landed (64 arrivals + transaction bytes)
	uint64_t empty[BWD_STAGES];            // mbarrier: all 8 warps are done reading the stage
	float wgt[8 * 32];                     // (1, qx, qy, qx^2, qx*qy, qy^2, 0, 0) of the 32 warp-local pixels: B operand of the moment product
	BwdWarp wp[8];
};
static_assert(offsetof(BwdWarp, u) == BW_OFF_U && offsetof(BwdWarp, rowid) == BW_OFF_ROWID && offsetof(BwdWarp, queue) == BW_OFF_QUEUE, "BwdWarp layout");
static_assert(sizeof(BwdWarp) % 16 == 0, "BwdWarp alignment");


__global__ void __launch_bounds__(256, 4) render_backward_kernel(const uint2* __restrict__ ranges,
	const uint32_t* __restrict__ point_list,
	int W, int H, const float4* __restrict__ rec, const float* __restrict__ bg,
	const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ tile_max,
	const float* __restrict__ dL_dpixels, float* __restrict__ acc)
{
	extern __shared__ __align__(16) unsigned char s_dyn_raw[];
	BwdSmem& S = *reinterpret_cast<BwdSmem*>(s_dyn_raw);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile = blockIdx.y * gridDim.x + blockIdx.x;
	const uint32_t hi = tile_max[tile];
	if (hi == 0) return;
	const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);       // tile-local pixel of this lane
	const int tx0 = blockIdx.x * GSB_TILE_X, ty0 = blockIdx.y * GSB_TILE_Y;
	const int px = tx0 + lx, py = ty0 + ly;
	const bool inside = px < W && py < H;
	const float pxf = (float)px, pyf = (float)py;
	const float rx0 = (float)(tx0 + (warp & 1) * 8), rx1 = rx0 + 7.0f, ry0 = (float)(ty0 + (warp >> 1) * 4), ry1 = ry0 + 3.0f;
	const uint2 range = ranges[tile];
	const size_t pid = (size_t)W * py + px, N = (size_t)W * H;
	uint32_t sbase0 = (uint32_t)__cvta_generic_to_shared(&S.rec[0][0]);
	asm volatile("" : "+r"(sbase0));
	BwdWarp& Wp = S.wp[warp];
	uint32_t wbase = (uint32_t)__cvta_generic_to_shared(&Wp);           // the warp's private block (see BwdWarp)
	asm volatile("" : "+r"(wbase));
	const unsigned lt_mask = (1u << lane) - 1u;

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

	// B fragments (m16n8k8: b0 = B[k = t][n = g], b1 = B[k = t + 4][n = g]; k = pixel lane of the k-step, n = output column) are
	// rebuilt inside flush_rows (a few shuffles per 16 Gaussians) instead of living in 24 registers: occupancy matters more.
	const int fg = lane >> 2, ft = lane & 3;
	float* sw = Wp.w;
	float* su = Wp.u;
	uint32_t nrows = 0;                                                           // warp-uniform: stashed Gaussians
	int buf = 0;

	// Pixel coordinates inside the moment product are WARP-local and centred (qx = lane%8 - 3.5, qy = lane/8 - 1.5: exact in
	// TF32, identical for every warp), and X, Y below are relative to the same centre: the shift back from moments to
	// sum w*dx^2 etc. then cancels as little as possible.
	if (tid < 32)
	{
		const float qx = (float)(tid & 7) - 3.5f, qy = (float)(tid >> 3) - 1.5f;
		S.wgt[0 * 32 + tid] = 1.0f; S.wgt[1 * 32 + tid] = qx; S.wgt[2 * 32 + tid] = qy;
		S.wgt[3 * 32 + tid] = qx * qx; S.wgt[4 * 32 + tid] = qx * qy; S.wgt[5 * 32 + tid] = qy * qy;
		S.wgt[6 * 32 + tid] = 0.0f; S.wgt[7 * 32 + tid] = 0.0f;
	}
	Wp.dlp[lane] = dLp0; Wp.dlp[32 + lane] = dLp1; Wp.dlp[64 + lane] = dLp2;
	const float cxw = rx0 + 3.5f, cyw = ry0 + 1.5f;                                // centre of the warp's 8x4 pixel block
	__syncthreads();

	// Pixel sums of the stashed rows on the tensor cores; lane (g, t = 0) then owns rows g and g + 8: it converts the moments
	// to the reference's sums and issues ONE set of global reductions per (warp, Gaussian) — no shared-memory accumulators,
	// no CTA-wide flush phase, no barrier besides the one that hands over the staging buffers.
	auto flush_rows = [&]() {
		if (nrows == 0) return;
		for (uint32_t r = nrows; r < 16; r++) { sw[r * STASH_LD + lane] = 0.f; su[r * STASH_LD + lane] = 0.f; }
		__syncwarp();
		if (ft == 0)                                                                 // the epilogue re-reads the rows' records: pull them into L1 behind the MMAs
		{
			if (fg < nrows) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 3 * (size_t)Wp.rowid[fg]));
			if (fg + 8 < nrows) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 3 * (size_t)Wp.rowid[fg + 8]));
		}
		float dw[4] = { 0.f, 0.f, 0.f, 0.f }, du[4] = { 0.f, 0.f, 0.f, 0.f };
		const float* dl = Wp.dlp + (fg < 3 ? fg : 0) * 32;
#pragma unroll
		for (int kk = 0; kk < 4; kk++)
		{
			const int c0 = 8 * kk + ft, c1 = c0 + 4;
			// B fragments: b0 = B[k = c0][n = fg], b1 = B[k = c1][n = fg]
			const uint32_t bw0 = __float_as_uint(S.wgt[fg * 32 + c0]), bw1 = __float_as_uint(S.wgt[fg * 32 + c1]);
			uint32_t buh0, buh1, bul0, bul1;
			split_tf32(fg < 3 ? dl[c0] : 0.0f, buh0, bul0);
			split_tf32(fg < 3 ? dl[c1] : 0.0f, buh1, bul1);
			uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
			split_tf32(sw[fg * STASH_LD + c0], h0, l0); split_tf32(sw[(fg + 8) * STASH_LD + c0], h1, l1);
			split_tf32(sw[fg * STASH_LD + c1], h2, l2); split_tf32(sw[(fg + 8) * STASH_LD + c1], h3, l3);
			mma_tf32(dw, h0, h1, h2, h3, bw0, bw1);
			mma_tf32(dw, l0, l1, l2, l3, bw0, bw1);
			split_tf32(su[fg * STASH_LD + c0], h0, l0); split_tf32(su[(fg + 8) * STASH_LD + c0], h1, l1);
			split_tf32(su[fg * STASH_LD + c1], h2, l2); split_tf32(su[(fg + 8) * STASH_LD + c1], h3, l3);
			mma_tf32(du, h0, h1, h2, h3, buh0, buh1);
			mma_tf32(du, l0, l1, l2, l3, buh0, buh1);
			mma_tf32(du, h0, h1, h2, h3, bul0, bul1);
		}
		// D fragment: d[0] = D[g][2t], d[1] = D[g][2t+1], d[2] = D[g+8][2t], d[3] = D[g+8][2t+1]; columns of the w product:
		// (M0 Mx | My Mxx | Mxy Myy) in lanes t = 0 | 1 | 2, of the u product (c0 c1 | c2 -) in lanes t = 0 | 1.
		const int q1 = (lane & ~3) | 1, q2 = (lane & ~3) | 2;
#pragma unroll
		for (int h = 0; h < 2; h++)
		{
			const float My = __shfl_sync(0xffffffffu, dw[2 * h], q1), Mxx = __shfl_sync(0xffffffffu, dw[2 * h + 1], q1);
			const float Mxy = __shfl_sync(0xffffffffu, dw[2 * h], q2), Myy = __shfl_sync(0xffffffffu, dw[2 * h + 1], q2);
			const float c2 = __shfl_sync(0xffffffffu, du[2 * h], q1);
			const uint32_t row = fg + 8 * h;
			if (ft == 0 && row < nrows)
			{
				const uint32_t gid = Wp.rowid[row];
				const float4 r0 = __ldg(rec + 3 * (size_t)gid), r1 = __ldg(rec + 3 * (size_t)gid + 1);      // L2-resident: staged moments ago
				const float X = r1.x - cxw, Y = r1.y - cyw, o = r1.z;
				const float M0 = dw[2 * h], Mx = dw[2 * h + 1];
				const float Sx = X * M0 - Mx, Sy = Y * M0 - My;
				const float Sxx = X * X * M0 - 2.0f * X * Mx + Mxx;
				const float Sxy = X * Y * M0 - X * My - Y * Mx + Mxy;
				const float Syy = Y * Y * M0 - 2.0f * Y * My + Myy;
				float* a = acc + 12 * (size_t)gid;
				red_add_v4(a, du[2 * h], du[2 * h + 1], c2, M0);
				red_add_v4(a + 4, -o * (r0.x * Sx + r0.y * Sy), -o * (r0.z * Sy + r0.y * Sx), o * Sxx, o * Sxy);
				atomicAdd(a + 8, o * Syy);
			}
		}
		__syncwarp();
		nrows = 0;
	};

	// ---- staging ring: warps do NOT run in lockstep.  A batch of 96 list entries is gathered by 96 threads, one 48-byte TMA
	// bulk copy (cp.async.bulk, completion by mbarrier transaction bytes) per entry, BWD_STAGES - 1 batches ahead of the
	// slowest warp; a warp that finishes a batch early moves on to the next stage instead of waiting at a CTA barrier.
	auto list_id = [&](uint32_t bi) -> uint32_t {                                  // entry `tid` of batch bi (counted from the back of the list)
		const uint32_t k = bi * BWD_BATCH + tid;
		return (tid < BWD_BATCH && k < hi) ? point_list[range.x + (hi - 1 - k)] : 0xffffffffu;
	};
	auto issue = [&](uint32_t bi, uint32_t id) {                                   // threads tid < BWD_BATCH
		const int st = bi % BWD_STAGES;
		if (id != 0xffffffffu)
		{
			mbar_arrive_expect_tx(&S.full[st], 48u);
			tma_bulk_g2s(&S.rec[st][3 * tid], rec + 3 * (size_t)id, 48u, &S.full[st]);
		}
		else mbar_arrive(&S.full[st]);
	};
	const uint32_t nb = (hi + BWD_BATCH - 1) / BWD_BATCH;
	if (tid == 0)
	{
		for (int k = 0; k < BWD_STAGES; k++) { mbar_init(&S.full[k], BWD_BATCH); mbar_init(&S.empty[k], 8); }
		mbar_fence_init();
	}
	__syncthreads();
	uint32_t id_next = 0xffffffffu;
	if (tid < BWD_BATCH)
	{
		for (uint32_t k = 0; k < BWD_STAGES - 1 && k < nb; k++) issue(k, list_id(k));
		id_next = list_id(BWD_STAGES - 1);
	}
	for (uint32_t bi = 0; bi < nb; bi++)
	{
		const uint32_t b = bi * BWD_BATCH;
		const int n = min((uint32_t)BWD_BATCH, hi - b);
		buf = bi % BWD_STAGES;
		if (tid < BWD_BATCH && bi + BWD_STAGES - 1 < nb)
		{
			// the stage that batch bi + STAGES - 1 goes into was last read for batch bi - 1
			if (bi >= 1) mbar_wait(&S.empty[(bi - 1) % BWD_STAGES], ((bi - 1) / BWD_STAGES) & 1u);
			issue(bi + BWD_STAGES - 1, id_next);
			id_next = list_id(bi + BWD_STAGES);
		}
		mbar_wait(&S.full[buf], (bi / BWD_STAGES) & 1u);
		const uint32_t sbase = sbase0 + (uint32_t)(buf * BWD_BATCH * 3) * 16u;
		// ---- cull the batch into the warp's work queue (see render_forward_kernel) ----
		uint32_t qn = 0;
		for (int c0 = 0; c0 < n; c0 += 32)
		{
			const int j = c0 + lane;
			bool keep = false;
			if (j < n && (hi - 1 - b - j) < wmax)
			{
				const float4 r0 = lds128(sbase + j * SREC_BYTES), r1 = lds128(sbase + j * SREC_BYTES + 16);
				keep = rect_may_contribute(r1.x, r1.y, r0.x, r0.y, r0.z, r0.w, rx0, rx1, ry0, ry1);
			}
			const unsigned mask = __ballot_sync(0xffffffffu, keep);
			if (keep) Wp.queue[qn + __popc(mask & lt_mask)] = (uint8_t)j;
			qn += __popc(mask);
		}
		__syncwarp();
		uint32_t posb = hi - 1 - b;                                            // list position of batch entry 0 (entries run backwards)
		asm volatile("" : "+r"(posb));
		uint32_t jj = 0, jnext;
		if (qn) asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(jj) : "r"(wbase), "n"(BW_OFF_QUEUE));
		for (uint32_t q = 0; q < qn; q++, jj = jnext)
		{
			asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(jnext) : "r"(wbase + q), "n"(BW_OFF_QUEUE + 1));   // next entry's index: off the critical path
			const uint32_t pos = posb - jj;
			const uint32_t addr = sbase + jj * SREC_BYTES;
			const float4 r0 = lds128(addr), r1 = lds128(addr + 16);
			const float dx = __fsub_rn(r1.x, pxf), dy = __fsub_rn(r1.y, pyf);
			const float power = pair_power(r0.x, r0.y, r0.z, dx, dy);
			const float G = exp_loop(power);
			const float alpha = fminf(0.99f, __fmul_rn(r1.z, G));
			// backward.cu:524-539: same skips as the forward (pos < last_contributor replaces the `contributor` countdown)
			const bool active = (pos < last_contributor) && !(power > 0.0f) && !(power < r0.w) && !(alpha < 1.0f / 255.0f);
			if (!__any_sync(0xffffffffu, active)) continue;
			float wv = 0.f, uv = 0.f;
			const float4 r2 = lds128(addr + 32);                               // (G, B, depth, Gaussian id)
			if (active)
			{
				// backward.cu:541 T = T / (1 - alpha): 1 - alpha lies in [0.01, 1], MUFU.RCP (1 ulp) is ample for a gradient that is
				// tolerance-compared; the IEEE-rounded reciprocal costs 12 more instructions per pair (range check + Newton step)
				const float inv = rcp_approx(1.0f - alpha);
				T = T * inv;
				uv = alpha * T;
				const float cr = r1.w, cg = r2.x, cb = r2.y;
				const float oml = 1.0f - last_alpha;
				ar0 = last_alpha * lc0 + oml * ar0; lc0 = cr;
				ar1 = last_alpha * lc1 + oml * ar1; lc1 = cg;
				ar2 = last_alpha * lc2 + oml * ar2; lc2 = cb;
				float dL_dalpha = (cr - ar0) * dLp0 + (cg - ar1) * dLp1 + (cb - ar2) * dLp2;
				dL_dalpha *= T;
				last_alpha = alpha;
				dL_dalpha += (-T_final * inv) * bg_dot_dpixel;                      // backward.cu:569-572
				wv = G * dL_dalpha;
			}
			{
				const uint32_t sa = wbase + (nrows * STASH_LD + lane) * 4;
				asm volatile("st.shared.f32 [%0], %1;" ::"r"(sa), "f"(wv) : "memory");
				asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(sa), "f"(uv), "n"(BW_OFF_U) : "memory");
				if (lane == 0) asm volatile("st.shared.f32 [%0+%2], %1;" ::"r"(wbase + nrows * 4), "f"(r2.w), "n"(BW_OFF_ROWID) : "memory");
			}
			nrows++;
			if (nrows == 16) flush_rows();
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&S.empty[buf]);           // this warp no longer reads the stage (stashed rows carry their own data)
	}
	flush_rows();                                            // only the tile's tail is a partial block
}

// ------------------------------------------------------------------------------------------------
int launch_render_forward(const ImageState& img, const BinningState& b, const GeomState& g, int W, int H, const float* bg,
	float* out_color, int32_t* touched_pixels, float* transmittance, cudaStream_t stream)
{
	const dim3 grid((W + GSB_TILE_X - 1) / GSB_TILE_X, (H + GSB_TILE_Y - 1) / GSB_TILE_Y);
	ProfScope prof(K_RENDER_FWD, stream);
	if (touched_pixels && transmittance)
		render_forward_kernel<true><<<grid, 256, 0, stream>>>(img.ranges, b.point_list, W, H, g.rec, bg,
			img.final_T, img.n_contrib, out_color, img.tile_max_contrib, touched_pixels, transmittance);
	else
		render_forward_kernel<false><<<grid, 256, 0, stream>>>(img.ranges, b.point_list, W, H, g.rec, bg,
			img.final_T, img.n_contrib, out_color, img.tile_max_contrib, nullptr, nullptr);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

int launch_render_backward(const ImageState& img, const BinningState& b, const GeomState& g, int P, int W, int H, const float* bg,
	const float* dL_dpix, float* acc, cudaStream_t stream)
{
	const dim3 grid((W + GSB_TILE_X - 1) / GSB_TILE_X, (H + GSB_TILE_Y - 1) / GSB_TILE_Y);
	if (int e = ensure_dyn_smem((const void*)render_backward_kernel, (int)sizeof(BwdSmem))) return e;
	ProfScope prof(K_RENDER_BWD, stream);
	// the per-Gaussian accumulator the kernel reduces into (12 floats per Gaussian, inside the geometry blob)
	GSB_CUDA_OK(cudaMemsetAsync(acc, 0, size_t(P) * 48, stream));
	render_backward_kernel<<<grid, 256, sizeof(BwdSmem), stream>>>(img.ranges, b.point_list, W, H, g.rec, bg,
		img.final_T, img.n_contrib, img.tile_max_contrib, dL_dpix, acc);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

} // namespace gsb
