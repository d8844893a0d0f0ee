// gsb_binning.cu — tile binning: prefix sum, key emission, stable LSD radix sort, tile ranges (sm_100a).
//
// Replaces, with hand-written kernels (no CUB):
//   cub::DeviceScan::InclusiveSum           rasterizer_impl.cu:441   -> scan_kernel (single pass, decoupled look-back)
//   duplicateWithKeys                       rasterizer_impl.cu:78-119 -> emit_keys_kernel
//   cub::DeviceRadixSort::SortPairs         rasterizer_impl.cu:468   -> sort_hist / sort_plan / sort_pass (onesweep:
//                                            one histogram sweep, then one read+write sweep per 8-bit digit;
//                                            digits in which every key agrees are skipped on the device)
//   cudaMemset + identifyTileRanges         rasterizer_impl.cu:475-482 -> tile_ranges_kernel
// All of it is integer work and bit-exact by construction: keys are (tile << 32 | depth bits), the sort is
// stable, so ties keep emission order (ascending Gaussian index).
#include "gsb_common.cuh"

namespace gsb {

// ------------------------------------------------------------------------------------------------
// Inclusive scan of tiles_touched, 2048 items per block, chained through 64-bit look-back cells
// (flag << 32 | value; flag 1 = block aggregate, 2 = inclusive prefix).  Block order = ticket order.
#define SCAN_ITEMS 8
__global__ void __launch_bounds__(256) scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n,
	unsigned long long* state, uint32_t* counters)
{
	__shared__ uint32_t s_warp[8];
	__shared__ uint32_t s_block, s_prefix;
	if (threadIdx.x == 0) s_block = atomicAdd(&counters[2], 1u);
	__syncthreads();
	const uint32_t bid = s_block;
	const long long base = (long long)bid * (256 * SCAN_ITEMS) + threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS];
	uint32_t sum = 0;
	if (base + SCAN_ITEMS <= n)
	{
		const uint4 a = reinterpret_cast<const uint4*>(in + base)[0], b = reinterpret_cast<const uint4*>(in + base)[1];
		v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
	}
	else
	{
#pragma unroll
		for (int i = 0; i < SCAN_ITEMS; i++) v[i] = (base + i < n) ? in[base + i] : 0u;
	}
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; i++) { sum += v[i]; v[i] = sum; }
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	uint32_t warp_excl = 0, block_total = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) { const uint32_t t = s_warp[w]; if (w < warp) warp_excl += t; block_total += t; }
	if (threadIdx.x == 0)
	{
		uint32_t prefix = 0;
		if (bid > 0)
		{
			atomicExch(&state[bid], (1ull << 32) | block_total);
			long long j = (long long)bid - 1;
			while (true)
			{
				unsigned long long c;
				do { c = *reinterpret_cast<volatile unsigned long long*>(&state[j]); } while ((c >> 32) == 0);
				prefix += (uint32_t)c;
				if ((c >> 32) == 2) break;
				j--;
			}
		}
		__threadfence();
		atomicExch(&state[bid], (2ull << 32) | (uint32_t)(prefix + block_total));
		s_prefix = prefix;
		if ((long long)(bid + 1) * (256 * SCAN_ITEMS) >= n) counters[0] = prefix + block_total;   // num_rendered
	}
	__syncthreads();
	const uint32_t off = s_prefix + warp_excl + (incl - sum);
	if (base + SCAN_ITEMS <= n)
	{
		reinterpret_cast<uint4*>(out + base)[0] = make_uint4(v[0] + off, v[1] + off, v[2] + off, v[3] + off);
		reinterpret_cast<uint4*>(out + base)[1] = make_uint4(v[4] + off, v[5] + off, v[6] + off, v[7] + off);
	}
	else
	{
#pragma unroll
		for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) out[base + i] = v[i] + off;
	}
}

// ------------------------------------------------------------------------------------------------
// duplicateWithKeys: one warp per 32 Gaussians; a Gaussian's tiles are written by the whole warp when it
// covers many tiles (no single-thread serial loop over a large splat), otherwise by its own lane.
__global__ void __launch_bounds__(256) emit_keys_kernel(int P, const float4* __restrict__ rec, const uint2* __restrict__ rect,
	const uint32_t* __restrict__ tiles_touched, const uint32_t* __restrict__ offsets, int gx,
	uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	uint32_t t = 0, off = 0, dbits = 0; uint2 rc = make_uint2(0, 0);
	if (idx < P)
	{
		t = tiles_touched[idx];
		if (t)
		{
			off = offsets[idx] - t;
			rc = rect[idx];
			dbits = __float_as_uint(rec[3 * (size_t)idx + 2].z);
		}
	}
	const uint32_t minx = rc.x & 0xffffu, maxx = rc.x >> 16, miny = rc.y & 0xffffu;
	const uint32_t w = maxx - minx;
	const bool big = t > 16;
	if (t && !big)
	{
		uint32_t x = minx, y = miny;
		for (uint32_t k = 0; k < t; k++)
		{
			keys[off + k] = ((uint64_t)(y * gx + x) << 32) | dbits;
			vals[off + k] = (uint32_t)idx;
			if (++x == maxx) { x = minx; y++; }
		}
	}
	unsigned bigmask = __ballot_sync(0xffffffffu, big);
	while (bigmask)
	{
		const int src = __ffs(bigmask) - 1; bigmask &= bigmask - 1;
		const uint32_t bt = __shfl_sync(0xffffffffu, t, src), boff = __shfl_sync(0xffffffffu, off, src);
		const uint32_t bminx = __shfl_sync(0xffffffffu, minx, src), bminy = __shfl_sync(0xffffffffu, miny, src);
		const uint32_t bw = __shfl_sync(0xffffffffu, w, src), bd = __shfl_sync(0xffffffffu, dbits, src);
		const uint32_t bidx = (uint32_t)(idx - lane + src);
		for (uint32_t k = lane; k < bt; k += 32)
		{
			const uint32_t y = bminy + k / bw, x = bminx + k % bw;
			keys[boff + k] = ((uint64_t)(y * gx + x) << 32) | bd;
			vals[boff + k] = bidx;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Radix sort, 8-bit digits.
__global__ void __launch_bounds__(256) sort_hist_kernel(const uint64_t* __restrict__ keys, long long R, int passes, uint32_t* __restrict__ hist)
{
	__shared__ uint32_t s_h[GSB_SORT_MAX_PASSES * 256];
	for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) s_h[i] = 0;
	__syncthreads();
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (long long)gridDim.x * blockDim.x)
	{
		const uint64_t k = keys[i];
		for (int p = 0; p < passes; p++) atomicAdd(&s_h[p * 256 + (uint32_t)((k >> (8 * p)) & 0xff)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) { const uint32_t c = s_h[i]; if (c) atomicAdd(&hist[i], c); }
}

// One block: exclusive digit offsets per pass, skip flags (all keys share the digit) and ping-pong schedule.
__global__ void __launch_bounds__(256) sort_plan_kernel(const uint32_t* __restrict__ hist, long long R, int passes, SortPlan* plan)
{
	__shared__ uint32_t s_scan[256];
	__shared__ uint32_t s_skip[GSB_SORT_MAX_PASSES];
	const int d = threadIdx.x;
	if (d < GSB_SORT_MAX_PASSES) s_skip[d] = 0;
	__syncthreads();
	for (int p = 0; p < passes; p++)
	{
		const uint32_t c = hist[p * 256 + d];
		if (c == (uint32_t)R) s_skip[p] = 1;           // at most one thread per pass can see this
		if (d == 0 && R == 0) s_skip[p] = 1;
		s_scan[d] = c;
		__syncthreads();
		for (int o = 1; o < 256; o <<= 1)
		{
			const uint32_t t = d >= o ? s_scan[d - o] : 0u;
			__syncthreads();
			s_scan[d] += t;
			__syncthreads();
		}
		plan->digit_base[p][d] = s_scan[d] - c;
		__syncthreads();
	}
	if (d == 0)
	{
		uint32_t cur = 0;
		for (int p = 0; p < passes; p++)
		{
			const uint32_t sk = s_skip[p] == 1;
			plan->skip[p] = sk; plan->src[p] = cur;
			if (!sk) cur ^= 1u;
		}
		plan->final_buf = cur;
	}
}

__global__ void sort_plan_init_kernel(SortPlan* plan)
{
	if (threadIdx.x < GSB_SORT_MAX_PASSES) { plan->skip[threadIdx.x] = 0; plan->src[threadIdx.x] = 0; }
	if (threadIdx.x == 0) plan->final_buf = 0;
}

// One onesweep pass.  Tile = 4096 consecutive keys handled by 256 threads; warp w ranks keys
// [w*512, (w+1)*512) in 16 warp-wide steps with match.any (stable), then digit counts are chained across tiles.
#define SORT_THREADS 256
#define SORT_ITEMS 16
#define LB_AGG 0x40000000u
#define LB_INC 0x80000000u
#define LB_VAL 0x3fffffffu
__global__ void __launch_bounds__(SORT_THREADS) sort_pass_kernel(uint64_t* keys0, uint64_t* keys1, uint32_t* vals0, uint32_t* vals1,
	long long R, int pass, const SortPlan* __restrict__ plan, uint32_t* lookback_all, uint32_t* tickets, size_t n_tiles)
{
	if (plan->skip[pass]) return;
	const uint32_t srcb = plan->src[pass];
	const uint64_t* __restrict__ kin = srcb ? keys1 : keys0;
	uint64_t* __restrict__ kout = srcb ? keys0 : keys1;
	const uint32_t* __restrict__ vin = srcb ? vals1 : vals0;
	uint32_t* __restrict__ vout = srcb ? vals0 : vals1;
	uint32_t* lookback = lookback_all + (size_t)pass * n_tiles * 256;

	__shared__ uint32_t s_whist[8][256];           // per-warp digit counts, later per-warp exclusive offsets
	__shared__ uint32_t s_dstart[256];             // tile-local start of each digit run
	__shared__ uint32_t s_gbase[256];              // global start of this tile's run of each digit
	__shared__ uint64_t s_keys[GSB_SORT_TILE];
	__shared__ uint32_t s_tile;
	uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys);   // reused after the key write-out

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(&tickets[pass], 1u);
	for (int i = tid; i < 8 * 256; i += SORT_THREADS) (&s_whist[0][0])[i] = 0;
	__syncthreads();
	const uint32_t tile = s_tile;
	const long long tbase = (long long)tile * GSB_SORT_TILE;
	const int count = (int)min((long long)GSB_SORT_TILE, R - tbase);
	const int shift = 8 * pass;

	uint64_t key[SORT_ITEMS];
	uint32_t rank[SORT_ITEMS];
	const unsigned lt = (1u << lane) - 1u;
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++)
	{
		const int local = warp * (32 * SORT_ITEMS) + i * 32 + lane;
		const bool valid = local < count;
		key[i] = valid ? kin[tbase + local] : ~0ull;
		const uint32_t d = valid ? (uint32_t)((key[i] >> shift) & 0xff) : 256u;
		const unsigned vmask = __ballot_sync(0xffffffffu, valid);
		unsigned m = __match_any_sync(0xffffffffu, d) & vmask;
		if (valid)
		{
			const int leader = __ffs(m) - 1;
			uint32_t old = 0;
			if (lane == leader) { old = s_whist[warp][d]; s_whist[warp][d] = old + __popc(m); }
			old = __shfl_sync(m, old, leader);
			rank[i] = old + __popc(m & lt);
		}
		__syncwarp();
	}
	__syncthreads();
	// digit `tid`: per-warp exclusive offsets + tile total
	uint32_t total = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) { const uint32_t c = s_whist[w][tid]; s_whist[w][tid] = total; total += c; }
	// chained scan over tiles (decoupled look-back) for digit `tid`
	uint32_t excl = 0;
	if (tile == 0) lookback[tid] = LB_INC | total;
	else
	{
		atomicExch(&lookback[(size_t)tile * 256 + tid], LB_AGG | total);
		long long j = (long long)tile - 1;
		while (true)
		{
			uint32_t c;
			do { c = *reinterpret_cast<volatile uint32_t*>(&lookback[(size_t)j * 256 + tid]); } while (c == 0);
			excl += c & LB_VAL;
			if (c & LB_INC) break;
			j--;
		}
		atomicExch(&lookback[(size_t)tile * 256 + tid], LB_INC | (excl + total));
	}
	// tile-local exclusive scan of digit totals (256 entries) -> s_dstart
	s_dstart[tid] = total;
	__syncthreads();
	for (int o = 1; o < 256; o <<= 1)
	{
		const uint32_t t = tid >= o ? s_dstart[tid - o] : 0u;
		__syncthreads();
		s_dstart[tid] += t;
		__syncthreads();
	}
	const uint32_t dstart = s_dstart[tid] - total;
	__syncthreads();
	s_dstart[tid] = dstart;
	s_gbase[tid] = plan->digit_base[pass][tid] + excl - dstart;    // global index = s_gbase[d] + tile-local sorted position
	__syncthreads();
	// scatter keys into tile-local sorted order
	uint32_t pos[SORT_ITEMS];
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++)
	{
		const int local = warp * (32 * SORT_ITEMS) + i * 32 + lane;
		if (local < count)
		{
			const uint32_t d = (uint32_t)((key[i] >> shift) & 0xff);
			pos[i] = s_dstart[d] + s_whist[warp][d] + rank[i];
			s_keys[pos[i]] = key[i];
		}
	}
	__syncthreads();
	uint32_t gpos[SORT_ITEMS];
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++)
	{
		const int p = i * SORT_THREADS + tid;
		if (p < count)
		{
			const uint64_t k = s_keys[p];
			gpos[i] = s_gbase[(uint32_t)((k >> shift) & 0xff)] + p;
			kout[gpos[i]] = k;
		}
	}
	__syncthreads();
	// values: load in the original arrangement, route through the same tile-local positions
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++)
	{
		const int local = warp * (32 * SORT_ITEMS) + i * 32 + lane;
		if (local < count) s_vals[pos[i]] = vin[tbase + local];
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++)
	{
		const int p = i * SORT_THREADS + tid;
		if (p < count) vout[gpos[i]] = s_vals[p];
	}
}

// identifyTileRanges (rasterizer_impl.cu:124-146); ranges zeroed by the caller (cudaMemsetAsync).
__global__ void __launch_bounds__(256) tile_ranges_kernel(long long L, const uint64_t* keys0, const uint64_t* keys1,
	const SortPlan* __restrict__ plan, uint2* __restrict__ ranges)
{
	const uint64_t* __restrict__ keys = plan->final_buf ? keys1 : keys0;
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= L) return;
	const uint32_t cur = (uint32_t)(keys[idx] >> 32);
	if (idx == 0) ranges[cur].x = 0;
	else
	{
		const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
		if (cur != prev) { ranges[prev].y = (uint32_t)idx; ranges[cur].x = (uint32_t)idx; }
	}
	if (idx == L - 1) ranges[cur].y = (uint32_t)L;
}

// ------------------------------------------------------------------------------------------------
static uint32_t higher_msb(uint32_t n)        // rasterizer_impl.cu:41-58 getHigherMsb
{
	uint32_t msb = sizeof(n) * 4, step = msb;
	while (step > 1) { step /= 2; if (n >> msb) msb += step; else msb -= step; }
	if (n >> msb) msb++;
	return msb;
}

int launch_scan(const GeomState& g, int P, cudaStream_t stream)
{
	const int blocks = (P + 256 * SCAN_ITEMS - 1) / (256 * SCAN_ITEMS);
	ProfScope prof(K_SCAN, stream);
	scan_kernel<<<blocks, 256, 0, stream>>>(g.tiles_touched, g.point_offsets, P, g.scan_state, g.counters);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

int launch_binning(const GeomState& g, const BinningState& b, char* bin_blob, const ImageState& img, int P, long long R, int W, int H, cudaStream_t stream)
{
	const int gx = (W + GSB_TILE_X - 1) / GSB_TILE_X, gy = (H + GSB_TILE_Y - 1) / GSB_TILE_Y;
	GSB_CUDA_OK(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)gx * gy, stream));
	GSB_CUDA_OK(cudaMemsetAsync(bin_blob + b.zero_begin, 0, b.zero_bytes, stream));
	sort_plan_init_kernel<<<1, 32, 0, stream>>>(b.plan);
	GSB_LAUNCHED();
	if (R == 0) { GSB_CUDA_OK(cudaGetLastError()); return GSB_OK; }
	{ ProfScope prof(K_EMIT_KEYS, stream);
	emit_keys_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g.rec, g.rect, g.tiles_touched, g.point_offsets, gx, b.keys[0], b.vals[0]); }
	GSB_LAUNCHED();
	const int bits = 32 + (int)higher_msb((uint32_t)(gx * gy));                        // rasterizer_impl.cu:465-473
	const int passes = (bits + 7) / 8;
	const size_t n_tiles = BinningState::sort_tiles(R);
	const int hist_blocks = (int)((n_tiles < 148 * 8) ? n_tiles : 148 * 8);
	{ ProfScope prof(K_SORT_HIST, stream);
	sort_hist_kernel<<<hist_blocks, 256, 0, stream>>>(b.keys[0], R, passes, b.hist); }
	GSB_LAUNCHED();
	{ ProfScope prof(K_SORT_PLAN, stream);
	sort_plan_kernel<<<1, 256, 0, stream>>>(b.hist, R, passes, b.plan); }
	GSB_LAUNCHED();
	for (int p = 0; p < passes; p++)
	{
		ProfScope prof(K_SORT_PASS, stream);
		sort_pass_kernel<<<(unsigned)n_tiles, SORT_THREADS, 0, stream>>>(b.keys[0], b.keys[1], b.vals[0], b.vals[1], R, p, b.plan,
			b.lookback, b.tickets, n_tiles);
		GSB_LAUNCHED();
	}
	{ ProfScope prof(K_TILE_RANGES, stream);
	tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, b.keys[0], b.keys[1], b.plan, img.ranges); }
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

} // namespace gsb
