// gsb_backward.cu — per-Gaussian backward of the preprocess stage (sm_100a).
//
// One fused kernel replaces reference backward.cu:177-307 computeCov2DCUDA + backward.cu:380-434 preprocessCUDA
// (with :20-172 computeColorFromSH and :311-374 computeCov3D), the nonZeroMask/cub::DeviceReduce/cudaMalloc/D2H
// sequence of rasterizer_impl.cu:549-571 (the visible count was produced by the forward preprocess), and the nine
// torch::zeros of rasterize_points.cu:259-267: every output element is written exactly once (zeros for culled
// Gaussians and inactive SH bands), so the caller allocates with torch.empty and nothing is memset.
// Gradients are fp32 and tolerance-compared (the reference's atomicAdd order makes its own bits non-deterministic).
#include "gsb_common.cuh"

namespace gsb {

struct BwdArgs {
	int P, M, W, H;
	float mod, tan_fovx, tan_fovy, focal_x, focal_y, lambda;
	const float* means3D; const float* scales; const float* rotations; const float* cov3D_precomp;
	const float* shs; const float* colors_precomp; const int32_t* degrees; const int32_t* radii;
	const float* view; const float* proj; const float* campos;
	int quant; GsbQuant q;
	GeomState g; const float* acc;
	GsbGrads out;
};

// ACC (view-batch accumulation): the sums are formed by the L2 with fire-and-forget reductions (RED.ADD, no value returns to
// the SM): a load-add-store in the kernel serialises one DRAM round trip per output element behind the previous store
// (measured: preprocess backward 0.30 -> 0.76 ms at 3 M Gaussians).  One kernel per view runs at a time on the stream, so every
// element receives exactly one addition per view, in view order: the result is deterministic.  Adding zero is skipped — culled
// Gaussians and inactive SH bands are most of the rows; the overwrite mode stores them (every element written once, no memset).
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d)
{
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <bool ACC> __device__ __forceinline__ void put(float* p, float v) { if (ACC) { if (v != 0.0f) atomicAdd(p, v); } else *p = v; }

// Coalesced store of one small per-Gaussian output ([P,WD]) for the 32 Gaussians of a warp: lanes park their WD values in
// shared memory, then the warp writes the 32*WD contiguous floats with unit-stride stores.
template <int WD, bool ACC>
__device__ __forceinline__ void warp_store(float* __restrict__ dst, long long base, int n_valid, const float* v, float* tmp, int lane)
{
#pragma unroll
	for (int k = 0; k < WD; k++) tmp[lane * WD + k] = v[k];
	__syncwarp();
#pragma unroll
	for (int i = 0; i < WD; i++)
	{
		const int f = i * 32 + lane;
		if (f < n_valid * WD) put<ACC>(dst + base * WD + f, tmp[f]);
	}
	__syncwarp();
}

// One warp per 32 consecutive Gaussians (lane = Gaussian).  The SH rows of the warp (32 x 3M floats, contiguous in HBM) are
// staged through shared memory with unit-stride loads, overwritten in place by the SH gradients and written back with
// unit-stride stores; the seven small outputs go through warp_store.  Every output element is written exactly once.
template <bool QUANT, bool ACC>
__global__ void __launch_bounds__(256) preprocess_backward_kernel(const BwdArgs a)
{
	extern __shared__ float s_dyn[];
	float* s_cb = s_dyn;                                                  // QUANT: [20][256]
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int RL = 3 * a.M, RS = RL + 1;                                  // row length / padded stride (conflict-free per-lane rows)
	float* s_row = s_dyn + (QUANT ? GSB_NUM_CODEBOOKS * GSB_CODEBOOK_SIZE : 0) + warp * (32 * RS + 32 * 6);
	float* s_tmp = s_row + 32 * RS;
	if (QUANT)
	{
		for (int i = threadIdx.x; i < GSB_NUM_CODEBOOKS * GSB_CODEBOOK_SIZE; i += blockDim.x)
		{
			float v = a.q.centers[i];
			if (i / GSB_CODEBOOK_SIZE == 17) v = exp_ref(v);
			s_cb[i] = v;
		}
		__syncthreads();
	}
	// rasterizer_impl.cu:549-571: sh_sparsity_multiplier = lambda / (n_visible * 15 * 3)
	const float mult = a.lambda != 0.0f ? a.lambda / (float)((int)a.g.counters[1] * 15 * 3) : 0.0f;
	// colours given by the caller (override_color): the SH coefficients were not used by the forward, their gradient is zero
	const bool have_sh = (QUANT || a.shs != nullptr) && a.out.dL_dsh != nullptr && a.colors_precomp == nullptr;
	const bool have_scales = QUANT || a.scales != nullptr;
	for (long long base = ((long long)blockIdx.x * 8 + warp) * 32; base < a.P; base += (long long)gridDim.x * 8 * 32)
	{
		const long long idx = base + lane;
		const int n_valid = (int)min((long long)32, a.P - base);
		const bool valid = lane < n_valid;
		const bool vis = valid && a.radii[idx] > 0;
		// ---- stage the warp's SH rows -------------------------------------------------------------
		if (have_sh && !QUANT && RL == 48 && n_valid == 32)
		{
			// M == 16 fast path: 12 independent 128-bit loads per lane (6 KB contiguous per warp), then scatter to padded rows
			const float4* src4 = reinterpret_cast<const float4*>(a.shs + base * 48);
			float4 v[12];
#pragma unroll
			for (int i = 0; i < 12; i++) v[i] = src4[i * 32 + lane];
#pragma unroll
			for (int i = 0; i < 12; i++)
			{
				const int f = (i * 32 + lane) * 4, row = f / 48, col = f - row * 48;
				float* d = s_row + row * RS + col;
				d[0] = v[i].x; d[1] = v[i].y; d[2] = v[i].z; d[3] = v[i].w;
			}
		}
		else if (have_sh && !QUANT)
		{
			const float* src = a.shs + base * RL;
			int row = 0, col = lane;
			while (col >= RL) { col -= RL; row++; }
			for (int f = lane; f < n_valid * RL; f += 32)
			{
				s_row[row * RS + col] = src[f];
				col += 32;
				while (col >= RL) { col -= RL; row++; }
			}
		}
		__syncwarp();
		float o_m2[3] = { 0, 0, 0 }, o_col[3] = { 0, 0, 0 }, o_m3[3] = { 0, 0, 0 }, o_cov[6] = { 0, 0, 0, 0, 0, 0 };
		float o_sc[3] = { 0, 0, 0 }, o_rot[4] = { 0, 0, 0, 0 }, o_con[4] = { 0, 0, 0, 0 }, o_op[1] = { 0 };
		float* myrow = s_row + lane * RS;
		// every per-Gaussian input is requested before the visibility flag is known (the flag is itself a load): radius -> accumulator
		// -> position -> ids used to be a chain of DRAM round trips; a culled Gaussian now costs ~90 wasted bytes
		float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0; float cyy = 0.f, mx = 0.f, my = 0.f, mz = 0.f, opac = 0.f;
		float sc[3] = { 0, 0, 0 }, qr = 1, qx = 0, qy = 0, qz = 0;
		uint32_t isb[3] = { 0, 0, 0 }, irw = 0; int deg_in = 0; unsigned cl_in = 0;
		if (valid)
		{
			acc0 = reinterpret_cast<const float4*>(a.acc)[3 * idx];
			acc1 = reinterpret_cast<const float4*>(a.acc)[3 * idx + 1];
			cyy = a.acc[12 * idx + 8];
			mx = a.means3D[3 * idx]; my = a.means3D[3 * idx + 1]; mz = a.means3D[3 * idx + 2];
			opac = a.g.rec[3 * idx + 1].z;
			if (QUANT)
			{
				const uint8_t* is = a.q.ids_scaling + 3 * idx;
				isb[0] = is[0]; isb[1] = is[1]; isb[2] = is[2];
				irw = reinterpret_cast<const uint32_t*>(a.q.ids_rot)[idx];
			}
			else if (!a.cov3D_precomp)
			{
				for (int k = 0; k < 3; k++) sc[k] = a.scales[3 * idx + k];
				const float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
				qr = q.x; qx = q.y; qy = q.z; qz = q.w;
			}
			if (have_sh) { deg_in = a.degrees[idx]; cl_in = a.g.clamped[idx]; }
		}
		if (!vis)
		{
			if (have_sh) for (int k = 0; k < RL; k++) myrow[k] = 0.f;
		}
		else
		{
			// constant factors of backward.cu:498-499, 583-589 applied once per Gaussian
			const float g2x = acc1.x * (0.5f * a.W), g2y = acc1.y * (0.5f * a.H);
			const float dconx = -0.5f * acc1.z, dcony = -0.5f * acc1.w, dconz = -0.5f * cyy;
			float cov3D[6];
			if (QUANT)
			{
				for (int k = 0; k < 3; k++) sc[k] = s_cb[17 * 256 + isb[k]];
				qr = s_cb[18 * 256 + (irw & 0xffu)]; qx = s_cb[19 * 256 + ((irw >> 8) & 0xffu)]; qy = s_cb[19 * 256 + ((irw >> 16) & 0xffu)]; qz = s_cb[19 * 256 + (irw >> 24)];
				normalize_quat(qr, qx, qy, qz);
				compute_cov3D(sc[0], sc[1], sc[2], a.mod, qr, qx, qy, qz, cov3D);
			}
			else if (a.cov3D_precomp) { for (int k = 0; k < 6; k++) cov3D[k] = a.cov3D_precomp[6 * idx + k]; }
			else compute_cov3D(sc[0], sc[1], sc[2], a.mod, qr, qx, qy, qz, cov3D);   // the forward computed exactly this; recomputing is bit-identical
			float dmean[3], dcov[6];
			// ---------------- computeCov2DCUDA, backward.cu:177-307 ----------------
			{
				const float* v = a.view;
				float tx = v[0] * mx + v[4] * my + v[8] * mz + v[12];
				float ty = v[1] * mx + v[5] * my + v[9] * mz + v[13];
				const float tz = v[2] * mx + v[6] * my + v[10] * mz + v[14];
				const float limx = 1.3f * a.tan_fovx, limy = 1.3f * a.tan_fovy;
				const float txtz = tx / tz, tytz = ty / tz;
				tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
				ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
				const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
				const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
				const float h_x = a.focal_x, h_y = a.focal_y;
				const float J00 = h_x / tz, J02 = -(h_x * tx) / (tz * tz), J11 = h_y / tz, J12 = -(h_y * ty) / (tz * tz);
				// T = W*J (column-major): T[0][r] = W[0][r]*J00 + W[2][r]*J02 ; T[1][r] = W[1][r]*J11 + W[2][r]*J12 ; W[k][r] = view[4r+k]
				float T0[3], T1[3];
				for (int r = 0; r < 3; r++) { T0[r] = v[4 * r] * J00 + v[4 * r + 2] * J02; T1[r] = v[4 * r + 1] * J11 + v[4 * r + 2] * J12; }
				const float V[3][3] = { { cov3D[0], cov3D[1], cov3D[2] }, { cov3D[1], cov3D[3], cov3D[4] }, { cov3D[2], cov3D[4], cov3D[5] } };
				float TV0[3], TV1[3];      // (T[0] . V[c]), (T[1] . V[c])
				for (int c = 0; c < 3; c++) { TV0[c] = T0[0] * V[c][0] + T0[1] * V[c][1] + T0[2] * V[c][2]; TV1[c] = T1[0] * V[c][0] + T1[1] * V[c][1] + T1[2] * V[c][2]; }
				const float ca = TV0[0] * T0[0] + TV0[1] * T0[1] + TV0[2] * T0[2] + 0.3f;
				const float cb = TV1[0] * T0[0] + TV1[1] * T0[1] + TV1[2] * T0[2];
				const float cc = TV1[0] * T1[0] + TV1[1] * T1[1] + TV1[2] * T1[2] + 0.3f;
				const float denom = ca * cc - cb * cb;
				float dL_da = 0, dL_db = 0, dL_dc = 0;
				const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
				if (denom2inv != 0)
				{
					dL_da = denom2inv * (-cc * cc * dconx + 2 * cb * cc * dcony + (denom - ca * cc) * dconz);
					dL_dc = denom2inv * (-ca * ca * dconz + 2 * ca * cb * dcony + (denom - ca * cc) * dconx);
					dL_db = denom2inv * 2 * (cb * cc * dconx - (denom + 2 * cb * cb) * dcony + ca * cb * dconz);
					dcov[0] = (T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc);
					dcov[3] = (T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc);
					dcov[5] = (T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc);
					dcov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
					dcov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
					dcov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
				}
				else { for (int i = 0; i < 6; i++) dcov[i] = 0; }
				const float dL_dT00 = 2 * TV0[0] * dL_da + TV1[0] * dL_db, dL_dT01 = 2 * TV0[1] * dL_da + TV1[1] * dL_db, dL_dT02 = 2 * TV0[2] * dL_da + TV1[2] * dL_db;
				const float dL_dT10 = 2 * TV1[0] * dL_dc + TV0[0] * dL_db, dL_dT11 = 2 * TV1[1] * dL_dc + TV0[1] * dL_db, dL_dT12 = 2 * TV1[2] * dL_dc + TV0[2] * dL_db;
				// W[c][r] = view[4r + c]
				const float dL_dJ00 = v[0] * dL_dT00 + v[4] * dL_dT01 + v[8] * dL_dT02;
				const float dL_dJ02 = v[2] * dL_dT00 + v[6] * dL_dT01 + v[10] * dL_dT02;
				const float dL_dJ11 = v[1] * dL_dT10 + v[5] * dL_dT11 + v[9] * dL_dT12;
				const float dL_dJ12 = v[2] * dL_dT10 + v[6] * dL_dT11 + v[10] * dL_dT12;
				const float itz = 1.f / tz, tz2 = itz * itz, tz3 = tz2 * itz;
				const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
				const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
				const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * tx) * tz3 * dL_dJ02 + (2 * h_y * ty) * tz3 * dL_dJ12;
				dmean[0] = v[0] * dL_dtx + v[1] * dL_dty + v[2] * dL_dtz;                // transformVec4x3Transpose
				dmean[1] = v[4] * dL_dtx + v[5] * dL_dty + v[6] * dL_dtz;
				dmean[2] = v[8] * dL_dtx + v[9] * dL_dty + v[10] * dL_dtz;
			}
			// ---------------- preprocessCUDA, backward.cu:406-423 ----------------
			{
				const float* p = a.proj;
				const float m_hw = p[3] * mx + p[7] * my + p[11] * mz + p[15];
				const float m_w = 1.0f / (m_hw + 0.0000001f);
				const float mul1 = (p[0] * mx + p[4] * my + p[8] * mz + p[12]) * m_w * m_w;
				const float mul2 = (p[1] * mx + p[5] * my + p[9] * mz + p[13]) * m_w * m_w;
				dmean[0] += (p[0] * m_w - p[3] * mul1) * g2x + (p[1] * m_w - p[3] * mul2) * g2y;
				dmean[1] += (p[4] * m_w - p[7] * mul1) * g2x + (p[5] * m_w - p[7] * mul2) * g2y;
				dmean[2] += (p[8] * m_w - p[11] * mul1) * g2x + (p[9] * m_w - p[11] * mul2) * g2y;
			}
			// ---------------- SH backward, backward.cu:20-172 ----------------
			if (have_sh)
			{
				const int deg = deg_in;
				const uint8_t* idc = QUANT ? a.q.ids_dc + 3 * idx : nullptr;
				const uint8_t* irest = QUANT ? a.q.ids_rest + 45 * idx : nullptr;
				auto sh = [&](int k, int c) -> float {
					if (QUANT) return k == 0 ? s_cb[idc[c]] : s_cb[k * 256 + irest[3 * (k - 1) + c]];
					return myrow[3 * k + c];
				};
				const float dox = mx - a.campos[0], doy = my - a.campos[1], doz = mz - a.campos[2];
				const float len = sqrtf(dox * dox + doy * doy + doz * doz);
				const float x = dox / len, y = doy / len, z = doz / len;
				const unsigned cl = cl_in;
				const float dRGB[3] = { (cl & 1u) ? 0.f : acc0.x, (cl & 2u) ? 0.f : acc0.y, (cl & 4u) ? 0.f : acc0.z };
				// gradient of coefficient k (in place over the staged value; the sparsity term needs the value's sign first)
				auto wr = [&](int k, float w) {
#pragma unroll
					for (int c = 0; c < 3; c++)
					{
						float g = w * dRGB[c];
						if (mult != 0.f && k > 0) { const float sv = sh(k, c); g += mult * (float)((0.f < sv) - (sv < 0.f)); }
						myrow[3 * k + c] = g;
					}
				};
				float dRx[3] = { 0, 0, 0 }, dRy[3] = { 0, 0, 0 }, dRz[3] = { 0, 0, 0 };
				if (deg > 0)
				{
					for (int c = 0; c < 3; c++) { dRx[c] = -kSH_C1 * sh(3, c); dRy[c] = -kSH_C1 * sh(1, c); dRz[c] = kSH_C1 * sh(2, c); }
					wr(1, -kSH_C1 * y); wr(2, kSH_C1 * z); wr(3, -kSH_C1 * x);
					if (deg > 1)
					{
						const float C20 = 1.0925484305920792f, C21 = -1.0925484305920792f, C22 = 0.31539156525252005f, C23 = -1.0925484305920792f, C24 = 0.5462742152960396f;
						const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
						for (int c = 0; c < 3; c++)
						{
							const float s4 = sh(4, c), s5 = sh(5, c), s6 = sh(6, c), s7 = sh(7, c), s8 = sh(8, c);
							dRx[c] += C20 * y * s4 + C22 * 2.f * -x * s6 + C23 * z * s7 + C24 * 2.f * x * s8;
							dRy[c] += C20 * x * s4 + C21 * z * s5 + C22 * 2.f * -y * s6 + C24 * 2.f * -y * s8;
							dRz[c] += C21 * y * s5 + C22 * 2.f * 2.f * z * s6 + C23 * x * s7;
						}
						wr(4, C20 * xy); wr(5, C21 * yz); wr(6, C22 * (2.f * zz - xx - yy)); wr(7, C23 * xz); wr(8, C24 * (xx - yy));
						if (deg > 2)
						{
							const float C30 = -0.5900435899266435f, C31 = 2.890611442640554f, C32 = -0.4570457994644658f, C33 = 0.3731763325901154f,
								C34 = -0.4570457994644658f, C35 = 1.445305721320277f, C36 = -0.5900435899266435f;
							for (int c = 0; c < 3; c++)
							{
								const float s9 = sh(9, c), s10 = sh(10, c), s11 = sh(11, c), s12 = sh(12, c), s13 = sh(13, c), s14 = sh(14, c), s15 = sh(15, c);
								dRx[c] += (C30 * s9 * 3.f * 2.f * xy + C31 * s10 * yz + C32 * s11 * -2.f * xy + C33 * s12 * -3.f * 2.f * xz +
									C34 * s13 * (-3.f * xx + 4.f * zz - yy) + C35 * s14 * 2.f * xz + C36 * s15 * 3.f * (xx - yy));
								dRy[c] += (C30 * s9 * 3.f * (xx - yy) + C31 * s10 * xz + C32 * s11 * (-3.f * yy + 4.f * zz - xx) + C33 * s12 * -3.f * 2.f * yz +
									C34 * s13 * -2.f * xy + C35 * s14 * -2.f * yz + C36 * s15 * -3.f * 2.f * xy);
								dRz[c] += (C31 * s10 * xy + C32 * s11 * 4.f * 2.f * yz + C33 * s12 * 3.f * (2.f * zz - xx - yy) + C34 * s13 * 4.f * 2.f * xz +
									C35 * s14 * (xx - yy));
							}
							wr(9, C30 * y * (3.f * xx - yy)); wr(10, C31 * xy * z); wr(11, C32 * y * (4.f * zz - xx - yy));
							wr(12, C33 * z * (2.f * zz - 3.f * xx - 3.f * yy)); wr(13, C34 * x * (4.f * zz - xx - yy));
							wr(14, C35 * z * (xx - yy)); wr(15, C36 * x * (xx - 3.f * yy));
						}
					}
				}
				wr(0, kSH_C0);
				{ const int nact = (deg + 1) * (deg + 1); for (int k = 3 * nact; k < RL; k++) myrow[k] = 0.f; }
				const float ddx = dRx[0] * dRGB[0] + dRx[1] * dRGB[1] + dRx[2] * dRGB[2];
				const float ddy = dRy[0] * dRGB[0] + dRy[1] * dRGB[1] + dRy[2] * dRGB[2];
				const float ddz = dRz[0] * dRGB[0] + dRz[1] * dRGB[1] + dRz[2] * dRGB[2];
				const float sum2 = dox * dox + doy * doy + doz * doz;                       // dnormvdv, auxiliary.h:107-117
				const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
				dmean[0] += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
				dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
				dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
			}
			// ---------------- cov3D -> scale / rotation, backward.cu:311-374 ----------------
			if (have_scales)
			{
				const float r = qr, x = qx, y = qy, z = qz;
				const float Rm[3][3] = { { 1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y) },
					{ 2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x) },
					{ 2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y) } };     // Rm[c][r]
				const float s[3] = { a.mod * sc[0], a.mod * sc[1], a.mod * sc[2] };
				float Mm[3][3];
				for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++) Mm[c][rr] = s[rr] * Rm[c][rr];
				const float dSig[3][3] = { { dcov[0], 0.5f * dcov[1], 0.5f * dcov[2] }, { 0.5f * dcov[1], dcov[3], 0.5f * dcov[4] }, { 0.5f * dcov[2], 0.5f * dcov[4], dcov[5] } };
				float dMt[3][3];   // dL_dMt[c][r] = dL_dM[r][c], dL_dM = 2 * M * dL_dSigma
				for (int c = 0; c < 3; c++) for (int rr = 0; rr < 3; rr++)
					dMt[rr][c] = 2.f * (Mm[0][rr] * dSig[c][0] + Mm[1][rr] * dSig[c][1] + Mm[2][rr] * dSig[c][2]);
				for (int k = 0; k < 3; k++) o_sc[k] = Rm[0][k] * dMt[k][0] + Rm[1][k] * dMt[k][1] + Rm[2][k] * dMt[k][2];   // Rt[k][j] = Rm[j][k]
				for (int k = 0; k < 3; k++) for (int rr = 0; rr < 3; rr++) dMt[k][rr] *= s[k];
				o_rot[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
				o_rot[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
				o_rot[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
				o_rot[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
			}
			o_m2[0] = g2x; o_m2[1] = g2y;
			o_col[0] = acc0.x; o_col[1] = acc0.y; o_col[2] = acc0.z;
			o_op[0] = acc0.w * (opac * (1.0f - opac));                                    // backward.cu:433
			for (int k = 0; k < 3; k++) o_m3[k] = dmean[k];
			for (int k = 0; k < 6; k++) o_cov[k] = dcov[k];
			o_con[0] = dconx; o_con[1] = dcony; o_con[3] = dconz;
		}
		__syncwarp();
		// ---- unit-stride write-back ------------------------------------------------------------------
		if (a.out.dL_dsh && RL == 48 && n_valid == 32)
		{
			float4* dst4 = reinterpret_cast<float4*>(a.out.dL_dsh + base * 48);
#pragma unroll
			for (int i = 0; i < 12; i++)
			{
				const int f = (i * 32 + lane) * 4, row = f / 48, col = f - row * 48;
				const float* sp = s_row + row * RS + col;
				float4 o = have_sh ? make_float4(sp[0], sp[1], sp[2], sp[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
				if (ACC)
				{
					if (o.x != 0.f || o.y != 0.f || o.z != 0.f || o.w != 0.f) red_add_f32x4(reinterpret_cast<float*>(dst4 + i * 32 + lane), o.x, o.y, o.z, o.w);
				}
				else dst4[i * 32 + lane] = o;
			}
		}
		else if (a.out.dL_dsh)
		{
			float* dst = a.out.dL_dsh + base * RL;
			int row = 0, col = lane;
			while (col >= RL) { col -= RL; row++; }
			for (int f = lane; f < n_valid * RL; f += 32)
			{
				put<ACC>(dst + f, have_sh ? s_row[row * RS + col] : 0.f);
				col += 32;
				while (col >= RL) { col -= RL; row++; }
			}
		}
		warp_store<3, ACC>(a.out.dL_dmeans2D, base, n_valid, o_m2, s_tmp, lane);
		// accumulate mode: THIS view's screen-space gradient on its own (densification statistics are ||.|| per view, gaussian_model.py:693-695)
		if (ACC && a.out.dL_dmeans2D_view) warp_store<3, false>(a.out.dL_dmeans2D_view, base, n_valid, o_m2, s_tmp, lane);
		warp_store<3, ACC>(a.out.dL_dcolors, base, n_valid, o_col, s_tmp, lane);
		warp_store<1, ACC>(a.out.dL_dopacity, base, n_valid, o_op, s_tmp, lane);
		warp_store<3, ACC>(a.out.dL_dmeans3D, base, n_valid, o_m3, s_tmp, lane);
		warp_store<6, ACC>(a.out.dL_dcov3D, base, n_valid, o_cov, s_tmp, lane);
		warp_store<3, ACC>(a.out.dL_dscales, base, n_valid, o_sc, s_tmp, lane);
		warp_store<4, ACC>(a.out.dL_drotations, base, n_valid, o_rot, s_tmp, lane);
		if (a.out.dL_dconic) warp_store<4, ACC>(a.out.dL_dconic, base, n_valid, o_con, s_tmp, lane);
		__syncwarp();
	}
}

int launch_preprocess_backward(const GsbScene* s, const GsbCamera* cam, const GeomState& g, const int32_t* radii, const float* acc,
	const GsbGrads* grads, float lambda, cudaStream_t stream)
{
	BwdArgs a{};
	a.P = s->P; a.M = s->M; a.W = cam->width; a.H = cam->height;
	a.mod = s->scale_modifier; a.tan_fovx = cam->tan_fovx; a.tan_fovy = cam->tan_fovy;
	a.focal_y = cam->height / (2.0f * cam->tan_fovy); a.focal_x = cam->width / (2.0f * cam->tan_fovx);   // rasterizer_impl.cu:573-574
	a.lambda = lambda;
	a.means3D = s->means3D; a.scales = s->scales; a.rotations = s->rotations; a.cov3D_precomp = s->cov3D_precomp;
	a.shs = s->shs; a.colors_precomp = s->colors_precomp; a.degrees = s->degrees; a.radii = radii;
	a.view = cam->viewmatrix; a.proj = cam->projmatrix; a.campos = cam->campos;
	a.quant = s->quant != nullptr; if (s->quant) a.q = *s->quant;
	a.g = g; a.acc = acc; a.out = *grads;
	const int need = (s->P + 255) / 256;
	const int grid = need < 148 * 4 ? need : 148 * 4;
	const size_t smem = (a.quant ? GSB_NUM_CODEBOOKS * GSB_CODEBOOK_SIZE : 0) * sizeof(float) + 8 * (32 * (3 * s->M + 1) + 32 * 6) * sizeof(float);
	ProfScope prof(K_PREPROCESS_BWD, stream);
#define GSB_LAUNCH_PB(Q, A)                                                                                          \
	do {                                                                                                             \
		if (int e = ensure_dyn_smem((const void*)preprocess_backward_kernel<Q, A>, 160 * 1024)) return e;              \
		preprocess_backward_kernel<Q, A><<<grid, 256, smem, stream>>>(a);                                            \
	} while (0)
	if (a.quant) { if (grads->accumulate) GSB_LAUNCH_PB(true, true); else GSB_LAUNCH_PB(true, false); }
	else { if (grads->accumulate) GSB_LAUNCH_PB(false, true); else GSB_LAUNCH_PB(false, false); }
#undef GSB_LAUNCH_PB
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

} // namespace gsb
