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
	const float* shs; const int32_t* degrees; const int32_t* radii;
	const float* view; const float* proj; const float* campos;
	int quant; GsbQuant q;
	GeomState g; const float* acc;
	GsbGrads out;
};

template <bool ACC> __device__ __forceinline__ void put(float* p, float v) { if (ACC) *p += v; else *p = v; }

template <bool QUANT, bool ACC>
__global__ void __launch_bounds__(256) preprocess_backward_kernel(const BwdArgs a)
{
	extern __shared__ float s_cb[];
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
	const bool have_sh = QUANT || a.shs != nullptr;
	const bool have_scales = QUANT || a.scales != nullptr;
	for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.P; idx += (long long)gridDim.x * blockDim.x)
	{
		float* o_m2 = a.out.dL_dmeans2D + 3 * idx;
		float* o_col = a.out.dL_dcolors + 3 * idx;
		float* o_m3 = a.out.dL_dmeans3D + 3 * idx;
		float* o_cov = a.out.dL_dcov3D + 6 * idx;
		float* o_sh = a.out.dL_dsh ? a.out.dL_dsh + 3 * idx * a.M : nullptr;
		if (!(a.radii[idx] > 0))
		{
			if (!ACC)
			{
				for (int k = 0; k < 3; k++) { o_m2[k] = 0.f; o_col[k] = 0.f; o_m3[k] = 0.f; a.out.dL_dscales[3 * idx + k] = 0.f; }
				for (int k = 0; k < 6; k++) o_cov[k] = 0.f;
				for (int k = 0; k < 4; k++) a.out.dL_drotations[4 * idx + k] = 0.f;
				a.out.dL_dopacity[idx] = 0.f;
				if (o_sh) for (int k = 0; k < 3 * a.M; k++) o_sh[k] = 0.f;
				if (a.out.dL_dconic) reinterpret_cast<float4*>(a.out.dL_dconic)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
			}
			continue;
		}
		const float4 acc0 = reinterpret_cast<const float4*>(a.acc)[3 * idx];
		const float4 acc1 = reinterpret_cast<const float4*>(a.acc)[3 * idx + 1];
		const float cyy = a.acc[12 * idx + 8];
		// constant factors of backward.cu:498-499, 583-589 applied once per Gaussian
		const float g2x = acc1.x * (0.5f * a.W), g2y = acc1.y * (0.5f * a.H);
		const float dconx = -0.5f * acc1.z, dcony = -0.5f * acc1.w, dconz = -0.5f * cyy;
		const float mx = a.means3D[3 * idx], my = a.means3D[3 * idx + 1], mz = a.means3D[3 * idx + 2];
		// attributes
		float sc[3] = { 0, 0, 0 }, qr = 1, qx = 0, qy = 0, qz = 0;
		float cov3D[6];
		if (QUANT)
		{
			const uint8_t* is = a.q.ids_scaling + 3 * idx; const uint8_t* ir = a.q.ids_rot + 4 * idx;
			for (int k = 0; k < 3; k++) sc[k] = s_cb[17 * 256 + is[k]];
			qr = s_cb[18 * 256 + ir[0]]; qx = s_cb[19 * 256 + ir[1]]; qy = s_cb[19 * 256 + ir[2]]; qz = s_cb[19 * 256 + ir[3]];
			float n2 = __fmul_rn(qr, qr); n2 = __fmaf_rn(qx, qx, n2); n2 = __fmaf_rn(qy, qy, n2); n2 = __fmaf_rn(qz, qz, n2);
			const float n = fmaxf(__fsqrt_rn(n2), 1e-12f);
			qr = __fdiv_rn(qr, n); qx = __fdiv_rn(qx, n); qy = __fdiv_rn(qy, n); qz = __fdiv_rn(qz, n);
			compute_cov3D(sc[0], sc[1], sc[2], a.mod, qr, qx, qy, qz, cov3D);
		}
		else if (a.cov3D_precomp) { for (int k = 0; k < 6; k++) cov3D[k] = a.cov3D_precomp[6 * idx + k]; }
		else
		{
			for (int k = 0; k < 3; k++) sc[k] = a.scales[3 * idx + k];
			const float4 q = reinterpret_cast<const float4*>(a.rotations)[idx];
			qr = q.x; qx = q.y; qy = q.z; qz = q.w;
			compute_cov3D(sc[0], sc[1], sc[2], a.mod, qr, qx, qy, qz, cov3D);       // forward stored this; recomputing is bit-identical
		}
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
		if (have_sh && o_sh)
		{
			const int deg = a.degrees[idx];
			const uint8_t* idc = QUANT ? a.q.ids_dc + 3 * idx : nullptr;
			const uint8_t* irest = QUANT ? a.q.ids_rest + 45 * idx : nullptr;
			const float* shp = QUANT ? nullptr : a.shs + 3 * idx * a.M;
			auto sh = [&](int k, int c) -> float {
				if (QUANT) return k == 0 ? s_cb[idc[c]] : s_cb[k * 256 + irest[3 * (k - 1) + c]];
				return shp[3 * k + c];
			};
			const float dox = mx - a.campos[0], doy = my - a.campos[1], doz = mz - a.campos[2];
			const float len = sqrtf(dox * dox + doy * doy + doz * doz);
			const float x = dox / len, y = doy / len, z = doz / len;
			const unsigned cl = a.g.clamped[idx];
			float dRGB[3] = { (cl & 1u) ? 0.f : acc0.x, (cl & 2u) ? 0.f : acc0.y, (cl & 4u) ? 0.f : acc0.z };
			auto wr = [&](int k, float w) {
				for (int c = 0; c < 3; c++)
				{
					float g = w * dRGB[c];
					if (mult != 0.f && k > 0) { const float s = sh(k, c); g += mult * (float)((0.f < s) - (s < 0.f)); }
					put<ACC>(o_sh + 3 * k + c, g);
				}
			};
			float dRx[3] = { 0, 0, 0 }, dRy[3] = { 0, 0, 0 }, dRz[3] = { 0, 0, 0 };
			wr(0, kSH_C0);
			if (deg > 0)
			{
				wr(1, -kSH_C1 * y); wr(2, kSH_C1 * z); wr(3, -kSH_C1 * x);
				for (int c = 0; c < 3; c++) { dRx[c] = -kSH_C1 * sh(3, c); dRy[c] = -kSH_C1 * sh(1, c); dRz[c] = kSH_C1 * sh(2, c); }
				if (deg > 1)
				{
					const float C20 = 1.0925484305920792f, C21 = -1.0925484305920792f, C22 = 0.31539156525252005f, C23 = -1.0925484305920792f, C24 = 0.5462742152960396f;
					const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
					wr(4, C20 * xy); wr(5, C21 * yz); wr(6, C22 * (2.f * zz - xx - yy)); wr(7, C23 * xz); wr(8, C24 * (xx - yy));
					for (int c = 0; c < 3; c++)
					{
						const float s4 = sh(4, c), s5 = sh(5, c), s6 = sh(6, c), s7 = sh(7, c), s8 = sh(8, c);
						dRx[c] += C20 * y * s4 + C22 * 2.f * -x * s6 + C23 * z * s7 + C24 * 2.f * x * s8;
						dRy[c] += C20 * x * s4 + C21 * z * s5 + C22 * 2.f * -y * s6 + C24 * 2.f * -y * s8;
						dRz[c] += C21 * y * s5 + C22 * 2.f * 2.f * z * s6 + C23 * x * s7;
					}
					if (deg > 2)
					{
						const float C30 = -0.5900435899266435f, C31 = 2.890611442640554f, C32 = -0.4570457994644658f, C33 = 0.3731763325901154f,
							C34 = -0.4570457994644658f, C35 = 1.445305721320277f, C36 = -0.5900435899266435f;
						wr(9, C30 * y * (3.f * xx - yy)); wr(10, C31 * xy * z); wr(11, C32 * y * (4.f * zz - xx - yy));
						wr(12, C33 * z * (2.f * zz - 3.f * xx - 3.f * yy)); wr(13, C34 * x * (4.f * zz - xx - yy));
						wr(14, C35 * z * (xx - yy)); wr(15, C36 * x * (xx - 3.f * yy));
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
					}
				}
			}
			if (!ACC) { const int nact = (deg + 1) * (deg + 1); for (int k = 3 * nact; k < 3 * a.M; k++) o_sh[k] = 0.f; }
			const float ddx = dRx[0] * dRGB[0] + dRx[1] * dRGB[1] + dRx[2] * dRGB[2];
			const float ddy = dRy[0] * dRGB[0] + dRy[1] * dRGB[1] + dRy[2] * dRGB[2];
			const float ddz = dRz[0] * dRGB[0] + dRz[1] * dRGB[1] + dRz[2] * dRGB[2];
			const float sum2 = dox * dox + doy * doy + doz * doz;                       // dnormvdv, auxiliary.h:107-117
			const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
			dmean[0] += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
			dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
			dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
		}
		else if (o_sh && !ACC) { for (int k = 0; k < 3 * a.M; k++) o_sh[k] = 0.f; }
		// ---------------- cov3D -> scale / rotation, backward.cu:311-374 ----------------
		float dscale[3] = { 0, 0, 0 }, dq[4] = { 0, 0, 0, 0 };
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
			// Rt[k][j] = Rm[j][k]
			for (int k = 0; k < 3; k++) dscale[k] = Rm[0][k] * dMt[k][0] + Rm[1][k] * dMt[k][1] + Rm[2][k] * dMt[k][2];
			for (int k = 0; k < 3; k++) for (int rr = 0; rr < 3; rr++) dMt[k][rr] *= s[k];
			dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
			dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
			dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
			dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
		}
		// ---------------- stores ----------------
		const float opac = a.g.rec[3 * idx].w;
		put<ACC>(o_m2 + 0, g2x); put<ACC>(o_m2 + 1, g2y); if (!ACC) o_m2[2] = 0.f;
		put<ACC>(o_col + 0, acc0.x); put<ACC>(o_col + 1, acc0.y); put<ACC>(o_col + 2, acc0.z);
		put<ACC>(a.out.dL_dopacity + idx, acc0.w * (opac * (1.0f - opac)));                 // backward.cu:433
		for (int k = 0; k < 3; k++) put<ACC>(o_m3 + k, dmean[k]);
		for (int k = 0; k < 6; k++) put<ACC>(o_cov + k, dcov[k]);
		for (int k = 0; k < 3; k++) put<ACC>(a.out.dL_dscales + 3 * idx + k, dscale[k]);
		for (int k = 0; k < 4; k++) put<ACC>(a.out.dL_drotations + 4 * idx + k, dq[k]);
		if (a.out.dL_dconic) { float* c = a.out.dL_dconic + 4 * idx; put<ACC>(c, dconx); put<ACC>(c + 1, dcony); if (!ACC) c[2] = 0.f; put<ACC>(c + 3, dconz); }
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
	a.shs = s->shs; a.degrees = s->degrees; a.radii = radii;
	a.view = cam->viewmatrix; a.proj = cam->projmatrix; a.campos = cam->campos;
	a.quant = s->quant != nullptr; if (s->quant) a.q = *s->quant;
	a.g = g; a.acc = acc; a.out = *grads;
	const int need = (s->P + 255) / 256;
	const int grid = need < 148 * 8 ? need : 148 * 8;
	const int smem = a.quant ? GSB_NUM_CODEBOOKS * GSB_CODEBOOK_SIZE * (int)sizeof(float) : 0;
	ProfScope prof(K_PREPROCESS_BWD, stream);
	if (a.quant) { if (grads->accumulate) preprocess_backward_kernel<true, true><<<grid, 256, smem, stream>>>(a); else preprocess_backward_kernel<true, false><<<grid, 256, smem, stream>>>(a); }
	else { if (grads->accumulate) preprocess_backward_kernel<false, true><<<grid, 256, 0, stream>>>(a); else preprocess_backward_kernel<false, false><<<grid, 256, 0, stream>>>(a); }
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

} // namespace gsb
