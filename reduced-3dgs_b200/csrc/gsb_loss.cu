// gsb_loss.cu — the loss side of the training step: L1 + D-SSIM of the rendered image against the
// ground truth, forward and backward (reference utils/loss_utils.py:17-65 l1_loss / ssim / _ssim, combined in train.py:110-115:
// loss = (1 - lambda) * L1 + lambda * (1 - SSIM)).
//
// The reference evaluates SSIM with five grouped 11x11 conv2d calls, ~15 element-wise ops and autograd's transposed
// convolutions for the backward: ~20 passes over 25 MB images.  Here:
//   forward  kernel: one CTA per 16x16 tile and channel stages the (16+10)^2 neighbourhood of both images in shared memory,
//            runs the SEPARABLE 11-tap Gaussian (sigma 1.5, zero padding like conv2d(padding=5)) over x, y, x^2, y^2, xy,
//            evaluates the SSIM map, accumulates sum(SSIM) and sum|x - y| per CTA, and writes the three partial derivatives
//            d ssim / d (mu_x, E[x^2], E[xy]) that the backward needs;
//   backward kernel: the same separable filter over those three maps (the window is symmetric, so the adjoint of the
//            convolution is the convolution) and dL/dx = s * (F*g_mu + 2 x F*g_xx + y F*g_xy) + L1 term.
// Two launches, ~0.2 GB of traffic at 1080p.
#include "gsb_common.cuh"

namespace gsb {

#define LS_TILE 16
#define LS_HALO 5
#define LS_EXT (LS_TILE + 2 * LS_HALO)          // 26

// gaussian(11, 1.5) of loss_utils.py:23-25, normalised, as float32
__device__ __constant__ float kWin[11] = {
	0.00102838012f, 0.00759875821f, 0.0360007733f, 0.109360687f, 0.213005528f, 0.266011715f,
	0.213005528f, 0.109360687f, 0.0360007733f, 0.00759875821f, 0.00102838012f };

__global__ void __launch_bounds__(256) l1_ssim_forward_kernel(const float* __restrict__ img, const float* __restrict__ gt, int H, int W,
	float* __restrict__ g_mu, float* __restrict__ g_xx, float* __restrict__ g_xy, float* __restrict__ partial /* [blocks][2] */)
{
	__shared__ float sx[LS_EXT][LS_EXT + 1], sy[LS_EXT][LS_EXT + 1];
	__shared__ float sh[5][LS_EXT][LS_TILE + 1];
	__shared__ float s_red[2][8];
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	const int c = blockIdx.z;
	const int x0 = blockIdx.x * LS_TILE - LS_HALO, y0 = blockIdx.y * LS_TILE - LS_HALO;
	const size_t plane = (size_t)c * H * W;
	for (int i = tid; i < LS_EXT * LS_EXT; i += 256)
	{
		const int r = i / LS_EXT, q = i - r * LS_EXT;
		const int gy = y0 + r, gx = x0 + q;
		const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
		sx[r][q] = in ? img[plane + (size_t)gy * W + gx] : 0.0f;
		sy[r][q] = in ? gt[plane + (size_t)gy * W + gx] : 0.0f;
	}
	__syncthreads();
	for (int i = tid; i < LS_EXT * LS_TILE; i += 256)
	{
		const int r = i / LS_TILE, q = i - r * LS_TILE;
		float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
		for (int k = 0; k < 11; k++)
		{
			const float w = kWin[k], x = sx[r][q + k], y = sy[r][q + k];
			a += w * x; b += w * y; aa += w * x * x; bb += w * y * y; ab += w * x * y;
		}
		sh[0][r][q] = a; sh[1][r][q] = b; sh[2][r][q] = aa; sh[3][r][q] = bb; sh[4][r][q] = ab;
	}
	__syncthreads();
	float mu1 = 0.f, mu2 = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
	for (int k = 0; k < 11; k++)
	{
		const float w = kWin[k];
		mu1 += w * sh[0][ty + k][tx]; mu2 += w * sh[1][ty + k][tx];
		exx += w * sh[2][ty + k][tx]; eyy += w * sh[3][ty + k][tx]; exy += w * sh[4][ty + k][tx];
	}
	const int px = blockIdx.x * LS_TILE + tx, py = blockIdx.y * LS_TILE + ty;
	const bool inside = px < W && py < H;
	float ssim = 0.f, l1 = 0.f;
	if (inside)
	{
		const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
		const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
		const float s1 = exx - mu1_sq, s2 = eyy - mu2_sq, s12 = exy - mu12;
		const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
		const float inv = 1.0f / (B1 * B2);
		ssim = A1 * A2 * inv;
		// partial derivatives of the SSIM map w.r.t. the three filtered quantities that depend on x
		//   d/dmu1: dA1 = 2 mu2, dA2 = -2 mu2, dB1 = 2 mu1, dB2 = -2 mu1;  d/dExx: dB2 = 1;  d/dExy: dA2 = 2
		const float d_mu = (2.f * mu2 * (A2 - A1) * inv) - ssim * (2.f * mu1 * (B2 - B1) * inv);
		const float d_xx = -ssim / B2;
		const float d_xy = 2.f * A1 * inv;
		const size_t o = plane + (size_t)py * W + px;
		g_mu[o] = d_mu; g_xx[o] = d_xx; g_xy[o] = d_xy;
		l1 = fabsf(sx[ty + LS_HALO][tx + LS_HALO] - sy[ty + LS_HALO][tx + LS_HALO]);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { ssim += __shfl_xor_sync(0xffffffffu, ssim, o); l1 += __shfl_xor_sync(0xffffffffu, l1, o); }
	if ((tid & 31) == 0) { s_red[0][tid >> 5] = ssim; s_red[1][tid >> 5] = l1; }
	__syncthreads();
	if (tid == 0)
	{
		float a = 0.f, b = 0.f;
		for (int w = 0; w < 8; w++) { a += s_red[0][w]; b += s_red[1][w]; }
		const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
		partial[2 * blk] = a; partial[2 * blk + 1] = b;
	}
}

// dL/dimg = w_ssim * (F*g_mu + 2 x F*g_xx + y F*g_xy) + w_l1 * sign(x - y),  w_* = coef_* / N * (*up_*)  (N = C H W).
// up_l1 / up_ssim are DEVICE scalars (the upstream gradients of mean|x-y| and of mean SSIM; NULL = 1), so no host
// synchronisation is needed to read them; for loss = (1-l) L1 + l (1 - SSIM): coef_l1 = 1 - l, coef_ssim = -l, both ups = dL/dloss.
__global__ void __launch_bounds__(256) l1_ssim_backward_kernel(const float* __restrict__ img, const float* __restrict__ gt, int H, int W,
	const float* __restrict__ g_mu, const float* __restrict__ g_xx, const float* __restrict__ g_xy, float ssim_scale, float l1_scale,
	const float* __restrict__ up_l1, const float* __restrict__ up_ssim, float* __restrict__ dL)
{
	__shared__ float sg[3][LS_EXT][LS_EXT + 1];
	__shared__ float sh[3][LS_EXT][LS_TILE + 1];
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	const int c = blockIdx.z;
	const int x0 = blockIdx.x * LS_TILE - LS_HALO, y0 = blockIdx.y * LS_TILE - LS_HALO;
	const size_t plane = (size_t)c * H * W;
	for (int i = tid; i < LS_EXT * LS_EXT; i += 256)
	{
		const int r = i / LS_EXT, q = i - r * LS_EXT;
		const int gy = y0 + r, gx = x0 + q;
		const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
		const size_t o = plane + (size_t)gy * W + gx;
		sg[0][r][q] = in ? g_mu[o] : 0.0f; sg[1][r][q] = in ? g_xx[o] : 0.0f; sg[2][r][q] = in ? g_xy[o] : 0.0f;
	}
	__syncthreads();
	for (int i = tid; i < LS_EXT * LS_TILE; i += 256)
	{
		const int r = i / LS_TILE, q = i - r * LS_TILE;
		float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
		for (int k = 0; k < 11; k++) { const float w = kWin[k]; a += w * sg[0][r][q + k]; b += w * sg[1][r][q + k]; d += w * sg[2][r][q + k]; }
		sh[0][r][q] = a; sh[1][r][q] = b; sh[2][r][q] = d;
	}
	__syncthreads();
	const int px = blockIdx.x * LS_TILE + tx, py = blockIdx.y * LS_TILE + ty;
	if (px >= W || py >= H) return;
	float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
	for (int k = 0; k < 11; k++) { const float w = kWin[k]; a += w * sh[0][ty + k][tx]; b += w * sh[1][ty + k][tx]; d += w * sh[2][ty + k][tx]; }
	const size_t o = plane + (size_t)py * W + px;
	const float x = img[o], y = gt[o];
	const float w_l1 = l1_scale * (up_l1 ? up_l1[0] : 1.0f), w_ssim = ssim_scale * (up_ssim ? up_ssim[0] : 1.0f);
	const float sgn = x > y ? 1.0f : (x < y ? -1.0f : 0.0f);
	dL[o] = w_ssim * (a + 2.f * x * b + y * d) + w_l1 * sgn;
}

int launch_l1_ssim_forward(const float* img, const float* gt, int C, int H, int W, float* maps, float* partial, cudaStream_t stream)
{
	const dim3 grid((W + LS_TILE - 1) / LS_TILE, (H + LS_TILE - 1) / LS_TILE, C);
	const size_t N = (size_t)C * H * W;
	ProfScope prof(K_TOOLS, stream);
	l1_ssim_forward_kernel<<<grid, 256, 0, stream>>>(img, gt, H, W, maps, maps + N, maps + 2 * N, partial);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

int launch_l1_ssim_backward(const float* img, const float* gt, int C, int H, int W, const float* maps, float coef_l1, const float* up_l1,
	float coef_ssim, const float* up_ssim, float* dL, cudaStream_t stream)
{
	const dim3 grid((W + LS_TILE - 1) / LS_TILE, (H + LS_TILE - 1) / LS_TILE, C);
	const size_t N = (size_t)C * H * W;
	ProfScope prof(K_TOOLS, stream);
	l1_ssim_backward_kernel<<<grid, 256, 0, stream>>>(img, gt, H, W, maps, maps + N, maps + 2 * N, coef_ssim / (float)N, coef_l1 / (float)N,
		up_l1, up_ssim, dL);
	GSB_LAUNCHED();
	GSB_CUDA_OK(cudaGetLastError());
	return GSB_OK;
}

} // namespace gsb
