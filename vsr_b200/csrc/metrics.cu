// metrics.cu — fused loss forward+backward, denormalize+PSNR, denormalize+SSIM.
// All are single-pass, 8 (resp. 12) bytes per element, bound by HBM; reductions are two-level
// with a fixed order (deterministic).
#include "common.cuh"

namespace vsr {
namespace {

// ------------------------------------------------------------------------------------------
// losses: torch.nn.L1Loss / MSELoss (main.py:60-63), CharbonnierLoss (losses.py:23-34),
// HuberLoss (losses.py:5-20); mean reduction; grad = dL/dout * grad_scale.
// ------------------------------------------------------------------------------------------
template <int KIND, bool GRAD>
__global__ void loss_kernel(const float4* __restrict__ out, const float4* __restrict__ tgt, long n4,
                            const float* __restrict__ out_s, const float* __restrict__ tgt_s, long n,
                            float param, float gscale, float* __restrict__ partials,
                            float4* __restrict__ grad, float* __restrict__ grad_s) {
  __shared__ float red[32];
  float acc = 0.f;
  auto one = [&](float o, float t, float& g) {
    const float d = o - t;
    if (KIND == 0) {  // L1
      acc += fabsf(d);
      g = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
    } else if (KIND == 1) {  // MSE
      acc += d * d;
      g = 2.f * d * gscale;
    } else if (KIND == 2) {  // Charbonnier: sqrt(d^2 + eps)
      const float s = sqrtf(d * d + param);
      acc += s;
      g = d / s * gscale;
    } else {  // Huber: q = min(|d|, delta); 0.5 q^2 + delta (|d| - q)
      const float ad = fabsf(d);
      const float q = fminf(ad, param);
      acc += 0.5f * q * q + param * (ad - q);
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      g = (ad < param ? d : param * sgn) * gscale;
    }
  };
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 o = __ldg(out + i), t = __ldg(tgt + i);
    float4 g;
    one(o.x, t.x, g.x); one(o.y, t.y, g.y); one(o.z, t.z, g.z); one(o.w, t.w, g.w);
    if (GRAD) grad[i] = g;
  }
  // scalar tail (n % 4)
  for (long i = n4 * 4 + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float g;
    one(out_s[i], tgt_s[i], g);
    if (GRAD) grad_s[i] = g;
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__device__ __forceinline__ float denorm(float v, float mean, float std, bool on) {
  // (x * std + mean).round().clamp(0, 255): two separately rounded ops, round-half-even
  if (!on) return v;
  const float r = rintf(__fadd_rn(__fmul_rn(v, std), mean));
  return fminf(fmaxf(r, 0.f), 255.f);
}

// PSNR pass 1: grid (blocks_per_sample, n): partial sum of squared error
__global__ void psnr_partial_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                    long per_sample, float mean, float std, int denorm_on,
                                    float* __restrict__ ws) {
  __shared__ float red[32];
  const float* o = out + (size_t)blockIdx.y * per_sample;
  const float* t = tgt + (size_t)blockIdx.y * per_sample;
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per_sample; i += (long)gridDim.x * blockDim.x) {
    const float d = denorm(__ldg(o + i), mean, std, denorm_on) - denorm(__ldg(t + i), mean, std, denorm_on);
    acc = fmaf(d, d, acc);
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) ws[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
}
__global__ void psnr_final_kernel(const float* __restrict__ ws, int n, int bps, long per_sample,
                                  float max_value, float* __restrict__ psnr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < bps; ++b) s += ws[(size_t)i * bps + b];
  const float mse = s / (float)per_sample;
  psnr[i] = 10.f * log10f(max_value * max_value / (mse + 1e-10f));
}

// SSIM: block = 16 x 32 output pixels of one image (valid 11x11 separable window).
constexpr int kSH = 16, kSW = 32, kWin = 11;
__global__ void __launch_bounds__(kSH * kSW) ssim_partial_kernel(
    const float* __restrict__ out, const float* __restrict__ tgt, int h, int w,
    const float* __restrict__ win, float mean, float std, int denorm_on, float c1, float c2,
    int tiles_x, int tiles_y, float* __restrict__ ws) {
  __shared__ float sx[kSH + kWin - 1][kSW + kWin - 1];
  __shared__ float sy[kSH + kWin - 1][kSW + kWin - 1];
  __shared__ float hs[5][kSH + kWin - 1][kSW];
  __shared__ float g[kWin];
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int ni = blockIdx.y;
  const int oh = h - (kWin - 1), ow = w - (kWin - 1);
  const int y0 = ty * kSH, x0 = tx * kSW;
  const float* o = out + (size_t)ni * h * w;
  const float* t = tgt + (size_t)ni * h * w;
  if (tid < kWin) g[tid] = win[tid];
  for (int i = tid; i < (kSH + kWin - 1) * (kSW + kWin - 1); i += blockDim.x) {
    const int r = i / (kSW + kWin - 1), c = i % (kSW + kWin - 1);
    const int yy = y0 + r, xx = x0 + c;
    float a = 0.f, b = 0.f;
    if (yy < h && xx < w) {
      a = denorm(__ldg(o + (size_t)yy * w + xx), mean, std, denorm_on);
      b = denorm(__ldg(t + (size_t)yy * w + xx), mean, std, denorm_on);
    }
    sx[r][c] = a;
    sy[r][c] = b;
  }
  __syncthreads();
  for (int i = tid; i < (kSH + kWin - 1) * kSW; i += blockDim.x) {
    const int r = i / kSW, c = i % kSW;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float a = sx[r][c + k], b = sy[r][c + k], wk = g[k];
      m1 = fmaf(wk, a, m1);
      m2 = fmaf(wk, b, m2);
      s11 = fmaf(wk, a * a, s11);
      s22 = fmaf(wk, b * b, s22);
      s12 = fmaf(wk, a * b, s12);
    }
    hs[0][r][c] = m1; hs[1][r][c] = m2; hs[2][r][c] = s11; hs[3][r][c] = s22; hs[4][r][c] = s12;
  }
  __syncthreads();
  const int r = tid / kSW, c = tid % kSW;
  float val = 0.f;
  if (y0 + r < oh && x0 + c < ow) {
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float wk = g[k];
      m1 = fmaf(wk, hs[0][r + k][c], m1);
      m2 = fmaf(wk, hs[1][r + k][c], m2);
      s11 = fmaf(wk, hs[2][r + k][c], s11);
      s22 = fmaf(wk, hs[3][r + k][c], s22);
      s12 = fmaf(wk, hs[4][r + k][c], s12);
    }
    const float v1 = s11 - m1 * m1, v2 = s22 - m2 * m2, cov = s12 - m1 * m2;
    val = ((2.f * m1 * m2 + c1) * (2.f * cov + c2)) / ((m1 * m1 + m2 * m2 + c1) * (v1 + v2 + c2));
  }
  const float s = block_sum(val, red);
  if (tid == 0) ws[(size_t)ni * (tiles_x * tiles_y) + tile] = s;
}
__global__ void ssim_final_kernel(const float* __restrict__ ws, int n, int tiles, float inv_count,
                                  float* __restrict__ ssim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < tiles; ++b) s += ws[(size_t)i * tiles + b];
  ssim[i] = s * inv_count;
}

int psnr_bps(long per_sample, int n) {
  long b = (per_sample + 4095) / 4096;
  long want = ((long)num_sms() * 4 + n - 1) / n;
  if (b > want) b = want;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_loss_fwd_bwd(const float* out, const float* target, int64_t numel, int32_t kind, float param,
                                float grad_scale, float* loss_partials, float* grad, void* stream) {
  VSR_CHECK_ARG(out && target && loss_partials && numel > 0, "vsr_loss_fwd_bwd: bad arguments");
  VSR_CHECK_ARG(kind >= 0 && kind <= 3, "vsr_loss_fwd_bwd: kind must be 0..3");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long n4 = numel / 4;
  int grid = grid_for(n4 > 0 ? n4 : 1, 256, 4);
  if (grid > kPartialsLen) grid = kPartialsLen;
  const float4* o4 = reinterpret_cast<const float4*>(out);
  const float4* t4 = reinterpret_cast<const float4*>(target);
  float4* g4 = reinterpret_cast<float4*>(grad);
#define VSR_LOSS(K)                                                                                          \
  if (grad) loss_kernel<K, true><<<grid, 256, 0, s>>>(o4, t4, n4, out, target, numel, param, grad_scale, loss_partials, g4, grad); \
  else loss_kernel<K, false><<<grid, 256, 0, s>>>(o4, t4, n4, out, target, numel, param, grad_scale, loss_partials, nullptr, nullptr);
  switch (kind) {
    case 0: VSR_LOSS(0) break;
    case 1: VSR_LOSS(1) break;
    case 2: VSR_LOSS(2) break;
    default: VSR_LOSS(3) break;
  }
#undef VSR_LOSS
  VSR_CHECK_LAUNCH("vsr_loss_fwd_bwd");
  return VSR_OK;
}

extern "C" size_t vsr_metric_workspace(int32_t n, int64_t per_sample) {
  return (size_t)n * (size_t)(per_sample / 256 + 64) * sizeof(float);
}

extern "C" int vsr_psnr(const float* out, const float* target, int32_t n, int64_t per_sample, float mean,
                        float std, float max_value, float* psnr_out, void* workspace, size_t workspace_bytes,
                        void* stream) {
  VSR_CHECK_ARG(out && target && psnr_out && n > 0 && per_sample > 0, "vsr_psnr: bad arguments");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_metric_workspace(n, per_sample), "vsr_psnr: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int bps = psnr_bps(per_sample, n);
  float* ws = static_cast<float*>(workspace);
  psnr_partial_kernel<<<dim3(bps, n), 256, 0, s>>>(out, target, per_sample, mean, std, std > 0.f, ws);
  VSR_CHECK_LAUNCH("vsr_psnr");
  psnr_final_kernel<<<(n + 127) / 128, 128, 0, s>>>(ws, n, bps, per_sample, max_value, psnr_out);
  VSR_CHECK_LAUNCH("vsr_psnr_final");
  return VSR_OK;
}

extern "C" int vsr_ssim(const float* out, const float* target, int32_t n, int32_t h, int32_t w_,
                        const float* win11, float mean, float std, float c1, float c2, float* ssim_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(out && target && ssim_out && win11 && n > 0, "vsr_ssim: bad arguments");
  VSR_CHECK_ARG(h >= 11 && w_ >= 11, "vsr_ssim: image smaller than the 11x11 window");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_metric_workspace(n, (int64_t)h * w_), "vsr_ssim: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int oh = h - 10, ow = w_ - 10;
  const int tiles_x = (ow + kSW - 1) / kSW, tiles_y = (oh + kSH - 1) / kSH;
  VSR_CHECK_SUPPORTED((size_t)tiles_x * tiles_y <= (size_t)((int64_t)h * w_ / 256 + 64), "vsr_ssim: degenerate aspect ratio");
  float* ws = static_cast<float*>(workspace);
  ssim_partial_kernel<<<dim3(tiles_x * tiles_y, n), kSH * kSW, 0, s>>>(out, target, h, w_, win11, mean, std,
                                                                     std > 0.f, c1, c2, tiles_x, tiles_y, ws);
  VSR_CHECK_LAUNCH("vsr_ssim");
  ssim_final_kernel<<<(n + 127) / 128, 128, 0, s>>>(ws, n, tiles_x * tiles_y, 1.f / ((float)oh * (float)ow), ssim_out);
  VSR_CHECK_LAUNCH("vsr_ssim_final");
  return VSR_OK;
}
