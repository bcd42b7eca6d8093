// metrics.cu — fused loss forward+backward, denormalize+PSNR, denormalize+SSIM.
// All are single-pass, 8 (resp. 12) bytes per element, bound by HBM; reductions are two-level
// with a fixed order (deterministic).
#include <algorithm>

#include "common.cuh"

namespace vsr {
namespace {

// ------------------------------------------------------------------------------------------
// losses: torch.nn.L1Loss / MSELoss (main.py:60-63), CharbonnierLoss (losses.py:23-34),
// HuberLoss (losses.py:5-20); mean reduction; grad = dL/dout * grad_scale.
// ------------------------------------------------------------------------------------------
// One launch covers `gridDim.y` equally sized segments (the T frames of a step: acdc_vsr_trainer.py:86 averages
// per-frame means); segment s writes its partial sums to row s of `partials`.  VEC: 16-byte loads / stores with
// four of each in flight per thread (the host checks the alignment of every pointer and of the segment size).
template <int KIND, bool GRAD, bool VEC>
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ out, const float* __restrict__ tgt, long n,
                                                  float param, float gscale, float* __restrict__ partials,
                                                  float* __restrict__ grad) {
  __shared__ float red[32];
  const long base = (long)blockIdx.y * n;
  out += base;
  tgt += base;
  if (GRAD) grad += base;
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
  if (VEC) {
    const long n4 = n >> 2;
    const float4* o4 = reinterpret_cast<const float4*>(out);
    const float4* t4 = reinterpret_cast<const float4*>(tgt);
    float4* g4 = reinterpret_cast<float4*>(grad);
    for (long i0 = (long)blockIdx.x * 1024 + threadIdx.x; i0 < n4; i0 += (long)gridDim.x * 1024) {
      float4 o[4], t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long i = i0 + u * 256;
        if (i < n4) {
          o[u] = __ldcs(o4 + i);
          t[u] = __ldcs(t4 + i);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long i = i0 + u * 256;
        if (i < n4) {
          float4 g;
          one(o[u].x, t[u].x, g.x); one(o[u].y, t[u].y, g.y); one(o[u].z, t[u].z, g.z); one(o[u].w, t[u].w, g.w);
          if (GRAD) g4[i] = g;
        }
      }
    }
  } else {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
      float g;
      one(__ldg(out + i), __ldg(tgt + i), g);
      if (GRAD) grad[i] = g;
    }
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) partials[(size_t)blockIdx.y * kPartialsLen + blockIdx.x] = s;
}

__device__ __forceinline__ float denorm(float v, float mean, float std, bool on) {
  // (x * std + mean).round().clamp(0, 255): two separately rounded ops, round-half-even
  if (!on) return v;
  const float r = rintf(__fadd_rn(__fmul_rn(v, std), mean));
  return fminf(fmaxf(r, 0.f), 255.f);
}

// PSNR pass 1: grid (blocks_per_sample, n): partial sum of squared error; `vec`: 16-byte loads, four of each input
// in flight per thread (the host checks pointer alignment and per_sample % 4)
__global__ void __launch_bounds__(256) psnr_partial_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                                          long per_sample, float mean, float std, int denorm_on, int vec,
                                                          float* __restrict__ ws) {
  __shared__ float red[32];
  const float* o = out + (size_t)blockIdx.y * per_sample;
  const float* t = tgt + (size_t)blockIdx.y * per_sample;
  float acc = 0.f;
  if (vec) {
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* t4 = reinterpret_cast<const float4*>(t);
    const long n4 = per_sample >> 2;
    for (long i0 = (long)blockIdx.x * 1024 + threadIdx.x; i0 < n4; i0 += (long)gridDim.x * 1024) {
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long i = i0 + u * 256;
        if (i < n4) {
          a[u] = __ldg(o4 + i);
          b[u] = __ldg(t4 + i);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u * 256 < n4) {
          const float d0 = denorm(a[u].x, mean, std, denorm_on) - denorm(b[u].x, mean, std, denorm_on);
          const float d1 = denorm(a[u].y, mean, std, denorm_on) - denorm(b[u].y, mean, std, denorm_on);
          const float d2 = denorm(a[u].z, mean, std, denorm_on) - denorm(b[u].z, mean, std, denorm_on);
          const float d3 = denorm(a[u].w, mean, std, denorm_on) - denorm(b[u].w, mean, std, denorm_on);
          acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
        }
      }
    }
  } else {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per_sample; i += (long)gridDim.x * blockDim.x) {
      const float d = denorm(__ldg(o + i), mean, std, denorm_on) - denorm(__ldg(t + i), mean, std, denorm_on);
      acc = fmaf(d, d, acc);
    }
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

// SSIM (valid 11x11 separable window): one warp = a strip of 32 output columns x kSRows output rows.
// Per input row the lane forms the 5 horizontal sums (x, y, x^2, y^2, xy) of its column from a
// per-warp shared-memory row buffer and pushes them into an 11-deep register window; the vertical
// sum over the window gives one SSIM value per row.  The window is indexed statically (the row loop
// is unrolled by 11).  Partials per (image, strip) -> fixed-order final reduce.
constexpr int kWin = 11, kSRows = 32, kSWarps = 4;
__global__ void __launch_bounds__(kSWarps * 32) ssim_partial_kernel(
    const float* __restrict__ out, const float* __restrict__ tgt, int h, int w, const float* __restrict__ win,
    float mean, float std, int denorm_on, float c1, float c2, int tiles_x, int tiles_y, int srows,
    float* __restrict__ ws) {
  __shared__ float rowbuf[kSWarps][2][48];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tx = blockIdx.x * kSWarps + warp;             // column strip
  const int ty = blockIdx.y;                               // row tile
  const int ni = blockIdx.z;
  const int oh = h - (kWin - 1), ow = w - (kWin - 1);
  if (tx >= tiles_x) return;                               // (no block-level barriers below)
  float g[kWin];
#pragma unroll
  for (int k = 0; k < kWin; ++k) g[k] = __ldg(win + k);
  const int x0 = tx * 32, y0 = ty * srows;          // srows output rows per warp (host: enough warps to fill the GPU)
  const int rows_out = min(srows, oh - y0);
  const int rows_in = rows_out + kWin - 1;
  const float* o = out + (size_t)ni * h * w;
  const float* t = tgt + (size_t)ni * h * w;
  float* bx = rowbuf[warp][0];
  float* by = rowbuf[warp][1];
  float wv[5][kWin];
#pragma unroll
  for (int q = 0; q < 5; ++q)
#pragma unroll
    for (int k = 0; k < kWin; ++k) wv[q][k] = 0.f;
  float acc = 0.f;
  const bool col_ok = (x0 + lane) < ow;
  // the row segment [x0, x0 + 42) (denormalised) of input row y0 + r: lanes 0..31 and, for the last 10 columns, lanes 0..9
  const int xa = x0 + lane, xb = x0 + 32 + lane;
  const bool ok_a = xa < w, ok_b = lane < kWin - 1 && xb < w;
  auto fetch = [&](int r, float& a0, float& b0, float& a1, float& b1) {
    a0 = b0 = a1 = b1 = 0.f;
    if (r >= rows_in) return;
    const size_t ro = (size_t)(y0 + r) * w;
    if (ok_a) { a0 = __ldg(o + ro + xa); b0 = __ldg(t + ro + xa); }
    if (ok_b) { a1 = __ldg(o + ro + xb); b1 = __ldg(t + ro + xb); }
  };
  float na0, nb0, na1, nb1;                                // the next row, in flight while this one is filtered
  fetch(0, na0, nb0, na1, nb1);
  for (int base = 0; base < rows_in; base += kWin) {
#pragma unroll
    for (int j = 0; j < kWin; ++j) {
      const int r = base + j;
      if (r < rows_in) {                                   // warp-uniform
        const float a0 = ok_a ? denorm(na0, mean, std, denorm_on) : 0.f, b0 = ok_a ? denorm(nb0, mean, std, denorm_on) : 0.f;
        const float a1 = ok_b ? denorm(na1, mean, std, denorm_on) : 0.f, b1 = ok_b ? denorm(nb1, mean, std, denorm_on) : 0.f;
        fetch(r + 1, na0, nb0, na1, nb1);
        __syncwarp();
        bx[lane] = a0; by[lane] = b0;
        if (lane < kWin - 1) { bx[32 + lane] = a1; by[32 + lane] = b1; }
        __syncwarp();
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
          const float a = bx[lane + k], b = by[lane + k], wk = g[k];
          m1 = fmaf(wk, a, m1);
          m2 = fmaf(wk, b, m2);
          s11 = fmaf(wk, a * a, s11);
          s22 = fmaf(wk, b * b, s22);
          s12 = fmaf(wk, a * b, s12);
        }
        wv[0][j] = m1; wv[1][j] = m2; wv[2][j] = s11; wv[3][j] = s22; wv[4][j] = s12;
        if (r >= kWin - 1) {
          // window rows r-10 .. r live in slots (j+1) % 11 .. j; slot (j + 1 + k) % 11 has tap k
          float v[5];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            float sv = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) sv = fmaf(g[k], wv[q][(j + 1 + k) % kWin], sv);
            v[q] = sv;
          }
          const float var1 = v[2] - v[0] * v[0], var2 = v[3] - v[1] * v[1], cov = v[4] - v[0] * v[1];
          const float val = ((2.f * v[0] * v[1] + c1) * (2.f * cov + c2)) / ((v[0] * v[0] + v[1] * v[1] + c1) * (var1 + var2 + c2));
          acc += col_ok ? val : 0.f;
        }
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) ws[((size_t)ni * tiles_y + ty) * tiles_x + tx] = acc;
}
__global__ void ssim_final_kernel(const float* __restrict__ ws, int n, int tiles, float inv_count,
                                  float* __restrict__ ssim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < tiles; ++b) s += ws[(size_t)i * tiles + b];
  ssim[i] = s * inv_count;
}

// ---- SSIM over volumes (dim = 3): the 11^3 window is separable; three passes over the five moment
// maps (x, y, x^2, y^2, xy): along w (from the images, denormalised on load), along h, along d (fused
// with the SSIM formula and a per-block partial sum).  Maps are stored [5][n][d][h'][w'].
__global__ void __launch_bounds__(256) ssim3_w_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                                     long rows, int w, const float* __restrict__ win, float mean,
                                                     float std, int denorm_on, float* __restrict__ m) {
  const int ow = w - (kWin - 1);
  const long total = rows * ow;
  const long stride5 = total;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / ow;
    const int x = (int)(i - r * ow);
    const float* o = out + r * w + x;
    const float* t = tgt + r * w + x;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float a = denorm(__ldg(o + k), mean, std, denorm_on), b = denorm(__ldg(t + k), mean, std, denorm_on);
      const float wk = __ldg(win + k);
      m1 = fmaf(wk, a, m1); m2 = fmaf(wk, b, m2);
      s11 = fmaf(wk, a * a, s11); s22 = fmaf(wk, b * b, s22); s12 = fmaf(wk, a * b, s12);
    }
    m[i] = m1; m[stride5 + i] = m2; m[2 * stride5 + i] = s11; m[3 * stride5 + i] = s22; m[4 * stride5 + i] = s12;
  }
}
// valid 11-tap sum along the middle axis of [planes][len][inner] -> [planes][len-10][inner]
__global__ void __launch_bounds__(256) ssim3_axis_kernel(const float* __restrict__ src, long planes, int len, long inner,
                                                        const float* __restrict__ win, float* __restrict__ dst) {
  const int ol = len - (kWin - 1);
  const long total = planes * ol * inner;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long x = i % inner;
    const long q = i / inner;
    const int y = (int)(q % ol);
    const long pl = q / ol;
    const float* sp = src + (pl * len + y) * inner + x;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) acc = fmaf(__ldg(win + k), __ldg(sp + k * inner), acc);
    dst[i] = acc;
  }
}
// last axis (d) + SSIM formula; grid (blocks, n); src is [5][n][d][inner]; partial sums per (n, block)
__global__ void __launch_bounds__(256) ssim3_d_kernel(const float* __restrict__ src, int n, int d, long inner,
                                                     const float* __restrict__ win, float c1, float c2,
                                                     float* __restrict__ ws) {
  __shared__ float red[32];
  const int od = d - (kWin - 1);
  const long per = (long)od * inner;
  const long stride5 = (long)n * d * inner;
  const float* base = src + (size_t)blockIdx.y * d * inner;
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    const long x = i % inner;
    const int z = (int)(i / inner);
    float v[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float* sp = base + q * stride5 + (long)z * inner + x;
      float sv = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) sv = fmaf(__ldg(win + k), __ldg(sp + k * inner), sv);
      v[q] = sv;
    }
    const float var1 = v[2] - v[0] * v[0], var2 = v[3] - v[1] * v[1], cov = v[4] - v[0] * v[1];
    acc += ((2.f * v[0] * v[1] + c1) * (2.f * cov + c2)) / ((v[0] * v[0] + v[1] * v[1] + c1) * (var1 + var2 + c2));
  }
  const float sm = block_sum(acc, red);
  if (threadIdx.x == 0) ws[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = sm;
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

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int vsr_loss_fwd_bwd_seg(const float* out, const float* target, int64_t numel, int32_t n_segments, int32_t kind,
                                    float param, float grad_scale, float* loss_partials, float* grad, void* stream) {
  VSR_CHECK_ARG(out && target && loss_partials && numel > 0 && n_segments >= 1 && n_segments <= 65535,
                "vsr_loss_fwd_bwd: bad arguments");
  VSR_CHECK_ARG(kind >= 0 && kind <= 3, "vsr_loss_fwd_bwd: kind must be 0..3");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // 16-byte path only when every segment of every tensor starts on a 16-byte boundary (slices of a stacked buffer
  // with an odd frame size do not: ADVICE r1) - otherwise the scalar loop takes everything
  const bool vec = (numel & 3) == 0 && aligned16(out) && aligned16(target) && (!grad || aligned16(grad));
  long per_block = vec ? 4096 : 1024;                 // elements per block and loop trip
  long gx = (numel + per_block - 1) / per_block;
  const long cap = std::max<long>(1, (long)num_sms() * 8 / n_segments);
  if (gx > cap) gx = cap;
  if (gx > kPartialsLen) gx = kPartialsLen;
  const dim3 grid((unsigned)gx, (unsigned)n_segments);
#define VSR_LOSS2(K, G, V) loss_kernel<K, G, V><<<grid, 256, 0, s>>>(out, target, numel, param, grad_scale, loss_partials, grad)
#define VSR_LOSS(K)                                               \
  if (grad) { if (vec) VSR_LOSS2(K, true, true); else VSR_LOSS2(K, true, false); } \
  else { if (vec) VSR_LOSS2(K, false, true); else VSR_LOSS2(K, false, false); }
  switch (kind) {
    case 0: VSR_LOSS(0) break;
    case 1: VSR_LOSS(1) break;
    case 2: VSR_LOSS(2) break;
    default: VSR_LOSS(3) break;
  }
#undef VSR_LOSS
#undef VSR_LOSS2
  VSR_CHECK_LAUNCH("vsr_loss_fwd_bwd");
  return VSR_OK;
}

extern "C" int vsr_loss_fwd_bwd(const float* out, const float* target, int64_t numel, int32_t kind, float param,
                                float grad_scale, float* loss_partials, float* grad, void* stream) {
  return vsr_loss_fwd_bwd_seg(out, target, numel, 1, kind, param, grad_scale, loss_partials, grad, stream);
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
  const int vec = (per_sample & 3) == 0 && aligned16(out) && aligned16(target);
  psnr_partial_kernel<<<dim3(bps, n), 256, 0, s>>>(out, target, per_sample, mean, std, std > 0.f, vec, ws);
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
  const int tiles_x = (ow + 31) / 32;
  int srows = kSRows;                                   // shorter strips until >= 8 warps per SM exist
  while (srows > 8 && (long)tiles_x * ((oh + srows - 1) / srows) * n < (long)num_sms() * 8) srows >>= 1;
  if ((size_t)tiles_x * ((oh + srows - 1) / srows) > (size_t)((int64_t)h * w_ / 256 + 64)) srows = kSRows;
  const int tiles_y = (oh + srows - 1) / srows;
  VSR_CHECK_SUPPORTED(tiles_y <= 65535 && n <= 65535, "vsr_ssim: image too tall or batch too large");
  VSR_CHECK_SUPPORTED((size_t)tiles_x * tiles_y <= (size_t)((int64_t)h * w_ / 256 + 64), "vsr_ssim: degenerate aspect ratio");
  float* ws = static_cast<float*>(workspace);
  ssim_partial_kernel<<<dim3((tiles_x + kSWarps - 1) / kSWarps, tiles_y, n), kSWarps * 32, 0, s>>>(
      out, target, h, w_, win11, mean, std, std > 0.f, c1, c2, tiles_x, tiles_y, srows, ws);
  VSR_CHECK_LAUNCH("vsr_ssim");
  ssim_final_kernel<<<(n + 127) / 128, 128, 0, s>>>(ws, n, tiles_x * tiles_y, 1.f / ((float)oh * (float)ow), ssim_out);
  VSR_CHECK_LAUNCH("vsr_ssim_final");
  return VSR_OK;
}

extern "C" size_t vsr_ssim3d_workspace(int32_t n, int32_t d, int32_t h, int32_t w_) {
  const size_t m1 = (size_t)5 * n * d * h * (size_t)(w_ > 10 ? w_ - 10 : 0);
  const size_t m2 = (size_t)5 * n * d * (size_t)(h > 10 ? h - 10 : 0) * (size_t)(w_ > 10 ? w_ - 10 : 0);
  return (m1 + m2 + (size_t)n * 1024 + 64) * sizeof(float);
}

extern "C" int vsr_ssim3d(const float* out, const float* target, int32_t n, int32_t d, int32_t h, int32_t w_,
                          const float* win11, float mean, float std, float c1, float c2, float* ssim_out,
                          void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(out && target && ssim_out && win11 && n > 0, "vsr_ssim3d: bad arguments");
  VSR_CHECK_ARG(d >= 11 && h >= 11 && w_ >= 11, "vsr_ssim3d: volume smaller than the 11x11x11 window");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_ssim3d_workspace(n, d, h, w_), "vsr_ssim3d: workspace too small");
  VSR_CHECK_SUPPORTED(n <= 65535, "vsr_ssim3d: batch too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int od = d - 10, oh = h - 10, ow = w_ - 10;
  float* m1 = static_cast<float*>(workspace);
  float* m2 = m1 + (size_t)5 * n * d * h * ow;
  float* part = m2 + (size_t)5 * n * d * oh * ow;
  const long rows = (long)n * d * h;
  ssim3_w_kernel<<<grid_for(rows * ow, 256, 8), 256, 0, s>>>(out, target, rows, w_, win11, mean, std, std > 0.f, m1);
  VSR_CHECK_LAUNCH("vsr_ssim3d(w)");
  ssim3_axis_kernel<<<grid_for((long)5 * n * d * oh * ow, 256, 8), 256, 0, s>>>(m1, (long)5 * n * d, h, ow, win11, m2);
  VSR_CHECK_LAUNCH("vsr_ssim3d(h)");
  const long per = (long)od * oh * ow;
  int bps = (int)((per + 2047) / 2048);
  if (bps > 1024) bps = 1024;
  if (bps < 1) bps = 1;
  ssim3_d_kernel<<<dim3(bps, n), 256, 0, s>>>(m2, n, d, (long)oh * ow, win11, c1, c2, part);
  VSR_CHECK_LAUNCH("vsr_ssim3d(d)");
  ssim_final_kernel<<<(n + 127) / 128, 128, 0, s>>>(part, n, bps, 1.f / (float)per, ssim_out);
  VSR_CHECK_LAUNCH("vsr_ssim3d_final");
  return VSR_OK;
}
