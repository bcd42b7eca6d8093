// dense3d.cu — the bandwidth-bound kernels of the Conv3d network (DUFNet, reference duf_net.py:21-214):
// BatchNorm3d batch statistics / normalise + ReLU / their backward on channel windows of the dense-concat
// map, channel-window copies, and the dynamic-upsampling-filter tail (softmax over the 5x5 taps, local
// filtering of the centre frame, pixel shuffle, residual add) forward and backward.
//
// All maps are pixel-major [rows][ld] with the channel window [c0, c0+c) addressed explicitly, so the dense
// concatenation (duf_net.py:123-128) is one buffer that every layer reads a prefix of and appends a slice to.
// Every thread moves 16-byte vectors; a thread keeps the same channel vector for all of its rows, so the
// per-channel constants live in registers.  Reductions are two-pass with a fixed summation order.
#include "common.cuh"

namespace vsr {
namespace {

constexpr int kThreads = 256;

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int V = 4;
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int V = 8;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* v) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = q;
  }
};

// rows of a [rows][.] map split over blocks: block b of `nb` gets [lo, hi)
__device__ __forceinline__ void row_range(long rows, int b, int nb, long* lo, long* hi) {
  const long per = (rows + nb - 1) / nb;
  *lo = (long)b * per;
  *hi = *lo + per < rows ? *lo + per : rows;
}

// ---- channel-window copy ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) copy_window_kernel(const T* __restrict__ src, int lds, int s0,
                                                               T* __restrict__ dst, int ldd, int d0, int c, long rows) {
  constexpr int V = Vec<T>::V;
  const int tpr = c / V;
  const long total = rows * tpr;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long)gridDim.x * kThreads) {
    const long row = i / tpr;
    const int v = (int)(i - row * tpr) * V;
    *reinterpret_cast<uint4*>(dst + row * ldd + d0 + v) = *reinterpret_cast<const uint4*>(src + row * lds + s0 + v);
  }
}

// ---- batch statistics ------------------------------------------------------------------------------------
// grid = frames * bpf blocks; block (f, b) reduces its rows of frame f to [2][c] doubles in ws.
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const T* __restrict__ x, int ldx, int c0, int c,
                                                            long rows_per_frame, int bpf, double* __restrict__ ws) {
  constexpr int V = Vec<T>::V;
  extern __shared__ double sm[];   // [rpp][2][c]
  const int tpr = c / V, rpp = kThreads / tpr;
  const int f = blockIdx.x / bpf, b = blockIdx.x % bpf;
  const int r = threadIdx.x / tpr, v = (threadIdx.x % tpr) * V;
  long lo, hi;
  row_range(rows_per_frame, b, bpf, &lo, &hi);
  // two-level fp32 accumulation (runs of 32 rows, then a sum of runs) keeps E[x^2] - E[x]^2 well conditioned;
  // everything across threads and blocks is fp64
  float s[V], q[V], s2[V], q2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = q[i] = s2[i] = q2[i] = 0.f;
  if (r < rpp) {
    const T* base = x + ((long)f * rows_per_frame) * ldx + c0 + v;
    int run = 0;
    long row = lo + r;
    for (; row + 3L * rpp < hi; row += 4L * rpp) {   // four independent 16-byte loads in flight per thread
      float t[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) Vec<T>::ld(base + (row + (long)u * rpp) * ldx, t[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) { s[i] += t[u][i]; q[i] = fmaf(t[u][i], t[u][i], q[i]); }
      if (++run == 8) {
#pragma unroll
        for (int i = 0; i < V; ++i) { s2[i] += s[i]; q2[i] += q[i]; s[i] = q[i] = 0.f; }
        run = 0;
      }
    }
    for (; row < hi; row += rpp) {
      float t[V];
      Vec<T>::ld(base + row * ldx, t);
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] += t[i]; q[i] = fmaf(t[i], t[i], q[i]); }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sm[(r * 2 + 0) * c + v + i] = (double)s2[i] + (double)s[i];
      sm[(r * 2 + 1) * c + v + i] = (double)q2[i] + (double)q[i];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * c; j += kThreads) {
    double a = 0.0;
    for (int rr = 0; rr < rpp; ++rr) a += sm[rr * 2 * c + j];
    ws[(long)blockIdx.x * 2 * c + j] = a;
  }
}

// one warp per output: lanes walk the per-block partials, then a shuffle tree (fixed order)
__global__ void bn_stats_final_kernel(const double* __restrict__ ws, int frames, int bpf, int c,
                                      double* __restrict__ stats, int ld_stats, int s0) {
  const int total = frames * 2 * c;
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= total) return;
  const int f = i / (2 * c), j = i % (2 * c);
  double a = 0.0;
  for (int b = lane; b < bpf; b += 32) a += ws[((long)(f * bpf + b)) * 2 * c + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) stats[((long)f * 2 + j / c) * ld_stats + s0 + j % c] = a;
}

// scale/shift of one BatchNorm over `frames` frames of statistics (training) or from the running buffers.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int ld_stats, int s0, int frames, double count,
                                   int c, int cp, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* running_mean, float* running_var, int training,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_rstd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cp) return;
  if (j >= c) {
    scale_shift[j] = 0.f;
    scale_shift[cp + j] = 0.f;
    return;
  }
  double mean, var;
  if (training) {
    double s = 0.0, q = 0.0;
    for (int f = 0; f < frames; ++f) {
      s += stats[((long)f * 2 + 0) * ld_stats + s0 + j];
      q += stats[((long)f * 2 + 1) * ld_stats + s0 + j];
    }
    mean = s / count;
    var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (running_mean) {   // nn.BatchNorm3d: running stats with the unbiased variance (duf_net.py:198,201)
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[j] = (float)((1.0 - momentum) * running_mean[j] + momentum * mean);
      running_var[j] = (float)((1.0 - momentum) * running_var[j] + momentum * unb);
    }
  } else {
    mean = running_mean[j];
    var = running_var[j];
  }
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double sc = (double)gamma[j] * rstd;
  scale_shift[j] = (float)sc;
  scale_shift[cp + j] = (float)((double)beta[j] - mean * sc);
  mean_rstd[j] = (float)mean;
  mean_rstd[c + j] = (float)rstd;
}

// ---- normalise + ReLU ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_relu_kernel(const T* __restrict__ x, int ldx, int c0, int c, long rows,
                                                           const float* __restrict__ scale_shift, int cp,
                                                           T* __restrict__ y) {
  constexpr int V = Vec<T>::V;
  const int tpr = cp / V, rpp = kThreads / tpr;
  const int r = threadIdx.x / tpr, v = (threadIdx.x % tpr) * V;
  if (r >= rpp) return;
  float sc[V], sh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { sc[i] = scale_shift[v + i]; sh[i] = scale_shift[cp + v + i]; }
  const bool pad = v >= c;
  const long stride = (long)gridDim.x * rpp;
  long row = (long)blockIdx.x * rpp + r;
  if (pad) {
    float z[V];
#pragma unroll
    for (int i = 0; i < V; ++i) z[i] = 0.f;
    for (; row < rows; row += stride) Vec<T>::st(y + row * cp + v, z);
    return;
  }
  for (; row + 3 * stride < rows; row += 4 * stride) {   // four independent 16-byte loads in flight per thread
    float t[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) Vec<T>::ld(x + (row + u * stride) * ldx + c0 + v, t[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < V; ++i) t[u][i] = fmaxf(fmaf(t[u][i], sc[i], sh[i]), 0.f);
      Vec<T>::st(y + (row + u * stride) * cp + v, t[u]);
    }
  }
  for (; row < rows; row += stride) {
    float t[V];
    Vec<T>::ld(x + row * ldx + c0 + v, t);
#pragma unroll
    for (int i = 0; i < V; ++i) t[i] = fmaxf(fmaf(t[i], sc[i], sh[i]), 0.f);
    Vec<T>::st(y + row * cp + v, t);
  }
}

// backward, pass 1: per-block partial sums of g and g*xhat (g = dy where the forward output was positive)
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_relu_bwd_reduce_kernel(const T* __restrict__ dy, int ld_dy,
                                                                      const T* __restrict__ x, int ldx, int c0, int c,
                                                                      long rows, const float* __restrict__ scale_shift,
                                                                      int cp, const float* __restrict__ mean_rstd,
                                                                      float* __restrict__ ws) {
  constexpr int V = Vec<T>::V;
  extern __shared__ float smf[];   // [rpp][2][c]
  const int tpr = c / V, rpp = kThreads / tpr;
  const int r = threadIdx.x / tpr, v = (threadIdx.x % tpr) * V;
  long lo, hi;
  row_range(rows, blockIdx.x, gridDim.x, &lo, &hi);
  if (r < rpp) {
    float sc[V], sh[V], mu[V], rs[V], sg[V], sx[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sc[i] = scale_shift[v + i]; sh[i] = scale_shift[cp + v + i];
      mu[i] = mean_rstd[v + i]; rs[i] = mean_rstd[c + v + i];
      sg[i] = sx[i] = 0.f;
    }
    long row = lo + r;
    for (; row + rpp < hi; row += 2L * rpp) {   // four independent 16-byte loads in flight per thread
      float t[2][V], g[2][V];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        Vec<T>::ld(x + (row + (long)u * rpp) * ldx + c0 + v, t[u]);
        Vec<T>::ld(dy + (row + (long)u * rpp) * ld_dy + v, g[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float gi = fmaf(t[u][i], sc[i], sh[i]) > 0.f ? g[u][i] : 0.f;
          sg[i] += gi;
          sx[i] = fmaf(gi, (t[u][i] - mu[i]) * rs[i], sx[i]);
        }
    }
    for (; row < hi; row += rpp) {
      float t[V], g[V];
      Vec<T>::ld(x + row * ldx + c0 + v, t);
      Vec<T>::ld(dy + row * ld_dy + v, g);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float gi = fmaf(t[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
        sg[i] += gi;
        sx[i] = fmaf(gi, (t[i] - mu[i]) * rs[i], sx[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      smf[(r * 2 + 0) * c + v + i] = sx[i];   // row 0: d gamma
      smf[(r * 2 + 1) * c + v + i] = sg[i];   // row 1: d beta
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * c; j += kThreads) {
    float a = 0.f;
    for (int rr = 0; rr < rpp; ++rr) a += smf[rr * 2 * c + j];
    ws[(long)blockIdx.x * 2 * c + j] = a;
  }
}

__global__ void bn_relu_bwd_final_kernel(const float* __restrict__ ws, int blocks, int c, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per output
  if (j >= 2 * c) return;
  double a = 0.0;
  for (int b = lane; b < blocks; b += 32) a += ws[(long)b * 2 * c + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[j] = (float)a;
}

// backward, pass 2: dx = gamma*rstd * (g - mean(g) - xhat * mean(g*xhat)), written or accumulated into a window
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_relu_bwd_apply_kernel(const T* __restrict__ dy, int ld_dy,
                                                                     const T* __restrict__ x, int ldx, int c0, int c,
                                                                     long rows, const float* __restrict__ scale_shift,
                                                                     int cp, const float* __restrict__ mean_rstd,
                                                                     const float* __restrict__ sums, float inv_count,
                                                                     T* __restrict__ dx, int ld_dx, int c0_dx, int cp_dx,
                                                                     int accumulate) {
  constexpr int V = Vec<T>::V;
  const int tpr = cp_dx / V, rpp = kThreads / tpr;
  const int r = threadIdx.x / tpr, v = (threadIdx.x % tpr) * V;
  if (r >= rpp) return;
  const bool pad = v >= c;
  float sc[V], sh[V], mu[V], rs[V], mg[V], mx[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if (pad) { sc[i] = sh[i] = mu[i] = rs[i] = mg[i] = mx[i] = 0.f; continue; }
    sc[i] = scale_shift[v + i]; sh[i] = scale_shift[cp + v + i];
    mu[i] = mean_rstd[v + i]; rs[i] = mean_rstd[c + v + i];
    mx[i] = sums[v + i] * inv_count;       // mean of g * xhat
    mg[i] = sums[c + v + i] * inv_count;   // mean of g
  }
  const long stride = (long)gridDim.x * rpp;
  long row = (long)blockIdx.x * rpp + r;
  if (pad) {
    float z[V];
#pragma unroll
    for (int i = 0; i < V; ++i) z[i] = 0.f;
    for (; row < rows; row += stride) Vec<T>::st(dx + row * ld_dx + c0_dx + v, z);
    return;
  }
  for (; row + stride < rows; row += 2 * stride) {   // two rows = four to six independent 16-byte loads in flight
    float t[2][V], g[2][V], o[2][V];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      Vec<T>::ld(x + (row + u * stride) * ldx + c0 + v, t[u]);
      Vec<T>::ld(dy + (row + u * stride) * ld_dy + v, g[u]);
      if (accumulate) Vec<T>::ld(dx + (row + u * stride) * ld_dx + c0_dx + v, o[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float gi = fmaf(t[u][i], sc[i], sh[i]) > 0.f ? g[u][i] : 0.f;
        const float d = sc[i] * (gi - mg[i] - (t[u][i] - mu[i]) * rs[i] * mx[i]);
        o[u][i] = accumulate ? o[u][i] + d : d;
      }
      Vec<T>::st(dx + (row + u * stride) * ld_dx + c0_dx + v, o[u]);
    }
  }
  for (; row < rows; row += stride) {
    float t[V], g[V], o[V];
    T* dst = dx + row * ld_dx + c0_dx + v;
    Vec<T>::ld(x + row * ldx + c0 + v, t);
    Vec<T>::ld(dy + row * ld_dy + v, g);
    if (accumulate) Vec<T>::ld(dst, o);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float gi = fmaf(t[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
      const float d = sc[i] * (gi - mg[i] - (t[i] - mu[i]) * rs[i] * mx[i]);
      o[i] = accumulate ? o[i] + d : d;
    }
    Vec<T>::st(dst, o);
  }
}

// ---- dynamic upsampling filter tail (duf_net.py:66-97) -------------------------------------------------------
// One thread per (low-resolution pixel, sub-pixel p): softmax over the sf*sf taps of logits[pix][k*r*r + p],
// applied to the sf x sf neighbourhood of the centre frame (zero padded), + residual[pix][c*r*r + p], stored at
// the pixel-shuffled position of the NCHW output.
template <typename T>
__global__ void __launch_bounds__(kThreads) duf_filter_kernel(const T* __restrict__ logits, int ld_l,
                                                              const T* __restrict__ res, int ld_r,
                                                              const float* __restrict__ x, int n, int cin, int h, int w,
                                                              int sf, int r, float* __restrict__ y) {
  const int rr = r * r, half = sf / 2, taps = sf * sf;
  const long total = (long)n * h * w * rr;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long)gridDim.x * kThreads) {
    const int p = (int)(i % rr);
    const long pix = i / rr;
    const int x0 = (int)(pix % w), y0 = (int)((pix / w) % h), nn = (int)(pix / ((long)w * h));
    const T* lg = logits + pix * ld_l + p;
    float mx = -INFINITY;
    for (int k = 0; k < taps; ++k) mx = fmaxf(mx, Elem<T>::ld(lg + k * rr));
    const int Y = y0 * r + p / r, X = x0 * r + p % r;
    for (int c = 0; c < cin; ++c) {
      const float* xc = x + ((long)nn * cin + c) * h * w;
      float se = 0.f, sx = 0.f;
      for (int k = 0; k < taps; ++k) {
        const float e = __expf(Elem<T>::ld(lg + k * rr) - mx);
        const int yy = y0 + k / sf - half, xx = x0 + k % sf - half;
        const float xv = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(xc + (long)yy * w + xx) : 0.f;
        se += e;
        sx = fmaf(e, xv, sx);
      }
      y[(((long)nn * cin + c) * h * r + Y) * ((long)w * r) + X] = sx / se + Elem<T>::ld(res + pix * ld_r + c * rr + p);
    }
  }
}

// SF = filter size known at compile time (logits and tap gradients of a thread stay in registers: one read of the
// logits), SF = 0: any size (three passes).
template <typename T, int SF>
__global__ void __launch_bounds__(kThreads) duf_filter_bwd_kernel(const T* __restrict__ logits, int ld_l,
                                                                  const float* __restrict__ x,
                                                                  const float* __restrict__ dy, int n, int cin, int h,
                                                                  int w, int sf_rt, int r, T* __restrict__ dlogits,
                                                                  T* __restrict__ dres, int ld_r) {
  const int sf = SF > 0 ? SF : sf_rt;
  const int rr = r * r, half = sf / 2, taps = sf * sf;
  const long total = (long)n * h * w * rr;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long)gridDim.x * kThreads) {
    const int p = (int)(i % rr);
    const long pix = i / rr;
    const int x0 = (int)(pix % w), y0 = (int)((pix / w) % h), nn = (int)(pix / ((long)w * h));
    const T* lg = logits + pix * ld_l + p;
    const int Y = y0 * r + p / r, X = x0 * r + p % r;
    const float* dyp = dy + ((long)nn * cin * h * r + Y) * ((long)w * r) + X;   // + c * (h*r) * (w*r)
    const long dy_cs = (long)h * r * w * r;
    // g_k = sum_c dy_c * x_c[k];  dlogit_k = s_k * (g_k - sum_j s_j g_j)
    auto tap_grad = [&](int k) {
      const int yy = y0 + k / sf - half, xx = x0 + k % sf - half;
      float g = 0.f;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w)
        for (int c = 0; c < cin; ++c)
          g = fmaf(__ldg(dyp + c * dy_cs), __ldg(x + (((long)nn * cin + c) * h + yy) * w + xx), g);
      return g;
    };
    if constexpr (SF > 0) {
      constexpr int TAPS = SF * SF;
      float e[TAPS], gk[TAPS];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < TAPS; ++k) { e[k] = Elem<T>::ld(lg + k * rr); mx = fmaxf(mx, e[k]); }
      float se = 0.f, sg = 0.f;
#pragma unroll
      for (int k = 0; k < TAPS; ++k) {
        e[k] = __expf(e[k] - mx);
        gk[k] = tap_grad(k);
        se += e[k];
        sg = fmaf(e[k], gk[k], sg);
      }
      const float inv = 1.f / se, dot = sg * inv;
#pragma unroll
      for (int k = 0; k < TAPS; ++k) Elem<T>::st(dlogits + pix * ld_l + k * rr + p, e[k] * inv * (gk[k] - dot));
    } else {
      float mx = -INFINITY;
      for (int k = 0; k < taps; ++k) mx = fmaxf(mx, Elem<T>::ld(lg + k * rr));
      float se = 0.f, sg = 0.f;
      for (int k = 0; k < taps; ++k) {
        const float e = __expf(Elem<T>::ld(lg + k * rr) - mx);
        se += e;
        sg = fmaf(e, tap_grad(k), sg);
      }
      const float inv = 1.f / se, dot = sg * inv;
      for (int k = 0; k < taps; ++k)
        Elem<T>::st(dlogits + pix * ld_l + k * rr + p, __expf(Elem<T>::ld(lg + k * rr) - mx) * inv * (tap_grad(k) - dot));
    }
    for (int c = 0; c < cin; ++c) Elem<T>::st(dres + pix * ld_r + c * rr + p, __ldg(dyp + c * dy_cs));
  }
}

// ---- temporal shift-add: the three temporal taps of a 3x3x3 convolution computed as extra OUTPUT columns ----
// z[frame][row][kt*G + co] holds the (1,3,3) convolution of frame `frame` with temporal slice kt of the kernel;
//   out[f][row][c0_out + co] = bias[co] + sum_kt z[f + kt - t_pad][row][kt*G + co]   (frames outside: zero)
// and, when ws != NULL, the per-frame sum / sum of squares of the stored values (BatchNorm statistics of the
// new concat slice) as bn_stats_kernel leaves them.
template <typename T>
__global__ void __launch_bounds__(kThreads) tshift_add_kernel(const T* __restrict__ z, int ldz, int G, int frames_in,
                                                              long rows_per_frame, int t_pad,
                                                              const float* __restrict__ bias, T* __restrict__ out,
                                                              int ld_out, int c0_out, int bpf, double* __restrict__ ws) {
  constexpr int V = Vec<T>::V;
  extern __shared__ double sm[];   // [rpp][2][G]
  const int tpr = G / V, rpp = kThreads / tpr;
  const int f = blockIdx.x / bpf, b = blockIdx.x % bpf;
  const int r = threadIdx.x / tpr, v = (threadIdx.x % tpr) * V;
  long lo, hi;
  row_range(rows_per_frame, b, bpf, &lo, &hi);
  float s[V], q[V], bs[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = q[i] = 0.f; bs[i] = bias ? bias[v + i] : 0.f; }
  bool ok[3];
#pragma unroll
  for (int kt = 0; kt < 3; ++kt) ok[kt] = f + kt - t_pad >= 0 && f + kt - t_pad < frames_in;
  for (long row = lo + r; row < hi; row += rpp) {
    float acc[V], t[3][V];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt)
      if (ok[kt]) Vec<T>::ld(z + ((long)(f + kt - t_pad) * rows_per_frame + row) * ldz + kt * G + v, t[kt]);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = bs[i];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt)
      if (ok[kt]) {
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += t[kt][i];
      }
    T* dst = out + ((long)f * rows_per_frame + row) * ld_out + c0_out + v;
    Vec<T>::st(dst, acc);
    if (ws) {
      float rd[V];
      if constexpr (sizeof(T) == 2) {
#pragma unroll
        for (int i = 0; i < V; ++i) rd[i] = __bfloat162float(__float2bfloat16_rn(acc[i]));
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) rd[i] = acc[i];
      }
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] += rd[i]; q[i] = fmaf(rd[i], rd[i], q[i]); }
    }
  }
  if (!ws) return;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sm[(r * 2 + 0) * G + v + i] = (double)s[i];
    sm[(r * 2 + 1) * G + v + i] = (double)q[i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * G; j += kThreads) {
    double a = 0.0;
    for (int rr = 0; rr < rpp; ++rr) a += sm[rr * 2 * G + j];
    ws[(long)blockIdx.x * 2 * G + j] = a;
  }
}

// inverse of the shift-add, for the weight gradient of the same convolution in its (1,3,3) form:
//   dz[f][row][kt*G + co] = dy[f - kt + t_pad][row][c0 + co]   (zero for frames outside [0, frames_out) and for the
// channel padding [3G, ldz)), f < frames_in.
template <typename T>
__global__ void __launch_bounds__(kThreads) tshift_gather_kernel(const T* __restrict__ dy, int ld_dy, int c0, int G,
                                                                 int frames_out, long rows_per_frame, int t_pad,
                                                                 T* __restrict__ dz, int ldz, int frames_in) {
  constexpr int V = Vec<T>::V;
  const int tpr = ldz / V;
  const long total = (long)frames_in * rows_per_frame * tpr;
  for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long)gridDim.x * kThreads) {
    const long frow = i / tpr;
    const int ch = (int)(i - frow * tpr) * V;
    const int f = (int)(frow / rows_per_frame);
    const long row = frow - (long)f * rows_per_frame;
    const int kt = ch / G, fo = f - kt + t_pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (kt < 3 && fo >= 0 && fo < frames_out)
      v = *reinterpret_cast<const uint4*>(dy + ((long)fo * rows_per_frame + row) * ld_dy + c0 + (ch - kt * G));
    *reinterpret_cast<uint4*>(dz + frow * ldz + ch) = v;
  }
}

int stats_bpf(int frames, long rows_per_frame) {
  long bpf = (rows_per_frame + 255) / 256;
  const long cap = (8L * num_sms() + frames - 1) / frames;
  if (bpf > cap) bpf = cap;
  if (bpf < 1) bpf = 1;
  return (int)bpf;
}

int reduce_blocks(long rows) {
  long b = (rows + 255) / 256;
  const long cap = 2L * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T>
bool window_ok(int ld, int c0, int c) {
  constexpr int V = Vec<T>::V;
  return c > 0 && c % V == 0 && c0 % V == 0 && ld % V == 0 && c / V <= kThreads;
}

#define VSR_DISPATCH_DTYPE(dtype, who, ...)                         \
  if ((dtype) == VSR_F32) {                                         \
    typedef float T;                                                \
    __VA_ARGS__                                                     \
  } else if ((dtype) == VSR_BF16) {                                 \
    typedef __nv_bfloat16 T;                                        \
    __VA_ARGS__                                                     \
  } else {                                                          \
    VSR_CHECK_ARG(false, "%s: bad dtype %d", who, (int)(dtype));    \
  }

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_copy_window(const void* src, int32_t ld_src, int32_t c0_src, void* dst, int32_t ld_dst,
                               int32_t c0_dst, int32_t c, int64_t rows, int32_t dtype, void* stream) {
  VSR_CHECK_ARG(src && dst && rows > 0, "vsr_copy_window: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VSR_DISPATCH_DTYPE(dtype, "vsr_copy_window", {
    VSR_CHECK_ARG((window_ok<T>(ld_src, c0_src, c) && window_ok<T>(ld_dst, c0_dst, c)),
                  "vsr_copy_window: windows must be 16-byte aligned (ld %d/%d c0 %d/%d c %d)", ld_src, ld_dst, c0_src,
                  c0_dst, c);
    copy_window_kernel<T><<<grid_for(rows * (c / Vec<T>::V), kThreads), kThreads, 0, s>>>(
        (const T*)src, ld_src, c0_src, (T*)dst, ld_dst, c0_dst, c, rows);
  })
  VSR_CHECK_LAUNCH("vsr_copy_window");
  return VSR_OK;
}

extern "C" size_t vsr_bn_stats_workspace(int32_t frames, int64_t rows_per_frame, int32_t c) {
  if (frames <= 0 || rows_per_frame <= 0 || c <= 0) return 0;
  return (size_t)frames * stats_bpf(frames, rows_per_frame) * 2 * c * sizeof(double);
}

extern "C" int vsr_bn_stats(const void* x, int32_t dtype, int32_t ldx, int32_t c0, int32_t c, int32_t frames,
                            int64_t rows_per_frame, double* stats, int32_t ld_stats, int32_t s0, void* workspace,
                            size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(x && stats && frames > 0 && rows_per_frame > 0, "vsr_bn_stats: bad arguments");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_bn_stats_workspace(frames, rows_per_frame, c),
                "vsr_bn_stats: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int bpf = stats_bpf(frames, rows_per_frame);
  double* ws = static_cast<double*>(workspace);
  VSR_DISPATCH_DTYPE(dtype, "vsr_bn_stats", {
    VSR_CHECK_ARG(window_ok<T>(ldx, c0, c), "vsr_bn_stats: window must be 16-byte aligned (ld %d c0 %d c %d)", ldx, c0, c);
    const int rpp = kThreads / (c / Vec<T>::V);
    const size_t smem = (size_t)rpp * 2 * c * sizeof(double);
    VSR_CHECK_SUPPORTED(smem <= 48 * 1024, "vsr_bn_stats: window of %d channels too wide", c);
    bn_stats_kernel<T><<<frames * bpf, kThreads, smem, s>>>((const T*)x, ldx, c0, c, rows_per_frame, bpf, ws);
  })
  VSR_CHECK_LAUNCH("vsr_bn_stats");
  bn_stats_final_kernel<<<(frames * 2 * c + 7) / 8, 256, 0, s>>>(ws, frames, bpf, c, stats, ld_stats, s0);
  VSR_CHECK_LAUNCH("vsr_bn_stats_final");
  return VSR_OK;
}

extern "C" int vsr_tshift_add(const void* z, int32_t dtype, int32_t ldz, int32_t g, int32_t frames_in,
                              int64_t rows_per_frame, int32_t t_pad, const float* bias, void* out, int32_t ld_out,
                              int32_t c0_out, int32_t frames_out, double* stats, int32_t ld_stats, int32_t s0,
                              void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(z && out && frames_in > 0 && frames_out > 0 && rows_per_frame > 0 && (t_pad == 0 || t_pad == 1),
                "vsr_tshift_add: bad arguments");
  VSR_CHECK_ARG(ldz >= 3 * g, "vsr_tshift_add: z holds 3*g columns");
  VSR_CHECK_ARG(!stats || (workspace && workspace_bytes >= vsr_bn_stats_workspace(frames_out, rows_per_frame, g)),
                "vsr_tshift_add: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int bpf = stats_bpf(frames_out, rows_per_frame);
  double* ws = stats ? static_cast<double*>(workspace) : nullptr;
  VSR_DISPATCH_DTYPE(dtype, "vsr_tshift_add", {
    VSR_CHECK_ARG((window_ok<T>(ldz, 0, g) && window_ok<T>(ld_out, c0_out, g)),
                  "vsr_tshift_add: windows must be 16-byte aligned (ldz %d ld_out %d c0 %d g %d)", ldz, ld_out, c0_out, g);
    const int rpp = kThreads / (g / Vec<T>::V);
    const size_t smem = (size_t)rpp * 2 * g * sizeof(double);
    VSR_CHECK_SUPPORTED(smem <= 48 * 1024, "vsr_tshift_add: %d channels too wide", g);
    tshift_add_kernel<T><<<frames_out * bpf, kThreads, smem, s>>>((const T*)z, ldz, g, frames_in, rows_per_frame, t_pad,
                                                                 bias, (T*)out, ld_out, c0_out, bpf, ws);
  })
  VSR_CHECK_LAUNCH("vsr_tshift_add");
  if (stats) {
    bn_stats_final_kernel<<<(frames_out * 2 * g + 7) / 8, 256, 0, s>>>(ws, frames_out, bpf, g, stats, ld_stats, s0);
    VSR_CHECK_LAUNCH("vsr_tshift_add_final");
  }
  return VSR_OK;
}

extern "C" int vsr_tshift_gather(const void* dy, int32_t dtype, int32_t ld_dy, int32_t c0, int32_t g, int32_t frames_out,
                                 int64_t rows_per_frame, int32_t t_pad, void* dz, int32_t ldz, int32_t frames_in,
                                 void* stream) {
  VSR_CHECK_ARG(dy && dz && frames_in > 0 && frames_out > 0 && rows_per_frame > 0 && (t_pad == 0 || t_pad == 1) &&
                    ldz >= 3 * g, "vsr_tshift_gather: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VSR_DISPATCH_DTYPE(dtype, "vsr_tshift_gather", {
    VSR_CHECK_ARG((window_ok<T>(ld_dy, c0, g) && ldz % Vec<T>::V == 0),
                  "vsr_tshift_gather: windows must be 16-byte aligned (ld_dy %d c0 %d g %d ldz %d)", ld_dy, c0, g, ldz);
    tshift_gather_kernel<T><<<grid_for((long)frames_in * rows_per_frame * (ldz / Vec<T>::V), kThreads), kThreads, 0, s>>>(
        (const T*)dy, ld_dy, c0, g, frames_out, rows_per_frame, t_pad, (T*)dz, ldz, frames_in);
  })
  VSR_CHECK_LAUNCH("vsr_tshift_gather");
  return VSR_OK;
}

extern "C" int vsr_bn_finalize(const double* stats, int32_t ld_stats, int32_t s0, int32_t frames,
                               int64_t rows_per_frame, int32_t c, int32_t cp, const float* gamma, const float* beta,
                               float eps, float momentum, float* running_mean, float* running_var, int32_t training,
                               float* scale_shift, float* mean_rstd, void* stream) {
  VSR_CHECK_ARG(gamma && beta && scale_shift && mean_rstd && c > 0 && cp >= c, "vsr_bn_finalize: bad arguments");
  VSR_CHECK_ARG(training ? (stats != nullptr && frames > 0 && rows_per_frame > 0) : (running_mean && running_var),
                "vsr_bn_finalize: missing statistics");
  bn_finalize_kernel<<<(cp + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, ld_stats, s0, frames, (double)frames * (double)rows_per_frame, c, cp, gamma, beta, eps, momentum,
      running_mean, running_var, training, scale_shift, mean_rstd);
  VSR_CHECK_LAUNCH("vsr_bn_finalize");
  return VSR_OK;
}

extern "C" int vsr_bn_relu(const void* x, int32_t dtype, int32_t ldx, int32_t c0, int32_t c, int64_t rows,
                           const float* scale_shift, int32_t cp, void* y, void* stream) {
  VSR_CHECK_ARG(x && y && scale_shift && rows > 0 && cp >= c, "vsr_bn_relu: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  VSR_DISPATCH_DTYPE(dtype, "vsr_bn_relu", {
    VSR_CHECK_ARG((window_ok<T>(ldx, c0, c) && window_ok<T>(cp, 0, cp)),
                  "vsr_bn_relu: windows must be 16-byte aligned (ld %d c0 %d c %d cp %d)", ldx, c0, c, cp);
    const int rpp = kThreads / (cp / Vec<T>::V);
    bn_relu_kernel<T><<<grid_for(rows, rpp * 4), kThreads, 0, s>>>((const T*)x, ldx, c0, c, rows, scale_shift, cp, (T*)y);
  })
  VSR_CHECK_LAUNCH("vsr_bn_relu");
  return VSR_OK;
}

extern "C" size_t vsr_bn_relu_bwd_workspace(int64_t rows, int32_t c) {
  if (rows <= 0 || c <= 0) return 0;
  return (size_t)reduce_blocks(rows) * 2 * c * sizeof(float);
}

extern "C" int vsr_bn_relu_bwd(const void* dy, int32_t ld_dy, const void* x, int32_t dtype, int32_t ldx, int32_t c0,
                               int32_t c, int64_t rows, const float* scale_shift, int32_t cp, const float* mean_rstd,
                               float* dgamma_dbeta, void* dx, int32_t ld_dx, int32_t c0_dx, int32_t cp_dx,
                               int accumulate, int32_t phase, const float* sums, int64_t count, void* workspace,
                               size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(dy && x && scale_shift && mean_rstd && rows > 0, "vsr_bn_relu_bwd: bad arguments");
  VSR_CHECK_ARG(phase >= 1 && phase <= 3, "vsr_bn_relu_bwd: phase must be 1 (sums), 2 (dx) or 3 (both)");
  VSR_CHECK_ARG(!(phase & 1) || (dgamma_dbeta && workspace && workspace_bytes >= vsr_bn_relu_bwd_workspace(rows, c)),
                "vsr_bn_relu_bwd: sums need dgamma_dbeta and a workspace");
  VSR_CHECK_ARG(!(phase & 2) || (dx && (sums || dgamma_dbeta) && cp_dx >= c && !(accumulate && cp_dx != c)),
                "vsr_bn_relu_bwd: dx needs an output window and the sums");
  VSR_CHECK_ARG(cp >= c, "vsr_bn_relu_bwd: bad channel counts");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = reduce_blocks(rows);
  float* ws = static_cast<float*>(workspace);
  if (!sums) sums = dgamma_dbeta;
  if (count <= 0) count = rows;
  VSR_DISPATCH_DTYPE(dtype, "vsr_bn_relu_bwd", {
    VSR_CHECK_ARG((window_ok<T>(ldx, c0, c) && window_ok<T>(ld_dy, 0, c)), "vsr_bn_relu_bwd: windows must be 16-byte aligned");
    if (phase & 1) {
      const int rpp = kThreads / (c / Vec<T>::V);
      const size_t smem = (size_t)rpp * 2 * c * sizeof(float);
      VSR_CHECK_SUPPORTED(smem <= 48 * 1024, "vsr_bn_relu_bwd: window of %d channels too wide", c);
      bn_relu_bwd_reduce_kernel<T><<<blocks, kThreads, smem, s>>>((const T*)dy, ld_dy, (const T*)x, ldx, c0, c, rows,
                                                                  scale_shift, cp, mean_rstd, ws);
      VSR_CHECK_LAUNCH("vsr_bn_relu_bwd_reduce");
      bn_relu_bwd_final_kernel<<<(2 * c + 7) / 8, 256, 0, s>>>(ws, blocks, c, dgamma_dbeta);
      VSR_CHECK_LAUNCH("vsr_bn_relu_bwd_final");
    }
    if (phase & 2) {
      VSR_CHECK_ARG(window_ok<T>(ld_dx, c0_dx, cp_dx), "vsr_bn_relu_bwd: dx window must be 16-byte aligned");
      const int rpp2 = kThreads / (cp_dx / Vec<T>::V);
      bn_relu_bwd_apply_kernel<T><<<grid_for(rows, rpp2 * 4), kThreads, 0, s>>>(
          (const T*)dy, ld_dy, (const T*)x, ldx, c0, c, rows, scale_shift, cp, mean_rstd, sums,
          (float)(1.0 / (double)count), (T*)dx, ld_dx, c0_dx, cp_dx, accumulate);
      VSR_CHECK_LAUNCH("vsr_bn_relu_bwd_apply");
    }
  })
  return VSR_OK;
}

extern "C" int vsr_duf_filter(const void* logits, int32_t ld_logits, const void* res, int32_t ld_res, int32_t dtype,
                              const float* x, int32_t n, int32_t cin, int32_t h, int32_t w, int32_t size_filter,
                              int32_t r, float* y, void* stream) {
  VSR_CHECK_ARG(logits && res && x && y && n > 0 && cin > 0 && h > 0 && w > 0, "vsr_duf_filter: bad arguments");
  VSR_CHECK_ARG(size_filter >= 1 && size_filter % 2 == 1 && r >= 1 && ld_logits >= size_filter * size_filter * r * r &&
                    ld_res >= cin * r * r, "vsr_duf_filter: bad filter size / strides");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long total = (long)n * h * w * r * r;
  VSR_DISPATCH_DTYPE(dtype, "vsr_duf_filter", {
    duf_filter_kernel<T><<<grid_for(total, kThreads), kThreads, 0, s>>>((const T*)logits, ld_logits, (const T*)res, ld_res,
                                                                      x, n, cin, h, w, size_filter, r, y);
  })
  VSR_CHECK_LAUNCH("vsr_duf_filter");
  return VSR_OK;
}

extern "C" int vsr_duf_filter_bwd(const void* logits, int32_t ld_logits, int32_t dtype, const float* x, const float* dy,
                                  int32_t n, int32_t cin, int32_t h, int32_t w, int32_t size_filter, int32_t r,
                                  void* dlogits, void* dres, int32_t ld_res, void* stream) {
  VSR_CHECK_ARG(logits && x && dy && dlogits && dres && n > 0 && cin > 0 && h > 0 && w > 0, "vsr_duf_filter_bwd: bad arguments");
  VSR_CHECK_ARG(size_filter >= 1 && size_filter % 2 == 1 && r >= 1 && ld_logits >= size_filter * size_filter * r * r &&
                    ld_res >= cin * r * r, "vsr_duf_filter_bwd: bad filter size / strides");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long pix = (long)n * h * w;
  const size_t es = dtype == VSR_F32 ? 4 : 2;
  // channel padding of both gradient maps stays zero
  if (ld_logits > size_filter * size_filter * r * r) cudaMemsetAsync(dlogits, 0, pix * ld_logits * es, s);
  if (ld_res > cin * r * r) cudaMemsetAsync(dres, 0, pix * ld_res * es, s);
  VSR_DISPATCH_DTYPE(dtype, "vsr_duf_filter_bwd", {
    const int grid = grid_for(pix * r * r, kThreads);
    if (size_filter == 5)
      duf_filter_bwd_kernel<T, 5><<<grid, kThreads, 0, s>>>((const T*)logits, ld_logits, x, dy, n, cin, h, w, size_filter, r,
                                                          (T*)dlogits, (T*)dres, ld_res);
    else if (size_filter == 3)
      duf_filter_bwd_kernel<T, 3><<<grid, kThreads, 0, s>>>((const T*)logits, ld_logits, x, dy, n, cin, h, w, size_filter, r,
                                                          (T*)dlogits, (T*)dres, ld_res);
    else
      duf_filter_bwd_kernel<T, 0><<<grid, kThreads, 0, s>>>((const T*)logits, ld_logits, x, dy, n, cin, h, w, size_filter, r,
                                                          (T*)dlogits, (T*)dres, ld_res);
  })
  VSR_CHECK_LAUNCH("vsr_duf_filter_bwd");
  return VSR_OK;
}
