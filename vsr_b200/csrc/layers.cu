// layers.cu — the two bandwidth-bound boundary convolutions of the nets:
//   first: NCHW fp32 image (Cin small) -> pixel-major features   (drf_net.py:55-56)
//   last : phase-blocked features -> NCHW fp32 image (Cout small) (drf_net.py:144,147)
// and their backward passes.  K=9*Cin resp. N=Cout are far too thin for tensor cores
// (4 FLOP/B); these are coalesced CUDA-core kernels bound by HBM.
#include "common.cuh"

namespace vsr {
namespace {

constexpr int kMaxCin = 4;
constexpr int kMaxCoutLast = 4;

// ------------------------------------------------------------------------------------------
// first conv: one thread = one pixel x 8 output channels; a warp covers 256 consecutive channels
// of one pixel (coalesced 16-byte stores).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void conv_first_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                  const float* __restrict__ wt, const float* __restrict__ bias,
                                  const float* __restrict__ slope_p, T* __restrict__ y, int cout) {
  extern __shared__ float sw[];  // [cin*9][cout] transposed weights, then bias[cout]
  float* sb = sw + cin * 9 * cout;
  for (int i = threadIdx.x; i < cin * 9 * cout; i += blockDim.x) {
    const int co = i / (cin * 9), r = i % (cin * 9);
    sw[r * cout + co] = wt[i];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const Prelu a = make_prelu(slope_p ? __ldg(slope_p) : 1.f);
  const int groups = cout / 8;
  const long total = (long)n * h * w * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int gq = (int)(i % groups);
    const long pix = i / groups;
    const int px = (int)(pix % w);
    const long q = pix / w;
    const int py = (int)(q % h);
    const int ni = (int)(q / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sb[gq * 8 + j];
    for (int ci = 0; ci < cin; ++ci) {
      const float* xp = x + ((size_t)ni * cin + ci) * h * w;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = py + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = px + kx - 1;
          float v = 0.f;
          if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(xp + (size_t)yy * w + xx);
          const float* wp = sw + ((ci * 3 + ky) * 3 + kx) * cout + gq * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
        }
      }
    }
    prelu_store8(y + pix * cout + gq * 8, acc, a);      // slope 1 (no PReLU) is the identity
  }
}

// first conv backward: block b covers a pixel range, thread = output channel; partial
// dw[co][cin*9] and db[co] in registers -> ws[b][cout][cin*9+1]; fixed-order final reduce.
template <typename T>
__global__ void conv_first_bwd_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                      const T* __restrict__ dz, int cout, long pix_per_block,
                                      float* __restrict__ ws) {
  const long total = (long)n * h * w;
  const long p0 = blockIdx.x * pix_per_block;
  long p1 = p0 + pix_per_block;
  if (p1 > total) p1 = total;
  const int stride = cin * 9 + 1;
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    float acc[kMaxCin * 9 + 1];
#pragma unroll
    for (int i = 0; i < kMaxCin * 9 + 1; ++i) acc[i] = 0.f;
    for (long p = p0; p < p1; ++p) {
      const int px = (int)(p % w);
      const long q = p / w;
      const int py = (int)(q % h);
      const int ni = (int)(q / h);
      const float g = Elem<T>::ld(dz + p * cout + co);
      acc[kMaxCin * 9] += g;
#pragma unroll
      for (int ci = 0; ci < kMaxCin; ++ci) {
        if (ci < cin) {
          const float* xp = x + ((size_t)ni * cin + ci) * h * w;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int yy = py + ky - 1;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int xx = px + kx - 1;
              float v = 0.f;
              if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(xp + (size_t)yy * w + xx);
              acc[(ci * 3 + ky) * 3 + kx] = fmaf(g, v, acc[(ci * 3 + ky) * 3 + kx]);
            }
          }
        }
      }
    }
    float* o = ws + ((size_t)blockIdx.x * cout + co) * stride;
    for (int i = 0; i < cin * 9; ++i) o[i] = acc[i];
    o[cin * 9] = acc[kMaxCin * 9];
  }
}

__global__ void conv_first_bwd_final_kernel(const float* __restrict__ ws, int blocks, int cout, int cin,
                                            float* __restrict__ dw, float* __restrict__ db, int accumulate) {
  const int stride = cin * 9 + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * stride) return;
  const int co = i / stride, r = i % stride;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += ws[((size_t)b * cout + co) * stride + r];
  float* dst = (r == cin * 9) ? (db + co) : (dw + (size_t)co * cin * 9 + r);
  *dst = accumulate ? *dst + s : s;
}

// ------------------------------------------------------------------------------------------
// first conv, v2 (cin == 1, cout <= 256): a lane owns 8 output channels and keeps their 72 weights
// and 8 biases in registers; a warp = one pixel at a time (9 broadcast loads, one 16/32-byte
// store per lane -> 512 contiguous bytes per warp).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) conv_first2_kernel(const float* __restrict__ x, int n, int h, int w,
                                                         const float* __restrict__ wt, const float* __restrict__ bias,
                                                         const float* __restrict__ slope_p, T* __restrict__ y, int cout) {
  __shared__ float sw2[256 * 9 + 256];
  for (int i = threadIdx.x; i < cout * 9; i += blockDim.x) sw2[i] = wt[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sw2[256 * 9 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int groups = cout >> 3;
  // a warp covers ppw pixels x `groups` 8-channel groups: 1 pixel at 256 channels, 4 at 64 (no idle lanes)
  const int ppw = (groups <= 16 && (groups & (groups - 1)) == 0) ? 32 / groups : 1;
  const int gl = ppw > 1 ? lane % groups : lane, sub = ppw > 1 ? lane / groups : 0;
  const bool live = gl < groups;
  float wr[8][9], br[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    br[j] = live ? sw2[256 * 9 + gl * 8 + j] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = live ? sw2[(gl * 8 + j) * 9 + t] : 0.f;
  }
  const Prelu a = make_prelu(slope_p ? __ldg(slope_p) : 1.f);
  const long total = (long)n * h * w;
  const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long pix0 = warp0 * ppw; pix0 < total; pix0 += nwarps * ppw) {
    const long pix = pix0 + sub;
    if (pix >= total) continue;
    const int px = (int)(pix % w);
    const long q = pix / w;
    const int py = (int)(q % h);
    const float* xp = x + (q / h) * (size_t)h * w;
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = py + ky - 1, xx = px + kx - 1;
        const bool ok = yy >= 0 && yy < h && xx >= 0 && xx < w;
        const float t = __ldg(xp + (size_t)min(max(yy, 0), h - 1) * w + min(max(xx, 0), w - 1));
        v[ky * 3 + kx] = ok ? t : 0.f;
      }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = br[j];
#pragma unroll
      for (int t = 0; t < 9; ++t) s = fmaf(v[t], wr[j][t], s);
      acc[j] = s;
    }
    if (live) prelu_store8(y + pix * cout + gl * 8, acc, a);
  }
}

// first conv backward, v2 (cin == 1): a block walks 64-pixel chunks (chunk = blockIdx.x, += gridDim.x); its 1024
// threads are LANES pixel-lanes x (1024 / LANES) channel threads (4 x 256 for up to 256 channels, 16 x 64 for up to
// 64); a thread accumulates dw[co][9], db[co] over its pixels of every chunk in registers; fixed-order smem fold over
// the pixel-lanes -> ws[block][cout][10]; parallel fixed-order final reduce over the (few) blocks.
constexpr int kF2Pix = 64;
template <typename T, int LANES>
__global__ void __launch_bounds__(1024) conv_first2_bwd_kernel(const float* __restrict__ x, int n, int h, int w,
                                                              const T* __restrict__ dz, int cout, float* __restrict__ ws) {
  constexpr int CT = 1024 / LANES;       // channel threads
  constexpr int PPL = kF2Pix / LANES;    // pixels per lane and chunk
  __shared__ float xs[kF2Pix][9];
  __shared__ float part[1024][10];       // [lane][co]
  const long total = (long)n * h * w;
  const long chunks = (total + kF2Pix - 1) / kF2Pix;
  const int co = threadIdx.x % CT, pl4 = threadIdx.x / CT;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  for (long chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    const long p0 = chunk * kF2Pix;
    __syncthreads();
    for (int i = threadIdx.x; i < kF2Pix * 9; i += blockDim.x) {
      const int pl = i / 9, t = i % 9;
      const long p = p0 + pl;
      float v = 0.f;
      if (p < total) {
        const int px = (int)(p % w);
        const long q = p / w;
        const int py = (int)(q % h);
        const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(x + (q / h) * (size_t)h * w + (size_t)yy * w + xx);
      }
      xs[pl][t] = v;
    }
    __syncthreads();
    if (co < cout) {
      float g[PPL];
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        const long p = p0 + pl4 * PPL + i;
        g[i] = p < total ? Elem<T>::ld(dz + p * cout + co) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        const int pl = pl4 * PPL + i;
        acc[9] += g[i];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = fmaf(g[i], xs[pl][t], acc[t]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 10; ++i) part[pl4 * CT + co][i] = acc[i];
  __syncthreads();
  for (int i = threadIdx.x; i < cout * 10; i += blockDim.x) {
    const int c = i / 10, r = i % 10;
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < LANES; ++l) t += part[l * CT + c][r];
    ws[(size_t)blockIdx.x * cout * 10 + i] = t;
  }
}
// out[i] (+)= sum_b ws[b][i], i < n: 256-thread blocks, 64 outputs x 4 block-lanes, fixed order
__global__ void __launch_bounds__(256) rows_reduce_kernel(const float* __restrict__ ws, int blocks, int n,
                                                         float* __restrict__ dw, float* __restrict__ db, int cout,
                                                         int accumulate) {
  __shared__ float sm[4][64];
  const int col = threadIdx.x & 63, part = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + col;
  float s = 0.f;
  if (i < n) {
    int b = part;
    for (; b + 28 < blocks; b += 32) {     // eight loads in flight, additions in the order b = part, part+4, ...
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(ws + (size_t)(b + 4 * u) * n + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; b < blocks; b += 4) s += ws[(size_t)b * n + i];
  }
  sm[part][col] = s;
  __syncthreads();
  if (part == 0 && i < n) {
    const float t = ((sm[0][col] + sm[1][col]) + sm[2][col]) + sm[3][col];
    const int c = i / 10, r = i % 10;
    float* dst = (r == 9) ? (db + c) : (dw + (size_t)c * 9 + r);
    *dst = accumulate ? *dst + t : t;
  }
}

// ------------------------------------------------------------------------------------------
// last conv: one warp = one high-resolution output pixel; lanes stride the C channels.
// x is phase-blocked [n][h][w][r*r][C]; slot_of[py*r+px] gives the phase slot.
// ------------------------------------------------------------------------------------------
struct LastGeom {
  int n, h, w, r, c, cout;
  int slot_of[64];  // r*r <= 64
};

template <typename T>
__device__ __forceinline__ const T* hr_ptr(const T* x, const LastGeom& g, int ni, int Y, int X) {
  const int y = Y / g.r, py = Y % g.r, xx = X / g.r, px = X % g.r;
  const int slot = g.slot_of[py * g.r + px];
  return x + ((((size_t)ni * g.h + y) * g.w + xx) * (g.r * g.r) + slot) * g.c;
}

template <typename T>
__global__ void conv_last_kernel(const T* __restrict__ x, const __grid_constant__ LastGeom g,
                                 const float* __restrict__ wt, const float* __restrict__ bias,
                                 float* __restrict__ y) {
  extern __shared__ float sw[];  // [cout][9][C]
  for (int i = threadIdx.x; i < g.cout * g.c * 9; i += blockDim.x) {
    const int co = i / (g.c * 9), rem = i % (g.c * 9), ci = rem / 9, tap = rem % 9;
    sw[(co * 9 + tap) * g.c + ci] = wt[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int H = g.h * g.r, W = g.w * g.r;
  const long total = (long)g.n * H * W;
  const long warp0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long p = warp0; p < total; p += nwarps) {
    const int X = (int)(p % W);
    const long q = p / W;
    const int Y = (int)(q % H);
    const int ni = (int)(q / H);
    float acc[kMaxCoutLast];
#pragma unroll
    for (int co = 0; co < kMaxCoutLast; ++co) acc[co] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = Y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = X + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const T* xp = hr_ptr(x, g, ni, yy, xx);
        for (int ci = lane; ci < g.c; ci += 32) {
          const float v = Elem<T>::ld(xp + ci);
#pragma unroll
          for (int co = 0; co < kMaxCoutLast; ++co)
            if (co < g.cout) acc[co] = fmaf(v, sw[(co * 9 + ky * 3 + kx) * g.c + ci], acc[co]);
        }
      }
    }
#pragma unroll
    for (int co = 0; co < kMaxCoutLast; ++co) {
      if (co < g.cout) {
        const float s = warp_sum(acc[co]);
        if (lane == 0) y[(((size_t)ni * g.cout + co) * H + Y) * W + X] = s + (bias ? __ldg(bias + co) : 0.f);
      }
    }
  }
}

// last conv backward: per HR pixel P (one warp): dx(P,c) = sum_{co,tap} dy(co, P - off(tap)) * w[co][c][tap]
// and dw[co][c][tap] += dy(co,P) * x(P + off(tap), c); db[co] += dy(co,P).
// dw/db partials: per-warp registers -> per-block smem (fixed warp order) -> ws[block][...] -> final.
template <typename T>
__global__ void conv_last_bwd_kernel(const T* __restrict__ x, const __grid_constant__ LastGeom g,
                                     const float* __restrict__ wt, const float* __restrict__ dy,
                                     T* __restrict__ dx, float* __restrict__ ws, int ch_per_lane) {
  // smem: weights [cout][9][C] then block partial [nwarps][cout*9*C + cout]
  extern __shared__ float sm[];
  float* sw = sm;
  const int wsz = g.cout * 9 * g.c;
  for (int i = threadIdx.x; i < g.cout * g.c * 9; i += blockDim.x) {
    const int co = i / (g.c * 9), rem = i % (g.c * 9), ci = rem / 9, tap = rem % 9;
    sw[(co * 9 + tap) * g.c + ci] = wt[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int H = g.h * g.r, W = g.w * g.r;
  const long total = (long)g.n * H * W;
  const long warp0 = (long)blockIdx.x * nw + warp;
  const long nwarps = (long)gridDim.x * nw;
  // per-lane dw accumulators: channels ci = lane + 32*k, k < ch_per_lane (<=4), cout==1..kMaxCoutLast
  // restricted to cout*ch_per_lane <= 4 groups of 9 to bound registers.
  float dwacc[4][9];
  float dbacc[kMaxCoutLast];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) dwacc[i][t] = 0.f;
#pragma unroll
  for (int co = 0; co < kMaxCoutLast; ++co) dbacc[co] = 0.f;

  for (long p = warp0; p < total; p += nwarps) {
    const int X = (int)(p % W);
    const long q = p / W;
    const int Y = (int)(q % H);
    const int ni = (int)(q / H);
    // ---- dx ----
    float dxa[4] = {0.f, 0.f, 0.f, 0.f};
    for (int co = 0; co < g.cout; ++co) {
      const float* dyp = dy + ((size_t)ni * g.cout + co) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = Y - (ky - 1);
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = X - (kx - 1);
          if (xx < 0 || xx >= W) continue;
          const float gval = __ldg(dyp + (size_t)yy * W + xx);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ci = lane + 32 * k;
            if (k < ch_per_lane && ci < g.c) dxa[k] = fmaf(gval, sw[(co * 9 + ky * 3 + kx) * g.c + ci], dxa[k]);
          }
        }
      }
    }
    T* dxp = const_cast<T*>(hr_ptr(dx, g, ni, Y, X));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ci = lane + 32 * k;
      if (k < ch_per_lane && ci < g.c) Elem<T>::st(dxp + ci, dxa[k]);
    }
    // ---- dw / db (cout * ch_per_lane <= 4) ----
    for (int co = 0; co < g.cout; ++co) {
      const float gval = __ldg(dy + (((size_t)ni * g.cout + co) * H + Y) * W + X);
      dbacc[co] += gval;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = Y + ky - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = X + kx - 1;
          if (xx < 0 || xx >= W) continue;
          const T* xp = hr_ptr(x, g, ni, yy, xx);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ci = lane + 32 * k;
            const int slot = co * ch_per_lane + k;
            if (k < ch_per_lane && ci < g.c && slot < 4)
              dwacc[slot][ky * 3 + kx] = fmaf(gval, Elem<T>::ld(xp + ci), dwacc[slot][ky * 3 + kx]);
          }
        }
      }
    }
  }
  // block reduce in fixed warp order
  float* part = sm + wsz;  // [nw][wsz + cout]
  const int psz = wsz + g.cout;
  for (int co = 0; co < g.cout; ++co) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ci = lane + 32 * k;
      const int slot = co * ch_per_lane + k;
      if (k < ch_per_lane && ci < g.c && slot < 4)
#pragma unroll
        for (int t = 0; t < 9; ++t) part[(size_t)warp * psz + (co * 9 + t) * g.c + ci] = dwacc[slot][t];
    }
    if (lane == 0) part[(size_t)warp * psz + wsz + co] = dbacc[co];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < psz; i += blockDim.x) {
    float s = 0.f;
    for (int wv = 0; wv < nw; ++wv) s += part[(size_t)wv * psz + i];
    ws[(size_t)blockIdx.x * psz + i] = s;
  }
}

__global__ void conv_last_bwd_final_kernel(const float* __restrict__ ws, int blocks, int cout, int c,
                                           float* __restrict__ dw, float* __restrict__ db, int accumulate) {
  const int wsz = cout * 9 * c, psz = wsz + cout;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= psz) return;
  float s = 0.f;
  int b = 0;
  for (; b + 8 <= blocks; b += 8) {        // eight loads in flight; the additions keep the order b = 0, 1, 2, ...
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(ws + (size_t)(b + u) * psz + i);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; b < blocks; ++b) s += ws[(size_t)b * psz + i];
  float* dst;
  if (i < wsz) {
    const int co = i / (9 * c), rem = i % (9 * c), tap = rem / c, ci = rem % c;
    dst = dw + ((size_t)co * c + ci) * 9 + tap;
  } else {
    dst = db + (i - wsz);
  }
  *dst = accumulate ? *dst + s : s;
}

// ------------------------------------------------------------------------------------------
// last conv, v2 (cout == 1, C in {32, 64, 128}): one warp = one SB x SB sub-block of HR pixels
// (SB = 4, or r for r < 4) with its 1-pixel halo held in registers; every lane owns CPL channels
// of every halo pixel (coalesced 64/128-byte loads), accumulates the SB*SB partial dot products
// and the warp folds them with a 16-shuffle transpose-reduce.
// ------------------------------------------------------------------------------------------
constexpr int kSBMax = 4;
constexpr int kHalo = (kSBMax + 2) * (kSBMax + 2);   // 36

struct LastGeom2 {
  int n, h, w, r, c, sb, nsb;       // nsb = sub-blocks per LR block side (r / sb)
  int slot_of[64];
};

template <typename T, int CPL>
__device__ __forceinline__ void load_cpl(const T* p, float* f) {
  if constexpr (sizeof(T) == 2) {
    if constexpr (CPL == 2) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
      f[0] = bf16_lo(u); f[1] = bf16_hi(u);
    } else {
      f[0] = __bfloat162float(p[0]);
    }
  } else {
    if constexpr (CPL == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p));
      f[0] = v.x; f[1] = v.y;
    } else {
      f[0] = __ldg(p);
    }
  }
}
template <typename T, int CPL>
__device__ __forceinline__ void store_cpl(T* p, const float* f) {
  if constexpr (sizeof(T) == 2) {
    if constexpr (CPL == 2) *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(f[0], f[1]);
    else p[0] = __float2bfloat16_rn(f[0]);
  } else {
    if constexpr (CPL == 2) *reinterpret_cast<float2*>(p) = make_float2(f[0], f[1]);
    else p[0] = f[0];
  }
}

// fold 16 per-lane values over the 32 lanes: afterwards lanes 2q and 2q+1 hold the sum of value q
__device__ __forceinline__ float transpose_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 8];
    const float keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 4];
    const float keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 2];
    const float keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct SubBlock {
  int ni, Y0, X0;   // top-left HR pixel of the sub-block
  int by, bx;       // its LR block
};
__device__ __forceinline__ SubBlock decode_sb(const LastGeom2& g, long item) {
  SubBlock b;
  const int sx = (int)(item % g.nsb); item /= g.nsb;
  const int sy = (int)(item % g.nsb); item /= g.nsb;
  const int x = (int)(item % g.w); item /= g.w;
  const int y = (int)(item % g.h);
  b.ni = (int)(item / g.h);
  b.Y0 = y * g.r + sy * g.sb;
  b.X0 = x * g.r + sx * g.sb;
  b.by = y;
  b.bx = x;
  return b;
}
template <typename T>
__device__ __forceinline__ const T* hr_ptr2(const T* x, const LastGeom2& g, int ni, int Y, int X) {
  const int y = Y / g.r, py = Y - y * g.r, xx = X / g.r, px = X - xx * g.r;
  return x + ((((size_t)ni * g.h + y) * g.w + xx) * (g.r * g.r) + g.slot_of[py * g.r + px]) * g.c;
}

// per-item halo geometry: element offset (in pixels*slots) of every halo pixel and its validity;
// coordinates are clamped so that every load is legal and can be issued unconditionally.
struct HaloIdx {
  size_t rowbase[kSBMax + 2];   // ((ni*h + y) * w) * r2  for halo row i
  int rowphase[kSBMax + 2];     // py * r
  int colbase[kSBMax + 2];      // x * r2 for halo column j
  int colphase[kSBMax + 2];     // px
  uint32_t rowok, colok;        // validity bit masks
};
// slot of phase (py, px) in the nested (Z-order) layout of a 4x up-scaled map (drf_plan.phase_table(4))
__host__ __device__ constexpr int zslot4(int py, int px) {
  return ((py >> 1) << 3) | ((px >> 1) << 2) | ((py & 1) << 1) | (px & 1);
}
// ZR4: r == 4, Z-order slots, sub-block == LR block: every phase below is a compile-time constant, the
// halo rows / columns 1..4 share the centre block and the address arithmetic folds to 9 block bases.
template <bool ZR4>
__device__ __forceinline__ void halo_index(const LastGeom2& g, const SubBlock& b, HaloIdx* hi) {
  if constexpr (ZR4) {
    const int H = g.h * 4, W = g.w * 4;
    hi->rowok = hi->colok = 0;
#pragma unroll
    for (int i = 0; i < kSBMax + 2; ++i) {
      const int Y = b.Y0 + i - 1, X = b.X0 + i - 1;
      const int yy = min(max(b.by + (i == 0 ? -1 : (i == 5 ? 1 : 0)), 0), g.h - 1);
      const int xx = min(max(b.bx + (i == 0 ? -1 : (i == 5 ? 1 : 0)), 0), g.w - 1);
      hi->rowbase[i] = ((size_t)b.ni * g.h + yy) * g.w * 16;
      hi->colbase[i] = xx * 16;
      hi->rowphase[i] = 0;
      hi->colphase[i] = 0;
      hi->rowok |= ((Y >= 0 && Y < H) ? 1u : 0u) << i;
      hi->colok |= ((X >= 0 && X < W) ? 1u : 0u) << i;
    }
    return;
  }
  // the sub-block lies inside one LR block: a halo coordinate is at most one block away, so the
  // (block, phase) split needs comparisons only
  const int H = g.h * g.r, W = g.w * g.r, r2 = g.r * g.r, HT = g.sb + 2;
  const int y = b.by, x = b.bx, ly0 = b.Y0 - b.by * g.r, lx0 = b.X0 - b.bx * g.r;
  hi->rowok = hi->colok = 0;
#pragma unroll
  for (int i = 0; i < kSBMax + 2; ++i) {
    const int Y = b.Y0 + i - 1, X = b.X0 + i - 1;
    const bool yo = i < HT && Y >= 0 && Y < H, xo = i < HT && X >= 0 && X < W;
    int ly = ly0 + i - 1, yy = y, lx = lx0 + i - 1, xx = x;
    if (ly < 0) { ly += g.r; --yy; } else if (ly >= g.r) { ly -= g.r; ++yy; }
    if (lx < 0) { lx += g.r; --xx; } else if (lx >= g.r) { lx -= g.r; ++xx; }
    yy = min(max(yy, 0), g.h - 1);
    xx = min(max(xx, 0), g.w - 1);
    hi->rowbase[i] = ((size_t)b.ni * g.h + yy) * g.w * r2;
    hi->rowphase[i] = ly * g.r;
    hi->colbase[i] = xx * r2;
    hi->colphase[i] = lx;
    hi->rowok |= (yo ? 1u : 0u) << i;
    hi->colok |= (xo ? 1u : 0u) << i;
  }
}
template <bool ZR4>
__device__ __forceinline__ size_t halo_off(const LastGeom2& g, const HaloIdx& hi, int i, int j) {
  if constexpr (ZR4) return (hi.rowbase[i] + hi.colbase[j] + zslot4((i + 3) & 3, (j + 3) & 3)) * (size_t)g.c;
  return (hi.rowbase[i] + hi.colbase[j] + g.slot_of[hi.rowphase[i] + hi.colphase[j]]) * (size_t)g.c;
}

// rolling 3-row window over the sub-block's halo: row i of the halo for channel block c0
template <typename T, int CPL, bool ZR4>
__device__ __forceinline__ void load_halo_row(const T* __restrict__ x, const LastGeom2& g, const HaloIdx& hi, int i,
                                              int c0, float (&row)[kSBMax + 2][CPL]) {
#pragma unroll
  for (int j = 0; j < kSBMax + 2; ++j) load_cpl<T, CPL>(x + halo_off<ZR4>(g, hi, i, j) + c0, row[j]);
#pragma unroll
  for (int j = 0; j < kSBMax + 2; ++j) {
    const bool ok = ((hi.rowok >> i) & 1u) && ((hi.colok >> j) & 1u);
#pragma unroll
    for (int q = 0; q < CPL; ++q) row[j][q] = ok ? row[j][q] : 0.f;
  }
}

template <typename T, int CPL, bool ZR4 = false>
__global__ void __launch_bounds__(256, 2) conv_last2_kernel(const T* __restrict__ x, const __grid_constant__ LastGeom2 g,
                                                           const float* __restrict__ wt, const float* __restrict__ bias,
                                                           float* __restrict__ y) {
  extern __shared__ float sw[];    // [9][C]
  for (int i = threadIdx.x; i < g.c * 9; i += blockDim.x) sw[(i % 9) * g.c + i / 9] = wt[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int H = g.h * g.r, W = g.w * g.r;
  const long total = (long)g.n * g.h * g.w * g.nsb * g.nsb;
  const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const float b0 = bias ? __ldg(bias) : 0.f;
  for (long item = warp0; item < total; item += nwarps) {
    const SubBlock sbk = decode_sb(g, item);
    HaloIdx hi;
    halo_index<ZR4>(g, sbk, &hi);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int cb = 0; cb < g.c; cb += 32 * CPL) {
      const int c0 = cb + lane * CPL;
      float wr[9][CPL];
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int q = 0; q < CPL; ++q) wr[t][q] = sw[t * g.c + c0 + q];
      float rows[3][kSBMax + 2][CPL];
      load_halo_row<T, CPL, ZR4>(x, g, hi, 0, c0, rows[0]);
      load_halo_row<T, CPL, ZR4>(x, g, hi, 1, c0, rows[1]);
#pragma unroll
      for (int oy = 0; oy < kSBMax; ++oy) {
        load_halo_row<T, CPL, ZR4>(x, g, hi, oy + 2, c0, rows[(oy + 2) % 3]);
#pragma unroll
        for (int ox = 0; ox < kSBMax; ++ox) {
          float sacc = acc[oy * kSBMax + ox];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int q = 0; q < CPL; ++q) sacc = fmaf(rows[(oy + ky) % 3][ox + kx][q], wr[ky * 3 + kx][q], sacc);
          acc[oy * kSBMax + ox] = sacc;
        }
      }
    }
    const float tot = transpose_reduce16(acc, lane);
    const int q = lane >> 1, oy = q / kSBMax, ox = q % kSBMax;
    if ((lane & 1) == 0 && oy < g.sb && ox < g.sb)
      y[((size_t)sbk.ni * H + sbk.Y0 + oy) * W + sbk.X0 + ox] = tot + b0;
  }
}

// backward, part 1 (cout == 1): dx(P,c) = sum_tap dy(P - off(tap)) * w[c][tap]; one warp per sub-block,
// the dy halo in registers (broadcast loads), CPL channels per lane, coalesced stores.
template <typename T, int CPL, bool ZR4 = false>
__global__ void __launch_bounds__(256, 2) conv_last2_dx_kernel(const __grid_constant__ LastGeom2 g,
                                                              const float* __restrict__ wt, const float* __restrict__ dy,
                                                              T* __restrict__ dx) {
  extern __shared__ float sw[];    // [9][C]
  for (int i = threadIdx.x; i < g.c * 9; i += blockDim.x) sw[(i % 9) * g.c + i / 9] = wt[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int H = g.h * g.r, W = g.w * g.r;
  const long total = (long)g.n * g.h * g.w * g.nsb * g.nsb;
  const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long item = warp0; item < total; item += nwarps) {
    const SubBlock sbk = decode_sb(g, item);
    HaloIdx hi;
    halo_index<ZR4>(g, sbk, &hi);
    float gy[kHalo];
#pragma unroll
    for (int i = 0; i < kSBMax + 2; ++i)
#pragma unroll
      for (int j = 0; j < kSBMax + 2; ++j) {
        const int Yc = min(max(sbk.Y0 + i - 1, 0), H - 1), Xc = min(max(sbk.X0 + j - 1, 0), W - 1);
        const float t = __ldg(dy + ((size_t)sbk.ni * H + Yc) * W + Xc);
        gy[i * (kSBMax + 2) + j] = (((hi.rowok >> i) & 1u) && ((hi.colok >> j) & 1u)) ? t : 0.f;
      }
    for (int cb = 0; cb < g.c; cb += 32 * CPL) {
      const int c0 = cb + lane * CPL;
      float wr[9][CPL];
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int q = 0; q < CPL; ++q) wr[t][q] = sw[t * g.c + c0 + q];
#pragma unroll
      for (int oy = 0; oy < kSBMax; ++oy)
#pragma unroll
        for (int ox = 0; ox < kSBMax; ++ox) {
          if (oy < g.sb && ox < g.sb) {
            float d[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) d[q] = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const float gn = gy[(oy + 2 - ky) * (kSBMax + 2) + ox + 2 - kx];
#pragma unroll
                for (int q = 0; q < CPL; ++q) d[q] = fmaf(gn, wr[ky * 3 + kx][q], d[q]);
              }
            store_cpl<T, CPL>(dx + halo_off<ZR4>(g, hi, oy + 1, ox + 1) + c0, d);
          }
        }
    }
  }
}

// backward, part 2: dw[c][tap] += dy(P) * x(P + off(tap), c), db += dy(P): the forward's rolling x
// window; per-lane accumulators -> block fold (fixed warp order) -> ws -> final reduce.
template <typename T, int CPL, bool ZR4 = false>
__global__ void __launch_bounds__(256, 2) conv_last2_dw_kernel(const T* __restrict__ x, const __grid_constant__ LastGeom2 g,
                                                              const float* __restrict__ dy, float* __restrict__ ws) {
  extern __shared__ float sm[];    // block partial [nw][9*C + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int H = g.h * g.r, W = g.w * g.r;
  const long total = (long)g.n * g.h * g.w * g.nsb * g.nsb;
  const long warp0 = (long)blockIdx.x * nw + warp;
  const long nwarps = (long)gridDim.x * nw;
  constexpr int NCB = 4 / CPL;            // channel blocks of 32*CPL (C <= 128)
  float dwacc[NCB][9][CPL];
#pragma unroll
  for (int b = 0; b < NCB; ++b)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int q = 0; q < CPL; ++q) dwacc[b][t][q] = 0.f;
  float dbacc = 0.f;
  for (long item = warp0; item < total; item += nwarps) {
    const SubBlock sbk = decode_sb(g, item);
    HaloIdx hi;
    halo_index<ZR4>(g, sbk, &hi);
    float gc[16];
#pragma unroll
    for (int oy = 0; oy < kSBMax; ++oy)
#pragma unroll
      for (int ox = 0; ox < kSBMax; ++ox) {
        const bool ok = oy < g.sb && ox < g.sb;
        const int Yc = min(sbk.Y0 + oy, H - 1), Xc = min(sbk.X0 + ox, W - 1);
        const float t = __ldg(dy + ((size_t)sbk.ni * H + Yc) * W + Xc);
        gc[oy * kSBMax + ox] = ok ? t : 0.f;
        dbacc += ok ? t : 0.f;
      }
#pragma unroll
    for (int b = 0; b < NCB; ++b) {
      const int cb = b * 32 * CPL;
      if (cb < g.c) {
        const int c0 = cb + lane * CPL;
        float rows[3][kSBMax + 2][CPL];
        load_halo_row<T, CPL, ZR4>(x, g, hi, 0, c0, rows[0]);
        load_halo_row<T, CPL, ZR4>(x, g, hi, 1, c0, rows[1]);
#pragma unroll
        for (int oy = 0; oy < kSBMax; ++oy) {
          load_halo_row<T, CPL, ZR4>(x, g, hi, oy + 2, c0, rows[(oy + 2) % 3]);
#pragma unroll
          for (int ox = 0; ox < kSBMax; ++ox) {
            const float gv = gc[oy * kSBMax + ox];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int q = 0; q < CPL; ++q)
                  dwacc[b][ky * 3 + kx][q] = fmaf(gv, rows[(oy + ky) % 3][ox + kx][q], dwacc[b][ky * 3 + kx][q]);
          }
        }
      }
    }
  }
  const int wsz = 9 * g.c, psz = wsz + 1;
#pragma unroll
  for (int b = 0; b < NCB; ++b) {
    const int cb = b * 32 * CPL;
    if (cb < g.c)
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int q = 0; q < CPL; ++q) sm[(size_t)warp * psz + t * g.c + cb + lane * CPL + q] = dwacc[b][t][q];
  }
  if (lane == 0) sm[(size_t)warp * psz + wsz] = dbacc;   // every lane accumulated the same db
  __syncthreads();
  for (int i = threadIdx.x; i < psz; i += blockDim.x) {
    float s2 = 0.f;
    for (int wv = 0; wv < nw; ++wv) s2 += sm[(size_t)wv * psz + i];
    ws[(size_t)blockIdx.x * psz + i] = s2;
  }
}

bool last2_supported(int r, int c, int cout) { return cout == 1 && (c == 32 || c == 64 || c == 128) && r >= 2 && r <= 8; }

int fill_geom2(LastGeom2* g, int n, int h, int w, int r, int c, const int32_t* phase_yx_host) {
  g->n = n; g->h = h; g->w = w; g->r = r; g->c = c;
  g->sb = (r % 4 == 0) ? 4 : r;
  if (g->sb > kSBMax) return -1;
  g->nsb = r / g->sb;
  for (int i = 0; i < 64; ++i) g->slot_of[i] = 0;
  for (int s = 0; s < r * r; ++s) {
    const int py = phase_yx_host[2 * s], px = phase_yx_host[2 * s + 1];
    if (py < 0 || py >= r || px < 0 || px >= r) return -1;
    g->slot_of[py * r + px] = s;
  }
  return 0;
}

bool geom_is_zr4(const LastGeom2& g) {
  if (g.r != 4 || g.sb != 4) return false;
  for (int py = 0; py < 4; ++py)
    for (int px = 0; px < 4; ++px)
      if (g.slot_of[py * 4 + px] != zslot4(py, px)) return false;
  return true;
}

int first_bwd_blocks(long pixels) {
  long b = (pixels + 31) / 32;
  const long cap = (long)num_sms() * 4;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

constexpr int kLastBwdThreads = 256;
int last_bwd_blocks() { return num_sms() * 2; }

int fill_geom(LastGeom* g, int n, int h, int w, int r, int c, int cout, const int32_t* phase_yx_host) {
  g->n = n; g->h = h; g->w = w; g->r = r; g->c = c; g->cout = cout;
  for (int i = 0; i < 64; ++i) g->slot_of[i] = 0;
  for (int s = 0; s < r * r; ++s) {
    const int py = phase_yx_host[2 * s], px = phase_yx_host[2 * s + 1];
    if (py < 0 || py >= r || px < 0 || px >= r) return -1;
    g->slot_of[py * r + px] = s;
  }
  return 0;
}

}  // namespace
}  // namespace vsr

namespace vsr {      // firstconv_mma.cu: bf16 output, cin == 1 on the tensor cores (mma.sync, split-bf16 operands)
bool firstconv_mma_supported(int dtype, int cin, int cout);
bool firstconv_mma_bwd_supported(int dtype, int cin, int cout);
void firstconv_mma_fwd(const float* x, int n, int h, int w, const float* wt, const float* bias, const float* slope, void* y,
                       int cout, cudaStream_t s);
int firstconv_mma_bwd(const float* x, int n, int h, int w, const void* dz, int cout, float* ws, int max_blocks, cudaStream_t s);
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_conv3x3_first(const float* x, int32_t n, int32_t cin, int32_t h, int32_t w_, const float* w,
                                 const float* bias, const float* slope, void* y, int32_t dtype, int32_t cout,
                                 void* stream) {
  VSR_CHECK_ARG(x && w && y && n > 0 && h > 0 && w_ > 0, "vsr_conv3x3_first: bad arguments");
  VSR_CHECK_SUPPORTED(cin >= 1 && cin <= kMaxCin, "vsr_conv3x3_first: cin must be in [1,%d]", kMaxCin);
  VSR_CHECK_SUPPORTED(cout % 8 == 0 && cout > 0, "vsr_conv3x3_first: cout must be a multiple of 8");
  const size_t smem = ((size_t)cin * 9 * cout + cout) * sizeof(float);
  VSR_CHECK_SUPPORTED(smem <= 48 * 1024, "vsr_conv3x3_first: cin*cout too large for shared memory");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (firstconv_mma_supported(dtype, cin, cout) && (long)n * h * w_ < 2147000000l) {
    firstconv_mma_fwd(x, n, h, w_, w, bias, slope, y, cout, s);
    VSR_CHECK_LAUNCH("vsr_conv3x3_first(mma)");
    return VSR_OK;
  }
  if (cin == 1 && cout <= 256 && (dtype == VSR_F32 || dtype == VSR_BF16)) {
    const int grid2 = grid_for((long)n * h * w_, 8 * 8, 2);      // >= 8 pixels per warp
    if (dtype == VSR_F32) conv_first2_kernel<float><<<grid2, 256, 0, s>>>(x, n, h, w_, w, bias, slope, (float*)y, cout);
    else conv_first2_kernel<__nv_bfloat16><<<grid2, 256, 0, s>>>(x, n, h, w_, w, bias, slope, (__nv_bfloat16*)y, cout);
    VSR_CHECK_LAUNCH("vsr_conv3x3_first(v2)");
    return VSR_OK;
  }
  const long total = (long)n * h * w_ * (cout / 8);
  const int grid = grid_for(total, 256, 4);
  if (dtype == VSR_F32)
    conv_first_kernel<float><<<grid, 256, smem, s>>>(x, n, cin, h, w_, w, bias, slope, (float*)y, cout);
  else if (dtype == VSR_BF16)
    conv_first_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(x, n, cin, h, w_, w, bias, slope, (__nv_bfloat16*)y, cout);
  else
    VSR_CHECK_ARG(false, "vsr_conv3x3_first: bad dtype %d", dtype);
  VSR_CHECK_LAUNCH("vsr_conv3x3_first");
  return VSR_OK;
}

extern "C" size_t vsr_conv3x3_first_bwd_workspace(int32_t n, int32_t cin, int32_t h, int32_t w_, int32_t cout) {
  const long pixels = (long)n * h * w_;
  size_t v1 = (size_t)first_bwd_blocks(pixels) * cout * (cin * 9 + 1) * sizeof(float);
  size_t v2 = (size_t)((pixels + kF2Pix - 1) / kF2Pix) * cout * 10 * sizeof(float);
  return v1 > v2 ? v1 : v2;
}

extern "C" int vsr_conv3x3_first_bwd(const float* x, int32_t n, int32_t cin, int32_t h, int32_t w_,
                                     const void* dz, int32_t dtype, int32_t cout, float* dw, float* db,
                                     int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(x && dz && dw && db && n > 0 && h > 0 && w_ > 0 && cout > 0, "vsr_conv3x3_first_bwd: bad arguments");
  VSR_CHECK_SUPPORTED(cin >= 1 && cin <= kMaxCin, "vsr_conv3x3_first_bwd: cin must be in [1,%d]", kMaxCin);
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_conv3x3_first_bwd_workspace(n, cin, h, w_, cout),
                "vsr_conv3x3_first_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long pixels = (long)n * h * w_;
  if (firstconv_mma_bwd_supported(dtype, cin, cout) && pixels < 2147000000l) {
    float* wsm = static_cast<float*>(workspace);
    const int blocks = firstconv_mma_bwd(x, n, h, w_, dz, cout, wsm, (int)((pixels + kF2Pix - 1) / kF2Pix), s);
    VSR_CHECK_LAUNCH("vsr_conv3x3_first_bwd(mma)");
    const int nout = cout * 10;
    rows_reduce_kernel<<<(nout + 63) / 64, 256, 0, s>>>(wsm, blocks, nout, dw, db, cout, accumulate);
    VSR_CHECK_LAUNCH("vsr_conv3x3_first_bwd_final(mma)");
    return VSR_OK;
  }
  if (cin == 1 && cout <= 256 && (dtype == VSR_F32 || dtype == VSR_BF16)) {
    int blocks2 = (int)((pixels + kF2Pix - 1) / kF2Pix);
    if (blocks2 > 2 * num_sms()) blocks2 = 2 * num_sms();     // two 1024-thread blocks per SM; every block walks its chunks
    float* ws2 = static_cast<float*>(workspace);
    if (cout <= 64) {
      if (dtype == VSR_F32) conv_first2_bwd_kernel<float, 16><<<blocks2, 1024, 0, s>>>(x, n, h, w_, (const float*)dz, cout, ws2);
      else conv_first2_bwd_kernel<__nv_bfloat16, 16><<<blocks2, 1024, 0, s>>>(x, n, h, w_, (const __nv_bfloat16*)dz, cout, ws2);
    } else {
      if (dtype == VSR_F32) conv_first2_bwd_kernel<float, 4><<<blocks2, 1024, 0, s>>>(x, n, h, w_, (const float*)dz, cout, ws2);
      else conv_first2_bwd_kernel<__nv_bfloat16, 4><<<blocks2, 1024, 0, s>>>(x, n, h, w_, (const __nv_bfloat16*)dz, cout, ws2);
    }
    VSR_CHECK_LAUNCH("vsr_conv3x3_first_bwd(v2)");
    const int nout = cout * 10;
    rows_reduce_kernel<<<(nout + 63) / 64, 256, 0, s>>>(ws2, blocks2, nout, dw, db, cout, accumulate);
    VSR_CHECK_LAUNCH("vsr_conv3x3_first_bwd_final(v2)");
    return VSR_OK;
  }
  const int blocks = first_bwd_blocks(pixels);
  const long ppb = (pixels + blocks - 1) / blocks;
  const int threads = cout >= 256 ? 256 : ((cout + 31) / 32) * 32;
  float* ws = static_cast<float*>(workspace);
  if (dtype == VSR_F32)
    conv_first_bwd_kernel<float><<<blocks, threads, 0, s>>>(x, n, cin, h, w_, (const float*)dz, cout, ppb, ws);
  else if (dtype == VSR_BF16)
    conv_first_bwd_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(x, n, cin, h, w_, (const __nv_bfloat16*)dz, cout, ppb, ws);
  else
    VSR_CHECK_ARG(false, "vsr_conv3x3_first_bwd: bad dtype %d", dtype);
  VSR_CHECK_LAUNCH("vsr_conv3x3_first_bwd");
  const int tot = cout * (cin * 9 + 1);
  conv_first_bwd_final_kernel<<<(tot + 127) / 128, 128, 0, s>>>(ws, blocks, cout, cin, dw, db, accumulate);
  VSR_CHECK_LAUNCH("vsr_conv3x3_first_bwd_final");
  return VSR_OK;
}

namespace vsr {      // lastconv_mma.cu: bf16, cout == 1 on the tensor cores (mma.sync)
bool lastconv_mma_supported(int dtype, int r, int c, int cout);
int lastconv_mma_blocks();
int lastconv_mma_fwd(const void* x, int n, int h, int w, int r, int c, const int32_t* phase_yx, const float* wt,
                     const float* bias, float* y, cudaStream_t s);
int lastconv_mma_bwd(const void* x, int n, int h, int w, int r, int c, const int32_t* phase_yx, const float* wt,
                     const float* dy, void* dx, float* ws, cudaStream_t s);
}  // namespace vsr

extern "C" int vsr_conv3x3_last(const void* x, int32_t dtype, int32_t n, int32_t h, int32_t w_, int32_t r,
                                int32_t c, const int32_t* phase_yx, const float* w, const float* bias,
                                float* y, int32_t cout, void* stream) {
  VSR_CHECK_ARG(x && w && y && phase_yx && n > 0 && h > 0 && w_ > 0, "vsr_conv3x3_last: bad arguments");
  VSR_CHECK_SUPPORTED(r >= 1 && r <= 8, "vsr_conv3x3_last: r must be in [1,8]");
  VSR_CHECK_SUPPORTED(cout >= 1 && cout <= kMaxCoutLast, "vsr_conv3x3_last: cout must be in [1,%d]", kMaxCoutLast);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (lastconv_mma_supported(dtype, r, c, cout)) return lastconv_mma_fwd(x, n, h, w_, r, c, phase_yx, w, bias, y, s);
  if (last2_supported(r, c, cout) && (dtype == VSR_F32 || dtype == VSR_BF16)) {
    LastGeom2 g2;
    VSR_CHECK_ARG(fill_geom2(&g2, n, h, w_, r, c, phase_yx) == 0, "vsr_conv3x3_last: bad phase table");
    const long items = (long)n * h * w_ * g2.nsb * g2.nsb;
    const int grid2 = grid_for(items, 8, 8);
    const size_t smem2 = (size_t)9 * c * sizeof(float);
    if (dtype == VSR_BF16) {
      // (c == 64 / 128 go to lastconv_mma.cu above; this branch serves c == 32)
      if (c % 64 == 0) conv_last2_kernel<__nv_bfloat16, 2><<<grid2, 256, smem2, s>>>((const __nv_bfloat16*)x, g2, w, bias, y);
      else conv_last2_kernel<__nv_bfloat16, 1><<<grid2, 256, smem2, s>>>((const __nv_bfloat16*)x, g2, w, bias, y);
    } else {
      if (c % 64 == 0) conv_last2_kernel<float, 2><<<grid2, 256, smem2, s>>>((const float*)x, g2, w, bias, y);
      else conv_last2_kernel<float, 1><<<grid2, 256, smem2, s>>>((const float*)x, g2, w, bias, y);
    }
    VSR_CHECK_LAUNCH("vsr_conv3x3_last(v2)");
    return VSR_OK;
  }
  LastGeom g;
  VSR_CHECK_ARG(fill_geom(&g, n, h, w_, r, c, cout, phase_yx) == 0, "vsr_conv3x3_last: bad phase table");
  const size_t smem = (size_t)cout * 9 * c * sizeof(float);
  VSR_CHECK_SUPPORTED(smem <= 48 * 1024, "vsr_conv3x3_last: cout*c too large");
  const long total = (long)n * h * r * w_ * r;
  const int grid = grid_for(total, 8, 8);
  if (dtype == VSR_F32)
    conv_last_kernel<float><<<grid, 256, smem, s>>>((const float*)x, g, w, bias, y);
  else if (dtype == VSR_BF16)
    conv_last_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>((const __nv_bfloat16*)x, g, w, bias, y);
  else
    VSR_CHECK_ARG(false, "vsr_conv3x3_last: bad dtype %d", dtype);
  VSR_CHECK_LAUNCH("vsr_conv3x3_last");
  return VSR_OK;
}

extern "C" size_t vsr_conv3x3_last_bwd_workspace(int32_t n, int32_t h, int32_t w_, int32_t r, int32_t c,
                                                 int32_t cout) {
  (void)n; (void)h; (void)w_; (void)r;
  return (size_t)last_bwd_blocks() * ((size_t)cout * 9 * c + cout) * sizeof(float);
}

extern "C" int vsr_conv3x3_last_bwd(const void* x, int32_t dtype, int32_t n, int32_t h, int32_t w_, int32_t r,
                                    int32_t c, const int32_t* phase_yx, const float* w, const float* dy,
                                    int32_t cout, void* dx, float* dw, float* db, int accumulate,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(x && w && dy && dx && dw && db && phase_yx, "vsr_conv3x3_last_bwd: bad arguments");
  VSR_CHECK_SUPPORTED(r >= 1 && r <= 8, "vsr_conv3x3_last_bwd: r must be in [1,8]");
  if (lastconv_mma_supported(dtype, r, c, cout)) {
    VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_conv3x3_last_bwd_workspace(n, h, w_, r, c, cout),
                  "vsr_conv3x3_last_bwd: workspace too small");
    cudaStream_t sm = static_cast<cudaStream_t>(stream);
    float* wsm = static_cast<float*>(workspace);
    int rc = lastconv_mma_bwd(x, n, h, w_, r, c, phase_yx, w, dy, dx, wsm, sm);
    if (rc != VSR_OK) return rc;
    const int pszm = 9 * c + 1;
    conv_last_bwd_final_kernel<<<(pszm + 127) / 128, 128, 0, sm>>>(wsm, lastconv_mma_blocks(), 1, c, dw, db, accumulate);
    VSR_CHECK_LAUNCH("vsr_conv3x3_last_bwd_final(mma)");
    return VSR_OK;
  }
  if (last2_supported(r, c, cout) && (dtype == VSR_F32 || dtype == VSR_BF16)) {
    VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_conv3x3_last_bwd_workspace(n, h, w_, r, c, cout),
                  "vsr_conv3x3_last_bwd: workspace too small");
    LastGeom2 g2;
    VSR_CHECK_ARG(fill_geom2(&g2, n, h, w_, r, c, phase_yx) == 0, "vsr_conv3x3_last_bwd: bad phase table");
    cudaStream_t s2 = static_cast<cudaStream_t>(stream);
    const int blocks2 = last_bwd_blocks();
    const int nw2 = kLastBwdThreads / 32;
    const size_t psz2 = (size_t)9 * c + 1;
    const size_t smem_w = (size_t)9 * c * sizeof(float);
    const size_t smem_p = (size_t)nw2 * psz2 * sizeof(float);
    float* ws2 = static_cast<float*>(workspace);
    const long items = (long)n * h * w_ * g2.nsb * g2.nsb;
    const int grid_dx = grid_for(items, 8, 8);
    if (dtype == VSR_BF16) {
      using B = __nv_bfloat16;
      if (c % 64 == 0 && geom_is_zr4(g2)) {
        conv_last2_dx_kernel<B, 2, true><<<grid_dx, 256, smem_w, s2>>>(g2, w, dy, (B*)dx);
        conv_last2_dw_kernel<B, 2, true><<<blocks2, kLastBwdThreads, smem_p, s2>>>((const B*)x, g2, dy, ws2);
      } else if (c % 64 == 0) {
        conv_last2_dx_kernel<B, 2><<<grid_dx, 256, smem_w, s2>>>(g2, w, dy, (B*)dx);
        conv_last2_dw_kernel<B, 2><<<blocks2, kLastBwdThreads, smem_p, s2>>>((const B*)x, g2, dy, ws2);
      } else {
        conv_last2_dx_kernel<B, 1><<<grid_dx, 256, smem_w, s2>>>(g2, w, dy, (B*)dx);
        conv_last2_dw_kernel<B, 1><<<blocks2, kLastBwdThreads, smem_p, s2>>>((const B*)x, g2, dy, ws2);
      }
    } else {
      if (c % 64 == 0) {
        conv_last2_dx_kernel<float, 2><<<grid_dx, 256, smem_w, s2>>>(g2, w, dy, (float*)dx);
        conv_last2_dw_kernel<float, 2><<<blocks2, kLastBwdThreads, smem_p, s2>>>((const float*)x, g2, dy, ws2);
      } else {
        conv_last2_dx_kernel<float, 1><<<grid_dx, 256, smem_w, s2>>>(g2, w, dy, (float*)dx);
        conv_last2_dw_kernel<float, 1><<<blocks2, kLastBwdThreads, smem_p, s2>>>((const float*)x, g2, dy, ws2);
      }
    }
    VSR_CHECK_LAUNCH("vsr_conv3x3_last_bwd(v2)");
    conv_last_bwd_final_kernel<<<((int)psz2 + 127) / 128, 128, 0, s2>>>(ws2, blocks2, 1, c, dw, db, accumulate);
    VSR_CHECK_LAUNCH("vsr_conv3x3_last_bwd_final");
    return VSR_OK;
  }
  const int cpl = (c + 31) / 32;
  VSR_CHECK_SUPPORTED(cout >= 1 && cout * cpl <= 4, "vsr_conv3x3_last_bwd: cout*ceil(c/32) must be <= 4 (got %d*%d)", cout, cpl);
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_conv3x3_last_bwd_workspace(n, h, w_, r, c, cout),
                "vsr_conv3x3_last_bwd: workspace too small");
  LastGeom g;
  VSR_CHECK_ARG(fill_geom(&g, n, h, w_, r, c, cout, phase_yx) == 0, "vsr_conv3x3_last_bwd: bad phase table");
  const int nw = kLastBwdThreads / 32;
  const size_t psz = (size_t)cout * 9 * c + cout;
  const size_t smem = ((size_t)cout * 9 * c + nw * psz) * sizeof(float);
  VSR_CHECK_SUPPORTED(smem <= 48 * 1024, "vsr_conv3x3_last_bwd: shared memory need too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = last_bwd_blocks();
  float* ws = static_cast<float*>(workspace);
  if (dtype == VSR_F32)
    conv_last_bwd_kernel<float><<<blocks, kLastBwdThreads, smem, s>>>((const float*)x, g, w, dy, (float*)dx, ws, cpl);
  else if (dtype == VSR_BF16)
    conv_last_bwd_kernel<__nv_bfloat16><<<blocks, kLastBwdThreads, smem, s>>>((const __nv_bfloat16*)x, g, w, dy, (__nv_bfloat16*)dx, ws, cpl);
  else
    VSR_CHECK_ARG(false, "vsr_conv3x3_last_bwd: bad dtype %d", dtype);
  VSR_CHECK_LAUNCH("vsr_conv3x3_last_bwd");
  conv_last_bwd_final_kernel<<<((int)psz + 127) / 128, 128, 0, s>>>(ws, blocks, cout, c, dw, db, accumulate);
  VSR_CHECK_LAUNCH("vsr_conv3x3_last_bwd_final");
  return VSR_OK;
}
