// lastconv_mma.cu — the last 3x3 convolution of the output block (F -> 1 channel, drf_net.py:144,147; srfb_net.py:150)
// on bf16 phase-blocked feature maps, forward / data gradient / weight gradient.
//
// The layer moves 128 bytes per high-resolution pixel for 576 MACs: it is bound by HBM, but the CUDA-core kernels it
// replaces needed ~28 warp instructions per pixel (bf16 unpack + FMA per channel pair + a 16-value shuffle reduction) and
// re-read every pixel 2.25 times as a halo; they ran at 0.20-0.27 of the copy bandwidth.  Here the channel contraction
// runs on the tensor cores with warp-level mma.sync.m16n8k16 (no shared-memory operand staging is worth it for a K = 64,
// N = 9 problem):
//   forward : z[p][tap] = sum_c x[p][c] * w[c][tap]   (M = 16 pixels, N = 9 taps in two n-tiles, K = C) for every pixel
//             of a 32x32 tile plus its 1-pixel ring, each pixel read ONCE; then y[P] = b + sum_tap z[P + d(tap)][tap]
//             from shared memory (9 loads per output);
//   dx      : dx[p][c] = sum_tap dy[p - d(tap)] * w[c][tap]   (M = 16 pixels, K = 9 taps padded to 16, N = C);
//   dw      : dw[c][tap] = sum_p dy[p - d(tap)] * x[p][c]     (M = taps, N = 8 channels, K = 16 pixels), the x fragments
//             transposed in registers with movmatrix.
// A-operand rows are arbitrary pixels (every lane computes its own pixel address), so tiles are cut on the
// high-resolution grid whatever the phase-blocked layout underneath; a lane loads 16 contiguous bytes of a pixel's
// channel vector and the channel <-> k-index permutation this implies is applied to the weight fragments too.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace vsr {
namespace {

constexpr int kTH = 32, kTW = 32;                  // output tile (high-resolution pixels)
constexpr int kHH = kTH + 2, kHW = kTW + 2;        // with its 1-pixel ring
constexpr int kHalo = kHH * kHW;                   // 1156
constexpr int kThreads = 256;

struct LcGeom {
  int n, h, w, r, c;
  int H, W;                 // high-resolution size
  int tiles_x, tiles_y;
  int slot_of[64];          // (py * r + px) -> phase slot
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// element offset of the channel vector of high-resolution pixel (Y, X) of image ni; slot table in shared memory
__device__ __forceinline__ size_t px_off(const LcGeom& g, const int* slot_s, int ni, int Y, int X) {
  const int ly = Y / g.r, lx = X / g.r;
  const int slot = slot_s[(Y - ly * g.r) * g.r + (X - lx * g.r)];
  return (((size_t)ni * g.h + ly) * g.w + lx) * ((size_t)g.r * g.r * g.c) + (size_t)slot * g.c;
}

struct Tile {
  int ni, Y0, X0;
};
__device__ __forceinline__ Tile decode(const LcGeom& g, long t) {
  Tile r;
  const int tx = (int)(t % g.tiles_x);
  t /= g.tiles_x;
  const int ty = (int)(t % g.tiles_y);
  r.ni = (int)(t / g.tiles_y);
  r.Y0 = ty * kTH;
  r.X0 = tx * kTW;
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// forward.  NH = C / 32 (a lane loads 16 bytes = 8 channels of each 32-channel half of a pixel).
// ------------------------------------------------------------------------------------------------------------------
template <int NH>
__global__ void __launch_bounds__(kThreads, 4) lastconv_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const __grid_constant__ LcGeom g,
                                                                const float* __restrict__ wt, const float* __restrict__ bias,
                                                                float* __restrict__ y) {
  __shared__ float z[kHalo * 9];
  __shared__ int slot_s[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t4 = lane & 3;
  if (tid < 64) slot_s[tid] = g.slot_of[tid];
  // weight fragments: n-tile 0 holds taps 0..7 (n = gq), n-tile 1 tap 8 in column 0.  k-step s of half hf covers, for
  // this lane, channels hf*32 + 8*t4 + 4*s + {0,1} (k = 2t, 2t+1) and + {2,3} (k = 2t+8, 2t+9): the order in which the
  // lane's 16-byte pixel load delivers them.
  uint32_t bw[2][NH][2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int hf = 0; hf < NH; ++hf)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int tap = nt * 8 + gq;
        const int c0 = hf * 32 + 8 * t4 + 4 * s;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
        if (tap < 9) {
          w0 = __ldg(wt + (size_t)(c0 + 0) * 9 + tap); w1 = __ldg(wt + (size_t)(c0 + 1) * 9 + tap);
          w2 = __ldg(wt + (size_t)(c0 + 2) * 9 + tap); w3 = __ldg(wt + (size_t)(c0 + 3) * 9 + tap);
        }
        bw[nt][hf][s][0] = pack_bf16x2(w0, w1);
        bw[nt][hf][s][1] = pack_bf16x2(w2, w3);
      }
  const float b0 = bias ? __ldg(bias) : 0.f;
  const long tiles = (long)g.n * g.tiles_y * g.tiles_x;
  constexpr int kGroups = (kHalo + 15) / 16;
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const Tile tc = decode(g, tile);
    __syncthreads();                       // slot table staged / previous tile's z consumed
    // two 16-pixel groups per trip: the loads of both are issued before the first MMA (bytes in flight per warp)
    for (int q0 = warp; q0 < kGroups; q0 += 2 * (kThreads / 32)) {
      uint4 v[2][2][NH];
      int p[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u * (kThreads / 32);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          p[u][rr] = q < kGroups ? q * 16 + gq + 8 * rr : kHalo;
          const int hy = p[u][rr] / kHW, hx = p[u][rr] - hy * kHW;
          const int Y = tc.Y0 - 1 + hy, X = tc.X0 - 1 + hx;
          const bool ok = p[u][rr] < kHalo && Y >= 0 && Y < g.H && X >= 0 && X < g.W;
          const __nv_bfloat16* src = x + (ok ? px_off(g, slot_s, tc.ni, Y, X) : 0) + 8 * t4;
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) {
            const uint4 ld = __ldg(reinterpret_cast<const uint4*>(src + hf * 32));
            v[u][rr][hf] = ok ? ld : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
          const uint32_t a0[4] = {v[u][0][hf].x, v[u][1][hf].x, v[u][0][hf].y, v[u][1][hf].y};
          const uint32_t a1[4] = {v[u][0][hf].z, v[u][1][hf].z, v[u][0][hf].w, v[u][1][hf].w};
          mma16816(d0, a0, bw[0][hf][0]);
          mma16816(d0, a1, bw[0][hf][1]);
          mma16816(d1, a0, bw[1][hf][0]);
          mma16816(d1, a1, bw[1][hf][1]);
        }
        if (p[u][0] < kHalo) {
          z[p[u][0] * 9 + 2 * t4] = d0[0];
          z[p[u][0] * 9 + 2 * t4 + 1] = d0[1];
          if (t4 == 0) z[p[u][0] * 9 + 8] = d1[0];
        }
        if (p[u][1] < kHalo) {
          z[p[u][1] * 9 + 2 * t4] = d0[2];
          z[p[u][1] * 9 + 2 * t4 + 1] = d0[3];
          if (t4 == 0) z[p[u][1] * 9 + 8] = d1[2];
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < kTH * kTW; i += kThreads) {
      const int oy = i / kTW, ox = i - oy * kTW;
      const int Y = tc.Y0 + oy, X = tc.X0 + ox;
      if (Y < g.H && X < g.W) {
        float s = b0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) s += z[((oy + ky) * kHW + ox + kx) * 9 + ky * 3 + kx];
        y[((size_t)tc.ni * g.H + Y) * g.W + X] = s;
      }
    }
  }
}

// stage the (zero-padded) 34x34 window of dy around the tile in shared memory
__device__ __forceinline__ void stage_dy(const float* __restrict__ dy, const LcGeom& g, const Tile& tc, float* dys) {
  for (int i = threadIdx.x; i < kHalo; i += kThreads) {
    const int hy = i / kHW, hx = i - hy * kHW;
    const int Y = tc.Y0 - 1 + hy, X = tc.X0 - 1 + hx;
    dys[i] = (Y >= 0 && Y < g.H && X >= 0 && X < g.W) ? __ldg(dy + ((size_t)tc.ni * g.H + Y) * g.W + X) : 0.f;
  }
}
// dy[P - d(tap)] for output position (oy, ox) of the tile: the tap's neighbour in the staged window
__device__ __forceinline__ float dy_tap(const float* dys, int oy, int ox, int tap) {
  const int ky = tap / 3, kx = tap - ky * 3;
  return dys[(oy - ky + 2) * kHW + ox - kx + 2];
}

// ------------------------------------------------------------------------------------------------------------------
// data gradient.  A: [16 pixels x 16 (9 taps, zero-padded)] of dy neighbours (bf16), B: w[c][tap], one k-step; n-tile j,
// column n stands for channel 16 * (n / 2) + 2 * j + (n & 1) of a 64-channel block, so that a lane ends up with 16
// consecutive channels of its pixel (two 16-byte stores).  NB = C / 64.
// ------------------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(kThreads) lastconv_dx_kernel(const __grid_constant__ LcGeom g, const float* __restrict__ wt,
                                                               const float* __restrict__ dy, __nv_bfloat16* __restrict__ dx) {
  __shared__ float dys[kHalo];
  __shared__ int slot_s[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t4 = lane & 3;
  if (tid < 64) slot_s[tid] = g.slot_of[tid];
  uint32_t bw[NB][8][2];
#pragma unroll
  for (int cb = 0; cb < NB; ++cb)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = cb * 64 + 16 * (gq >> 1) + 2 * j + (gq & 1);
      bw[cb][j][0] = pack_bf16x2(__ldg(wt + (size_t)ch * 9 + 2 * t4), __ldg(wt + (size_t)ch * 9 + 2 * t4 + 1));
      bw[cb][j][1] = t4 == 0 ? pack_bf16x2(__ldg(wt + (size_t)ch * 9 + 8), 0.f) : 0u;
    }
  const long tiles = (long)g.n * g.tiles_y * g.tiles_x;
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const Tile tc = decode(g, tile);
    __syncthreads();
    stage_dy(dy, g, tc, dys);
    __syncthreads();
    // 16-pixel groups: half a row of the tile each
    for (int q = warp; q < kTH * kTW / 16; q += kThreads / 32) {
      const int oy = q / (kTW / 16), ox0 = (q - oy * (kTW / 16)) * 16;
      uint32_t a[4];
      a[0] = pack_bf16x2(dy_tap(dys, oy, ox0 + gq, 2 * t4), dy_tap(dys, oy, ox0 + gq, 2 * t4 + 1));
      a[1] = pack_bf16x2(dy_tap(dys, oy, ox0 + gq + 8, 2 * t4), dy_tap(dys, oy, ox0 + gq + 8, 2 * t4 + 1));
      a[2] = t4 == 0 ? pack_bf16x2(dy_tap(dys, oy, ox0 + gq, 8), 0.f) : 0u;
      a[3] = t4 == 0 ? pack_bf16x2(dy_tap(dys, oy, ox0 + gq + 8, 8), 0.f) : 0u;
      const int Y = tc.Y0 + oy;
#pragma unroll
      for (int cb = 0; cb < NB; ++cb) {
        uint32_t o[2][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(d, a, bw[cb][j]);
          o[0][j] = pack_bf16x2(d[0], d[1]);
          o[1][j] = pack_bf16x2(d[2], d[3]);
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int X = tc.X0 + ox0 + gq + 8 * rr;
          if (Y < g.H && X < g.W) {
            uint4* dst = reinterpret_cast<uint4*>(dx + px_off(g, slot_s, tc.ni, Y, X) + cb * 64 + 16 * t4);
            dst[0] = make_uint4(o[rr][0], o[rr][1], o[rr][2], o[rr][3]);
            dst[1] = make_uint4(o[rr][4], o[rr][5], o[rr][6], o[rr][7]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight / bias gradient: D[tap (M = 16), channel (N = 8)] += dyN[tap][16 pixels] * x[16 pixels][8 channels].
// A lane's 16-byte load of a pixel covers channels hf*32 + 8*t4 + {0..7}; register i of it, transposed across the warp
// with movmatrix, is the B fragment of the 8 channels {hf*32 + 8*(n/2) + 2*i + (n&1)}, so accumulator (hf, i) column
// pair (2*t4, 2*t4+1) is channels hf*32 + 8*t4 + 2*i + {0, 1}.  Per-block partial sums -> ws[block][9*C + 1]
// (tap-major, then db), reduced in a fixed order by conv_last_bwd_final_kernel (layers.cu).
// ------------------------------------------------------------------------------------------------------------------
template <int NH>
__global__ void __launch_bounds__(kThreads) lastconv_dw_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const __grid_constant__ LcGeom g,
                                                               const float* __restrict__ dy, float* __restrict__ ws) {
  __shared__ float dys[kHalo];
  __shared__ int slot_s[64];
  extern __shared__ float red[];            // [warps][9*C + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t4 = lane & 3;
  if (tid < 64) slot_s[tid] = g.slot_of[tid];
  float acc[NH][4][4];
#pragma unroll
  for (int hf = 0; hf < NH; ++hf)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[hf][i][e] = 0.f;
  float dbacc = 0.f;
  const long tiles = (long)g.n * g.tiles_y * g.tiles_x;
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const Tile tc = decode(g, tile);
    __syncthreads();
    stage_dy(dy, g, tc, dys);
    __syncthreads();
    // two 16-pixel groups per trip (64 groups per tile, 8 warps): all four pixel rows are requested before the first MMA
    for (int q0 = warp; q0 < kTH * kTW / 16; q0 += 2 * (kThreads / 32)) {
      uint4 v[2][2][NH];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u * (kThreads / 32);
        const int oy = q / (kTW / 16), ox0 = (q - oy * (kTW / 16)) * 16;
        const int Y = tc.Y0 + oy;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int X = tc.X0 + ox0 + gq + 8 * rr;
          const bool ok = Y < g.H && X < g.W;
          const __nv_bfloat16* src = x + (ok ? px_off(g, slot_s, tc.ni, Y, X) : 0) + 8 * t4;
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) {
            const uint4 ld = __ldg(reinterpret_cast<const uint4*>(src + hf * 32));
            v[u][rr][hf] = ok ? ld : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u * (kThreads / 32);
        const int oy = q / (kTW / 16), ox0 = (q - oy * (kTW / 16)) * 16;
        const int Y = tc.Y0 + oy;
        // A: rows = taps (gq, gq + 8), k = the group's 16 pixels; out-of-image pixels carry x = 0, so their dy columns
        // may hold anything the window holds (zero outside the image, real values inside)
        uint32_t a[4];
        a[0] = pack_bf16x2(dy_tap(dys, oy, ox0 + 2 * t4, gq), dy_tap(dys, oy, ox0 + 2 * t4 + 1, gq));
        a[2] = pack_bf16x2(dy_tap(dys, oy, ox0 + 2 * t4 + 8, gq), dy_tap(dys, oy, ox0 + 2 * t4 + 9, gq));
        a[1] = gq == 0 ? pack_bf16x2(dy_tap(dys, oy, ox0 + 2 * t4, 8), dy_tap(dys, oy, ox0 + 2 * t4 + 1, 8)) : 0u;
        a[3] = gq == 0 ? pack_bf16x2(dy_tap(dys, oy, ox0 + 2 * t4 + 8, 8), dy_tap(dys, oy, ox0 + 2 * t4 + 9, 8)) : 0u;
        if (lane < 16) {                      // db: every pixel of the group once
          const int X = tc.X0 + ox0 + lane;
          dbacc += (Y < g.H && X < g.W) ? dys[(oy + 1) * kHW + ox0 + lane + 1] : 0.f;
        }
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
          const uint32_t r0[4] = {v[u][0][hf].x, v[u][0][hf].y, v[u][0][hf].z, v[u][0][hf].w};
          const uint32_t r1[4] = {v[u][1][hf].x, v[u][1][hf].y, v[u][1][hf].z, v[u][1][hf].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t b[2] = {movmatrix_trans(r0[i]), movmatrix_trans(r1[i])};   // pixels 0-7, 8-15 of the group
            mma16816(acc[hf][i], a, b);
          }
        }
      }
    }
  }
  // fold: per-warp partials in shared memory, then fixed-order sum over the warps of the block
  const int C = g.c, psz = 9 * C + 1, nw = kThreads / 32;
  __syncthreads();
#pragma unroll
  for (int hf = 0; hf < NH; ++hf)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ch = hf * 32 + 8 * t4 + 2 * i;
      red[(size_t)warp * psz + gq * C + ch] = acc[hf][i][0];           // tap gq (0..7)
      red[(size_t)warp * psz + gq * C + ch + 1] = acc[hf][i][1];
      if (gq == 0) {                                                    // tap 8 (row 8 of the accumulator)
        red[(size_t)warp * psz + 8 * C + ch] = acc[hf][i][2];
        red[(size_t)warp * psz + 8 * C + ch + 1] = acc[hf][i][3];
      }
    }
  dbacc = warp_sum(dbacc);
  if (lane == 0) red[(size_t)warp * psz + 9 * C] = dbacc;
  __syncthreads();
  for (int i = tid; i < psz; i += kThreads) {
    float s = 0.f;
    for (int wv = 0; wv < nw; ++wv) s += red[(size_t)wv * psz + i];
    ws[(size_t)blockIdx.x * psz + i] = s;
  }
}

int fill(LcGeom* g, int n, int h, int w, int r, int c, const int32_t* phase_yx) {
  g->n = n; g->h = h; g->w = w; g->r = r; g->c = c;
  g->H = h * r; g->W = w * r;
  g->tiles_x = (g->W + kTW - 1) / kTW;
  g->tiles_y = (g->H + kTH - 1) / kTH;
  for (int i = 0; i < 64; ++i) g->slot_of[i] = 0;
  for (int s = 0; s < r * r; ++s) {
    const int py = phase_yx[2 * s], px = phase_yx[2 * s + 1];
    if (py < 0 || py >= r || px < 0 || px >= r) return -1;
    g->slot_of[py * r + px] = s;
  }
  return 0;
}

}  // namespace

bool lastconv_mma_supported(int dtype, int r, int c, int cout) {
  return dtype == VSR_BF16 && cout == 1 && (c == 64 || c == 128) && r >= 1 && r <= 8;
}
int lastconv_mma_blocks() { return num_sms() * 2; }

int lastconv_mma_fwd(const void* x, int n, int h, int w, int r, int c, const int32_t* phase_yx, const float* wt,
                     const float* bias, float* y, cudaStream_t s) {
  LcGeom g;
  VSR_CHECK_ARG(fill(&g, n, h, w, r, c, phase_yx) == 0, "vsr_conv3x3_last: bad phase table");
  const long tiles = (long)n * g.tiles_x * g.tiles_y;
  const int grid = (int)std::min<long>(tiles, (long)num_sms() * 4);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  if (c == 64) lastconv_fwd_kernel<2><<<grid, kThreads, 0, s>>>(xb, g, wt, bias, y);
  else lastconv_fwd_kernel<4><<<grid, kThreads, 0, s>>>(xb, g, wt, bias, y);
  VSR_CHECK_LAUNCH("vsr_conv3x3_last(mma)");
  return VSR_OK;
}

// dx and the per-block partials of dw / db (ws: lastconv_mma_blocks() x (9*c + 1) floats); the caller reduces ws
int lastconv_mma_bwd(const void* x, int n, int h, int w, int r, int c, const int32_t* phase_yx, const float* wt,
                     const float* dy, void* dx, float* ws, cudaStream_t s) {
  LcGeom g;
  VSR_CHECK_ARG(fill(&g, n, h, w, r, c, phase_yx) == 0, "vsr_conv3x3_last_bwd: bad phase table");
  const long tiles = (long)n * g.tiles_x * g.tiles_y;
  const int grid = (int)std::min<long>(tiles, (long)num_sms() * 4);
  const int blocks = lastconv_mma_blocks();
  const size_t smem = (size_t)(kThreads / 32) * (9 * c + 1) * sizeof(float);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dxb = static_cast<__nv_bfloat16*>(dx);
  if (c == 64) {
    lastconv_dx_kernel<1><<<grid, kThreads, 0, s>>>(g, wt, dy, dxb);
    lastconv_dw_kernel<2><<<blocks, kThreads, smem, s>>>(xb, g, dy, ws);
  } else {
    lastconv_dx_kernel<2><<<grid, kThreads, 0, s>>>(g, wt, dy, dxb);
    lastconv_dw_kernel<4><<<blocks, kThreads, smem, s>>>(xb, g, dy, ws);
  }
  VSR_CHECK_LAUNCH("vsr_conv3x3_last_bwd(mma)");
  return VSR_OK;
}

}  // namespace vsr
