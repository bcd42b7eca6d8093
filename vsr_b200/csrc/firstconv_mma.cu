// firstconv_mma.cu — the first 3x3 convolution of the input block (1 -> 4F channels + PReLU, drf_net.py:58-60) and its
// weight / bias gradient, bf16 feature maps, on the tensor cores (warp-level mma.sync.m16n8k16).
//
// The layer writes (reads, backward) 512 bytes per pixel for 2304 MACs: on the FP32 pipes that is as long as the HBM time
// itself and the CUDA-core kernels ran at 0.15 of the copy bandwidth.  K = 9 taps is tiny, so the contraction is done at
// near-fp32 accuracy with split operands: x = xh + xl, w = wh + wl (bf16 pairs; products of bf16 values are exact in the
// fp32 accumulator, the dropped xl * wl term is 2^-16 relative), i.e.
//   forward : y[p][co] = sum_k A[p][k] * B[k][co],  A = [xh(9) | xl(9) | xh(9) | 0(5)],  B = [wh | wh | wl | 0]   (K = 32)
//   backward: dw[co][t] = sum_p dz[p][co] * V[p][t],  V = [xh(9) | 1 | xl(9) | 0(5)]   (N = 24: the "1" column is db;
//             dz is bf16 already, so two products give the fp32-exact sum)
// Forward: a CTA stages A for 256 pixels in shared memory (ldmatrix-ready rows); a warp owns a 32-channel block, keeps
// its B fragments and biases in registers and walks the 16-pixel m-tiles: 2 ldmatrix + 8 mma + two 16-byte stores per
// thread (columns of the n-tiles are permuted so that a thread ends up with 8 consecutive channels of a pixel).
// Backward: dz chunks of 32 pixels arrive by cp.async (double-buffered), A fragments by ldmatrix.trans, accumulators
// (2 m-tiles x 3 n-tiles per warp) stay in registers over all pixels of the CTA; fixed-order partials per CTA.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace vsr {
namespace {

constexpr int kFcThreads = 256;
constexpr int kFcPix = 256;        // forward: pixels per CTA step
constexpr int kFcRow = 40;         // bf16 elements per staged A row (32 used; 80-byte pitch: conflict-free ldmatrix)
constexpr int kBwPix = 32;         // backward: pixels per chunk (two k-steps)
constexpr int kBwCols = 24;        // V columns (three n-tiles)
constexpr int kBwVRow = 40;        // bf16 elements per V column (32 used)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;                     // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// x = hi + lo with hi, lo in bf16
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// the nine zero-padded input samples around pixel p of the flattened [n][h][w] map (p < 2^31: checked on the host)
__device__ __forceinline__ void taps_of(const float* __restrict__ x, unsigned p, unsigned total, int h, int w, float (&v)[9]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) v[t] = 0.f;
  if (p >= total) return;
  const unsigned q = p / (unsigned)w;
  const int px = (int)(p - q * (unsigned)w);
  const unsigned img = q / (unsigned)h;
  const int py = (int)(q - img * (unsigned)h);
  const float* xp = x + (size_t)img * h * w;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = py + ky - 1, xx = px + kx - 1;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) v[ky * 3 + kx] = __ldg(xp + yy * w + xx);
    }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFcThreads) firstconv_fwd_kernel(const float* __restrict__ x, int n, int h, int w,
                                                                   const float* __restrict__ wt, const float* __restrict__ bias,
                                                                   const float* __restrict__ slope_p,
                                                                   __nv_bfloat16* __restrict__ y, int cout) {
  __shared__ __align__(16) __nv_bfloat16 As[kFcPix * kFcRow];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int nblk = cout >> 5;                        // 32-channel blocks: 1, 2, 4 or 8
  const int blk = warp % nblk, mgrp = warp / nblk, mstride = 8 / nblk;
  // B = [wh(9) | wh(9) | wl(9) | 0(5)] per channel, split once per CTA into the (not yet used) A buffer: Bs[co][32]
  for (int co = threadIdx.x; co < cout; co += kFcThreads) {
    __nv_bfloat16 wh[9], wl[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) split_bf16(__ldg(wt + co * 9 + t), wh[t], wl[t]);
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    uint4* row = reinterpret_cast<uint4*>(As + co * kFcRow);
    row[0] = make_uint4(pack2(wh[0], wh[1]), pack2(wh[2], wh[3]), pack2(wh[4], wh[5]), pack2(wh[6], wh[7]));
    row[1] = make_uint4(pack2(wh[8], wh[0]), pack2(wh[1], wh[2]), pack2(wh[3], wh[4]), pack2(wh[5], wh[6]));
    row[2] = make_uint4(pack2(wh[7], wh[8]), pack2(wl[0], wl[1]), pack2(wl[2], wl[3]), pack2(wl[4], wl[5]));
    row[3] = make_uint4(pack2(wl[6], wl[7]), pack2(wl[8], z), pack2(z, z), pack2(z, z));
  }
  __syncthreads();
  // fragments of this warp's channel block: n-tile j, column n <-> channel 32 blk + 8 (n / 2) + 2 j + (n % 2)
  uint32_t bf[4][2][2];
  float br[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co = 32 * blk + 8 * (g >> 1) + 2 * j + (g & 1);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      bf[j][s][0] = *reinterpret_cast<const uint32_t*>(As + co * kFcRow + 16 * s + 2 * q);
      bf[j][s][1] = *reinterpret_cast<const uint32_t*>(As + co * kFcRow + 16 * s + 2 * q + 8);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) br[i] = bias ? __ldg(bias + 32 * blk + 8 * q + i) : 0.f;
  const Prelu a = make_prelu(slope_p ? __ldg(slope_p) : 1.f);
  const unsigned total = (unsigned)n * h * w;
  const unsigned tiles = (total + kFcPix - 1) / kFcPix;
  for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const unsigned p0 = tile * kFcPix;
    __syncthreads();                                 // the previous tile's fragments (first: the weights) have been read
    {
      float v[9];
      taps_of(x, p0 + threadIdx.x, total, h, w, v);
      __nv_bfloat16 xh[9], xl[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) split_bf16(v[t], xh[t], xl[t]);
      const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
      uint4* row = reinterpret_cast<uint4*>(As + threadIdx.x * kFcRow);
      // [xh0..8 | xl0..8 | xh0..8 | 0 x 5]
      row[0] = make_uint4(pack2(xh[0], xh[1]), pack2(xh[2], xh[3]), pack2(xh[4], xh[5]), pack2(xh[6], xh[7]));
      row[1] = make_uint4(pack2(xh[8], xl[0]), pack2(xl[1], xl[2]), pack2(xl[3], xl[4]), pack2(xl[5], xl[6]));
      row[2] = make_uint4(pack2(xl[7], xl[8]), pack2(xh[0], xh[1]), pack2(xh[2], xh[3]), pack2(xh[4], xh[5]));
      row[3] = make_uint4(pack2(xh[6], xh[7]), pack2(xh[8], z), pack2(z, z), pack2(z, z));
    }
    __syncthreads();
    for (int mt = mgrp; mt < kFcPix / 16; mt += mstride) {
      if (p0 + mt * 16 >= total) break;
      uint32_t af[2][4];
      const __nv_bfloat16* ap = As + (mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kFcRow + 8 * (lane >> 4);
      ldmatrix_x4(af[0], ap);
      ldmatrix_x4(af[1], ap + 16);
      float c[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
        mma16816(c[j], af[0], bf[j][0]);
        mma16816(c[j], af[1], bf[j][1]);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const unsigned p = p0 + mt * 16 + g + 8 * half;
        if (p < total) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[2 * j] = c[j][2 * half] + br[2 * j];
            o[2 * j + 1] = c[j][2 * half + 1] + br[2 * j + 1];
          }
          prelu_store8(y + (size_t)p * cout + 32 * blk + 8 * q, o, a);
        }
      }
    }
  }
}

// ---- weight / bias gradient ------------------------------------------------------------------------------------------
// ws[block][co * 10 + r] (r < 9: taps, r = 9: bias) - the layout of conv_first2_bwd_kernel, reduced by rows_reduce_kernel.
// MT = m-tiles (16 channels) per warp: cout = 128 * MT.
template <int MT>
__global__ void __launch_bounds__(kFcThreads) firstconv_dw_kernel(const float* __restrict__ x, int n, int h, int w,
                                                                  const __nv_bfloat16* __restrict__ dz,
                                                                  float* __restrict__ ws) {
  constexpr int COUT = 128 * MT;
  constexpr int DZROW = COUT + 8;                    // bf16 elements per staged dz row (+16 bytes: conflict-free ldmatrix)
  constexpr int DZBUF = kBwPix * DZROW;
  constexpr int SMEM_E = 2 * DZBUF > COUT * kBwCols * 2 ? 2 * DZBUF : COUT * kBwCols * 2;   // also holds fp32 [COUT][24]
  __shared__ __align__(16) __nv_bfloat16 dzs[SMEM_E];
  __shared__ __align__(16) __nv_bfloat16 vs[2][kBwCols * kBwVRow];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const unsigned total = (unsigned)n * h * w;
  const unsigned chunks = (total + kBwPix - 1) / kBwPix;
  // constant columns of V: 9 = one (the bias gradient), 19..23 = zero
  for (int i = threadIdx.x; i < 2 * kBwCols * kBwVRow; i += kFcThreads) {
    const int col = (i % (kBwCols * kBwVRow)) / kBwVRow;
    (&vs[0][0])[i] = __float2bfloat16_rn(col == 9 ? 1.f : 0.f);
  }
  auto stage = [&](unsigned chunk, int buf) {
    const unsigned p0 = chunk * kBwPix;
    // dz rows: COUT * 2 / 16 vectors per pixel
    constexpr int VPR = COUT / 8;
#pragma unroll
    for (int i = threadIdx.x; i < kBwPix * VPR; i += kFcThreads) {
      const int pl = i / VPR, v = i % VPR;
      const unsigned p = p0 + pl;
      const bool ok = p < total;
      cp_async16(dzs + buf * DZBUF + pl * DZROW + v * 8, dz + (size_t)(ok ? p : 0) * COUT + v * 8, ok);
    }
    cp_async_commit();
    if (warp == 0) {                                 // one pixel per lane: its nine samples, split
      float v[9];
      taps_of(x, p0 + lane, total, h, w, v);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        __nv_bfloat16 hi, lo;
        split_bf16(v[t], hi, lo);
        vs[buf][t * kBwVRow + lane] = hi;
        vs[buf][(10 + t) * kBwVRow + lane] = lo;
      }
    }
  };
  float acc[MT][3][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
  __syncthreads();
  int buf = 0;
  if (blockIdx.x < chunks) stage(blockIdx.x, 0);
  for (unsigned chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x, buf ^= 1) {
    const unsigned next = chunk + gridDim.x;
    if (next < chunks) {
      stage(next, buf ^ 1);                          // the other buffer was released by the barrier ending the last step
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kBwPix / 16; ++ks) {
      uint32_t bfr[3][2];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const __nv_bfloat16* vp = &vs[buf][(nt * 8 + g) * kBwVRow + ks * 16 + 2 * q];
        bfr[nt][0] = *reinterpret_cast<const uint32_t*>(vp);
        bfr[nt][1] = *reinterpret_cast<const uint32_t*>(vp + 8);
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int m0 = (warp * MT + m) * 16;
        uint32_t af[4];
        // matrices: (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15), each transposed on load
        const int mi = lane >> 3;
        ldmatrix_x4_trans(af, dzs + buf * DZBUF + (ks * 16 + 8 * (mi >> 1) + (lane & 7)) * DZROW + m0 + 8 * (mi & 1));
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma16816(acc[m][nt], af, bfr[nt]);
      }
    }
    __syncthreads();
  }
  // fold hi + lo columns through shared memory: part[co][24] fp32
  float* part = reinterpret_cast<float*>(dzs);
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const int m0 = (warp * MT + m) * 16;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      part[(m0 + g) * kBwCols + nt * 8 + 2 * q] = acc[m][nt][0];
      part[(m0 + g) * kBwCols + nt * 8 + 2 * q + 1] = acc[m][nt][1];
      part[(m0 + g + 8) * kBwCols + nt * 8 + 2 * q] = acc[m][nt][2];
      part[(m0 + g + 8) * kBwCols + nt * 8 + 2 * q + 1] = acc[m][nt][3];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * 10; i += kFcThreads) {
    const int co = i / 10, r = i % 10;
    const float v = r < 9 ? part[co * kBwCols + r] + part[co * kBwCols + 10 + r] : part[co * kBwCols + 9];
    ws[(size_t)blockIdx.x * COUT * 10 + i] = v;
  }
}

}  // namespace

// (callers also require n * h * w < 2^31 - 2^16: 32-bit pixel arithmetic in the kernels)
bool firstconv_mma_supported(int dtype, int cin, int cout) {
  return dtype == VSR_BF16 && cin == 1 && (cout == 32 || cout == 64 || cout == 128 || cout == 256) && tunables().fc_simt != 1;
}
bool firstconv_mma_bwd_supported(int dtype, int cin, int cout) {
  return dtype == VSR_BF16 && cin == 1 && (cout == 128 || cout == 256) && tunables().fc_simt != 1;
}

void firstconv_mma_fwd(const float* x, int n, int h, int w, const float* wt, const float* bias, const float* slope, void* y,
                       int cout, cudaStream_t s) {
  const long tiles = ((long)n * h * w + kFcPix - 1) / kFcPix;
  const int grid = (int)std::min<long>(tiles, 4l * num_sms());       // 62 registers: four CTAs per SM
  firstconv_fwd_kernel<<<grid, kFcThreads, 0, s>>>(x, n, h, w, wt, bias, slope, static_cast<__nv_bfloat16*>(y), cout);
}

// returns the number of partial rows written to ws ([blocks][cout * 10])
int firstconv_mma_bwd(const float* x, int n, int h, int w, const void* dz, int cout, float* ws, int max_blocks, cudaStream_t s) {
  const long chunks = ((long)n * h * w + kBwPix - 1) / kBwPix;
  const int blocks = (int)std::min<long>(std::min<long>(chunks, 4l * num_sms()), max_blocks);
  if (cout == 256) firstconv_dw_kernel<2><<<blocks, kFcThreads, 0, s>>>(x, n, h, w, static_cast<const __nv_bfloat16*>(dz), ws);
  else firstconv_dw_kernel<1><<<blocks, kFcThreads, 0, s>>>(x, n, h, w, static_cast<const __nv_bfloat16*>(dz), ws);
  return blocks;
}

}  // namespace vsr
