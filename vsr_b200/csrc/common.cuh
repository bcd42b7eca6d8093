// common.cuh — shared host/device helpers for libvsr_sm100.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vsr_b200.h"

namespace vsr {

// ---- error reporting (host) -------------------------------------------------------------
void set_error(const char* fmt, ...);
int num_sms();

// tuning overrides from the environment (VSR_TC_*, VSR_WG_*, VSR_PDL; not part of the ABI): read once, -1 = unset.
// vsr_reload_tunables() re-reads them (tests / tools); the launch path never calls getenv.
struct Tunables {
  int tc_debug, tc_tall, tc_tall_stages, tc_resident, tc_stages, tc_grid, tc_square, tc_pair, tc_epibuf, pdl, wg_debug, wg_tall, up_generic, fc_simt;
};
const Tunables& tunables();      // tma_host.cu

#define VSR_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      vsr::set_error(__VA_ARGS__);    \
      return VSR_ERR_BAD_ARG;         \
    }                                 \
  } while (0)

#define VSR_CHECK_SUPPORTED(cond, ...) \
  do {                                 \
    if (!(cond)) {                     \
      vsr::set_error(__VA_ARGS__);     \
      return VSR_ERR_UNSUPPORTED;      \
    }                                  \
  } while (0)

#define VSR_CHECK_LAUNCH(what)                                                        \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      vsr::set_error("%s: launch failed: %s", what, cudaGetErrorString(e__));         \
      return VSR_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

constexpr int kPartialsLen = 1024;  // floats per slope/loss partial row (>= any grid we launch)

// ---- device helpers -----------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  // value with [tag] in the mantissa LSB / the LSB of a stored value (PReLU with a negative slope, below)
  static __device__ __forceinline__ void st_tag(float* p, float v, bool tag) {
    *p = __uint_as_float((__float_as_uint(v) & ~1u) | (tag ? 1u : 0u));
  }
  static __device__ __forceinline__ bool ld_tag(const float* p) { return (__float_as_uint(*p) & 1u) != 0u; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ void st_tag(__nv_bfloat16* p, float v, bool tag) {
    const unsigned short b = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    *p = __ushort_as_bfloat16((unsigned short)((b & 0xfffeu) | (tag ? 1u : 0u)));
  }
  static __device__ __forceinline__ bool ld_tag(const __nv_bfloat16* p) { return (__bfloat16_as_ushort(*p) & 1u) != 0u; }
};

// ---- PReLU with an unconstrained learnable slope (nn.PReLU(num_parameters=1), drf_net.py:56) ----------
// The backward pass reads the stored post-activation y, never the pre-activation x:
//   slope > 0 : sign(y) == sign(x) and x = y / slope on the non-positive side;
//   slope == 0: the forward multiplies the non-positive side by kPreluTiny instead of 0 (changes y by at most
//               6e-8 |x|), so x = y / kPreluTiny stays recoverable for d(slope); dx uses the true slope (0);
//   slope < 0 : y > 0 on both sides, so the forward stores [x > 0] in the mantissa LSB of y (<= 1 ulp of the
//               storage type) and the backward reads the branch from there.
constexpr float kPreluTiny = 5.9604644775390625e-08f;   // 2^-24
struct Prelu {
  float slope;   // true slope: dx = slope * dy where x <= 0
  float fwd;     // factor the forward applies where x <= 0 (slope, or kPreluTiny when slope == 0)
  float inv;     // 1 / fwd
  bool tag;      // slope < 0: the branch travels in the LSB of y
};
__device__ __forceinline__ Prelu make_prelu(float slope) {
  Prelu a;
  a.slope = slope;
  a.fwd = slope == 0.f ? kPreluTiny : slope;
  a.inv = 1.f / a.fwd;
  a.tag = slope < 0.f;
  return a;
}
// y = PReLU(x) stored to p (tagged when the slope is negative); returns y as the consumers will read it back
template <typename T>
__device__ __forceinline__ void prelu_store(T* p, float x, const Prelu& a) {
  const bool pos = x > 0.f;
  const float y = pos ? x : a.fwd * x;
  if (a.tag) Elem<T>::st_tag(p, y, pos); else Elem<T>::st(p, y);
}
// backward through a stored y: g <- dL/dx, returns the d(slope) contribution g * x * [x <= 0]
template <typename T>
__device__ __forceinline__ float prelu_bwd(const T* yp, float& g, const Prelu& a) {
  const float y = Elem<T>::ld(yp);
  const bool pos = a.tag ? Elem<T>::ld_tag(yp) : y > 0.f;
  const float da = pos ? 0.f : g * (y * a.inv);
  g = pos ? g : a.slope * g;
  return da;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum in a fixed order; result valid on thread 0. `red` holds >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += red[i];
  }
  return s;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// PReLU of 8 consecutive channels stored as one 16-byte (bf16) / two 16-byte (fp32) vectors
template <typename T>
__device__ __forceinline__ void prelu_store8(T* p, const float (&x)[8], const Prelu& a) {
  float y[8];
  uint32_t pos = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    pos |= (x[j] > 0.f ? 1u : 0u) << j;
    y[j] = x[j] > 0.f ? x[j] : a.fwd * x[j];
  }
  if constexpr (sizeof(T) == 2) {
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      o[q] = pack_bf16x2(y[2 * q], y[2 * q + 1]);
      if (a.tag) o[q] = (o[q] & 0xfffefffeu) | ((pos >> (2 * q)) & 1u) | (((pos >> (2 * q + 1)) & 1u) << 16);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
    if (a.tag) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = __uint_as_float((__float_as_uint(y[j]) & ~1u) | ((pos >> j) & 1u));
    }
    reinterpret_cast<float4*>(p)[0] = make_float4(y[0], y[1], y[2], y[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(y[4], y[5], y[6], y[7]);
  }
}

// grid sized in whole waves of the SM count
inline int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace vsr
