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
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

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

// grid sized in whole waves of the SM count
inline int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace vsr
