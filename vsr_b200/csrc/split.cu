// split.cu — the strict mode on tensor cores ("bf16x3"): fp32 products emulated by error-compensated bf16 pairs.
//
// An fp32 value x is carried as two bf16 planes xh = bf16(x), xl = bf16(x - xh) (16 significant bits, |x - xh - xl| <=
// 2^-17 |x|); a product x*w is xh*wh + xl*wh + xh*wl (the dropped xl*wl term is 2^-16 relative to 2^-8-sized terms), each
// an exact bf16 x bf16 product accumulated in fp32 by tcgen05 (kind::f16).  The tap-GEMM expresses the three terms as
// three taps (bit 3 of a tap's source index selects the low-order plane, the weight buffer holds [wh | wh | wl] slabs per
// tap: VSR_BF16X2 in include/vsr_b200.h), writes the raw fp32 accumulators, and the kernels below do the rest in fp32:
//   vsr_split_planes  fp32 map -> the two bf16 planes a tap-GEMM / weight-gradient launch reads,
//   vsr_tap_epilogue  bias / scale / residual / (P)ReLU forward and backward / second output, in place on the accumulators
//                     (the same order and arithmetic as the epilogue of the CUDA-core kernel, tapgemm_simt.cu),
//   vsr_gather_split  flat fp32 parameters -> packed bf16 slabs, high or low part per element.
// Replaces aten.convolution / convolution_backward under drf_net.py:55-106,141-147 in precision='bf16x3'.
#include "common.cuh"

namespace vsr {
namespace {

__device__ __forceinline__ void split2(float x0, float x1, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
  const float r0 = x0 - __bfloat162float(h0), r1 = x1 - __bfloat162float(h1);      // exact in fp32
  *hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  *lo = pack_bf16x2(r0, r1);
}

// 8 elements per thread: two 16-byte loads, one 16-byte store per plane
__global__ void __launch_bounds__(256) split_planes_kernel(const float4* __restrict__ x, uint4* __restrict__ hi,
                                                           uint4* __restrict__ lo, long n8) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
    uint4 h, l;
    split2(a.x, a.y, &h.x, &l.x);
    split2(a.z, a.w, &h.y, &l.y);
    split2(b.x, b.y, &h.z, &l.z);
    split2(b.z, b.w, &h.w, &l.w);
    hi[i] = h;
    lo[i] = l;
  }
}

struct EpiArgs {
  float* out;
  const float* bias;
  const float* slope;
  const float* residual;
  const float* aux_y;
  float* out2;
  const float* res2;
  float* slope_partials;
  uint2* planes;   // optional: the two bf16 planes of `out` ([2][numel], what vsr_split_planes(out) would give), written here
  uint2* planes2;  // optional: the same for `out2`
  float out_scale;
  int epi;
  int c4;          // channels / 4
  long n4;         // elements / 4
};

template <int EPI>
__global__ void __launch_bounds__(256) tap_epilogue_kernel(const EpiArgs a) {
  __shared__ float red[32];
  const int epi = EPI >= 0 ? EPI : a.epi;
  const Prelu pr = make_prelu((epi & (VSR_EPI_PRELU | VSR_EPI_PRELU_BWD)) ? __ldg(a.slope) : 1.f);
  float slope_acc = 0.f;
  float4* out4 = reinterpret_cast<float4*>(a.out);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < a.n4; i += (long)gridDim.x * blockDim.x) {
    const float4 acc = out4[i];
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    if (epi & VSR_EPI_BIAS) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias) + (int)(i % a.c4));
      v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
    if (epi & VSR_EPI_SCALE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= a.out_scale;
    }
    if (epi & VSR_EPI_RES_PRE) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(a.residual) + i);
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    if (epi & VSR_EPI_RELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (epi & VSR_EPI_PRELU_BWD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) slope_acc += prelu_bwd(a.aux_y + 4 * i + j, v[j], pr);
    } else if (epi & VSR_EPI_RELU_BWD) {
      const float4 y = __ldg(reinterpret_cast<const float4*>(a.aux_y) + i);
      v[0] = y.x > 0.f ? v[0] : 0.f; v[1] = y.y > 0.f ? v[1] : 0.f;
      v[2] = y.z > 0.f ? v[2] : 0.f; v[3] = y.w > 0.f ? v[3] : 0.f;
    }
    if (epi & VSR_EPI_PRELU) {
      float y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool pos = v[j] > 0.f;
        v[j] = pos ? v[j] : pr.fwd * v[j];
        // a negative slope: the stored y carries [x > 0] in its LSB (Elem<float>::st_tag); out2 takes the untagged value
        y[j] = pr.tag ? __uint_as_float((__float_as_uint(v[j]) & ~1u) | (pos ? 1u : 0u)) : v[j];
      }
      out4[i] = make_float4(y[0], y[1], y[2], y[3]);
    } else {
      out4[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (a.planes) {
      uint2 h, l;
      split2(v[0], v[1], &h.x, &l.x);
      split2(v[2], v[3], &h.y, &l.y);
      a.planes[i] = h;
      a.planes[a.n4 + i] = l;
    }
    if (epi & VSR_EPI_OUT2) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(a.res2) + i);
      const float s2 = (epi & VSR_EPI_OUT2_SUB) ? -1.f : 1.f;
      const float4 o2 = make_float4(fmaf(s2, r.x, v[0]), fmaf(s2, r.y, v[1]), fmaf(s2, r.z, v[2]), fmaf(s2, r.w, v[3]));
      reinterpret_cast<float4*>(a.out2)[i] = o2;
      if (a.planes2) {
        uint2 h, l;
        split2(o2.x, o2.y, &h.x, &l.x);
        split2(o2.z, o2.w, &h.y, &l.y);
        a.planes2[i] = h;
        a.planes2[a.n4 + i] = l;
      }
    }
  }
  if (epi & VSR_EPI_PRELU_BWD) {
    const float s = block_sum(slope_acc, red);
    if (threadIdx.x == 0) a.slope_partials[blockIdx.x] = s;
  }
}

constexpr int kLoFlag = 1 << 30;

__global__ void gather_split_kernel(const float* __restrict__ src, const int* __restrict__ idx, __nv_bfloat16* __restrict__ dst,
                                    long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int j = __ldg(idx + i);
    float v = 0.f;
    if (j >= 0) {
      const float w = __ldg(src + (j & (kLoFlag - 1)));
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      v = (j & kLoFlag) ? w - __bfloat162float(h) : __bfloat162float(h);
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_split_planes(const float* x, void* planes, int64_t numel, void* stream) {
  VSR_CHECK_ARG(x && planes && numel > 0 && numel % 8 == 0, "vsr_split_planes: bad arguments (numel must be a multiple of 8)");
  VSR_CHECK_ARG(((uintptr_t)x | (uintptr_t)planes) % 16 == 0, "vsr_split_planes: pointers must be 16-byte aligned");
  const long n8 = numel / 8;
  uint4* hi = static_cast<uint4*>(planes);
  split_planes_kernel<<<grid_for(n8, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), hi, hi + n8, n8);
  VSR_CHECK_LAUNCH("vsr_split_planes");
  return VSR_OK;
}

extern "C" int vsr_tap_epilogue(float* out, int64_t rows, int32_t c, const float* bias, int32_t epi, float out_scale,
                                const float* slope, const float* residual, const float* aux_y, float* out2, const float* res2,
                                float* slope_partials, void* planes, void* planes2, void* stream) {
  VSR_CHECK_ARG(out && rows > 0 && c > 0 && c % 4 == 0, "vsr_tap_epilogue: bad arguments (c must be a multiple of 4)");
  if (epi & VSR_EPI_BIAS) VSR_CHECK_ARG(bias, "vsr_tap_epilogue: BIAS without bias");
  if (epi & VSR_EPI_RES_PRE) VSR_CHECK_ARG(residual, "vsr_tap_epilogue: RES_PRE without residual");
  if (epi & (VSR_EPI_PRELU | VSR_EPI_PRELU_BWD)) VSR_CHECK_ARG(slope, "vsr_tap_epilogue: PReLU without slope");
  if (epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) VSR_CHECK_ARG(aux_y, "vsr_tap_epilogue: *_BWD without aux_y");
  if (epi & VSR_EPI_PRELU_BWD) VSR_CHECK_ARG(slope_partials, "vsr_tap_epilogue: PRELU_BWD without slope_partials");
  if (epi & VSR_EPI_OUT2) VSR_CHECK_ARG(out2 && res2, "vsr_tap_epilogue: OUT2 without out2/res2");
  if (epi == 0 && !planes) return VSR_OK;
  VSR_CHECK_ARG(!planes2 || (epi & VSR_EPI_OUT2), "vsr_tap_epilogue: planes2 without OUT2");
  EpiArgs a;
  a.planes = static_cast<uint2*>(planes);
  a.planes2 = static_cast<uint2*>(planes2);
  a.out = out; a.bias = bias; a.slope = slope; a.residual = residual; a.aux_y = aux_y; a.out2 = out2; a.res2 = res2;
  a.slope_partials = slope_partials; a.out_scale = out_scale; a.epi = epi; a.c4 = c / 4; a.n4 = rows * (long)(c / 4);
  int grid = grid_for(a.n4, 256, 6);
  if (grid > kPartialsLen) grid = kPartialsLen;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the flag sets of the DRFNet schedule are compile-time constants; anything else takes the generic kernel
  switch (epi) {
#define VSR_EPI_CASE(E) case (E): tap_epilogue_kernel<(E)><<<grid, 256, 0, s>>>(a); break;
    VSR_EPI_CASE(VSR_EPI_BIAS)
    VSR_EPI_CASE(VSR_EPI_BIAS | VSR_EPI_PRELU)
    VSR_EPI_CASE(VSR_EPI_BIAS | VSR_EPI_PRELU | VSR_EPI_OUT2)
    VSR_EPI_CASE(VSR_EPI_PRELU_BWD)
    VSR_EPI_CASE(VSR_EPI_PRELU_BWD | VSR_EPI_RES_PRE)
    VSR_EPI_CASE(VSR_EPI_RES_PRE)
#undef VSR_EPI_CASE
    default: tap_epilogue_kernel<-1><<<grid, 256, 0, s>>>(a); break;
  }
  VSR_CHECK_LAUNCH("vsr_tap_epilogue");
  return VSR_OK;
}

extern "C" int vsr_gather_split(const float* src, const int32_t* idx, void* dst, int64_t n, void* stream) {
  VSR_CHECK_ARG(src && idx && dst && n >= 0, "vsr_gather_split: bad arguments");
  if (n == 0) return VSR_OK;
  gather_split_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, idx, static_cast<__nv_bfloat16*>(dst), n);
  VSR_CHECK_LAUNCH("vsr_gather_split");
  return VSR_OK;
}
