// core.cu — error state, device query and the small bandwidth-bound helper kernels
// (gather/pack, cast, add, activation backward, column sums, partial reductions, Adam).
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace vsr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

namespace {

template <typename D>
__global__ void gather_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                              D* __restrict__ dst, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int j = __ldg(idx + i);
    Elem<D>::st(dst + i, j >= 0 ? __ldg(src + j) : 0.f);
  }
}

__global__ void gather_add_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                  float* __restrict__ dst, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int j = __ldg(idx + i);
    if (j >= 0) dst[i] += __ldg(src + j);
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    Elem<D>::st(dst + i, Elem<S>::ld(src + i));
}

// 16-byte vectorised elementwise add (fp32 x4 or bf16 x8)
__global__ void add_f32_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                               float4* __restrict__ o, long n4) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 x = __ldg(a + i), y = __ldg(b + i);
    o[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}
// any length / alignment (odd frame sizes of x3 nets): one element per thread and trip
template <typename T>
__global__ void add_scalar_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    Elem<T>::st(o + i, Elem<T>::ld(a + i) + Elem<T>::ld(b + i));
}
__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                uint4* __restrict__ o, long n8) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const uint4 x = __ldg(a + i), y = __ldg(b + i);
    uint4 r;
    r.x = pack_bf16x2(bf16_lo(x.x) + bf16_lo(y.x), bf16_hi(x.x) + bf16_hi(y.x));
    r.y = pack_bf16x2(bf16_lo(x.y) + bf16_lo(y.y), bf16_hi(x.y) + bf16_hi(y.y));
    r.z = pack_bf16x2(bf16_lo(x.z) + bf16_lo(y.z), bf16_hi(x.z) + bf16_hi(y.z));
    r.w = pack_bf16x2(bf16_lo(x.w) + bf16_lo(y.w), bf16_hi(x.w) + bf16_hi(y.w));
    o[i] = r;
  }
}

// dz = y>0 ? dy : a*dy ; slope partial += dy * y/a on the non-positive side. 4 elements/thread/step.
template <typename T, bool kPrelu>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz,
                               long n, const float* __restrict__ slope_p, float* __restrict__ partials) {
  __shared__ float red[32];
  const Prelu a = make_prelu(kPrelu ? __ldg(slope_p) : 1.f);
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float g = Elem<T>::ld(dy + i);
    if (kPrelu) {
      acc += prelu_bwd(y + i, g, a);
    } else {
      g = Elem<T>::ld(y + i) > 0.f ? g : 0.f;
    }
    Elem<T>::st(dz + i, g);
  }
  if (kPrelu) {
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
  }
}

// the same with 16-byte vectors (8 bf16 / 4 fp32 per thread and step); block 0 also takes the n % VEC tail elements
template <typename T, bool kPrelu>
__global__ void __launch_bounds__(256) act_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz,
                                                         long n, const float* __restrict__ slope_p,
                                                         float* __restrict__ partials) {
  constexpr int VEC = 16 / sizeof(T);
  union U {
    uint4 q;
    T e[VEC];
  };
  __shared__ float red[32];
  const Prelu a = make_prelu(kPrelu ? __ldg(slope_p) : 1.f);
  float acc = 0.f;
  const long nvec = n / VEC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
    U g_in, y_in, out;
    g_in.q = __ldg(reinterpret_cast<const uint4*>(dy) + i);
    y_in.q = __ldg(reinterpret_cast<const uint4*>(y) + i);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float g = Elem<T>::ld(&g_in.e[j]);
      if (kPrelu) {
        acc += prelu_bwd(&y_in.e[j], g, a);
      } else {
        g = Elem<T>::ld(&y_in.e[j]) > 0.f ? g : 0.f;
      }
      Elem<T>::st(&out.e[j], g);
    }
    reinterpret_cast<uint4*>(dz)[i] = out.q;
  }
  if (blockIdx.x == 0 && threadIdx.x < n - nvec * VEC) {
    const long i = nvec * VEC + threadIdx.x;
    float g = Elem<T>::ld(dy + i);
    if (kPrelu) {
      acc += prelu_bwd(y + i, g, a);
    } else {
      g = Elem<T>::ld(y + i) > 0.f ? g : 0.f;
    }
    Elem<T>::st(dz + i, g);
  }
  if (kPrelu) {
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
  }
}

// column sums, pass 1: block b sums rows [b*rpb, (b+1)*rpb); 16-byte vector loads, each thread owns
// VEC consecutive columns of every (blockDim/TPR)-th row; fixed-order shared-memory fold.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long rows, int c, long rows_per_block,
                                                    float* __restrict__ ws) {
  constexpr int VEC = 16 / sizeof(T);
  extern __shared__ float sm[];           // [rpp][c]
  const int tpr = c / VEC;                // threads per row
  const int rpp = blockDim.x / tpr;       // rows per pass
  const int cg = threadIdx.x % tpr, ro = threadIdx.x / tpr;
  const long r0 = blockIdx.x * rows_per_block;
  long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  if (ro < rpp) {
    auto add = [&](const uint4& q) {
      if constexpr (sizeof(T) == 2) {
        acc[0] += bf16_lo(q.x); acc[1] += bf16_hi(q.x); acc[2] += bf16_lo(q.y); acc[3] += bf16_hi(q.y);
        acc[4] += bf16_lo(q.z); acc[5] += bf16_hi(q.z); acc[6] += bf16_lo(q.w); acc[7] += bf16_hi(q.w);
      } else {
        acc[0] += __uint_as_float(q.x); acc[1] += __uint_as_float(q.y);
        acc[2] += __uint_as_float(q.z); acc[3] += __uint_as_float(q.w);
      }
    };
    long r = r0 + ro;
    for (; r + 3l * rpp < r1; r += 4l * rpp) {             // four loads in flight, added in row order
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(x + (r + (long)u * rpp) * c + cg * VEC));
#pragma unroll
      for (int u = 0; u < 4; ++u) add(q[u]);
    }
    for (; r < r1; r += rpp) add(__ldg(reinterpret_cast<const uint4*>(x + r * c + cg * VEC)));
#pragma unroll
    for (int i = 0; i < VEC; ++i) sm[ro * c + cg * VEC + i] = acc[i];
  }
  __syncthreads();
  for (int col = threadIdx.x; col < c; col += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rpp; ++r) s += sm[r * c + col];
    ws[(size_t)blockIdx.x * c + col] = s;
  }
}
// generic (any c, scalar loads)
template <typename T>
__global__ void colsum_scalar_kernel(const T* __restrict__ x, long rows, int c, long rows_per_block,
                                     float* __restrict__ ws) {
  const long r0 = blockIdx.x * rows_per_block;
  long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  for (int col = threadIdx.x; col < c; col += blockDim.x) {
    float s = 0.f;
    for (long r = r0; r < r1; ++r) s += Elem<T>::ld(x + r * c + col);
    ws[(size_t)blockIdx.x * c + col] = s;
  }
}
// pass 2: a block owns 64 columns; thread (part, col) sums every 4th partial row (eight loads in flight, fixed order), then
// a fixed-order fold over the four parts
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ ws, int blocks, int c,
                                                          float* __restrict__ db, int accumulate) {
  __shared__ float sm[4][64];
  const int col = blockIdx.x * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;
  float s = 0.f;
  if (col < c) {
    int b = part;
    for (; b + 28 < blocks; b += 32) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(ws + (size_t)(b + 4 * u) * c + col);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; b < blocks; b += 4) s += ws[(size_t)b * c + col];
  }
  sm[part][threadIdx.x & 63] = s;
  __syncthreads();
  if (threadIdx.x < 64 && col < c) {
    float t = accumulate ? db[col] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) t += sm[q][threadIdx.x];
    db[col] = t;
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partials, int rows, int len,
                                       const int* __restrict__ row_dst, float* __restrict__ dst) {
  // one warp per row; lanes stride the row, then a fixed butterfly; rows mapping to the same
  // destination are serialised by a single thread afterwards to keep the order fixed.
  extern __shared__ float row_sums[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < rows; r += nw) {
    float s = 0.f;
    for (int i = lane; i < len; i += 32) s += partials[(size_t)r * len + i];
    s = warp_sum(s);
    if (lane == 0) row_sums[r] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int r = 0; r < rows; ++r) dst[row_dst[r]] += row_sums[r];
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
                            float wd, float bc1, float bc2_sqrt, float gscale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

__global__ void scale_kernel(float* __restrict__ x, long n, float alpha) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] *= alpha;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_bc1, float b1, float b2,
                                         float bc2_sqrt, float eps, float wd, float gscale) {
  float gi = g * gscale;
  if (wd != 0.f) gi = fmaf(wd, p, gi);
  m = b1 * m + (1.f - b1) * gi;
  v = b2 * v + (1.f - b2) * gi * gi;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - lr_bc1 * (m / denom);
}

// VEC = 4: 16-byte accesses (all four pointers 16-byte aligned), block 0 takes the n % 4 tail; VEC = 1: any alignment
template <int VEC>
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                      float* __restrict__ v, long n, const float* __restrict__ hyper) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step = hyper[5],
              gscale = hyper[6];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float lr_bc1 = lr / bc1;
  const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
  if constexpr (VEC == 4) {
    const long nvec = n / 4;
    for (long i = tid; i < nvec; i += nthr) {
      float4 pi = reinterpret_cast<float4*>(p)[i];
      const float4 gi = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
      adam_one(pi.x, gi.x, mi.x, vi.x, lr_bc1, b1, b2, bc2_sqrt, eps, wd, gscale);
      adam_one(pi.y, gi.y, mi.y, vi.y, lr_bc1, b1, b2, bc2_sqrt, eps, wd, gscale);
      adam_one(pi.z, gi.z, mi.z, vi.z, lr_bc1, b1, b2, bc2_sqrt, eps, wd, gscale);
      adam_one(pi.w, gi.w, mi.w, vi.w, lr_bc1, b1, b2, bc2_sqrt, eps, wd, gscale);
      reinterpret_cast<float4*>(m)[i] = mi;
      reinterpret_cast<float4*>(v)[i] = vi;
      reinterpret_cast<float4*>(p)[i] = pi;
    }
    if (blockIdx.x == 0 && threadIdx.x < n - nvec * 4) {
      const long i = nvec * 4 + threadIdx.x;
      adam_one(p[i], g[i], m[i], v[i], lr_bc1, b1, b2, bc2_sqrt, eps, wd, gscale);
    }
  } else {
    for (long i = tid; i < n; i += nthr) adam_one(p[i], g[i], m[i], v[i], lr_bc1, b1, b2, bc2_sqrt, eps, wd, gscale);
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_abi_version(void) { return VSR_ABI_VERSION; }
extern "C" const char* vsr_last_error(void) { return g_err; }
extern "C" int vsr_partials_len(void) { return kPartialsLen; }

extern "C" int64_t vsr_slab_index(int32_t j, int32_t k) {
  return (int64_t)j * 64 + (((k >> 3) ^ (j & 7)) << 3) + (k & 7);
}

extern "C" int vsr_gather(const float* src, const int32_t* idx, void* dst, int32_t dst_dtype, int64_t n,
                          void* stream) {
  VSR_CHECK_ARG(src && idx && dst && n >= 0, "vsr_gather: bad arguments");
  if (n == 0) return VSR_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, 256);
  if (dst_dtype == VSR_F32)
    gather_kernel<float><<<grid, 256, 0, s>>>(src, idx, static_cast<float*>(dst), n);
  else if (dst_dtype == VSR_BF16)
    gather_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, idx, static_cast<__nv_bfloat16*>(dst), n);
  else
    VSR_CHECK_ARG(false, "vsr_gather: bad dtype %d", dst_dtype);
  VSR_CHECK_LAUNCH("vsr_gather");
  return VSR_OK;
}

extern "C" int vsr_gather_add(const float* src, const int32_t* idx, float* dst, int64_t n, void* stream) {
  VSR_CHECK_ARG(src && idx && dst && n >= 0, "vsr_gather_add: bad arguments");
  if (n == 0) return VSR_OK;
  gather_add_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, idx, dst, n);
  VSR_CHECK_LAUNCH("vsr_gather_add");
  return VSR_OK;
}

extern "C" int vsr_cast(const void* src, int32_t sd, void* dst, int32_t dd, int64_t n, void* stream) {
  VSR_CHECK_ARG(src && dst && n >= 0, "vsr_cast: bad arguments");
  if (n == 0) return VSR_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n, 256);
  if (sd == VSR_F32 && dd == VSR_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (sd == VSR_BF16 && dd == VSR_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (sd == VSR_F32 && dd == VSR_F32)
    cast_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, (float*)dst, n);
  else if (sd == VSR_BF16 && dd == VSR_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else
    VSR_CHECK_ARG(false, "vsr_cast: bad dtypes %d -> %d", sd, dd);
  VSR_CHECK_LAUNCH("vsr_cast");
  return VSR_OK;
}

namespace vsr {
namespace {
template <typename T>
__global__ void axpby_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long n, float alpha, float beta) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = alpha * Elem<T>::ld(a + i);
    if (b) v = fmaf(beta, Elem<T>::ld(b + i), v);
    Elem<T>::st(o + i, v);
  }
}
__global__ void axpby_bf16x8_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o, long n8,
                                    float alpha, float beta) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const uint4 x = __ldg(a + i);
    const uint4 y = b ? __ldg(b + i) : make_uint4(0u, 0u, 0u, 0u);
    uint4 r;
    r.x = pack_bf16x2(fmaf(beta, bf16_lo(y.x), alpha * bf16_lo(x.x)), fmaf(beta, bf16_hi(y.x), alpha * bf16_hi(x.x)));
    r.y = pack_bf16x2(fmaf(beta, bf16_lo(y.y), alpha * bf16_lo(x.y)), fmaf(beta, bf16_hi(y.y), alpha * bf16_hi(x.y)));
    r.z = pack_bf16x2(fmaf(beta, bf16_lo(y.z), alpha * bf16_lo(x.z)), fmaf(beta, bf16_hi(y.z), alpha * bf16_hi(x.z)));
    r.w = pack_bf16x2(fmaf(beta, bf16_lo(y.w), alpha * bf16_lo(x.w)), fmaf(beta, bf16_hi(y.w), alpha * bf16_hi(x.w)));
    o[i] = r;
  }
}
}  // namespace
}  // namespace vsr

extern "C" int vsr_axpby(const void* a, const void* b, void* out, int32_t dtype, int64_t numel, float alpha, float beta,
                         void* stream) {
  VSR_CHECK_ARG(a && out && numel >= 0 && (b || beta == 0.f), "vsr_axpby: bad arguments");
  if (numel == 0) return VSR_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (beta == 0.f) b = nullptr;
  const bool al16 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (dtype == VSR_F32) {
    axpby_kernel<float><<<grid_for(numel, 256), 256, 0, s>>>((const float*)a, (const float*)b, (float*)out, numel, alpha, beta);
  } else if (dtype == VSR_BF16) {
    using B = __nv_bfloat16;
    if (al16 && numel % 8 == 0)
      axpby_bf16x8_kernel<<<grid_for(numel / 8, 256), 256, 0, s>>>((const uint4*)a, (const uint4*)b, (uint4*)out, numel / 8, alpha, beta);
    else
      axpby_kernel<B><<<grid_for(numel, 256), 256, 0, s>>>((const B*)a, (const B*)b, (B*)out, numel, alpha, beta);
  } else {
    VSR_CHECK_ARG(false, "vsr_axpby: bad dtype %d", dtype);
  }
  VSR_CHECK_LAUNCH("vsr_axpby");
  return VSR_OK;
}

extern "C" int vsr_add(const void* a, const void* b, void* out, int32_t dtype, int64_t numel, void* stream) {
  VSR_CHECK_ARG(a && b && out && numel >= 0, "vsr_add: bad arguments");
  if (numel == 0) return VSR_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool al16 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (dtype == VSR_F32) {
    if (al16 && numel % 4 == 0)
      add_f32_kernel<<<grid_for(numel / 4, 256), 256, 0, s>>>((const float4*)a, (const float4*)b, (float4*)out, numel / 4);
    else
      add_scalar_kernel<float><<<grid_for(numel, 256), 256, 0, s>>>((const float*)a, (const float*)b, (float*)out, numel);
  } else if (dtype == VSR_BF16) {
    using B = __nv_bfloat16;
    if (al16 && numel % 8 == 0)
      add_bf16_kernel<<<grid_for(numel / 8, 256), 256, 0, s>>>((const uint4*)a, (const uint4*)b, (uint4*)out, numel / 8);
    else
      add_scalar_kernel<B><<<grid_for(numel, 256), 256, 0, s>>>((const B*)a, (const B*)b, (B*)out, numel);
  } else {
    VSR_CHECK_ARG(false, "vsr_add: bad dtype %d", dtype);
  }
  VSR_CHECK_LAUNCH("vsr_add");
  return VSR_OK;
}

extern "C" int vsr_act_bwd(const void* dy, const void* y, void* dz, int32_t dtype, int64_t numel,
                           const float* slope, float* slope_partials, void* stream) {
  VSR_CHECK_ARG(dy && y && dz && numel >= 0, "vsr_act_bwd: bad arguments");
  VSR_CHECK_ARG(!slope || slope_partials, "vsr_act_bwd: PReLU needs slope_partials");
  if (numel == 0) return VSR_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int grid = grid_for(numel, 256 * 4, 4);
  if (grid > kPartialsLen) grid = kPartialsLen;
  const bool vec = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0;
  if (dtype == VSR_F32) {
    using F = float;
    if (vec && slope) act_bwd_vec_kernel<F, true><<<grid, 256, 0, s>>>((const F*)dy, (const F*)y, (F*)dz, numel, slope, slope_partials);
    else if (vec) act_bwd_vec_kernel<F, false><<<grid, 256, 0, s>>>((const F*)dy, (const F*)y, (F*)dz, numel, nullptr, nullptr);
    else if (slope) act_bwd_kernel<F, true><<<grid, 256, 0, s>>>((const F*)dy, (const F*)y, (F*)dz, numel, slope, slope_partials);
    else act_bwd_kernel<F, false><<<grid, 256, 0, s>>>((const F*)dy, (const F*)y, (F*)dz, numel, nullptr, nullptr);
  } else if (dtype == VSR_BF16) {
    using B = __nv_bfloat16;
    if (vec && slope) act_bwd_vec_kernel<B, true><<<grid, 256, 0, s>>>((const B*)dy, (const B*)y, (B*)dz, numel, slope, slope_partials);
    else if (vec) act_bwd_vec_kernel<B, false><<<grid, 256, 0, s>>>((const B*)dy, (const B*)y, (B*)dz, numel, nullptr, nullptr);
    else if (slope) act_bwd_kernel<B, true><<<grid, 256, 0, s>>>((const B*)dy, (const B*)y, (B*)dz, numel, slope, slope_partials);
    else act_bwd_kernel<B, false><<<grid, 256, 0, s>>>((const B*)dy, (const B*)y, (B*)dz, numel, nullptr, nullptr);
  } else {
    VSR_CHECK_ARG(false, "vsr_act_bwd: bad dtype %d", dtype);
  }
  VSR_CHECK_LAUNCH("vsr_act_bwd");
  return VSR_OK;
}

static int colsum_blocks(int64_t rows) {
  int64_t b = (rows + 255) / 256;
  const int64_t cap = 4 * (int64_t)num_sms();            // four 256-thread blocks per SM, four 16-byte loads in flight each
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" size_t vsr_colsum_workspace(int64_t rows, int32_t c) {
  return (size_t)colsum_blocks(rows) * (size_t)c * sizeof(float);
}

template <typename T>
static int colsum_launch(const T* x, int64_t rows, int32_t c, float* ws, int blocks, long rpb, cudaStream_t s) {
  constexpr int VEC = 16 / sizeof(T);
  const int tpr = c / VEC;
  if (c % VEC == 0 && tpr >= 1 && tpr <= 256) {
    const int rpp = 256 / tpr;
    const size_t smem = (size_t)rpp * c * sizeof(float);
    if (smem <= 48 * 1024) {
      colsum_kernel<T><<<blocks, 256, smem, s>>>(x, rows, c, rpb, ws);
      return 0;
    }
  }
  colsum_scalar_kernel<T><<<blocks, 256, 0, s>>>(x, rows, c, rpb, ws);
  return 0;
}

extern "C" int vsr_colsum(const void* x, int32_t dtype, int64_t rows, int32_t c, float* db, int accumulate,
                          void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(x && db && rows > 0 && c > 0, "vsr_colsum: bad arguments");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_colsum_workspace(rows, c), "vsr_colsum: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = colsum_blocks(rows);
  const long rpb = (rows + blocks - 1) / blocks;
  float* ws = static_cast<float*>(workspace);
  if (dtype == VSR_F32)
    colsum_launch<float>((const float*)x, rows, c, ws, blocks, rpb, s);
  else if (dtype == VSR_BF16)
    colsum_launch<__nv_bfloat16>((const __nv_bfloat16*)x, rows, c, ws, blocks, rpb, s);
  else
    VSR_CHECK_ARG(false, "vsr_colsum: bad dtype %d", dtype);
  VSR_CHECK_LAUNCH("vsr_colsum");
  colsum_final_kernel<<<(c + 63) / 64, 256, 0, s>>>(ws, blocks, c, db, accumulate);
  VSR_CHECK_LAUNCH("vsr_colsum_final");
  return VSR_OK;
}

extern "C" int vsr_reduce_partials(const float* partials, int32_t rows, int32_t len, const int32_t* row_dst,
                                   float* dst, void* stream) {
  VSR_CHECK_ARG(partials && row_dst && dst && rows >= 0 && len > 0, "vsr_reduce_partials: bad arguments");
  VSR_CHECK_SUPPORTED(rows <= 8192, "vsr_reduce_partials: at most 8192 rows per call");
  if (rows == 0) return VSR_OK;
  reduce_partials_kernel<<<1, 1024, rows * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      partials, rows, len, row_dst, dst);
  VSR_CHECK_LAUNCH("vsr_reduce_partials");
  return VSR_OK;
}

extern "C" int vsr_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                             void* stream) {
  VSR_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "vsr_adam_flat: bad arguments");
  if (n == 0) return VSR_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
  VSR_CHECK_LAUNCH("vsr_adam_flat");
  return VSR_OK;
}

extern "C" int vsr_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                                 void* stream) {
  VSR_CHECK_ARG(p && g && m && v && hyper && n >= 0, "vsr_adam_flat_dev: bad arguments");
  if (n == 0) return VSR_OK;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) adam_dev_kernel<4><<<grid_for((n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, hyper);
  else adam_dev_kernel<1><<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, hyper);
  VSR_CHECK_LAUNCH("vsr_adam_flat_dev");
  return VSR_OK;
}

extern "C" int vsr_scale(float* x, int64_t n, float alpha, void* stream) {
  VSR_CHECK_ARG(x && n >= 0, "vsr_scale: bad arguments");
  if (n == 0) return VSR_OK;
  scale_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, alpha);
  VSR_CHECK_LAUNCH("vsr_scale");
  return VSR_OK;
}
