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
  const float a = kPrelu ? __ldg(slope_p) : 0.f;
  const float inv_a = a != 0.f ? 1.f / a : 0.f;
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float g = Elem<T>::ld(dy + i), v = Elem<T>::ld(y + i);
    const bool pos = v > 0.f;
    if (kPrelu) acc += pos ? 0.f : g * (v * inv_a);
    Elem<T>::st(dz + i, pos ? g : a * g);
  }
  if (kPrelu) {
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
  }
}

// column sums, pass 1: block b sums rows [b*rpb, (b+1)*rpb) for all c columns -> ws[b][c]
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long rows, int c, long rows_per_block,
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
__global__ void colsum_final_kernel(const float* __restrict__ ws, int blocks, int c,
                                    float* __restrict__ db, int accumulate) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= c) return;
  float s = accumulate ? db[col] : 0.f;
  for (int b = 0; b < blocks; ++b) s += ws[(size_t)b * c + col];
  db[col] = s;
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

extern "C" int vsr_add(const void* a, const void* b, void* out, int32_t dtype, int64_t numel, void* stream) {
  VSR_CHECK_ARG(a && b && out && numel >= 0, "vsr_add: bad arguments");
  if (numel == 0) return VSR_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == VSR_F32) {
    VSR_CHECK_ARG(numel % 4 == 0, "vsr_add: fp32 numel must be a multiple of 4");
    add_f32_kernel<<<grid_for(numel / 4, 256), 256, 0, s>>>((const float4*)a, (const float4*)b, (float4*)out, numel / 4);
  } else if (dtype == VSR_BF16) {
    VSR_CHECK_ARG(numel % 8 == 0, "vsr_add: bf16 numel must be a multiple of 8");
    add_bf16_kernel<<<grid_for(numel / 8, 256), 256, 0, s>>>((const uint4*)a, (const uint4*)b, (uint4*)out, numel / 8);
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
  if (dtype == VSR_F32) {
    if (slope) act_bwd_kernel<float, true><<<grid, 256, 0, s>>>((const float*)dy, (const float*)y, (float*)dz, numel, slope, slope_partials);
    else act_bwd_kernel<float, false><<<grid, 256, 0, s>>>((const float*)dy, (const float*)y, (float*)dz, numel, nullptr, nullptr);
  } else if (dtype == VSR_BF16) {
    using B = __nv_bfloat16;
    if (slope) act_bwd_kernel<B, true><<<grid, 256, 0, s>>>((const B*)dy, (const B*)y, (B*)dz, numel, slope, slope_partials);
    else act_bwd_kernel<B, false><<<grid, 256, 0, s>>>((const B*)dy, (const B*)y, (B*)dz, numel, nullptr, nullptr);
  } else {
    VSR_CHECK_ARG(false, "vsr_act_bwd: bad dtype %d", dtype);
  }
  VSR_CHECK_LAUNCH("vsr_act_bwd");
  return VSR_OK;
}

static int colsum_blocks(int64_t rows) {
  int64_t b = (rows + 63) / 64;
  const int64_t cap = (int64_t)num_sms() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" size_t vsr_colsum_workspace(int64_t rows, int32_t c) {
  return (size_t)colsum_blocks(rows) * (size_t)c * sizeof(float);
}

extern "C" int vsr_colsum(const void* x, int32_t dtype, int64_t rows, int32_t c, float* db, int accumulate,
                          void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(x && db && rows > 0 && c > 0, "vsr_colsum: bad arguments");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_colsum_workspace(rows, c), "vsr_colsum: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = colsum_blocks(rows);
  const long rpb = (rows + blocks - 1) / blocks;
  const int threads = c >= 256 ? 256 : ((c + 31) / 32) * 32;
  float* ws = static_cast<float*>(workspace);
  if (dtype == VSR_F32)
    colsum_kernel<float><<<blocks, threads, 0, s>>>((const float*)x, rows, c, rpb, ws);
  else if (dtype == VSR_BF16)
    colsum_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>((const __nv_bfloat16*)x, rows, c, rpb, ws);
  else
    VSR_CHECK_ARG(false, "vsr_colsum: bad dtype %d", dtype);
  VSR_CHECK_LAUNCH("vsr_colsum");
  colsum_final_kernel<<<(c + 127) / 128, 128, 0, s>>>(ws, blocks, c, db, accumulate);
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
