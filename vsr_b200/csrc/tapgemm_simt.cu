// tapgemm_simt.cu — CUDA-core tap-GEMM (fp32 accumulate, fp32 or bf16 storage): the strict-fp32
// mode of every convolution of the nets, and the weight-gradient of the tap-GEMM for both
// modes.  Same semantics as tapgemm_tc.cu (see include/vsr_b200.h); shapes are unrestricted.
#include "common.cuh"

namespace vsr {
namespace {

constexpr int kTM = 64;   // pixels per CTA tile
constexpr int kTN = 64;   // output channels per CTA tile
constexpr int kTK = 32;   // channels per smem step
constexpr int kThreads = 256;

template <typename T>
struct SimtArgs {
  const T* srcs[VSR_MAX_SRCS];
  int src_c[VSR_MAX_SRCS];
  const int4* tap_tab;
  const int4* group_tab;
  const T* w;
  const float* bias;
  const float* slope;
  const T* residual;
  const T* aux_y;
  T* out;
  T* out2;
  const T* res2;
  float* slope_partials;
  float out_scale;
  int epi;
  int kc, nt, n_groups;
  int N, H, W, Cout;
  int m_tiles, n_tiles;   // tiles over pixels / over nt
  long total_pix;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) tapgemm_simt_kernel(const __grid_constant__ SimtArgs<T> a) {
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Bs[kTK][kTN + 4];
  __shared__ int pix_n[kTM], pix_y[kTM], pix_x[kTM];
  __shared__ float red[32];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 threads, 4x4 outputs each
  const Prelu prelu = make_prelu((a.epi & (VSR_EPI_PRELU | VSR_EPI_PRELU_BWD)) ? __ldg(a.slope) : 1.f);
  float slope_acc = 0.f;

  const long tiles = (long)a.m_tiles * a.n_tiles * a.n_groups;
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int g = (int)(tile % a.n_groups);
    long rest = tile / a.n_groups;
    const int nb = (int)(rest % a.n_tiles);
    const long mb = rest / a.n_tiles;
    const int4 grp = __ldg(a.group_tab + g);
    const int j0 = nb * kTN;

    __syncthreads();
    if (tid < kTM) {
      const long p = mb * kTM + tid;
      if (p < a.total_pix) {
        const int x = (int)(p % a.W);
        const long q = p / a.W;
        pix_x[tid] = x;
        pix_y[tid] = (int)(q % a.H);
        pix_n[tid] = (int)(q / a.H);
      } else {
        pix_n[tid] = -1; pix_y[tid] = 0; pix_x[tid] = 0;
      }
    }
    __syncthreads();

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < grp.z; ++t) {
      const int4 tap = __ldg(a.tap_tab + grp.y + t);
      const T* src = a.srcs[tap.x];
      const int sc = a.src_c[tap.x];
      const T* wslab = a.w + (size_t)(grp.y + t) * a.nt * a.kc;
      for (int k0 = 0; k0 < a.kc; k0 += kTK) {
        // A: 64 pixels x 32 channels, 8 per thread
        {
          const int p = tid >> 2, kk = (tid & 3) * 8;
          const int n = pix_n[p];
          const int y = pix_y[p] + tap.y, x = pix_x[p] + tap.z;
          const bool ok = n >= 0 && y >= 0 && y < a.H && x >= 0 && x < a.W;
          const T* sp = src + (((size_t)n * a.H + y) * a.W + x) * sc + tap.w + k0 + kk;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v = 0.f;
            if (ok && (k0 + kk + i) < a.kc) v = Elem<T>::ld(sp + i);
            As[kk + i][p] = v;
          }
        }
        // B: 64 out channels x 32 k
        {
          const int j = tid >> 2, kk = (tid & 3) * 8;
          const bool ok = (j0 + j) < a.nt;
          const T* wp = wslab + (size_t)(j0 + j) * a.kc + k0 + kk;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v = 0.f;
            if (ok && (k0 + kk + i) < a.kc) v = Elem<T>::ld(wp + i);
            Bs[kk + i][j] = v;
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTK; ++k) {
          const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          const float aa[4] = {av.x, av.y, av.z, av.w};
          const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
      }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = ty * 4 + i;
      const int n = pix_n[p];
      if (n < 0) continue;
      const size_t rowoff = (((size_t)n * a.H + pix_y[p]) * a.W + pix_x[p]) * (size_t)a.Cout + grp.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = j0 + tx * 4 + j;
        if (jj >= a.nt) continue;
        float v = acc[i][j];
        if (a.epi & VSR_EPI_BIAS) v += __ldg(a.bias + grp.x + jj);
        if (a.epi & VSR_EPI_SCALE) v *= a.out_scale;
        if (a.epi & VSR_EPI_RES_PRE) v += Elem<T>::ld(a.residual + rowoff + jj);
        if (a.epi & VSR_EPI_RELU) v = fmaxf(v, 0.f);
        if (a.epi & VSR_EPI_PRELU_BWD) {
          slope_acc += prelu_bwd(a.aux_y + rowoff + jj, v, prelu);
        } else if (a.epi & VSR_EPI_RELU_BWD) {
          const float y = Elem<T>::ld(a.aux_y + rowoff + jj);
          v = y > 0.f ? v : 0.f;
        }
        if (a.epi & VSR_EPI_PRELU) {
          prelu_store(a.out + rowoff + jj, v, prelu);      // (the branch of a negative slope travels in y's LSB)
          v = v > 0.f ? v : prelu.fwd * v;
        } else {
          Elem<T>::st(a.out + rowoff + jj, v);
        }
        if (a.epi & VSR_EPI_OUT2) {
          const float r2 = Elem<T>::ld(a.res2 + rowoff + jj);
          Elem<T>::st(a.out2 + rowoff + jj, (a.epi & VSR_EPI_OUT2_SUB) ? v - r2 : v + r2);
        }
      }
    }
  }
  if (a.epi & VSR_EPI_PRELU_BWD) {
    const float s = block_sum(slope_acc, red);
    if (tid == 0) a.slope_partials[blockIdx.x] = s;
  }
}

template <typename T>
int launch_simt(const VsrTapGemmDesc* d, cudaStream_t stream) {
  SimtArgs<T> a;
  memset(&a, 0, sizeof(a));
  for (int s = 0; s < d->n_srcs; ++s) {
    a.srcs[s] = static_cast<const T*>(d->srcs[s].ptr);
    a.src_c[s] = d->srcs[s].c;
  }
  a.tap_tab = reinterpret_cast<const int4*>(d->tap_tab);
  a.group_tab = reinterpret_cast<const int4*>(d->group_tab);
  a.w = static_cast<const T*>(d->w);
  a.bias = d->bias;
  a.slope = d->slope;
  a.residual = static_cast<const T*>(d->residual);
  a.aux_y = static_cast<const T*>(d->aux_y);
  a.out = static_cast<T*>(d->out.ptr);
  a.out2 = static_cast<T*>(d->out2);
  a.res2 = static_cast<const T*>(d->res2);
  a.slope_partials = d->slope_partials;
  a.out_scale = d->out_scale;
  a.epi = d->epi;
  a.kc = d->kc; a.nt = d->nt; a.n_groups = d->n_groups;
  a.N = d->out.n; a.H = d->out.h; a.W = d->out.w; a.Cout = d->out.c;
  a.total_pix = (long)a.N * a.H * a.W;
  a.m_tiles = (int)((a.total_pix + kTM - 1) / kTM);
  a.n_tiles = (a.nt + kTN - 1) / kTN;
  const long tiles = (long)a.m_tiles * a.n_tiles * a.n_groups;
  long grid = tiles;
  const long cap = (a.epi & VSR_EPI_PRELU_BWD) ? kPartialsLen : (long)num_sms() * 16;
  if (grid > cap) grid = cap;
  if (grid > kPartialsLen && (a.epi & VSR_EPI_PRELU_BWD)) grid = kPartialsLen;
  tapgemm_simt_kernel<T><<<(int)grid, kThreads, 0, stream>>>(a);
  VSR_CHECK_LAUNCH("tapgemm_simt");
  return VSR_OK;
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dw[t][j][k] = sum_pix dz[pix][o0+j] * src[pix+off][c0+k]
// CTA = (tap, 64 j, 64 k) x one pixel split; partials to workspace, fixed-order reduce.
// ---------------------------------------------------------------------------------------------
constexpr int kWP = 16;  // pixels per smem step

template <typename T>
struct WgradArgs {
  const T* srcs[VSR_MAX_SRCS];
  int src_c[VSR_MAX_SRCS];
  const int4* tap_tab;
  const int4* group_tab;
  const T* dz;
  float* ws;          // [splits][n_taps_total*nt*kc]
  int kc, nt, n_groups, n_taps_total;
  int N, H, W, Cout;
  int j_tiles, k_tiles, splits;
  long total_pix, pix_per_split;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) tapgemm_wgrad_kernel(const __grid_constant__ WgradArgs<T> a) {
  __shared__ float Zs[kWP][kTN + 4];   // dz   [pix][j]
  __shared__ float Xs[kWP][kTN + 4];   // src  [pix][k]
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;   // ty -> j, tx -> k

  int tile = blockIdx.x;
  const int kb = tile % a.k_tiles; tile /= a.k_tiles;
  const int jb = tile % a.j_tiles; tile /= a.j_tiles;
  const int tapi = tile;                 // global tap index
  const int split = blockIdx.y;

  // find the group of this tap (few groups: linear scan)
  int o0 = 0;
  for (int g = 0; g < a.n_groups; ++g) {
    const int4 grp = __ldg(a.group_tab + g);
    if (tapi >= grp.y && tapi < grp.y + grp.z) { o0 = grp.x; break; }
  }
  const int4 tap = __ldg(a.tap_tab + tapi);
  const T* src = a.srcs[tap.x];
  const int sc = a.src_c[tap.x];
  const int j0 = jb * kTN, k0 = kb * kTN;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const long p_begin = (long)split * a.pix_per_split;
  long p_end = p_begin + a.pix_per_split;
  if (p_end > a.total_pix) p_end = a.total_pix;

  for (long pb = p_begin; pb < p_end; pb += kWP) {
    // each thread loads 4 dz + 4 src elements: pixel = tid/16, 4 channels at (tid%16)*4
    {
      const int pl = tid >> 4, c4 = (tid & 15) * 4;
      const long p = pb + pl;
      const bool pin = p < p_end;
      int x = 0, y = 0, n = 0;
      if (pin) {
        x = (int)(p % a.W);
        const long q = p / a.W;
        y = (int)(q % a.H);
        n = (int)(q / a.H);
      }
      const T* zp = a.dz + (size_t)p * a.Cout + o0 + j0 + c4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (pin && (j0 + c4 + i) < a.nt) v = Elem<T>::ld(zp + i);
        Zs[pl][c4 + i] = v;
      }
      const int sy = y + tap.y, sx = x + tap.z;
      const bool ok = pin && sy >= 0 && sy < a.H && sx >= 0 && sx < a.W;
      const T* sp = src + (((size_t)n * a.H + sy) * a.W + sx) * sc + tap.w + k0 + c4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (ok && (k0 + c4 + i) < a.kc) v = Elem<T>::ld(sp + i);
        Xs[pl][c4 + i] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < kWP; ++p) {
      const float4 zv = *reinterpret_cast<const float4*>(&Zs[p][ty * 4]);
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[p][tx * 4]);
      const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
      const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zz[i], xx[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* wsp = a.ws + (size_t)split * a.n_taps_total * a.nt * a.kc + (size_t)tapi * a.nt * a.kc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + ty * 4 + i;
    if (j >= a.nt) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int k = k0 + tx * 4 + jj;
      if (k < a.kc) wsp[(size_t)j * a.kc + k] = acc[i][jj];
    }
  }
}

__global__ void reduce_splits_kernel(const float* __restrict__ ws, float* __restrict__ dw, long n,
                                     int splits, int accumulate) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float s = accumulate ? dw[i] : 0.f;
    for (int k = 0; k < splits; ++k) s += ws[(size_t)k * n + i];
    dw[i] = s;
  }
}

int wgrad_splits(const VsrTapGemmDesc* d) {
  const long total_pix = (long)d->out.n * d->out.h * d->out.w;
  const long tiles = (long)d->n_taps_total * ((d->nt + kTN - 1) / kTN) * ((d->kc + kTN - 1) / kTN);
  long s = (4l * 148 + tiles - 1) / tiles;
  const long max_s = (total_pix + 255) / 256;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return (int)s;
}

template <typename T>
int launch_wgrad(const VsrTapGemmDesc* d, float* dw, int accumulate, void* workspace,
                 cudaStream_t stream) {
  WgradArgs<T> a;
  memset(&a, 0, sizeof(a));
  for (int s = 0; s < d->n_srcs; ++s) {
    a.srcs[s] = static_cast<const T*>(d->srcs[s].ptr);
    a.src_c[s] = d->srcs[s].c;
  }
  a.tap_tab = reinterpret_cast<const int4*>(d->tap_tab);
  a.group_tab = reinterpret_cast<const int4*>(d->group_tab);
  a.dz = static_cast<const T*>(d->out.ptr);
  a.ws = static_cast<float*>(workspace);
  a.kc = d->kc; a.nt = d->nt; a.n_groups = d->n_groups; a.n_taps_total = d->n_taps_total;
  a.N = d->out.n; a.H = d->out.h; a.W = d->out.w; a.Cout = d->out.c;
  a.j_tiles = (a.nt + kTN - 1) / kTN;
  a.k_tiles = (a.kc + kTN - 1) / kTN;
  a.splits = wgrad_splits(d);
  a.total_pix = (long)a.N * a.H * a.W;
  a.pix_per_split = (a.total_pix + a.splits - 1) / a.splits;
  dim3 grid(a.n_taps_total * a.j_tiles * a.k_tiles, a.splits);
  tapgemm_wgrad_kernel<T><<<grid, kThreads, 0, stream>>>(a);
  VSR_CHECK_LAUNCH("tapgemm_wgrad");
  const long n = (long)a.n_taps_total * a.nt * a.kc;
  reduce_splits_kernel<<<grid_for(n, 256), 256, 0, stream>>>(a.ws, dw, n, a.splits, accumulate);
  VSR_CHECK_LAUNCH("tapgemm_wgrad_reduce");
  return VSR_OK;
}

}  // namespace

int tapgemm_tc2_launch(const VsrTapGemmDesc* d, cudaStream_t stream);  // tapgemm_tc2.cu
bool wgrad_tc_supported(const VsrTapGemmDesc* d);                       // wgrad_tc.cu
size_t wgrad_tc_workspace(const VsrTapGemmDesc* d);
int wgrad_tc_launch(const VsrTapGemmDesc* d, float* dw, float* db, int db_period, int accumulate, void* workspace,
                    cudaStream_t stream);
int wgrad_tc_partial(const VsrTapGemmDesc* d, int want_bias, int slice, int n_slices, void* workspace, cudaStream_t stream);
int wgrad_tc_finish(const VsrTapGemmDesc* d, float* dw, float* db, int db_period, int accumulate, int used_slices,
                    int n_slices, void* workspace, cudaStream_t stream);
bool wgrad_tc_bias_ok(const VsrTapGemmDesc* d, int db_period);

int validate_desc(const VsrTapGemmDesc* d, const char* who) {
  VSR_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
  VSR_CHECK_ARG(d->dtype == VSR_F32 || d->dtype == VSR_BF16 || d->dtype == VSR_BF16X2, "%s: bad dtype %d", who, d->dtype);
  VSR_CHECK_ARG(d->n_srcs >= 1 && d->n_srcs <= VSR_MAX_SRCS, "%s: n_srcs=%d out of range", who, d->n_srcs);
  VSR_CHECK_ARG(d->kc >= 1 && d->nt >= 1, "%s: kc/nt must be positive", who);
  VSR_CHECK_ARG(d->n_groups >= 1 && d->n_taps_total >= 1, "%s: empty group/tap table", who);
  VSR_CHECK_ARG(d->group_tab && d->tap_tab, "%s: null tap/group table", who);
  VSR_CHECK_ARG(d->out.ptr && d->out.n > 0 && d->out.h > 0 && d->out.w > 0 && d->out.c > 0,
                "%s: bad output tensor", who);
  for (int s = 0; s < d->n_srcs; ++s) {
    VSR_CHECK_ARG(d->srcs[s].ptr != nullptr, "%s: src %d is null", who, s);
    VSR_CHECK_ARG(d->srcs[s].n == d->out.n && d->srcs[s].h == d->out.h && d->srcs[s].w == d->out.w,
                  "%s: src %d pixel grid [%d,%d,%d] differs from out [%d,%d,%d]", who, s, d->srcs[s].n,
                  d->srcs[s].h, d->srcs[s].w, d->out.n, d->out.h, d->out.w);
  }
  return VSR_OK;
}

}  // namespace vsr

extern "C" int vsr_tapgemm(const VsrTapGemmDesc* d, void* stream) {
  using namespace vsr;
  int rc = validate_desc(d, "vsr_tapgemm");
  if (rc != VSR_OK) return rc;
  VSR_CHECK_ARG(d->w != nullptr, "vsr_tapgemm: null weights");
  if (d->epi & VSR_EPI_BIAS) VSR_CHECK_ARG(d->bias, "vsr_tapgemm: BIAS without bias");
  if (d->epi & VSR_EPI_RES_PRE) VSR_CHECK_ARG(d->residual, "vsr_tapgemm: RES_PRE without residual");
  if (d->epi & (VSR_EPI_PRELU | VSR_EPI_PRELU_BWD)) VSR_CHECK_ARG(d->slope, "vsr_tapgemm: PReLU without slope");
  if (d->epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) VSR_CHECK_ARG(d->aux_y, "vsr_tapgemm: *_BWD without aux_y");
  if (d->epi & VSR_EPI_PRELU_BWD) VSR_CHECK_ARG(d->slope_partials, "vsr_tapgemm: PRELU_BWD without slope_partials");
  if (d->epi & VSR_EPI_OUT2) VSR_CHECK_ARG(d->out2 && d->res2, "vsr_tapgemm: OUT2 without out2/res2");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d->dtype == VSR_BF16 || d->dtype == VSR_BF16X2) return tapgemm_tc2_launch(d, s);
  return launch_simt<float>(d, s);
}

// test hook: the CUDA-core kernel on bf16 storage (used to cross-check the tcgen05 path; the
// weights here are plain row-major bf16 slabs, not the swizzled image).
extern "C" int vsr_tapgemm_simt_bf16(const VsrTapGemmDesc* d, void* stream) {
  using namespace vsr;
  int rc = validate_desc(d, "vsr_tapgemm_simt_bf16");
  if (rc != VSR_OK) return rc;
  return launch_simt<__nv_bfloat16>(d, static_cast<cudaStream_t>(stream));
}

extern "C" size_t vsr_tapgemm_wgrad_workspace(const VsrTapGemmDesc* d) {
  if (!d) return 0;
  if (vsr::wgrad_tc_supported(d)) return vsr::wgrad_tc_workspace(d);
  return (size_t)vsr::wgrad_splits(d) * d->n_taps_total * d->nt * d->kc * sizeof(float);
}

extern "C" int vsr_tapgemm_wgrad(const VsrTapGemmDesc* d, float* dw, int accumulate, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  using namespace vsr;
  int rc = validate_desc(d, "vsr_tapgemm_wgrad");
  if (rc != VSR_OK) return rc;
  VSR_CHECK_ARG(dw != nullptr, "vsr_tapgemm_wgrad: null dw");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_tapgemm_wgrad_workspace(d),
                "vsr_tapgemm_wgrad: workspace too small (%zu < %zu)", workspace_bytes,
                vsr_tapgemm_wgrad_workspace(d));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (wgrad_tc_supported(d)) return wgrad_tc_launch(d, dw, nullptr, 0, accumulate, workspace, s);
  VSR_CHECK_SUPPORTED(d->dtype != VSR_BF16X2, "vsr_tapgemm_wgrad: VSR_BF16X2 is a forward / data-gradient mode (weight gradients take the planes one by one as VSR_BF16)");
  if (d->dtype == VSR_BF16) return launch_wgrad<__nv_bfloat16>(d, dw, accumulate, workspace, s);
  return launch_wgrad<float>(d, dw, accumulate, workspace, s);
}

// test hook: the CUDA-core weight gradient on bf16 storage (cross-check for the tcgen05 path)
extern "C" size_t vsr_tapgemm_wgrad_simt_workspace(const VsrTapGemmDesc* d) {
  if (!d) return 0;
  return (size_t)vsr::wgrad_splits(d) * d->n_taps_total * d->nt * d->kc * sizeof(float);
}
extern "C" int vsr_tapgemm_wgrad_simt(const VsrTapGemmDesc* d, float* dw, int accumulate, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  using namespace vsr;
  int rc = validate_desc(d, "vsr_tapgemm_wgrad_simt");
  if (rc != VSR_OK) return rc;
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_tapgemm_wgrad_simt_workspace(d), "vsr_tapgemm_wgrad_simt: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d->dtype == VSR_BF16) return launch_wgrad<__nv_bfloat16>(d, dw, accumulate, workspace, s);
  return launch_wgrad<float>(d, dw, accumulate, workspace, s);
}

// weight gradient + bias gradient (db[q] (+)= sum over pixels and channels c = q mod period of dz).
// Returns 1 if the bias gradient was fused, 0 if the caller still has to run vsr_colsum.
extern "C" int vsr_tapgemm_wgrad_bias(const VsrTapGemmDesc* d, float* dw, float* db, int32_t db_period,
                                      int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace vsr;
  int rc = validate_desc(d, "vsr_tapgemm_wgrad_bias");
  if (rc != VSR_OK) return rc;
  VSR_CHECK_ARG(dw != nullptr && db != nullptr && db_period > 0, "vsr_tapgemm_wgrad_bias: bad arguments");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_tapgemm_wgrad_workspace(d),
                "vsr_tapgemm_wgrad_bias: workspace too small (%zu < %zu)", workspace_bytes,
                vsr_tapgemm_wgrad_workspace(d));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (wgrad_tc_supported(d) && d->out.c <= 1024 && d->out.c % db_period == 0) {
    rc = wgrad_tc_launch(d, dw, db, db_period, accumulate, workspace, s);
    return rc == VSR_OK ? 1 : rc;
  }
  rc = vsr_tapgemm_wgrad(d, dw, accumulate, workspace, workspace_bytes, stream);
  return rc;
}

// Deferred reduction: the same layer is differentiated once per frame with identical shapes, so every
// frame writes its per-split partial sums into its own slice of a per-layer workspace
// (vsr_tapgemm_wgrad_partial, slice < n_slices; workspace >= n_slices * vsr_tapgemm_wgrad_workspace)
// and one fixed-order pass reduces all slices (vsr_tapgemm_wgrad_finish).  partial returns 1 if the
// tcgen05 kernel took it (bias included), 0 if the shape is unsupported (use vsr_tapgemm_wgrad).
extern "C" int vsr_tapgemm_wgrad_partial(const VsrTapGemmDesc* d, int32_t db_period, int32_t slice, int32_t n_slices,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  using namespace vsr;
  int rc = validate_desc(d, "vsr_tapgemm_wgrad_partial");
  if (rc != VSR_OK) return rc;
  if (!(wgrad_tc_supported(d) && wgrad_tc_bias_ok(d, db_period))) return 0;
  VSR_CHECK_ARG(slice >= 0 && slice < n_slices, "vsr_tapgemm_wgrad_partial: slice out of range");
  VSR_CHECK_ARG(workspace && workspace_bytes >= (size_t)n_slices * vsr_tapgemm_wgrad_workspace(d),
                "vsr_tapgemm_wgrad_partial: workspace too small");
  rc = wgrad_tc_partial(d, 1, slice, n_slices, workspace, static_cast<cudaStream_t>(stream));
  return rc == VSR_OK ? 1 : rc;
}

extern "C" int vsr_tapgemm_wgrad_finish(const VsrTapGemmDesc* d, float* dw, float* db, int32_t db_period,
                                        int accumulate, int32_t used_slices, int32_t n_slices, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  using namespace vsr;
  VSR_CHECK_ARG(d && dw && db && workspace, "vsr_tapgemm_wgrad_finish: bad arguments");
  VSR_CHECK_ARG(wgrad_tc_supported(d) && wgrad_tc_bias_ok(d, db_period), "vsr_tapgemm_wgrad_finish: shape was not accepted by _partial");
  VSR_CHECK_ARG(used_slices >= 1 && used_slices <= n_slices, "vsr_tapgemm_wgrad_finish: bad slice count");
  VSR_CHECK_ARG(workspace_bytes >= (size_t)n_slices * vsr_tapgemm_wgrad_workspace(d), "vsr_tapgemm_wgrad_finish: workspace too small");
  return wgrad_tc_finish(d, dw, db, db_period, accumulate, used_slices, n_slices, workspace, static_cast<cudaStream_t>(stream));
}
