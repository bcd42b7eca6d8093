// wgrad_shared.cu — weight gradients of SEVERAL 1x1 convolutions that read the same feature list, in one pass over the
// maps (tcgen05, sm_100a).
//
// The dense connections of the feedback block (drf_net.py:89-105: group g's 1x1 convolution reads the concatenation of the
// feature maps 0..g) make the weight gradients of layers g = 1..G-1 read map j once per layer g >= j: at config 2 the
// down-projection 1x1 layers read 20 high-resolution maps (335 MB each, all frames stacked) plus their 5 gradient maps,
// five HBM-bound launches at the copy bandwidth.  Here one CTA loads, per 64-pixel tile, every source tile ONCE and the
// gradient tiles of all layers of the launch, and issues one MMA set per (layer, source pair):
//     dw_g[t][j][k] = sum_pix dz_g[pix][j] * src_t[pix][k],   t < ntaps_g
// accumulated in TMEM over the CTA's pixel range (<= 8 accumulators of [2 taps x 64] x 64), fp32 partials per pixel split,
// fixed-order reduction (deterministic).  The column sums of every dz tile (bias gradients) ride along as in wgrad_tc.cu.
// Operands are MN-major like there: a TMA box [64 pixels x 64 channels] (128-byte rows, 128B swizzle) is a K = 64 x MN = 64
// tile; two neighbouring source tiles form the M = 128 A operand.
//
// Replaces the weight half of aten.convolution_backward for f_block.{up,down}_blocks[g].conv1 (drf_net.py:90,97).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vsr {

int get_src_map_pub(const VsrTensor4& t, int bw, int bh, CUtensorMap* out);   // tma_host.cu

namespace {

constexpr int kPx = 64;                  // pixels per tile
constexpr int kTileB = kPx * 128;        // 8 KiB: [64 px][64 ch] bf16
constexpr int kMaxSrc = VSR_WS_MAX_SRCS, kMaxDz = VSR_WS_MAX_DZ, kMaxAcc = 8;
constexpr int kCtrl = 1024;
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;            // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: column sums + epilogue
constexpr int kTmemCols = 512;

struct SharedArgs {
  CUtensorMap src_maps[kMaxSrc];
  CUtensorMap dz_maps[kMaxDz];
  float* ws;            // [splits][total_taps][64][64]
  float* bias_ws;       // [splits][n_dz][64] (or NULL)
  int n_src, n_dz, total_taps, stages;
  int ntaps[kMaxDz], tap_base[kMaxDz], acc_base[kMaxDz];
  int N, H, W, bw, bh, tiles_x, tiles_y, num_ptiles, splits;
};

__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo) { return ptx::make_sw128_desc(addr, lbo, 1024); }

__global__ void __launch_bounds__(kThreads, 1) wgrad_shared_kernel(const __grid_constant__ SharedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t full_bar = base, empty_bar = base + 64, done_bar = base + 128, tmem_slot = base + 136;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 136);
  const int n_tiles = a.n_src + a.n_dz;
  const uint32_t stage_bytes = static_cast<uint32_t>(n_tiles) * kTileB;
  const uint32_t stage_base = base + kCtrl;
  const bool do_cs = a.bias_ws != nullptr;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int per = (a.num_ptiles + a.splits - 1) / a.splits;
  const int pt0 = split * per, pt1 = min(pt0 + per, a.num_ptiles);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, do_cs ? 5 : 1);      // MMA commit (+ the 4 column-sum warps)
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::pdl_wait();            // nothing above touches tensors written by earlier kernels
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ---- TMA producer: per pixel tile every source tile once, then the gradient tile of every layer ----
    const bool leader = ptx::elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = pt0; pt < pt1; ++pt) {
      int q = pt;
      const int tx = q % a.tiles_x;
      q /= a.tiles_x;
      const int ty = q % a.tiles_y, n = q / a.tiles_y;
      const int x0 = tx * a.bw, y0 = ty * a.bh;
      ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1u);
      if (leader) {
        const uint32_t fb = full_bar + 8 * stage, sa = stage_base + stage * stage_bytes;
        ptx::mbar_arrive_expect_tx(fb, stage_bytes);
        for (int i = 0; i < a.n_src; ++i) ptx::tma_load_4d(sa + i * kTileB, &a.src_maps[i], fb, 0, x0, y0, n);
        for (int g = 0; g < a.n_dz; ++g) ptx::tma_load_4d(sa + (a.n_src + g) * kTileB, &a.dz_maps[g], fb, 0, x0, y0, n);
      }
      if (++stage == a.stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: (layer g, source pair p) -> accumulator acc_base[g] + p, summed over the CTA's pixel tiles ----
    const bool leader = ptx::elect_one();
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 1, 1);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int pt = pt0; pt < pt1; ++pt, ++it) {
      ptx::mbar_wait(full_bar + 8 * stage, phase);
      ptx::tc_fence_after();
      const uint32_t sa = stage_base + stage * stage_bytes;
      if (leader) {
        for (int g = 0; g < a.n_dz; ++g) {
          const uint32_t sb = sa + (a.n_src + g) * kTileB;
          const int pairs = (a.ntaps[g] + 1) >> 1;
          for (int p = 0; p < pairs; ++p) {
            // (an odd tap count pairs the last source with the tile that follows it in the stage: its 64 accumulator rows
            //  are never read)
#pragma unroll
            for (int k = 0; k < kPx / 16; ++k)
              ptx::mma_bf16_ss(tmem_u + (a.acc_base[g] + p) * 64, mn_desc(sa + 2 * p * kTileB + k * 2048, kTileB),
                               mn_desc(sb + k * 2048, kTileB), idesc, (it | k) != 0);
          }
        }
        ptx::mma_commit(empty_bar + 8 * stage);
      }
      if (++stage == a.stages) { stage = 0; phase ^= 1u; }
    }
    if (leader) ptx::mma_commit(done_bar);
  } else {
    // ---- column sums of the gradient tiles while the MMAs run, then the accumulators -> fp32 partials ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;          // TMEM lane: (tap parity, k)
    const int cg = row >> 5, cp = row & 31;       // column-sum duty: gradient map cg, channels 2 cp, 2 cp + 1
    float cs0 = 0.f, cs1 = 0.f;
    if (do_cs) {
      int stage = 0;
      uint32_t phase = 0;
      const int cc = 2 * cp;
      for (int pt = pt0; pt < pt1; ++pt) {
        ptx::mbar_wait(full_bar + 8 * stage, phase);
        if (cg < a.n_dz) {
          const uint32_t tb = stage_base + stage * stage_bytes + (a.n_src + cg) * kTileB;
#pragma unroll 8
          for (int r = 0; r < kPx; ++r) {
            uint32_t u;
            const uint32_t addr = tb + r * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(addr) : "memory");
            cs0 += bf16_lo(u);
            cs1 += bf16_hi(u);
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(empty_bar + 8 * stage);
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      }
    }
    if (pt1 > pt0) {
      ptx::mbar_wait(done_bar, 0);
      ptx::tc_fence_after();
    }
    const int half = row >> 6, k = row & 63;
    float* wsp = a.ws + (size_t)split * a.total_taps * 4096;
    for (int g = 0; g < a.n_dz; ++g) {
      const int pairs = (a.ntaps[g] + 1) >> 1;
      for (int p = 0; p < pairs; ++p) {
        const int t = 2 * p + half;
        const bool live = t < a.ntaps[g];
        const uint32_t taddr = tmem_base + (a.acc_base[g] + p) * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
        for (int c = 0; c < 64; c += 16) {
          uint32_t r[16];
          if (pt1 > pt0) {
            ptx::tmem_ld16(taddr + c, r);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = 0u;        // empty split: contributes zeros
          }
          if (live) {
            float* o = wsp + ((size_t)(a.tap_base[g] + t) * 64 + c) * 64 + k;
#pragma unroll
            for (int i = 0; i < 16; ++i) o[(size_t)i * 64] = __uint_as_float(r[i]);
          }
        }
      }
    }
    if (do_cs && cg < a.n_dz) {
      float* o = a.bias_ws + ((size_t)split * a.n_dz + cg) * 64 + 2 * cp;
      o[0] = cs0;
      o[1] = cs1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

struct ReduceArgs {
  const float* ws;
  const float* bias_ws;
  float* dw[kMaxDz];
  float* db[kMaxDz];
  int tap_base[kMaxDz], ntaps[kMaxDz];
  int n_dz, total_taps, splits, accumulate;
};

// dw_g[i] (+)= sum over the splits, in order; the last block reduces the bias partials
__global__ void __launch_bounds__(256) wgrad_shared_reduce_kernel(const ReduceArgs a) {
  const long n4 = (long)a.total_taps * 1024;
  if (blockIdx.x == gridDim.x - 1) {
    if (a.bias_ws != nullptr)
      for (int i = threadIdx.x; i < a.n_dz * 64; i += blockDim.x) {
        const int g = i >> 6, c = i & 63;
        if (a.db[g] == nullptr) continue;
        float s = a.accumulate ? a.db[g][c] : 0.f;
        for (int k = 0; k < a.splits; ++k) s += a.bias_ws[((size_t)k * a.n_dz + g) * 64 + c];
        a.db[g][c] = s;
      }
    return;
  }
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int tap = (int)(i >> 10);
  int g = 0;
  while (g + 1 < a.n_dz && tap >= a.tap_base[g + 1]) ++g;
  float4* dst = reinterpret_cast<float4*>(a.dw[g]) + (i - (long)a.tap_base[g] * 1024);
  const float4* src = reinterpret_cast<const float4*>(a.ws) + i;
  float4 s = a.accumulate ? *dst : make_float4(0.f, 0.f, 0.f, 0.f);
  int k = 0;
  for (; k + 4 <= a.splits; k += 4) {          // four independent 16-byte loads in flight, added in split order
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (size_t)(k + u) * n4);
#pragma unroll
    for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
  }
  for (; k < a.splits; ++k) {
    const float4 v = __ldg(src + (size_t)k * n4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *dst = s;
}

struct Plan {
  int bw, bh, tiles_x, tiles_y, num_ptiles, splits, stages, total_taps, n_acc;
};

int make_plan(const VsrWgradSharedDesc* d, Plan* p) {
  VSR_CHECK_ARG(d && d->n_srcs >= 1 && d->n_srcs <= kMaxSrc && d->n_dz >= 1 && d->n_dz <= kMaxDz, "vsr_wgrad_shared: bad source / gradient counts");
  const VsrTensor4& z = d->dzs[0];
  VSR_CHECK_ARG(z.n > 0 && z.h > 0 && z.w > 0, "vsr_wgrad_shared: bad pixel grid");
  p->total_taps = p->n_acc = 0;
  for (int g = 0; g < d->n_dz; ++g) {
    VSR_CHECK_ARG(d->ntaps[g] >= 1 && d->ntaps[g] <= d->n_srcs, "vsr_wgrad_shared: ntaps[%d] = %d out of range", g, d->ntaps[g]);
    VSR_CHECK_ARG(d->dzs[g].ptr && d->dzs[g].c == 64 && d->dzs[g].n == z.n && d->dzs[g].h == z.h && d->dzs[g].w == z.w && d->dw[g],
                  "vsr_wgrad_shared: gradient map %d must be [n,h,w,64] on the common pixel grid", g);
    p->total_taps += d->ntaps[g];
    p->n_acc += (d->ntaps[g] + 1) / 2;
  }
  for (int i = 0; i < d->n_srcs; ++i)
    VSR_CHECK_ARG(d->srcs[i].ptr && d->srcs[i].c == 64 && d->srcs[i].n == z.n && d->srcs[i].h == z.h && d->srcs[i].w == z.w,
                  "vsr_wgrad_shared: source map %d must be [n,h,w,64] on the common pixel grid", i);
  VSR_CHECK_SUPPORTED(p->n_acc <= kMaxAcc, "vsr_wgrad_shared: %d accumulators (> %d): split the layer set", p->n_acc, kMaxAcc);
  int bw = 1;
  while (bw * 2 <= z.w && bw * 2 <= kPx) bw *= 2;
  p->bw = bw;
  p->bh = kPx / bw;
  p->tiles_x = (z.w + p->bw - 1) / p->bw;
  p->tiles_y = (z.h + p->bh - 1) / p->bh;
  p->num_ptiles = z.n * p->tiles_x * p->tiles_y;
  p->splits = num_sms() < p->num_ptiles ? num_sms() : p->num_ptiles;
  p->stages = (227 * 1024 - kCtrl - 1024) / ((d->n_srcs + d->n_dz) * kTileB);
  if (p->stages > kMaxStages) p->stages = kMaxStages;
  VSR_CHECK_SUPPORTED(p->stages >= 2, "vsr_wgrad_shared: %d tiles per stage do not fit twice in shared memory", d->n_srcs + d->n_dz);
  return VSR_OK;
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" size_t vsr_wgrad_shared_workspace(const VsrWgradSharedDesc* d) {
  Plan p;
  if (make_plan(d, &p) != VSR_OK) return 0;
  return (size_t)p.splits * ((size_t)p.total_taps * 4096 + (size_t)d->n_dz * 64) * sizeof(float);
}

extern "C" int vsr_wgrad_shared(const VsrWgradSharedDesc* d, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  Plan p;
  int rc = make_plan(d, &p);
  if (rc != VSR_OK) return rc;
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_wgrad_shared_workspace(d), "vsr_wgrad_shared: workspace too small");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(wgrad_shared smem) failed: %s", cudaGetErrorString(e));
      return VSR_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SharedArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < d->n_srcs; ++i)
    if ((rc = get_src_map_pub(d->srcs[i], p.bw, p.bh, &a.src_maps[i])) != VSR_OK) return rc;
  for (int g = 0; g < d->n_dz; ++g)
    if ((rc = get_src_map_pub(d->dzs[g], p.bw, p.bh, &a.dz_maps[g])) != VSR_OK) return rc;
  bool any_db = false;
  for (int g = 0; g < d->n_dz; ++g) any_db = any_db || d->db[g] != nullptr;
  float* ws = static_cast<float*>(workspace);
  a.ws = ws;
  a.bias_ws = any_db ? ws + (size_t)p.splits * p.total_taps * 4096 : nullptr;
  a.n_src = d->n_srcs; a.n_dz = d->n_dz; a.total_taps = p.total_taps; a.stages = p.stages;
  ReduceArgs r;
  memset(&r, 0, sizeof(r));
  int tb = 0, ab = 0;
  for (int g = 0; g < d->n_dz; ++g) {
    a.ntaps[g] = r.ntaps[g] = d->ntaps[g];
    a.tap_base[g] = r.tap_base[g] = tb;
    a.acc_base[g] = ab;
    tb += d->ntaps[g];
    ab += (d->ntaps[g] + 1) / 2;
    r.dw[g] = d->dw[g];
    r.db[g] = d->db[g];
  }
  a.N = d->dzs[0].n; a.H = d->dzs[0].h; a.W = d->dzs[0].w;
  a.bw = p.bw; a.bh = p.bh; a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y; a.num_ptiles = p.num_ptiles; a.splits = p.splits;
  const int smem = kCtrl + 1024 + p.stages * (d->n_srcs + d->n_dz) * kTileB;
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(p.splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    if (tunables().pdl != 0) {
      cfg.attrs = attr;
      cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_shared_kernel, a);
    if (e != cudaSuccess) {
      set_error("wgrad_shared: launch failed: %s", cudaGetErrorString(e));
      return VSR_ERR_CUDA;
    }
  }
  VSR_CHECK_LAUNCH("wgrad_shared");
  r.ws = a.ws; r.bias_ws = a.bias_ws; r.n_dz = d->n_dz; r.total_taps = p.total_taps; r.splits = p.splits; r.accumulate = accumulate;
  const long n4 = (long)p.total_taps * 1024;
  wgrad_shared_reduce_kernel<<<(int)((n4 + 255) / 256) + 1, 256, 0, s>>>(r);
  VSR_CHECK_LAUNCH("wgrad_shared_reduce");
  return VSR_OK;
}
