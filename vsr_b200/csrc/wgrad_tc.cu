// wgrad_tc.cu — bf16 weight-gradient of the tap-GEMM on tcgen05 tensor cores (sm_100a).
//
//   dw[t][j][k] = sum_pix dz[pix][o0+j] * src_t[pix + off_t][c0_t + k]
//
// The reduction runs over pixels, so both operands are "MN-major" for the tensor core: a TMA box
// [128 pixels x 64 channels] (128-byte rows, 128B swizzle) is a K=128 x MN=64 operand tile.
// A work item is (group, N-slice of <=128 output channels, chunk of <=8 taps, pixel split):
// per 128-pixel tile the producer loads the dz tile once (B operand, N = 64 or 128) and streams
// tap *pairs* (A operand, M = 128 = 2 taps x 64 channels); the MMA thread accumulates every pair
// into its own TMEM accumulator [128 x N] over all pixel tiles of the split; the epilogue writes
// fp32 partials to the workspace, and a fixed-order kernel reduces the splits (deterministic).
//
// Replaces the weight half of aten.convolution_backward under drf_net.py:55-106,141-147.
#include <cuda.h>

#include <vector>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vsr {

// from tma_host.cu
int get_src_map_pub(const VsrTensor4& t, int bw, int bh, CUtensorMap* out);
void pick_box_pub(int h, int w, int* bw, int* bh);

namespace {

#ifdef VSR_ATTRIB        // timing attribution (tools/wg_attrib.py) exists only in builds with -DVSR_ATTRIB
constexpr bool kAttrib = true;
#else
constexpr bool kAttrib = false;
#endif

constexpr int kTile = 128;
constexpr int kTileBytes = kTile * 128;   // one [128 px x 64 ch] bf16 box
constexpr int kMaxAStages = 12;           // ring of tap pairs (2 boxes each); as many as shared memory holds
constexpr int kMaxChunk = 8;              // taps per work item
constexpr int kCtrl = 1024;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;

constexpr int kMaxPerm = 256;             // taps whose pairing order can travel in the kernel arguments

struct WgArgs {
  CUtensorMap src_maps[VSR_MAX_SRCS];
  CUtensorMap tall_maps[VSR_MAX_SRCS];   // shared-pair mode: box of bh + 1 pixel rows
  CUtensorMap dz_map;
  uint16_t perm[kMaxPerm];               // position in the pairing order -> tap index (identity if unused)
  int use_perm;                          // perm[] is valid
  int tall;                              // every pair (2p, 2p+1) is (tap, same tap one pixel row lower): one A load
  int row_bytes;                         // bw * 128
  int stages;                            // tap-pair ring depth
  int slot_bytes;                        // bytes of one ring slot
  int debug;                             // VSR_WG_DEBUG: 32 = per-role cycle counts of block 0
  const int4* tap_tab;
  const int4* group_tab;
  float* ws;              // [splits][n_taps_total][nt][64]
  float* bias_ws;         // [splits][cout] column sums of dz (bias gradient), or NULL
  int cout;

  int n_items, splits, chunks;   // item = (group, N-slice, chunk of <= kMaxChunk taps)
  int nt, ncta;           // group width, per-item N (64 or 128)
  int n_taps_total;
  int N, H, W;
  int bw, bh, tiles_x, tiles_y, num_ptiles;
};

__device__ __forceinline__ void tile_coord(const WgArgs& a, int pt, int* n, int* y0, int* x0) {
  const int tx = pt % a.tiles_x;
  pt /= a.tiles_x;
  const int ty = pt % a.tiles_y;
  *n = pt / a.tiles_y;
  *x0 = tx * a.bw;
  *y0 = ty * a.bh;
}

// MN-major 128B-swizzled operand: 64-element MN atoms are `lbo` bytes apart, 8-row K groups 1024 B.
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo) {
  return ptx::make_sw128_desc(addr, lbo, 1024);
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t a_full = base, a_empty = base + 128, b_full = base + 256, b_empty = base + 272;   // 16 slots each for a_*
  const uint32_t done_bar = base + 288, tmem_slot = base + 296;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 296);
  const uint32_t slot_bytes = static_cast<uint32_t>(a.slot_bytes);   // one tap pair: two boxes, or one box of bh + 1 rows
  const int nb = a.ncta / 64;                                // dz boxes per tile
  const uint32_t b_bytes = nb * kTileBytes;
  const uint32_t b_base = base + kCtrl;                      // 2 buffers
  const uint32_t a_base = b_base + 2 * b_bytes;              // a.stages x 2 boxes

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item_i = blockIdx.x % a.n_items, split = blockIdx.x / a.n_items;
  const int halves = a.nt / a.ncta;
  const int chunk = item_i % a.chunks;
  const int4 grp = __ldg(a.group_tab + (item_i / a.chunks) / halves);
  int4 item;                                   // {-, n0, first tap (global), taps in this chunk}
  item.x = 0;
  item.y = ((item_i / a.chunks) % halves) * a.ncta;
  item.z = grp.y + chunk * kMaxChunk;
  item.w = max(0, min(kMaxChunk, grp.z - chunk * kMaxChunk));
  const int n_pairs = (item.w + 1) >> 1;
  const int per = (a.num_ptiles + a.splits - 1) / a.splits;
  const int pt0 = split * per;
  const int pt1 = min(pt0 + per, a.num_ptiles);
  // the first chunk of every (group, N-slice) also sums the columns of its dz tiles (bias gradient)
  const bool do_cs = a.bias_ws != nullptr && chunk == 0;

  float cs0_out = 0.f, cs1_out = 0.f;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(a_full + 8 * s, 1);
      ptx::mbar_init(a_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(b_full + 8 * b, 1);
      ptx::mbar_init(b_empty + 8 * b, do_cs ? 5 : 1);   // MMA commit (+ the 4 column-sum warps)
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
    // The whole warp walks the (uniform) loops and one elected lane issues: the operands stay on the
    // uniform datapath (see tapgemm_tc2.cu).
    {
      const bool leader = ptx::elect_one();
      const bool prof = kAttrib && (a.debug & 32) != 0;
      long long p_wb = 0, p_wa = 0, p_t0 = kAttrib ? clock64() : 0, c0 = 0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int4 taps[kMaxChunk];                     // the chunk's taps, fetched once (independent loads)
#pragma unroll
      for (int i = 0; i < kMaxChunk; ++i) {
        const int pos = item.z + (i < item.w ? i : 0);
        int4 t = __ldg(a.tap_tab + (a.use_perm ? (int)a.perm[pos] : pos));
        t.x = __shfl_sync(0xffffffffu, t.x, 0);
        t.y = __shfl_sync(0xffffffffu, t.y, 0);
        t.z = __shfl_sync(0xffffffffu, t.z, 0);
        t.w = __shfl_sync(0xffffffffu, t.w, 0);
        taps[i] = t;
      }
      for (int pt = pt0; pt < pt1; ++pt, ++it) {
        int n, y0, x0;
        tile_coord(a, pt, &n, &y0, &x0);
        const int bb = it & 1;
        if (prof) c0 = clock64();
        ptx::mbar_wait(b_empty + 8 * bb, ((it >> 1) & 1) ^ 1u);
        if (prof) p_wb += clock64() - c0;
        if (leader) {
          ptx::mbar_arrive_expect_tx(b_full + 8 * bb, b_bytes);
          for (int i = 0; i < nb; ++i)
            ptx::tma_load_4d(b_base + bb * b_bytes + i * kTileBytes, &a.dz_map, b_full + 8 * bb,
                             grp.x + item.y + i * 64, x0, y0, n);
        }
        for (int p = 0; p < n_pairs; ++p) {
          if (prof) c0 = clock64();
          ptx::mbar_wait(a_empty + 8 * stage, phase ^ 1u);
          if (prof) p_wa += clock64() - c0;
          const uint32_t sa = a_base + stage * slot_bytes;
          if (a.tall) {
            // the pair is (tap, tap shifted one pixel row down): one box of bh + 1 rows serves both
            int4 tap = taps[0];
#pragma unroll
            for (int i = 1; i < kMaxChunk; ++i) tap = (i == 2 * p) ? taps[i] : tap;
            if (leader) {
              ptx::mbar_arrive_expect_tx(a_full + 8 * stage, kTileBytes + a.row_bytes);
              ptx::tma_load_4d(sa, &a.tall_maps[tap.x], a_full + 8 * stage, tap.w, x0 + tap.z, y0 + tap.y, n);
            }
          } else {
            if (leader) ptx::mbar_arrive_expect_tx(a_full + 8 * stage, 2 * kTileBytes);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              int tl = 2 * p + h;
              if (tl >= item.w) tl = 2 * p;                      // odd count: duplicate (rows ignored)
              int4 tap = taps[0];
#pragma unroll
              for (int i = 1; i < kMaxChunk; ++i) tap = (i == tl) ? taps[i] : tap;
              if (leader)
                ptx::tma_load_4d(sa + h * kTileBytes, &a.src_maps[tap.x], a_full + 8 * stage, tap.w, x0 + tap.z,
                                 y0 + tap.y, n);
            }
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof && blockIdx.x == 0 && leader)
        printf("wg-prof producer: total %lld cyc, %d tiles x %d pairs, wait(b_empty) %lld, wait(a_empty) %lld, tall %d\n",
               clock64() - p_t0, it, n_pairs, p_wb, p_wa, a.tall);
    }
  } else if (warp == 1) {
    {
      const bool leader = ptx::elect_one();
      const uint32_t idesc = ptx::make_idesc_bf16(128, a.ncta, 1, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t lbo = a.tall ? a.row_bytes : kTileBytes;
      const bool prof = kAttrib && (a.debug & 32) != 0;
      long long m_wb = 0, m_wa = 0, m_is = 0, m_t0 = kAttrib ? clock64() : 0, c0 = 0, c1 = 0;
      const uint64_t m_g0 = kAttrib ? ptx::globaltimer_ns() : 0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pt = pt0; pt < pt1; ++pt, ++it) {
        const int bb = it & 1;
        if (prof) c0 = clock64();
        ptx::mbar_wait(b_full + 8 * bb, (it >> 1) & 1);
        ptx::tc_fence_after();
        if (prof) m_wb += clock64() - c0;
        const uint32_t sb = b_base + bb * b_bytes;
        for (int p = 0; p < n_pairs; ++p) {
          if (prof) c0 = clock64();
          ptx::mbar_wait(a_full + 8 * stage, phase);
          ptx::tc_fence_after();
          if (prof) { c1 = clock64(); m_wa += c1 - c0; }
          const uint32_t sa = a_base + stage * slot_bytes;
          if (leader) {
#pragma unroll
            for (int k = 0; k < kTile / 16; ++k) {
              // 16 pixels = two 8-row K groups = 2048 bytes
              const uint64_t ad = mn_desc(sa + k * 2048, lbo);
              const uint64_t bd = mn_desc(sb + k * 2048, kTileBytes);
              ptx::mma_bf16_ss(tmem_u + p * a.ncta, ad, bd, idesc, (it | k) != 0);
            }
            ptx::mma_commit(a_empty + 8 * stage);
          }
          if (prof) m_is += clock64() - c1;
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        if (leader) ptx::mma_commit(b_empty + 8 * bb);
      }
      if (leader) ptx::mma_commit(done_bar);
      if (prof && blockIdx.x == 0 && leader)
        printf("wg-prof mma: total %lld cyc in %llu ns, wait(b_full) %lld, wait(a_full) %lld, issue %lld\n", clock64() - m_t0,
               (unsigned long long)(ptx::globaltimer_ns() - m_g0), m_wb, m_wa, m_is);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;       // (tap half, k)
    const int half = row >> 6, k = row & 63;
    float cs0 = 0.f, cs1 = 0.f;
    const int cpairs = a.ncta >> 1;            // column pairs of the dz tile (32 or 64)
    const int cp = row % cpairs, part = row / cpairs;
    if (do_cs) {
      const int rows_per = kTile / (kTile / cpairs);
      const int c = 2 * cp, cc = c & 63;
      int it = 0;
      for (int pt = pt0; pt < pt1; ++pt, ++it) {
        const int bb = it & 1;
        ptx::mbar_wait(b_full + 8 * bb, (it >> 1) & 1);
        const uint32_t tb = b_base + bb * b_bytes + (c >> 6) * kTileBytes;
#pragma unroll 8
        for (int r = part * rows_per; r < (part + 1) * rows_per; ++r) {
          uint32_t u;
          const uint32_t addr = tb + r * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(addr) : "memory");
          cs0 += bf16_lo(u);
          cs1 += bf16_hi(u);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(b_empty + 8 * bb);
      }
    }
    const long long e_t0 = kAttrib ? clock64() : 0;
    if (pt1 > pt0) {
      ptx::mbar_wait(done_bar, 0);
      ptx::tc_fence_after();
    }
    const long long e_t1 = kAttrib ? clock64() : 0;
    float* wsp = a.ws + (size_t)split * a.n_taps_total * a.nt * 64;
    for (int p = 0; p < n_pairs; ++p) {
      const int tl = 2 * p + half;
      const bool live = tl < item.w;
      const int ti = a.use_perm ? (int)a.perm[item.z + (live ? tl : 0)] : item.z + tl;
      const uint32_t taddr = tmem_base + p * a.ncta + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int c = 0; c < a.ncta; c += 16) {
        uint32_t r[16];
        if (pt1 > pt0) {
          ptx::tmem_ld16(taddr + c, r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = 0u;      // empty split: contributes zeros
        }
        if (live) {
          float* o = wsp + ((size_t)ti * a.nt + item.y + c) * 64 + k;
#pragma unroll
          for (int i = 0; i < 16; ++i) o[(size_t)i * 64] = __uint_as_float(r[i]);
        }
      }
    }
    if (kAttrib && (a.debug & 32) && blockIdx.x == 0 && warp == 2 && lane == 0)
      printf("wg-prof epilogue: wait(done) %lld cyc, drain %lld cyc\n", e_t1 - e_t0, clock64() - e_t1);
    cs0_out = cs0;
    cs1_out = cs1;
  }
  if (do_cs && warp >= 2) {
    // fold the row parts in fixed order through smem (all MMAs / TMA loads of this CTA are done)
    const int quarter = warp & 3, row = quarter * 32 + lane;
    float2* scr = reinterpret_cast<float2*>(gen + kCtrl);
    const int cpairs = a.ncta >> 1, cp = row % cpairs, part = row / cpairs, parts = kTile / cpairs;
    scr[row] = make_float2(cs0_out, cs1_out);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (part == 0) {
      float s0 = 0.f, s1 = 0.f;
      for (int q = 0; q < parts; ++q) { s0 += scr[q * cpairs + cp].x; s1 += scr[q * cpairs + cp].y; }
      float* o = a.bias_ws + (size_t)split * a.cout + grp.x + item.y + 2 * cp;
      o[0] = s0;
      o[1] = s1;
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

// dw[i] (+)= sum_s ws[s][i]: block = 64 float4 columns x 4 split lanes; every lane sums the splits
// s = lane (mod 4) in increasing order, then a fixed-order fold -> deterministic.
// The last block (if bias_ws) reduces the bias partials: db[q] (+)= sum_s sum_{c = q mod period} bias_ws[s][c].
__global__ void __launch_bounds__(256) wg_reduce_kernel(const float4* __restrict__ ws, float4* __restrict__ dw, long n4,
                                                       int splits, int accumulate, const float* __restrict__ bias_ws,
                                                       int cout, int period, float* __restrict__ db) {
  __shared__ float4 sm[4][64];
  if (bias_ws != nullptr && blockIdx.x == gridDim.x - 1) {
    float* colsum = reinterpret_cast<float*>(sm);          // [cout <= 1024]
    for (int c = threadIdx.x; c < cout; c += blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < splits; ++k) s += bias_ws[(size_t)k * cout + c];
      colsum[c] = s;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < period; q += blockDim.x) {
      float s = accumulate ? db[q] : 0.f;
      for (int c = q; c < cout; c += period) s += colsum[c];
      db[q] = s;
    }
    return;
  }
  const int col = threadIdx.x & 63, part = threadIdx.x >> 6;
  const long i = (long)blockIdx.x * 64 + col;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    // eight independent 16-byte loads in flight per thread; the additions keep the order k = part, part+4, ...
    int k = part;
    for (; k + 28 < splits; k += 32) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(ws + (size_t)(k + 4 * u) * n4 + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; k < splits; k += 4) {
      const float4 v = __ldg(ws + (size_t)k * n4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  sm[part][col] = s;
  __syncthreads();
  if (part == 0 && i < n4) {
    float4 t = accumulate ? dw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) { t.x += sm[q][col].x; t.y += sm[q][col].y; t.z += sm[q][col].z; t.w += sm[q][col].w; }
    dw[i] = t;
  }
}

// Host: order the taps as pairs (t, t') with equal (src, c0, dx) and dy' = dy + 1.  False if some tap has no partner.
bool pair_rows(const int32_t* taps, int n, uint16_t* perm) {
  std::vector<char> used(n, 0);
  int out = 0;
  for (int i = 0; i < n; ++i) {
    if (used[i]) continue;
    // i must be the upper tap of its pair: find the partner one row lower; if i is itself a lower row of an
    // unused tap, that tap comes first
    int up = i;
    for (int j = 0; j < n; ++j)
      if (!used[j] && j != i && taps[4 * j] == taps[4 * i] && taps[4 * j + 3] == taps[4 * i + 3] &&
          taps[4 * j + 2] == taps[4 * i + 2] && taps[4 * j + 1] == taps[4 * i + 1] - 1) {
        // prefer pairing (j, i) only if i has no lower partner of its own
        bool has_lower = false;
        for (int k = 0; k < n; ++k)
          if (!used[k] && taps[4 * k] == taps[4 * i] && taps[4 * k + 3] == taps[4 * i + 3] &&
              taps[4 * k + 2] == taps[4 * i + 2] && taps[4 * k + 1] == taps[4 * i + 1] + 1)
            has_lower = true;
        if (!has_lower) up = j;
        break;
      }
    int low = -1;
    for (int k = 0; k < n; ++k)
      if (!used[k] && k != up && taps[4 * k] == taps[4 * up] && taps[4 * k + 3] == taps[4 * up + 3] &&
          taps[4 * k + 2] == taps[4 * up + 2] && taps[4 * k + 1] == taps[4 * up + 1] + 1) {
        low = k;
        break;
      }
    if (low < 0) return false;
    used[up] = used[low] = 1;
    perm[out++] = (uint16_t)up;
    perm[out++] = (uint16_t)low;
  }
  return out == n;
}

struct WgPlan {
  int ncta, n_items, splits, chunks, bw, bh, num_ptiles;
};

WgPlan make_plan(const VsrTapGemmDesc* d) {
  WgPlan p;
  p.ncta = d->nt % 128 == 0 ? 128 : 64;
  pick_box_pub(d->out.h, d->out.w, &p.bw, &p.bh);
  const int tx = (d->out.w + p.bw - 1) / p.bw, ty = (d->out.h + p.bh - 1) / p.bh;
  p.num_ptiles = d->out.n * tx * ty;
  const int mg = d->max_group_taps > 0 ? d->max_group_taps : (d->n_taps_total + d->n_groups - 1) / d->n_groups;
  p.chunks = (mg + kMaxChunk - 1) / kMaxChunk;
  p.n_items = d->n_groups * (d->nt / p.ncta) * p.chunks;
  int s = num_sms() / p.n_items;   // ONE wave of equal-work CTAs: items * splits must not exceed the SM count
                                   // (rounding up cost a second, nearly empty wave = 2x the kernel time)
  if (s > p.num_ptiles) s = p.num_ptiles;
  if (s < 1) s = 1;
  p.splits = s;
  return p;
}

}  // namespace

bool wgrad_tc_supported(const VsrTapGemmDesc* d) {
  return d->dtype == VSR_BF16 && d->kc == 64 && d->nt % 64 == 0 && d->nt <= 256 && d->out.c % 8 == 0 &&
         (d->max_group_taps > 0 || d->n_taps_total % d->n_groups == 0);
}

size_t wgrad_tc_workspace(const VsrTapGemmDesc* d) {
  const WgPlan p = make_plan(d);
  return (size_t)p.splits * ((size_t)d->n_taps_total * d->nt * 64 + d->out.c) * sizeof(float);
}

// partial pass: leaves per-split partial sums (weights, and bias columns if want_bias) in workspace
int wgrad_tc_partial(const VsrTapGemmDesc* d, int want_bias, int slice, int n_slices, void* workspace,
                     cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(wgrad smem) failed: %s", cudaGetErrorString(e));
      return VSR_ERR_CUDA;
    }
    attr_set = true;
  }
  const WgPlan p = make_plan(d);
  WgArgs a;
  memset(&a, 0, sizeof(a));
  for (int s = 0; s < d->n_srcs; ++s) {
    int rc = get_src_map_pub(d->srcs[s], p.bw, p.bh, &a.src_maps[s]);
    if (rc != VSR_OK) return rc;
  }
  int rc = get_src_map_pub(d->out, p.bw, p.bh, &a.dz_map);
  if (rc != VSR_OK) return rc;
  a.tap_tab = reinterpret_cast<const int4*>(d->tap_tab);
  a.group_tab = reinterpret_cast<const int4*>(d->group_tab);
  a.row_bytes = p.bw * 128;
  {
    a.debug = tunables().wg_debug > 0 ? tunables().wg_debug : 0;
  }
  {
    // shared-pair mode: reorder the taps of every group so that positions (2p, 2p+1) hold a tap and the tap
    // with the same source, channel slice and dx one pixel row lower; all pairs of the launch must be such
    if (tunables().wg_tall != 0 && d->tap_tab_host && d->n_taps_total <= kMaxPerm && d->n_groups == 1 &&
        p.bw >= 8 && p.bw * p.bh == kTile && d->n_taps_total % 2 == 0 && pair_rows(d->tap_tab_host, d->n_taps_total, a.perm)) {
      a.use_perm = 1;
      a.tall = 1;
      for (int s = 0; s < d->n_srcs; ++s) {
        int rc = get_src_map_pub(d->srcs[s], p.bw, p.bh + 1, &a.tall_maps[s]);
        if (rc != VSR_OK) return rc;
      }
    }
  }
  // workspace: [n_slices * splits][n] weight partials, then [n_slices * splits][cout] bias partials
  const size_t n_w = (size_t)d->n_taps_total * d->nt * 64;
  float* ws0 = static_cast<float*>(workspace);
  a.ws = ws0 + (size_t)slice * p.splits * n_w;
  a.bias_ws = want_bias ? ws0 + (size_t)n_slices * p.splits * n_w + (size_t)slice * p.splits * d->out.c : nullptr;
  a.cout = d->out.c;
  a.n_items = p.n_items;
  a.splits = p.splits;
  a.chunks = p.chunks;
  a.nt = d->nt;
  a.ncta = p.ncta;
  a.n_taps_total = d->n_taps_total;
  a.N = d->out.n; a.H = d->out.h; a.W = d->out.w;
  a.bw = p.bw; a.bh = p.bh;
  a.tiles_x = (a.W + p.bw - 1) / p.bw;
  a.tiles_y = (a.H + p.bh - 1) / p.bh;
  a.num_ptiles = p.num_ptiles;
  const int nb = p.ncta / 64;
  // the kernel is bound by the bytes it keeps in flight (frame-batched maps no longer fit the L2): a
  // shared-pair slot is 20 KB instead of 32 KB, so the ring is 1.5x deeper in tap pairs
  a.slot_bytes = a.tall ? kTileBytes + a.row_bytes : 2 * kTileBytes;
  a.stages = (227 * 1024 - kCtrl - 1024 - 2 * nb * kTileBytes) / a.slot_bytes;
  if (a.stages > kMaxAStages) a.stages = kMaxAStages;
  const int smem = kCtrl + 1024 + 2 * nb * kTileBytes + a.stages * a.slot_bytes;
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(p.n_items * p.splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    if (tunables().pdl != 0) {
      cfg.attrs = attr;
      cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_tc_kernel, a);
    if (e != cudaSuccess) {
      set_error("wgrad_tc: launch failed: %s", cudaGetErrorString(e));
      return VSR_ERR_CUDA;
    }
  }
  VSR_CHECK_LAUNCH("wgrad_tc");
  return VSR_OK;
}

// final pass: dw (+)= sum over splits, db[q] (+)= sum over splits and columns c = q (mod period)
int wgrad_tc_finish(const VsrTapGemmDesc* d, float* dw, float* db, int db_period, int accumulate, int used_slices,
                    int n_slices, void* workspace, cudaStream_t stream) {
  WgPlan p = make_plan(d);
  const float* ws = static_cast<const float*>(workspace);
  const bool with_bias = db != nullptr;
  const float* bias_ws = with_bias ? ws + (size_t)n_slices * p.splits * d->n_taps_total * d->nt * 64 : nullptr;
  p.splits *= used_slices;                       // slices 0..used-1 are contiguous rows
  const long n = (long)d->n_taps_total * d->nt * 64;
  const long n4 = n / 4;     // nt * 64 is a multiple of 4
  wg_reduce_kernel<<<(int)((n4 + 63) / 64) + (with_bias ? 1 : 0), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(ws), reinterpret_cast<float4*>(dw), n4, p.splits, accumulate, bias_ws,
      d->out.c, db_period, db);
  VSR_CHECK_LAUNCH("wgrad_tc_reduce");
  return VSR_OK;
}

bool wgrad_tc_bias_ok(const VsrTapGemmDesc* d, int db_period) {
  return d->out.c <= 1024 && db_period > 0 && d->out.c % db_period == 0;
}

int wgrad_tc_launch(const VsrTapGemmDesc* d, float* dw, float* db, int db_period, int accumulate, void* workspace,
                    cudaStream_t stream) {
  const bool with_bias = db != nullptr && wgrad_tc_bias_ok(d, db_period);
  int rc = wgrad_tc_partial(d, with_bias, 0, 1, workspace, stream);
  if (rc != VSR_OK) return rc;
  return wgrad_tc_finish(d, dw, with_bias ? db : nullptr, db_period, accumulate, 1, 1, workspace, stream);
}

}  // namespace vsr
