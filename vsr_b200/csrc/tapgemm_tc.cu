// tapgemm_tc.cu — bf16 tap-GEMM on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM.  A work tile is 128 output pixels (a bw x bh box of one image)
// times the `nt` output channels of one group.  Per tap the producer thread issues one 4-D TMA
// box load of the (shifted) source pixels [128 x 64 ch] (out-of-image pixels are zero-filled by
// the TMA unit, which is the convolution padding) and one bulk copy of the tap's pre-swizzled
// weight slab [nt x 64]; a single thread issues 4 tcgen05.mma (128 x nt x 16) per tap into a
// TMEM accumulator; four epilogue warps drain the other TMEM accumulator of the previous tile
// (bias / residual / PReLU / PReLU-backward / second output) while the next tile's MMAs run.
//
// Replaces aten.convolution / convolution_backward(data) under drf_net.py:55-106,141-147.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vsr {
namespace {

constexpr int kBlockM = 128;
constexpr int kKc = 64;                       // bf16 channels per tap = one 128-byte row
constexpr int kATileBytes = kBlockM * 128;    // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kStageTileBytes = 4096;         // per epilogue warp: 32 rows x 128 B staging tile
constexpr int kEpiWarps = 8;                  // two epilogue groups, one per TMEM accumulator buffer
constexpr int kBiasBytes = 1024;              // bias of outputs with <= 256 channels is staged in smem
constexpr int kMaxSmemGroups = 48;            // group table rows cached in smem (16 B each, ctrl[256..1024))
constexpr int kMaxSmemTaps = 256;             // packed tap entries cached in smem (4 B each)
constexpr int kTapBytes = kMaxSmemTaps * 4;
constexpr int kCtrlBytes = 1024 + kTapBytes + kBiasBytes + kEpiWarps * kStageTileBytes;
constexpr int kTmemCols = 512;
constexpr int kThreads = 64 + 32 * 8;          // warp0 TMA, warp1 MMA, warps 2-5 / 6-9 epilogue groups

struct TcArgs {
  CUtensorMap maps[VSR_MAX_SRCS];
  const int4* tap_tab;
  const int4* group_tab;
  const uint8_t* w;
  const float* bias;
  const float* slope;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* aux_y;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  const __nv_bfloat16* res2;
  float* slope_partials;
  float out_scale;
  int epi;
  int nt, n_groups;
  int N, H, W, Cout;
  int bw, bh, tiles_x, tiles_y;
  int bw_shift;            // bw is a power of two
  int num_tiles;
  int stages;
  int n_taps_total;
  int resident;            // 1: the current group's weight slabs stay in smem across tiles
  int res_bytes;           // size of the resident weight region
  int m_tiles;             // pixel tiles per group
  int debug;               // timing-attribution switches (VSR_TC_DEBUG), wrong results when non-zero
};

// tap entry packed into 32 bits: src[0:4) | dy+8 [4:8) | dx+8 [8:12) | c0/8 [12:32)
__device__ __forceinline__ uint32_t pack_tap(const int4& t) {
  return (uint32_t)t.x | ((uint32_t)(t.y + 8) << 4) | ((uint32_t)(t.z + 8) << 8) | ((uint32_t)(t.w >> 3) << 12);
}
__device__ __forceinline__ int4 unpack_tap(uint32_t p) {
  return make_int4((int)(p & 15u), (int)((p >> 4) & 15u) - 8, (int)((p >> 8) & 15u) - 8, (int)(p >> 12) << 3);
}

struct TileCoord {
  int g, n, y0, x0;
};

// non-resident: groups vary fastest (concurrent CTAs share A tiles in L2);
// resident: group-major, every CTA walks a contiguous tile range (the group rarely changes).
__device__ __forceinline__ TileCoord decode_tile(const TcArgs& a, int tile) {
  TileCoord t;
  int mt;
  if (a.resident) {
    t.g = tile / a.m_tiles;
    mt = tile - t.g * a.m_tiles;
  } else {
    t.g = tile % a.n_groups;
    mt = tile / a.n_groups;
  }
  const int tx = mt % a.tiles_x;
  mt /= a.tiles_x;
  const int ty = mt % a.tiles_y;
  t.n = mt / a.tiles_y;
  t.x0 = tx * a.bw;
  t.y0 = ty * a.bh;
  return t;
}

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x);
  f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z);
  f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]);
  q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]);
  q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// staging tile: row r (0..31), 16-byte chunk q (0..7) lives at r*128 + ((q ^ (r & 7)) << 4)
__device__ __forceinline__ uint32_t stg_addr(uint32_t base, int r, int q) {
  return base + r * 128 + ((q ^ (r & 7)) << 4);
}
// coalesced read of this warp's [32 rows x 64 ch] bf16 block of a pixel-major map into f[64] of
// the row-owning thread (lane = row): 4 rows x 128 B per load instruction, via the staging tile.
__device__ __forceinline__ void load_rows64(const __nv_bfloat16* __restrict__ src, const size_t (&rowoff)[8],
                                            uint32_t valid_mask, int col0, uint32_t stg, int lane, float* f) {
  const int q = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((valid_mask >> it) & 1u) v = __ldg(reinterpret_cast<const uint4*>(src + rowoff[it] + col0) + q);
    st_shared_v4(stg_addr(stg, it * 4 + (lane >> 3), q), v);
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 8; ++c) unpack8(ld_shared_v4(stg_addr(stg, lane, c)), f + 8 * c);
  __syncwarp();
}
// the reverse: v[64] of the row-owning thread -> bf16 -> coalesced store
__device__ __forceinline__ void store_rows64(__nv_bfloat16* __restrict__ dst, const size_t (&rowoff)[8],
                                             uint32_t valid_mask, int col0, uint32_t stg, int lane, const float* v) {
  const int q = lane & 7;
#pragma unroll
  for (int c = 0; c < 8; ++c) st_shared_v4(stg_addr(stg, lane, c), pack8(v + 8 * c));
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const uint4 o = ld_shared_v4(stg_addr(stg, it * 4 + (lane >> 3), q));
    if ((valid_mask >> it) & 1u) *(reinterpret_cast<uint4*>(dst + rowoff[it] + col0) + q) = o;
  }
  __syncwarp();
}

// FIXED_EPI >= 0: the epilogue flag set is a compile-time constant (dead variants are not emitted, the
// hot loop stays small); FIXED_EPI < 0: flags are read from the arguments (any combination).
template <int FIXED_EPI>
__global__ void __launch_bounds__(kThreads, 1) tapgemm_tc_kernel(const __grid_constant__ TcArgs a) {
  const int epi = FIXED_EPI >= 0 ? FIXED_EPI : a.epi;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  uint8_t* smem_gen = smem_raw;
  if (smem_base & 1023u) __trap();             // the swizzled tiles need a 1024-byte aligned window

  // control block
  const uint32_t full_bar = smem_base;                     // kMaxStages x 8 B
  const uint32_t empty_bar = smem_base + 64;               // kMaxStages x 8 B
  const uint32_t tfull_bar = smem_base + 128;              // 2 x 8 B
  const uint32_t tempty_bar = smem_base + 144;             // 2 x 8 B
  const uint32_t tmem_slot = smem_base + 160;              // u32
  const uint32_t bres_full = smem_base + 168, bres_empty = smem_base + 176;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + 160);
  float* red = reinterpret_cast<float*>(smem_gen + 192);   // 8 floats
  int4* grp_s = reinterpret_cast<int4*>(smem_gen + 256);        // group table (<= 48 rows)
  uint32_t* tap_s = reinterpret_cast<uint32_t*>(smem_gen + 1024);   // packed tap table (<= 256 entries)
  float* bias_s = reinterpret_cast<float*>(smem_gen + 1024 + kTapBytes);   // bias when Cout <= 256
  const uint32_t stg_base = smem_base + 1024 + kTapBytes + kBiasBytes;     // 8 x 4 KiB staging tiles
  const bool grp_in_smem = a.n_groups <= kMaxSmemGroups;
  const bool taps_in_smem = a.n_taps_total <= kMaxSmemTaps;
  if (grp_in_smem)
    for (int i = threadIdx.x; i < a.n_groups; i += blockDim.x) grp_s[i] = __ldg(a.group_tab + i);
  if (taps_in_smem)
    for (int i = threadIdx.x; i < a.n_taps_total; i += blockDim.x) tap_s[i] = pack_tap(__ldg(a.tap_tab + i));
  const bool bias_in_smem = (epi & VSR_EPI_BIAS) && a.Cout <= kBiasBytes / 4;
  if (bias_in_smem)
    for (int i = threadIdx.x; i < a.Cout; i += blockDim.x) bias_s[i] = a.bias[i];
  const uint32_t res_base = smem_base + kCtrlBytes;        // resident weight slabs (resident mode)
  const uint32_t stage_base = res_base + a.res_bytes;
  const uint32_t b_bytes = static_cast<uint32_t>(a.nt) * 128u;
  const uint32_t stage_bytes = a.resident ? kATileBytes : kATileBytes + b_bytes;
  // tile walk of this CTA
  const int tile_begin = a.resident ? (int)((long)a.num_tiles * blockIdx.x / gridDim.x) : (int)blockIdx.x;
  const int tile_end = a.resident ? (int)((long)a.num_tiles * (blockIdx.x + 1) / gridDim.x) : a.num_tiles;
  const int tile_step = a.resident ? 1 : (int)gridDim.x;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar + 8 * b, 1);
      ptx::mbar_init(tempty_bar + 8 * b, 128);
    }
    ptx::mbar_init(bres_full, 1);
    ptx::mbar_init(bres_empty, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int cur_g = -1;
      uint32_t gcount = 0;
      const bool prof = (a.debug & 32) != 0;
      long long p_wait = 0, p_issue = 0, p_n = 0, p_t0 = clock64();
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const TileCoord tc = decode_tile(a, tile);
        const int4 grp = grp_in_smem ? grp_s[tc.g] : __ldg(a.group_tab + tc.g);
        if (a.resident && tc.g != cur_g) {
          // (re)load this group's weight slabs once all MMAs of the previous group are done
          ptx::mbar_wait(bres_empty, (gcount & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(bres_full, static_cast<uint32_t>(grp.z) * b_bytes);
          for (int t = 0; t < grp.z; ++t)
            ptx::bulk_load(res_base + t * b_bytes, a.w + static_cast<size_t>(grp.y + t) * b_bytes, b_bytes, bres_full);
          cur_g = tc.g;
          ++gcount;
        }
        for (int t = 0; t < grp.z; ++t) {
          const int4 tap = taps_in_smem ? unpack_tap(tap_s[grp.y + t]) : __ldg(a.tap_tab + grp.y + t);
          long long c0 = 0;
          if (prof) c0 = clock64();
          ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1u);
          if (prof) { const long long c1 = clock64(); p_wait += c1 - c0; ++p_n; }
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t sa = stage_base + stage * stage_bytes;
          if (a.debug & 6) {
            // attribution runs: skip the A (2) and/or B (4) transfer, keep the barrier protocol
            uint32_t tx = 0;
            if (!(a.debug & 2)) tx += kATileBytes;
            if (!(a.debug & 4) && !a.resident) tx += b_bytes;
            if (tx == 0) { ptx::mbar_arrive(fb); } else { ptx::mbar_arrive_expect_tx(fb, tx); }
            if (!(a.debug & 2)) ptx::tma_load_4d(sa, &a.maps[tap.x], fb, tap.w, tc.x0 + tap.z, tc.y0 + tap.y, tc.n);
            if (!(a.debug & 4) && !a.resident)
              ptx::bulk_load(sa + kATileBytes, a.w + static_cast<size_t>(grp.y + t) * b_bytes, b_bytes, fb);
            if (++stage == a.stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          ptx::mbar_arrive_expect_tx(fb, stage_bytes);
          ptx::tma_load_4d(sa, &a.maps[tap.x], fb, tap.w, tc.x0 + tap.z, tc.y0 + tap.y, tc.n);
          if (!a.resident)
            ptx::bulk_load(sa + kATileBytes, a.w + static_cast<size_t>(grp.y + t) * b_bytes, b_bytes, fb);
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof && blockIdx.x == 0)
        printf("tc-prof producer: total %lld cyc, %lld taps, wait(empty) %lld\n", clock64() - p_t0, p_n, p_wait);
      (void)p_issue;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kBlockM, a.nt, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int cur_g = -1;
      uint32_t gcount = 0;
      const bool prof = (a.debug & 32) != 0;
      long long m_wfull = 0, m_wtmem = 0, m_issue = 0, m_commit = 0, m_n = 0, m_t0 = clock64();
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
        const TileCoord tc = decode_tile(a, tile);
        const int4 grp = grp_in_smem ? grp_s[tc.g] : __ldg(a.group_tab + tc.g);
        if (a.resident && tc.g != cur_g) {
          ptx::mbar_wait(bres_full, gcount & 1u);
          cur_g = tc.g;
          ++gcount;
        }
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (prof) c0 = clock64();
        ptx::mbar_wait(tempty_bar + 8 * buf, bphase ^ 1u);
        ptx::tc_fence_after();
        if (prof) m_wtmem += clock64() - c0;
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * a.nt);
        for (int t = 0; t < grp.z; ++t) {
          if (prof) c0 = clock64();
          ptx::mbar_wait(full_bar + 8 * stage, phase);
          ptx::tc_fence_after();
          if (prof) c1 = clock64();
          const uint32_t sa = stage_base + stage * stage_bytes;
          const uint64_t adesc = ptx::make_sw128_desc(sa, 16, 1024);
          const uint64_t bdesc = ptx::make_sw128_desc(a.resident ? res_base + t * b_bytes : sa + kATileBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kKc / 16; ++k) {
            if (a.debug & 8) break;
            // advancing K by 16 bf16 = 32 bytes = 2 descriptor address units
            ptx::mma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (t | k) != 0);
          }
          if (prof) c2 = clock64();
          ptx::mma_commit(empty_bar + 8 * stage);
          if (prof) { c3 = clock64(); m_wfull += c1 - c0; m_issue += c2 - c1; m_commit += c3 - c2; ++m_n; }
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        ptx::mma_commit(tfull_bar + 8 * buf);
        if (a.resident) {
          const int next = tile + tile_step;
          if (next >= tile_end || decode_tile(a, next).g != cur_g) ptx::mma_commit(bres_empty);
        }
      }
      if (prof && blockIdx.x == 0)
        printf("tc-prof mma: total %lld cyc, %lld taps, %d tiles, wait(full) %lld, issue %lld, commit %lld, wait(tmem) %lld\n",
               clock64() - m_t0, m_n, it, m_wfull, m_issue, m_commit, m_wtmem);
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int quarter = warp & 3;
    const int egroup = (warp - 2) >> 2;          // drains TMEM buffer `egroup` (tiles with it&1 == egroup)
    const int row = quarter * 32 + lane;
    const int ry = row >> a.bw_shift, rx = row & (a.bw - 1);
    const float slope = (epi & (VSR_EPI_PRELU | VSR_EPI_PRELU_BWD)) ? __ldg(a.slope) : 0.f;
    const float inv_slope = slope != 0.f ? 1.f / slope : 0.f;
    float slope_acc = 0.f;
    int it = 0;
    const bool prof = (a.debug & 32) != 0;
    long long e_wait = 0, e_work = 0, e_t0 = clock64(), e_ld = 0, e_math = 0, e_st = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
      if ((it & 1) != egroup) continue;
      const TileCoord tc = decode_tile(a, tile);
      const int4 grp = grp_in_smem ? grp_s[tc.g] : __ldg(a.group_tab + tc.g);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int y = tc.y0 + ry, x = tc.x0 + rx;
      const bool valid = (y < a.H) && (x < a.W);
      const size_t rowoff =
          ((static_cast<size_t>(tc.n) * a.H + y) * a.W + x) * static_cast<size_t>(a.Cout) + grp.x;
      long long ec0 = 0, ec1 = 0;
      if (prof) ec0 = clock64();
      ptx::mbar_wait(tfull_bar + 8 * buf, bphase);
      ptx::tc_fence_after();
      if (prof) { ec1 = clock64(); e_wait += ec1 - ec0; }
      const uint32_t taddr =
          tmem_base + static_cast<uint32_t>(buf * a.nt) + (static_cast<uint32_t>(quarter * 32) << 16);
      if (a.debug & 1) {
        // attribution run: no epilogue work
      } else if ((a.nt & 63) == 0) {
        // ---- v2: 64 columns at a time, coalesced global traffic through the staging tile ----
        const uint32_t stg = stg_base + (warp - 2) * kStageTileBytes;
        size_t ro[8];
        uint32_t vmask = 0;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = quarter * 32 + it * 4 + (lane >> 3);
          const int yy = tc.y0 + (rr >> a.bw_shift), xx = tc.x0 + (rr & (a.bw - 1));
          const bool ok = (yy < a.H) && (xx < a.W);
          vmask |= (ok ? 1u : 0u) << it;
          ro[it] = ((static_cast<size_t>(tc.n) * a.H + yy) * a.W + xx) * static_cast<size_t>(a.Cout) + grp.x;
        }
        for (int c = 0; c < a.nt; c += 64) {
          float v[64];
          long long q0 = 0, q1 = 0, q2 = 0;
          if (prof) q0 = clock64();
          {
            uint32_t r[64];
            ptx::tmem_ld64(taddr + c, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
          }
          if (prof) q1 = clock64();
          if (epi & VSR_EPI_BIAS) {
            const float4* bp = bias_in_smem ? reinterpret_cast<const float4*>(bias_s + grp.x + c)
                                            : reinterpret_cast<const float4*>(a.bias + grp.x + c);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 b = bp[i];
              v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          }
          if (epi & VSR_EPI_SCALE) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] *= a.out_scale;
          }
          if (epi & VSR_EPI_RES_PRE) {
            float f[64];
            load_rows64(a.residual, ro, vmask, c, stg, lane, f);
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] += f[i];
          }
          if (epi & VSR_EPI_PRELU) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = v[i] > 0.f ? v[i] : slope * v[i];
          }
          if (epi & VSR_EPI_RELU) {
#pragma unroll
            for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) {
            float f[64];
            load_rows64(a.aux_y, ro, vmask, c, stg, lane, f);
            if (epi & VSR_EPI_PRELU_BWD) {
#pragma unroll
              for (int i = 0; i < 64; ++i) {
                const bool pos = f[i] > 0.f;
                if (valid) slope_acc += pos ? 0.f : v[i] * (f[i] * inv_slope);
                v[i] = pos ? v[i] : slope * v[i];
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i) v[i] = f[i] > 0.f ? v[i] : 0.f;
            }
          }
          if (prof) q2 = clock64();
          store_rows64(a.out, ro, (a.debug & 16) ? 0u : vmask, c, stg, lane, v);
          if (prof) { e_ld += q1 - q0; e_math += q2 - q1; e_st += clock64() - q2; }
          if (epi & VSR_EPI_OUT2) {
            float f[64];
            load_rows64(a.res2, ro, vmask, c, stg, lane, f);
#pragma unroll
            for (int i = 0; i < 64; ++i) f[i] += v[i];
            store_rows64(a.out2, ro, vmask, c, stg, lane, f);
          }
        }
      } else
      for (int c = 0; c < a.nt; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(taddr + c, r);
        ptx::tmem_ld_wait();
        if (valid) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          if (epi & VSR_EPI_BIAS) {
            const float4* bp = reinterpret_cast<const float4*>(a.bias + grp.x + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b = __ldg(bp + i);
              v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
            }
          }
          if (epi & VSR_EPI_SCALE) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= a.out_scale;
          }
          if (epi & VSR_EPI_RES_PRE) {
            const uint4* rp = reinterpret_cast<const uint4*>(a.residual + rowoff + c);
            float f[16];
            unpack8(__ldg(rp), f);
            unpack8(__ldg(rp + 1), f + 8);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += f[i];
          }
          if (epi & VSR_EPI_PRELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : slope * v[i];
          }
          if (epi & VSR_EPI_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) {
            const uint4* yp = reinterpret_cast<const uint4*>(a.aux_y + rowoff + c);
            float f[16];
            unpack8(__ldg(yp), f);
            unpack8(__ldg(yp + 1), f + 8);
            if (epi & VSR_EPI_PRELU_BWD) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const bool pos = f[i] > 0.f;
                slope_acc += pos ? 0.f : v[i] * (f[i] * inv_slope);
                v[i] = pos ? v[i] : slope * v[i];
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = f[i] > 0.f ? v[i] : 0.f;
            }
          }
          uint4* op = reinterpret_cast<uint4*>(a.out + rowoff + c);
          op[0] = pack8(v);
          op[1] = pack8(v + 8);
          if (epi & VSR_EPI_OUT2) {
            const uint4* rp = reinterpret_cast<const uint4*>(a.res2 + rowoff + c);
            float f[16];
            unpack8(__ldg(rp), f);
            unpack8(__ldg(rp + 1), f + 8);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += v[i];
            uint4* o2 = reinterpret_cast<uint4*>(a.out2 + rowoff + c);
            o2[0] = pack8(f);
            o2[1] = pack8(f + 8);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(tempty_bar + 8 * buf);
      if (prof) e_work += clock64() - ec1;
    }
    if (prof && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6))
      printf("tc-prof epilogue warp %d: total %lld cyc, wait(tfull) %lld, work %lld (tmem-ld %lld, math %lld, store %lld)\n", warp, clock64() - e_t0, e_wait, e_work, e_ld, e_math, e_st);
    if (epi & VSR_EPI_PRELU_BWD) {
      slope_acc = warp_sum(slope_acc);
      if (lane == 0) red[warp - 2] = slope_acc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0)
        a.slope_partials[blockIdx.x] =
            ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int n, h, w, c, bw, bh;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && n == o.n && h == o.h && w == o.w && c == o.c && bw == o.bw && bh == o.bh;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.n); mix(k.h); mix(k.w); mix(k.c); mix(k.bw); mix(k.bh);
    return h;
  }
};

std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

// 4-D bf16 map over a dense [n][h][w][c] map; box = 64 channels x bw x bh x 1, 128B swizzle.
int get_src_map(const VsrTensor4& t, int bw, int bh, CUtensorMap* out) {
  MapKey key{t.ptr, t.n, t.h, t.w, t.c, bw, bh};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
      *out = it->second;
      return VSR_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return VSR_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)t.c, (cuuint64_t)t.w, (cuuint64_t)t.h, (cuuint64_t)t.n};
  cuuint64_t strides[3] = {(cuuint64_t)t.c * 2, (cuuint64_t)t.w * t.c * 2,
                           (cuuint64_t)t.h * t.w * t.c * 2};
  cuuint32_t box[4] = {(cuuint32_t)kKc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] box %dx%d", (int)r, t.n, t.h,
              t.w, t.c, bw, bh);
    return VSR_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_map_cache.size() > 16384) g_map_cache.clear();
    g_map_cache.emplace(key, m);
  }
  *out = m;
  return VSR_OK;
}

// pick the pixel box (bw x bh = 128) that wastes the fewest MMA rows
void pick_box(int h, int w, int* bw_out, int* bh_out) {
  long best = -1;
  int best_bw = 128;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    const int bh = kBlockM / bw;
    if (bh > 256) break;
    const long tx = (w + bw - 1) / bw, ty = (h + bh - 1) / bh;
    const long cost = tx * ty;
    // ties: the widest box for maps up to 32 pixels wide (what the headline shapes were tuned on), the squarest one
    // for wider maps - less halo per tile for tables with row shifts (64x64 frames of the Conv3d path: 30.4 -> 28.0
    // ms per step; 32x32: +1.4 %, DRFNet step -0.5 %).  VSR_TC_SQUARE=0/1 forces either.
    static const char* env_sq = getenv("VSR_TC_SQUARE");
    const bool square = env_sq ? env_sq[0] == '1' : w > 32;
    if (best < 0 || cost < best || (square && cost == best && bw >= 8 && bw >= bh)) {
      best = cost;
      best_bw = bw;
    }
  }
  *bw_out = best_bw;
  *bh_out = kBlockM / best_bw;
}

}  // namespace

int get_src_map_pub(const VsrTensor4& t, int bw, int bh, CUtensorMap* out) { return get_src_map(t, bw, bh, out); }
void pick_box_pub(int h, int w, int* bw, int* bh) { pick_box(h, w, bw, bh); }

int tapgemm_tc2_launch(const VsrTapGemmDesc* d, cudaStream_t stream);   // tapgemm_tc2.cu

int tapgemm_tc_launch(const VsrTapGemmDesc* d, cudaStream_t stream) {
  {
    // second-generation kernel for every shape it covers; VSR_TC_V1=1 keeps this one (A/B timing)
    const char* env_v1 = getenv("VSR_TC_V1");
    if (!(env_v1 && env_v1[0] == '1') && d->kc == kKc && d->nt >= 64 && d->nt <= 256 && d->nt % 64 == 0 &&
        d->out.c >= 64)
      return tapgemm_tc2_launch(d, stream);
  }
  VSR_CHECK_SUPPORTED(d->kc == kKc, "tapgemm(bf16): kc must be 64, got %d", d->kc);
  VSR_CHECK_SUPPORTED(d->nt >= 16 && d->nt <= 256 && d->nt % 16 == 0,
                      "tapgemm(bf16): nt must be a multiple of 16 in [16,256], got %d", d->nt);
  VSR_CHECK_ARG(d->out.c % 8 == 0, "tapgemm(bf16): out.c must be a multiple of 8");
  typedef void (*KernelFn)(const TcArgs);
  static const struct { int epi; KernelFn fn; } kVariants[] = {
      {-1, tapgemm_tc_kernel<-1>},
      {0, tapgemm_tc_kernel<0>},
      {VSR_EPI_BIAS, tapgemm_tc_kernel<VSR_EPI_BIAS>},
      {VSR_EPI_RES_PRE, tapgemm_tc_kernel<VSR_EPI_RES_PRE>},
      {VSR_EPI_BIAS | VSR_EPI_PRELU, tapgemm_tc_kernel<VSR_EPI_BIAS | VSR_EPI_PRELU>},
      {VSR_EPI_PRELU_BWD, tapgemm_tc_kernel<VSR_EPI_PRELU_BWD>},
      {VSR_EPI_PRELU_BWD | VSR_EPI_RES_PRE, tapgemm_tc_kernel<VSR_EPI_PRELU_BWD | VSR_EPI_RES_PRE>},
  };
  static bool attr_set = false;
  if (!attr_set) {
    for (const auto& v : kVariants) {
      cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
        return VSR_ERR_CUDA;
      }
    }
    attr_set = true;
  }
  KernelFn kernel = kVariants[0].fn;
  for (const auto& v : kVariants)
    if (v.epi == d->epi) kernel = v.fn;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  int bw, bh;
  pick_box(d->out.h, d->out.w, &bw, &bh);
  for (int s = 0; s < d->n_srcs; ++s) {
    VSR_CHECK_ARG(d->srcs[s].c % 8 == 0, "tapgemm(bf16): src channels must be a multiple of 8");
    int rc = get_src_map(d->srcs[s], bw, bh, &a.maps[s]);
    if (rc != VSR_OK) return rc;
  }
  a.tap_tab = reinterpret_cast<const int4*>(d->tap_tab);
  a.group_tab = reinterpret_cast<const int4*>(d->group_tab);
  a.w = static_cast<const uint8_t*>(d->w);
  a.bias = d->bias;
  a.slope = d->slope;
  a.residual = static_cast<const __nv_bfloat16*>(d->residual);
  a.aux_y = static_cast<const __nv_bfloat16*>(d->aux_y);
  a.out = static_cast<__nv_bfloat16*>(d->out.ptr);
  a.out2 = static_cast<__nv_bfloat16*>(d->out2);
  a.res2 = static_cast<const __nv_bfloat16*>(d->res2);
  a.slope_partials = d->slope_partials;
  a.out_scale = d->out_scale;
  a.epi = d->epi;
  a.nt = d->nt;
  a.n_groups = d->n_groups;
  a.N = d->out.n; a.H = d->out.h; a.W = d->out.w; a.Cout = d->out.c;
  a.bw = bw; a.bh = bh;
  a.bw_shift = 0;
  while ((1 << a.bw_shift) < bw) ++a.bw_shift;
  a.tiles_x = (a.W + bw - 1) / bw;
  a.tiles_y = (a.H + bh - 1) / bh;
  const long tiles = (long)a.n_groups * a.N * a.tiles_x * a.tiles_y;
  VSR_CHECK_SUPPORTED(tiles < (1l << 30), "tapgemm(bf16): too many tiles");
  a.num_tiles = (int)tiles;
  a.n_taps_total = d->n_taps_total;
  a.m_tiles = a.N * a.tiles_x * a.tiles_y;
  const int b_bytes = d->nt * 128;
  const long res_need = (long)d->max_group_taps * b_bytes;
  a.resident = d->max_group_taps > 0 && res_need <= kSmemBudget - kCtrlBytes - 4 * kATileBytes;
  static const char* env_res = getenv("VSR_TC_RESIDENT");      // tuning overrides (not part of the ABI)
  static const char* env_stg = getenv("VSR_TC_STAGES");
  static const char* env_grid = getenv("VSR_TC_GRID");
  if (!(env_res && env_res[0] == '1')) a.resident = 0;   // opt-in: measured no gain (profiles/README.md)
  {
    const char* env_dbg = getenv("VSR_TC_DEBUG");            // re-read per launch: attribution sweeps flip it
    a.debug = env_dbg ? atoi(env_dbg) : 0;
    const char* env_res2 = getenv("VSR_TC_RESIDENT2");
    if (env_res2 && env_res2[0] == '1' && d->max_group_taps > 0 &&
        res_need <= kSmemBudget - kCtrlBytes - 4 * kATileBytes) a.resident = 1;
  }
  a.res_bytes = a.resident ? (int)res_need : 0;
  const int stage_bytes = a.resident ? kATileBytes : kATileBytes + b_bytes;
  int stages = (kSmemBudget - kCtrlBytes - a.res_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (env_stg && atoi(env_stg) >= 1 && atoi(env_stg) < stages) stages = atoi(env_stg);
  if (stages < 1) stages = 1;
  a.stages = stages;
  const int smem = kCtrlBytes + a.res_bytes + stages * stage_bytes;
  int grid = num_sms();
  if (env_grid && atoi(env_grid) >= 1) grid = atoi(env_grid);
  if (grid > a.num_tiles) grid = a.num_tiles;
  if (grid > kPartialsLen) grid = kPartialsLen;
  kernel<<<grid, kThreads, smem, stream>>>(a);
  VSR_CHECK_LAUNCH("tapgemm_tc");
  return VSR_OK;
}

}  // namespace vsr
