// tapgemm_tc2.cu — bf16 tap-GEMM on tcgen05 tensor cores (sm_100a), second generation.
//
// Same contract as tapgemm_tc.cu (include/vsr_b200.h, vsr_tapgemm) for nt % 64 == 0.  What changed
// and why (profiles/r01_attrib_*.log): with ~227 KB of shared memory per CTA the L1 is gone, so every
// local-memory access (spilled registers, address arrays) and every plain global load in the epilogue
// was an L2 round trip; the epilogue, not the tensor pipe or HBM, bounded the kernel.  Here
//   * the kernel is specialised on the epilogue flag set (no dead variants in the instruction stream),
//   * four epilogue warps keep everything in registers (192 threads -> up to 255 registers each),
//   * outputs leave through TMA stores from a 128B-swizzled staging tile (no address arithmetic, no
//     predicates: the TMA unit clips partial tiles), epilogue operands (residual, saved activation,
//     second residual) arrive through TMA loads that are issued one 64-channel chunk ahead,
//   * the bias of up to 1024 output channels is staged in shared memory once per CTA.
// Producer / MMA roles, the tap and group tables, the weight-slab format and the optional
// weight-resident mode are those of the first kernel.
//
// Replaces aten.convolution / convolution_backward(data) under drf_net.py:55-106,141-147.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace vsr {

int get_src_map_pub(const VsrTensor4& t, int bw, int bh, CUtensorMap* out);   // tma_host.cu
int get_f32_map_pub(const VsrTensor4& t, int bw, int bh, CUtensorMap* out);   // fp32 map, box of 32 channels
void pick_box_pub(int h, int w, int* bw, int* bh);

namespace {

// timing attribution (tools/attrib.py: parts of the kernel switched off through VSR_TC_DEBUG bits, per-role clock64
// sums, printf) exists only in builds with -DVSR_ATTRIB (python -m vsr_b200.build --attrib); the product kernel carries
// none of it
#ifdef VSR_ATTRIB
constexpr bool kAttrib = true;
#else
constexpr bool kAttrib = false;
#endif

constexpr int kBlockM = 128;
constexpr int kKc = 64;                       // bf16 channels per tap = one 128-byte row
constexpr int kATileBytes = kBlockM * 128;    // 16 KiB
constexpr int kMaxStages = 12;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kEpiWarps = 8;                  // two per TMEM lane quarter: even / odd 64-channel chunks
constexpr int kTileBytes = 4096;              // epilogue tile: 32 pixels x 64 channels bf16
constexpr int kMaxSmemGroups = 20;            // group table rows cached in smem (16 B each, ctrl[448..768))
constexpr int kMaxSmemTaps = 256;             // column entries cached in smem (8 B each)
constexpr int kMaxParamCols = 96;             // host-built columns travel in the kernel arguments
constexpr int kMaxTallGroups = 8;
constexpr int kBiasFloats = 1024;
constexpr int kCtrlBytes = 1024 + kMaxSmemTaps * 8 + kBiasFloats * 4;   // 7 KiB, keeps 1024-byte alignment
constexpr int kTmemCols = 512;
constexpr int kProducers = 3;                 // TMA-issuing warps: tap t of the CTA's sequence belongs to warp t % 3
constexpr int kMmaWarp = kProducers;
constexpr int kEpiWarp0 = kProducers + 1;      // first epilogue warp (a multiple of 4: TMEM lane quarter = warp & 3)
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);

constexpr int kEpiIn = VSR_EPI_RES_PRE | VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD | VSR_EPI_OUT2;
// internal flag (never in VsrTapGemmDesc.epi): the split-bf16 mode (dtype VSR_BF16X2) writes the raw fp32 accumulators -
// one 32-pixel x 64-channel chunk leaves as two [32 x 32 fp32] tiles through an fp32 tensor map; vsr_tap_epilogue
// applies bias / activation / residuals afterwards in fp32
constexpr int kEpiF32Out = 0x4000;

struct Tc2Args {
  CUtensorMap maps[VSR_MAX_SRCS];
  CUtensorMap tall_maps[VSR_MAX_SRCS];   // shared-load mode: box of mb*bh + ndy_max - 1 rows
  CUtensorMap out_map, res_map, aux_map, out2_map, res2_map;
  uint2 cols[kMaxParamCols];             // shared-load mode: {packed tap (dy = first row shift), slab0 | stride<<12 | ndy<<24}
  int4 tgroups[kMaxTallGroups];          // shared-load mode: {o0, first column, columns, first tap | taps<<16}
  const int4* tap_tab;
  const int4* group_tab;
  const uint8_t* w;
  const float* bias;
  const float* slope;
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
  int contig;              // 1: every CTA walks a contiguous, group-major tile range (resident weights, > 1 group)
  int res_bytes;           // size of the resident weight region
  int epi_bytes;           // size of the epilogue tile region
  uint32_t mg_groups, mg_mtiles, mg_tx, mg_ty;   // ceil(2^32 / d) of n_groups, m_tiles, tiles_x, tiles_y (0: d == 1)
  int m_tiles;             // pixel tiles per group
  int debug;               // timing-attribution switches (VSR_TC_DEBUG); results are wrong when non-zero
  int tall;                // 1: shared-load mode (columns from `cols`, one group)
  int n_cols;              // columns in `cols`
  int mb;                  // 128-pixel sub-tiles stacked in y per CTA tile (1 or 2)
  int a_bytes;             // bytes of one A box
  int row_bytes;           // bw * 128: bytes of one pixel row of the A box
  int stage_bytes;         // A box + the weight slabs of one column (none in resident mode)
  int epi_obuf;            // output staging tiles per epilogue warp (1 or 2: the store of a chunk overlaps the next chunk)
  int epi_ibuf;            // epilogue operand sets per warp (1 or 2: operands are fetched two chunks ahead)
  int pair;                // 1: CTA pairs (clusters of 2, tcgen05 cta_group::2): m_tiles / num_tiles count PAIRS of pixel tiles
};

static_assert(sizeof(Tc2Args) <= 4096, "kernel arguments are limited to 4 KB");

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
// n / d and n % d through a host-made reciprocal (magic = ceil(2^32 / d), 0 for d == 1): every role decodes
// every tile, and five emulated 32-bit divisions per tile were a third of the MMA warp's per-tile time on
// the HBM-bound 1x1 convolutions (profiles/README.md)
__device__ __forceinline__ void fast_divmod(int n, int d, uint32_t magic, int* q, int* r) {
  if (magic == 0u) {
    *q = n;
    *r = 0;
    return;
  }
  int qq = (int)__umulhi((uint32_t)n, magic);      // floor(n / d) or one more
  int rr = n - qq * d;
  if (rr < 0) {
    --qq;
    rr += d;
  }
  *q = qq;
  *r = rr;
}
// pair mode: `tile` counts pairs of pixel tiles, CTA `rank` of the pair takes pixel tile 2 * pair + rank (an odd tile
// count leaves the last pair's second CTA with n == N: its loads are zero-filled and its stores clipped by the TMA unit)
__device__ __forceinline__ TileCoord decode_tile(const Tc2Args& a, int tile, int rank) {
  TileCoord t;
  int mt, tx, ty;
  if (a.contig) {
    fast_divmod(tile, a.m_tiles, a.mg_mtiles, &t.g, &mt);
  } else {
    fast_divmod(tile, a.n_groups, a.mg_groups, &mt, &t.g);
  }
  if (a.pair) mt = 2 * mt + rank;
  fast_divmod(mt, a.tiles_x, a.mg_tx, &mt, &tx);
  fast_divmod(mt, a.tiles_y, a.mg_ty, &t.n, &ty);
  t.x0 = tx * a.bw;
  t.y0 = ty * a.bh * a.mb;
  return t;
}

// two fp32 additions in one instruction (FADD2: sm_100 packed fp32, same rounding as two FADDs)
__device__ __forceinline__ void add2(float& x, float& y, float bx, float by) {
  const float2 r = __fadd2_rn(make_float2(x, y), make_float2(bx, by));
  x = r.x;
  y = r.y;
}
// PReLU' of one element whose stored activation is f (slope > 0 or = 0: the branch is the sign of f): where f <= 0 the
// slope sum takes g * f and g is scaled by the slope.  Two predicated instructions behind one FSETP.
__device__ __forceinline__ void prelu_bwd_elem(float& acc, float& g, float f, float slope) {
  asm("{\n\t.reg .pred pp;\n\tsetp.gt.f32 pp, %2, 0f00000000;\n\t"
      "@!pp fma.rn.f32 %0, %1, %2, %0;\n\t@!pp mul.f32 %1, %1, %3;\n\t}"
      : "+f"(acc), "+f"(g) : "f"(f), "f"(slope));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// 128B-swizzled [32 pixels x 64 channels] tile (the TMA SWIZZLE_128B image): 16-byte chunk q of row r
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int r, int q) {
  return base + r * 128 + ((q ^ (r & 7)) << 4);
}
__device__ __forceinline__ void tma_store_4d(const void* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::
                   "l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// FIXED_EPI >= 0: the epilogue flag set is a compile-time constant; FIXED_EPI < 0: read from the arguments.
// PAIR: two CTAs of one TPC (a cluster of 2) work on two pixel tiles of the same group with ONE stream of M = 256
// tcgen05.mma.cta_group::2 instructions issued by the leader CTA (rank 0): every CTA loads its own A box and only HALF of
// every weight slab (rows [rank * nt/2, +nt/2)), so the weight traffic L2 -> SM, the resident-weight footprint and the
// shared-memory reads per MMA of the B operand halve.  Both CTAs run every role on their own tile; the peer's MMA warp
// relays "my stage is loaded" to the leader (remote mbarrier arrive), the leader's commits arrive in both CTAs
// (multicast), and the peer's epilogue frees the accumulator on the leader's barrier.
template <int FIXED_EPI, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) tapgemm_tc2_kernel(const __grid_constant__ Tc2Args a) {
  const int epi = FIXED_EPI >= 0 ? FIXED_EPI : a.epi;
  const int dbg = kAttrib ? a.debug : 0;
  const uint64_t g_t0 = kAttrib ? ptx::globaltimer_ns() : 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  uint8_t* smem_gen = smem_raw;
  if (smem_base & 1023u) __trap();             // the swizzled tiles need a 1024-byte aligned window

  // control block
  const uint32_t full_bar = smem_base;                     // kMaxStages x 8 B (room for 16)
  const uint32_t empty_bar = smem_base + 128;              // kMaxStages x 8 B (room for 16)
  const uint32_t tfull_bar = smem_base + 256;              // 2 x 8 B
  const uint32_t tempty_bar = smem_base + 272;             // 2 x 8 B
  const uint32_t tmem_slot = smem_base + 288;              // u32
  const uint32_t bres_full = smem_base + 296, bres_empty = smem_base + 304;
  const uint32_t in_bar0 = smem_base + 768;                // kEpiWarps x 2 x 8 B (..896): operand sets of the epilogue warps
  const uint32_t peer_bres_full = smem_base + 416;         // pair mode, leader: the peer's resident slabs have landed
  const uint32_t peer_full_bar = smem_base + 896;          // pair mode, leader: kMaxStages x 8 B, the peer's stage is loaded
  const int cta_rank = PAIR ? (int)ptx::cluster_ctarank() : 0;
  const bool leader_cta = cta_rank == 0;
  const int walker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // tile walker: a CTA, or a CTA pair
  const int n_walkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + 288);
  float* red = reinterpret_cast<float*>(smem_gen + 384);   // kEpiWarps floats
  int4* grp_s = reinterpret_cast<int4*>(smem_gen + 448);        // group table (<= 36 rows)
  uint2* col_s = reinterpret_cast<uint2*>(smem_gen + 1024);     // column table (<= 256 entries)
  float* bias_s = reinterpret_cast<float*>(smem_gen + 1024 + kMaxSmemTaps * 8);   // bias when Cout <= 1024
  const uint32_t epi_base = smem_base + kCtrlBytes;        // epilogue tiles
  const uint32_t res_base = epi_base + a.epi_bytes;        // resident weight slabs (resident mode)
  const uint32_t stage_base = res_base + a.res_bytes;
  const bool grp_in_smem = a.n_groups <= kMaxSmemGroups;
  // (shared-load mode: the host-built columns - at most kMaxParamCols - always sit in shared memory, however long the tap
  //  table is; only the plain mode falls back to the global tap table for tables beyond kMaxSmemTaps)
  const bool taps_in_smem = a.tall || a.n_taps_total <= kMaxSmemTaps;
  // A "column" is one A load feeding ndy taps (weight slabs slab0 + j*stride) whose row shifts are
  // dy0 + j; plain mode: every tap is a column of its own.
  if (a.tall) {
    if (threadIdx.x < a.n_groups) {
      int4 g = a.tgroups[threadIdx.x];
      if (g.x < 0) g.x = __ldg(a.group_tab + threadIdx.x).x;      // single group: the slice start lives in the device table
      grp_s[threadIdx.x] = g;
    }
    for (int i = threadIdx.x; i < a.n_cols; i += blockDim.x) col_s[i] = a.cols[i];
  } else {
    if (grp_in_smem)
      for (int i = threadIdx.x; i < a.n_groups; i += blockDim.x) grp_s[i] = __ldg(a.group_tab + i);
    if (taps_in_smem)
      for (int i = threadIdx.x; i < a.n_taps_total; i += blockDim.x)
        col_s[i] = make_uint2(pack_tap(__ldg(a.tap_tab + i)), (uint32_t)i | (1u << 24));
  }
  const bool bias_in_smem = (epi & VSR_EPI_BIAS) && a.Cout <= kBiasFloats;
  const uint32_t b_full = static_cast<uint32_t>(a.nt) * 128u;      // one weight slab in global memory
  const uint32_t b_bytes = PAIR ? b_full / 2 : b_full;             // the rows of it this CTA holds
  const uint8_t* const wbase = a.w + (PAIR ? cta_rank * b_bytes : 0u);
  const uint32_t a_bytes = static_cast<uint32_t>(a.a_bytes);
  const uint32_t stage_bytes = static_cast<uint32_t>(a.stage_bytes);
  // tile walk of this CTA
  // (one group: resident weights never change, so the CTAs keep the interleaved walk whose concurrent
  //  tiles are neighbours in memory - measured 5-10 % faster than contiguous ranges on the 1x1 convolutions)
  const int tile_begin = a.contig ? (int)((long)a.num_tiles * walker / n_walkers) : walker;
  const int tile_end = a.contig ? (int)((long)a.num_tiles * (walker + 1) / n_walkers) : a.num_tiles;
  const int tile_step = a.contig ? 1 : n_walkers;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool prof = (dbg & 32) != 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(full_bar + 8 * s, 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar + 8 * b, 1);
      // (pair mode: the epilogue warps of BOTH CTAs arrive on the leader's barrier)
      ptx::mbar_init(tempty_bar + 8 * b, (a.nt > 64 ? kEpiWarps : kEpiWarps / 2) * (PAIR ? 2 : 1));
    }
    ptx::mbar_init(bres_full, 1);
    ptx::mbar_init(bres_empty, 1);
    if (PAIR) {
      for (int s = 0; s < a.stages; ++s) ptx::mbar_init(peer_full_bar + 8 * s, 1);
      ptx::mbar_init(peer_bres_full, 1);
    }
    for (int w = 0; w < 2 * kEpiWarps; ++w) ptx::mbar_init(in_bar0 + 8 * w, 1);
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    if (PAIR) {
      ptx::tmem_alloc2(tmem_slot, kTmemCols);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(tmem_slot, kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  // Everything above reads only launch-time constants (tables uploaded when the layer plan was built,
  // kernel arguments) and this CTA's own shared / tensor memory; tensors written by earlier kernels
  // are touched only below this point.
  // The CTA-wide (pair-wide) synchronisation belongs to the prologue as well: after pdl_wait() the producers issue their
  // first loads at once.  (The bias - packed by an earlier kernel of the step - used to be staged by all threads between the
  // wait and this synchronisation: one global-load latency on the critical path of every launch.  Only the epilogue warps
  // need it; they stage it among themselves below while the first loads and MMAs are in flight.)
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync();       // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  ptx::pdl_wait();
  ptx::pdl_launch_dependents();
  const uint64_t g_t1 = kAttrib ? ptx::globaltimer_ns() : 0;

  if (warp < kProducers) {
    // ===================== TMA producers =====================
    // The whole warp walks the (uniform) loops; one elected lane issues.  Keeping the control flow and
    // the operands warp-uniform lets the compiler stay on the uniform datapath (no per-lane loops).
    {
      const bool leader = ptx::elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int cur_g = -1;
      uint32_t gcount = 0;
      // Producers take the columns in turn.  A producer only waits for the slots it fills itself, so it
      // must not get a whole ring ahead of a slot it never waited on (the parity of an mbarrier phase
      // aliases after two completions): at most `stages` producers may be active.
      const int n_prod = a.stages < kProducers ? a.stages : kProducers;
      int turn = 0;                               // position in the column sequence modulo n_prod
      long long p_wait = 0, p_n = 0, p_t0 = kAttrib ? clock64() : 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const TileCoord tc = decode_tile(a, tile, cta_rank);
        int4 grp = grp_in_smem ? grp_s[tc.g] : __ldg(a.group_tab + tc.g);
        grp.y = __shfl_sync(0xffffffffu, grp.y, 0);
        grp.z = __shfl_sync(0xffffffffu, grp.z, 0);
        grp.w = __shfl_sync(0xffffffffu, grp.w, 0);
        if (a.resident && tc.g != cur_g && warp == 0) {
          // (re)load this group's weight slabs once all MMAs of the previous group are done
          const int tap0 = a.tall ? (grp.w & 0xffff) : grp.y, ntap = a.tall ? (grp.w >> 16) : grp.z;
          ptx::mbar_wait(bres_empty, (gcount & 1u) ^ 1u);
          if (leader) {
            ptx::mbar_arrive_expect_tx(bres_full, static_cast<uint32_t>(ntap) * b_bytes);
            for (int t = 0; t < ntap; ++t)
              ptx::bulk_load(res_base + t * b_bytes, wbase + static_cast<size_t>(tap0 + t) * b_full, b_bytes, bres_full);
          }
          cur_g = tc.g;
          ++gcount;
        }
        for (int t = 0; t < grp.z; ++t) {
          const bool mine = turn == warp;
          if (++turn == n_prod) turn = 0;
          if (!mine) {
            if (++stage == a.stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          uint2 col = taps_in_smem ? col_s[grp.y + t]
                                   : make_uint2(pack_tap(__ldg(a.tap_tab + grp.y + t)), (uint32_t)(grp.y + t) | (1u << 24));
          col.x = __shfl_sync(0xffffffffu, col.x, 0);
          col.y = __shfl_sync(0xffffffffu, col.y, 0);
          const int4 tap = unpack_tap(col.x);
          const int slab0 = (int)(col.y & 0xfffu), sstride = (int)((col.y >> 12) & 0xfffu), ndy = (int)(col.y >> 24);
          long long c0 = 0;
          if (prof) c0 = clock64();
          ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1u);
          if (prof) { p_wait += clock64() - c0; ++p_n; }
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t sa = stage_base + stage * stage_bytes;
          // (split-bf16 mode: bit 3 of the source index selects the low-order plane = images [N, 2N) of the map)
          const CUtensorMap* map = (a.tall ? a.tall_maps : a.maps) + (tap.x & 7);
          const int n_img = tc.n + (tap.x >> 3) * a.N;
          if (leader) {
            // attribution runs skip the A (debug & 2) and/or B (debug & 4) transfer, keeping the protocol
            const bool do_a = !(dbg & 2), do_b = !(dbg & 4) && !a.resident;
            const uint32_t tx = (do_a ? a_bytes : 0u) + (do_b ? static_cast<uint32_t>(ndy) * b_bytes : 0u);
            if (tx == 0) { ptx::mbar_arrive(fb); } else { ptx::mbar_arrive_expect_tx(fb, tx); }
            if (do_a) ptx::tma_load_4d(sa, map, fb, tap.w, tc.x0 + tap.z, tc.y0 + tap.y, n_img);
            if (do_b)
              for (int j = 0; j < ndy; ++j)
                ptx::bulk_load(sa + a_bytes + j * b_bytes, wbase + static_cast<size_t>(slab0 + j * sstride) * b_full, b_bytes, fb);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
      }
      if (prof && blockIdx.x == 0 && leader && warp == 0)
        printf("tc2-prof producer: total %lld cyc, %lld taps, wait(empty) %lld\n", clock64() - p_t0, p_n, p_wait);
    }
  } else if (warp == kMmaWarp && PAIR && !leader_cta) {
    // ===================== pair mode, peer CTA: relay "my stage is loaded" to the leader =====================
    // Same walk as the leader's MMA issuer; instead of issuing, arrive on the leader's peer_full barrier of the stage
    // (and on peer_bres_full when this CTA's resident slabs have landed).  The stage cannot be refilled before the
    // leader's commit of its MMAs arrives on this CTA's empty barrier, so no arrival is ever ahead by a whole phase.
    {
      const bool leader = ptx::elect_one();
      const uint32_t r_full = ptx::mapa(peer_full_bar, 0), r_bres = ptx::mapa(peer_bres_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      int cur_g = -1;
      uint32_t gcount = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const TileCoord tc = decode_tile(a, tile, cta_rank);
        int4 grp = grp_in_smem ? grp_s[tc.g] : __ldg(a.group_tab + tc.g);
        grp.z = __shfl_sync(0xffffffffu, grp.z, 0);
        if (a.resident && tc.g != cur_g) {
          ptx::mbar_wait(bres_full, gcount & 1u);
          if (leader) ptx::mbar_arrive_remote(r_bres);
          cur_g = tc.g;
          ++gcount;
        }
        for (int t = 0; t < grp.z; ++t) {
          ptx::mbar_wait(full_bar + 8 * stage, phase);
          if (leader) ptx::mbar_arrive_remote(r_full + 8 * stage);
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (one elected lane, warp-uniform control flow) =====================
    {
      const bool leader = ptx::elect_one();
      const uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, a.nt, 0, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int cur_g = -1;
      uint32_t gcount = 0;
      long long m_wfull = 0, m_wtmem = 0, m_issue = 0, m_n = 0, m_t0 = kAttrib ? clock64() : 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
        const TileCoord tc = decode_tile(a, tile, cta_rank);
        int4 grp = grp_in_smem ? grp_s[tc.g] : __ldg(a.group_tab + tc.g);
        grp.y = __shfl_sync(0xffffffffu, grp.y, 0);
        grp.z = __shfl_sync(0xffffffffu, grp.z, 0);
        grp.w = __shfl_sync(0xffffffffu, grp.w, 0);
        const int tap0 = a.tall ? (grp.w & 0xffff) : grp.y;      // first weight slab of the group
        if (a.resident && tc.g != cur_g) {
          ptx::mbar_wait(bres_full, gcount & 1u);
          if (PAIR) ptx::mbar_wait(peer_bres_full, gcount & 1u);
          cur_g = tc.g;
          ++gcount;
        }
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        long long c0 = 0, c1 = 0;
        if (prof) c0 = clock64();
        ptx::mbar_wait(tempty_bar + 8 * buf, bphase ^ 1u);
        ptx::tc_fence_after();
        if (prof) m_wtmem += clock64() - c0;
        const uint32_t d_tmem = tmem_u + static_cast<uint32_t>(buf * a.mb * a.nt);
        for (int t = 0; t < grp.z; ++t) {
          uint2 col = taps_in_smem ? col_s[grp.y + t] : make_uint2(0u, (uint32_t)(grp.y + t) | (1u << 24));
          col.y = __shfl_sync(0xffffffffu, col.y, 0);
          const int slab0 = (int)(col.y & 0xfffu), sstride = (int)((col.y >> 12) & 0xfffu), ndy = (int)(col.y >> 24);
          if (prof) c0 = clock64();
          ptx::mbar_wait(full_bar + 8 * stage, phase);
          if (PAIR) ptx::mbar_wait(peer_full_bar + 8 * stage, phase);
          ptx::tc_fence_after();
          if (prof) c1 = clock64();
          const uint32_t sa = stage_base + stage * stage_bytes;
          if (leader) {
            if (!(dbg & 8)) {
              for (int j = 0; j < ndy; ++j) {
                const uint64_t bdesc = ptx::make_sw128_desc(
                    a.resident ? res_base + (slab0 + j * sstride - tap0) * b_bytes : sa + a_bytes + j * b_bytes, 16, 1024);
                for (int m = 0; m < a.mb; ++m) {
                  // rows shifted by j (tap) and m*bh (sub-tile) inside the shared A box: whole pixel rows,
                  // i.e. multiples of 1024 bytes, so the swizzle phase of the descriptor is unchanged
                  const uint64_t adesc = ptx::make_sw128_desc(sa + (j + m * a.bh) * a.row_bytes, 16, 1024);
#pragma unroll
                  for (int k = 0; k < kKc / 16; ++k) {
                    // advancing K by 16 bf16 = 32 bytes = 2 descriptor address units
                    if (PAIR) ptx::mma_bf16_ss2(d_tmem + m * a.nt, adesc + 2 * k, bdesc + 2 * k, idesc, (t | j | k) != 0);
                    else ptx::mma_bf16_ss(d_tmem + m * a.nt, adesc + 2 * k, bdesc + 2 * k, idesc, (t | j | k) != 0);
                  }
                }
              }
            }
            if (PAIR) ptx::mma_commit2(empty_bar + 8 * stage);     // frees the stage in both CTAs
            else ptx::mma_commit(empty_bar + 8 * stage);
          }
          if (prof) { m_wfull += c1 - c0; m_issue += clock64() - c1; ++m_n; }
          if (++stage == a.stages) { stage = 0; phase ^= 1u; }
        }
        if (leader) {
          if (PAIR) ptx::mma_commit2(tfull_bar + 8 * buf);
          else ptx::mma_commit(tfull_bar + 8 * buf);
        }
        if (a.resident) {
          const int next = tile + tile_step;
          if ((next >= tile_end || decode_tile(a, next, cta_rank).g != cur_g) && leader) {
            if (PAIR) ptx::mma_commit2(bres_empty);
            else ptx::mma_commit(bres_empty);
          }
        }
      }
      if (prof && blockIdx.x == 0 && leader)
        printf("tc2-prof mma: total %lld cyc, %lld taps, %d tiles, wait(full) %lld, issue+commit %lld, wait(tmem) %lld\n",
               clock64() - m_t0, m_n, it, m_wfull, m_issue, m_wtmem);
    }
  } else {
    // ===================== epilogue (8 warps: TMEM lane quarter x chunk parity) =====================
    const int ew = warp - kEpiWarp0;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32) = tile pixels
    // nt > 64: the two warps of a lane quarter split every tile by chunk parity (c = 64*chalf mod 128);
    // nt == 64 (one chunk per tile): they alternate tiles instead, i.e. warp set `chalf` owns TMEM
    // buffer `chalf`, which doubles the time each warp has to hide its operand loads
    const int chalf = ew >> 2;
    const bool by_tile = a.nt <= 64;
    const int c_first = by_tile ? 0 : 64 * chalf;
    const int my_step = by_tile ? 2 * tile_step : tile_step;       // distance between this warp's tiles
    const int my_begin = by_tile ? tile_begin + chalf * tile_step : tile_begin;
    // this warp's 32 pixels as a sub-box of the bw x bh tile (the epilogue maps have box ew x eh)
    const int sub_x = (quarter * 32) & (a.bw - 1);
    const int sub_y = (quarter * 32) >> a.bw_shift;
    const bool skip = (dbg & 1) != 0;
    const bool has_in = (epi & kEpiIn) != 0 && !skip;
    const int n_in = ((epi & VSR_EPI_RES_PRE) ? 1 : 0) + ((epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) ? 1 : 0) +
                     ((epi & VSR_EPI_OUT2) ? 1 : 0);
    // staging of this warp: [obuf output tiles][ibuf sets of n_in operand tiles].  With two output tiles the TMA store of
    // a chunk reads its tile while the next chunk is computed; with two operand sets the residual / saved-activation
    // tiles of chunk q + 2 are fetched while chunks q and q + 1 are processed (one set: chunk q + 1 while q is stored,
    // which exposed most of the L2 latency in the chunk-parity mode - the PReLU' data gradients ran 15 us behind
    // their forward twins).
    const int obuf = a.epi_obuf, ibuf = a.epi_ibuf;
    const uint32_t warp_t = epi_base + ew * (obuf + ibuf * n_in) * kTileBytes;
    const uint32_t in_t0 = warp_t + obuf * kTileBytes;
    const uint32_t aux_o = (epi & VSR_EPI_RES_PRE) ? kTileBytes : 0;
    const uint32_t res2_o = aux_o + ((epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) ? kTileBytes : 0);
    const uint32_t in_bar = in_bar0 + 16 * ew;
    // (slope == 0 and slope < 0: see Prelu in common.cuh)
    const Prelu pr = make_prelu((epi & (VSR_EPI_PRELU | VSR_EPI_PRELU_BWD)) ? __ldg(a.slope) : 1.f);
    const bool slope01 = pr.fwd >= 0.f && pr.fwd <= 1.f;
    float slope_acc = 0.f, slope_acc1 = 0.f;      // (two chains: even / odd channels)
    long long e_wait = 0, e_in = 0, e_ld = 0, e_math = 0, e_st = 0, e_iss = 0, e_t0 = kAttrib ? clock64() : 0;

    // the chunks of this warp in processing order: (tile, sub-tile m, first channel c)
    struct Chunk { int tile, m, c; };
    auto next_chunk = [&](Chunk& k) {
      if (k.c + 128 < a.nt) { k.c += 128; return true; }
      k.c = c_first;
      if (k.m + 1 < a.mb) { ++k.m; return true; }
      k.m = 0;
      k.tile += my_step;
      return k.tile < tile_end;
    };
    // one elected lane fetches the epilogue operands of a 64-channel chunk into operand set b through TMA
    auto issue_chunk = [&](const Chunk& k, int b) {
      const TileCoord t = decode_tile(a, k.tile, cta_rank);
      const int4 g = grp_in_smem ? grp_s[t.g] : __ldg(a.group_tab + t.g);
      const uint32_t bar = in_bar + 8 * b, base = in_t0 + b * n_in * kTileBytes;
      const int c0 = g.x + k.c, x0 = t.x0 + sub_x, y0 = t.y0 + k.m * a.bh + sub_y;
      ptx::mbar_arrive_expect_tx(bar, static_cast<uint32_t>(n_in) * kTileBytes);
      if (epi & VSR_EPI_RES_PRE) ptx::tma_load_4d(base, &a.res_map, bar, c0, x0, y0, t.n);
      if (epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) ptx::tma_load_4d(base + aux_o, &a.aux_map, bar, c0, x0, y0, t.n);
      if (epi & VSR_EPI_OUT2) ptx::tma_load_4d(base + res2_o, &a.res2_map, bar, c0, x0, y0, t.n);
    };
    Chunk la = {my_begin, 0, c_first};             // look-ahead: the next chunk whose operands are to be fetched
    bool la_ok = has_in && my_begin < tile_end;
    for (int d = 0; d < ibuf && la_ok; ++d) {
      if (lane == 0) issue_chunk(la, d);
      la_ok = next_chunk(la);
    }
    // (after the operand prefetch above: the bias load must not delay it)
    if (bias_in_smem) {
      for (int i = ew * 32 + lane; i < a.Cout; i += kEpiWarps * 32) bias_s[i] = __ldg(a.bias + i);
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    uint32_t q = 0;                                // chunks processed by this warp
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++it) {
      if (by_tile && (it & 1) != chalf) continue;
      const TileCoord tc = decode_tile(a, tile, cta_rank);
      const int4 grp = grp_in_smem ? grp_s[tc.g] : __ldg(a.group_tab + tc.g);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      long long q0 = 0, q1 = 0;
      if (prof) q0 = clock64();
      ptx::mbar_wait(tfull_bar + 8 * buf, bphase);
      ptx::tc_fence_after();
      if (prof) e_wait += clock64() - q0;
      if (!skip) {
#pragma unroll 1
        for (int mc = 0; mc < a.mb; ++mc)
#pragma unroll 1
        for (int c = c_first; c < a.nt; c += 128) {
          const int m = mc;                       // 128-pixel sub-tile of the CTA tile
          const uint32_t taddr = tmem_base + static_cast<uint32_t>((buf * a.mb + m) * a.nt) +
                                 (static_cast<uint32_t>(quarter * 32) << 16);
          if (prof) q0 = clock64();
          const int ib = (int)(q & (uint32_t)(ibuf - 1)), ob = (int)(q & (uint32_t)(obuf - 1));
          const uint32_t out_t = warp_t + ob * kTileBytes;
          const uint32_t res_t = in_t0 + ib * n_in * kTileBytes, aux_t = res_t + aux_o, res2_t = res_t + res2_o;
          if (has_in) ptx::mbar_wait(in_bar + 8 * ib, (q >> (ibuf - 1)) & 1u);
          // the TMA store that last used this staging tile must have read it before it is rewritten
          if (lane == 0) {
            if (obuf == 2 && !(epi & kEpiF32Out)) bulk_wait_read1();
            else bulk_wait_read0();
          }
          __syncwarp();
          if (prof) { q1 = clock64(); e_in += q1 - q0; }
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {               // two 32-channel halves keep the live set small
            if (prof) q1 = clock64();
            float v[32];
            {
              uint32_t r[32];
              ptx::tmem_ld32(taddr + c + 32 * h, r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            }
            if (prof) { q0 = clock64(); e_ld += q0 - q1; }
            if (epi & VSR_EPI_BIAS) {
              if (bias_in_smem) {
                const float4* bp = reinterpret_cast<const float4*>(bias_s + grp.x + c + 32 * h);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = bp[i];
                  add2(v[4 * i], v[4 * i + 1], b.x, b.y);
                  add2(v[4 * i + 2], v[4 * i + 3], b.z, b.w);
                }
              } else {
                const float4* bp = reinterpret_cast<const float4*>(a.bias + grp.x + c + 32 * h);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b = __ldg(bp + i);
                  add2(v[4 * i], v[4 * i + 1], b.x, b.y);
                  add2(v[4 * i + 2], v[4 * i + 3], b.z, b.w);
                }
              }
            }
            if (epi & VSR_EPI_SCALE) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= a.out_scale;
            }
            if (epi & VSR_EPI_RES_PRE) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 rq = ld_shared_v4(tile_addr(res_t, lane, 4 * h + j));
                v[8 * j + 0] += bf16_lo(rq.x); v[8 * j + 1] += bf16_hi(rq.x);
                v[8 * j + 2] += bf16_lo(rq.y); v[8 * j + 3] += bf16_hi(rq.y);
                v[8 * j + 4] += bf16_lo(rq.z); v[8 * j + 5] += bf16_hi(rq.z);
                v[8 * j + 6] += bf16_lo(rq.w); v[8 * j + 7] += bf16_hi(rq.w);
              }
            }
            uint32_t posmask = 0u;                   // negative slope only: [x > 0] travels in the LSB of y
            if (epi & VSR_EPI_PRELU) {
              if (slope01) {
                // for 0 <= a <= 1: PReLU(v) = max(v, a*v) exactly (two instructions per element instead of three)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float2 t = __fmul2_rn(make_float2(v[i], v[i + 1]), make_float2(pr.fwd, pr.fwd));
                  v[i] = fmaxf(v[i], t.x);
                  v[i + 1] = fmaxf(v[i + 1], t.y);
                }
              } else {
                if (pr.tag) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) posmask |= (v[i] > 0.f ? 1u : 0u) << i;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : pr.fwd * v[i];
              }
            }
            if (epi & VSR_EPI_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) {
              // The branch of an element is decided once per warp-uniform case OUTSIDE the element loop and the two
              // updates are predicated on it (unpack, FSETP, @!pos FFMA, @!pos FMUL: ~4 instructions per element).  With
              // the tag test inside the loop and select-style arithmetic the compiler materialised every predicate as an
              // integer and re-tested it (~12 instructions per element: the epilogue warps, not the MMAs or the
              // memory system, bounded the PReLU' data gradients).
              // out-of-image pixels read f = 0 from the TMA zero fill and contribute exactly 0.
              // d(slope) = sum over x <= 0 of g * x with x = y / slope: accumulate g * y here and scale by 1 / slope
              // once per CTA.
              if ((epi & VSR_EPI_PRELU_BWD) && pr.tag) {         // negative slope: the branch travels in the LSB of y
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 aq = ld_shared_v4(tile_addr(aux_t, lane, 4 * h + j));
                  const uint32_t w4[4] = {aq.x, aq.y, aq.z, aq.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                      const int i = 8 * j + 2 * q + p;
                      const float f = p ? bf16_hi(w4[q]) : bf16_lo(w4[q]);
                      if (!(w4[q] & (1u << (16 * p)))) {
                        slope_acc = fmaf(v[i], f, slope_acc);
                        v[i] *= pr.slope;
                      }
                    }
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint4 aq = ld_shared_v4(tile_addr(aux_t, lane, 4 * h + j));
                  const uint32_t w4[4] = {aq.x, aq.y, aq.z, aq.w};
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                      const int i = 8 * j + 2 * q + p;
                      const float f = p ? bf16_hi(w4[q]) : bf16_lo(w4[q]);
                      if (epi & VSR_EPI_PRELU_BWD) {
                        // (written as predicated PTX: from the C++ form the compiler built FFMA + FSEL, one serial
                        // 8-cycle link per element in the slope sum)
                        if (p) prelu_bwd_elem(slope_acc1, v[i], f, pr.slope);
                        else prelu_bwd_elem(slope_acc, v[i], f, pr.slope);
                      } else if (!(f > 0.f)) {
                        v[i] = 0.f;
                      }
                    }
                  }
                }
              }
            }
            if (prof) { q1 = clock64(); e_math += q1 - q0; }
            if (epi & kEpiF32Out) {
              // raw accumulators: half h = 32 channels x 4 bytes = one 128-byte row of staging tile h
#pragma unroll
              for (int j = 0; j < 8; ++j)
                st_shared_v4(tile_addr(warp_t + h * kTileBytes, lane, j),
                             make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                        __float_as_uint(v[4 * j + 3])));
              continue;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              if ((epi & VSR_EPI_PRELU) && pr.tag) {
                const uint32_t m8 = posmask >> (8 * j);
                o.x = (o.x & 0xfffefffeu) | (m8 & 1u) | (((m8 >> 1) & 1u) << 16);
                o.y = (o.y & 0xfffefffeu) | ((m8 >> 2) & 1u) | (((m8 >> 3) & 1u) << 16);
                o.z = (o.z & 0xfffefffeu) | ((m8 >> 4) & 1u) | (((m8 >> 5) & 1u) << 16);
                o.w = (o.w & 0xfffefffeu) | ((m8 >> 6) & 1u) | (((m8 >> 7) & 1u) << 16);
              }
              st_shared_v4(tile_addr(out_t, lane, 4 * h + j), o);
            }
            if (epi & VSR_EPI_OUT2) {
              // out2 = v + res2 (v - res2 with OUT2_SUB), staged in place of the res2 tile (every thread owns its row)
              const float s2 = (epi & VSR_EPI_OUT2_SUB) ? -1.f : 1.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                const uint4 sq = ld_shared_v4(tile_addr(res2_t, lane, 4 * h + j));
                o.x = pack_bf16x2(fmaf(s2, bf16_lo(sq.x), v[8 * j + 0]), fmaf(s2, bf16_hi(sq.x), v[8 * j + 1]));
                o.y = pack_bf16x2(fmaf(s2, bf16_lo(sq.y), v[8 * j + 2]), fmaf(s2, bf16_hi(sq.y), v[8 * j + 3]));
                o.z = pack_bf16x2(fmaf(s2, bf16_lo(sq.z), v[8 * j + 4]), fmaf(s2, bf16_hi(sq.z), v[8 * j + 5]));
                o.w = pack_bf16x2(fmaf(s2, bf16_lo(sq.w), v[8 * j + 6]), fmaf(s2, bf16_hi(sq.w), v[8 * j + 7]));
                st_shared_v4(tile_addr(res2_t, lane, 4 * h + j), o);
              }
            }
            if (prof) e_st += clock64() - q1;
          }
          if (prof) q1 = clock64();
          if (has_in && !(epi & VSR_EPI_OUT2)) {
            // this operand set has been consumed: refill it with the look-ahead chunk's operands
            __syncwarp();
            if (la_ok) {
              if (lane == 0) issue_chunk(la, ib);
              la_ok = next_chunk(la);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (epi & kEpiF32Out) {
              tma_store_4d(&a.out_map, warp_t, grp.x + c, tc.x0 + sub_x, tc.y0 + m * a.bh + sub_y, tc.n);
              tma_store_4d(&a.out_map, warp_t + kTileBytes, grp.x + c + 32, tc.x0 + sub_x, tc.y0 + m * a.bh + sub_y, tc.n);
            } else if (!(dbg & 16)) {
              tma_store_4d(&a.out_map, out_t, grp.x + c, tc.x0 + sub_x, tc.y0 + m * a.bh + sub_y, tc.n);
              if (epi & VSR_EPI_OUT2)
                tma_store_4d(&a.out2_map, res2_t, grp.x + c, tc.x0 + sub_x, tc.y0 + m * a.bh + sub_y, tc.n);
            }
            bulk_commit();
            if (epi & VSR_EPI_OUT2) {
              // the res2 tile doubles as the out2 staging tile (one operand set): refill it only after the store has read it
              bulk_wait_read0();
              if (la_ok) issue_chunk(la, ib);
            }
          }
          if ((epi & VSR_EPI_OUT2) && has_in && la_ok) la_ok = next_chunk(la);
          ++q;
          if (prof) e_iss += clock64() - q1;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // (pair mode: the accumulator of BOTH CTAs is rewritten by the leader's next MMAs - free it on the leader's barrier)
        if (PAIR && !leader_cta) ptx::mbar_arrive_remote(ptx::mapa(tempty_bar + 8 * buf, 0));
        else ptx::mbar_arrive(tempty_bar + 8 * buf);
      }
    }
    // all output bytes are written before the CTA retires (waiting only for the TMA unit to have READ the staging tiles -
    // cp.async.bulk.wait_group.read - measured the same step time on one box, so the stronger wait stays)
    if (lane == 0) bulk_wait0();
    if (prof && blockIdx.x == 0 && lane == 0 && (ew & 3) == 0)
      printf("tc2-prof epilogue warp %d: total %lld cyc, wait(tfull) %lld, operands %lld, tmem-ld %lld, math %lld, pack+sts %lld, fence+issue %lld\n",
             warp, clock64() - e_t0, e_wait, e_in, e_ld, e_math, e_st, e_iss);
    if (epi & VSR_EPI_PRELU_BWD) {
      slope_acc = warp_sum(slope_acc + slope_acc1) * pr.inv;
      if (lane == 0) red[ew] = slope_acc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ew == 0 && lane == 0)
        a.slope_partials[blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    }
  }

  const uint64_t g_t2 = kAttrib ? ptx::globaltimer_ns() : 0;
  ptx::tc_fence_before();
  if (PAIR) ptx::cluster_sync();       // no CTA of the pair leaves while the other may still touch its shared / tensor memory
  else __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    if (PAIR) ptx::tmem_dealloc2(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
  if (prof && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    printf("tc2-prof block %d: start %llu ns, setup %llu ns, role done (producer) +%llu ns, exit +%llu ns\n", (int)blockIdx.x,
           (unsigned long long)g_t0, (unsigned long long)(g_t1 - g_t0), (unsigned long long)(g_t2 - g_t0),
           (unsigned long long)(ptx::globaltimer_ns() - g_t0));
}

// Host: fold the taps of a (single-group) table into columns = taps with equal (src, c0, dx), consecutive dy
// and evenly spaced slab indices.  Returns the longest column (0 if the table does not fit the argument array).
// `split` (bf16x3 tables): the group holds its taps three times over - [high plane x wh | low plane x wh | high plane x wl] -
// so the first and the last third have equal (src, c0, dx, dy); the third a tap belongs to joins the sort key, otherwise the
// two copies would interleave and no column could form.
int build_columns(const int32_t* taps, int n_taps, int idx_base, int max_cols, uint2* cols, int* n_cols, bool split) {
  struct T { int src, dy, dx, c0, idx, cls; };
  std::vector<T> v(n_taps);
  const int third = split && n_taps % 3 == 0 ? n_taps / 3 : n_taps;
  for (int i = 0; i < n_taps; ++i)
    v[i] = T{taps[4 * i], taps[4 * i + 1], taps[4 * i + 2], taps[4 * i + 3], idx_base + i, i / third};
  for (const T& t : v)
    if (t.src < 0 || t.src > 15 || t.dy < -8 || t.dy > 7 || t.dx < -8 || t.dx > 7 || (t.c0 & 7)) return 0;
  std::stable_sort(v.begin(), v.end(), [](const T& x, const T& y) {
    if (x.cls != y.cls) return x.cls < y.cls;
    if (x.src != y.src) return x.src < y.src;
    if (x.c0 != y.c0) return x.c0 < y.c0;
    if (x.dx != y.dx) return x.dx < y.dx;
    return x.dy < y.dy;
  });
  int n = 0, longest = 0;
  size_t i = 0;
  while (i < v.size()) {
    size_t j = i + 1;
    int stride = 0;
    while (j < v.size() && j - i < 3 && v[j].cls == v[i].cls && v[j].src == v[i].src && v[j].c0 == v[i].c0 &&
           v[j].dx == v[i].dx && v[j].dy == v[j - 1].dy + 1) {
      const int st = v[j].idx - v[j - 1].idx;
      if (j == i + 1) stride = st;
      if (st != stride || st <= 0 || st > 4095) break;
      ++j;
    }
    const int ndy = (int)(j - i);
    if (n >= max_cols) return 0;
    const int4 first = make_int4(v[i].src, v[i].dy, v[i].dx, v[i].c0);
    cols[n].x = (uint32_t)first.x | ((uint32_t)(first.y + 8) << 4) | ((uint32_t)(first.z + 8) << 8) | ((uint32_t)(first.w >> 3) << 12);
    cols[n].y = (uint32_t)v[i].idx | ((uint32_t)(ndy > 1 ? stride : 0) << 12) | ((uint32_t)ndy << 24);
    ++n;
    if (ndy > longest) longest = ndy;
    i = j;
  }
  *n_cols = n;
  return longest;
}

template <int EPI, bool PAIR>
cudaError_t prepare() {
  return cudaFuncSetAttribute(tapgemm_tc2_kernel<EPI, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
}

}  // namespace

int tapgemm_tc2_launch(const VsrTapGemmDesc* d, cudaStream_t stream) {
  VSR_CHECK_SUPPORTED(d->kc == kKc, "tapgemm(bf16): kc must be 64, got %d", d->kc);
  VSR_CHECK_SUPPORTED(d->nt >= 64 && d->nt <= 256 && d->nt % 64 == 0,
                      "tapgemm(bf16, v2): nt must be a multiple of 64 in [64,256], got %d", d->nt);
  VSR_CHECK_ARG(d->out.c % 8 == 0 && d->out.c >= 64, "tapgemm(bf16): out.c must be a multiple of 8 and >= 64");
  typedef void (*KernelFn)(const Tc2Args);
#define VSR_TC2_VARIANT(E) {E, tapgemm_tc2_kernel<E, false>, tapgemm_tc2_kernel<E, true>, prepare<E, false>, prepare<E, true>}
  static const struct { int epi; KernelFn fn, fn_pair; cudaError_t (*prep)(); cudaError_t (*prep_pair)(); } kVariants[] = {
      VSR_TC2_VARIANT(-1),
      VSR_TC2_VARIANT(0),
      VSR_TC2_VARIANT(VSR_EPI_BIAS),
      VSR_TC2_VARIANT(VSR_EPI_RES_PRE),
      VSR_TC2_VARIANT(VSR_EPI_BIAS | VSR_EPI_PRELU),
      VSR_TC2_VARIANT(VSR_EPI_BIAS | VSR_EPI_PRELU | VSR_EPI_OUT2),
      VSR_TC2_VARIANT(VSR_EPI_PRELU_BWD),
      VSR_TC2_VARIANT(VSR_EPI_PRELU_BWD | VSR_EPI_RES_PRE),
      VSR_TC2_VARIANT(kEpiF32Out),
  };
#undef VSR_TC2_VARIANT
  static bool attr_set = false;
  if (!attr_set) {
    for (const auto& v : kVariants) {
      cudaError_t e = v.prep();
      if (e == cudaSuccess) e = v.prep_pair();
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
        return VSR_ERR_CUDA;
      }
    }
    attr_set = true;
  }
  const bool split = d->dtype == VSR_BF16X2;
  if (split) {
    VSR_CHECK_SUPPORTED(d->epi == 0, "tapgemm(bf16x2): the tensor-core pass writes raw accumulators; run vsr_tap_epilogue for epi %d", d->epi);
    VSR_CHECK_SUPPORTED(d->n_srcs <= 8, "tapgemm(bf16x2): at most 8 sources (bit 3 of a tap's source index is the plane)");
  }
  const int epi_k = split ? kEpiF32Out : d->epi;
  const auto* variant = &kVariants[0];
  for (const auto& v : kVariants)
    if (v.epi == epi_k) variant = &v;

  Tc2Args a;
  memset(&a, 0, sizeof(a));
  int bw, bh;
  pick_box_pub(d->out.h, d->out.w, &bw, &bh);
  // split-bf16 mode: a source is two bf16 planes [2][n][h][w][c] = one map of 2n images
  auto planes = [split](VsrTensor4 t) { if (split) t.n *= 2; return t; };
  for (int s = 0; s < d->n_srcs; ++s) {
    VSR_CHECK_ARG(d->srcs[s].c % 8 == 0, "tapgemm(bf16): src channels must be a multiple of 8");
    int rc = get_src_map_pub(planes(d->srcs[s]), bw, bh, &a.maps[s]);
    if (rc != VSR_OK) return rc;
  }
  // epilogue tensors: one warp = 32 consecutive tile pixels = an ew x eh sub-box
  const int ew = bw < 32 ? bw : 32, eh = 32 / ew;
  {
    int rc = split ? get_f32_map_pub(d->out, ew, eh, &a.out_map) : get_src_map_pub(d->out, ew, eh, &a.out_map);
    if (rc != VSR_OK) return rc;
    VsrTensor4 t = d->out;
    if (d->epi & VSR_EPI_RES_PRE) {
      VSR_CHECK_ARG(d->residual != nullptr, "tapgemm: RES_PRE without residual");
      t.ptr = const_cast<void*>(d->residual);
      if ((rc = get_src_map_pub(t, ew, eh, &a.res_map)) != VSR_OK) return rc;
    }
    if (d->epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) {
      VSR_CHECK_ARG(d->aux_y != nullptr, "tapgemm: activation backward without aux_y");
      t.ptr = const_cast<void*>(d->aux_y);
      if ((rc = get_src_map_pub(t, ew, eh, &a.aux_map)) != VSR_OK) return rc;
    }
    if (d->epi & VSR_EPI_OUT2) {
      VSR_CHECK_ARG(d->out2 != nullptr && d->res2 != nullptr, "tapgemm: OUT2 without out2/res2");
      t.ptr = d->out2;
      if ((rc = get_src_map_pub(t, ew, eh, &a.out2_map)) != VSR_OK) return rc;
      t.ptr = const_cast<void*>(d->res2);
      if ((rc = get_src_map_pub(t, ew, eh, &a.res2_map)) != VSR_OK) return rc;
    }
  }
  a.tap_tab = reinterpret_cast<const int4*>(d->tap_tab);
  a.group_tab = reinterpret_cast<const int4*>(d->group_tab);
  a.w = static_cast<const uint8_t*>(d->w);
  a.bias = d->bias;
  a.slope = d->slope;
  a.slope_partials = d->slope_partials;
  a.out_scale = d->out_scale;
  a.epi = epi_k;
  a.nt = d->nt;
  a.n_groups = d->n_groups;
  a.N = d->out.n; a.H = d->out.h; a.W = d->out.w; a.Cout = d->out.c;
  a.bw = bw; a.bh = bh;
  a.bw_shift = 0;
  while ((1 << a.bw_shift) < bw) ++a.bw_shift;
  a.n_taps_total = d->n_taps_total;
  a.row_bytes = bw * 128;
  const int n_in = ((d->epi & VSR_EPI_RES_PRE) ? 1 : 0) + ((d->epi & (VSR_EPI_PRELU_BWD | VSR_EPI_RELU_BWD)) ? 1 : 0) +
                   ((d->epi & VSR_EPI_OUT2) ? 1 : 0);
  const Tunables& tn = tunables();
  // ---- CTA pairs (cta_group::2): every CTA holds half of each weight slab.  Worth it where the B operand weighs: wide
  // slabs (nt >= 128: their shared-memory reads and their resident footprint halve) and long tables of narrow ones (the
  // strided k x k convolutions: slab traffic L2 -> SM halves); the short nt = 64 tables of the 1x1 convolutions are
  // HBM-bound and keep single CTAs.  VSR_TC_PAIR=0 / 1 forces none / all.
  {
    const long px_tiles = (long)d->out.n * ((d->out.w + bw - 1) / bw) * ((d->out.h + bh - 1) / bh);
    bool want = d->nt >= 128 || (d->max_group_taps >= 16 && d->n_groups == 1);
    if (tn.tc_pair == 0) want = false;
    if (tn.tc_pair == 1) want = true;
    a.pair = want && d->nt % 32 == 0 && px_tiles >= 2 && num_sms() >= 2;
  }
  const int b_bytes = d->nt * 128 / (a.pair ? 2 : 1);      // bytes of a weight slab held by one CTA
  // ---- epilogue staging: in the chunk-parity mode (nt > 64: a warp processes chunk after chunk of the same tile) a
  // second output tile lets the TMA store of a chunk overlap the next chunk, and a second operand set lets the
  // residual / saved-activation tiles arrive two chunks ahead; taken when >= 3 pipeline stages remain
  // (nt <= 64: the warp sets alternate tiles, which already hides both)
  int cand[4][2], n_cand = 0;
  if (split) {
    cand[n_cand][0] = 2; cand[n_cand++][1] = 1;       // two fp32 half-chunk tiles per warp
  } else if (d->nt > 64 && !(d->epi & VSR_EPI_OUT2) && tn.tc_epibuf != 1) {      // (VSR_TC_EPIBUF=1: single staging, for A/B runs)
    if (n_in > 0) {
      cand[n_cand][0] = 2; cand[n_cand++][1] = 2;
      cand[n_cand][0] = 1; cand[n_cand++][1] = 2;
    }
    cand[n_cand][0] = 2; cand[n_cand++][1] = 1;
  }
  if (!split) { cand[n_cand][0] = 1; cand[n_cand++][1] = 1; }
  int ndy_max = 1;
  // Pass 0 evaluates the smallest epilogue staging alone, to learn whether the shared-load mode engages at all; pass 1 picks
  // the candidate.  For long tables (>= 64 taps in a group) a candidate whose staging squeezes the shared-load mode out is
  // skipped: double-buffered operands hide L2 latency in the epilogue, but plain columns double the L2 -> SM operand
  // stream of every tap (the PReLU' data gradient of the 128-channel strided convolution ran 121 us against its forward
  // twin's 60 us, profiles/r02b_sweep_config5.json).
  bool tall_min = false;
  for (int pass = 0; pass < 2; ++pass)
  for (int ci = pass == 0 ? n_cand - 1 : 0; ci < n_cand; ++ci) {
    a.epi_obuf = cand[ci][0];
    a.epi_ibuf = cand[ci][1];
    a.epi_bytes = kEpiWarps * (a.epi_obuf + a.epi_ibuf * n_in) * kTileBytes;
    a.tall = 0;
    a.n_cols = 0;
    ndy_max = 1;
    const long avail = kSmemBudget - kCtrlBytes - a.epi_bytes;
    a.debug = tn.tc_debug > 0 ? tn.tc_debug : 0;

    // ---- shared-load mode: taps that differ only by a row shift read one A box; with nt <= 128 two pixel
    // tiles stacked in y also share every weight slab.  Needs the host copies of the tables
    // (d->tap_tab_host, and d->group_tab_host when there are several groups).
    a.mb = 1;
    a.a_bytes = kATileBytes;
    const long res_need = (long)d->max_group_taps * b_bytes;
    {
      const bool want = tn.tc_tall != 0;
      bool ok = want && d->tap_tab_host != nullptr && bw >= 8 && bw * bh == kBlockM && d->n_taps_total <= 4095 &&
                d->n_groups <= kMaxTallGroups && (d->n_groups == 1 || d->group_tab_host != nullptr);
      int n_cols = 0, longest = 0;
      for (int gi = 0; ok && gi < d->n_groups; ++gi) {
        const int o0 = d->n_groups == 1 ? 0 : d->group_tab_host[4 * gi];
        const int begin = d->n_groups == 1 ? 0 : d->group_tab_host[4 * gi + 1];
        const int count = d->n_groups == 1 ? d->n_taps_total : d->group_tab_host[4 * gi + 2];
        int nc = 0;
        // (bf16 tables keep the 64-column limit they were tuned with; the tripled tables of the bf16x3 mode may use all 96)
        const int lg = build_columns(d->tap_tab_host + 4 * begin, count, begin, (split ? kMaxParamCols : 64) - n_cols,
                                     a.cols + n_cols, &nc, split);
        if (lg == 0 || begin > 0xffff || count > 0x7fff) { ok = false; break; }
        a.tgroups[gi] = make_int4(o0, n_cols, nc, begin | (count << 16));
        n_cols += nc;
        if (lg > longest) longest = lg;
      }
      if (ok && d->n_groups == 1) {
        // the output slice of a single group comes from the device table (group_tab_host is optional)
        a.tgroups[0].x = -1;
      }
      if (ok && longest >= 2) {
        const int mb = (d->nt <= 128 && a.H >= 2 * bh) ? 2 : 1;
        const int rows = mb * bh + longest - 1;
        const long abox = (long)rows * a.row_bytes;
        const bool res_ok = d->max_group_taps > 0 && d->nt > 64 && res_need <= avail - 3 * abox;
        // streamed slabs: three stages of {shared box + its slabs} when they fit; two are enough when a stage is long
        // (3 taps x mb sub-tiles x 4 MMAs) and the alternative is one load per tap (3x the L2 -> SM traffic: the
        // 3x3(x3) convolutions of the Conv3d path, profiles/README.md).  VSR_TC_TALL_STAGES=3 restores the old rule.
        const int min_stages = tn.tc_tall_stages == 3 ? 3 : 2;
        const bool stream_ok = min_stages * (abox + (long)longest * b_bytes) <= avail;
        if (rows <= 256 && (res_ok || stream_ok)) {
          a.tall = 1;
          a.mb = mb;
          a.n_cols = n_cols;
          a.a_bytes = (int)abox;
          ndy_max = longest;
          for (int s = 0; s < d->n_srcs; ++s) {
            int rc = get_src_map_pub(planes(d->srcs[s]), bw, rows, &a.tall_maps[s]);
            if (rc != VSR_OK) return rc;
          }
        }
      }
    }
    a.tiles_x = (a.W + bw - 1) / bw;
    a.tiles_y = (a.H + bh * a.mb - 1) / (bh * a.mb);
    const long tiles = (long)a.n_groups * a.N * a.tiles_x * a.tiles_y;
    VSR_CHECK_SUPPORTED(tiles < (1l << 30), "tapgemm(bf16): too many tiles");
    a.num_tiles = (int)tiles;
    a.m_tiles = a.N * a.tiles_x * a.tiles_y;
    // weight-resident mode: the group's slabs stay in smem and >= 3 A stages remain; worth it when the
    // slabs are large next to the A tile (nt > 64) and every CTA sees few groups
    // (short nt=64 tables - the 1x1 convolutions on concatenations - were measured 5-10% slower resident:
    // tools/hr_sweep.py, so they stream their 8 KB slabs with the A tiles)
    a.resident = d->max_group_taps > 0 && d->nt > 64 && res_need <= avail - 3 * (long)a.a_bytes &&
                 tiles >= 2 * (long)num_sms();
    // (resident slabs for the short nt = 64 tables of the 1x1 convolutions were measured neutral to slightly slower,
    //  with either tile walk: tools/hr_sweep.py; VSR_TC_RESIDENT=1 forces them)
    {
      if (tn.tc_resident == 0) a.resident = 0;
      if (tn.tc_resident == 1 && d->max_group_taps > 0 && res_need <= avail - 2 * (long)a.a_bytes)
        a.resident = 1;
    }
    if (a.tall && !a.resident && 2 * ((long)a.a_bytes + (long)ndy_max * b_bytes) > avail) {
      // the shared box plus its slabs does not fit twice without resident weights: plain columns
      a.tall = 0;
      a.mb = 1;
      a.a_bytes = kATileBytes;
      ndy_max = 1;
      a.tiles_y = (a.H + bh - 1) / bh;
      a.num_tiles = (int)((long)a.n_groups * a.N * a.tiles_x * a.tiles_y);
      a.m_tiles = a.N * a.tiles_x * a.tiles_y;
    }
    if (a.pair) {
      // the walk counts PAIRS of pixel tiles (decode_tile: CTA `rank` takes pixel tile 2 * pair + rank)
      a.m_tiles = (a.m_tiles + 1) / 2;
      a.num_tiles = a.n_groups * a.m_tiles;
    }
    {
      auto magic = [](int d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d); };
      // exact for every n < 2^32: ceil(2^32/d) overestimates 2^32/d by < 1, so the quotient by < n / 2^32 < 1
      a.mg_groups = magic(a.n_groups);
      a.mg_mtiles = magic(a.m_tiles);
      a.mg_tx = magic(a.tiles_x);
      a.mg_ty = magic(a.tiles_y);
    }
    a.contig = a.resident && d->n_groups > 1;
    a.res_bytes = a.resident ? (int)res_need : 0;
    a.stage_bytes = a.resident ? a.a_bytes : a.a_bytes + ndy_max * b_bytes;
    int stages = (int)((avail - a.res_bytes) / a.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (tn.tc_stages >= 1 && tn.tc_stages < stages) stages = tn.tc_stages;
    if (stages < 1 && ci + 1 < n_cand) continue;
    VSR_CHECK_SUPPORTED(stages >= 1, "tapgemm(bf16, v2): no room for a pipeline stage");
    a.stages = stages;
    if (pass == 0) {
      tall_min = a.tall != 0;
      break;
    }
    if (tall_min && !a.tall && d->max_group_taps >= 64 && ci + 1 < n_cand) continue;
    if (stages >= 3 || ci + 1 == n_cand) break;
  }
  const int smem = kCtrlBytes + a.epi_bytes + a.res_bytes + a.stages * a.stage_bytes;
  int grid = num_sms();
  if (tn.tc_grid >= 1) grid = tn.tc_grid;
  if (grid > kPartialsLen) grid = kPartialsLen;
  KernelFn kernel = a.pair ? variant->fn_pair : variant->fn;
  if (a.pair) {
    // one CTA pair per TPC - but not every TPC can host a cluster (GPCs with an odd number of live SMs): a persistent
    // kernel must not launch more pairs than are co-resident, or the rest would run as a second wave
    static int max_pairs = -1;
    if (max_pairs < 0) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(num_sms());
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = kSmemBudget;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &q) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = num_sms() / 2;
      }
      max_pairs = n;
    }
    int walkers = grid / 2;
    if (walkers > max_pairs) walkers = max_pairs;
    if (walkers > a.num_tiles) walkers = a.num_tiles;
    grid = 2 * walkers;
  } else if (grid > a.num_tiles) {
    grid = a.num_tiles;
  }
  if (a.debug & 64) {
    fprintf(stderr, "tc2 plan: epi bufs %d/%d pair %d nt %d groups %d taps %d | tall %d mb %d cols %d a_bytes %d | resident %d res_bytes %d | stages %d x %d, epi %d, smem %d, tiles %d (m %d), grid %d\n",
            a.epi_obuf, a.epi_ibuf, a.pair, a.nt, a.n_groups, a.n_taps_total, a.tall, a.mb, a.n_cols, a.a_bytes, a.resident, a.res_bytes, a.stages,
            a.stage_bytes, a.epi_bytes, smem, a.num_tiles, a.m_tiles, grid);
    if (a.tall)
      for (int gi = 0; gi < a.n_groups; ++gi)
        fprintf(stderr, "  group %d: o0 %d cols [%d, +%d) taps [%d, +%d)\n", gi, a.tgroups[gi].x, a.tgroups[gi].y,
                a.tgroups[gi].z, a.tgroups[gi].w & 0xffff, a.tgroups[gi].w >> 16);
  }
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (tn.pdl != 0) {
      attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
      ++n_attr;
    }
    if (a.pair) {
      attr[n_attr].id = cudaLaunchAttributeClusterDimension;
      attr[n_attr].val.clusterDim.x = 2;
      attr[n_attr].val.clusterDim.y = 1;
      attr[n_attr].val.clusterDim.z = 1;
      ++n_attr;
    }
    if (n_attr) {
      cfg.attrs = attr;
      cfg.numAttrs = n_attr;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a);
    if (e != cudaSuccess) {
      set_error("tapgemm_tc2: launch failed: %s", cudaGetErrorString(e));
      return VSR_ERR_CUDA;
    }
  }
  VSR_CHECK_LAUNCH("tapgemm_tc2");
  return VSR_OK;
}

}  // namespace vsr
