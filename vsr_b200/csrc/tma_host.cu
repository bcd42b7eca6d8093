// tma_host.cu — host-side helpers shared by the tcgen05 kernels: the TMA tensor-map cache, the pixel-box choice and
// the tuning overrides read from the environment ONCE (vsr_reload_tunables() re-reads them: tests and tools flip them
// between launches; the launch path itself never calls getenv).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace vsr {

namespace {
constexpr int kBlockM = 128;
constexpr int kKc = 64;

Tunables g_tunables;
std::once_flag g_tunables_once;

int env_int(const char* name) {
  const char* v = getenv(name);
  return (v && v[0]) ? atoi(v) : -1;
}
void load_tunables() {
  Tunables t;
  t.tc_debug = env_int("VSR_TC_DEBUG");
  t.tc_tall = env_int("VSR_TC_TALL");
  t.tc_tall_stages = env_int("VSR_TC_TALL_STAGES");
  t.tc_resident = env_int("VSR_TC_RESIDENT");
  t.tc_stages = env_int("VSR_TC_STAGES");
  t.tc_grid = env_int("VSR_TC_GRID");
  t.tc_square = env_int("VSR_TC_SQUARE");
  t.tc_pair = env_int("VSR_TC_PAIR");
  t.tc_epibuf = env_int("VSR_TC_EPIBUF");
  t.pdl = env_int("VSR_PDL");
  t.wg_debug = env_int("VSR_WG_DEBUG");
  t.wg_tall = env_int("VSR_WG_TALL");
  t.fc_simt = env_int("VSR_FC_SIMT");           // 1: CUDA-core first convolution in bf16 mode too (A/B)
  t.up_generic = env_int("VSR_UP_GENERIC");     // 1: skip the integer-ratio up-sampling kernels (A/B)
  g_tunables = t;
}
}  // namespace

const Tunables& tunables() {
  std::call_once(g_tunables_once, load_tunables);
  return g_tunables;
}

namespace {

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
  int f32;      // 1: fp32 map with a box of 32 channels (raw-accumulator output of the split-bf16 mode)
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && n == o.n && h == o.h && w == o.w && c == o.c && bw == o.bw && bh == o.bh && f32 == o.f32;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.n); mix(k.h); mix(k.w); mix(k.c); mix(k.bw); mix(k.bh); mix(k.f32);
    return h;
  }
};

std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

// 4-D bf16 map over a dense [n][h][w][c] map; box = 64 channels x bw x bh x 1, 128B swizzle
// (f32: fp32 elements, box = 32 channels - the same 128-byte rows).
int get_src_map(const VsrTensor4& t, int bw, int bh, CUtensorMap* out, int f32 = 0) {
  MapKey key{t.ptr, t.n, t.h, t.w, t.c, bw, bh, f32};
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
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t strides[3] = {(cuuint64_t)t.c * es, (cuuint64_t)t.w * t.c * es,
                           (cuuint64_t)t.h * t.w * t.c * es};
  cuuint32_t box[4] = {(cuuint32_t)(f32 ? kKc / 2 : kKc), (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, dims, strides, box, estr,
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
    const int sq = tunables().tc_square;
    const bool square = sq >= 0 ? sq == 1 : w > 32;
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
int get_f32_map_pub(const VsrTensor4& t, int bw, int bh, CUtensorMap* out) { return get_src_map(t, bw, bh, out, 1); }
void pick_box_pub(int h, int w, int* bw, int* bh) { pick_box(h, w, bw, bh); }


}  // namespace vsr

extern "C" void vsr_reload_tunables() {
  (void)vsr::tunables();
  vsr::load_tunables();
}
