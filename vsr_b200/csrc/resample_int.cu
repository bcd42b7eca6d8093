// resample_int.cu — bilinear / trilinear up-sampling by an integer power-of-two factor with align_corners = False, forward
// and backward: the case of the nets (F.interpolate(scale_factor = r, mode = 'bilinear', align_corners = False) is the global
// skip of srfb_net.py:47 / drf_sisr_net.py; trilinear is its 3-D sibling).
//
// With an integer ratio R the interpolation weights are periodic: output R*i + p reads input i with weight wo(p) and the
// neighbour i - 1 (p < R/2) or i + 1 (p >= R/2) with weight 1 - wo(p), where wo(p) = 1 - |(p + 0.5)/R - 0.5|; the clamp at
// the borders folds the neighbour back onto i.  For R in {2, 4, 8} those weights are dyadic, so they equal PyTorch's
// `scale * (o + 0.5) - 0.5` arithmetic exactly.  No coordinate arithmetic is left per element:
//   * a thread owns one input column: forward it forms the R horizontally interpolated values of a row from its value and
//     the two neighbours' (warp shuffles), walks down its strip of input rows keeping three such rows in registers and
//     writes R output rows of R values (one 16-byte store each for R = 4) per input row;
//   * backward it loads the R output gradients above its column as one vector, sends the parts that belong to the
//     neighbour columns through shuffles, and walks the output rows of its strip carrying the parts that belong to the
//     rows above / below in registers.
//   * a warp spans 32 columns of which the outer two are halo (tile stride 30), so no lane needs a second load; a strip
//     re-reads one input row (forward) or R output rows (backward) per 8 - 16 input rows.
//   * 3-D: the two input slices of an output slice are blended at load time (forward); backward a thread sums the up to 2R
//     output slices that touch its input slice while loading (adjacent input slices share them through the L2).
// Bytes moved = input + output; everything else is registers.
#include "common.cuh"

namespace vsr {
namespace {

template <int R>
struct Phase {
  // weight of the own input sample for phase p (the other sample is the previous one for p < R/2, else the next one)
  __host__ __device__ static constexpr float own(int p) {
    const float f = (p + 0.5f) / R - 0.5f;
    return f < 0.f ? 1.f + f : 1.f - f;
  }
};

constexpr int kLanesOut = 30;                  // useful columns per warp (lanes 1..30); lanes 0 and 31 are halo

template <int R>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[R]) {
  if constexpr (R == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int i = 0; i < R; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}
template <int R>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[R]) {
  if constexpr (R == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < R; i += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
// grid (ceil(w / 30), ceil(h / (8 * S)), nc * od), block (32, 8).  Z: trilinear (od = R * d), else d == od == 1.
template <int R, bool Z>
__global__ void __launch_bounds__(256) up_int_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int d, int h,
                                                        int w, int S) {
  const int lane = threadIdx.x;
  const int ix = blockIdx.x * kLanesOut + lane - 1;
  const int ixc = min(max(ix, 0), w - 1);
  const int ya = (blockIdx.y * 8 + threadIdx.y) * S;
  if (ya >= h) return;                                              // whole warp
  const int yb = min(ya + S, h);
  const int ow = R * w, oh = R * h;
  const float* p0;
  const float* p1 = nullptr;
  float wz0 = 1.f, wz1 = 0.f;
  if constexpr (Z) {
    const int od = R * d;
    const int c = blockIdx.z / od, oz = blockIdx.z - c * od;
    const int iz = oz / R, s = oz - iz * R;
    const float f = (s + 0.5f) / R - 0.5f;
    const int izs = min(max(iz + (f < 0.f ? -1 : 1), 0), d - 1);
    wz0 = f < 0.f ? 1.f + f : 1.f - f;
    wz1 = 1.f - wz0;
    p0 = x + ((size_t)c * d + iz) * h * w;
    p1 = x + ((size_t)c * d + izs) * h * w;
  } else {
    p0 = x + (size_t)blockIdx.z * h * w;
  }
  float* yp = y + (size_t)blockIdx.z * oh * ow + (size_t)R * ixc;
  const bool writer = lane >= 1 && lane <= kLanesOut && ix < w;

  auto load = [&](int iy) -> float {
    const size_t o = (size_t)min(max(iy, 0), h - 1) * w + ixc;
    float v = __ldg(p0 + o);
    if constexpr (Z) v = wz0 * v + wz1 * __ldg(p1 + o);
    return v;
  };
  auto hrow = [&](float v, float (&hv)[R]) {
    const float l = __shfl_up_sync(0xffffffffu, v, 1), r = __shfl_down_sync(0xffffffffu, v, 1);
#pragma unroll
    for (int p = 0; p < R; ++p) {
      const float wo = Phase<R>::own(p);
      hv[p] = wo * v + (1.f - wo) * (2 * p < R ? l : r);
    }
  };

  float hp[R], hc[R], hn[R];
  hrow(load(ya - 1), hp);
  hrow(load(ya), hc);
  float vn = load(ya + 1);
  for (int iy = ya; iy < yb; ++iy) {
    const float vnn = load(iy + 2);                                 // in flight while this row's outputs are formed
    hrow(vn, hn);
    if (writer) {
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const float wo = Phase<R>::own(q);
        float o[R];
#pragma unroll
        for (int p = 0; p < R; ++p) o[p] = wo * hc[p] + (1.f - wo) * (2 * q < R ? hp[p] : hn[p]);
        store_vec<R>(yp + (size_t)(R * iy + q) * ow, o);
      }
    }
#pragma unroll
    for (int p = 0; p < R; ++p) { hp[p] = hc[p]; hc[p] = hn[p]; }
    vn = vnn;
  }
}

// ---- backward --------------------------------------------------------------------------------------------------------
// grid (ceil(w / 30), ceil(h / (8 * S)), nc * d), block (32, 8)
template <int R, bool Z>
__global__ void __launch_bounds__(256) up_int_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int d, int h,
                                                        int w, int S) {
  const int lane = threadIdx.x;
  const int ix = blockIdx.x * kLanesOut + lane - 1;
  const bool inside = ix >= 0 && ix < w;
  const int ya = (blockIdx.y * 8 + threadIdx.y) * S;
  if (ya >= h) return;
  const int yb = min(ya + S, h);
  const int ow = R * w, oh = R * h;
  // output slices that touch this input slice, with their weights
  constexpr int NZ = Z ? 2 * R : 1;
  const float* sl[NZ];
  float wz[NZ];
  if constexpr (Z) {
    const int c = blockIdx.z / d, iz = blockIdx.z - c * d;
    const int od = R * d;
#pragma unroll
    for (int k = 0; k < NZ; ++k) {
      const int oz = R * iz - R / 2 + k;
      const bool ok = oz >= 0 && oz < od;
      const int j = ok ? oz / R : 0, s = ok ? oz - j * R : 0;
      const float f = (s + 0.5f) / R - 0.5f;
      const float wo = f < 0.f ? 1.f + f : 1.f - f;
      float wk;
      if (j == iz) wk = wo + (((f < 0.f && iz == 0) || (f >= 0.f && iz == d - 1)) ? 1.f - wo : 0.f);
      else wk = 1.f - wo;
      wz[k] = ok ? wk : 0.f;
      sl[k] = dy + ((size_t)c * od + (ok ? oz : 0)) * oh * ow;
    }
  } else {
    sl[0] = dy + (size_t)blockIdx.z * oh * ow;
    wz[0] = 1.f;
  }
  const size_t col = (size_t)R * min(max(ix, 0), w - 1);

  // gradient of the horizontally interpolated row oy with respect to input column ix
  auto tcol = [&](int oy) -> float {
    float v[R];
    if constexpr (Z) {
#pragma unroll
      for (int p = 0; p < R; ++p) v[p] = 0.f;
#pragma unroll
      for (int k = 0; k < NZ; ++k) {
        if (wz[k] != 0.f) {                                         // uniform over the block
          float t[R];
          load_vec<R>(sl[k] + (size_t)oy * ow + col, t);
#pragma unroll
          for (int p = 0; p < R; ++p) v[p] = fmaf(wz[k], t[p], v[p]);
        }
      }
    } else {
      load_vec<R>(sl[0] + (size_t)oy * ow + col, v);
    }
    float own = 0.f, to_l = 0.f, to_r = 0.f;
#pragma unroll
    for (int p = 0; p < R; ++p) {
      const float wo = Phase<R>::own(p);
      const float g = inside ? v[p] : 0.f;
      own = fmaf(wo, g, own);
      if (2 * p < R) to_l = fmaf(1.f - wo, g, to_l); else to_r = fmaf(1.f - wo, g, to_r);
    }
    const float from_r = __shfl_down_sync(0xffffffffu, to_l, 1), from_l = __shfl_up_sync(0xffffffffu, to_r, 1);
    float t = own + from_l + from_r;
    if (ix == 0) t += to_l;                                         // clamp at the borders folds back
    if (ix == w - 1) t += to_r;
    return t;
  };
  // row j of the input: own / up (to row j - 1) / down (to row j + 1) parts of its R output rows
  auto rows = [&](int j, bool want_first_half, bool want_second_half, float& own, float& up, float& down) {
    own = up = down = 0.f;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const bool first = 2 * q < R;
      if ((first && !want_first_half) || (!first && !want_second_half)) continue;
      const float wo = Phase<R>::own(q);
      const float t = tcol(R * j + q);
      own = fmaf(wo, t, own);
      if (first) up = fmaf(1.f - wo, t, up); else down = fmaf(1.f - wo, t, down);
    }
  };

  const bool writer = lane >= 1 && lane <= kLanesOut && inside;
  float* dxp = dx + (size_t)blockIdx.z * h * w + min(max(ix, 0), w - 1);
  float own, up, down, carry = 0.f;
  if (ya > 0) {
    rows(ya - 1, false, true, own, up, down);
    carry = down;
  }
  float pend = 0.f;
  for (int j = ya; j < yb; ++j) {
    rows(j, true, true, own, up, down);
    if (j == 0) own += up;                                          // fold at the top border
    else if (j > ya && writer) dxp[(size_t)(j - 1) * w] = pend + up;
    // (j == ya > 0: `up` belongs to the strip above, which computes it itself)
    pend = own + carry;
    carry = down;
  }
  float last_up = 0.f;
  if (yb < h) {
    rows(yb, true, false, own, up, down);
    last_up = up;
  } else {
    last_up = carry;                                                // fold at the bottom border
  }
  if (writer) dxp[(size_t)(yb - 1) * w] = pend + last_up;
}

int pick_strip(int h, long planes, long xtiles) {
  // rows per warp strip: long strips amortise the halo rows, short ones give the machine enough warps (32 per SM)
  int S = 16;
  while (S > 2 && planes * xtiles * ((h + S - 1) / S) < 32l * num_sms()) S >>= 1;
  return S;
}

template <int R>
void launch_fwd(const float* x, float* y, int nc, int d, int h, int w, bool z, cudaStream_t s) {
  const int od = z ? R * d : 1;
  const long xt = (w + kLanesOut - 1) / kLanesOut;
  const int S = pick_strip(h, (long)nc * od, xt);
  dim3 grid((unsigned)xt, (h + 8 * S - 1) / (8 * S), nc * od), block(32, 8);
  if (z) up_int_fwd_kernel<R, true><<<grid, block, 0, s>>>(x, y, d, h, w, S);
  else up_int_fwd_kernel<R, false><<<grid, block, 0, s>>>(x, y, d, h, w, S);
}
template <int R>
void launch_bwd(const float* dy, float* dx, int nc, int d, int h, int w, bool z, cudaStream_t s) {
  const long xt = (w + kLanesOut - 1) / kLanesOut;
  const int S = pick_strip(h, (long)nc * d, xt);
  dim3 grid((unsigned)xt, (h + 8 * S - 1) / (8 * S), nc * d), block(32, 8);
  if (z) up_int_bwd_kernel<R, true><<<grid, block, 0, s>>>(dy, dx, d, h, w, S);
  else up_int_bwd_kernel<R, false><<<grid, block, 0, s>>>(dy, dx, d, h, w, S);
}

}  // namespace

// 0 if the shape is not an integer power-of-two up-scaling this file covers, else R
int upsample_int_ratio(const void* big, int nc, int d, int h, int w, int od, int oh, int ow, int align_corners) {
  if (align_corners || h <= 0 || w <= 0 || oh % h || ow % w) return 0;
  const int r = oh / h;
  if (ow / w != r || !(r == 2 || r == 4 || r == 8)) return 0;
  if (!((d == 1 && od == 1) || od == r * d)) return 0;
  if ((reinterpret_cast<uintptr_t>(big) & 15) != 0) return 0;
  const long planes = (long)nc * (od == 1 ? 1 : od);
  if (planes > 65535 || h > 8 * 65535) return 0;
  return r;
}

bool upsample_int_fwd(const float* x, float* y, int nc, int d, int h, int w, int od, int oh, int ow, int ac, cudaStream_t s) {
  const int r = upsample_int_ratio(y, nc, d, h, w, od, oh, ow, ac);
  if (!r) return false;
  const bool z = !(d == 1 && od == 1);
  if (r == 2) launch_fwd<2>(x, y, nc, d, h, w, z, s);
  else if (r == 4) launch_fwd<4>(x, y, nc, d, h, w, z, s);
  else launch_fwd<8>(x, y, nc, d, h, w, z, s);
  return true;
}

bool upsample_int_bwd(const float* dy, float* dx, int nc, int d, int h, int w, int od, int oh, int ow, int ac, cudaStream_t s) {
  const int r = upsample_int_ratio(dy, nc, d, h, w, od, oh, ow, ac);
  if (!r) return false;
  const bool z = !(d == 1 && od == 1);
  if (r == 2) launch_bwd<2>(dy, dx, nc, d, h, w, z, s);
  else if (r == 4) launch_bwd<4>(dy, dx, nc, d, h, w, z, s);
  else launch_bwd<8>(dy, dx, nc, d, h, w, z, s);
  return true;
}

}  // namespace vsr
