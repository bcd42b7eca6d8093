// resample.cu — standalone bandwidth kernels: nn.PixelShuffle and its inverse on NCHW
// (drf_net.py:142; the nets themselves never launch it — the shuffle is a reinterpretation of
// the phase-blocked layout) and F.interpolate(bilinear / trilinear) forward + backward
// (srfb_net.py:47; trilinear has no reference call site).
#include "common.cuh"

namespace vsr {
namespace {

// y[n][c][h*r+i][w*r+j] = x[n][c*r*r + i*r + j][h][w].  grid = (x tiles, output rows, n*c planes):
// no per-element divisions; a thread writes 4 consecutive outputs (16-byte store) when W % 4 == 0.
constexpr int kRowsPerBlock = 32;

__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const float* __restrict__ x, float* __restrict__ y, int c,
                                                           int h, int w, int r, int inverse) {
  const int W = w * r, H = h * r;
  const int plane = blockIdx.z;                  // n * c + ci
  const size_t hw = (size_t)h * w;
  const int Y0 = blockIdx.y * kRowsPerBlock;
  const int rows = min(kRowsPerBlock, H - Y0);
  const bool vec = (r == 2) && ((W & 3) == 0);
  const int per_row = vec ? (W >> 2) : W;
  const float* xplane = x + (size_t)plane * r * r * hw;      // un-shuffled planes of this (n, c)
  float* yplane = y + (size_t)plane * (size_t)H * W;          // shuffled plane
  const float* splane = x + (size_t)plane * (size_t)H * W;    // (inverse) shuffled source
  float* dplane = y + (size_t)plane * r * r * hw;             // (inverse) un-shuffled destination
  for (int idx = threadIdx.x; idx < rows * per_row; idx += blockDim.x) {
    const int rr = idx / per_row, xi = idx - rr * per_row;
    const int Y = Y0 + rr, yy = Y / r, i = Y - yy * r;
    if (vec) {
      const int X4 = xi * 4;
      const size_t a_off = ((size_t)i * 2) * hw + (size_t)yy * w + X4 / 2;
      if (!inverse) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(xplane + a_off));
        const float2 b = __ldg(reinterpret_cast<const float2*>(xplane + a_off + hw));
        *reinterpret_cast<float4*>(yplane + (size_t)Y * W + X4) = make_float4(a.x, b.x, a.y, b.y);
      } else {
        const float4 v = __ldg(reinterpret_cast<const float4*>(splane + (size_t)Y * W + X4));
        *reinterpret_cast<float2*>(dplane + a_off) = make_float2(v.x, v.z);
        *reinterpret_cast<float2*>(dplane + a_off + hw) = make_float2(v.y, v.w);
      }
    } else {
      const int X = xi, xx = X / r, j = X - xx * r;
      const size_t u_off = ((size_t)i * r + j) * hw + (size_t)yy * w + xx;
      if (!inverse) yplane[(size_t)Y * W + X] = __ldg(xplane + u_off);
      else dplane[u_off] = __ldg(splane + (size_t)Y * W + X);
    }
  }
}

struct LinCoord {
  int i0, i1;
  float w0, w1;
};
// torch's area_pixel_compute_source_index for linear modes
__device__ __forceinline__ LinCoord lin_coord(int o, int in_size, int out_size, int align_corners) {
  LinCoord c;
  float src;
  if (align_corners) {
    const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    src = scale * o;
  } else {
    const float scale = (float)in_size / (float)out_size;
    src = scale * (o + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  c.i0 = (int)src;
  if (c.i0 > in_size - 1) c.i0 = in_size - 1;
  c.i1 = c.i0 + (c.i0 < in_size - 1 ? 1 : 0);
  c.w1 = src - (float)c.i0;
  c.w0 = 1.f - c.w1;
  return c;
}

// grid = (x tiles of 1024, row tiles, nc*od).  The x interpolation coordinates of the tile are
// computed once per block into shared memory and reused for all rows of the tile; a thread
// produces 4 consecutive outputs per row (16-byte store when aligned).
constexpr int kUpTile = 1024;
__global__ void __launch_bounds__(256) upsample_linear_kernel(const float* __restrict__ x, float* __restrict__ y, int d,
                                                             int h, int w, int od, int oh, int ow, int ac) {
  __shared__ int xi0[kUpTile], xi1[kUpTile];
  __shared__ float xw1[kUpTile];
  const int x0 = blockIdx.x * kUpTile;
  const int nx = min(kUpTile, ow - x0);
  for (int i = threadIdx.x; i < nx; i += blockDim.x) {
    const LinCoord cx = lin_coord(x0 + i, w, ow, ac);
    xi0[i] = cx.i0; xi1[i] = cx.i1; xw1[i] = cx.w1;
  }
  __syncthreads();
  const int cz = blockIdx.z;                     // c * od + oz
  const int c = cz / od, oz = cz - c * od;
  const bool three_d = !(d == 1 && od == 1);
  LinCoord czc;
  czc.i0 = czc.i1 = 0; czc.w0 = 1.f; czc.w1 = 0.f;
  if (three_d) czc = lin_coord(oz, d, od, ac);
  const float* p0 = x + ((size_t)c * d + czc.i0) * h * w;
  const float* p1 = x + ((size_t)c * d + czc.i1) * h * w;
  const bool vec = ((ow & 3) == 0);
  for (int oy = blockIdx.y * kRowsPerBlock; oy < min((int)(blockIdx.y + 1) * kRowsPerBlock, oh); ++oy) {
    const LinCoord cy = lin_coord(oy, h, oh, ac);
    const float* r00 = p0 + (size_t)cy.i0 * w;
    const float* r01 = p0 + (size_t)cy.i1 * w;
    const float* r10 = p1 + (size_t)cy.i0 * w;
    const float* r11 = p1 + (size_t)cy.i1 * w;
    float* yp = y + ((size_t)cz * oh + oy) * ow + x0;
    for (int i4 = threadIdx.x * 4; i4 < nx; i4 += blockDim.x * 4) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = min(i4 + q, nx - 1);
        const int a0 = xi0[i], a1 = xi1[i];
        const float w1 = xw1[i], w0 = 1.f - w1;
        float t = cy.w0 * (w0 * __ldg(r00 + a0) + w1 * __ldg(r00 + a1)) + cy.w1 * (w0 * __ldg(r01 + a0) + w1 * __ldg(r01 + a1));
        if (three_d) {
          const float t1 = cy.w0 * (w0 * __ldg(r10 + a0) + w1 * __ldg(r10 + a1)) + cy.w1 * (w0 * __ldg(r11 + a0) + w1 * __ldg(r11 + a1));
          t = czc.w0 * t + czc.w1 * t1;
        }
        v[q] = t;
      }
      if (vec && i4 + 3 < nx) {
        *reinterpret_cast<float4*>(yp + i4) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i4 + q < nx) yp[i4 + q] = v[q];
      }
    }
  }
}

// backward as a gather (deterministic): each input element sums the output elements whose
// stencils touch it.  For an integer-ish scale the candidate output range per axis is small;
// we bound it by scanning outputs o with src(o) in (i-1, i+1).
__device__ __forceinline__ void out_range(int i, int in_size, int out_size, int ac, int* lo, int* hi) {
  float scale, inv;
  if (ac) {
    scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    if (scale == 0.f) { *lo = 0; *hi = out_size - 1; return; }
    inv = 1.f / scale;
    *lo = (int)floorf((i - 1) * inv) - 1;
    *hi = (int)ceilf((i + 1) * inv) + 1;
  } else {
    scale = (float)in_size / (float)out_size;
    inv = 1.f / scale;
    *lo = (int)floorf((i - 1 + 0.5f) * inv - 0.5f) - 1;
    *hi = (int)ceilf((i + 1 + 0.5f) * inv - 0.5f) + 1;
    if (i == 0) *lo = 0;  // clamped negative sources all land on index 0
  }
  if (*lo < 0) *lo = 0;
  if (*hi > out_size - 1) *hi = out_size - 1;
}

__global__ void __launch_bounds__(256) upsample_linear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                                 int d, int h, int w, int od, int oh, int ow, int ac) {
  const int cz = blockIdx.z;                     // c * d + iz
  const int c = cz / d, iz = cz - c * d;
  const bool three_d = !(d == 1 && od == 1);
  for (int iy = blockIdx.y * kRowsPerBlock; iy < min((int)(blockIdx.y + 1) * kRowsPerBlock, h); ++iy) {
  int yl, yh, zl = 0, zh = 0;
  out_range(iy, h, oh, ac, &yl, &yh);
  if (three_d) out_range(iz, d, od, ac, &zl, &zh);
  const float* g = dy + (size_t)c * od * oh * ow;
  float* dp = dx + ((size_t)cz * h + iy) * w;
  for (int ix = blockIdx.x * blockDim.x + threadIdx.x; ix < w; ix += gridDim.x * blockDim.x) {
    int xl, xh;
    out_range(ix, w, ow, ac, &xl, &xh);
    float s = 0.f;
    for (int oz = zl; oz <= zh; ++oz) {
      float wz = 1.f;
      if (three_d) {
        const LinCoord cc = lin_coord(oz, d, od, ac);
        wz = (cc.i0 == iz ? cc.w0 : 0.f) + (cc.i1 == iz ? cc.w1 : 0.f);
        if (wz == 0.f) continue;
      }
      for (int oy = yl; oy <= yh; ++oy) {
        const LinCoord cy = lin_coord(oy, h, oh, ac);
        const float wy = (cy.i0 == iy ? cy.w0 : 0.f) + (cy.i1 == iy ? cy.w1 : 0.f);
        if (wy == 0.f) continue;
        const float* gr = g + ((size_t)oz * oh + oy) * ow;
        for (int ox = xl; ox <= xh; ++ox) {
          const LinCoord cx = lin_coord(ox, w, ow, ac);
          const float wx = (cx.i0 == ix ? cx.w0 : 0.f) + (cx.i1 == ix ? cx.w1 : 0.f);
          if (wx != 0.f) s = fmaf(wz * wy * wx, __ldg(gr + ox), s);
        }
      }
    }
    dp[ix] = s;
  }
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_pixel_shuffle(const float* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w_,
                                 int32_t r, int inverse, void* stream) {
  VSR_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w_ > 0 && r >= 1, "vsr_pixel_shuffle: bad arguments");
  VSR_CHECK_SUPPORTED((long)n * c <= 65535 && (long)h * r <= 65535, "vsr_pixel_shuffle: n*c and h*r must be <= 65535");
  dim3 grid(1, (h * r + kRowsPerBlock - 1) / kRowsPerBlock, n * c);
  pixel_shuffle_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, c, h, w_, r, inverse);
  VSR_CHECK_LAUNCH("vsr_pixel_shuffle");
  return VSR_OK;
}

extern "C" int vsr_upsample_linear(const float* x, float* y, int32_t nc, int32_t d, int32_t h, int32_t w_,
                                   int32_t od, int32_t oh, int32_t ow, int align_corners, void* stream) {
  VSR_CHECK_ARG(x && y && nc > 0 && d > 0 && h > 0 && w_ > 0 && od > 0 && oh > 0 && ow > 0, "vsr_upsample_linear: bad arguments");
  VSR_CHECK_SUPPORTED((long)nc * od <= 65535 && oh <= 65535, "vsr_upsample_linear: nc*od and oh must be <= 65535");
  dim3 grid((ow + kUpTile - 1) / kUpTile, (oh + kRowsPerBlock - 1) / kRowsPerBlock, nc * od);
  upsample_linear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, d, h, w_, od, oh, ow, align_corners);
  VSR_CHECK_LAUNCH("vsr_upsample_linear");
  return VSR_OK;
}

extern "C" int vsr_upsample_linear_bwd(const float* dy, float* dx, int32_t nc, int32_t d, int32_t h, int32_t w_,
                                       int32_t od, int32_t oh, int32_t ow, int align_corners, void* stream) {
  VSR_CHECK_ARG(dy && dx && nc > 0 && d > 0 && h > 0 && w_ > 0 && od > 0 && oh > 0 && ow > 0, "vsr_upsample_linear_bwd: bad arguments");
  VSR_CHECK_SUPPORTED((long)nc * d <= 65535 && h <= 65535, "vsr_upsample_linear_bwd: nc*d and h must be <= 65535");
  dim3 grid((w_ + 255) / 256, (h + kRowsPerBlock - 1) / kRowsPerBlock, nc * d);
  upsample_linear_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, d, h, w_, od, oh, ow, align_corners);
  VSR_CHECK_LAUNCH("vsr_upsample_linear_bwd");
  return VSR_OK;
}

// ---- device-side data front end ---------------------------------------------------------------------------
// One batch of training items from cine volumes resident in device memory: temporal window, horizontal / vertical
// flip, crop (taken from the flipped image) and normalisation, written as collated frames.
// tab: int32 [n][5 + nf] = {sequence, flip_x, flip_y, y0, x0, frame_0 .. frame_{nf-1}} (y0, x0 in LR pixels).
namespace vsr {
namespace {
__global__ void __launch_bounds__(256) cine_gather_kernel(const float* __restrict__ vol, int T, int H, int W,
                                                         const int* __restrict__ tab, int ld, int n, int r, int f_first,
                                                         int f_count, int PH, int PW, float mean, float std,
                                                         float* __restrict__ out) {
  const long total = (long)f_count * n * PH * PW;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % PW);
    long q = i / PW;
    const int y = (int)(q % PH);
    q /= PH;
    const int item = (int)(q % n), f = (int)(q / n);
    const int* t = tab + (long)item * ld;
    int sy = __ldg(t + 3) * r + y, sx = __ldg(t + 4) * r + x;
    if (__ldg(t + 2)) sy = H - 1 - sy;
    if (__ldg(t + 1)) sx = W - 1 - sx;
    const float v = __ldg(vol + (((long)__ldg(t) * T + __ldg(t + 5 + f_first + f)) * H + sy) * W + sx);
    out[i] = __fdiv_rn(__fsub_rn(v, mean), std);      // Normalize (transforms.py:154-168) in the host's fp32 order
  }
}
}  // namespace
}  // namespace vsr

extern "C" int vsr_cine_gather(const float* vol, int32_t seqs, int32_t frames, int32_t h, int32_t w_, const int32_t* tab,
                               int32_t n, int32_t nf, int32_t r, int32_t f_first, int32_t f_count, int32_t ph, int32_t pw,
                               float mean, float std, float* out, void* stream) {
  using namespace vsr;
  VSR_CHECK_ARG(vol && tab && out && seqs > 0 && frames > 0 && h > 0 && w_ > 0 && n > 0 && nf > 0 && r >= 1,
                "vsr_cine_gather: bad arguments");
  VSR_CHECK_ARG(f_first >= 0 && f_count > 0 && f_first + f_count <= nf && ph > 0 && pw > 0 && ph * r <= h && pw * r <= w_,
                "vsr_cine_gather: frame range / patch outside the volume");
  VSR_CHECK_ARG(std != 0.f, "vsr_cine_gather: std must not be zero");
  const long total = (long)f_count * n * ph * r * pw * r;
  cine_gather_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      vol, frames, h, w_, tab, 5 + nf, n, r, f_first, f_count, ph * r, pw * r, mean, std, out);
  VSR_CHECK_LAUNCH("vsr_cine_gather");
  return VSR_OK;
}
