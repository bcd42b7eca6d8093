// resample.cu — standalone bandwidth kernels: nn.PixelShuffle and its inverse on NCHW
// (drf_net.py:142; the nets themselves never launch it — the shuffle is a reinterpretation of
// the phase-blocked layout) and F.interpolate(bilinear / trilinear) forward + backward
// (srfb_net.py:47; trilinear has no reference call site).
#include "common.cuh"

namespace vsr {
namespace {

// y[n][c][h*r+i][w*r+j] = x[n][c*r*r + i*r + j][h][w]; thread = one output element (coalesced
// writes; reads are r-strided within r*r planes, served by L1/L2).
__global__ void pixel_shuffle_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int c,
                                     int h, int w, int r, int inverse) {
  const long total = (long)n * c * h * r * w * r;
  const int W = w * r, H = h * r;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    if (!inverse) {
      const int X = (int)(i % W);
      long q = i / W;
      const int Y = (int)(q % H);
      q /= H;
      const int ci = (int)(q % c);
      const int ni = (int)(q / c);
      const size_t src = ((((size_t)ni * c + ci) * r * r + (Y % r) * r + (X % r)) * h + Y / r) * w + X / r;
      y[i] = __ldg(x + src);
    } else {
      // inverse: x is [n][c][H][W], y is [n][c*r*r][h][w]; thread = one output element of y
      const int xx = (int)(i % w);
      long q = i / w;
      const int yy = (int)(q % h);
      q /= h;
      const int ch = (int)(q % (c * r * r));
      const int ni = (int)(q / (c * r * r));
      const int ci = ch / (r * r), ph = ch % (r * r);
      const size_t src = (((size_t)ni * c + ci) * H + yy * r + ph / r) * W + xx * r + ph % r;
      y[i] = __ldg(x + src);
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

__global__ void upsample_linear_kernel(const float* __restrict__ x, float* __restrict__ y, int nc, int d,
                                       int h, int w, int od, int oh, int ow, int ac) {
  const long total = (long)nc * od * oh * ow;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    long q = i / ow;
    const int oy = (int)(q % oh);
    q /= oh;
    const int oz = (int)(q % od);
    const int c = (int)(q / od);
    const LinCoord cx = lin_coord(ox, w, ow, ac), cy = lin_coord(oy, h, oh, ac);
    const float* p = x + (size_t)c * d * h * w;
    float v;
    if (d == 1 && od == 1) {
      v = cy.w0 * (cx.w0 * __ldg(p + (size_t)cy.i0 * w + cx.i0) + cx.w1 * __ldg(p + (size_t)cy.i0 * w + cx.i1)) +
          cy.w1 * (cx.w0 * __ldg(p + (size_t)cy.i1 * w + cx.i0) + cx.w1 * __ldg(p + (size_t)cy.i1 * w + cx.i1));
    } else {
      const LinCoord cz = lin_coord(oz, d, od, ac);
      const float* p0 = p + (size_t)cz.i0 * h * w;
      const float* p1 = p + (size_t)cz.i1 * h * w;
      auto plane = [&](const float* pp) {
        return cy.w0 * (cx.w0 * __ldg(pp + (size_t)cy.i0 * w + cx.i0) + cx.w1 * __ldg(pp + (size_t)cy.i0 * w + cx.i1)) +
               cy.w1 * (cx.w0 * __ldg(pp + (size_t)cy.i1 * w + cx.i0) + cx.w1 * __ldg(pp + (size_t)cy.i1 * w + cx.i1));
      };
      v = cz.w0 * plane(p0) + cz.w1 * plane(p1);
    }
    y[i] = v;
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

__global__ void upsample_linear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int nc,
                                           int d, int h, int w, int od, int oh, int ow, int ac) {
  const long total = (long)nc * d * h * w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ix = (int)(i % w);
    long q = i / w;
    const int iy = (int)(q % h);
    q /= h;
    const int iz = (int)(q % d);
    const int c = (int)(q / d);
    int xl, xh, yl, yh, zl = 0, zh = 0;
    out_range(ix, w, ow, ac, &xl, &xh);
    out_range(iy, h, oh, ac, &yl, &yh);
    const bool three_d = !(d == 1 && od == 1);
    if (three_d) out_range(iz, d, od, ac, &zl, &zh);
    const float* g = dy + (size_t)c * od * oh * ow;
    float s = 0.f;
    for (int oz = zl; oz <= zh; ++oz) {
      float wz = 1.f;
      if (three_d) {
        const LinCoord cz = lin_coord(oz, d, od, ac);
        wz = (cz.i0 == iz ? cz.w0 : 0.f) + (cz.i1 == iz ? cz.w1 : 0.f);
        if (wz == 0.f) continue;
      }
      for (int oy = yl; oy <= yh; ++oy) {
        const LinCoord cy = lin_coord(oy, h, oh, ac);
        const float wy = (cy.i0 == iy ? cy.w0 : 0.f) + (cy.i1 == iy ? cy.w1 : 0.f);
        if (wy == 0.f) continue;
        for (int ox = xl; ox <= xh; ++ox) {
          const LinCoord cx = lin_coord(ox, w, ow, ac);
          const float wx = (cx.i0 == ix ? cx.w0 : 0.f) + (cx.i1 == ix ? cx.w1 : 0.f);
          if (wx == 0.f) continue;
          s = fmaf(wz * wy * wx, __ldg(g + ((size_t)oz * oh + oy) * ow + ox), s);
        }
      }
    }
    dx[i] = s;
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_pixel_shuffle(const float* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w_,
                                 int32_t r, int inverse, void* stream) {
  VSR_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w_ > 0 && r >= 1, "vsr_pixel_shuffle: bad arguments");
  const long total = (long)n * c * h * r * w_ * r;
  pixel_shuffle_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, c, h, w_, r, inverse);
  VSR_CHECK_LAUNCH("vsr_pixel_shuffle");
  return VSR_OK;
}

extern "C" int vsr_upsample_linear(const float* x, float* y, int32_t nc, int32_t d, int32_t h, int32_t w_,
                                   int32_t od, int32_t oh, int32_t ow, int align_corners, void* stream) {
  VSR_CHECK_ARG(x && y && nc > 0 && d > 0 && h > 0 && w_ > 0 && od > 0 && oh > 0 && ow > 0, "vsr_upsample_linear: bad arguments");
  const long total = (long)nc * od * oh * ow;
  upsample_linear_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, nc, d, h, w_, od, oh, ow, align_corners);
  VSR_CHECK_LAUNCH("vsr_upsample_linear");
  return VSR_OK;
}

extern "C" int vsr_upsample_linear_bwd(const float* dy, float* dx, int32_t nc, int32_t d, int32_t h, int32_t w_,
                                       int32_t od, int32_t oh, int32_t ow, int align_corners, void* stream) {
  VSR_CHECK_ARG(dy && dx && nc > 0 && d > 0 && h > 0 && w_ > 0 && od > 0 && oh > 0 && ow > 0, "vsr_upsample_linear_bwd: bad arguments");
  const long total = (long)nc * d * h * w_;
  upsample_linear_bwd_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, nc, d, h, w_, od, oh, ow, align_corners);
  VSR_CHECK_LAUNCH("vsr_upsample_linear_bwd");
  return VSR_OK;
}
