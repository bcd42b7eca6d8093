// flow.cu — the element-wise pieces of the flow-based recurrent net FRVSRNet (frvsr_net.py:11-239) that the tap-GEMM does
// not cover: 2x2 max pooling and bilinear x2 up-sampling of pixel-major feature maps (FNet, :117-135,166-172), tanh on the
// flow head (:142), the spatial-transformer warp (STN: mesh + flow, grid_sample bilinear / border, :194-226) with its
// gradient with respect to the flow, and space-to-depth + concatenation in front of SRNet (:47-48,88,175-191).
// fp32 maps (the net runs in the strict mode).  None of these is on BASELINE's headline path; they are written for
// coalesced access (a thread per channel of a pixel-major map, per pixel of an image), not tuned further.
#include "common.cuh"

namespace vsr {
namespace {

// ---- 2x2 max pooling, [n][h][w][c] -> [n][h/2][w/2][c]; idx keeps the window position (0..3) of the maximum ------------
__global__ void maxpool2_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ y,
                                uint8_t* __restrict__ idx) {
  const int oh = h / 2, ow = w / 2;
  const long total = (long)n * oh * ow * c;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long q = i / c;
    const int ox = (int)(q % ow);
    q /= ow;
    const int oy = (int)(q % oh), b = (int)(q / oh);
    const float* p = x + (((size_t)b * h + 2 * oy) * w + 2 * ox) * c + ch;
    // scan order of torch.nn.MaxPool2d: a later element wins only when strictly greater (or NaN)
    float m = p[0];
    int k = 0;
    const float v1 = p[c], v2 = p[(size_t)w * c], v3 = p[(size_t)w * c + c];
    if (v1 > m || v1 != v1) { m = v1; k = 1; }
    if (v2 > m || v2 != v2) { m = v2; k = 2; }
    if (v3 > m || v3 != v3) { m = v3; k = 3; }
    y[i] = m;
    idx[i] = (uint8_t)k;
  }
}
__global__ void maxpool2_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx, int n, int h, int w, int c,
                                    float* __restrict__ dx) {
  const int oh = h / 2, ow = w / 2;
  const long total = (long)n * oh * ow * c;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long q = i / c;
    const int ox = (int)(q % ow);
    q /= ow;
    const int oy = (int)(q % oh), b = (int)(q / oh);
    float* p = dx + (((size_t)b * h + 2 * oy) * w + 2 * ox) * c + ch;
    const int k = idx[i];
    const float g = dy[i];
    p[0] = k == 0 ? g : 0.f;
    p[c] = k == 1 ? g : 0.f;
    p[(size_t)w * c] = k == 2 ? g : 0.f;
    p[(size_t)w * c + c] = k == 3 ? g : 0.f;
  }
}

// ---- bilinear x2 (align_corners = False) of a pixel-major map: output 2i + p reads i (0.75) and i -+ 1 (0.25), clamped ---
__global__ void up2_nhwc_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ y) {
  const int oh = 2 * h, ow = 2 * w;
  const long total = (long)n * oh * ow * c;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long q = i / c;
    const int ox = (int)(q % ow);
    q /= ow;
    const int oy = (int)(q % oh), b = (int)(q / oh);
    const int iy = oy >> 1, ix = ox >> 1;
    const int sy = min(max(iy + ((oy & 1) ? 1 : -1), 0), h - 1), sx = min(max(ix + ((ox & 1) ? 1 : -1), 0), w - 1);
    const float* xb = x + (size_t)b * h * w * c + ch;
    const float a = xb[((size_t)iy * w + ix) * c], bq = xb[((size_t)iy * w + sx) * c];
    const float cq = xb[((size_t)sy * w + ix) * c], d = xb[((size_t)sy * w + sx) * c];
    y[i] = 0.75f * (0.75f * a + 0.25f * bq) + 0.25f * (0.75f * cq + 0.25f * d);
  }
}
// dx(iy, ix) = sum over the outputs that read it; per axis: rows 2i, 2i+1 with 0.75, rows 2i-1 and 2i+2 with 0.25, the
// clamped neighbour of a border row folds back onto it
__global__ void up2_nhwc_bwd_kernel(const float* __restrict__ dy, int n, int h, int w, int c, float* __restrict__ dx) {
  const int oh = 2 * h, ow = 2 * w;
  const long total = (long)n * h * w * c;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long q = i / c;
    const int ix = (int)(q % w);
    q /= w;
    const int iy = (int)(q % h), b = (int)(q / h);
    const float* gb = dy + (size_t)b * oh * ow * c + ch;
    // weights of output rows 2iy-1 .. 2iy+2 onto input row iy (and the same for columns)
    float wy[4] = {0.25f, 0.75f, 0.75f, 0.25f}, wx[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    if (iy == 0) { wy[0] = 0.f; wy[1] = 1.f; }            // output row 0 reads (0.75 + 0.25) of input row 0
    if (iy == h - 1) { wy[3] = 0.f; wy[2] = 1.f; }
    if (ix == 0) { wx[0] = 0.f; wx[1] = 1.f; }
    if (ix == w - 1) { wx[3] = 0.f; wx[2] = 1.f; }
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int oy = 2 * iy - 1 + a;
      if (wy[a] == 0.f) continue;
      float r = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ox = 2 * ix - 1 + e;
        if (wx[e] != 0.f) r = fmaf(wx[e], gb[((size_t)oy * ow + ox) * c], r);
      }
      s = fmaf(wy[a], r, s);
    }
    dx[i] = s;
  }
}

// ---- flow head: flow[b][k][y][x] = tanh(z[b][y0 + y][x0 + x][k]), k < 2 (z: [n][hp][wp][cz], the crop undoes FNet's padding)
__global__ void flow_tanh_kernel(const float* __restrict__ z, int n, int hp, int wp, int cz, int y0, int x0, int h, int w,
                                 float* __restrict__ flow) {
  const long total = (long)n * 2 * h * w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    long q = i / w;
    const int y = (int)(q % h);
    q /= h;
    const int k = (int)(q % 2), b = (int)(q / 2);
    flow[i] = tanhf(z[(((size_t)b * hp + y0 + y) * wp + x0 + x) * cz + k]);
  }
}
// dz = dflow * (1 - flow^2) inside the crop and for k < 2, zero elsewhere (every element of dz is written)
__global__ void flow_tanh_bwd_kernel(const float* __restrict__ dflow, const float* __restrict__ flow, int n, int hp, int wp,
                                     int cz, int y0, int x0, int h, int w, float* __restrict__ dz) {
  const long total = (long)n * hp * wp * cz;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % cz);
    long q = i / cz;
    const int xp = (int)(q % wp);
    q /= wp;
    const int yp = (int)(q % hp), b = (int)(q / hp);
    const int y = yp - y0, x = xp - x0;
    float g = 0.f;
    if (k < 2 && y >= 0 && y < h && x >= 0 && x < w) {
      const size_t j = (((size_t)b * 2 + k) * h + y) * w + x;
      const float f = flow[j];
      g = dflow[j] * (1.f - f * f);
    }
    dz[i] = g;
  }
}

// ---- STN warp of a one-channel image: grid = linspace(-1, 1) mesh + flow, grid_sample(bilinear, border, align_corners =
// False) as ATen computes it (grid_sampler_2d: unnormalise, clip to [0, size - 1], corner weights) ----------------------
struct Sample {
  int x0, y0;
  float fx, fy;      // fractional parts
  float mx, my;      // d(unnormalised, clipped coordinate) / d(grid coordinate)
};
__device__ __forceinline__ Sample locate(float gx, float gy, int H, int W) {
  Sample s;
  float x = ((gx + 1.f) * W - 1.f) * 0.5f, y = ((gy + 1.f) * H - 1.f) * 0.5f;
  s.mx = 0.5f * W;
  s.my = 0.5f * H;
  if (x <= 0.f) { x = 0.f; s.mx = 0.f; } else if (x >= (float)(W - 1)) { x = (float)(W - 1); s.mx = 0.f; }
  if (y <= 0.f) { y = 0.f; s.my = 0.f; } else if (y >= (float)(H - 1)) { y = (float)(H - 1); s.my = 0.f; }
  const float xf = floorf(x), yf = floorf(y);
  s.x0 = (int)xf; s.y0 = (int)yf;
  s.fx = x - xf; s.fy = y - yf;
  return s;
}
__device__ __forceinline__ float mesh_at(int i, int n) {       // np.linspace(-1, 1, n)[i]
  return n > 1 ? (float)(-1.0 + 2.0 * (double)i / (double)(n - 1)) : -1.f;
}
__global__ void grid_warp_kernel(const float* __restrict__ img, const float* __restrict__ flow, int n, int H, int W,
                                 float* __restrict__ out) {
  const long total = (long)n * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long q = i / W;
    const int y = (int)(q % H), b = (int)(q / H);
    const float* fb = flow + (size_t)b * 2 * H * W + (size_t)y * W + x;
    const Sample s = locate(mesh_at(x, W) + fb[0], mesh_at(y, H) + fb[(size_t)H * W], H, W);
    const float* ib = img + (size_t)b * H * W;
    auto at = [&](int yy, int xx) -> float { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? ib[(size_t)yy * W + xx] : 0.f; };
    const float nw = at(s.y0, s.x0), ne = at(s.y0, s.x0 + 1), sw = at(s.y0 + 1, s.x0), se = at(s.y0 + 1, s.x0 + 1);
    out[i] = nw * (1.f - s.fx) * (1.f - s.fy) + ne * s.fx * (1.f - s.fy) + sw * (1.f - s.fx) * s.fy + se * s.fx * s.fy;
  }
}
// gradient with respect to the flow only (the warped image is data or a detached output, frvsr_net.py:47,53)
__global__ void grid_warp_bwd_kernel(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ dout,
                                     int n, int H, int W, float* __restrict__ dflow) {
  const long total = (long)n * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long q = i / W;
    const int y = (int)(q % H), b = (int)(q / H);
    const float* fb = flow + (size_t)b * 2 * H * W + (size_t)y * W + x;
    const Sample s = locate(mesh_at(x, W) + fb[0], mesh_at(y, H) + fb[(size_t)H * W], H, W);
    const float* ib = img + (size_t)b * H * W;
    auto at = [&](int yy, int xx) -> float { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? ib[(size_t)yy * W + xx] : 0.f; };
    const float nw = at(s.y0, s.x0), ne = at(s.y0, s.x0 + 1), sw = at(s.y0 + 1, s.x0), se = at(s.y0 + 1, s.x0 + 1);
    const float g = dout[i];
    const float gx = (-nw * (1.f - s.fy) + ne * (1.f - s.fy) - sw * s.fy + se * s.fy) * g;
    const float gy = (-nw * (1.f - s.fx) - ne * s.fx + sw * (1.f - s.fx) + se * s.fx) * g;
    float* db = dflow + (size_t)b * 2 * H * W + (size_t)y * W + x;
    db[0] = s.mx * gx;
    db[(size_t)H * W] = s.my * gy;
  }
}

// ---- space-to-depth of the warped HR image + the LR frame -> pixel-major SRNet input [n][h][w][cpad]:
// channel py * r + px = hr[y * r + py][x * r + px] (SpaceToDepth.forward), channel r * r = lr, the rest zero ---------------
__global__ void s2d_cat_kernel(const float* __restrict__ hr, const float* __restrict__ lr, int n, int h, int w, int r, int cpad,
                               float* __restrict__ out) {
  const long total = (long)n * h * w * cpad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % cpad);
    long q = i / cpad;
    const int x = (int)(q % w);
    q /= w;
    const int y = (int)(q % h), b = (int)(q / h);
    float v = 0.f;
    if (ch < r * r) v = hr[((size_t)b * h * r + (size_t)y * r + ch / r) * ((size_t)w * r) + (size_t)x * r + ch % r];
    else if (ch == r * r) v = lr[((size_t)b * h + y) * w + x];
    out[i] = v;
  }
}
__global__ void s2d_cat_bwd_kernel(const float* __restrict__ dout, int n, int h, int w, int r, int cpad, float* __restrict__ dhr) {
  const long total = (long)n * h * r * w * r;
  const int W = w * r;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const long q = i / W;
    const int Y = (int)(q % (h * r)), b = (int)(q / (h * r));
    dhr[i] = dout[(((size_t)b * h + Y / r) * w + X / r) * cpad + (Y % r) * r + X % r];
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

#define FLOW_LAUNCH(kernel, items, ...)                                                         \
  do {                                                                                          \
    kernel<<<grid_for((items), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__); \
    VSR_CHECK_LAUNCH(#kernel);                                                                  \
  } while (0)

extern "C" int vsr_maxpool2x2(const float* x, int32_t n, int32_t h, int32_t w_, int32_t c, float* y, uint8_t* idx, void* stream) {
  VSR_CHECK_ARG(x && y && idx && n > 0 && h > 0 && w_ > 0 && c > 0, "vsr_maxpool2x2: bad arguments");
  VSR_CHECK_SUPPORTED(h % 2 == 0 && w_ % 2 == 0, "vsr_maxpool2x2: even sizes only (%d x %d)", h, w_);
  FLOW_LAUNCH(maxpool2_kernel, (long)n * (h / 2) * (w_ / 2) * c, x, n, h, w_, c, y, idx);
  return VSR_OK;
}
extern "C" int vsr_maxpool2x2_bwd(const float* dy, const uint8_t* idx, int32_t n, int32_t h, int32_t w_, int32_t c, float* dx,
                                  void* stream) {
  VSR_CHECK_ARG(dy && dx && idx && n > 0 && h > 0 && w_ > 0 && c > 0 && h % 2 == 0 && w_ % 2 == 0, "vsr_maxpool2x2_bwd: bad arguments");
  FLOW_LAUNCH(maxpool2_bwd_kernel, (long)n * (h / 2) * (w_ / 2) * c, dy, idx, n, h, w_, c, dx);
  return VSR_OK;
}
extern "C" int vsr_upsample2x_nhwc(const float* x, int32_t n, int32_t h, int32_t w_, int32_t c, float* y, void* stream) {
  VSR_CHECK_ARG(x && y && n > 0 && h > 0 && w_ > 0 && c > 0, "vsr_upsample2x_nhwc: bad arguments");
  FLOW_LAUNCH(up2_nhwc_kernel, (long)n * 4 * h * w_ * c, x, n, h, w_, c, y);
  return VSR_OK;
}
extern "C" int vsr_upsample2x_nhwc_bwd(const float* dy, int32_t n, int32_t h, int32_t w_, int32_t c, float* dx, void* stream) {
  VSR_CHECK_ARG(dy && dx && n > 0 && h > 0 && w_ > 0 && c > 0, "vsr_upsample2x_nhwc_bwd: bad arguments");
  FLOW_LAUNCH(up2_nhwc_bwd_kernel, (long)n * h * w_ * c, dy, n, h, w_, c, dx);
  return VSR_OK;
}
extern "C" int vsr_flow_tanh(const float* z, int32_t n, int32_t hp, int32_t wp, int32_t cz, int32_t y0, int32_t x0, int32_t h,
                             int32_t w_, float* flow, void* stream) {
  VSR_CHECK_ARG(z && flow && n > 0 && cz >= 2 && y0 >= 0 && x0 >= 0 && h > 0 && w_ > 0 && y0 + h <= hp && x0 + w_ <= wp,
                "vsr_flow_tanh: bad arguments");
  FLOW_LAUNCH(flow_tanh_kernel, (long)n * 2 * h * w_, z, n, hp, wp, cz, y0, x0, h, w_, flow);
  return VSR_OK;
}
extern "C" int vsr_flow_tanh_bwd(const float* dflow, const float* flow, int32_t n, int32_t hp, int32_t wp, int32_t cz, int32_t y0,
                                 int32_t x0, int32_t h, int32_t w_, float* dz, void* stream) {
  VSR_CHECK_ARG(dflow && flow && dz && n > 0 && cz >= 2 && y0 >= 0 && x0 >= 0 && h > 0 && w_ > 0 && y0 + h <= hp && x0 + w_ <= wp,
                "vsr_flow_tanh_bwd: bad arguments");
  FLOW_LAUNCH(flow_tanh_bwd_kernel, (long)n * hp * wp * cz, dflow, flow, n, hp, wp, cz, y0, x0, h, w_, dz);
  return VSR_OK;
}
extern "C" int vsr_grid_warp(const float* img, const float* flow, int32_t n, int32_t h, int32_t w_, float* out, void* stream) {
  VSR_CHECK_ARG(img && flow && out && n > 0 && h > 0 && w_ > 0, "vsr_grid_warp: bad arguments");
  FLOW_LAUNCH(grid_warp_kernel, (long)n * h * w_, img, flow, n, h, w_, out);
  return VSR_OK;
}
extern "C" int vsr_grid_warp_bwd(const float* img, const float* flow, const float* dout, int32_t n, int32_t h, int32_t w_,
                                 float* dflow, void* stream) {
  VSR_CHECK_ARG(img && flow && dout && dflow && n > 0 && h > 0 && w_ > 0, "vsr_grid_warp_bwd: bad arguments");
  FLOW_LAUNCH(grid_warp_bwd_kernel, (long)n * h * w_, img, flow, dout, n, h, w_, dflow);
  return VSR_OK;
}
extern "C" int vsr_s2d_cat(const float* hr, const float* lr, int32_t n, int32_t h, int32_t w_, int32_t r, int32_t cpad, float* out,
                           void* stream) {
  VSR_CHECK_ARG(hr && lr && out && n > 0 && h > 0 && w_ > 0 && r >= 1 && cpad >= r * r + 1, "vsr_s2d_cat: bad arguments");
  FLOW_LAUNCH(s2d_cat_kernel, (long)n * h * w_ * cpad, hr, lr, n, h, w_, r, cpad, out);
  return VSR_OK;
}
extern "C" int vsr_s2d_cat_bwd(const float* dout, int32_t n, int32_t h, int32_t w_, int32_t r, int32_t cpad, float* dhr, void* stream) {
  VSR_CHECK_ARG(dout && dhr && n > 0 && h > 0 && w_ > 0 && r >= 1 && cpad >= r * r + 1, "vsr_s2d_cat_bwd: bad arguments");
  FLOW_LAUNCH(s2d_cat_bwd_kernel, (long)n * h * r * w_ * r, dout, n, h, w_, r, cpad, dhr);
  return VSR_OK;
}
