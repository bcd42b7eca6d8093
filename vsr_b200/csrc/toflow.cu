// toflow.cu — the non-convolutional pieces of TOFlowNet (reference: src/model/nets/toflow_net.py:33-138): bicubic
// up-sampling of the input frames, padding with the batch minimum, the 2x2 average-pooling pyramid of SpyNet, the flow warp
// (pixel-unit flow, bilinear, zeros outside) fused with the concatenation that feeds a SpyNet block / the output block, the
// flow update, and the output head (+ reference frame, crop).  Every convolution and BatchNorm runs on the tap-GEMM /
// vsr_bn_* kernels; maps are pixel-major [n][h][w][c], images and flows planar [n][k][h][w].
#include "common.cuh"

namespace vsr {
namespace {

// ---- F.interpolate(mode='bicubic', align_corners=False, scale_factor=r) as ATen's upsample_bicubic2d computes it:
// source = (dst + 0.5) / r - 0.5 (not clamped), cubic convolution coefficients with A = -0.75, indices clamped ----------
__device__ __forceinline__ float cc1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cc2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coef(float t, float* c) {
  const float A = -0.75f;
  c[0] = cc2(t + 1.f, A);
  c[1] = cc1(t, A);
  c[2] = cc1(1.f - t, A);
  c[3] = cc2(2.f - t, A);
}
__global__ void bicubic_up_kernel(const float* __restrict__ x, int nc, int h, int w, int r, float* __restrict__ y) {
  const int H = h * r, W = w * r;
  const long total = (long)nc * H * W;
  const float scale = 1.f / (float)r;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W);
    const long q = i / W;
    const int oy = (int)(q % H), b = (int)(q / H);
    const float sx = scale * ((float)ox + 0.5f) - 0.5f, sy = scale * ((float)oy + 0.5f) - 0.5f;
    const float fx = floorf(sx), fy = floorf(sy);
    const int ix = (int)fx, iy = (int)fy;
    float cx[4], cy[4];
    cubic_coef(sx - fx, cx);
    cubic_coef(sy - fy, cy);
    const float* xb = x + (size_t)b * h * w;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int yy = min(max(iy - 1 + j, 0), h - 1);
      float row = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) row += xb[(size_t)yy * w + min(max(ix - 1 + k, 0), w - 1)] * cx[k];
      acc += row * cy[j];
    }
    y[i] = acc;
  }
}

// ---- x.min() over a whole tensor, kept on the device: every block writes its partial minimum (block 0 fills the unused
// rows of the kPartialsLen-long vector with +inf); the consumer folds the vector ------------------------------------------
__global__ void __launch_bounds__(256) min_partials_kernel(const float* __restrict__ x, long n, float* __restrict__ partials) {
  __shared__ float red[8];
  float m = __int_as_float(0x7f800000);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) m = fminf(m, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) m = fminf(m, red[k]);
    partials[blockIdx.x] = m;
  }
  if (blockIdx.x == 0)
    for (int k = gridDim.x + threadIdx.x; k < kPartialsLen; k += blockDim.x) partials[k] = __int_as_float(0x7f800000);
}
// F.pad(x, (x0, wp - w - x0, y0, hp - h - y0), value=x.min()) for planar images [nc][h][w] -> [nc][hp][wp]
__global__ void __launch_bounds__(256) pad_fill_kernel(const float* __restrict__ x, int nc, int h, int w, int y0, int x0, int hp,
                                                       int wp, const float* __restrict__ partials, float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float fill_s;
  float m = __int_as_float(0x7f800000);
  for (int k = threadIdx.x; k < kPartialsLen; k += blockDim.x) m = fminf(m, partials[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) m = fminf(m, red[k]);
    fill_s = m;
  }
  __syncthreads();
  const float fill = fill_s;
  const long total = (long)nc * hp * wp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % wp) - x0;
    const long q = i / wp;
    const int yy = (int)(q % hp) - y0, b = (int)(q / hp);
    out[i] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? x[((size_t)b * h + yy) * w + xx] : fill;
  }
}

// F.avg_pool2d(x, 2, 2) of planar images (even h, w)
__global__ void avgpool2x2_kernel(const float* __restrict__ x, int nc, int h, int w, float* __restrict__ y) {
  const int ho = h / 2, wo = w / 2;
  const long total = (long)nc * ho * wo;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % wo);
    const long q = i / wo;
    const int oy = (int)(q % ho), b = (int)(q / ho);
    const float* p = x + ((size_t)b * h + 2 * oy) * w + 2 * ox;
    y[i] = (p[0] + p[1] + p[w] + p[w + 1]) * 0.25f;
  }
}

// ---- flow_warp (toflow_net.py:117-138): vgrid = pixel mesh + flow, normalised by (size - 1), grid_sample(bilinear, zeros,
// align_corners = False) as ATen un-normalises it; `scale` multiplies the stored flow (flow_up * 2.0, :82) ------------------
struct Warp {
  int x0, y0;
  float fx, fy;
  float nw, ne, sw, se;
};
__device__ __forceinline__ Warp warp_at(const float* __restrict__ img, int H, int W, int x, int y, float flx, float fly) {
  Warp s;
  const float gx = 2.f * ((float)x + flx) / (float)max(W - 1, 1) - 1.f, gy = 2.f * ((float)y + fly) / (float)max(H - 1, 1) - 1.f;
  const float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;
  const float xf = floorf(ix), yf = floorf(iy);
  s.x0 = (int)xf; s.y0 = (int)yf;
  s.fx = ix - xf; s.fy = iy - yf;
  auto at = [&](int yy, int xx) -> float { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(size_t)yy * W + xx] : 0.f; };
  s.nw = at(s.y0, s.x0); s.ne = at(s.y0, s.x0 + 1); s.sw = at(s.y0 + 1, s.x0); s.se = at(s.y0 + 1, s.x0 + 1);
  return s;
}
// out[pix][c_ref] = ref[pix] (if ref), out[pix][c_w] = warp(nbr, scale * flow)[pix] (if nbr), out[pix][c_flow + k] = scale *
// flow[k][pix] (if c_flow >= 0); the other channels of `out` are left as they are (the caller clears the buffer once)
__global__ void warp_cat_kernel(float* __restrict__ out, int n, int H, int W, int cpad, int c_ref, const float* __restrict__ ref,
                                int c_w, const float* __restrict__ nbr, const float* __restrict__ flow, float scale, int c_flow) {
  const long total = (long)n * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long q = i / W;
    const int y = (int)(q % H), b = (int)(q / H);
    float* o = out + (size_t)i * cpad;
    if (ref) o[c_ref] = ref[i];
    float flx = 0.f, fly = 0.f;
    if (flow) {
      const float* fb = flow + (size_t)b * 2 * H * W + (size_t)y * W + x;
      flx = scale * fb[0];
      fly = scale * fb[(size_t)H * W];
    }
    if (nbr) {
      const Warp s = warp_at(nbr + (size_t)b * H * W, H, W, x, y, flx, fly);
      o[c_w] = s.nw * (1.f - s.fx) * (1.f - s.fy) + s.ne * s.fx * (1.f - s.fy) + s.sw * (1.f - s.fx) * s.fy + s.se * s.fx * s.fy;
    }
    if (c_flow >= 0) { o[c_flow] = flx; o[c_flow + 1] = fly; }
  }
}
// gradient with respect to the STORED flow: dflow[k] = scale * (dout[c_w] * d warp / d flow_k + dout[c_flow + k])
__global__ void warp_cat_bwd_kernel(const float* __restrict__ dout, int n, int H, int W, int cpad, int c_w,
                                    const float* __restrict__ nbr, const float* __restrict__ flow, float scale, int c_flow,
                                    float* __restrict__ dflow) {
  const long total = (long)n * H * W;
  const float mx = (float)W / (float)max(W - 1, 1), my = (float)H / (float)max(H - 1, 1);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long q = i / W;
    const int y = (int)(q % H), b = (int)(q / H);
    const float* fb = flow + (size_t)b * 2 * H * W + (size_t)y * W + x;
    const float* d = dout + (size_t)i * cpad;
    float gx = 0.f, gy = 0.f;
    if (nbr) {
      const Warp s = warp_at(nbr + (size_t)b * H * W, H, W, x, y, scale * fb[0], scale * fb[(size_t)H * W]);
      const float g = d[c_w];
      gx = (-s.nw * (1.f - s.fy) + s.ne * (1.f - s.fy) - s.sw * s.fy + s.se * s.fy) * g * mx;
      gy = (-s.nw * (1.f - s.fx) - s.ne * s.fx + s.sw * (1.f - s.fx) + s.se * s.fx) * g * my;
    }
    if (c_flow >= 0) { gx += d[c_flow]; gy += d[c_flow + 1]; }
    float* db = dflow + (size_t)b * 2 * H * W + (size_t)y * W + x;
    db[0] = scale * gx;
    db[(size_t)H * W] = scale * gy;
  }
}

// flow[b][k][y][x] = scale * flow_up[b][k][y][x] + z[b][y][x][k], k < 2 (toflow_net.py:83: flow = flow_up + block(...))
__global__ void flow_add_kernel(const float* __restrict__ z, int n, int H, int W, int cz, const float* __restrict__ flow_up,
                                float scale, float* __restrict__ flow) {
  const long total = (long)n * 2 * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i % ((long)H * W);
    const long q = i / ((long)H * W);
    const int k = (int)(q % 2), b = (int)(q / 2);
    flow[i] = scale * flow_up[i] + z[((size_t)b * H * W + pix) * cz + k];
  }
}

// dz[b][yp][xp][k] = d[b][k][yp - y0][xp - x0] inside the window and for k < kc, zero elsewhere: the gradient of "take the
// first kc channels of a pixel-major map (and crop)" - the flow update (kc = 2, no crop) and the output head (kc = 1)
__global__ void planar_to_nhwc_kernel(const float* __restrict__ d, int n, int kc, int hp, int wp, int cz, int y0, int x0, int h,
                                      int w, float* __restrict__ dz) {
  const long total = (long)n * hp * wp * cz;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % cz);
    long q = i / cz;
    const int xp = (int)(q % wp);
    q /= wp;
    const int yp = (int)(q % hp), b = (int)(q / hp);
    const int y = yp - y0, x = xp - x0;
    dz[i] = (k < kc && y >= 0 && y < h && x >= 0 && x < w) ? d[(((size_t)b * kc + k) * h + y) * w + x] : 0.f;
  }
}

// out[b][0][y][x] = z[b][y + y0][x + x0][0] + xref[b][y + y0][x + x0]  (toflow_net.py:59-65: out_block(x) + x_ref, crop)
__global__ void head_add_kernel(const float* __restrict__ z, int n, int hp, int wp, int cz, const float* __restrict__ xref, int y0,
                                int x0, int h, int w, float* __restrict__ out) {
  const long total = (long)n * h * w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long q = i / w;
    const int y = (int)(q % h), b = (int)(q / h);
    const size_t p = ((size_t)b * hp + y + y0) * wp + x + x0;
    out[i] = z[p * cz] + xref[p];
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

#define TOF_LAUNCH(kernel, items, ...)                                                          \
  do {                                                                                          \
    kernel<<<grid_for((items), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(__VA_ARGS__); \
    VSR_CHECK_LAUNCH(#kernel);                                                                  \
  } while (0)

extern "C" int vsr_upsample_bicubic(const float* x, int32_t nc, int32_t h, int32_t w_, int32_t r, float* y, void* stream) {
  VSR_CHECK_ARG(x && y && nc > 0 && h > 0 && w_ > 0 && r >= 1, "vsr_upsample_bicubic: bad arguments");
  TOF_LAUNCH(bicubic_up_kernel, (int64_t)nc * h * r * w_ * r, x, nc, h, w_, r, y);
  return VSR_OK;
}

extern "C" int vsr_min_partials(const float* x, int64_t numel, float* partials, void* stream) {
  VSR_CHECK_ARG(x && partials && numel > 0, "vsr_min_partials: bad arguments");
  int grid = grid_for(numel, 256 * 8);
  if (grid > kPartialsLen) grid = kPartialsLen;
  min_partials_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, numel, partials);
  VSR_CHECK_LAUNCH("vsr_min_partials");
  return VSR_OK;
}

extern "C" int vsr_pad_fill(const float* x, int32_t nc, int32_t h, int32_t w_, int32_t y0, int32_t x0, int32_t hp, int32_t wp,
                            const float* partials, float* out, void* stream) {
  VSR_CHECK_ARG(x && out && partials && nc > 0 && y0 >= 0 && x0 >= 0 && y0 + h <= hp && x0 + w_ <= wp, "vsr_pad_fill: bad arguments");
  TOF_LAUNCH(pad_fill_kernel, (int64_t)nc * hp * wp, x, nc, h, w_, y0, x0, hp, wp, partials, out);
  return VSR_OK;
}

extern "C" int vsr_avgpool2x2(const float* x, int32_t nc, int32_t h, int32_t w_, float* y, void* stream) {
  VSR_CHECK_ARG(x && y && nc > 0 && h > 0 && w_ > 0 && h % 2 == 0 && w_ % 2 == 0, "vsr_avgpool2x2: even sizes expected");
  TOF_LAUNCH(avgpool2x2_kernel, (int64_t)nc * (h / 2) * (w_ / 2), x, nc, h, w_, y);
  return VSR_OK;
}

extern "C" int vsr_warp_cat(float* out, int32_t n, int32_t h, int32_t w_, int32_t cpad, int32_t c_ref, const float* ref,
                            int32_t c_w, const float* nbr, const float* flow, float scale, int32_t c_flow, void* stream) {
  VSR_CHECK_ARG(out && n > 0 && h > 0 && w_ > 0 && cpad > 0, "vsr_warp_cat: bad arguments");
  VSR_CHECK_ARG((!ref || (c_ref >= 0 && c_ref < cpad)) && (!nbr || (c_w >= 0 && c_w < cpad)) && c_flow + 1 < cpad &&
                    (c_flow < 0 || flow), "vsr_warp_cat: channel out of range");
  TOF_LAUNCH(warp_cat_kernel, (int64_t)n * h * w_, out, n, h, w_, cpad, c_ref, ref, c_w, nbr, flow, scale, c_flow);
  return VSR_OK;
}

extern "C" int vsr_warp_cat_bwd(const float* dout, int32_t n, int32_t h, int32_t w_, int32_t cpad, int32_t c_w, const float* nbr,
                                const float* flow, float scale, int32_t c_flow, float* dflow, void* stream) {
  VSR_CHECK_ARG(dout && flow && dflow && n > 0 && h > 0 && w_ > 0 && cpad > 0 && c_flow + 1 < cpad, "vsr_warp_cat_bwd: bad arguments");
  TOF_LAUNCH(warp_cat_bwd_kernel, (int64_t)n * h * w_, dout, n, h, w_, cpad, c_w, nbr, flow, scale, c_flow, dflow);
  return VSR_OK;
}

extern "C" int vsr_flow_add(const float* z, int32_t n, int32_t h, int32_t w_, int32_t cz, const float* flow_up, float scale,
                            float* flow, void* stream) {
  VSR_CHECK_ARG(z && flow_up && flow && n > 0 && h > 0 && w_ > 0 && cz >= 2, "vsr_flow_add: bad arguments");
  TOF_LAUNCH(flow_add_kernel, (int64_t)n * 2 * h * w_, z, n, h, w_, cz, flow_up, scale, flow);
  return VSR_OK;
}

extern "C" int vsr_planar_to_nhwc(const float* d, int32_t n, int32_t kc, int32_t hp, int32_t wp, int32_t cz, int32_t y0,
                                  int32_t x0, int32_t h, int32_t w_, float* dz, void* stream) {
  VSR_CHECK_ARG(d && dz && n > 0 && kc >= 1 && kc <= cz && y0 >= 0 && x0 >= 0 && y0 + h <= hp && x0 + w_ <= wp,
                "vsr_planar_to_nhwc: bad arguments");
  TOF_LAUNCH(planar_to_nhwc_kernel, (int64_t)n * hp * wp * cz, d, n, kc, hp, wp, cz, y0, x0, h, w_, dz);
  return VSR_OK;
}

extern "C" int vsr_head_add(const float* z, int32_t n, int32_t hp, int32_t wp, int32_t cz, const float* xref, int32_t y0,
                            int32_t x0, int32_t h, int32_t w_, float* out, void* stream) {
  VSR_CHECK_ARG(z && xref && out && n > 0 && y0 >= 0 && x0 >= 0 && y0 + h <= hp && x0 + w_ <= wp, "vsr_head_add: bad arguments");
  TOF_LAUNCH(head_add_kernel, (int64_t)n * h * w_, z, n, hp, wp, cz, xref, y0, x0, h, w_, out);
  return VSR_OK;
}
