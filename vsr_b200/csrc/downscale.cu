// downscale.cu — the reference's `Downscale` (acdc_preprocess.py:102-180: centred k-space truncation, |.|, round,
// cv2.INTER_CUBIC resize by 1/r, round, clip to [0, 255]) for a batch of frames resident in device memory, so that the
// low-resolution side of a synthetic / pre-loaded cine dataset is produced without the host.
//
// The k-space truncation is linear and separable: with P_n the n x n complex matrix of
//   v -> fftshift(ifft(ifftshift(mask_n * fftshift(fft(ifftshift(v))))))
// (built once per (n, r) on the host in double precision) the low-passed frame is  P_h * X * P_w^T : two small dense
// complex GEMMs per frame instead of four FFTs - no cuFFT.  They run in FP64: the result is rounded to integers, and a
// frame must come out bit-identical to the reference's numpy (double precision, its pinned numpy 1.16) whenever the exact
// value is farther than ~1e-10 from x.5.  Integer-ratio bicubic with OpenCV's a = -0.75 kernel has the fixed weights
// (-3/32, 19/32, 19/32, -3/32) (even r: sample positions at .5) or (0, 1, 0, 0) (odd r) - exact in floating point for the
// integer-valued low-passed frame.
#include "common.cuh"

namespace vsr {
namespace {

constexpr int kT = 16;

// T[b] = P (h x h complex) * X[b] (h x w real).  grid (w / 16, h / 16, batch), block 16 x 16.
__global__ void __launch_bounds__(kT * kT) lowpass_rows_kernel(const double2* __restrict__ P, const float* __restrict__ x,
                                                              int h, int w, double2* __restrict__ t) {
  __shared__ double2 ps[kT][kT + 1];
  __shared__ double xs[kT][kT + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.y * kT + ty, j = blockIdx.x * kT + tx;
  const float* xb = x + (size_t)blockIdx.z * h * w;
  double re = 0.0, im = 0.0;
  for (int k0 = 0; k0 < h; k0 += kT) {
    ps[ty][tx] = (i < h && k0 + tx < h) ? P[(size_t)i * h + k0 + tx] : make_double2(0.0, 0.0);
    xs[ty][tx] = (k0 + ty < h && j < w) ? (double)xb[(size_t)(k0 + ty) * w + j] : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kT; ++k) {
      const double2 p = ps[ty][k];
      const double v = xs[k][tx];
      re = fma(p.x, v, re);
      im = fma(p.y, v, im);
    }
    __syncthreads();
  }
  if (i < h && j < w) t[((size_t)blockIdx.z * h + i) * w + j] = make_double2(re, im);
}

// Y[b](i, j) = rint(| sum_k T[b](i, k) * Q(j, k) |)   (Q = P_w), stored as fp32 (integer-valued)
__global__ void __launch_bounds__(kT * kT) lowpass_cols_kernel(const double2* __restrict__ t, const double2* __restrict__ Q,
                                                              int h, int w, float* __restrict__ y) {
  __shared__ double2 ts[kT][kT + 1];
  __shared__ double2 qs[kT][kT + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.y * kT + ty, j = blockIdx.x * kT + tx;
  const double2* tb = t + (size_t)blockIdx.z * h * w;
  double re = 0.0, im = 0.0;
  for (int k0 = 0; k0 < w; k0 += kT) {
    ts[ty][tx] = (i < h && k0 + tx < w) ? tb[(size_t)i * w + k0 + tx] : make_double2(0.0, 0.0);
    const int jq = blockIdx.x * kT + ty;                 // row of Q staged by this thread row
    qs[ty][tx] = (jq < w && k0 + tx < w) ? Q[(size_t)jq * w + k0 + tx] : make_double2(0.0, 0.0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kT; ++k) {
      const double2 a = ts[ty][k], b = qs[tx][k];
      re = fma(a.x, b.x, re);
      re = fma(-a.y, b.y, re);
      im = fma(a.x, b.y, im);
      im = fma(a.y, b.x, im);
    }
    __syncthreads();
  }
  if (i < h && j < w) y[((size_t)blockIdx.z * h + i) * w + j] = (float)rint(sqrt(re * re + im * im));
}

// cv2.resize(INTER_CUBIC) by the integer ratio r (source index clamped at the borders) + round (half to even) + clip
__global__ void __launch_bounds__(256) bicubic_down_kernel(const float* __restrict__ y, int n, int h, int w, int r,
                                                          float* __restrict__ out) {
  const int oh = h / r, ow = w / r;
  const long total = (long)n * oh * ow;
  const bool even = (r & 1) == 0;
  const double c0 = even ? -0.09375 : 0.0, c1 = even ? 0.59375 : 1.0, c2 = even ? 0.59375 : 0.0, c3 = even ? -0.09375 : 0.0;
  const double cw[4] = {c0, c1, c2, c3};
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % ow);
    const long q = idx / ow;
    const int oy = (int)(q % oh), b = (int)(q / oh);
    const int sx = even ? r * ox + r / 2 - 1 : r * ox + (r - 1) / 2;
    const int sy = even ? r * oy + r / 2 - 1 : r * oy + (r - 1) / 2;
    const float* yb = y + (size_t)b * h * w;
    // horizontal pass of the four source rows, then vertical (OpenCV's order); every partial sum is exact in double
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(sy - 1 + a, 0), h - 1);
      double row = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int xx = min(max(sx - 1 + c, 0), w - 1);
        row += cw[c] * (double)__ldg(yb + (size_t)yy * w + xx);
      }
      acc += cw[a] * row;
    }
    out[idx] = (float)fmin(fmax(rint(acc), 0.0), 255.0);
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" size_t vsr_downscale_workspace(int32_t n, int32_t h, int32_t w_) {
  return (size_t)n * h * w_ * (sizeof(double2) + sizeof(float));
}

// hr: [n][h][w] fp32 (integer-valued intensities); ph / pw: the h x h and w x w complex low-pass matrices (interleaved
// re, im; fp64; host-built per (size, r)); lr: [n][h / r][w / r] fp32.  h and w must be multiples of r.
extern "C" int vsr_downscale(const float* hr, int32_t n, int32_t h, int32_t w_, int32_t r, const double* ph, const double* pw,
                             float* lr, void* workspace, size_t workspace_bytes, void* stream) {
  VSR_CHECK_ARG(hr && ph && pw && lr && n > 0 && h > 0 && w_ > 0 && r >= 1, "vsr_downscale: bad arguments");
  VSR_CHECK_SUPPORTED(h % r == 0 && w_ % r == 0, "vsr_downscale: the frame size (%d x %d) must be a multiple of the factor %d", h, w_, r);
  VSR_CHECK_SUPPORTED(n <= 65535, "vsr_downscale: at most 65535 frames per call");
  VSR_CHECK_ARG(workspace && workspace_bytes >= vsr_downscale_workspace(n, h, w_), "vsr_downscale: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double2* t = static_cast<double2*>(workspace);
  float* y = reinterpret_cast<float*>(t + (size_t)n * h * w_);
  const dim3 grid((w_ + kT - 1) / kT, (h + kT - 1) / kT, n), block(kT, kT);
  lowpass_rows_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const double2*>(ph), hr, h, w_, t);
  VSR_CHECK_LAUNCH("vsr_downscale(rows)");
  lowpass_cols_kernel<<<grid, block, 0, s>>>(t, reinterpret_cast<const double2*>(pw), h, w_, y);
  VSR_CHECK_LAUNCH("vsr_downscale(cols)");
  const long total = (long)n * (h / r) * (w_ / r);
  bicubic_down_kernel<<<grid_for(total, 256), 256, 0, s>>>(y, n, h, w_, r, lr);
  VSR_CHECK_LAUNCH("vsr_downscale(bicubic)");
  return VSR_OK;
}
