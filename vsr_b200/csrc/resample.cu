// resample.cu — standalone bandwidth kernels: nn.PixelShuffle and its inverse on NCHW
// (drf_net.py:142; the nets themselves never launch it — the shuffle is a reinterpretation of
// the phase-blocked layout) and F.interpolate(bilinear / trilinear) forward + backward
// (srfb_net.py:47; trilinear has no reference call site).
#include <algorithm>

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

// ---- staged forward (up-scaling, the case of the nets: bilinear x r of srfb_net.py:47, trilinear) -------------------
// The gather kernel above issues 4 (8) scalar global loads per output and exposes their latency on every row.  Here a block
// stages ALL input rows its 32 output rows need - the tile's input segment, the two z planes already blended - in shared
// memory with one round of coalesced loads, then every thread forms 4 consecutive outputs per row from shared memory
// (x table read as 16-byte vectors, 16-byte stores): one global latency per block instead of one per row.
// Bytes moved = input + output.
constexpr int kUpStage = 6144;                   // staged input elements per block (24 KB)
__global__ void __launch_bounds__(256) upsample_linear_staged_kernel(const float* __restrict__ x, float* __restrict__ y, int d,
                                                                    int h, int w, int od, int oh, int ow, int ac) {
  __shared__ __align__(16) int xi0[kUpTile];
  __shared__ __align__(16) int xi1[kUpTile];
  __shared__ __align__(16) float xw1[kUpTile];
  __shared__ float raw[kUpStage];
  const int x0 = blockIdx.x * kUpTile;
  const int nx = min(kUpTile, ow - x0);
  for (int i = threadIdx.x; i < kUpTile; i += blockDim.x) {
    const LinCoord cx = lin_coord(min(x0 + i, ow - 1), w, ow, ac);
    xi0[i] = cx.i0; xi1[i] = cx.i1; xw1[i] = cx.w1;
  }
  __syncthreads();
  const int lo = xi0[0], ns = xi1[nx - 1] - lo + 1;            // input segment of the tile
  const int cz = blockIdx.z;                                    // c * od + oz
  const int c = cz / od, oz = cz - c * od;
  const bool three_d = !(d == 1 && od == 1);
  LinCoord czc;
  czc.i0 = czc.i1 = 0; czc.w0 = 1.f; czc.w1 = 0.f;
  if (three_d) czc = lin_coord(oz, d, od, ac);
  const float* p0 = x + ((size_t)c * d + czc.i0) * h * w + lo;
  const float* p1 = x + ((size_t)c * d + czc.i1) * h * w + lo;
  const int oy_first = blockIdx.y * kRowsPerBlock, oy_last = min(oy_first + kRowsPerBlock, oh) - 1;
  const int iy_lo = lin_coord(oy_first, h, oh, ac).i0, iy_hi = lin_coord(oy_last, h, oh, ac).i1;
  const int rows_in = iy_hi - iy_lo + 1;                        // rows_in * ns <= kUpStage: host check
  for (int i = threadIdx.x; i < rows_in * ns; i += blockDim.x) {
    const int rr = i / ns, xi = i - rr * ns;
    float t = __ldg(p0 + (size_t)(iy_lo + rr) * w + xi);
    if (three_d) t = czc.w0 * t + czc.w1 * __ldg(p1 + (size_t)(iy_lo + rr) * w + xi);
    raw[i] = t;
  }
  __syncthreads();
  const int nx4 = (nx + 3) >> 2;
  const bool vec = ((ow & 3) == 0);
  const int n_rows = oy_last - oy_first + 1;
  for (int i = threadIdx.x; i < n_rows * nx4; i += blockDim.x) {
    const int rr = i / nx4, i4 = (i - rr * nx4) * 4;
    const LinCoord cy = lin_coord(oy_first + rr, h, oh, ac);
    const float* r0 = raw + (cy.i0 - iy_lo) * ns - lo;
    const float* r1 = raw + (cy.i1 - iy_lo) * ns - lo;
    const int4 a0 = *reinterpret_cast<const int4*>(xi0 + i4);
    const int4 a1 = *reinterpret_cast<const int4*>(xi1 + i4);
    const float4 w1 = *reinterpret_cast<const float4*>(xw1 + i4);
    float v[4];
    v[0] = cy.w0 * ((1.f - w1.x) * r0[a0.x] + w1.x * r0[a1.x]) + cy.w1 * ((1.f - w1.x) * r1[a0.x] + w1.x * r1[a1.x]);
    v[1] = cy.w0 * ((1.f - w1.y) * r0[a0.y] + w1.y * r0[a1.y]) + cy.w1 * ((1.f - w1.y) * r1[a0.y] + w1.y * r1[a1.y]);
    v[2] = cy.w0 * ((1.f - w1.z) * r0[a0.z] + w1.z * r0[a1.z]) + cy.w1 * ((1.f - w1.z) * r1[a0.z] + w1.z * r1[a1.z]);
    v[3] = cy.w0 * ((1.f - w1.w) * r0[a0.w] + w1.w * r0[a1.w]) + cy.w1 * ((1.f - w1.w) * r1[a0.w] + w1.w * r1[a1.w]);
    float* yp = y + ((size_t)cz * oh + oy_first + rr) * ow + x0;
    if (vec && i4 + 3 < nx) {
      *reinterpret_cast<float4*>(yp + i4) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (i4 + q < nx) yp[i4 + q] = v[q];
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

// ---- staged backward: the exact transposed stencil (deterministic) --------------------------------------------------
// A block owns kBwdRows input rows x 256 input columns of one (plane, iz).  Every thread owns one column: the outputs
// ox in [xl, xh] whose stencil touches it and their weights are tabulated ONCE per block (shared memory, at most kBwdK
// per column - up-scaling by <= 8); then for every output row that touches the block's input rows the row segment of dy
// is staged in shared memory (coalesced), the thread forms the horizontal sum h = sum_k wx[k] * dy[oy][xl + k] and adds
// wy * h to the accumulators of the (at most two) input rows the output row interpolates between.  No lin_coord in
// the inner loop; bytes moved = dy + dx (+ the 2r halo rows per block, from L2).
constexpr int kBwdRows = 8, kBwdK = 18, kBwdRowsInFlight = 3, kBwdSeg = 2112;
__global__ void __launch_bounds__(256) upsample_linear_bwd_staged_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                                        int d, int h, int w, int od, int oh, int ow, int ac) {
  __shared__ float wxs[kBwdK][256];
  __shared__ float seg[kBwdRowsInFlight][kBwdSeg];
  __shared__ int blk_lo, blk_hi;
  const int tid = threadIdx.x;
  const int cz = blockIdx.z;                     // c * d + iz
  const int c = cz / d, iz = cz - c * d;
  const bool three_d = !(d == 1 && od == 1);
  const int ix = blockIdx.x * 256 + tid;
  const int iy0 = blockIdx.y * kBwdRows;
  const int nrows = min(kBwdRows, h - iy0);
  int xl = 0, xh = -1;
  if (ix < w) out_range(ix, w, ow, ac, &xl, &xh);
  // trim to the outputs that really touch ix and tabulate their weights
  int first = -1, nk = 0;
  for (int ox = xl; ox <= xh; ++ox) {
    const LinCoord cx = lin_coord(ox, w, ow, ac);
    const float wx = (cx.i0 == ix ? cx.w0 : 0.f) + (cx.i1 == ix ? cx.w1 : 0.f);
    if (wx != 0.f || (first >= 0 && nk < kBwdK)) {
      if (first < 0) first = ox;
      if (ox - first < kBwdK) {
        wxs[ox - first][tid] = wx;
        nk = ox - first + 1;
      }
    }
  }
  for (int k = nk; k < kBwdK; ++k) wxs[k][tid] = 0.f;
  if (first < 0) first = 0;
  if (tid == 0) {                                // output-column range of the block: columns are monotone in ix
    int l, hdummy, dummy, hh;
    out_range(blockIdx.x * 256, w, ow, ac, &l, &hdummy);
    out_range(min((int)blockIdx.x * 256 + 255, w - 1), w, ow, ac, &dummy, &hh);
    blk_lo = l;
    blk_hi = hh;
  }
  __syncthreads();
  const int slo = blk_lo, ns = blk_hi - blk_lo + 1;            // ns <= kBwdSeg: host check
  int nk_blk = nk;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nk_blk = max(nk_blk, __shfl_xor_sync(0xffffffffu, nk_blk, o));
  int yl, yh, ydummy, zl = 0, zh = 0;
  out_range(iy0, h, oh, ac, &yl, &ydummy);
  out_range(iy0 + nrows - 1, h, oh, ac, &ydummy, &yh);
  if (three_d) out_range(iz, d, od, ac, &zl, &zh);
  float acc[kBwdRows];
#pragma unroll
  for (int r = 0; r < kBwdRows; ++r) acc[r] = 0.f;
  const float* g = dy + (size_t)c * od * oh * ow;
  for (int oz = zl; oz <= zh; ++oz) {
    float wz = 1.f;
    if (three_d) {
      const LinCoord cc = lin_coord(oz, d, od, ac);
      wz = (cc.i0 == iz ? cc.w0 : 0.f) + (cc.i1 == iz ? cc.w1 : 0.f);
      if (wz == 0.f) continue;                   // (uniform over the block)
    }
    for (int oy0 = yl; oy0 <= yh; oy0 += kBwdRowsInFlight) {
      const int nr = min(kBwdRowsInFlight, yh - oy0 + 1);
      __syncthreads();
      for (int i = tid; i < nr * ns; i += 256) {
        const int rr = i / ns, xi = i - rr * ns;
        seg[rr][xi] = __ldg(g + ((size_t)oz * oh + oy0 + rr) * ow + slo + xi);
      }
      __syncthreads();
      for (int rr = 0; rr < nr; ++rr) {
        const LinCoord cy = lin_coord(oy0 + rr, h, oh, ac);
        const float* sr = &seg[rr][first - slo];
        float hsum = 0.f;
        for (int k = 0; k < nk_blk; ++k) hsum = fmaf(wxs[k][tid], (k < nk) ? sr[k] : 0.f, hsum);
        const int r0 = cy.i0 - iy0, r1 = cy.i1 - iy0;
        const float a0 = wz * cy.w0 * hsum, a1 = wz * cy.w1 * hsum;
#pragma unroll
        for (int r = 0; r < kBwdRows; ++r) acc[r] += (r == r0 ? a0 : 0.f) + (r == r1 ? a1 : 0.f);
      }
    }
  }
  if (ix < w)
#pragma unroll
    for (int r = 0; r < kBwdRows; ++r)
      if (r < nrows) dx[((size_t)cz * h + iy0 + r) * w + ix] = acc[r];
}

// ---- tile-staged backward (up-scaling by at most 4 per axis: the nets' bilinear x2..x4, trilinear x2) ----------------
// The row-staged kernel above still exposes one global latency per group of output rows.  Here a block owns 64 input
// columns x 8 input rows, stages EVERY dy row that touches them (the block's output-column segment) in shared memory with
// one round of coalesced loads per z plane, and then only reads shared memory: a thread owns one column and two rows,
// keeps its column's stencil weights in registers and, for every staged output row that interpolates from one of its
// rows, adds wy * sum_k wx[k] * dy[oy][first + k].
constexpr int kB2X = 64, kB2Y = 8, kB2K = 12, kB2Rows = 48, kB2Chunk = 16, kB2Seg = 288;
__global__ void __launch_bounds__(256) upsample_linear_bwd_tile_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                                      int d, int h, int w, int od, int oh, int ow, int ac) {
  // separable: horizontal sums hs[output row][input column] first (every staged dy element is used once), then the
  // vertical sums over the output rows that touch an input row - O(2r + 2r) multiply-adds per input element instead of
  // O(2r * 2r)
  __shared__ float stage[kB2Chunk][kB2Seg];
  __shared__ float hs[kB2Rows][kB2X];
  __shared__ float wxs[kB2K][kB2X];
  __shared__ int yi0[kB2Rows], yi1[kB2Rows];
  __shared__ float yw0[kB2Rows], yw1[kB2Rows];
  const int tid = threadIdx.x, col = tid & (kB2X - 1), rg = tid >> 6, lane = tid & 31, warp = tid >> 5;
  const int cz = blockIdx.z;                     // c * d + iz
  const int c = cz / d, iz = cz - c * d;
  const bool three_d = !(d == 1 && od == 1);
  const int ix0 = blockIdx.x * kB2X, iy0 = blockIdx.y * kB2Y;
  const int ncols = min(kB2X, w - ix0), nrows = min(kB2Y, h - iy0);
  int slo, shi, yl, yh, tmp;
  out_range(ix0, w, ow, ac, &slo, &tmp);
  out_range(ix0 + ncols - 1, w, ow, ac, &tmp, &shi);
  out_range(iy0, h, oh, ac, &yl, &tmp);
  out_range(iy0 + nrows - 1, h, oh, ac, &tmp, &yh);
  const int ns = shi - slo + 1, nro = yh - yl + 1;              // ns <= kB2Seg, nro <= kB2Rows: host check
  if (ns > kB2Seg || nro > kB2Rows) __trap();                   // (never silently out of bounds)
  // this column's candidate outputs [xl, xl + kB2K) and their weights (zero where the stencil does not touch it); the
  // four row groups of a column each tabulate a quarter of them
  const int ix = ix0 + col;
  int xl = slo, xh = slo - 1;
  if (ix < w) out_range(ix, w, ow, ac, &xl, &xh);
  for (int k = rg; k < kB2K; k += 4) {
    const int ox = xl + k;
    float wx = 0.f;
    if (ix < w && ox <= xh) {
      const LinCoord cx = lin_coord(ox, w, ow, ac);
      wx = (cx.i0 == ix ? cx.w0 : 0.f) + (cx.i1 == ix ? cx.w1 : 0.f);
    }
    wxs[k][col] = wx;
  }
  if (tid < nro) {
    const LinCoord cy = lin_coord(yl + tid, h, oh, ac);
    yi0[tid] = cy.i0; yi1[tid] = cy.i1; yw0[tid] = cy.w0; yw1[tid] = cy.w1;
  }
  for (int i = tid; i < kB2Rows * kB2X; i += 256) (&hs[0][0])[i] = 0.f;
  __syncthreads();
  float wr[kB2K];
#pragma unroll
  for (int k = 0; k < kB2K; ++k) wr[k] = wxs[k][col];
  const int off = xl - slo;
  const int nk = min(kB2K, ns - off);                           // staged elements to the right of the column's first output
  int zl = 0, zh = 0;
  if (three_d) out_range(iz, d, od, ac, &zl, &zh);
  const float* g = dy + (size_t)c * od * oh * ow;
  for (int oz = zl; oz <= zh; ++oz) {
    float wz = 1.f;
    if (three_d) {
      const LinCoord cc = lin_coord(oz, d, od, ac);
      wz = (cc.i0 == iz ? cc.w0 : 0.f) + (cc.i1 == iz ? cc.w1 : 0.f);
      if (wz == 0.f) continue;                   // (uniform over the block)
    }
    for (int r0 = 0; r0 < nro; r0 += kB2Chunk) {
      const int nr = min(kB2Chunk, nro - r0);
      __syncthreads();
      for (int rr = warp; rr < nr; rr += 8) {    // a warp stages whole rows: no index division, coalesced
        const float* src = g + ((size_t)oz * oh + yl + r0 + rr) * ow + slo;
        for (int xi = lane; xi < ns; xi += 32) stage[rr][xi] = __ldg(src + xi);
      }
      __syncthreads();
      for (int rr = rg; rr < nr; rr += 4) {      // horizontal sums of this chunk's rows
        const float* sr = &stage[rr][off];
        float hsum = 0.f;
#pragma unroll
        for (int k = 0; k < kB2K; ++k) hsum = fmaf(wr[k], k < nk ? sr[k] : 0.f, hsum);
        hs[r0 + rr][col] += wz * hsum;
      }
    }
  }
  __syncthreads();
  // vertical sums: this thread's two input rows
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int ri = iy0 + rg + 4 * e;
    if (ri < iy0 + nrows && col < ncols) {
      int lo, hi;
      out_range(ri, h, oh, ac, &lo, &hi);
      lo = max(lo - yl, 0);
      hi = min(hi - yl, nro - 1);
      float acc = 0.f;
      for (int rr = lo; rr <= hi; ++rr) {
        const float wy = (yi0[rr] == ri ? yw0[rr] : 0.f) + (yi1[rr] == ri ? yw1[rr] : 0.f);
        acc = fmaf(wy, hs[rr][col], acc);
      }
      dx[((size_t)cz * h + ri) * w + ix0 + col] = acc;
    }
  }
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

namespace vsr {
// resample_int.cu: integer power-of-two ratios with align_corners = False; false when the shape is not covered
bool upsample_int_fwd(const float* x, float* y, int nc, int d, int h, int w, int od, int oh, int ow, int ac, cudaStream_t s);
bool upsample_int_bwd(const float* dy, float* dx, int nc, int d, int h, int w, int od, int oh, int ow, int ac, cudaStream_t s);
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
  if (tunables().up_generic != 1 && upsample_int_fwd(x, y, nc, d, h, w_, od, oh, ow, align_corners, static_cast<cudaStream_t>(stream))) {
    VSR_CHECK_LAUNCH("vsr_upsample_linear(integer ratio)");
    return VSR_OK;
  }
  VSR_CHECK_SUPPORTED((long)nc * od <= 65535 && oh <= 65535, "vsr_upsample_linear: nc*od and oh must be <= 65535");
  dim3 grid((ow + kUpTile - 1) / kUpTile, (oh + kRowsPerBlock - 1) / kRowsPerBlock, nc * od);
  // staged kernel when the input rows of a block (32 output rows x 1024 output columns) fit its staging buffer - any
  // up-scaling by 2 or more on both axes; gather kernel otherwise
  const long seg_x = (long)std::min(ow, kUpTile) * w_ / ow + 3, seg_y = (long)kRowsPerBlock * h / oh + 3;
  const bool staged = ow >= w_ && oh >= h && seg_x * seg_y <= kUpStage;
  if (staged)
    upsample_linear_staged_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, d, h, w_, od, oh, ow, align_corners);
  else
    upsample_linear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, d, h, w_, od, oh, ow, align_corners);
  VSR_CHECK_LAUNCH("vsr_upsample_linear");
  return VSR_OK;
}

extern "C" int vsr_upsample_linear_bwd(const float* dy, float* dx, int32_t nc, int32_t d, int32_t h, int32_t w_,
                                       int32_t od, int32_t oh, int32_t ow, int align_corners, void* stream) {
  VSR_CHECK_ARG(dy && dx && nc > 0 && d > 0 && h > 0 && w_ > 0 && od > 0 && oh > 0 && ow > 0, "vsr_upsample_linear_bwd: bad arguments");
  if (tunables().up_generic != 1 && upsample_int_bwd(dy, dx, nc, d, h, w_, od, oh, ow, align_corners, static_cast<cudaStream_t>(stream))) {
    VSR_CHECK_LAUNCH("vsr_upsample_linear_bwd(integer ratio)");
    return VSR_OK;
  }
  VSR_CHECK_SUPPORTED((long)nc * d <= 65535 && h <= 65535, "vsr_upsample_linear_bwd: nc*d and h must be <= 65535");
  // staged transposed stencil for up-scaling by at most 8 per axis (at most kBwdK outputs touch an input column and the
  // dy segment of 256 input columns fits kBwdSeg); the generic gather kernel otherwise
  auto ratio_ok = [&](int in, int out) {       // outputs per input along an axis: 1 <= ratio <= 8 (both align modes)
    if (out < in) return false;
    if (align_corners) return in == 1 ? out <= 8 : (long)(out - 1) <= 8l * (in - 1);
    return (long)out <= 8l * in;
  };
  const bool staged = ratio_ok(w_, ow) && oh >= h && od >= d;
  // outputs per input along an axis (rounded up, both align modes)
  auto span = [&](int in, int out) -> long {
    if (align_corners) return in == 1 ? out : ((long)(out - 1) + (in - 2)) / (in - 1);
    return ((long)out + in - 1) / in;
  };
  const long sx = span(w_, ow), sy = span(h, oh);
  const bool tiled = staged && 2 * sx + 4 <= kB2K && (kB2Y * sy + 2 * sy + 2) <= kB2Rows &&
                     (kB2X * sx + 2 * sx + 2) <= kB2Seg;
  if (tiled) {
    dim3 grid((w_ + kB2X - 1) / kB2X, (h + kB2Y - 1) / kB2Y, nc * d);
    upsample_linear_bwd_tile_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, d, h, w_, od, oh, ow, align_corners);
  } else if (staged) {
    dim3 grid((w_ + 255) / 256, (h + kBwdRows - 1) / kBwdRows, nc * d);
    upsample_linear_bwd_staged_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, d, h, w_, od, oh, ow, align_corners);
  } else {
    dim3 grid((w_ + 255) / 256, (h + kRowsPerBlock - 1) / kRowsPerBlock, nc * d);
    upsample_linear_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, d, h, w_, od, oh, ow, align_corners);
  }
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
