/*
 * vsr_b200.h — C-ABI of libvsr_sm100.so: the sm_100a kernels under the drop-in SR nets.
 *
 * The reference (yangsenwxy/VSR) has no FFI for this path: its nets call torch.nn modules
 * (src/model/nets/drf_net.py:55-106,141-147), which dispatch to ATen/cuDNN.  Every entry
 * point below states which reference call site(s) it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *  - every device buffer (including workspaces) is owned by the caller; the library allocates
 *    nothing on the device and keeps only host-side caches (TMA descriptors).
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *    synchronises, and is CUDA-graph capturable.
 *  - return value: 0 on success, a negative VsrStatus otherwise; vsr_last_error() gives text.
 *    Nothing throws or exits across the ABI.
 *
 * Feature-map layout ("pixel-major"): dense [n][h][w][c], c contiguous.  A high-resolution map
 * of upscale r is kept *phase-blocked*: the r*r sub-pixels of low-resolution pixel (y,x) are
 * stored as r*r consecutive channel groups of that pixel, i.e. as a [n][h][w][r*r*C] map, so
 * that nn.PixelShuffle (drf_net.py:142,146) is a reinterpretation, the transposed convolution
 * (drf_net.py:81,93) is a dense GEMM into channel slices, and the strided convolution
 * (drf_net.py:86,100) reads channel slices of neighbouring low-resolution pixels.
 */
#ifndef VSR_B200_H_
#define VSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSR_ABI_VERSION 4
#define VSR_MAX_SRCS 8

typedef enum VsrStatus {
  VSR_OK = 0,
  VSR_ERR_BAD_ARG = -1,
  VSR_ERR_UNSUPPORTED = -2,
  VSR_ERR_CUDA = -3,
  VSR_ERR_DRIVER = -4
} VsrStatus;

/* VSR_BF16X2 (tap-GEMM only): the strict mode on tensor cores.  An fp32 map travels as two bf16 planes
 * [2][n][h][w][c] (x = plane 0 + plane 1 to 16 significant bits, vsr_split_planes); srcs[i].ptr points at plane 0 and
 * srcs[i].n stays the logical n; bit 3 of a tap's source index selects plane 1; `out` receives the raw fp32
 * accumulators (epi must be 0: vsr_tap_epilogue applies the epilogue in fp32 afterwards); weights are bf16 slabs as for
 * VSR_BF16.  The caller expresses x*w = xh*wh + xl*wh + xh*wl as three taps with slabs [wh | wh | wl] (vsr_gather_split). */
typedef enum VsrDType { VSR_F32 = 0, VSR_BF16 = 1, VSR_BF16X2 = 2 } VsrDType;

/* epilogue flags of the tap-GEMM (applied in this order) */
enum {
  VSR_EPI_BIAS = 1,       /* v += bias[o]                                              */
  VSR_EPI_RES_PRE = 2,    /* v += residual[pix][o]        (before the activation)      */
  VSR_EPI_PRELU = 4,      /* v = v > 0 ? v : a*v          (nn.PReLU, scalar a)         */
  VSR_EPI_RELU = 8,       /* v = max(v, 0)                                             */
  VSR_EPI_PRELU_BWD = 16, /* y=aux_y: v = y>0 ? v : a*v ; slope grad += v_in * y/a     */
  VSR_EPI_RELU_BWD = 32,  /* y=aux_y: v = y>0 ? v : 0                                  */
  VSR_EPI_OUT2 = 64,      /* out2[pix][o] = v + res2[pix][o]   (after the activation)  */
  VSR_EPI_SCALE = 128,    /* v *= out_scale   (after bias, before the residual)        */
  VSR_EPI_OUT2_SUB = 256  /* with OUT2: out2[pix][o] = v - res2[pix][o]  (rbp_net.py:274,284: l0 - x, h0 - x)  */
};

typedef struct VsrTensor4 {
  void* ptr;        /* device pointer, 16-byte aligned */
  int32_t n, h, w, c;
} VsrTensor4;

/*
 * Tap-GEMM: for every output pixel (n,y,x) of `out` and every group g
 *   out[n,y,x, o0_g + j] = epi( sum_{t in taps(g)} sum_{k<kc}
 *                               srcs[t.src][n, y+t.dy, x+t.dx, t.c0 + k] * w[t][j][k] ),  j < nt
 * with zero for out-of-range source pixels.  All srcs and out share n,h,w.
 * Replaces, with the right tap tables: nn.Conv2d 1x1 on torch.cat (drf_net.py:57,65,90,97,105,
 * 119-131: the cat is never materialised, each cat operand is one tap), nn.Conv2d 3x3
 * (drf_net.py:141,144,147), nn.ConvTranspose2d k,s=r,p=2 (drf_net.py:81,93), nn.Conv2d k,s=r,p=2
 * (drf_net.py:86,100), and the data-gradient of each of them.
 *
 * group_tab: int32[n_groups][4] = {o0, tap_begin, n_taps, reserved}
 * tap_tab:   int32[n_taps_total][4] = {src, dy, dx, c0}
 * w:         n_taps_total slabs of [nt][kc]; VSR_F32: plain row-major fp32;
 *            VSR_BF16: bf16, each slab in the 128-byte-swizzled K-major shared-memory image
 *            (16-byte chunk q of row j stored at chunk q ^ (j & 7)); see vsr_slab_index().
 */
typedef struct VsrTapGemmDesc {
  int32_t dtype; /* VsrDType of srcs / out / residual / aux / w */
  int32_t kc;    /* channels per tap                (BF16: 64) */
  int32_t nt;    /* output channels per group       (BF16: multiple of 16, <= 256) */
  int32_t n_srcs;
  VsrTensor4 srcs[VSR_MAX_SRCS];
  VsrTensor4 out;
  int32_t n_groups;
  int32_t n_taps_total;
  int32_t max_group_taps;   /* host copy of max over groups of n_taps (needed by vsr_tapgemm_wgrad) */
  const int32_t* group_tab; /* device */
  const int32_t* tap_tab;   /* device */
  const void* w;            /* device */
  const float* bias;        /* device, [out.c] fp32, or NULL */
  int32_t epi;              /* VSR_EPI_* flags */
  float out_scale;
  const float* slope;    /* device scalar (PRELU / PRELU_BWD) */
  const void* residual;  /* [n,h,w,out.c] dtype (RES_PRE) */
  const void* aux_y;     /* [n,h,w,out.c] dtype (PRELU_BWD / RELU_BWD) */
  void* out2;            /* [n,h,w,out.c] dtype (OUT2) */
  const void* res2;      /* [n,h,w,out.c] dtype (OUT2) */
  float* slope_partials; /* device, >= vsr_partials_len() floats: per-CTA partial slope grads */
  const int32_t* tap_tab_host; /* optional HOST copy of tap_tab (or NULL): lets the bf16 kernels plan shared loads
                                  for taps that differ only by a row shift; results do not depend on it */
  const int32_t* group_tab_host; /* optional HOST copy of group_tab (or NULL): shared loads for tables with several groups */
} VsrTapGemmDesc;

int vsr_abi_version(void);
const char* vsr_last_error(void);
/* number of floats a slope_partials row must hold (upper bound of any grid) */
int vsr_partials_len(void);
/* re-read the tuning overrides from the environment (VSR_TC_*, VSR_WG_*, VSR_PDL: attribution tools and tests;
 * not part of the reference-facing contract).  They are read once at the first launch otherwise. */
void vsr_reload_tunables(void);

/* forward / data-gradient tap-GEMM (see above) */
int vsr_tapgemm(const VsrTapGemmDesc* d, void* stream);

/*
 * Weight-gradient of a tap-GEMM: dw[t][j][k] (+)= sum_{n,y,x} dz[n,y,x,o0_g+j] *
 *   srcs[t.src][n,y+t.dy,x+t.dx,t.c0+k]; dw is fp32 [n_taps_total][nt][kc] plain row-major
 * (never swizzled).  `d->out.ptr` is dz; d->w, bias, epilogue fields are ignored.
 * workspace: >= vsr_tapgemm_wgrad_workspace(d) bytes.  accumulate!=0 adds to dw.
 * Replaces the weight half of aten.convolution_backward for the call sites listed above;
 * the summation order is fixed (deterministic, main.py:32 cudnn.deterministic).
 */
size_t vsr_tapgemm_wgrad_workspace(const VsrTapGemmDesc* d);
int vsr_tapgemm_wgrad(const VsrTapGemmDesc* d, float* dw, int accumulate, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Weight gradient with the bias gradient fused where the kernel supports it:
 * db[q] (+)= sum over pixels and channels c = q (mod db_period) of dz.  Returns 1 if db was produced,
 * 0 if only dw was (the caller then runs vsr_colsum), negative on error. */
int vsr_tapgemm_wgrad_bias(const VsrTapGemmDesc* d, float* dw, float* db, int32_t db_period, int accumulate,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Deferred reduction of the weight / bias gradient of one layer over several calls with identical
 * shapes (the T frames of a recurrent net): call k writes its per-split partial sums into slice k of
 * `workspace` (>= n_slices * vsr_tapgemm_wgrad_workspace(d) bytes) and returns 1, or returns 0 if the
 * shape is not supported by the tensor-core kernel (then use vsr_tapgemm_wgrad); _finish reduces the
 * first `used_slices` slices in a fixed order into dw / db.  `d->out` is dz as in vsr_tapgemm_wgrad. */
int vsr_tapgemm_wgrad_partial(const VsrTapGemmDesc* d, int32_t db_period, int32_t slice, int32_t n_slices,
                              void* workspace, size_t workspace_bytes, void* stream);
int vsr_tapgemm_wgrad_finish(const VsrTapGemmDesc* d, float* dw, float* db, int32_t db_period, int accumulate,
                             int32_t used_slices, int32_t n_slices, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- strict mode on tensor cores (precision='bf16x3'): error-compensated bf16 pairs, csrc/split.cu ----
 * Same call sites as vsr_tapgemm (drf_net.py:55-106,141-147) with fp32 maps on the caller's side:
 *   vsr_split_planes(x)                         -> planes the VSR_BF16X2 tap-GEMM (and, plane by plane as VSR_BF16 maps,
 *                                                  vsr_tapgemm_wgrad_partial) reads;  numel % 8 == 0
 *   vsr_tapgemm(dtype = VSR_BF16X2, epi = 0)    -> raw fp32 accumulators in `out`
 *   vsr_tap_epilogue(out, ...)                  -> the VSR_EPI_* flags of vsr_tapgemm applied in place, in fp32, in the
 *                                                  order listed above ([rows][c] maps, c % 4 == 0; bias[c]; PRELU_BWD
 *                                                  writes one partial per block into slope_partials); `planes` /
 *                                                  `planes2` (optional): the bf16 planes of `out` / `out2`, written in
 *                                                  the same pass (saves the next consumer's vsr_split_planes a read)
 *   vsr_gather_split(src, idx, dst, n)          -> dst[i] = bf16 high part of src[idx[i]], or its low part
 *                                                  bf16(src - high) when bit 30 of idx[i] is set; 0 for idx[i] < 0 */
int vsr_split_planes(const float* x, void* planes, int64_t numel, void* stream);
int vsr_tap_epilogue(float* out, int64_t rows, int32_t c, const float* bias, int32_t epi, float out_scale,
                     const float* slope, const float* residual, const float* aux_y, float* out2, const float* res2,
                     float* slope_partials, void* planes, void* planes2, void* stream);
int vsr_gather_split(const float* src, const int32_t* idx, void* dst, int64_t n, void* stream);

/* Weight (+ bias) gradients of SEVERAL 1x1 convolutions that read the same list of 64-channel maps, in one pass over
 * the maps (csrc/wgrad_shared.cu): layer g multiplies its gradient map dzs[g] with the first ntaps[g] sources,
 *   dw[g][t][j][k] (+)= sum_pix dzs[g][pix][j] * srcs[t][pix][k],  t < ntaps[g];   db[g][j] (+)= sum_pix dzs[g][pix][j]
 * (dw[g]: fp32 [ntaps[g]][64][64] plain row-major = the packed slab order of the layer; db[g] may be NULL).  Every source
 * tile is loaded once per launch instead of once per layer: the dense connections of the feedback block
 * (f_block.{up,down}_blocks[g].conv1 on torch.cat(lr_list / hr_list), drf_net.py:89-105) make the separate weight gradients
 * read map j once for every g >= j.  All maps: VSR_BF16 [n][h][w][64] on one pixel grid; sum over g of ceil(ntaps[g] / 2)
 * <= 8 (tensor-memory accumulators), otherwise VSR_ERR_UNSUPPORTED: split the layer set.  Fixed summation order. */
#define VSR_WS_MAX_SRCS 8
#define VSR_WS_MAX_DZ 4
typedef struct VsrWgradSharedDesc {
  int32_t n_srcs, n_dz;
  VsrTensor4 srcs[VSR_WS_MAX_SRCS];
  VsrTensor4 dzs[VSR_WS_MAX_DZ];
  int32_t ntaps[VSR_WS_MAX_DZ];
  float* dw[VSR_WS_MAX_DZ];
  float* db[VSR_WS_MAX_DZ];
} VsrWgradSharedDesc;
size_t vsr_wgrad_shared_workspace(const VsrWgradSharedDesc* d);
int vsr_wgrad_shared(const VsrWgradSharedDesc* d, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* colsum: db[c] (+)= sum over pixels x[pix][c]  (bias gradient; fixed order).
 * workspace >= vsr_colsum_workspace(rows, c) bytes. */
size_t vsr_colsum_workspace(int64_t rows, int32_t c);
int vsr_colsum(const void* x, int32_t dtype, int64_t rows, int32_t c, float* db, int accumulate,
               void* workspace, size_t workspace_bytes, void* stream);

/*
 * First layer (in_block.conv1, drf_net.py:55-56): 3x3 pad-1 convolution of a Cin-channel image
 * (NCHW fp32, Cin small) to `cout` channels + bias + PReLU, written pixel-major in `dtype`.
 * w: fp32 [cout][cin][3][3] (reference layout, used as is).
 */
int vsr_conv3x3_first(const float* x, int32_t n, int32_t cin, int32_t h, int32_t w_, const float* w,
                      const float* bias, const float* slope, void* y, int32_t dtype, int32_t cout,
                      void* stream);
/* its backward: dz is [n,h,w,cout] (already multiplied by PReLU'); produces dw (fp32
 * [cout][cin][3][3]) and db; dx is not needed (network input).  deterministic two-pass. */
size_t vsr_conv3x3_first_bwd_workspace(int32_t n, int32_t cin, int32_t h, int32_t w_, int32_t cout);
int vsr_conv3x3_first_bwd(const float* x, int32_t n, int32_t cin, int32_t h, int32_t w_,
                          const void* dz, int32_t dtype, int32_t cout, float* dw, float* db,
                          int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Last layer (out_block.conv{k}, drf_net.py:144,147): 3x3 pad-1 convolution from C channels to
 * cout (small) on a phase-blocked high-resolution map x [n,h,w,r*r*C] (phase order given by
 * `phase_yx`, a HOST int32[r*r][2] = (py,px) of each phase slot), producing NCHW fp32 [n,cout,r*h,r*w].
 * w: fp32 [cout][C][3][3]; bias fp32 [cout].
 */
int vsr_conv3x3_last(const void* x, int32_t dtype, int32_t n, int32_t h, int32_t w_, int32_t r,
                     int32_t c, const int32_t* phase_yx, const float* w, const float* bias,
                     float* y, int32_t cout, void* stream);
/* backward: dy NCHW fp32 [n,cout,rh,rw] -> dx phase-blocked [n,h,w,r*r*c] in dtype, and dw/db. */
size_t vsr_conv3x3_last_bwd_workspace(int32_t n, int32_t h, int32_t w_, int32_t r, int32_t c,
                                      int32_t cout);
int vsr_conv3x3_last_bwd(const void* x, int32_t dtype, int32_t n, int32_t h, int32_t w_, int32_t r,
                         int32_t c, const int32_t* phase_yx, const float* w, const float* dy,
                         int32_t cout, void* dx, float* dw, float* db, int accumulate,
                         void* workspace, size_t workspace_bytes, void* stream);

/* elementwise PReLU / ReLU backward on a pixel-major map: dz = y>0 ? dy : a*dy, partial slope
 * grads to slope_partials (nn.PReLU backward, drf_net.py:56-106). a == NULL means ReLU. */
int vsr_act_bwd(const void* dy, const void* y, void* dz, int32_t dtype, int64_t numel,
                const float* slope, float* slope_partials, void* stream);

/* out[i] = a[i] + b[i] (pixel-major maps, dtype) — global skip drf_net.py:46 when not fused */
int vsr_add(const void* a, const void* b, void* out, int32_t dtype, int64_t numel, void* stream);

/* x[i] *= alpha (fp32) — res_scale of the EDSR residual blocks applied to packed gradients
 * (edsr_net.py:50-52). */
int vsr_scale(float* x, int64_t n, float alpha, void* stream);

/* dst[row_dst[r]] += sum(partials[r][0..len))  for r < rows; fixed order. */
int vsr_reduce_partials(const float* partials, int32_t rows, int32_t len, const int32_t* row_dst,
                        float* dst, void* stream);

/* gather: dst[i] = idx[i] >= 0 ? cast(src[idx[i]]) : 0, i < n.  Packs reference-layout fp32
 * parameters into tap slabs (dst dtype) and un-packs slab gradients (fp32 -> fp32). */
int vsr_gather(const float* src, const int32_t* idx, void* dst, int32_t dst_dtype, int64_t n,
               void* stream);
/* out[i] = alpha * a[i] + beta * b[i]  (fp32 or bf16 maps; b may be NULL when beta == 0; out may alias a or b):
 * the residual sums / differences between feature maps of rbp_net.py:84-87,274-275,284-285 in the backward pass */
int vsr_axpby(const void* a, const void* b, void* out, int32_t dtype, int64_t numel, float alpha, float beta, void* stream);

/* dst[i] += src[idx[i]] for idx[i] >= 0 (fp32) */
int vsr_gather_add(const float* src, const int32_t* idx, float* dst, int64_t n, void* stream);

/* index of element (row j, k) inside a bf16 [nt][64] slab in the swizzled image */
int64_t vsr_slab_index(int32_t j, int32_t k);

/*
 * Fused loss forward+backward over one frame (torch.nn.L1Loss / MSELoss via main.py:60-63,
 * CharbonnierLoss losses.py:23-34, HuberLoss losses.py:5-20), mean reduction:
 *   partial sums -> loss_partials[vsr_partials_len()], grad[i] = dloss/dout[i] * grad_scale.
 * kind: 0 L1, 1 MSE, 2 Charbonnier(param=epsilon), 3 Huber(param=delta). grad may be NULL.
 */
int vsr_loss_fwd_bwd(const float* out, const float* target, int64_t numel, int32_t kind, float param,
                     float grad_scale, float* loss_partials, float* grad, void* stream);
/*
 * The same over `n_segments` equally sized, consecutive segments of `numel` elements each in ONE launch - the T
 * frames of a step, whose per-frame means the trainer averages (acdc_vsr_trainer.py:74-88) and the predictor logs
 * per frame (acdc_vsr_predictor.py:119-133): segment s writes row s of loss_partials
 * ([n_segments][vsr_partials_len()]).  16-byte vector path only when every segment of every tensor is 16-byte
 * aligned; any other shape / slice takes the scalar loop (never a misaligned access).
 */
int vsr_loss_fwd_bwd_seg(const float* out, const float* target, int64_t numel, int32_t n_segments, int32_t kind,
                         float param, float grad_scale, float* loss_partials, float* grad, void* stream);

/*
 * Fused denormalize + PSNR (src/utils.py:1-20 + metrics.py:20-36): per sample
 * mse over (x*std+mean).round().clamp(0,255) of both inputs, psnr = 10 log10(max^2/(mse+1e-10)).
 * psnr_out: [n] fp32.  std<=0 disables the denormalisation (inputs used as they are).
 * workspace (both metrics): >= vsr_metric_workspace(n, per_sample) bytes; partial sums, fixed-order reduce.
 */
size_t vsr_metric_workspace(int32_t n, int64_t per_sample);
int vsr_psnr(const float* out, const float* target, int32_t n, int64_t per_sample, float mean,
             float std, float max_value, float* psnr_out, void* workspace, size_t workspace_bytes,
             void* stream);

/*
 * Fused denormalize + SSIM (metrics.py:51-113, dim=2, channels=1): separable 11-tap window
 * (weights given by the caller, fp32[11]; the reference's window is exp(-((i-5)/(2 sigma))^2)
 * normalised, metrics.py:74-77), valid convolution, per-sample mean of the SSIM map.
 * ssim_out: [n] fp32.  imgs are [n][h][w] fp32.
 */
int vsr_ssim(const float* out, const float* target, int32_t n, int32_t h, int32_t w_,
             const float* win11, float mean, float std, float c1, float c2, float* ssim_out,
             void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same for volumes (metrics.py:51-113 with dim=3: F.conv3d with the 11x11x11 product window): three
 * separable passes over the five moment maps, valid convolution, per-sample mean.  vols are [n][d][h][w].
 * workspace >= vsr_ssim3d_workspace(n, d, h, w) bytes.
 */
size_t vsr_ssim3d_workspace(int32_t n, int32_t d, int32_t h, int32_t w_);
int vsr_ssim3d(const float* out, const float* target, int32_t n, int32_t d, int32_t h, int32_t w_,
               const float* win11, float mean, float std, float c1, float c2, float* ssim_out,
               void* workspace, size_t workspace_bytes, void* stream);

/* nn.PixelShuffle / its inverse on NCHW fp32 (standalone, for the kernel sweep; the nets never
 * launch it).  x [n, c*r*r, h, w] -> y [n, c, h*r, w*r]  (drf_net.py:142). */
int vsr_pixel_shuffle(const float* x, float* y, int32_t n, int32_t c, int32_t h, int32_t w_,
                      int32_t r, int inverse, void* stream);

/* F.interpolate(mode='bilinear'|'trilinear') forward on NCHW / NCDHW fp32 (srfb_net.py:47;
 * trilinear has no reference call site, torch is the oracle).  d==1 -> bilinear. */
int vsr_upsample_linear(const float* x, float* y, int32_t nc, int32_t d, int32_t h, int32_t w_,
                        int32_t od, int32_t oh, int32_t ow, int align_corners, void* stream);
int vsr_upsample_linear_bwd(const float* dy, float* dx, int32_t nc, int32_t d, int32_t h,
                            int32_t w_, int32_t od, int32_t oh, int32_t ow, int align_corners,
                            void* stream);

/* Fused Adam over a flat fp32 parameter bucket (torch.optim.Adam semantics, main.py:73-74):
 * grads are multiplied by grad_scale first (1/world_size after the NCCL sum). */
int vsr_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                  void* stream);

/* Same update with the hyper-parameters read from device memory, so that a captured CUDA graph of
 * the training step can be replayed while lr / step change:
 * hyper = float[7] {lr, beta1, beta2, eps, weight_decay, step (>=1), grad_scale}. */
int vsr_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                      void* stream);

/* cast fp32 <-> bf16 / layout helpers */
int vsr_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n,
             void* stream);

/*
 * ---- Conv3d network (DUFNet, duf_net.py:21-214): bandwidth-bound parts -----------------------------------
 * The dense concatenation (duf_net.py:123-128) is ONE pixel-major map [frames*n*h*w][ld] (time-major: all
 * pixels of frame 0, then frame 1, ...), so torch.cat is never executed: each layer reads the channel
 * prefix [0, c) and its 3x3x3 convolution (a tap-GEMM whose sources are frame-shifted views of one map)
 * appends a channel slice.  All windows below are [c0, c0+c) of rows with stride ld (elements); c0, c and
 * ld must be multiples of 16 bytes.
 */

/* dst[row][c0_dst + j] = src[row][c0_src + j], j < c  (torch.stack of the per-frame head features,
 * duf_net.py:57-61, into the concat map; and back for the head's weight gradient) */
int vsr_copy_window(const void* src, int32_t ld_src, int32_t c0_src, void* dst, int32_t ld_dst,
                    int32_t c0_dst, int32_t c, int64_t rows, int32_t dtype, void* stream);

/* nn.BatchNorm3d batch statistics (duf_net.py:114,198,201,207,210), kept PER FRAME so that the temporal
 * crop concat[:, :, 1:-1] (duf_net.py:126) re-uses them:
 *   stats[f][0][s0 + j] = sum over the rows of frame f of x[row][c0 + j],  stats[f][1][s0 + j] = sum of squares
 * (fp64, [frames][2][ld_stats]).  workspace >= vsr_bn_stats_workspace(frames, rows_per_frame, c). */
size_t vsr_bn_stats_workspace(int32_t frames, int64_t rows_per_frame, int32_t c);
int vsr_bn_stats(const void* x, int32_t dtype, int32_t ldx, int32_t c0, int32_t c, int32_t frames,
                 int64_t rows_per_frame, double* stats, int32_t ld_stats, int32_t s0, void* workspace,
                 size_t workspace_bytes, void* stream);

/* Temporal shift-add.  In tensor-core mode the three temporal taps of a 3x3x3 growth convolution
 * (duf_net.py:203,212) are computed as extra OUTPUT columns of one (1,3,3) tap-GEMM (N = 3*g instead of g: the
 * A operand is read once for all three), z[frame][row][kt*g + co]; this kernel finishes the convolution:
 *   out[f][row][c0_out + co] = bias[co] + sum_kt z[f + kt - t_pad][row][kt*g + co],  f < frames_out
 * (t_pad = 1: padding (1,1,1), frames_out = frames_in; t_pad = 0: padding (0,1,1), frames_out = frames_in - 2),
 * and, if stats != NULL, the per-frame BatchNorm statistics of the new slice exactly as vsr_bn_stats would
 * (workspace >= vsr_bn_stats_workspace(frames_out, rows_per_frame, g)). */
int vsr_tshift_add(const void* z, int32_t dtype, int32_t ldz, int32_t g, int32_t frames_in, int64_t rows_per_frame,
                   int32_t t_pad, const float* bias, void* out, int32_t ld_out, int32_t c0_out, int32_t frames_out,
                   double* stats, int32_t ld_stats, int32_t s0, void* workspace, size_t workspace_bytes, void* stream);

/* The gradient side of vsr_tshift_add, feeding the weight gradient of the (1,3,3) form:
 *   dz[f][row][kt*g + co] = dy[f - kt + t_pad][row][c0 + co]  (zero outside the frames_out frames of dy and for the
 * padding columns [3g, ldz)), f < frames_in. */
int vsr_tshift_gather(const void* dy, int32_t dtype, int32_t ld_dy, int32_t c0, int32_t g, int32_t frames_out,
                      int64_t rows_per_frame, int32_t t_pad, void* dz, int32_t ldz, int32_t frames_in, void* stream);

/* One BatchNorm's affine map from the statistics of `frames` frames (training != 0: biased batch variance,
 * running_mean / running_var updated with `momentum` and the unbiased variance when non-NULL) or from the
 * running buffers (training == 0):  scale_shift = float[2][cp] {gamma*rstd, beta - mean*gamma*rstd}, zero
 * for the padding channels [c, cp);  mean_rstd = float[2][c] (kept for the backward pass). */
int vsr_bn_finalize(const double* stats, int32_t ld_stats, int32_t s0, int32_t frames, int64_t rows_per_frame,
                    int32_t c, int32_t cp, const float* gamma, const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var, int32_t training, float* scale_shift,
                    float* mean_rstd, void* stream);

/* y[row][j] = max(0, x[row][c0 + j] * scale[j] + shift[j]) for j < c, 0 for c <= j < cp; y is dense
 * [rows][cp]  (BatchNorm3d + ReLU in front of every Conv3d, duf_net.py:198-203,207-212,114-115). */
int vsr_bn_relu(const void* x, int32_t dtype, int32_t ldx, int32_t c0, int32_t c, int64_t rows,
                const float* scale_shift, int32_t cp, void* y, void* stream);

/* Backward of the same (training mode): with g = dy where the forward output was positive, else 0, and
 * xhat = (x - mean) * rstd:
 *   phase & 1:  dgamma_dbeta = float[2][c] {sum g*xhat, sum g} over this call's rows (written; two passes, fixed order)
 *   phase & 2:  dx[row][c0_dx + j] (+)= gamma*rstd * (g - S_g/count - xhat * S_gx/count), zero for c <= j < cp_dx when
 *               not accumulating; {S_gx, S_g} = `sums` (float[2][c]; NULL = dgamma_dbeta), count <= 0 = rows.
 * phase 3 is the single-device backward.  With synchronised BatchNorm across ranks the caller runs phase 1, sums
 * dgamma_dbeta over the ranks (NCCL) into `sums`, then phase 2 with count = rows of all ranks.
 * dy is [rows][ld_dy] (channels [0, c)). */
size_t vsr_bn_relu_bwd_workspace(int64_t rows, int32_t c);
int vsr_bn_relu_bwd(const void* dy, int32_t ld_dy, const void* x, int32_t dtype, int32_t ldx, int32_t c0,
                    int32_t c, int64_t rows, const float* scale_shift, int32_t cp, const float* mean_rstd,
                    float* dgamma_dbeta, void* dx, int32_t ld_dx, int32_t c0_dx, int32_t cp_dx, int accumulate,
                    int32_t phase, const float* sums, int64_t count, void* workspace, size_t workspace_bytes,
                    void* stream);

/* Dynamic-upsampling-filter tail (duf_net.py:66-97): logits [n*h*w][ld_logits] with channel k*r*r + p
 * (k = tap of the size_filter^2 window, p = sub-pixel), softmax over k, applied to the neighbourhood of the
 * centre frame x (NCHW fp32 [n,cin,h,w], zero padded), + res[pix][c*r*r + p], pixel-shuffled into
 * y (NCHW fp32 [n,cin,r*h,r*w]). */
int vsr_duf_filter(const void* logits, int32_t ld_logits, const void* res, int32_t ld_res, int32_t dtype,
                   const float* x, int32_t n, int32_t cin, int32_t h, int32_t w, int32_t size_filter, int32_t r,
                   float* y, void* stream);
/* its backward w.r.t. logits and res (x is data): dlogits [pix][ld_logits], dres [pix][ld_res], padding
 * channels zeroed. */
int vsr_duf_filter_bwd(const void* logits, int32_t ld_logits, int32_t dtype, const float* x, const float* dy,
                       int32_t n, int32_t cin, int32_t h, int32_t w, int32_t size_filter, int32_t r, void* dlogits,
                       void* dres, int32_t ld_res, void* stream);

/*
 * Device-side data front end (the callers of the path).  One batch of training items from cine volumes
 * resident in device memory, vol = [seqs][frames][h][w] fp32 intensities (LR volume: r = 1; HR volume: r = upscale):
 * per item the temporal window (acdc_vsr_dataset.py:59-78), RandomHorizontalFlip / RandomVerticalFlip /
 * RandomCropPatch (transforms.py:321-450: the crop is taken from the flipped image), Normalize (:154-168) and the
 * default collate, written as frames out = [f_count][n][ph*r][pw*r] (frames f_first .. f_first+f_count-1 of the window).
 * tab: device int32 [n][5 + nf] = {sequence, flip_x, flip_y, y0, x0, frame_0 .. frame_{nf-1}}, y0 / x0 / ph / pw in
 * LR pixels.  The random decisions are the caller's (the reference draws them with Python's `random`).
 */
int vsr_cine_gather(const float* vol, int32_t seqs, int32_t frames, int32_t h, int32_t w_, const int32_t* tab,
                    int32_t n, int32_t nf, int32_t r, int32_t f_first, int32_t f_count, int32_t ph, int32_t pw,
                    float mean, float std, float* out, void* stream);

/*
 * The reference's offline `Downscale` (acdc_preprocess.py:102-180: centred k-space truncation to 1/r per axis, |.|, round,
 * cv2.INTER_CUBIC resize by 1/r, round, clip to [0, 255]) for n frames resident in device memory: the low-resolution side
 * of a dataset without the host.  hr: [n][h][w] fp32 integer-valued; ph / pw: the h x h / w x w complex matrices of the 1-D
 * low-pass operator (interleaved re, im, fp64, built per (size, r) by vsr_b200.data.lowpass_matrix); lr: [n][h/r][w/r] fp32.
 * h, w multiples of r.  FP64 arithmetic: bit-identical to the reference's numpy path unless an exact low-pass value lies
 * within ~1e-10 of x.5.  workspace >= vsr_downscale_workspace(n, h, w) bytes.
 */
size_t vsr_downscale_workspace(int32_t n, int32_t h, int32_t w);
int vsr_downscale(const float* hr, int32_t n, int32_t h, int32_t w, int32_t r, const double* ph, const double* pw, float* lr,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Element-wise pieces of the flow-based recurrent net FRVSRNet (frvsr_net.py:11-239), fp32 maps.
 *   vsr_maxpool2x2{,_bwd}      nn.MaxPool2d(2) of FNet (:127) on a pixel-major map [n][h][w][c] -> [n][h/2][w/2][c]; idx keeps
 *                              the window position of every maximum (torch's scan order: a later element wins when greater)
 *   vsr_upsample2x_nhwc{,_bwd} BilinerUp(2) = F.interpolate(x2, bilinear, align_corners=False) (:135,166-172) on a pixel-major map
 *   vsr_flow_tanh{,_bwd}       nn.Tanh on the two flow channels of the last FNet convolution (:142) + the crop that undoes FNet's
 *                              padding (:147-161): z [n][hp][wp][cz] -> flow [n][2][h][w] taken at offset (y0, x0)
 *   vsr_grid_warp{,_bwd}       STN.forward (:205-226): grid = linspace(-1, 1) mesh + flow, F.grid_sample(bilinear, border,
 *                              align_corners=False) of a one-channel image; the backward gives d(flow) only (the warped image is
 *                              data or a detached output, :47,53)
 *   vsr_s2d_cat{,_bwd}         SpaceToDepth(r) of the warped HR image + torch.cat with the LR frame (:47-48,88,175-191) as the
 *                              pixel-major SRNet input [n][h][w][cpad] (channels r*r+1 .. cpad-1 zero); backward: d(warped image)
 */
int vsr_maxpool2x2(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* y, uint8_t* idx, void* stream);
int vsr_maxpool2x2_bwd(const float* dy, const uint8_t* idx, int32_t n, int32_t h, int32_t w, int32_t c, float* dx, void* stream);
int vsr_upsample2x_nhwc(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* y, void* stream);
int vsr_upsample2x_nhwc_bwd(const float* dy, int32_t n, int32_t h, int32_t w, int32_t c, float* dx, void* stream);
int vsr_flow_tanh(const float* z, int32_t n, int32_t hp, int32_t wp, int32_t cz, int32_t y0, int32_t x0, int32_t h, int32_t w,
                  float* flow, void* stream);
int vsr_flow_tanh_bwd(const float* dflow, const float* flow, int32_t n, int32_t hp, int32_t wp, int32_t cz, int32_t y0, int32_t x0,
                      int32_t h, int32_t w, float* dz, void* stream);
int vsr_grid_warp(const float* img, const float* flow, int32_t n, int32_t h, int32_t w, float* out, void* stream);
int vsr_grid_warp_bwd(const float* img, const float* flow, const float* dout, int32_t n, int32_t h, int32_t w, float* dflow,
                      void* stream);
int vsr_s2d_cat(const float* hr, const float* lr, int32_t n, int32_t h, int32_t w, int32_t r, int32_t cpad, float* out, void* stream);
int vsr_s2d_cat_bwd(const float* dout, int32_t n, int32_t h, int32_t w, int32_t r, int32_t cpad, float* dhr, void* stream);

/*
 * ---- TOFlowNet (toflow_net.py:33-138): bicubic input up-sampling, padding with the batch minimum, SpyNet's pyramid, the flow
 * warp fused with the concatenations, the flow update and the output head.  Images and flows are planar ([n][k][h][w]),
 * feature maps pixel-major [n][h][w][c]; the convolutions and BatchNorm2d run on vsr_tapgemm / vsr_bn_*. --------------------- */
/* F.interpolate(scale_factor=r, mode='bicubic', align_corners=False) (toflow_net.py:34-36): [nc][h][w] -> [nc][rh][rw] */
int vsr_upsample_bicubic(const float* x, int32_t nc, int32_t h, int32_t w, int32_t r, float* y, void* stream);
/* x.min() kept on the device (toflow_net.py:47 pads with it): `partials` (vsr_partials_len() floats) receives partial minima
 * (+inf in the unused rows); vsr_pad_fill folds them and writes F.pad(x, (x0, wp-w-x0, y0, hp-h-y0), value=min) */
int vsr_min_partials(const float* x, int64_t numel, float* partials, void* stream);
int vsr_pad_fill(const float* x, int32_t nc, int32_t h, int32_t w, int32_t y0, int32_t x0, int32_t hp, int32_t wp,
                 const float* partials, float* out, void* stream);
/* F.avg_pool2d(x, 2, 2) of planar images, even sizes (SpyNet.forward, toflow_net.py:76-78) */
int vsr_avgpool2x2(const float* x, int32_t nc, int32_t h, int32_t w, float* y, void* stream);
/* torch.cat([ref, flow_warp(nbr, scale * flow), scale * flow]) into channels c_ref, c_w, c_flow..c_flow+1 of a pixel-major map
 * (toflow_net.py:83-85; ref / nbr may be NULL, c_flow < 0 = no flow channels; other channels are not written); flow_warp =
 * grid_sample(bilinear, zeros) of the pixel mesh + flow (toflow_net.py:117-138).  _bwd: gradient w.r.t. the stored flow. */
int vsr_warp_cat(float* out, int32_t n, int32_t h, int32_t w, int32_t cpad, int32_t c_ref, const float* ref, int32_t c_w,
                 const float* nbr, const float* flow, float scale, int32_t c_flow, void* stream);
int vsr_warp_cat_bwd(const float* dout, int32_t n, int32_t h, int32_t w, int32_t cpad, int32_t c_w, const float* nbr,
                     const float* flow, float scale, int32_t c_flow, float* dflow, void* stream);
/* flow[n][2][h][w] = scale * flow_up + z[..., 0:2] (toflow_net.py:82-83) */
int vsr_flow_add(const float* z, int32_t n, int32_t h, int32_t w, int32_t cz, const float* flow_up, float scale, float* flow,
                 void* stream);
/* dz[n][hp][wp][cz] = d[n][kc][h][w] in channels < kc of the window at (y0, x0), zero elsewhere (gradient of "first kc
 * channels of a pixel-major map, cropped": the flow update and the output head) */
int vsr_planar_to_nhwc(const float* d, int32_t n, int32_t kc, int32_t hp, int32_t wp, int32_t cz, int32_t y0, int32_t x0,
                       int32_t h, int32_t w, float* dz, void* stream);
/* out[n][1][h][w] = z[n][y+y0][x+x0][0] + xref[n][y+y0][x+x0] (toflow_net.py:59-65: out_block(x) + x_ref, un-padded) */
int vsr_head_add(const float* z, int32_t n, int32_t hp, int32_t wp, int32_t cz, const float* xref, int32_t y0, int32_t x0,
                 int32_t h, int32_t w, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSR_B200_H_ */
