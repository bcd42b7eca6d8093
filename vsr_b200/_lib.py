"""ctypes binding of libvsr_sm100.so (see include/vsr_b200.h).

The product has no CPU fallback: if the shared library is missing the import of any op raises.
"""
import ctypes as C
import os

from . import build as _build

VSR_F32, VSR_BF16, VSR_BF16X2 = 0, 1, 2
VSR_MAX_SRCS = 8

EPI_BIAS, EPI_RES_PRE, EPI_PRELU, EPI_RELU = 1, 2, 4, 8
EPI_PRELU_BWD, EPI_RELU_BWD, EPI_OUT2, EPI_SCALE = 16, 32, 64, 128
EPI_OUT2_SUB = 256


class VsrTensor4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32)]


class VsrTapGemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("kc", C.c_int32), ("nt", C.c_int32), ("n_srcs", C.c_int32),
        ("srcs", VsrTensor4 * VSR_MAX_SRCS), ("out", VsrTensor4),
        ("n_groups", C.c_int32), ("n_taps_total", C.c_int32), ("max_group_taps", C.c_int32),
        ("group_tab", C.c_void_p), ("tap_tab", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p),
        ("epi", C.c_int32), ("out_scale", C.c_float),
        ("slope", C.c_void_p), ("residual", C.c_void_p), ("aux_y", C.c_void_p),
        ("out2", C.c_void_p), ("res2", C.c_void_p), ("slope_partials", C.c_void_p),
        ("tap_tab_host", C.c_void_p), ("group_tab_host", C.c_void_p),
    ]


VSR_WS_MAX_SRCS, VSR_WS_MAX_DZ = 8, 4


class VsrWgradSharedDesc(C.Structure):
    _fields_ = [("n_srcs", C.c_int32), ("n_dz", C.c_int32), ("srcs", VsrTensor4 * VSR_WS_MAX_SRCS),
                ("dzs", VsrTensor4 * VSR_WS_MAX_DZ), ("ntaps", C.c_int32 * VSR_WS_MAX_DZ),
                ("dw", C.c_void_p * VSR_WS_MAX_DZ), ("db", C.c_void_p * VSR_WS_MAX_DZ)]


_SIGS = {
    "vsr_abi_version": (C.c_int, []),
    "vsr_last_error": (C.c_char_p, []),
    "vsr_partials_len": (C.c_int, []),
    "vsr_reload_tunables": (None, []),
    "vsr_tapgemm": (C.c_int, [C.POINTER(VsrTapGemmDesc), C.c_void_p]),
    "vsr_tapgemm_simt_bf16": (C.c_int, [C.POINTER(VsrTapGemmDesc), C.c_void_p]),
    "vsr_tapgemm_wgrad_workspace": (C.c_size_t, [C.POINTER(VsrTapGemmDesc)]),
    "vsr_tapgemm_wgrad": (C.c_int, [C.POINTER(VsrTapGemmDesc), C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_size_t, C.c_void_p]),
    "vsr_tapgemm_wgrad_bias": (C.c_int, [C.POINTER(VsrTapGemmDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_int,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_tapgemm_wgrad_partial": (C.c_int, [C.POINTER(VsrTapGemmDesc), C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_size_t, C.c_void_p]),
    "vsr_tapgemm_wgrad_finish": (C.c_int, [C.POINTER(VsrTapGemmDesc), C.c_void_p, C.c_void_p, C.c_int32, C.c_int,
                                           C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_wgrad_shared_workspace": (C.c_size_t, [C.POINTER(VsrWgradSharedDesc)]),
    "vsr_wgrad_shared": (C.c_int, [C.POINTER(VsrWgradSharedDesc), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_colsum_workspace": (C.c_size_t, [C.c_int64, C.c_int32]),
    "vsr_colsum": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int,
                             C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_conv3x3_first": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_int32, C.c_void_p]),
    "vsr_conv3x3_first_bwd_workspace": (C.c_size_t, [C.c_int32] * 5),
    "vsr_conv3x3_first_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_conv3x3_last": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_void_p]),
    "vsr_conv3x3_last_bwd_workspace": (C.c_size_t, [C.c_int32] * 6),
    "vsr_conv3x3_last_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_act_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "vsr_axpby": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "vsr_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "vsr_reduce_partials": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "vsr_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "vsr_gather_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "vsr_slab_index": (C.c_int64, [C.c_int32, C.c_int32]),
    "vsr_split_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "vsr_tap_epilogue": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_float] + [C.c_void_p] * 9),
    "vsr_gather_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "vsr_loss_fwd_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                   C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_loss_fwd_bwd_seg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_metric_workspace": (C.c_size_t, [C.c_int32, C.c_int64]),
    "vsr_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_float,
                           C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_ssim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                           C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                           C.c_size_t, C.c_void_p]),
    "vsr_ssim3d_workspace": (C.c_size_t, [C.c_int32] * 4),
    "vsr_ssim3d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                             C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                             C.c_size_t, C.c_void_p]),
    "vsr_pixel_shuffle": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int, C.c_void_p]),
    "vsr_upsample_linear": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_int,
                                                                                    C.c_void_p]),
    "vsr_upsample_linear_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_int,
                                                                                        C.c_void_p]),
    "vsr_adam_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_float,
                                C.c_void_p]),
    "vsr_adam_flat_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p]),
    "vsr_scale": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "vsr_cast": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "vsr_cine_gather": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p] + [C.c_int32] * 7 +
                        [C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "vsr_maxpool2x2": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 3),
    "vsr_maxpool2x2_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2),
    "vsr_upsample2x_nhwc": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2),
    "vsr_upsample2x_nhwc_bwd": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2),
    "vsr_flow_tanh": (C.c_int, [C.c_void_p] + [C.c_int32] * 8 + [C.c_void_p] * 2),
    "vsr_flow_tanh_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 8 + [C.c_void_p] * 2),
    "vsr_grid_warp": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 3 + [C.c_void_p] * 2),
    "vsr_grid_warp_bwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p] * 2),
    "vsr_s2d_cat": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p] * 2),
    "vsr_s2d_cat_bwd": (C.c_int, [C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p] * 2),
    "vsr_upsample_bicubic": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2),
    "vsr_min_partials": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vsr_pad_fill": (C.c_int, [C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p] * 3),
    "vsr_avgpool2x2": (C.c_int, [C.c_void_p] + [C.c_int32] * 3 + [C.c_void_p] * 2),
    "vsr_warp_cat": (C.c_int, [C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_int32,
                                                              C.c_void_p]),
    "vsr_warp_cat_bwd": (C.c_int, [C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_void_p,
                                                                  C.c_void_p]),
    "vsr_flow_add": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "vsr_planar_to_nhwc": (C.c_int, [C.c_void_p] + [C.c_int32] * 9 + [C.c_void_p] * 2),
    "vsr_head_add": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 2),
    "vsr_downscale_workspace": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "vsr_downscale": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_copy_window": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int64, C.c_int32, C.c_void_p]),
    "vsr_bn_stats_workspace": (C.c_size_t, [C.c_int32, C.c_int64, C.c_int32]),
    "vsr_bn_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                               C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_tshift_add": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                 C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vsr_tshift_gather": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                    C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "vsr_bn_finalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsr_bn_relu": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                              C.c_int32, C.c_void_p, C.c_void_p]),
    "vsr_bn_relu_bwd_workspace": (C.c_size_t, [C.c_int64, C.c_int32]),
    "vsr_bn_relu_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "vsr_duf_filter": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p] +
                       [C.c_int32] * 6 + [C.c_void_p, C.c_void_p]),
    "vsr_duf_filter_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 +
                           [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
}

# every symbol include/vsr_b200.h declares (vsr_tapgemm_simt_bf16 is a test hook, not in the header)
HEADER_SYMBOLS = [s for s in _SIGS if s not in ("vsr_tapgemm_simt_bf16",)]

_lib = None


class VsrError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        path = _build.LIBPATH
        if not os.path.exists(path):
            raise VsrError(f"{path} is missing: run `python -m vsr_b200.build` "
                           "(there is no CPU fallback for the vsr_b200 kernels)")
        handle = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.vsr_abi_version() != 4:
            raise VsrError("libvsr_sm100.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().vsr_last_error()
        raise VsrError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
