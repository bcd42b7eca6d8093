"""Tensor-level wrappers over the C-ABI (include/vsr_b200.h).  PyTorch is used here only for
device memory and streams; every function launches hand-written sm_100a kernels and raises
VsrError if the shared library is missing or a call fails.  No CPU path exists.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (EPI_BIAS, EPI_OUT2, EPI_OUT2_SUB, EPI_PRELU, EPI_PRELU_BWD, EPI_RELU, EPI_RELU_BWD,  # noqa: F401
                   EPI_RES_PRE, EPI_SCALE, VSR_BF16, VSR_F32, VsrTapGemmDesc, VsrTensor4, check)

_DT = {torch.float32: VSR_F32, torch.bfloat16: VSR_BF16}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.VsrError("vsr_b200 ops need CUDA tensors (there is no CPU fallback)")
        if t is not None and not t.is_contiguous():
            raise _lib.VsrError("vsr_b200 ops need contiguous tensors")


@dataclass
class TapTable:
    """Host description of a tap-GEMM: groups of (o0, [(src, dy, dx, c0), ...])."""
    kc: int
    nt: int
    groups: List[Tuple[int, List[Tuple[int, int, int, int]]]]
    _dev: dict = field(default_factory=dict, repr=False)
    useful: float = 1.0      # share of the table's MACs that are algorithmic (< 1: structural-zero weight rows)

    @property
    def n_groups(self):
        return len(self.groups)

    @property
    def n_taps_total(self):
        return sum(len(t) for _, t in self.groups)

    def flat_taps(self):
        return [t for _, taps in self.groups for t in taps]

    def device_tabs(self, device):
        key = str(device)
        if key not in self._dev:
            g_rows, t_rows, begin = [], [], 0
            for o0, taps in self.groups:
                g_rows.append([o0, begin, len(taps), 0])
                begin += len(taps)
                t_rows.extend([list(t) for t in taps])
            self._dev[key] = (torch.tensor(g_rows, dtype=torch.int32, device=device),
                              torch.tensor(t_rows, dtype=torch.int32, device=device))
        return self._dev[key]

    def host_taps(self):
        """Host copy of the tap table (kept alive with the table): VsrTapGemmDesc.tap_tab_host."""
        if "host" not in self._dev:
            self._dev["host"] = torch.tensor([list(t) for t in self.flat_taps()], dtype=torch.int32).contiguous()
        return self._dev["host"]

    def host_groups(self):
        """Host copy of the group table: VsrTapGemmDesc.group_tab_host."""
        if "host_groups" not in self._dev:
            rows, begin = [], 0
            for o0, taps in self.groups:
                rows.append([o0, begin, len(taps), 0])
                begin += len(taps)
            self._dev["host_groups"] = torch.tensor(rows, dtype=torch.int32).contiguous()
        return self._dev["host_groups"]


def _tensor4(t: torch.Tensor) -> VsrTensor4:
    n, h, w, c = t.shape
    return VsrTensor4(t.data_ptr(), n, h, w, c)


def _make_desc(tab: TapTable, srcs: Sequence[torch.Tensor], out: torch.Tensor):
    if len(srcs) > _lib.VSR_MAX_SRCS:
        raise _lib.VsrError(f"tap-GEMM takes at most {_lib.VSR_MAX_SRCS} sources")
    _need_cuda(out, *srcs)
    d = VsrTapGemmDesc()
    d.dtype = _DT[out.dtype]
    d.kc, d.nt, d.n_srcs = tab.kc, tab.nt, len(srcs)
    for i, s in enumerate(srcs):
        if s.dtype != out.dtype:
            raise _lib.VsrError("tap-GEMM sources and output must share a dtype")
        d.srcs[i] = _tensor4(s)
    d.out = _tensor4(out)
    gt, tt = tab.device_tabs(out.device)
    d.n_groups, d.n_taps_total = tab.n_groups, tab.n_taps_total
    d.max_group_taps = max(len(t) for _, t in tab.groups)
    d.group_tab, d.tap_tab = gt.data_ptr(), tt.data_ptr()
    d.tap_tab_host = tab.host_taps().data_ptr()
    d.group_tab_host = tab.host_groups().data_ptr()
    return d


class _TimedLib:
    """bench.py's per-kernel pass: the ctypes handle with a CUDA-event pair around every kernel-launching entry
    point.  Records (entry name, start, end, meta, return code) where meta = (family, flops, shape signature, algorithmic bytes) for
    the tap-GEMM family (set by the calling CudaOps method) and None for the others."""

    _QUERIES = ("vsr_abi_version", "vsr_last_error", "vsr_partials_len", "vsr_slab_index", "vsr_metric_workspace",
                "vsr_reload_tunables")

    def __init__(self, lib, ops, sink):
        self._lib, self._ops, self._sink = lib, ops, sink

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("vsr_") or name.endswith("_workspace") or name in self._QUERIES:
            return fn
        ops, sink = self._ops, self._sink

        def timed(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            sink.append((name, e0, e1, ops._meta, rc))
            return rc
        return timed


class CudaOps:
    """The product backend: every method is one or two kernel launches on the current stream."""
    name = "cuda"

    def __init__(self):
        self.lib = _lib.lib()
        self.partials_len = self.lib.vsr_partials_len()
        self.launches = 0
        self.timing = None      # a list while bench.py's per-kernel pass runs (start_timing / stop_timing)
        self._meta = None

    def start_timing(self):
        """CUDA events around every kernel-launching C-ABI call from now on (bench.py; not for production runs)"""
        self.timing = []
        self.lib = _TimedLib(_lib.lib(), self, self.timing)

    def stop_timing(self):
        t, self.timing, self.lib = self.timing, None, _lib.lib()
        return t

    @staticmethod
    def gemm_records(records):
        """(family, flops, start, end, shape signature, algorithmic bytes) of the tap-GEMM / weight-gradient launches"""
        return [(m[0], m[1], e0, e1, m[2], m[3]) for name, e0, e1, m, rc in records
                if m is not None and not (name == "vsr_tapgemm_wgrad_partial" and rc != 1)]

    # ---- tap-GEMM ----------------------------------------------------------------------
    def tapgemm(self, tab, srcs, out, w, bias=None, epi=0, out_scale=1.0, slope=None, residual=None,
                aux_y=None, out2=None, res2=None, slope_partials=None, force_simt=False):
        d = _make_desc(tab, srcs, out)
        _need_cuda(w, bias, slope, residual, aux_y, out2, res2, slope_partials)
        d.w, d.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
        d.epi, d.out_scale = epi, out_scale
        d.slope = slope.data_ptr() if slope is not None else None
        d.residual = residual.data_ptr() if residual is not None else None
        d.aux_y = aux_y.data_ptr() if aux_y is not None else None
        d.out2 = out2.data_ptr() if out2 is not None else None
        d.res2 = res2.data_ptr() if res2 is not None else None
        d.slope_partials = slope_partials.data_ptr() if slope_partials is not None else None
        fn = self.lib.vsr_tapgemm_simt_bf16 if force_simt else self.lib.vsr_tapgemm
        if self.timing is not None:
            pix = out.shape[0] * out.shape[1] * out.shape[2]
            sig = f"taps{tab.n_taps_total}_nt{tab.nt}_g{tab.n_groups}_px{pix}_epi{epi}"
            es = out.element_size()
            nbytes = es * (sum(pix * s.shape[-1] for s in srcs) + pix * out.shape[-1])
            # epilogue operands read (residual, saved activation, second residual) / written (second output)
            n_extra = ((epi & EPI_RES_PRE) != 0) + ((epi & (EPI_PRELU_BWD | EPI_RELU_BWD)) != 0) + 2 * ((epi & EPI_OUT2) != 0)
            nbytes += es * pix * out.shape[-1] * n_extra
            self._meta = ("tapgemm", 2.0 * pix * tab.n_taps_total * tab.nt * tab.kc * tab.useful, sig, nbytes)
        check(fn(C.byref(d), _stream()), "vsr_tapgemm")
        self._meta = None
        self.launches += 1

    def tapgemm_wgrad_workspace(self, tab, srcs, dz):
        return self.lib.vsr_tapgemm_wgrad_workspace(C.byref(_make_desc(tab, srcs, dz)))

    # (the names SplitOps gives to the CUDA-core path of mixed nets; here they are the methods themselves)
    def tapgemm_plain(self, tab, srcs, out, w, **kw):
        return self.tapgemm(tab, srcs, out, w, **kw)

    def tapgemm_wgrad_plain(self, tab, srcs, dz, dw, accumulate, workspace, db=None, db_period=0):
        return self.tapgemm_wgrad(tab, srcs, dz, dw, accumulate, workspace, db=db, db_period=db_period)

    def tapgemm_wgrad_workspace_plain(self, tab, srcs, dz):
        return self.tapgemm_wgrad_workspace(tab, srcs, dz)

    def tapgemm_wgrad(self, tab, srcs, dz, dw, accumulate, workspace, db=None, db_period=0):
        """weight gradient; with db, also tries to fuse the bias gradient — returns True if db was produced."""
        d = _make_desc(tab, srcs, dz)
        _need_cuda(dw, workspace, db)
        if self.timing is not None:
            self._meta = self._wgrad_meta(tab, srcs, dz)
        fused = False
        nbytes_ws = workspace.numel() * workspace.element_size()
        if db is not None:
            rc = self.lib.vsr_tapgemm_wgrad_bias(C.byref(d), _p(dw), _p(db), db_period, int(accumulate),
                                                 _p(workspace), nbytes_ws, _stream())
            if rc < 0:
                check(rc, "vsr_tapgemm_wgrad_bias")
            fused = rc == 1
        else:
            check(self.lib.vsr_tapgemm_wgrad(C.byref(d), _p(dw), int(accumulate), _p(workspace), nbytes_ws,
                                             _stream()), "vsr_tapgemm_wgrad")
        self._meta = None
        self.launches += 2
        return fused

    @staticmethod
    def _wgrad_meta(tab, srcs, dz):
        pix = dz.shape[0] * dz.shape[1] * dz.shape[2]
        sig = f"taps{tab.n_taps_total}_nt{tab.nt}_g{tab.n_groups}_px{pix}"
        nbytes = dz.element_size() * (sum(pix * s.shape[-1] for s in srcs) + pix * dz.shape[-1])
        return ("wgrad", 2.0 * pix * tab.n_taps_total * tab.nt * tab.kc, sig, nbytes)

    def tapgemm_wgrad_partial(self, tab, srcs, dz, workspace, slice_, n_slices, db_period):
        """per-split partials of dw / db into slice `slice_` of `workspace`; False if unsupported."""
        d = _make_desc(tab, srcs, dz)
        _need_cuda(workspace)
        if self.timing is not None:
            self._meta = self._wgrad_meta(tab, srcs, dz)
        rc = self.lib.vsr_tapgemm_wgrad_partial(C.byref(d), db_period, int(slice_), int(n_slices), _p(workspace),
                                                workspace.numel() * workspace.element_size(), _stream())
        self._meta = None
        if rc < 0:
            check(rc, "vsr_tapgemm_wgrad_partial")
        if rc == 1:
            self.launches += 1
        return rc == 1

    def tapgemm_wgrad_finish(self, tab, srcs, dz, dw, db, db_period, accumulate, used, n_slices, workspace):
        d = _make_desc(tab, srcs, dz)
        _need_cuda(dw, db, workspace)
        check(self.lib.vsr_tapgemm_wgrad_finish(C.byref(d), _p(dw), _p(db), db_period, int(accumulate), int(used),
                                                int(n_slices), _p(workspace),
                                                workspace.numel() * workspace.element_size(), _stream()),
              "vsr_tapgemm_wgrad_finish")
        self.launches += 1

    # ---- weight gradients of several 1x1 convolutions over one feature list, in one pass -----
    MAX_SHARED_ACC, MAX_SHARED_DZ, MAX_SHARED_SRCS = 8, _lib.VSR_WS_MAX_DZ, _lib.VSR_WS_MAX_SRCS

    def _shared_desc(self, srcs, dzs, ntaps, dws, dbs):
        d = _lib.VsrWgradSharedDesc()
        d.n_srcs, d.n_dz = len(srcs), len(dzs)
        for i, t in enumerate(srcs):
            d.srcs[i] = _tensor4(t)
        for g, (z, nt, dw, db) in enumerate(zip(dzs, ntaps, dws, dbs)):
            d.dzs[g], d.ntaps[g] = _tensor4(z), nt
            d.dw[g] = dw.data_ptr()
            d.db[g] = db.data_ptr() if db is not None else None
        return d

    def wgrad_shared_ok(self, srcs, dzs, ntaps):
        """can `wgrad_shared` take this layer set in ONE launch?  (bf16 maps of 64 channels, <= 8 accumulators)"""
        return (all(t.dtype == torch.bfloat16 and t.shape[-1] == 64 for t in (*srcs, *dzs)) and len(srcs) <= self.MAX_SHARED_SRCS and
                len(dzs) <= self.MAX_SHARED_DZ and sum((n + 1) // 2 for n in ntaps) <= self.MAX_SHARED_ACC)

    def wgrad_shared(self, srcs, dzs, ntaps, dws, dbs, accumulate, workspace_of):
        """dws[g][t] (+)= dzs[g]^T srcs[t] for t < ntaps[g], dbs[g] (+)= column sums of dzs[g]; every source is read once.
        `workspace_of(nbytes)` returns a workspace tensor of at least that many bytes."""
        _need_cuda(*srcs, *dzs, *dws, *[b for b in dbs if b is not None])
        d = self._shared_desc(srcs, dzs, ntaps, dws, dbs)
        ws = workspace_of(self.lib.vsr_wgrad_shared_workspace(C.byref(d)))
        if self.timing is not None:
            pix = dzs[0].numel() // 64
            nbytes = 2 * pix * 64 * (len(srcs) + len(dzs))
            self._meta = ("wgrad", 2.0 * pix * sum(ntaps) * 64 * 64, f"shared_srcs{len(srcs)}_dz{len(dzs)}_taps{sum(ntaps)}_px{pix}", nbytes)
        check(self.lib.vsr_wgrad_shared(C.byref(d), int(accumulate), _p(ws), ws.numel() * ws.element_size(), _stream()),
              "vsr_wgrad_shared")
        self._meta = None
        self.launches += 2

    # ---- small kernels -----------------------------------------------------------------
    def colsum_workspace(self, rows, c):
        return self.lib.vsr_colsum_workspace(rows, c)

    def colsum(self, x, rows, c, db, accumulate, workspace):
        _need_cuda(x, db, workspace)
        check(self.lib.vsr_colsum(_p(x), _DT[x.dtype], rows, c, _p(db), int(accumulate), _p(workspace),
                                  workspace.numel() * workspace.element_size(), _stream()), "vsr_colsum")
        self.launches += 2

    def conv3x3_first(self, x, w, bias, slope, y):
        _need_cuda(x, w, bias, slope, y)
        n, cin, h, w_ = x.shape
        check(self.lib.vsr_conv3x3_first(_p(x), n, cin, h, w_, _p(w), _p(bias), _p(slope), _p(y),
                                         _DT[y.dtype], y.shape[-1], _stream()), "vsr_conv3x3_first")
        self.launches += 1

    def conv3x3_first_bwd_workspace(self, x, cout):
        n, cin, h, w_ = x.shape
        return self.lib.vsr_conv3x3_first_bwd_workspace(n, cin, h, w_, cout)

    def conv3x3_first_bwd(self, x, dz, dw, db, accumulate, workspace):
        _need_cuda(x, dz, dw, db, workspace)
        n, cin, h, w_ = x.shape
        check(self.lib.vsr_conv3x3_first_bwd(_p(x), n, cin, h, w_, _p(dz), _DT[dz.dtype], dz.shape[-1],
                                             _p(dw), _p(db), int(accumulate), _p(workspace),
                                             workspace.numel() * workspace.element_size(), _stream()),
              "vsr_conv3x3_first_bwd")
        self.launches += 2

    @staticmethod
    def _phase_arr(phase_yx):
        flat = [v for yx in phase_yx for v in yx]
        return (C.c_int32 * len(flat))(*flat)

    def conv3x3_last(self, x, r, c, phase_yx, w, bias, y):
        _need_cuda(x, w, bias, y)
        n, h, w_, _ = x.shape
        check(self.lib.vsr_conv3x3_last(_p(x), _DT[x.dtype], n, h, w_, r, c, self._phase_arr(phase_yx),
                                        _p(w), _p(bias), _p(y), y.shape[1], _stream()), "vsr_conv3x3_last")
        self.launches += 1

    def conv3x3_last_bwd_workspace(self, x, r, c, cout):
        n, h, w_, _ = x.shape
        return self.lib.vsr_conv3x3_last_bwd_workspace(n, h, w_, r, c, cout)

    def conv3x3_last_bwd(self, x, r, c, phase_yx, w, dy, dx, dw, db, accumulate, workspace):
        _need_cuda(x, w, dy, dx, dw, db, workspace)
        n, h, w_, _ = x.shape
        check(self.lib.vsr_conv3x3_last_bwd(_p(x), _DT[x.dtype], n, h, w_, r, c, self._phase_arr(phase_yx),
                                            _p(w), _p(dy), dy.shape[1], _p(dx), _p(dw), _p(db),
                                            int(accumulate), _p(workspace),
                                            workspace.numel() * workspace.element_size(), _stream()),
              "vsr_conv3x3_last_bwd")
        self.launches += 2

    def act_bwd(self, dy, y, dz, slope=None, slope_partials=None):
        _need_cuda(dy, y, dz, slope, slope_partials)
        check(self.lib.vsr_act_bwd(_p(dy), _p(y), _p(dz), _DT[dy.dtype], dy.numel(), _p(slope),
                                   _p(slope_partials), _stream()), "vsr_act_bwd")
        self.launches += 1

    def add(self, a, b, out):
        _need_cuda(a, b, out)
        check(self.lib.vsr_add(_p(a), _p(b), _p(out), _DT[a.dtype], a.numel(), _stream()), "vsr_add")
        self.launches += 1

    def axpby(self, a, b, out, alpha, beta):
        """out = alpha * a + beta * b (b may be None when beta == 0)"""
        _need_cuda(a, b, out)
        check(self.lib.vsr_axpby(_p(a), _p(b), _p(out), _DT[a.dtype], a.numel(), float(alpha), float(beta), _stream()), "vsr_axpby")
        self.launches += 1

    def reduce_partials(self, partials, rows, row_dst, dst):
        _need_cuda(partials, row_dst, dst)
        check(self.lib.vsr_reduce_partials(_p(partials), rows, partials.shape[-1], _p(row_dst), _p(dst),
                                           _stream()), "vsr_reduce_partials")
        self.launches += 1

    def gather(self, src, idx, dst):
        _need_cuda(src, idx, dst)
        check(self.lib.vsr_gather(_p(src), _p(idx), _p(dst), _DT[dst.dtype], idx.numel(), _stream()),
              "vsr_gather")
        self.launches += 1

    def gather_add(self, src, idx, dst):
        _need_cuda(src, idx, dst)
        check(self.lib.vsr_gather_add(_p(src), _p(idx), _p(dst), idx.numel(), _stream()), "vsr_gather_add")
        self.launches += 1

    def cast(self, src, dst):
        _need_cuda(src, dst)
        check(self.lib.vsr_cast(_p(src), _DT[src.dtype], _p(dst), _DT[dst.dtype], src.numel(), _stream()),
              "vsr_cast")
        self.launches += 1

    def zero_(self, t):
        t.zero_()  # cudaMemsetAsync on the current stream

    # ---- Conv3d network: BatchNorm3d + ReLU on channel windows, dynamic filter tail ---------
    @staticmethod
    def _rows(t):
        return t.numel() // t.shape[-1]

    def copy_window(self, src, c0_src, dst, c0_dst, c):
        _need_cuda(src, dst)
        check(self.lib.vsr_copy_window(_p(src), src.shape[-1], c0_src, _p(dst), dst.shape[-1], c0_dst, c,
                                       self._rows(src), _DT[src.dtype], _stream()), "vsr_copy_window")
        self.launches += 1

    def bn_stats_workspace(self, frames, rows_per_frame, c):
        return self.lib.vsr_bn_stats_workspace(frames, rows_per_frame, c)

    def bn_stats(self, x, c0, c, frames, stats, s0, workspace):
        """per-frame sum / sum of squares of x[..., c0:c0+c] (x: `frames` stacked frames) -> stats[f, :, s0:s0+c]"""
        _need_cuda(x, stats, workspace)
        check(self.lib.vsr_bn_stats(_p(x), _DT[x.dtype], x.shape[-1], c0, c, frames, self._rows(x) // frames,
                                    _p(stats), stats.shape[-1], s0, _p(workspace),
                                    workspace.numel() * workspace.element_size(), _stream()), "vsr_bn_stats")
        self.launches += 2

    def tshift_add(self, z, g, frames_in, t_pad, bias, out, c0_out, frames_out, stats, s0, workspace):
        """out[f, ..., c0_out + co] = bias[co] + sum_kt z[f + kt - t_pad, ..., kt*g + co] (+ the slice's BatchNorm statistics)"""
        _need_cuda(z, bias, out, stats, workspace)
        check(self.lib.vsr_tshift_add(_p(z), _DT[z.dtype], z.shape[-1], g, frames_in, self._rows(z) // frames_in, t_pad,
                                      _p(bias), _p(out), out.shape[-1], c0_out, frames_out, _p(stats),
                                      stats.shape[-1] if stats is not None else 0, s0, _p(workspace),
                                      workspace.numel() * workspace.element_size() if workspace is not None else 0,
                                      _stream()), "vsr_tshift_add")
        self.launches += 2 if stats is not None else 1

    def tshift_gather(self, dy, c0, g, frames_out, t_pad, dz, frames_in):
        """dz[f, ..., kt*g + co] = dy[f - kt + t_pad, ..., c0 + co] (zero outside dy's frames / in the column padding)"""
        _need_cuda(dy, dz)
        check(self.lib.vsr_tshift_gather(_p(dy), _DT[dy.dtype], dy.shape[-1], c0, g, frames_out,
                                         self._rows(dy) // frames_out, t_pad, _p(dz), dz.shape[-1], frames_in,
                                         _stream()), "vsr_tshift_gather")
        self.launches += 1

    def bn_finalize(self, stats, s0, frames, rows_per_frame, c, gamma, beta, eps, momentum, running_mean,
                    running_var, training, scale_shift, mean_rstd):
        _need_cuda(stats, gamma, beta, running_mean, running_var, scale_shift, mean_rstd)
        check(self.lib.vsr_bn_finalize(_p(stats), stats.shape[-1] if stats is not None else 0, s0, frames,
                                       rows_per_frame, c, scale_shift.shape[-1], _p(gamma), _p(beta), eps, momentum,
                                       _p(running_mean), _p(running_var), int(training), _p(scale_shift),
                                       _p(mean_rstd), _stream()), "vsr_bn_finalize")
        self.launches += 1

    def bn_relu(self, x, c0, c, scale_shift, y):
        _need_cuda(x, scale_shift, y)
        check(self.lib.vsr_bn_relu(_p(x), _DT[x.dtype], x.shape[-1], c0, c, self._rows(x), _p(scale_shift),
                                   y.shape[-1], _p(y), _stream()), "vsr_bn_relu")
        self.launches += 1

    def bn_relu_bwd_workspace(self, rows, c):
        return self.lib.vsr_bn_relu_bwd_workspace(rows, c)

    def bn_relu_bwd(self, dy, x, c0, c, scale_shift, mean_rstd, dgamma_dbeta, dx, c0_dx, cp_dx, accumulate, workspace,
                    phase=3, sums=None, count=0):
        """phase 1: dgamma_dbeta = {sum g*xhat, sum g} of this rank's rows; phase 2: dx from `sums` (default
        dgamma_dbeta) over `count` rows (default: this call's); 3: both (single device)."""
        _need_cuda(dy, x, scale_shift, mean_rstd, dgamma_dbeta, dx, workspace, sums)
        check(self.lib.vsr_bn_relu_bwd(_p(dy), dy.shape[-1], _p(x), _DT[x.dtype], x.shape[-1], c0, c, self._rows(x),
                                       _p(scale_shift), scale_shift.shape[-1], _p(mean_rstd), _p(dgamma_dbeta),
                                       _p(dx), dx.shape[-1] if dx is not None else 0, c0_dx, cp_dx, int(accumulate),
                                       phase, _p(sums), int(count), _p(workspace),
                                       workspace.numel() * workspace.element_size(), _stream()), "vsr_bn_relu_bwd")
        self.launches += (2 if phase & 1 else 0) + (1 if phase & 2 else 0)

    def duf_filter(self, logits, res, x, size_filter, r, y):
        _need_cuda(logits, res, x, y)
        n, cin, h, w_ = x.shape
        check(self.lib.vsr_duf_filter(_p(logits), logits.shape[-1], _p(res), res.shape[-1], _DT[logits.dtype], _p(x),
                                      n, cin, h, w_, size_filter, r, _p(y), _stream()), "vsr_duf_filter")
        self.launches += 1

    def duf_filter_bwd(self, logits, x, dy, size_filter, r, dlogits, dres):
        _need_cuda(logits, x, dy, dlogits, dres)
        n, cin, h, w_ = x.shape
        check(self.lib.vsr_duf_filter_bwd(_p(logits), logits.shape[-1], _DT[logits.dtype], _p(x), _p(dy), n, cin, h,
                                          w_, size_filter, r, _p(dlogits), _p(dres), dres.shape[-1], _stream()),
              "vsr_duf_filter_bwd")
        self.launches += 1

    # ---- device-side data front end --------------------------------------------------------
    def cine_gather(self, vol, tab, r, f_first, f_count, mean, std, out):
        """out[f, i] = normalised crop of frame tab[i, 5 + f_first + f] of sequence tab[i, 0] (flips, crop from `tab`)"""
        _need_cuda(vol, tab, out)
        s_, t_, h, w_ = vol.shape
        check(self.lib.vsr_cine_gather(_p(vol), s_, t_, h, w_, _p(tab), tab.shape[0], tab.shape[1] - 5, r, f_first, f_count,
                                       out.shape[-2] // r, out.shape[-1] // r, float(mean), float(std), _p(out), _stream()),
              "vsr_cine_gather")
        self.launches += 1

    def downscale(self, hr, r, ph, pw, lr):
        """lr[i] = Downscale(r)(hr[i]) of acdc_preprocess.py:102-180 for frames [n, h, w] on the device; ph / pw: the complex
        low-pass matrices of the two axes ([h, h, 2] / [w, w, 2] fp64, data.lowpass_matrix)"""
        _need_cuda(hr, ph, pw, lr)
        n, h, w_ = hr.shape
        nbytes = self.lib.vsr_downscale_workspace(n, h, w_)
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=hr.device)
        check(self.lib.vsr_downscale(_p(hr), n, h, w_, r, _p(ph), _p(pw), _p(lr), _p(ws), ws.numel() * 8, _stream()),
              "vsr_downscale")
        self.launches += 3

    # ---- flow-based recurrent net (FRVSRNet): pooling / up-sampling of pixel-major maps, flow head, warp ----------
    def maxpool2x2(self, x, y, idx):
        _need_cuda(x, y, idx)
        n, h, w_, c = x.shape
        check(self.lib.vsr_maxpool2x2(_p(x), n, h, w_, c, _p(y), _p(idx), _stream()), "vsr_maxpool2x2")
        self.launches += 1

    def maxpool2x2_bwd(self, dy, idx, dx):
        _need_cuda(dy, idx, dx)
        n, h, w_, c = dx.shape
        check(self.lib.vsr_maxpool2x2_bwd(_p(dy), _p(idx), n, h, w_, c, _p(dx), _stream()), "vsr_maxpool2x2_bwd")
        self.launches += 1

    def upsample2x_nhwc(self, x, y):
        _need_cuda(x, y)
        n, h, w_, c = x.shape
        check(self.lib.vsr_upsample2x_nhwc(_p(x), n, h, w_, c, _p(y), _stream()), "vsr_upsample2x_nhwc")
        self.launches += 1

    def upsample2x_nhwc_bwd(self, dy, dx):
        _need_cuda(dy, dx)
        n, h, w_, c = dx.shape
        check(self.lib.vsr_upsample2x_nhwc_bwd(_p(dy), n, h, w_, c, _p(dx), _stream()), "vsr_upsample2x_nhwc_bwd")
        self.launches += 1

    def flow_tanh(self, z, y0, x0, flow):
        """flow [n, 2, h, w] = tanh(z[:, y0:y0+h, x0:x0+w, :2]) for a pixel-major z [n, hp, wp, cz]"""
        _need_cuda(z, flow)
        n, hp, wp, cz = z.shape
        check(self.lib.vsr_flow_tanh(_p(z), n, hp, wp, cz, y0, x0, flow.shape[2], flow.shape[3], _p(flow), _stream()),
              "vsr_flow_tanh")
        self.launches += 1

    def flow_tanh_bwd(self, dflow, flow, y0, x0, dz):
        _need_cuda(dflow, flow, dz)
        n, hp, wp, cz = dz.shape
        check(self.lib.vsr_flow_tanh_bwd(_p(dflow), _p(flow), n, hp, wp, cz, y0, x0, flow.shape[2], flow.shape[3], _p(dz),
                                         _stream()), "vsr_flow_tanh_bwd")
        self.launches += 1

    def grid_warp(self, img, flow, out):
        _need_cuda(img, flow, out)
        n, _, h, w_ = img.shape
        check(self.lib.vsr_grid_warp(_p(img), _p(flow), n, h, w_, _p(out), _stream()), "vsr_grid_warp")
        self.launches += 1

    def grid_warp_bwd(self, img, flow, dout, dflow):
        _need_cuda(img, flow, dout, dflow)
        n, _, h, w_ = img.shape
        check(self.lib.vsr_grid_warp_bwd(_p(img), _p(flow), _p(dout), n, h, w_, _p(dflow), _stream()), "vsr_grid_warp_bwd")
        self.launches += 1

    def s2d_cat(self, hr, lr, r, out):
        _need_cuda(hr, lr, out)
        n, h, w_, cpad = out.shape
        check(self.lib.vsr_s2d_cat(_p(hr), _p(lr), n, h, w_, r, cpad, _p(out), _stream()), "vsr_s2d_cat")
        self.launches += 1

    def s2d_cat_bwd(self, dout, r, dhr):
        _need_cuda(dout, dhr)
        n, h, w_, cpad = dout.shape
        check(self.lib.vsr_s2d_cat_bwd(_p(dout), n, h, w_, r, cpad, _p(dhr), _stream()), "vsr_s2d_cat_bwd")
        self.launches += 1

    # ---- TOFlowNet: bicubic input up-sampling, min padding, pyramid, warp + concat, flow update, head ---------------------
    def upsample_bicubic(self, x, r, y):
        _need_cuda(x, y)
        h, w_ = x.shape[-2:]
        check(self.lib.vsr_upsample_bicubic(_p(x), x.numel() // (h * w_), h, w_, r, _p(y), _stream()), "vsr_upsample_bicubic")
        self.launches += 1

    def min_partials(self, x, partials):
        _need_cuda(x, partials)
        check(self.lib.vsr_min_partials(_p(x), x.numel(), _p(partials), _stream()), "vsr_min_partials")
        self.launches += 1

    def pad_fill(self, x, y0, x0, partials, out):
        _need_cuda(x, partials, out)
        h, w_ = x.shape[-2:]
        check(self.lib.vsr_pad_fill(_p(x), x.numel() // (h * w_), h, w_, y0, x0, out.shape[-2], out.shape[-1], _p(partials), _p(out),
                                    _stream()), "vsr_pad_fill")
        self.launches += 1

    def avgpool2x2(self, x, y):
        _need_cuda(x, y)
        h, w_ = x.shape[-2:]
        check(self.lib.vsr_avgpool2x2(_p(x), x.numel() // (h * w_), h, w_, _p(y), _stream()), "vsr_avgpool2x2")
        self.launches += 1

    def warp_cat(self, out, c_ref, ref, c_w, nbr, flow, scale, c_flow):
        """out[..., c_ref] = ref, out[..., c_w] = flow_warp(nbr, scale * flow), out[..., c_flow:c_flow+2] = scale * flow"""
        _need_cuda(out, ref, nbr, flow)
        n, h, w_, cpad = out.shape
        check(self.lib.vsr_warp_cat(_p(out), n, h, w_, cpad, c_ref, _p(ref), c_w, _p(nbr), _p(flow), float(scale), c_flow, _stream()),
              "vsr_warp_cat")
        self.launches += 1

    def warp_cat_bwd(self, dout, c_w, nbr, flow, scale, c_flow, dflow):
        _need_cuda(dout, nbr, flow, dflow)
        n, h, w_, cpad = dout.shape
        check(self.lib.vsr_warp_cat_bwd(_p(dout), n, h, w_, cpad, c_w, _p(nbr), _p(flow), float(scale), c_flow, _p(dflow), _stream()),
              "vsr_warp_cat_bwd")
        self.launches += 1

    def flow_add(self, z, flow_up, scale, flow):
        _need_cuda(z, flow_up, flow)
        n, h, w_, cz = z.shape
        check(self.lib.vsr_flow_add(_p(z), n, h, w_, cz, _p(flow_up), float(scale), _p(flow), _stream()), "vsr_flow_add")
        self.launches += 1

    def planar_to_nhwc(self, d, y0, x0, dz):
        """dz[n, hp, wp, cz] = d[n, kc, h, w] in the first kc channels of the window at (y0, x0), zero elsewhere"""
        _need_cuda(d, dz)
        n, kc, h, w_ = d.shape
        check(self.lib.vsr_planar_to_nhwc(_p(d), n, kc, dz.shape[1], dz.shape[2], dz.shape[3], y0, x0, h, w_, _p(dz), _stream()),
              "vsr_planar_to_nhwc")
        self.launches += 1

    def head_add(self, z, xref, y0, x0, out):
        _need_cuda(z, xref, out)
        n, hp, wp, cz = z.shape
        check(self.lib.vsr_head_add(_p(z), n, hp, wp, cz, _p(xref), y0, x0, out.shape[-2], out.shape[-1], _p(out), _stream()),
              "vsr_head_add")
        self.launches += 1

    # ---- loss / metrics ----------------------------------------------------------------
    def loss_fwd_bwd(self, out, target, kind, param, grad_scale, partials, grad):
        _need_cuda(out, target, partials, grad)
        check(self.lib.vsr_loss_fwd_bwd(_p(out), _p(target), out.numel(), kind, param, grad_scale,
                                        _p(partials), _p(grad), _stream()), "vsr_loss_fwd_bwd")
        self.launches += 1

    def loss_fwd_bwd_seg(self, out, target, n_segments, kind, param, grad_scale, partials, grad):
        """all `n_segments` equally sized frames of out / target (one contiguous tensor each) in one launch; row s of
        `partials` [n_segments, partials_len] receives the partial sums of frame s"""
        _need_cuda(out, target, partials, grad)
        check(self.lib.vsr_loss_fwd_bwd_seg(_p(out), _p(target), out.numel() // n_segments, n_segments, kind, param,
                                            grad_scale, _p(partials), _p(grad), _stream()), "vsr_loss_fwd_bwd_seg")
        self.launches += 1

    def metric_workspace(self, n, per_sample):
        return self.lib.vsr_metric_workspace(n, per_sample)

    def psnr(self, out, target, mean, std, max_value, psnr_out, workspace):
        _need_cuda(out, target, psnr_out, workspace)
        n = out.shape[0]
        check(self.lib.vsr_psnr(_p(out), _p(target), n, out.numel() // n, mean, std, max_value,
                                _p(psnr_out), _p(workspace), workspace.numel() * workspace.element_size(),
                                _stream()), "vsr_psnr")
        self.launches += 2

    def ssim(self, out, target, win11, mean, std, c1, c2, ssim_out, workspace):
        _need_cuda(out, target, win11, ssim_out, workspace)
        n, h, w_ = out.shape[0], out.shape[-2], out.shape[-1]
        check(self.lib.vsr_ssim(_p(out), _p(target), n, h, w_, _p(win11), mean, std, c1, c2, _p(ssim_out),
                                _p(workspace), workspace.numel() * workspace.element_size(), _stream()),
              "vsr_ssim")
        self.launches += 2

    def ssim3d_workspace(self, n, d, h, w_):
        return self.lib.vsr_ssim3d_workspace(n, d, h, w_)

    def ssim3d(self, out, target, win11, mean, std, c1, c2, ssim_out, workspace):
        _need_cuda(out, target, win11, ssim_out, workspace)
        n, d, h, w_ = out.shape
        check(self.lib.vsr_ssim3d(_p(out), _p(target), n, d, h, w_, _p(win11), mean, std, c1, c2, _p(ssim_out),
                                  _p(workspace), workspace.numel() * workspace.element_size(), _stream()),
              "vsr_ssim3d")
        self.launches += 4

    # ---- standalone resampling / optimiser ------------------------------------------------
    def pixel_shuffle(self, x, y, r, inverse=False):
        _need_cuda(x, y)
        if not inverse:
            n, cr2, h, w_ = x.shape
            c = cr2 // (r * r)
        else:
            n, c, hh, ww = x.shape
            h, w_ = hh // r, ww // r
        check(self.lib.vsr_pixel_shuffle(_p(x), _p(y), n, c, h, w_, r, int(inverse), _stream()),
              "vsr_pixel_shuffle")
        self.launches += 1

    def upsample_linear(self, x, y, align_corners):
        _need_cuda(x, y)
        d, h, w_ = (1, *x.shape[-2:]) if x.dim() == 4 else x.shape[-3:]
        od, oh, ow = (1, *y.shape[-2:]) if y.dim() == 4 else y.shape[-3:]
        nc = x.shape[0] * x.shape[1]
        check(self.lib.vsr_upsample_linear(_p(x), _p(y), nc, d, h, w_, od, oh, ow, int(align_corners),
                                           _stream()), "vsr_upsample_linear")
        self.launches += 1

    def upsample_linear_bwd(self, dy, dx, align_corners):
        _need_cuda(dy, dx)
        d, h, w_ = (1, *dx.shape[-2:]) if dx.dim() == 4 else dx.shape[-3:]
        od, oh, ow = (1, *dy.shape[-2:]) if dy.dim() == 4 else dy.shape[-3:]
        nc = dx.shape[0] * dx.shape[1]
        check(self.lib.vsr_upsample_linear_bwd(_p(dy), _p(dx), nc, d, h, w_, od, oh, ow,
                                               int(align_corners), _stream()), "vsr_upsample_linear_bwd")
        self.launches += 1

    def adam_flat(self, p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
        _need_cuda(p, g, m, v)
        check(self.lib.vsr_adam_flat(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps,
                                     weight_decay, step, grad_scale, _stream()), "vsr_adam_flat")
        self.launches += 1

    def scale_(self, x, alpha):
        _need_cuda(x)
        check(self.lib.vsr_scale(_p(x), x.numel(), float(alpha), _stream()), "vsr_scale")
        self.launches += 1

    def adam_flat_dev(self, p, g, m, v, hyper):
        _need_cuda(p, g, m, v, hyper)
        check(self.lib.vsr_adam_flat_dev(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(hyper), _stream()),
              "vsr_adam_flat_dev")
        self.launches += 1


LO_FLAG = 1 << 30      # vsr_gather_split: bit 30 of an index = the low-order bf16 part of the parameter


class SplitOps(CudaOps):
    """precision='bf16x3': the strict mode on tensor cores (csrc/split.cu, VSR_BF16X2 in include/vsr_b200.h).

    The engine keeps fp32 maps; a tap-GEMM here is  split every source into two bf16 planes -> the tcgen05 kernel on a
    3x-expanded tap table (xh*wh + xl*wh + xh*wl, raw fp32 accumulators into `out`) -> the epilogue in fp32, in place.
    A weight gradient is three launches of the tcgen05 weight-gradient kernel on the ORIGINAL table
    (src_h x dz_h, src_l x dz_h, src_h x dz_l: the planes are plain bf16 maps) reduced together in a fixed order.
    Weight buffers hold [wh | wh | wl] slabs per group (DrfPlan.split_index).  Everything else (first / last
    convolution, activations, losses, Adam) is the fp32 arithmetic of precision='fp32'."""
    name = "cuda-bf16x3"
    split = True

    # Plane cache: a map is split once per step, not once per consumer (an LR feature map of DRFNet feeds up to G 1x1
    # convolutions, a gradient map several data gradients and a weight gradient).  An entry is keyed by the byte range of
    # the fp32 map; it dies when a tap-GEMM of this backend writes an overlapping range (`out` / `out2`) and ALL entries
    # die on every other kernel-launching method (they may write maps: first / last convolution, activation backward, ...;
    # `gather_split`, the first launch of every step, clears the cache too).  Only methods that cannot write a map are
    # exempt (_READ_ONLY).  Tensors written behind this backend's back (torch ops) must be complete before the next
    # cache-clearing launch: the engines only do that to raw frames and to the stacked hidden state, both assembled before
    # the last convolution of the forward pass.
    _READ_ONLY = ("tapgemm", "tapgemm_wgrad", "tapgemm_wgrad_workspace", "tapgemm_wgrad_partial", "colsum", "colsum_workspace",
                  "reduce_partials", "gather", "gather_add", "conv3x3_first_bwd", "conv3x3_first_bwd_workspace",
                  "conv3x3_last_bwd_workspace", "metric_workspace", "start_timing", "stop_timing", "gemm_records", "table3",
                  "wgrad_shared", "wgrad_shared_ok", "tapgemm_plain", "tapgemm_wgrad_plain", "tapgemm_wgrad_workspace_plain")

    def __init__(self):
        super().__init__()
        self._wg_ws = None
        self._db_dummy = None
        self._retired = []
        self._cache = {}

    def __getattribute__(self, name):
        attr = object.__getattribute__(self, name)
        if name.startswith("_") or not callable(attr) or name in SplitOps._READ_ONLY:
            return attr
        object.__getattribute__(self, "_cache").clear()       # a launch that may write maps: every cached plane pair dies
        return attr

    def _invalidate(self, t):
        if t is None or not self._cache:
            return
        lo = t.data_ptr()
        hi = lo + t.numel() * t.element_size()
        for k in [k for k in self._cache if k[0] < hi and lo < k[1]]:
            del self._cache[k]

    def _planes(self, t, cached=True):
        """fp32 [n,h,w,c] -> bf16 [2,n,h,w,c]: t = planes[0] + planes[1] to 16 significant bits"""
        key = (t.data_ptr(), t.data_ptr() + t.numel() * 4)
        if cached and key in self._cache:
            return self._cache[key]
        _need_cuda(t)
        if t.dtype != torch.float32:
            raise _lib.VsrError("bf16x3 mode takes fp32 maps")
        pl = torch.empty((2, *t.shape), dtype=torch.bfloat16, device=t.device)
        check(self.lib.vsr_split_planes(_p(t), _p(pl), t.numel(), _stream()), "vsr_split_planes")
        self.launches += 1
        if cached:
            self._cache[key] = pl
        return pl

    @staticmethod
    def table3(tab: TapTable) -> TapTable:
        """every tap (s, dy, dx, c0) -> high plane x wh, low plane (source index | 8) x wh, high plane x wl; the slabs of a
        group are ordered [wh of all taps | wh of all taps | wl of all taps] like its taps"""
        t3 = tab._dev.get("x3")
        if t3 is None:
            groups = []
            for o0, taps in tab.groups:
                for s, _, _, _ in taps:
                    if s >= 8:
                        raise _lib.VsrError("bf16x3: at most 8 sources per tap-GEMM")
                groups.append((o0, list(taps) + [(s | 8, dy, dx, c0) for s, dy, dx, c0 in taps] + list(taps)))
            t3 = tab._dev["x3"] = TapTable(tab.kc, tab.nt, groups, useful=tab.useful)
        return t3

    def tapgemm(self, tab, srcs, out, w, bias=None, epi=0, out_scale=1.0, slope=None, residual=None,
                aux_y=None, out2=None, res2=None, slope_partials=None, force_simt=False):
        _need_cuda(out, w, bias, slope, residual, aux_y, out2, res2, slope_partials)
        if out.dtype != torch.float32 or w.dtype != torch.bfloat16:
            raise _lib.VsrError("bf16x3 mode: fp32 maps, bf16 [wh | wh | wl] weight slabs")
        self._invalidate(out)
        self._invalidate(out2)
        planes = [self._planes(s) for s in srcs]
        t3 = self.table3(tab)
        if w.numel() != t3.n_taps_total * t3.nt * t3.kc:
            raise _lib.VsrError("bf16x3 mode: the weight buffer must hold three slabs per tap")
        d = VsrTapGemmDesc()
        d.dtype = _lib.VSR_BF16X2
        d.kc, d.nt, d.n_srcs = t3.kc, t3.nt, len(srcs)
        for i, (s, pl) in enumerate(zip(srcs, planes)):
            n, h, w_, c = s.shape
            d.srcs[i] = VsrTensor4(pl.data_ptr(), n, h, w_, c)
        d.out = _tensor4(out)
        gt, tt = t3.device_tabs(out.device)
        d.n_groups, d.n_taps_total = t3.n_groups, t3.n_taps_total
        d.max_group_taps = max(len(t) for _, t in t3.groups)
        d.group_tab, d.tap_tab = gt.data_ptr(), tt.data_ptr()
        d.tap_tab_host = t3.host_taps().data_ptr()
        d.group_tab_host = t3.host_groups().data_ptr()
        d.w, d.epi, d.out_scale = w.data_ptr(), 0, 1.0
        pix = out.shape[0] * out.shape[1] * out.shape[2]
        if self.timing is not None:
            sig = f"x3_taps{tab.n_taps_total}_nt{tab.nt}_g{tab.n_groups}_px{pix}_epi{epi}"
            nbytes = 4 * (sum(pix * s.shape[-1] for s in srcs) + pix * out.shape[-1])
            self._meta = ("tapgemm", 2.0 * pix * tab.n_taps_total * tab.nt * tab.kc * tab.useful, sig, nbytes)
        check(self.lib.vsr_tapgemm(C.byref(d), _stream()), "vsr_tapgemm(bf16x2)")
        self._meta = None
        # the epilogue pass also writes the bf16 planes of what it produces (the next tap-GEMM that reads the map finds them
        # in the cache instead of re-reading the fp32 map); a launch without an epilogue leaves the raw accumulators as they are
        pl = pl2 = None
        if epi:
            pl = torch.empty((2, *out.shape), dtype=torch.bfloat16, device=out.device)
            if out2 is not None:
                pl2 = torch.empty((2, *out2.shape), dtype=torch.bfloat16, device=out.device)
        check(self.lib.vsr_tap_epilogue(_p(out), pix, out.shape[-1], _p(bias), epi, float(out_scale), _p(slope), _p(residual),
                                        _p(aux_y), _p(out2), _p(res2), _p(slope_partials), _p(pl), _p(pl2), _stream()),
              "vsr_tap_epilogue")
        if pl is not None:
            self._cache[(out.data_ptr(), out.data_ptr() + out.numel() * 4)] = pl
        if pl2 is not None:
            self._cache[(out2.data_ptr(), out2.data_ptr() + out2.numel() * 4)] = pl2
        self.launches += 2 if epi else 1

    def tapgemm_wgrad_workspace(self, tab, srcs, dz):
        return 16            # the three-slice workspace is this backend's own (tapgemm_wgrad)

    def tapgemm_wgrad_partial(self, tab, srcs, dz, workspace, slice_, n_slices, db_period):
        return False         # per-frame deferral is not used in this mode: one three-term launch set per call

    def tapgemm_wgrad(self, tab, srcs, dz, dw, accumulate, workspace, db=None, db_period=0):
        """dw (+)= src_h x dz_h + src_l x dz_h + src_h x dz_l; returns False: the bias gradient is the caller's fp32 colsum"""
        _need_cuda(dw, dz, *srcs)
        # (saved activations feed several weight gradients - the LR / HR feature lists of DRFNet - so their planes are cached
        # like the tap-GEMM sources; a gradient map feeds one weight gradient: split, not kept)
        # (a cached pair may have been made from another 4-D view of the same bytes: take the operand's own shape)
        pl_s, pl_z = [self._planes(s).view(2, *s.shape) for s in srcs], self._planes(dz, cached=False)
        hi, lo = [p[0] for p in pl_s], [p[1] for p in pl_s]
        period = dz.shape[-1]
        terms = ((hi, pl_z[0]), (lo, pl_z[0]), (hi, pl_z[1]))
        descs = [_make_desc(tab, ss, zz) for ss, zz in terms]
        need = 3 * self.lib.vsr_tapgemm_wgrad_workspace(C.byref(descs[0]))
        # (a workspace that has to grow is replaced, never freed: a captured CUDA graph of another net may still replay
        #  launches that write into the old one)
        if self._wg_ws is None or self._wg_ws.numel() * 4 < need or self._wg_ws.device != dz.device:
            self._retired.append(self._wg_ws)
            self._wg_ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=dz.device)
        if self._db_dummy is None or self._db_dummy.numel() < period or self._db_dummy.device != dz.device:
            self._retired.append(self._db_dummy)
            self._db_dummy = torch.empty(max(period, 1024), dtype=torch.float32, device=dz.device)
        ws, nbytes = self._wg_ws, self._wg_ws.numel() * 4
        if self.timing is not None:
            m = self._wgrad_meta(tab, srcs, dz)
            self._meta = (m[0], m[1], "x3_" + m[2], m[3])
        deferred = True
        for k, d in enumerate(descs):
            rc = self.lib.vsr_tapgemm_wgrad_partial(C.byref(d), period, k, 3, _p(ws), nbytes, _stream())
            if rc < 0:
                check(rc, "vsr_tapgemm_wgrad_partial")
            if rc != 1:          # a shape outside the one-reduction path (maps wider than 1024 channels, ...): term by term
                deferred = False
                break
        if deferred:
            check(self.lib.vsr_tapgemm_wgrad_finish(C.byref(descs[0]), _p(dw), _p(self._db_dummy), period, int(accumulate), 3, 3,
                                                    _p(ws), nbytes, _stream()), "vsr_tapgemm_wgrad_finish")
            self.launches += 4
        else:
            for k, d in enumerate(descs):
                check(self.lib.vsr_tapgemm_wgrad(C.byref(d), _p(dw), int(accumulate or k > 0), _p(ws), nbytes, _stream()),
                      "vsr_tapgemm_wgrad")
            self.launches += 6
        self._meta = None
        return False

    def wgrad_shared_ok(self, srcs, dzs, ntaps):
        return (all(t.dtype == torch.float32 and t.shape[-1] == 64 for t in (*srcs, *dzs)) and len(srcs) <= self.MAX_SHARED_SRCS and
                len(dzs) <= self.MAX_SHARED_DZ and sum((n + 1) // 2 for n in ntaps) <= self.MAX_SHARED_ACC)

    def wgrad_shared(self, srcs, dzs, ntaps, dws, dbs, accumulate, workspace_of):
        """the three bf16 terms of every product through the shared-source kernel: (src_h, dz_h), (src_l, dz_h), (src_h, dz_l);
        the bias gradients are the column sums of dz_h + dz_l"""
        pl_s = [self._planes(s).view(2, *s.shape) for s in srcs]
        pl_z = [self._planes(z, cached=False) for z in dzs]
        hi, lo = [p[0] for p in pl_s], [p[1] for p in pl_s]
        zh, zl = [p[0] for p in pl_z], [p[1] for p in pl_z]
        none = [None] * len(dzs)
        base = CudaOps.wgrad_shared
        base(self, hi, zh, ntaps, dws, dbs, accumulate, workspace_of)
        base(self, lo, zh, ntaps, dws, none, True, workspace_of)
        base(self, hi, zl, ntaps, dws, dbs, True, workspace_of)

    # the CUDA-core fp32 tap-GEMM through THIS backend (nets that mix narrow CUDA-core layers with 64-channel tensor-core
    # layers: FRVSRNet): the cache must see the write
    def tapgemm_plain(self, tab, srcs, out, w, **kw):
        self._invalidate(out)
        self._invalidate(kw.get("out2"))
        return CudaOps.tapgemm(self, tab, srcs, out, w, **kw)

    def tapgemm_wgrad_plain(self, tab, srcs, dz, dw, accumulate, workspace, db=None, db_period=0):
        return CudaOps.tapgemm_wgrad(self, tab, srcs, dz, dw, accumulate, workspace, db=db, db_period=db_period)

    def tapgemm_wgrad_workspace_plain(self, tab, srcs, dz):
        return CudaOps.tapgemm_wgrad_workspace(self, tab, srcs, dz)

    def gather_split(self, src, idx, dst):
        _need_cuda(src, idx, dst)
        check(self.lib.vsr_gather_split(_p(src), _p(idx), _p(dst), idx.numel(), _stream()), "vsr_gather_split")
        self.launches += 1


_OPS = None
_SPLIT_OPS = None


def cuda_ops() -> CudaOps:
    global _OPS
    if _OPS is None:
        _OPS = CudaOps()
    return _OPS


def split_ops() -> SplitOps:
    global _SPLIT_OPS
    if _SPLIT_OPS is None:
        _SPLIT_OPS = SplitOps()
    return _SPLIT_OPS


def slab_index(j, k):
    """Element offset of (row j, k) in the swizzled bf16 [nt][64] slab image (vsr_slab_index)."""
    return j * 64 + (((k >> 3) ^ (j & 7)) << 3) + (k & 7)
