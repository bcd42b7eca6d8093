"""DUFNet — the Conv3d network of the reference (src/model/nets/duf_net.py:9-214) on the tap-GEMM kernels.

Layout: every feature map is pixel-major and TIME-MAJOR, [frames, N, h, w, C] (C contiguous), so that
  * a temporal tap of a Conv3d is the same map seen through a pointer shifted by whole frames: the data gradient
    of a 3x3x3 growth convolution is a 27-tap tap-GEMM over three frame-shifted views of the concat gradient
    (a zero frame at both ends gives the temporal padding of `_denseBlock1`; `_denseBlock2` has two frames fewer),
  * its forward pass and weight gradient use the (1,3,3) COLUMN FORM: the three temporal taps are extra output
    columns of a 9-tap tap-GEMM (N = 3G instead of G, one source, the A operand read once for all three) and
    `vsr_tshift_add` / `vsr_tshift_gather` move between the [.., 3G] columns and the concat slice,
  * the temporal crop `concat[:, :, 1:-1]` (duf_net.py:126) is a pointer offset,
  * the dense concatenation (duf_net.py:123-128) is ONE buffer: each layer reads the channel prefix and
    its growth convolution writes its slice in place; `torch.cat` never runs.
BatchNorm3d statistics are kept per frame and per channel of that buffer (each slice is reduced once, when
it is produced, in the shift-add pass) and combined per layer over its frame range.  BatchNorm + ReLU is one
bandwidth-bound pass in front of every convolution (`vsr_bn_relu`), its backward two passes (`vsr_bn_relu_bwd`).
The tail — softmax over the 5x5 taps, local filtering of the centre frame, pixel shuffle, residual add
(duf_net.py:66-97) — is one kernel each way (`vsr_duf_filter`).

bf16 mode runs every convolution on the tcgen05 tap-GEMM: channel counts are padded to multiples of 64 with
structural-zero weights (64 + 32 i -> 64, 128, 128, 192, 192, 256; 3G = 96 columns -> 128); the concat gradient
buffer is one 64-channel window wider than the net needs so that every growth slice can be read as a window.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _flat
from ._lib import EPI_BIAS, EPI_RELU, EPI_RELU_BWD
from .drf_plan import DrfPlan, Layer, _split_nt
from .nets import _PRECISIONS, BaseNet
from .ops import TapTable

BACKBONES = {"_DenseLayer16": (3, 3, 32), "_DenseLayer28": (9, 3, 16), "_DenseLayer52": (21, 3, 16)}
BN_EPS, BN_MOMENTUM = 1e-5, 0.1          # nn.BatchNorm3d defaults (duf_net.py:114,198,201)


class DufPlan(DrfPlan):
    """Tap tables and weight packing of DUFNet; re-uses DrfPlan's slab / un-pack machinery."""

    def __init__(self, in_channels, num_frames, size_filter, r, backbone, bf16):
        self.variant = "duf"
        n1, n2, G = BACKBONES[backbone]
        if num_frames != 2 * n2 + 1:
            raise ValueError(f"DUFNet with {backbone} reduces {2 * n2 + 1} frames to one (duf_net.py:66,93); "
                             f"got num_frames={num_frames}")
        if size_filter % 2 != 1:
            raise ValueError("size_filter must be odd")
        self.cin, self.T, self.sf, self.r, self.bf16 = in_channels, num_frames, size_filter, r, bf16
        self.n1, self.n2, self.Gr, self.L = n1, n2, G, n1 + n2
        self.kc = 64 if bf16 else G
        self.C = [64 + G * i for i in range(self.L + 1)]       # input channels of layer i; C[L] feeds the tail
        self.ctot = self.C[-1]
        self.ccat = self.ctot + (64 - G if bf16 else 0)
        # growth convolutions, forward and weight gradient: the three temporal taps as output columns (3G, padded to a
        # multiple of 64 for the tensor-core kernels)
        self.ntz = -(-3 * G // 64) * 64 if bf16 else 3 * G
        self.cf, self.cr = size_filter * size_filter * r * r, in_channels * r * r
        self.cfp, self.crp = self.pad(self.cf), self.pad(self.cr)
        self.params, self.n_params, self.fwd, self.bwd = {}, 0, {}, {}
        self._declare_params()
        self._build_layers()
        self._finalize()

    def pad(self, c):
        return -(-c // self.kc) * self.kc

    def frames_of(self, i):
        """(first input frame, input frames, first output frame, output frames) of dense layer i; i == L: tail."""
        if i < self.n1:
            return 0, self.T, 0, self.T
        f0 = i - self.n1
        tin = self.T - 2 * f0
        return f0, tin, f0 + 1, tin - 2

    def _declare_params(self):
        C, G, add = self.C, self.Gr, self._add_param
        for i in range(self.L):
            p = f"denseLayer.conv{i}"
            add(p + ".bn1.weight", (C[i],)); add(p + ".bn1.bias", (C[i],))
            add(p + ".conv1.weight", (C[i], C[i], 1, 1, 1)); add(p + ".conv1.bias", (C[i],))
            add(p + ".bn2.weight", (C[i],)); add(p + ".bn2.bias", (C[i],))
            add(p + ".conv2.weight", (G, C[i], 3, 3, 3)); add(p + ".conv2.bias", (G,))
        add("denseLayer.tail.bn.weight", (self.ctot,)); add("denseLayer.tail.bn.bias", (self.ctot,))
        add("denseLayer.tail.conv.weight", (256, self.ctot, 1, 3, 3)); add("denseLayer.tail.conv.bias", (256,))
        add("head.weight", (64, self.cin, 3, 3)); add("head.bias", (64,))
        add("filterNet.conv1.weight", (512, 256, 1, 1, 1)); add("filterNet.conv1.bias", (512,))
        add("filterNet.conv2.weight", (self.cf, 512, 1, 1, 1)); add("filterNet.conv2.bias", (self.cf,))
        add("residualNet.conv1.weight", (256, 256, 1, 1, 1)); add("residualNet.conv1.bias", (256,))
        add("residualNet.conv2.weight", (self.cr, 256, 1, 1, 1)); add("residualNet.conv2.bias", (self.cr,))

    # ---- helpers ---------------------------------------------------------------------------
    @staticmethod
    def _midx(W, *ix):
        """flat parameter indices of W[ix] (broadcast), -1 (structural zero) where an index is out of range."""
        ix = np.broadcast_arrays(*[np.asarray(i) for i in ix])
        ok = np.ones(ix[0].shape, dtype=bool)
        cl = []
        for d, i in zip(W.shape, ix):
            ok &= (i >= 0) & (i < d)
            cl.append(np.clip(i, 0, d - 1))
        return np.where(ok, W.idx(*cl), -1)

    def _mbias(self, name, total, at=0):
        b = self.params[name + ".bias"]
        out = np.full(total, -1, dtype=np.int64)
        out[at:at + b.shape[0]] = b.offset + np.arange(b.shape[0])
        return out

    def _split(self, total):
        if self.bf16:
            for nt in (256, 192, 128, 64):
                if total % nt == 0:
                    return [(i * nt, nt) for i in range(total // nt)]
        return _split_nt(total)

    def _pointwise(self, lname, widx, cin_p, cout_p, bias, src_c0=0, out_c0=0, out_c=None, store=None):
        """1x1x1 convolution cin_p -> cout_p channels; widx(out channel, in channel) -> flat indices."""
        store = self.fwd if store is None else store
        splits = self._split(cout_p)
        nt = splits[0][1]
        j, k = self._jk(nt)
        groups, slabs = [], []
        for o0, _ in splits:
            taps = []
            for b in range(cin_p // self.kc):
                taps.append((0, 0, 0, src_c0 + b * self.kc))
                slabs.append(widx(o0 + j, b * self.kc + k))
            groups.append((out_c0 + o0, taps))
        store[lname] = Layer(lname, TapTable(self.kc, nt, groups), slabs, cout_p if out_c is None else out_c, bias)

    def _build_layers(self):
        C, G, kc, m = self.C, self.Gr, self.kc, self._midx
        for i in range(self.L):
            p = f"denseLayer.conv{i}"
            cp = self.pad(C[i])
            W1, W2 = self._W(p + ".conv1"), self._W(p + ".conv2")
            self._pointwise(f"c1_{i}", lambda o, c, W=W1: m(W, o, c, 0, 0, 0), cp, cp, self._mbias(p + ".conv1", cp))
            self._pointwise(f"c1_{i}", lambda o, c, W=W1: m(W, c, o, 0, 0, 0), cp, cp, None, store=self.bwd)
            # 3x3x3 growth convolution, forward AND weight gradient: the three temporal taps as extra output columns of a
            # (1,3,3) tap-GEMM (N = 3G: one source, the A operand read once for all three); vsr_tshift_add /
            # vsr_tshift_gather move between the [.., 3G] column form and the concat slice.  The bias rides in the
            # centre block (kt = 1: always a valid frame), so the shift-add adds it once and its gradient is the centre
            # block of the column sums of dz.
            j, k = self._jk(self.ntz)
            taps, slabs = [], []
            for ky in range(3):
                for kx in range(3):
                    for b in range(cp // kc):
                        taps.append((0, ky - 1, kx - 1, b * kc))
                        slabs.append(m(W2, np.where(j < 3 * G, j % G, -1), b * kc + k, j // G, ky, kx))
            self.fwd[f"c2z_{i}"] = Layer(f"c2z_{i}", TapTable(kc, self.ntz, [(0, taps)]), slabs, self.ntz,
                                         self._mbias(p + ".conv2", self.ntz, at=G))
            # its data gradient: sources 0..2 = the concat gradient seen through frame shifts +1, 0, -1
            splits = self._split(cp)
            nt = splits[0][1]
            j, k = self._jk(nt)
            groups, slabs = [], []
            for o0, _ in splits:
                taps = []
                for kt in range(3):
                    for ky in (2, 1, 0):          # row shifts ascending with the slab index: the tensor-core kernel
                        for kx in (2, 1, 0):      # then loads ONE taller box for the three taps of a column
                            taps.append((kt, -(ky - 1), -(kx - 1), C[i]))
                            slabs.append(m(W2, k, o0 + j, kt, ky, kx))
                groups.append((o0, taps))
            self.bwd[f"c2_{i}"] = Layer(f"c2_{i}", TapTable(kc, nt, groups), slabs, cp)
        # tail: (1,3,3) convolution ctot -> 256 (+ the ReLU both heads start with, duf_net.py:38,45)
        Wt = self._W("denseLayer.tail.conv")
        ctp = self.pad(self.ctot)
        for store, cin_p, cout_p, widx, bias in (
                (self.fwd, ctp, 256, lambda o, c, ky, kx: m(Wt, o, c, 0, ky, kx), self._mbias("denseLayer.tail.conv", 256)),
                (self.bwd, 256, ctp, lambda o, c, ky, kx: m(Wt, c, o, 0, 2 - ky, 2 - kx), None)):
            splits = self._split(cout_p)
            nt = splits[0][1]
            j, k = self._jk(nt)
            groups, slabs = [], []
            for o0, _ in splits:
                taps = []
                for ky in range(3):
                    for kx in range(3):
                        for b in range(cin_p // kc):
                            taps.append((0, ky - 1, kx - 1, b * kc))
                            slabs.append(widx(o0 + j, b * kc + k, ky, kx))
                groups.append((o0, taps))
            store["tail"] = Layer("tail", TapTable(kc, nt, groups), slabs, cout_p, bias)
        # first convolutions of the filter and residual heads share their input: one 256 -> 512 + 256 GEMM
        Wf1, Wr1 = self._W("filterNet.conv1"), self._W("residualNet.conv1")
        both = lambda o, c: np.where(o + 0 * c < 512, m(Wf1, o, c, 0, 0, 0), m(Wr1, o - 512, c, 0, 0, 0))
        b1 = np.concatenate([self._mbias("filterNet.conv1", 512), self._mbias("residualNet.conv1", 256)])
        self._pointwise("fr1", both, 256, 768, b1)
        self._pointwise("fr1", lambda o, c: both(c, o), 768, 256, None, store=self.bwd)
        Wf2, Wr2 = self._W("filterNet.conv2"), self._W("residualNet.conv2")
        self._pointwise("f2", lambda o, c: m(Wf2, o, c, 0, 0, 0), 512, self.cfp, self._mbias("filterNet.conv2", self.cfp))
        self._pointwise("f2", lambda o, c: m(Wf2, c, o, 0, 0, 0), self.cfp, 512, None, out_c=768, store=self.bwd)
        self._pointwise("r2", lambda o, c: m(Wr2, o, c, 0, 0, 0), 256, self.crp, self._mbias("residualNet.conv2", self.crp),
                        src_c0=512)
        self._pointwise("r2", lambda o, c: m(Wr2, c, o, 0, 0, 0), self.crp, 256, None, out_c0=512, out_c=768,
                        store=self.bwd)


# ---- parameter containers with the reference's names (duf_net.py:102-130,195-214) -----------------------
def _dense_block(cin, cout, pad):
    s = nn.Sequential()
    s.add_module("bn1", nn.BatchNorm3d(cin)); s.add_module("relu1", nn.ReLU())
    s.add_module("conv1", nn.Conv3d(cin, cin, kernel_size=1))
    s.add_module("bn2", nn.BatchNorm3d(cin)); s.add_module("relu2", nn.ReLU())
    s.add_module("conv2", nn.Conv3d(cin, cout, kernel_size=3, padding=pad))
    return s


class _DenseLayerParams(nn.Module):
    def __init__(self, n1, n2, G):
        super().__init__()
        F = 64
        for i in range(n1 + n2):
            setattr(self, f"conv{i}", _dense_block(F, G, 1 if i < n1 else (0, 1, 1)))
            F += G
        self.tail = nn.Sequential()
        self.tail.add_module("bn", nn.BatchNorm3d(F)); self.tail.add_module("relu", nn.ReLU())
        self.tail.add_module("conv", nn.Conv3d(F, 256, kernel_size=(1, 3, 3), padding=(0, 1, 1)))


class _DufFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, T, *args):
        net._pack(True)
        y, saved = net._forward([f.contiguous() for f in args[:T]], True)
        ctx.net, ctx.saved, ctx.T = net, saved, T
        return y

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        gflat = net._backward(ctx.saved, dy.contiguous())
        ctx.saved = None
        net.flat_grad = gflat
        pg = [gflat[p.offset:p.offset + int(np.prod(p.shape))].view(p.shape) for p in net._plan.params.values()]
        return (None, None) + (None,) * ctx.T + tuple(pg)


class DUFNet(BaseNet):
    """Deep video SR with dynamic upsampling filters (reference: duf_net.py:9-99).  Same constructor arguments
    (in_channels, out_channels, num_frames, size_filter, upscale_factor, backbone) and state_dict (BatchNorm3d
    running buffers included); forward(list of num_frames tensors [N,C,h,w]) -> tensor [N,C,r*h,r*w];
    precision 'fp32' (CUDA-core strict mode) | 'bf16' (tcgen05 mode)."""

    def __init__(self, in_channels, out_channels, num_frames, size_filter, upscale_factor, backbone, precision="fp32"):
        super().__init__()
        assert backbone in BACKBONES
        if precision not in _PRECISIONS:
            raise ValueError(f"precision should be one of {sorted(_PRECISIONS)}. Got {precision!r}.")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_frames, self.size_filter, self.upscale_factor = num_frames, size_filter, upscale_factor
        self.backbone, self.precision = backbone, precision
        n1, n2, G = BACKBONES[backbone]
        self.denseLayer = _DenseLayerParams(n1, n2, G)
        self.head = nn.Conv2d(in_channels, 64, kernel_size=3, padding=1)
        cf = size_filter ** 2 * upscale_factor ** 2
        self.filterNet, self.residualNet = nn.Sequential(), nn.Sequential()
        for seq, mid, out in ((self.filterNet, 512, cf), (self.residualNet, 256, in_channels * upscale_factor ** 2)):
            seq.add_module("relu1", nn.ReLU()); seq.add_module("conv1", nn.Conv3d(256, mid, kernel_size=1))
            seq.add_module("relu2", nn.ReLU()); seq.add_module("conv2", nn.Conv3d(mid, out, kernel_size=1))
        self._plan = DufPlan(in_channels, num_frames, size_filter, upscale_factor, backbone, precision == "bf16")
        assert [n for n, _ in self.named_parameters()] == list(self._plan.params), "parameter order differs"
        self._ops = None          # tests may set an emulated backend here; the product uses CudaOps
        self._dev_state = None
        self.flat = self.flat_grad = None
        self._sync_group, self._sync_world = None, 1
        self._flatten()

    def enable_sync_bn(self, process_group=None):
        """Synchronised BatchNorm for data parallelism: every BatchNorm3d normalises with the
        statistics of the GLOBAL batch — one small all-reduce of {sum, sum of squares} per BatchNorm in forward and
        of {sum g*xhat, sum g} in backward — so N ranks with batch B reproduce one device with batch N*B."""
        self._sync_group = process_group
        self._sync_world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        return self

    # ---- flat parameter bucket (same scheme as the DRF nets) ----
    def _flatten(self):
        params = list(self.parameters())
        dev, dt = params[0].device, params[0].dtype
        flat = torch.empty(self._plan.n_params, dtype=dt, device=dev)
        for p, ref in zip(params, self._plan.params.values()):
            n = p.numel()
            flat[ref.offset:ref.offset + n].copy_(p.data.reshape(-1))
            p.data = flat[ref.offset:ref.offset + n].view(ref.shape)
        self.flat = flat
        self._dev_state = None

    def _is_flat(self):
        return _flat.is_flat(self)      # cached: the module tree is walked only after a parameter registration

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._flatten()
        return out

    def _backend(self):
        if self._ops is not None:
            return self._ops
        if self.flat.device.type != "cuda":
            raise RuntimeError("vsr_b200 nets run on CUDA only (there is no CPU fallback); call .to('cuda')")
        from .ops import cuda_ops
        return cuda_ops()

    def _state(self):
        if self._dev_state is None:
            P, dev = self._plan, self.flat.device
            act = torch.float64 if self.flat.dtype == torch.float64 else _PRECISIONS[self.precision]
            st = {"act": act,
                  "fwd_w": torch.empty(P.fwd_w_numel, dtype=act, device=dev),
                  "bwd_w": torch.empty(P.bwd_w_numel, dtype=act, device=dev),
                  "fwd_b": torch.empty(P.fwd_b_numel, dtype=self.flat.dtype, device=dev),
                  "fwd_w_idx": torch.from_numpy(P.fwd_w_idx).to(dev), "bwd_w_idx": torch.from_numpy(P.bwd_w_idx).to(dev),
                  "fwd_b_idx": torch.from_numpy(P.fwd_b_idx).to(dev),
                  "unpack": [(lo, torch.from_numpy(i).to(dev)) for lo, i in P.unpack_passes], "ws": {}}
            b = P.bias_unpack_idx
            nz = (b >= 0).nonzero()[0]
            st["bias_unpack"] = (int(nz.min()), torch.from_numpy(b[nz.min():nz.max() + 1].copy()).to(dev))
            self._dev_state = st
        return self._dev_state

    def _ws(self, key, nbytes):
        st = self._state()["ws"]
        n = (max(int(nbytes), 16) + 7) // 8
        if key not in st or st[key].numel() < n:
            st[key] = torch.empty(n, dtype=torch.float64, device=self.flat.device)
        return st[key]

    def _pview(self, flat, name):
        p = self._plan.params[name]
        return flat[p.offset:p.offset + int(np.prod(p.shape))].view(p.shape)

    def _pack(self, need_bwd):
        st, ops = self._state(), self._backend()
        ops.gather(self.flat, st["fwd_w_idx"], st["fwd_w"])
        ops.gather(self.flat, st["fwd_b_idx"], st["fwd_b"])
        if need_bwd:
            ops.gather(self.flat, st["bwd_w_idx"], st["bwd_w"])

    def _conv(self, lname, srcs, out, epi=0, **kw):
        st, L = self._state(), self._plan.fwd[lname]
        if L.bias_idx is None:
            self._backend().tapgemm(L.table, srcs, out, st["fwd_w"][L.w_off:L.w_off + L.w_numel], epi=epi, **kw)
            return
        self._backend().tapgemm(L.table, srcs, out, st["fwd_w"][L.w_off:L.w_off + L.w_numel],
                                bias=st["fwd_b"][L.b_off:L.b_off + len(L.bias_idx)], epi=EPI_BIAS | epi, **kw)

    def _dgrad(self, lname, srcs, out, epi=0, **kw):
        st, L = self._state(), self._plan.bwd[lname]
        self._backend().tapgemm(L.table, srcs, out, st["bwd_w"][L.w_off:L.w_off + L.w_numel], epi=epi, **kw)

    # ---- BatchNorm3d + ReLU ------------------------------------------------------------------
    def _bn(self, bn, pname, stats, s0, frames, rows_per_frame, c, cp):
        """scale/shift [2,cp] and mean/rstd [2,c] of one BatchNorm3d (batch statistics when training)."""
        ops, dev = self._backend(), self.flat.device
        ss = torch.empty(2, cp, dtype=torch.float32, device=dev)
        mr = torch.empty(2, c, dtype=torch.float32, device=dev)
        if self.flat.dtype == torch.float64:
            ss, mr = ss.double(), mr.double()
        training = self.training                     # nn.BatchNorm3d(track_running_stats=True): batch statistics when training
        if training and self._sync_world > 1:
            stats = stats[:frames, :, s0:s0 + c].sum(0, keepdim=True).contiguous()     # this rank's sums, [1, 2, c]
            dist.all_reduce(stats, group=self._sync_group)
            s0, frames, rows_per_frame = 0, 1, frames * rows_per_frame * self._sync_world
        ops.bn_finalize(stats if training else None, s0, frames, rows_per_frame, c,
                        self._pview(self.flat, pname + ".weight"), self._pview(self.flat, pname + ".bias"), bn.eps,
                        bn.momentum if bn.momentum is not None else BN_MOMENTUM, bn.running_mean, bn.running_var,
                        training, ss, mr)
        if training:
            self._nbt.append(bn.num_batches_tracked)
        return ss, mr

    def _forward(self, frames, save):
        P, ops, st = self._plan, self._backend(), self._state()
        T, G, C = P.T, P.Gr, P.C
        N, cin, h, w = frames[0].shape
        dev, act, rpf = frames[0].device, st["act"], N * h * w
        new = lambda f, c: torch.empty(f, N, h, w, c, dtype=act, device=dev)
        m4 = lambda t: t.reshape(-1, h, w, t.shape[-1])                     # [frames, N, h, w, c] -> [frames*N, h, w, c]
        x = torch.stack(frames).reshape(T * N, cin, h, w)                   # duf_net.py:57-61 (time-major)
        centre = frames[T // 2 if T % 2 == 1 else T // 2 - 1]               # duf_net.py:53-54
        # the concat buffer (frame f of the net = index f + 1, mirroring its gradient buffer).  Every layer reads only
        # channels and frames written before it, so it is not cleared
        cat = torch.empty(T + 2, N, h, w, P.ccat, dtype=act, device=dev)
        need_stats = self.training
        # per-frame {sum, sum of squares} of every concat channel, and of every 1x1x1 convolution's output (one clear)
        stats = torch.zeros(T + P.L, 2, P.pad(P.ctot), dtype=torch.float64, device=dev) if need_stats else None
        stats, stats2 = (stats[:T], stats[T:]) if need_stats else (None, None)
        self._nbt = []
        head = new(T, 64)
        ops.conv3x3_first(x, self._pview(self.flat, "head.weight"), self._pview(self.flat, "head.bias"), None, m4(head))
        ops.copy_window(m4(head), 0, m4(cat[1:T + 1]), 0, 64)
        if need_stats:
            ops.bn_stats(m4(cat[1:T + 1]), 0, 64, T, stats, 0, self._ws("stats", ops.bn_stats_workspace(T, rpf, 64)))
        saved = []
        for i in range(P.L):
            f0, tin, fo, tout = P.frames_of(i)
            blk = getattr(self.denseLayer, f"conv{i}")
            pn = f"denseLayer.conv{i}"
            cp = P.pad(C[i])
            X = m4(cat[1 + f0:1 + f0 + tin])
            ss1, mr1 = self._bn(blk.bn1, pn + ".bn1", stats[f0:f0 + tin] if need_stats else None, 0, tin, rpf, C[i], cp)
            a = new(tin, cp)
            ops.bn_relu(X, 0, C[i], ss1, m4(a))
            b = new(tin, cp)
            self._conv(f"c1_{i}", [m4(a)], m4(b))
            st2 = stats2[i:i + 1] if need_stats else None
            if need_stats:
                ops.bn_stats(m4(b), 0, C[i], 1, st2, 0, self._ws("stats", ops.bn_stats_workspace(1, tin * rpf, C[i])))
            ss2, mr2 = self._bn(blk.bn2, pn + ".bn2", st2, 0, 1, tin * rpf, C[i], cp)
            c = new(tin, cp)
            ops.bn_relu(m4(b), 0, C[i], ss2, m4(c))
            out = m4(cat[1 + fo:1 + fo + tout])
            z = new(tin, P.ntz)
            self._conv(f"c2z_{i}", [m4(c)], m4(z))                          # (1,3,3) column form ...
            ops.tshift_add(m4(z), G, tin, 1 if i < P.n1 else 0, None, out, C[i],     # ... summed over the temporal taps
                           tout, stats[fo:fo + tout] if need_stats else None, C[i],
                           self._ws("stats", ops.bn_stats_workspace(tout, rpf, G)) if need_stats else None)
            saved.append((a, b, c, ss1, mr1, ss2, mr2))
        f0, tin, _, _ = P.frames_of(P.L)
        ctp = P.pad(P.ctot)
        Xt = m4(cat[1 + f0:1 + f0 + tin])
        sst, mrt = self._bn(self.denseLayer.tail.bn, "denseLayer.tail.bn", stats[f0:f0 + tin] if need_stats else None, 0, tin,
                            rpf, P.ctot, ctp)
        at = new(tin, ctp)
        ops.bn_relu(Xt, 0, P.ctot, sst, m4(at))
        feat = new(tin, 256)
        self._conv("tail", [m4(at)], m4(feat), epi=EPI_RELU)               # tail conv + relu1 of both heads
        fr1 = new(tin, 768)
        self._conv("fr1", [m4(feat)], m4(fr1), epi=EPI_RELU)               # filterNet/residualNet conv1 + relu2
        logits, res = new(tin, P.cfp), new(tin, P.crp)
        self._conv("f2", [m4(fr1)], m4(logits))
        self._conv("r2", [m4(fr1)], m4(res))
        y = torch.empty(N, cin, h * P.r, w * P.r, dtype=self.flat.dtype, device=dev)
        ops.duf_filter(m4(logits), m4(res), centre.contiguous(), P.sf, P.r, y)
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)                              # num_batches_tracked of all BatchNorms, one launch
        keep = (x, centre, cat, saved, sst, mrt, at, feat, fr1, logits) if save else None
        return y, keep

    def _backward(self, keep, dy):
        P, ops, st = self._plan, self._backend(), self._state()
        x, centre, cat, saved, sst, mrt, at, feat, fr1, logits = keep
        T, G, C = P.T, P.Gr, P.C
        _, N, h, w, _ = cat.shape
        dev, act, pd, rpf = cat.device, st["act"], self.flat.dtype, N * h * w
        new = lambda f, c: torch.empty(f, N, h, w, c, dtype=act, device=dev)
        m4 = lambda t: t.reshape(-1, h, w, t.shape[-1])
        gflat = torch.zeros(P.n_params, dtype=pd, device=dev)
        dw = torch.zeros(P.fwd_w_numel, dtype=pd, device=dev)
        db = torch.zeros(P.fwd_b_numel, dtype=pd, device=dev)

        def wgrad(lname, srcs, dz):
            L = P.fwd[lname]
            ws = self._ws("wgrad", ops.tapgemm_wgrad_workspace(L.table, srcs, dz))
            dbl = db[L.b_off:L.b_off + L.bias_c]
            if not ops.tapgemm_wgrad(L.table, srcs, dz, dw[L.w_off:L.w_off + L.w_numel], True, ws, db=dbl,
                                     db_period=L.bias_c):
                rows = dz.numel() // L.bias_c
                ops.colsum(dz, rows, L.bias_c, dbl, True, self._ws("colsum", ops.colsum_workspace(rows, L.bias_c)))

        def bn_bwd(pname, dyv, xv, c, ss, mr, dxv, c0_dx, cp_dx, accumulate):
            pw = P.params[pname + ".weight"]                                # weight then bias: adjacent in the bucket
            rows = xv.numel() // xv.shape[-1]
            ws = self._ws("bnbwd", ops.bn_relu_bwd_workspace(rows, c))
            gsl = gflat[pw.offset:pw.offset + 2 * c]
            if self._sync_world > 1:                                        # synchronised BatchNorm: global sums for dx
                ops.bn_relu_bwd(dyv, xv, 0, c, ss, mr, gsl, None, 0, c, False, ws, phase=1)
                tot = gsl.clone()
                dist.all_reduce(tot, group=self._sync_group)
                ops.bn_relu_bwd(dyv, xv, 0, c, ss, mr, gsl, dxv, c0_dx, cp_dx, accumulate, ws, phase=2, sums=tot,
                                count=rows * self._sync_world)
            else:
                ops.bn_relu_bwd(dyv, xv, 0, c, ss, mr, gsl, dxv, c0_dx, cp_dx, accumulate, ws)

        f0, tin, _, _ = P.frames_of(P.L)
        dlogits, dres = new(tin, P.cfp), new(tin, P.crp)
        ops.duf_filter_bwd(m4(logits), centre.contiguous(), dy, P.sf, P.r, m4(dlogits), m4(dres))
        wgrad("f2", [m4(fr1)], m4(dlogits))
        wgrad("r2", [m4(fr1)], m4(dres))
        dfr1 = new(tin, 768)
        self._dgrad("f2", [m4(dlogits)], m4(dfr1), epi=EPI_RELU_BWD, aux_y=m4(fr1))
        self._dgrad("r2", [m4(dres)], m4(dfr1), epi=EPI_RELU_BWD, aux_y=m4(fr1))
        wgrad("fr1", [m4(feat)], m4(dfr1))
        dfeat = new(tin, 256)
        self._dgrad("fr1", [m4(dfr1)], m4(dfeat), epi=EPI_RELU_BWD, aux_y=m4(feat))
        wgrad("tail", [m4(at)], m4(dfeat))
        dat = new(tin, at.shape[-1])
        self._dgrad("tail", [m4(dfeat)], m4(dat))
        dcat = torch.zeros(T + 2, N, h, w, P.ccat, dtype=act, device=dev)
        bn_bwd("denseLayer.tail.bn", m4(dat), m4(cat[1 + f0:1 + f0 + tin]), P.ctot, sst, mrt,
               m4(dcat[1 + f0:1 + f0 + tin]), 0, P.ctot, True)
        for i in reversed(range(P.L)):
            f0, tin, fo, tout = P.frames_of(i)
            pn = f"denseLayer.conv{i}"
            a, b, c, ss1, mr1, ss2, mr2 = saved[i]
            cp = a.shape[-1]
            dz = m4(dcat[1 + fo:1 + fo + tout])
            if i < P.n1:
                gviews = [m4(dcat[1 + f0 + 1 - kt:1 + f0 + 1 - kt + tin]) for kt in range(3)]
            else:
                gviews = [m4(dcat[f0 + 2 - kt:f0 + 2 - kt + tin]) for kt in range(3)]
            dzz = new(tin, P.ntz)
            ops.tshift_gather(dz, C[i], G, tout, 1 if i < P.n1 else 0, m4(dzz), tin)
            wgrad(f"c2z_{i}", [m4(c)], m4(dzz))
            dc = new(tin, cp)
            self._dgrad(f"c2_{i}", gviews, m4(dc))
            dbm = new(tin, cp)
            bn_bwd(pn + ".bn2", m4(dc), m4(b), C[i], ss2, mr2, m4(dbm), 0, cp, False)
            wgrad(f"c1_{i}", [m4(a)], m4(dbm))
            da = new(tin, cp)
            self._dgrad(f"c1_{i}", [m4(dbm)], m4(da))
            bn_bwd(pn + ".bn1", m4(da), m4(cat[1 + f0:1 + f0 + tin]), C[i], ss1, mr1,
                   m4(dcat[1 + f0:1 + f0 + tin]), 0, C[i], True)
        dhead = new(T, 64)
        ops.copy_window(m4(dcat[1:T + 1]), 0, m4(dhead), 0, 64)
        ws = self._ws("first", ops.conv3x3_first_bwd_workspace(x, 64))
        ops.conv3x3_first_bwd(x, m4(dhead), self._pview(gflat, "head.weight"), self._pview(gflat, "head.bias"), True, ws)
        for lo, idx in st["unpack"]:
            ops.gather_add(dw, idx, gflat[lo:lo + idx.numel()])
        lo, idx = st["bias_unpack"]
        ops.gather_add(db, idx, gflat[lo:lo + idx.numel()])
        return gflat

    def forward(self, inputs):
        frames = list(inputs)
        if len(frames) != self.num_frames:
            raise ValueError(f"expected {self.num_frames} frames, got {len(frames)}")
        for f in frames:
            if f.dim() != 4 or f.shape[1] != self.in_channels:
                raise ValueError(f"expected frames of shape [N,{self.in_channels},h,w], got {tuple(f.shape)}")
        if not self._is_flat():
            self._flatten()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if not self.training:
                raise NotImplementedError("DUFNet backward needs training mode (batch statistics)")
            return _DufFunction.apply(self, len(frames), *frames, *self.parameters())
        self._pack(False)
        return self._forward([f.contiguous() for f in frames], False)[0]
