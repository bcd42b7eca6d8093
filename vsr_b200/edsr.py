"""EDSRNet on the tap-GEMM kernels (reference: src/model/nets/edsr_net.py:8-67).

head 3x3 (first-layer kernel, no activation) -> B residual blocks [3x3 + ReLU, 3x3, x res_scale, + x]
(the scaled residual add is the epilogue `v*out_scale + residual`) -> 3x3 + head skip -> tail: 3x3 to
4F / 9F channels stored phase-major (= nn.PixelShuffle by reinterpretation) ... -> last-layer kernel.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _flat
from ._lib import EPI_BIAS, EPI_RELU, EPI_RELU_BWD, EPI_RES_PRE, EPI_SCALE
from .drf_plan import DrfPlan, Layer, phase_table
from .nets import _PRECISIONS, _TC_LAYOUT, BaseNet, pack_weights, packed_weight_state
from .ops import TapTable


class EdsrPlan(DrfPlan):
    """Reuses DrfPlan's packing machinery (slab index maps, out-level tables, un-pack passes)."""

    def __init__(self, in_channels, out_channels, B, F, r, bf16):
        if r not in (2, 3, 4, 8):
            raise NotImplementedError
        self.variant, self.B = "edsr", B
        self.cin, self.cout, self.F, self.G, self.r, self.bf16 = in_channels, out_channels, F, 0, r, bf16
        from .drf_plan import _kc
        self.kc = _kc(F, bf16)
        self.kb = F // self.kc
        self.phases = phase_table(r)
        self.slot_of = {yx: i for i, yx in enumerate(self.phases)}
        self.params, self.n_params, self.fwd, self.bwd = {}, 0, {}, {}
        self._declare_params()
        self._build_layers()
        self._finalize()

    def _declare_params(self):
        F, B = self.F, self.B
        conv = lambda p, o, i: (self._add_param(p + ".weight", (o, i, 3, 3)), self._add_param(p + ".bias", (o,)))
        conv("head.0", F, self.cin)
        for b in range(B):
            conv(f"body.{b}.body.conv1", F, F)
            conv(f"body.{b}.body.conv2", F, F)
        conv("body.conv", F, F)
        if self.r == 3:
            conv("tail.0.conv1", 9 * F, F)
            self.out_levels = 1
        else:
            self.out_levels = int(math.log2(self.r))
            for i in range(self.out_levels):
                conv(f"tail.0.conv{i + 1}", 4 * F, F)
        conv("tail.conv", self.cout, F)
        self.last_name = "tail.conv"

    def _conv3x3(self, lname, wname):
        W = self._W(wname)
        F = self.F
        j, k = self._jk(F)
        taps, slabs, btaps, bslabs = [], [], [], []
        for ky in range(3):
            for kx in range(3):
                for b in range(self.kb):
                    taps.append((0, ky - 1, kx - 1, b * self.kc))
                    slabs.append(W.idx(j, b * self.kc + k, ky, kx))
                    btaps.append((0, -(ky - 1), -(kx - 1), b * self.kc))       # data-gradient: flipped taps
                    bslabs.append(W.idx(b * self.kc + k, j, ky, kx))
        self.fwd[lname] = Layer(lname, TapTable(self.kc, F, [(0, taps)]), slabs, F, self._bias_idx(wname))
        # reversed: row shifts ascending with the slab index, so the tensor-core kernel shares one taller A box per column
        self.bwd[lname] = Layer(lname, TapTable(self.kc, F, [(0, btaps[::-1])]), bslabs[::-1], F)

    def _build_layers(self):
        # order matters: all conv2 layers are contiguous in the packed gradient buffers (res_scale)
        for b in range(self.B):
            self._conv3x3(f"c1_{b}", f"body.{b}.body.conv1")
        for b in range(self.B):
            self._conv3x3(f"c2_{b}", f"body.{b}.body.conv2")
        self._conv3x3("cb", "body.conv")
        for lv in range(self.out_levels):
            self._out_level(lv, f"tail.0.conv{lv + 1}")


class _EdsrFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        net._pack(True)
        y, saved = net._forward(x.contiguous(), True)
        ctx.net, ctx.saved = net, saved
        return y

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        gflat = net._backward(ctx.saved, dy.contiguous())
        ctx.saved = None
        net.flat_grad = gflat
        pg = []
        for p in net._plan.params.values():
            n = int(np.prod(p.shape))
            pg.append(gflat[p.offset:p.offset + n].view(p.shape))
        return (None, None) + tuple(pg)


class _ResBlockParams(nn.Module):
    def __init__(self, F):
        super().__init__()
        self.body = nn.Sequential()
        self.body.add_module("conv1", nn.Conv2d(F, F, 3, padding=1))
        self.body.add_module("relu1", nn.ReLU())
        self.body.add_module("conv2", nn.Conv2d(F, F, 3, padding=1))


class EDSRNet(BaseNet):
    """Enhanced Deep Residual Network (reference: edsr_net.py:8-38).  Same constructor arguments
    (in_channels, out_channels, num_resblocks, num_features, upscale_factor, res_scale=0.1) and
    state_dict; forward(tensor [N,C,h,w]) -> tensor [N,C,r*h,r*w]; precision 'fp32' | 'bf16'."""

    def __init__(self, in_channels, out_channels, num_resblocks, num_features, upscale_factor, res_scale=0.1,
                 precision="fp32"):
        super().__init__()
        if precision not in _PRECISIONS:
            raise ValueError(f"precision should be one of {sorted(_PRECISIONS)}. Got {precision!r}.")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_resblocks, self.num_features = num_resblocks, num_features
        self.upscale_factor, self.res_scale, self.precision = upscale_factor, res_scale, precision
        F = num_features
        self.head = nn.Sequential(nn.Conv2d(in_channels, F, 3, padding=1))
        self.body = nn.Sequential(*[_ResBlockParams(F) for _ in range(num_resblocks)])
        self.body.add_module("conv", nn.Conv2d(F, F, 3, padding=1))
        up = nn.Sequential()
        if math.log(upscale_factor, 2) % 1 == 0:
            for i in range(int(math.log(upscale_factor, 2))):
                up.add_module(f"conv{i + 1}", nn.Conv2d(F, 4 * F, 3, padding=1))
                up.add_module(f"deconv{i + 1}", nn.PixelShuffle(2))
        elif upscale_factor == 3:
            up.add_module("conv1", nn.Conv2d(F, 9 * F, 3, padding=1))
            up.add_module("deconv1", nn.PixelShuffle(3))
        else:
            raise NotImplementedError
        self.tail = nn.Sequential(up)
        self.tail.add_module("conv", nn.Conv2d(F, out_channels, 3, padding=1))
        self._plan = EdsrPlan(in_channels, out_channels, num_resblocks, F, upscale_factor, precision in _TC_LAYOUT)
        assert [n for n, _ in self.named_parameters()] == list(self._plan.params), "parameter order differs"
        self._ops = None
        self._dev_state = None
        self.flat = self.flat_grad = None
        self._flatten()

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
        from .ops import cuda_ops, split_ops
        return split_ops() if self.precision in ("bf16x3", "tf32") else cuda_ops()

    def _state(self):
        if self._dev_state is None:
            P, dev = self._plan, self.flat.device
            act = torch.float64 if self.flat.dtype == torch.float64 else _PRECISIONS[self.precision]
            st = {"act": act,
                  "fwd_b": torch.empty(P.fwd_b_numel, dtype=self.flat.dtype, device=dev),
                  "fwd_b_idx": torch.from_numpy(P.fwd_b_idx).to(dev),
                  "unpack": [(lo, torch.from_numpy(i).to(dev)) for lo, i in P.unpack_passes], "ws": {}}
            st.update(packed_weight_state(self, P, dev, act))
            b = P.bias_unpack_idx
            nz = (b >= 0).nonzero()[0]
            st["bias_unpack"] = (int(nz.min()), torch.from_numpy(b[nz.min():nz.max() + 1].copy()).to(dev))
            self._dev_state = st
        return self._dev_state

    def _ws(self, key, nbytes):
        st = self._state()["ws"]
        n = (max(int(nbytes), 16) + 3) // 4
        if key not in st or st[key].numel() < n:
            st[key] = torch.empty(n, dtype=torch.float32, device=self.flat.device)
        return st[key]

    def _pview(self, flat, name):
        p = self._plan.params[name]
        return flat[p.offset:p.offset + int(np.prod(p.shape))].view(p.shape)

    def _pack(self, need_bwd):
        pack_weights(self, self._state(), need_bwd)

    def _conv(self, lname, src, out, epi=0, **kw):
        st, L = self._state(), self._plan.fwd[lname]
        self._backend().tapgemm(L.table, [src], out, st["fwd_w"][L.w_off:L.w_off + L.w_numel],
                                bias=st["fwd_b"][L.b_off:L.b_off + L.out_c], epi=EPI_BIAS | epi, **kw)

    def _dgrad(self, lname, src, out, epi=0, **kw):
        st, L = self._state(), self._plan.bwd[lname]
        self._backend().tapgemm(L.table, [src], out, st["bwd_w"][L.w_off:L.w_off + L.w_numel], epi=epi, **kw)

    def _forward(self, x, save):
        P, ops, st = self._plan, self._backend(), self._state()
        N, _, h, w = x.shape
        F, r = P.F, P.r
        new = lambda c: torch.empty(N, h, w, c, dtype=st["act"], device=x.device)
        head = new(F)
        ops.conv3x3_first(x, self._pview(self.flat, "head.0.weight"), self._pview(self.flat, "head.0.bias"), None, head)
        xs, ts = [head], []
        for b in range(P.B):
            t = new(F)
            self._conv(f"c1_{b}", xs[-1], t, epi=EPI_RELU)                       # edsr_net.py:46-47
            nxt = new(F)
            self._conv(f"c2_{b}", t, nxt, epi=EPI_SCALE | EPI_RES_PRE, out_scale=self.res_scale, residual=xs[-1])  # :50-52
            ts.append(t)
            xs.append(nxt)
        body = new(F)
        self._conv("cb", xs[-1], body, epi=EPI_RES_PRE, residual=head)           # :36
        s = [body]
        for lv in range(P.out_levels):
            L = P.fwd[f"out{lv + 1}"]
            nxt = new(L.out_c)
            self._conv(L.name, s[-1], nxt)
            s.append(nxt)
        y = torch.empty(N, P.cout, h * r, w * r, dtype=self.flat.dtype, device=x.device)
        ops.conv3x3_last(s[-1], r, F, P.phases, self._pview(self.flat, "tail.conv.weight"),
                         self._pview(self.flat, "tail.conv.bias"), y)
        return y, ((x, xs, ts, s) if save else None)

    def _backward(self, saved, dy):
        P, ops, st = self._plan, self._backend(), self._state()
        x, xs, ts, s = saved
        N, _, h, w = x.shape
        F, r = P.F, P.r
        dev, pd = x.device, self.flat.dtype
        new = lambda c: torch.empty(N, h, w, c, dtype=st["act"], device=dev)
        gflat = torch.zeros(P.n_params, dtype=pd, device=dev)
        dw = torch.zeros(P.fwd_w_numel, dtype=pd, device=dev)
        db = torch.zeros(P.fwd_b_numel, dtype=pd, device=dev)

        def wgrad(lname, src, dz):
            L = P.fwd[lname]
            ws = self._ws("wgrad", ops.tapgemm_wgrad_workspace(L.table, [src], dz))
            dbl = db[L.b_off:L.b_off + L.bias_c]
            if not ops.tapgemm_wgrad(L.table, [src], dz, dw[L.w_off:L.w_off + L.w_numel], True, ws, db=dbl,
                                     db_period=L.bias_c):
                rows = dz.numel() // L.bias_c
                ops.colsum(dz, rows, L.bias_c, dbl, True, self._ws("colsum", ops.colsum_workspace(rows, L.bias_c)))

        d_s = new(s[-1].shape[-1])
        ws = self._ws("last", ops.conv3x3_last_bwd_workspace(s[-1], r, F, P.cout))
        ops.conv3x3_last_bwd(s[-1], r, F, P.phases, self._pview(self.flat, "tail.conv.weight"), dy, d_s,
                             self._pview(gflat, "tail.conv.weight"), self._pview(gflat, "tail.conv.bias"), True, ws)
        for lv in reversed(range(P.out_levels)):
            lname = f"out{lv + 1}"
            wgrad(lname, s[lv], d_s)
            d_prev = new(s[lv].shape[-1])
            self._dgrad(lname, d_s, d_prev)
            d_s = d_prev
        d_bo = d_s                                        # gradient of body(head) + head
        wgrad("cb", xs[-1], d_bo)
        d_x = new(F)
        self._dgrad("cb", d_bo, d_x)
        for b in reversed(range(P.B)):
            wgrad(f"c2_{b}", ts[b], d_x)                  # x res_scale is applied to the packed buffers below
            dz1 = new(F)
            self._dgrad(f"c2_{b}", d_x, dz1, epi=EPI_SCALE | EPI_RELU_BWD, out_scale=self.res_scale, aux_y=ts[b])
            wgrad(f"c1_{b}", xs[b], dz1)
            d_prev = new(F)
            self._dgrad(f"c1_{b}", dz1, d_prev, epi=EPI_RES_PRE, residual=d_x)
            d_x = d_prev
        d_head = new(F)
        ops.add(d_x, d_bo, d_head)
        ws = self._ws("first", ops.conv3x3_first_bwd_workspace(x, F))
        ops.conv3x3_first_bwd(x, d_head, self._pview(gflat, "head.0.weight"), self._pview(gflat, "head.0.bias"), True, ws)
        if P.B > 0:     # res = body(x).mul(res_scale): conv2 weight / bias gradients carry the factor
            L0, L1 = P.fwd["c2_0"], P.fwd[f"c2_{P.B - 1}"]
            ops.scale_(dw[L0.w_off:L1.w_off + L1.w_numel], self.res_scale)
            ops.scale_(db[L0.b_off:L1.b_off + L1.out_c], self.res_scale)
        for lo, idx in st["unpack"]:
            ops.gather_add(dw, idx, gflat[lo:lo + idx.numel()])
        lo, idx = st["bias_unpack"]
        ops.gather_add(db, idx, gflat[lo:lo + idx.numel()])
        return gflat

    def forward(self, input):
        if input.dim() != 4 or input.shape[1] != self.in_channels:
            raise ValueError(f"expected input of shape [N,{self.in_channels},h,w], got {tuple(input.shape)}")
        if not self._is_flat():
            self._flatten()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _EdsrFunction.apply(self, input, *self.parameters())
        self._pack(False)
        return self._forward(input.contiguous(), False)[0]
