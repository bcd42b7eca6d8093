"""FRVSRNet (Frame-Recurrent Video Super-Resolution) on the tap-GEMM kernels — reference: src/model/nets/frvsr_net.py:11-239
(after RBPNet: "FNet encoder-decoder with max pooling / bilinear up-sampling, STN warp, SRNet of
residual blocks and two stride-2 transposed convolutions").

Same constructor arguments, forward I/O (list of T frames [N,C,h,w] -> (sr_imgs, lr_imgs), or sr_imgs alone with
`is_prediction`) and state_dict (keys, shapes; Xavier-uniform convolution weights like :33-37) as the reference.  Strict
accuracy only: `precision='fp32'` runs every layer on the CUDA-core tap-GEMM; `precision='bf16x3'` (alias 'tf32') runs
SRNet's 64 -> 64 layers - 0.92 of the MACs - on the tcgen05 tap-GEMM as three bf16 products per product (fp32 maps,
ops.SplitOps); the flow net and SRNet's 17-channel head stay on the CUDA cores (the warp amplifies a flow error by the image
gradient times half the frame width: fp32 products are needed there to hold the 1e-4 output bar).

Layout: every feature map is pixel-major [N, h, w, c].  SRNet's two ConvTranspose2d(k=3, s=2, p=1, output_padding=1)
(:82-85) write the phase-blocked high-resolution layout of the other nets: the first one LR -> four 2x phase
slots, the second one 2x phase-blocked -> sixteen 4x slots in the nested order, so that the last 3x3 convolution is the
N = 1 kernel the other nets use.  The 17-channel SRNet input (space-to-depth of the warped previous output + the frame,
:47-48) and the 2-channel flow head are padded with structural zeros to a multiple of the K / N chunk.

Backward: the forward pass records one entry per launch; the backward pass walks the record in reverse (the warped images
are data or detached outputs, :47,53, so there is no gradient path between frames except through the shared weights).
Nothing runs on the CPU and no torch arithmetic op is launched on feature maps (torch.cat / F.pad act on the raw input
frames only).
"""
import numpy as np
import torch
import torch.nn as nn

from ._lib import EPI_BIAS, EPI_PRELU, EPI_RELU, EPI_RES_PRE
from .drf_plan import Layer, _split_nt, phase_table
from .nets import BaseNet
from .ops import TapTable
from .rbpn import RbpPlan

_LRELU = 0.2
# FNet: (name, cin, cout, what follows the pair of convolutions)
_FNET = [("1", 2, 32, "pool"), ("2", 32, 64, "pool"), ("3", 64, 128, "pool"), ("4", 128, 256, "up"), ("5", 256, 128, "up"),
         ("6", 128, 64, "up")]


class _ResBlockP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.body = nn.Sequential()
        self.body.add_module("conv1", nn.Conv2d(c, c, 3, padding=1))
        self.body.add_module("conv2", nn.Conv2d(c, c, 3, padding=1))


class _SRNetP(nn.Module):
    def __init__(self, cin, cout, r, R):
        super().__init__()
        self.head = nn.Sequential()
        self.head.add_module("conv", nn.Conv2d(cin * (r * r + 1), 64, 3, padding=1))
        self.body = nn.Sequential(*[_ResBlockP(64) for _ in range(R)])
        self.tail = nn.Sequential()
        self.tail.add_module("deconv1", nn.ConvTranspose2d(64, 64, 3, stride=2, padding=1, output_padding=1))
        self.tail.add_module("deconv2", nn.ConvTranspose2d(64, 64, 3, stride=2, padding=1, output_padding=1))
        self.tail.add_module("conv", nn.Conv2d(64, cout, 3, padding=1))


class _FNetP(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.body = nn.Sequential()
        for name, ci, co, _ in _FNET:
            self.body.add_module(f"conv{name}_1", nn.Conv2d(cin * 2 if name == "1" else ci, co, 3, padding=1))
            self.body.add_module(f"conv{name}_2", nn.Conv2d(co, co, 3, padding=1))
        self.tail = nn.Sequential()
        self.tail.add_module("conv1", nn.Conv2d(64, 32, 3, padding=1))
        self.tail.add_module("conv2", nn.Conv2d(32, 2, 3, padding=1))


class FrvsrPlan(RbpPlan):
    """Tap tables and packing maps of FRVSRNet; reuses the packing machinery of DrfPlan / RbpPlan."""
    KC = 32
    FLOW_PAD = 32       # output channels of the flow head (2 real; one K chunk for its data gradient)
    SIN_PAD = 32        # input channels of SRNet's head (r*r + 1 = 17 real)

    def __init__(self, named_shapes, R, subset=None, kc=None, bf16=False):
        """subset: None = every layer (strict fp32 mode); 'wide' = the layers with 64-channel blocks on both sides (the
        tensor-core tables of precision='bf16x3': kc = 64, bf16 slab layout), 'narrow' = the others (CUDA-core tables)"""
        self.variant, self.R, self.r, self.bf16 = "frvsr", R, 4, bf16
        self.subset = subset
        self.kc = kc or self.KC
        self.F, self.Fe, self.B, self.G = 64, 64, 64, 0
        self.phases = phase_table(4)
        self.slot_of = {yx: i for i, yx in enumerate(self.phases)}
        self.slot2 = {yx: i for i, yx in enumerate(phase_table(2))}
        self.params, self.n_params, self.fwd, self.bwd, self.act = {}, 0, {}, {}, {}
        for name, shape in named_shapes:
            self._add_param(name, shape)
        self._build_layers()
        self._finalize()

    # 3x3 convolution on a pixel-major map; channels beyond cin / cout (up to cin_pad / cout_pad) are structural zeros
    def _conv3(self, lname, wname, cin, cout, act, cin_pad=None, cout_pad=None, need_dgrad=True):
        cin_pad, cout_pad = cin_pad or cin, cout_pad or cout
        if not self._take(lname, cin_pad, cout_pad):
            return
        W, kc = self._W(wname), self.kc
        self.act[lname] = act
        groups, slabs = [], []
        for (o0, nt) in _split_nt(cout_pad):
            j, k = self._jk(nt)
            taps = []
            for ky in range(3):
                for kx in range(3):
                    for b in range(cin_pad // kc):
                        taps.append((0, ky - 1, kx - 1, b * kc))
                        ok = ((o0 + j) < cout) & ((b * kc + k) < cin)
                        slabs.append(np.where(ok, W.idx(np.minimum(o0 + j, cout - 1), np.minimum(b * kc + k, cin - 1), ky, kx), -1))
            groups.append((o0, taps))
        bias = self._bias_idx(wname, cout_pad)
        bias = np.where(np.arange(cout_pad) < cout, bias, -1)
        self.fwd[lname] = Layer(lname, TapTable(kc, _split_nt(cout_pad)[0][1], groups), slabs, cout_pad, bias)
        if not need_dgrad:
            return
        groups, slabs = [], []
        for (o0, nt) in _split_nt(cin_pad):
            j, k = self._jk(nt)
            taps = []
            for ky in (2, 1, 0):
                for kx in (2, 1, 0):
                    for b in range(cout_pad // kc if cout_pad >= kc else 1):
                        taps.append((0, -(ky - 1), -(kx - 1), b * kc))
                        ok = ((b * kc + k) < cout) & ((o0 + j) < cin)
                        slabs.append(np.where(ok, W.idx(np.minimum(b * kc + k, cout - 1), np.minimum(o0 + j, cin - 1), ky, kx), -1))
            groups.append((o0, taps))
        self.bwd[lname] = Layer(lname, TapTable(kc, _split_nt(cin_pad)[0][1], groups), slabs, cin_pad)

    # ConvTranspose2d(C, C, k=3, s=2, p=1, output_padding=1) from resolution level `lvl` (1 or 2) to 2 * lvl on phase-blocked
    # maps.  Along an axis output 2I + q reads input I with kernel tap 1 (q = 0), or I with tap 2 and I + 1 with tap 0 (q = 1).
    def _take(self, lname, ci, co):
        # SRNet only: the flow net keeps fp32 products - a flow error of 1e-5 (the 16 significant bits of a bf16 pair) is
        # multiplied by the image gradient and half the frame width in the warp and shows as 7e-4 in the next frame's output
        wide = lname.startswith("s_") and ci % 64 == 0 and co % 64 == 0
        return self.subset is None or (self.subset == "wide") == wide

    def _deconv2(self, lname, wname, C, lvl):
        if not self._take(lname, C, C):
            return
        WT, kc = self._W(wname), self.kc                     # [Cin, Cout, 3, 3]
        self.act[lname] = "relu"
        Q = {0: [(0, 1)], 1: [(0, 2), (1, 0)]}               # q -> [(input offset, kernel tap)]
        D = [(-1, 1, 0), (0, 0, 1), (0, 1, 2)]               # kernel tap k: input I receives from output 2 (I + dI) + q
        in_slot = {(0, 0): 0} if lvl == 1 else self.slot2
        out_phases = phase_table(2 * lvl)
        out_slot = {yx: i for i, yx in enumerate(out_phases)}
        j, k = self._jk(C)
        groups, slabs = [], []
        for sl, (Py, Px) in enumerate(out_phases):
            ay, qy, ax, qx = Py // 2, Py % 2, Px // 2, Px % 2
            taps = []
            for (diy, ky) in Q[qy]:
                dY, py = divmod(ay + diy, lvl)
                for (dix, kx) in Q[qx]:
                    dX, px = divmod(ax + dix, lvl)
                    for b in range(C // kc):
                        taps.append((0, dY, dX, in_slot[(py, px)] * C + b * kc))
                        slabs.append(WT.idx(b * kc + k, j, ky, kx))
            groups.append((sl * C, taps))
        n_out = len(out_phases) * C
        bias = self._bias_idx(wname, n_out, perm=lambda q: q % C)
        self.fwd[lname] = Layer(lname, TapTable(kc, C, groups), slabs, n_out, bias)
        # data gradient: the input slot (ay, ax) of level `lvl` gathers from the output positions 2 I - 1 + k
        groups, slabs = [], []
        for sl, (ay, ax) in enumerate(phase_table(lvl)):
            taps = []
            for (_, _, ky) in D:
                dY, Py = divmod(2 * ay - 1 + ky, 2 * lvl)
                for (_, _, kx) in D:
                    dX, Px = divmod(2 * ax - 1 + kx, 2 * lvl)
                    for b in range(C // kc):
                        taps.append((0, dY, dX, out_slot[(Py, Px)] * C + b * kc))
                        slabs.append(WT.idx(j, b * kc + k, ky, kx))
            groups.append((sl * C, taps))
        self.bwd[lname] = Layer(lname, TapTable(kc, C, groups), slabs, len(phase_table(lvl)) * C)

    def _build_layers(self):
        for name, ci, co, _ in _FNET:
            if name != "1":          # conv1_1 (2 -> 32) runs on the first-convolution kernel (K = 9 * Cin)
                self._conv3(f"f{name}_1", f"fnet.body.conv{name}_1", ci, co, "lrelu")
            self._conv3(f"f{name}_2", f"fnet.body.conv{name}_2", co, co, "lrelu")
        self._conv3("ft_1", "fnet.tail.conv1", 64, 32, "lrelu")
        self._conv3("ft_2", "fnet.tail.conv2", 32, 2, None, cout_pad=self.FLOW_PAD)
        self._conv3("s_head", "srnet.head.conv", 17, 64, "relu", cin_pad=self.SIN_PAD)
        for i in range(self.R):
            self._conv3(f"s_b{i}_1", f"srnet.body.{i}.body.conv1", 64, 64, "relu")
            self._conv3(f"s_b{i}_2", f"srnet.body.{i}.body.conv2", 64, 64, None)
        self._deconv2("s_d1", "srnet.tail.deconv1", 64, 1)
        self._deconv2("s_d2", "srnet.tail.deconv2", 64, 2)


class _FrvsrFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, T, *args):
        net._pack(True)
        sr, lr, saved = net._forward([f.contiguous() for f in args[:T]], True)
        ctx.net, ctx.saved, ctx.T = net, saved, T
        return tuple(sr) + tuple(lr)

    @staticmethod
    def backward(ctx, *grads):
        net, T = ctx.net, ctx.T
        gflat = net._backward(ctx.saved, grads[:T], grads[T:])
        ctx.saved = None
        net.flat_grad = gflat
        pg = [gflat[p.offset:p.offset + int(np.prod(p.shape))].view(p.shape) for p in net._plan.params.values()]
        return (None, None) + (None,) * T + tuple(pg)


class FRVSRNet(BaseNet):
    """Args as the reference (frvsr_net.py:25): in_channels, out_channels, upscale_factor, is_prediction, num_resblocks;
    in_channels = out_channels = 1 (single-channel cine MRI) and upscale_factor = 4 (SRNet's tail is two stride-2
    transposed convolutions whatever the factor, so the reference itself only runs at 4).
    forward(list of T tensors [N,1,h,w]) -> (list of T [N,1,4h,4w], list of T [N,1,h,w])."""

    def __init__(self, in_channels, out_channels, upscale_factor, is_prediction=False, num_resblocks=10, precision="fp32"):
        super().__init__()
        if precision not in ("fp32", "bf16x3", "tf32"):
            raise ValueError("FRVSRNet: precision 'fp32' (CUDA cores) or 'bf16x3' / 'tf32' (the strict bar with the 64-channel "
                             "layers on the tensor cores)")
        if in_channels != 1 or out_channels != 1:
            raise NotImplementedError("FRVSRNet: in_channels = out_channels = 1 (single-channel cine MRI)")
        if upscale_factor != 4:
            raise ValueError(f"SRNet up-scales by 2 x 2 (frvsr_net.py:82-85): the upscale factor should be 4. Got {upscale_factor}.")
        self.in_channels, self.out_channels, self.upscale_factor = in_channels, out_channels, upscale_factor
        self.is_prediction, self.num_resblocks, self.precision = is_prediction, num_resblocks, precision
        self.srnet = _SRNetP(in_channels, out_channels, upscale_factor, num_resblocks)
        self.fnet = _FNetP(in_channels)
        for m in self.modules():                                     # frvsr_net.py:33-37
            if m.__class__.__name__.find("Conv") != -1:
                nn.init.xavier_uniform_(m.weight)
        shapes = [(n, tuple(q.shape)) for n, q in self.named_parameters()]
        # precision='bf16x3': SRNet's residual blocks and transposed convolutions (64 -> 64: 0.92 of the MACs) run on the tensor
        # cores as three bf16 products per product (ops.SplitOps); SRNet's 17-channel head and the whole
        # flow net stay on the CUDA-core tap-GEMM (FrvsrPlan._take); maps are fp32 throughout
        self._hybrid = precision != "fp32"
        self._plan = FrvsrPlan(shapes, num_resblocks, subset="narrow" if self._hybrid else None)
        self._planB = FrvsrPlan(shapes, num_resblocks, subset="wide", kc=64, bf16=True) if self._hybrid else None
        self._ops = None
        self._dev_state = None
        self.flat = self.flat_grad = None
        self._flatten()

    # the flat-bucket plumbing is RBPNet's
    from .rbpn import RBPNet as _R
    _flatten, _is_flat, _backend, _ws, _pview, _make_state = _R._flatten, _R._is_flat, _R._backend, _R._ws, _R._pview, _R._make_state
    del _R

    def _pack(self, need_bwd):
        from .nets import pack_weights
        st = self._state()
        pack_weights(self, st, need_bwd)
        if self._hybrid:
            pack_weights(self, st["B"], need_bwd)

    def _of(self, lname):
        """(plan, state, tap-GEMM, weight gradient, its workspace size) of a layer: the tensor-core tables of the bf16x3 mode
        or the CUDA-core ones"""
        ops, st = self._backend(), self._state()
        if self._hybrid and lname in self._planB.fwd:
            return self._planB, st["B"], ops.tapgemm, ops.tapgemm_wgrad, ops.tapgemm_wgrad_workspace
        return (self._plan, st, getattr(ops, "tapgemm_plain", ops.tapgemm), getattr(ops, "tapgemm_wgrad_plain", ops.tapgemm_wgrad),
                getattr(ops, "tapgemm_wgrad_workspace_plain", ops.tapgemm_wgrad_workspace))

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._flatten()
        return out

    def _state(self):
        if self._dev_state is None:
            st = self._make_state(self._plan, False if self._hybrid else None)
            if self._hybrid:
                if self.flat.dtype != torch.float32 or not getattr(self._backend(), "split", False):
                    raise RuntimeError("FRVSRNet(precision='bf16x3') runs on the CUDA backend in fp32 storage only")
                st["B"] = self._make_state(self._planB, True)
            st["lrelu"] = torch.tensor([_LRELU], dtype=self.flat.dtype, device=self.flat.device)
            self._dev_state = st
        return self._dev_state

    # ---- forward ----
    def _forward(self, frames, save):
        P, ops, st = self._plan, self._backend(), self._state()
        N, _, h, w = frames[0].shape
        r, dev, act = 4, frames[0].device, st["act"]
        tape = [] if save else None
        new = lambda hh, ww, c: torch.empty(N, hh, ww, c, dtype=act, device=dev)
        rec = (lambda *e: tape.append(e)) if save else (lambda *e: None)

        def conv(lname, src, residual=None):
            P_, st_, tapgemm, _, _ = self._of(lname)
            L = P_.fwd[lname]
            a = P_.act[lname]
            out = new(src.shape[1], src.shape[2], L.out_c)
            epi, kw = EPI_BIAS, {}
            if a == "lrelu":
                epi |= EPI_PRELU
                kw["slope"] = st["lrelu"]
            elif a == "relu":
                epi |= EPI_RELU
            if residual is not None:
                epi |= EPI_RES_PRE
                kw["residual"] = residual
            tapgemm(L.table, [src], out, st_["fwd_w"][L.w_off:L.w_off + L.w_numel],
                    bias=st_["fwd_b"][L.b_off:L.b_off + L.out_c], epi=epi, **kw)
            rec("conv", lname, src, out, residual)
            return out

        def fnet(a, b):
            x = torch.cat([a, b], dim=1)                               # raw frames (frvsr_net.py:145)
            H, W = x.shape[-2:]
            y0 = x0 = 0
            if H % 8 != 0 or W % 8 != 0:                               # :149-156
                hd = 8 - H % 8 if H % 8 != 0 else 0
                wd = 8 - W % 8 if W % 8 != 0 else 0
                y0, x0 = hd // 2, wd // 2
                x = torch.nn.functional.pad(x, (wd // 2, wd - wd // 2, hd // 2, hd - hd // 2), value=float(x.min()))
            x = x.contiguous()
            Hp, Wp = x.shape[-2:]
            z = new(Hp, Wp, 32)
            ops.conv3x3_first(x, self._pview(self.flat, "fnet.body.conv1_1.weight"), self._pview(self.flat, "fnet.body.conv1_1.bias"),
                              st["lrelu"], z)
            rec("first", "fnet.body.conv1_1", x, z)
            for name, _, _, nxt in _FNET:
                if name != "1":
                    z = conv(f"f{name}_1", z)
                z = conv(f"f{name}_2", z)
                n_, hh, ww, c = z.shape
                if nxt == "pool":
                    y, idx = new(hh // 2, ww // 2, c), torch.empty(n_, hh // 2, ww // 2, c, dtype=torch.uint8, device=dev)
                    ops.maxpool2x2(z, y, idx)
                    rec("pool", z, y, idx)
                else:
                    y = new(2 * hh, 2 * ww, c)
                    ops.upsample2x_nhwc(z, y)
                    rec("up", z, y)
                z = y
            z = conv("ft_2", conv("ft_1", z))
            flow = torch.empty(N, 2, H, W, dtype=act, device=dev)
            ops.flow_tanh(z, y0, x0, flow)
            rec("tanh", z, flow, y0, x0)
            return flow

        def warp(img, flow):
            out = torch.empty_like(img)
            ops.grid_warp(img, flow, out)
            rec("warp", img, flow, out)
            return out

        sr_imgs, lr_imgs = [], []
        lr_last = frames[0]
        sr_last = torch.zeros(N, 1, h * r, w * r, dtype=act, device=dev)
        wl, bl = self._pview(self.flat, "srnet.tail.conv.weight"), self._pview(self.flat, "srnet.tail.conv.bias")
        for x in frames:
            lr_flow = fnet(lr_last, x)
            sr_flow = torch.empty(N, 2, h * r, w * r, dtype=act, device=dev)
            ops.upsample_linear(lr_flow, sr_flow, True)               # :46
            rec("flowup", lr_flow, sr_flow)
            warped = warp(sr_last, sr_flow)                            # :47 (sr_last detached: no gradient into it)
            sin = new(h, w, P.SIN_PAD)
            ops.s2d_cat(warped, x, r, sin)
            rec("s2d", warped, sin)
            z = conv("s_head", sin)
            for i in range(P.R):
                z = conv(f"s_b{i}_2", conv(f"s_b{i}_1", z), residual=z)
            z = conv("s_d2", conv("s_d1", z))
            sr = torch.empty(N, 1, h * r, w * r, dtype=self.flat.dtype, device=dev)
            ops.conv3x3_last(z, r, 64, P.phases, wl, bl, sr)
            rec("last", z, sr)
            sr_imgs.append(sr)
            sr_last = sr
            lr_imgs.append(warp(lr_last, lr_flow))                     # :53
            lr_last = x
        return sr_imgs, lr_imgs, tape

    # ---- backward: reverse walk of the record ----
    def _backward(self, tape, d_sr, d_lr):
        P, ops, st = self._plan, self._backend(), self._state()
        sr_outs = [e[2] for e in tape if e[0] == "last"]
        lr_outs = [e[3] for e in tape if e[0] == "warp"][1::2]         # per frame: the SR warp first, then the LR warp
        dev, pd = sr_outs[0].device, self.flat.dtype
        gflat = torch.zeros(P.n_params, dtype=pd, device=dev)
        plans = [(P, st)] + ([(self._planB, st["B"])] if self._hybrid else [])
        dwb = {id(p_): (torch.zeros(p_.fwd_w_numel, dtype=pd, device=dev), torch.zeros(p_.fwd_b_numel, dtype=pd, device=dev))
               for p_, _ in plans}
        scratch = torch.zeros(ops.partials_len, dtype=pd, device=dev)  # slope-gradient partials of the constant LeakyReLU
        G = {}

        def acc(t, g):
            k = t.data_ptr()
            g = g.reshape(-1)
            if k not in G:
                G[k] = [g, False]
            else:
                cur, owned = G[k]
                dst = cur if owned else torch.empty_like(cur)
                ops.axpby(cur, g, dst, 1.0, 1.0)
                G[k] = [dst, True]

        def act_bwd(g, y, a):
            if a is None:
                return g.view(y.shape)
            dz = torch.empty_like(y)
            if a == "lrelu":
                ops.act_bwd(g.view(y.shape), y, dz, slope=st["lrelu"], slope_partials=scratch)
            else:
                ops.act_bwd(g.view(y.shape), y, dz)
            return dz

        for outs, grads in ((sr_outs, d_sr), (lr_outs, d_lr)):
            for o, g in zip(outs, grads):
                if g is not None:
                    acc(o, g.contiguous())
        wl = self._pview(self.flat, "srnet.tail.conv.weight")
        gwl, gbl = self._pview(gflat, "srnet.tail.conv.weight"), self._pview(gflat, "srnet.tail.conv.bias")
        for e in reversed(tape):
            kind = e[0]
            if kind == "last":
                _, z, sr = e
                ent = G.pop(sr.data_ptr(), None)
                if ent is None:
                    continue
                dz = torch.empty_like(z)
                ws = self._ws("last", ops.conv3x3_last_bwd_workspace(z, 4, 64, 1))
                ops.conv3x3_last_bwd(z, 4, 64, P.phases, wl, ent[0].view(sr.shape), dz, gwl, gbl, True, ws)
                acc(z, dz)
            elif kind == "conv":
                _, lname, src, out, residual = e
                ent = G.pop(out.data_ptr(), None)
                if ent is None:
                    continue
                P_, st_, tapgemm, wgrad, wgrad_ws = self._of(lname)
                dw, db = dwb[id(P_)]
                L = P_.fwd[lname]
                dz = act_bwd(ent[0], out, P_.act[lname])
                if residual is not None:
                    acc(residual, dz)
                ws = self._ws("wgrad", wgrad_ws(L.table, [src], dz))
                dbl = db[L.b_off:L.b_off + L.bias_c]
                if not wgrad(L.table, [src], dz, dw[L.w_off:L.w_off + L.w_numel], True, ws, db=dbl, db_period=L.bias_c):
                    rows = dz.numel() // L.bias_c
                    ops.colsum(dz, rows, L.bias_c, dbl, True, self._ws("colsum", ops.colsum_workspace(rows, L.bias_c)))
                Lb = P_.bwd[lname]
                ds = torch.empty_like(src)
                tapgemm(Lb.table, [dz], ds, st_["bwd_w"][Lb.w_off:Lb.w_off + Lb.w_numel], epi=0)
                acc(src, ds)
            elif kind == "first":
                _, prefix, xin, y = e
                ent = G.pop(y.data_ptr(), None)
                if ent is None:
                    continue
                dz = act_bwd(ent[0], y, "lrelu")
                ws = self._ws("first", ops.conv3x3_first_bwd_workspace(xin, y.shape[-1]))
                ops.conv3x3_first_bwd(xin, dz, self._pview(gflat, prefix + ".weight"), self._pview(gflat, prefix + ".bias"), True, ws)
            elif kind == "pool":
                _, x, y, idx = e
                ent = G.pop(y.data_ptr(), None)
                if ent is not None:
                    dx = torch.empty_like(x)
                    ops.maxpool2x2_bwd(ent[0].view(y.shape), idx, dx)
                    acc(x, dx)
            elif kind == "up":
                _, x, y = e
                ent = G.pop(y.data_ptr(), None)
                if ent is not None:
                    dx = torch.empty_like(x)
                    ops.upsample2x_nhwc_bwd(ent[0].view(y.shape), dx)
                    acc(x, dx)
            elif kind == "tanh":
                _, z, flow, y0, x0 = e
                ent = G.pop(flow.data_ptr(), None)
                if ent is not None:
                    dz = torch.empty_like(z)
                    ops.flow_tanh_bwd(ent[0].view(flow.shape), flow, y0, x0, dz)
                    acc(z, dz)
            elif kind == "flowup":
                _, lr_flow, sr_flow = e
                ent = G.pop(sr_flow.data_ptr(), None)
                if ent is not None:
                    d = torch.empty_like(lr_flow)
                    ops.upsample_linear_bwd(ent[0].view(sr_flow.shape), d, True)
                    acc(lr_flow, d)
            elif kind == "warp":
                _, img, flow, out = e
                ent = G.pop(out.data_ptr(), None)
                if ent is not None:
                    d = torch.empty_like(flow)
                    ops.grid_warp_bwd(img, flow, ent[0].view(out.shape), d)
                    acc(flow, d)
            elif kind == "s2d":
                _, warped, sin = e
                ent = G.pop(sin.data_ptr(), None)
                if ent is not None:
                    d = torch.empty_like(warped)
                    ops.s2d_cat_bwd(ent[0].view(sin.shape), 4, d)
                    acc(warped, d)
        for p_, st_ in plans:
            dw, db = dwb[id(p_)]
            for lo, idx in st_["unpack"]:
                ops.gather_add(dw, idx, gflat[lo:lo + idx.numel()])
            lo, idx = st_["bias_unpack"]
            ops.gather_add(db, idx, gflat[lo:lo + idx.numel()])
        return gflat

    def forward(self, inputs):
        inputs = list(inputs)
        for f in inputs:
            if f.dim() != 4 or f.shape[1] != self.in_channels:
                raise ValueError(f"expected frames of shape [N,{self.in_channels},h,w], got {tuple(f.shape)}")
        if not self._is_flat():
            self._flatten()
        T = len(inputs)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            outs = _FrvsrFunction.apply(self, T, *inputs, *self.parameters())
            sr, lr = list(outs[:T]), list(outs[T:])
        else:
            self._pack(False)
            sr, lr, _ = self._forward([f.contiguous() for f in inputs], False)
        return sr if self.is_prediction else (sr, lr)
