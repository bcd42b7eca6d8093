"""TOFlowNet (Task-Oriented Flow) on the tap-GEMM kernels — reference: src/model/nets/toflow_net.py:8-138 (SURVEY.md §8f
rank 4, after RBPNet and FRVSRNet: "SpyNet pyramid with 7x7 convolutions and BatchNorm2d, flow warp, bicubic resize").

Same constructor arguments, forward I/O (list of `num_frames` frames [N,1,h,w] -> ONE frame [N,1,rh,rw]) and state_dict
(keys, shapes, BatchNorm buffers, PyTorch's default initialisation in the reference's construction order) as the reference.
Strict fp32 mode only (`precision='fp32'`: the CUDA-core tap-GEMM; SpyNet's 4 / 16 / 32-channel levels do not fit the
64-channel tcgen05 tiles and the net is not on BASELINE's headline path).

Pipeline (csrc/toflow.cu for everything that is not a convolution or a BatchNorm):
  frames -> bicubic x r (one launch for all frames) -> padded to multiples of 16 with the batch minimum (device-side
  reduction, no host read-back) -> 2x2 average-pooling pyramid of all frames (three launches) -> per neighbour frame:
  SpyNet = 4 levels of {bilinear x2 of the flow (align_corners=True), warp of the neighbour + concatenation [ref, warped,
  flow] in ONE kernel, five 7x7 convolutions (49 taps x cin / 32 each) with BatchNorm2d (batch statistics, running buffers
  updated once per call like the module) + ReLU between them, flow update} -> warp of the full-resolution neighbour straight
  into its channel of the output block's input -> 9x9, 9x9, 1x1, 1x1 convolutions (ReLU in the epilogues) -> + reference
  frame, crop.
Layout: feature maps pixel-major [N, h, w, c] with the channel counts padded to multiples of 32 by structural zeros; images
and flows planar.  Backward: the forward pass records one entry per launch and the backward pass walks the record in
reverse; the up-sampled / padded frames are data (no parameter precedes them), so gradients stop at the warps' flow inputs.
"""
import numpy as np
import torch
import torch.nn as nn

from ._lib import EPI_BIAS, EPI_RELU
from .drf_plan import Layer, _split_nt
from .nets import BaseNet
from .ops import TapTable
from .rbpn import RbpPlan

_SPY = [(None, 32), (32, 64), (64, 32), (32, 16), (16, 2)]          # (cin, cout) of a SpyNet block's convolutions; first cin = 2 C + 2


def _spy_block(cin):
    mods, c = [], cin
    for i, (_, co) in enumerate(_SPY):
        mods.append(nn.Conv2d(c, co, kernel_size=7, stride=1, padding=3))
        if i < 4:
            mods += [nn.BatchNorm2d(co), nn.ReLU(inplace=True)]
        c = co
    m = nn.Module()
    m.block = nn.Sequential(*mods)
    return m


class ToflowPlan(RbpPlan):
    """Tap tables and packing maps of TOFlowNet; reuses the packing machinery of DrfPlan / RbpPlan."""
    KC = 32

    def __init__(self, named_shapes, T):
        self.variant, self.T, self.r, self.bf16 = "toflow", T, 1, False
        self.kc = self.KC
        self.F, self.Fe, self.B, self.G, self.R = 64, 64, 64, 0, 0
        self.params, self.n_params, self.fwd, self.bwd, self.act = {}, 0, {}, {}, {}
        for name, shape in named_shapes:
            self._add_param(name, shape)
        self._build_layers()
        self._finalize()

    @staticmethod
    def pad(c):
        return -(-c // ToflowPlan.KC) * ToflowPlan.KC

    # k x k convolution (stride 1, padding k // 2) on a pixel-major map; channels beyond cin / cout are structural zeros
    def _conv(self, lname, wname, cin, cout, k, act):
        W, kc, p = self._W(wname), self.kc, k // 2
        cin_pad, cout_pad = self.pad(cin), self.pad(cout)
        self.act[lname] = act
        groups, slabs = [], []
        for (o0, nt) in _split_nt(cout_pad):
            j, kk = self._jk(nt)
            taps = []
            for ky in range(k):
                for kx in range(k):
                    for b in range(cin_pad // kc):
                        taps.append((0, ky - p, kx - p, b * kc))
                        ok = ((o0 + j) < cout) & ((b * kc + kk) < cin)
                        slabs.append(np.where(ok, W.idx(np.minimum(o0 + j, cout - 1), np.minimum(b * kc + kk, cin - 1), ky, kx), -1))
            groups.append((o0, taps))
        bias = self._bias_idx(wname, cout_pad)
        bias = np.where(np.arange(cout_pad) < cout, bias, -1)
        self.fwd[lname] = Layer(lname, TapTable(kc, _split_nt(cout_pad)[0][1], groups), slabs, cout_pad, bias)
        groups, slabs = [], []
        for (o0, nt) in _split_nt(cin_pad):
            j, kk = self._jk(nt)
            taps = []
            for ky in reversed(range(k)):
                for kx in reversed(range(k)):
                    for b in range(cout_pad // kc):
                        taps.append((0, -(ky - p), -(kx - p), b * kc))
                        ok = ((b * kc + kk) < cout) & ((o0 + j) < cin)
                        slabs.append(np.where(ok, W.idx(np.minimum(b * kc + kk, cout - 1), np.minimum(o0 + j, cin - 1), ky, kx), -1))
            groups.append((o0, taps))
        self.bwd[lname] = Layer(lname, TapTable(kc, _split_nt(cin_pad)[0][1], groups), slabs, cin_pad)

    def _build_layers(self):
        for lv in range(4):
            cin = 4
            for i, (_, co) in enumerate(_SPY):
                self._conv(f"s{lv}_{i}", f"spy_net.blocks.{lv}.block.{3 * i}", cin, co, 7, None)
                cin = co
        self._conv("o0", "out_block.0", self.T, 64, 9, "relu")
        self._conv("o1", "out_block.2", 64, 64, 9, "relu")
        self._conv("o2", "out_block.4", 64, 64, 1, "relu")
        self._conv("o3", "out_block.6", 64, 1, 1, None)


class _ToflowFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, T, *args):
        net._pack(True)
        out, tape = net._forward([f.contiguous() for f in args[:T]], True)
        ctx.net, ctx.tape = net, tape
        return out

    @staticmethod
    def backward(ctx, grad):
        net = ctx.net
        gflat = net._backward(ctx.tape, grad)
        ctx.tape = None
        net.flat_grad = gflat
        pg = [gflat[p.offset:p.offset + int(np.prod(p.shape))].view(p.shape) for p in net._plan.params.values()]
        return (None, None) + (None,) * net.num_frames + tuple(pg)


class TOFlowNet(BaseNet):
    """Args as the reference (toflow_net.py:15): in_channels, out_channels, num_frames, upscale_factor; in_channels =
    out_channels = 1 (single-channel cine MRI).  forward(list of num_frames tensors [N,1,h,w]) -> [N,1,rh,rw]."""

    def __init__(self, in_channels, out_channels, num_frames, upscale_factor, precision="fp32"):
        super().__init__()
        if precision != "fp32":
            raise ValueError("TOFlowNet runs in the strict fp32 mode only (precision='fp32')")
        if in_channels != 1 or out_channels != 1:
            raise NotImplementedError("TOFlowNet: in_channels = out_channels = 1 (single-channel cine MRI)")
        if num_frames > ToflowPlan.KC:
            raise NotImplementedError(f"TOFlowNet: at most {ToflowPlan.KC} frames")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_frames, self.upscale_factor, self.precision = num_frames, upscale_factor, precision
        self.ref_idx = num_frames // 2 if num_frames % 2 == 1 else num_frames // 2 - 1      # toflow_net.py:21
        spy = nn.Module()
        spy.blocks = nn.ModuleList([_spy_block(2 * in_channels + 2) for _ in range(4)])     # :73
        self.spy_net = spy
        self.out_block = nn.Sequential(nn.Conv2d(in_channels * num_frames, 64, 9, 1, 4), nn.ReLU(inplace=True),
                                       nn.Conv2d(64, 64, 9, 1, 4), nn.ReLU(inplace=True),
                                       nn.Conv2d(64, 64, 1), nn.ReLU(inplace=True),
                                       nn.Conv2d(64, out_channels, 1))                       # :24-30
        self._plan = ToflowPlan([(n, tuple(q.shape)) for n, q in self.named_parameters()], num_frames)
        self._ops = None
        self._dev_state = None
        self.flat = self.flat_grad = None
        self._flatten()

    # the flat-bucket plumbing is RBPNet's
    from .rbpn import RBPNet as _R
    _flatten, _is_flat, _backend, _ws, _pview, _pack, _state, _make_state = (_R._flatten, _R._is_flat, _R._backend, _R._ws,
                                                                             _R._pview, _R._pack, _R._state, _R._make_state)
    del _R

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._flatten()
        return out

    def enable_sync_bn(self, process_group=None):
        """(MISRTrainStep calls this under data parallelism.)  TOFlowNet keeps rank-local BatchNorm2d statistics - what
        torch's DistributedDataParallel does with plain nn.BatchNorm2d; gradients are all-reduced like every other net's."""

    # ---- forward ----
    def _forward(self, frames, save):
        P, ops, st = self._plan, self._backend(), self._state()
        T, r = self.num_frames, self.upscale_factor
        N, _, h, w = frames[0].shape
        dev, act = frames[0].device, st["act"]
        tape = [] if save else None
        rec = (lambda *e: tape.append(e)) if save else (lambda *e: None)
        new = lambda *shape: torch.empty(*shape, dtype=act, device=dev)
        training = self.training

        # frames -> bicubic x r -> padded to multiples of 16 with the minimum over all frames (toflow_net.py:34-48)
        lr_all = torch.stack(frames).reshape(T * N, h, w)
        H0, W0 = h * r, w * r
        up = new(T, N, H0, W0)
        ops.upsample_bicubic(lr_all, r, up)
        hd = (16 - H0 % 16) % 16
        wd = (16 - W0 % 16) % 16
        y0, x0, H, W = hd // 2, wd // 2, H0 + hd, W0 + wd
        if hd or wd:
            partials = new(ops.partials_len)
            ops.min_partials(up, partials)
            xs = new(T, N, H, W)
            ops.pad_fill(up, y0, x0, partials, xs)
        else:
            xs = up
        pyr = [xs]                                                             # SpyNet.forward :76-78, all frames at once
        for _ in range(3):
            p = new(T, N, pyr[-1].shape[2] // 2, pyr[-1].shape[3] // 2)
            ops.avgpool2x2(pyr[-1], p)
            pyr.append(p)
        pyr = pyr[::-1]                                                        # coarsest first

        def conv(lname, src):
            L = P.fwd[lname]
            out = new(N, src.shape[1], src.shape[2], L.out_c)
            epi = EPI_BIAS | (EPI_RELU if P.act[lname] == "relu" else 0)
            ops.tapgemm(L.table, [src], out, st["fwd_w"][L.w_off:L.w_off + L.w_numel],
                        bias=st["fwd_b"][L.b_off:L.b_off + L.out_c], epi=epi)
            rec("conv", lname, src, out)
            return out

        def bn_relu(z, bn, pname, c):
            rows = z.numel() // z.shape[-1]
            ss, mr = new(2, z.shape[-1]), new(2, c)
            stats = None
            if training:
                stats = torch.zeros(1, 2, z.shape[-1], dtype=torch.float64, device=dev)
                ops.bn_stats(z, 0, c, 1, stats, 0, self._ws("stats", ops.bn_stats_workspace(1, rows, c)))
            ops.bn_finalize(stats, 0, 1, rows, c, self._pview(self.flat, pname + ".weight"), self._pview(self.flat, pname + ".bias"),
                            bn.eps, bn.momentum, bn.running_mean, bn.running_var, training, ss, mr)
            if training:
                bn.num_batches_tracked += 1
            a = new(*z.shape)
            ops.bn_relu(z, 0, c, ss, a)
            rec("bn", pname, z, a, ss, mr, c)
            return a

        yin = torch.zeros(N, H, W, P.pad(T), dtype=act, device=dev)            # input of the output block: one channel per frame
        for i in range(T):
            if i == self.ref_idx:
                ops.warp_cat(yin, i, xs[i], 0, None, None, 1.0, -1)             # x_warped[ref] = x_ref (:55)
                continue
            flow = torch.zeros(N, 2, H // 16, W // 16, dtype=act, device=dev)  # :80
            for lv in range(4):
                hl, wl = pyr[lv].shape[2:]
                fup = new(N, 2, hl, wl)
                ops.upsample_linear(flow, fup, True)                           # :82 (the factor 2.0 rides in `scale` below)
                rec("flowup", flow, fup)
                sin = torch.zeros(N, hl, wl, P.pad(4), dtype=act, device=dev)
                ops.warp_cat(sin, 0, pyr[lv][self.ref_idx], 1, pyr[lv][i], fup, 2.0, 2)       # :83-85
                rec("warpcat", sin, 1, pyr[lv][i], fup, 2.0, 2)
                z = sin
                blk = self.spy_net.blocks[lv].block
                for k in range(5):
                    z = conv(f"s{lv}_{k}", z)
                    if k < 4:
                        z = bn_relu(z, blk[3 * k + 1], f"spy_net.blocks.{lv}.block.{3 * k + 1}", _SPY[k][1])
                flow = new(N, 2, hl, wl)
                ops.flow_add(z, fup, 2.0, flow)                                # :83
                rec("flowadd", z, fup, 2.0, flow)
            ops.warp_cat(yin, 0, None, i, xs[i], flow, 1.0, -1)                # :60-61
            rec("warpcat", yin, i, xs[i], flow, 1.0, -1)
        z = conv("o3", conv("o2", conv("o1", conv("o0", yin))))                # :63
        out = torch.empty(N, 1, H0, W0, dtype=self.flat.dtype, device=dev)
        ops.head_add(z, xs[self.ref_idx], y0, x0, out)                         # :63-65
        rec("head", z, out, y0, x0)
        return out, tape

    # ---- backward: reverse walk of the record ----
    def _backward(self, tape, d_out):
        P, ops, st = self._plan, self._backend(), self._state()
        if not self.training:
            raise NotImplementedError("TOFlowNet: backward through BatchNorm2d in eval mode is not implemented")
        dev, pd = d_out.device, self.flat.dtype
        gflat = torch.zeros(P.n_params, dtype=pd, device=dev)
        dw = torch.zeros(P.fwd_w_numel, dtype=pd, device=dev)
        db = torch.zeros(P.fwd_b_numel, dtype=pd, device=dev)
        G = {}

        def acc(t, g):
            k = t.data_ptr()
            if k not in G:
                G[k] = [g, False]
            else:
                cur, owned = G[k]
                dst = cur if owned else torch.empty_like(cur)
                ops.axpby(cur, g, dst, 1.0, 1.0)
                G[k] = [dst, True]

        for e in reversed(tape):
            kind = e[0]
            if kind == "head":
                _, z, out, y0, x0 = e
                dz = torch.empty_like(z)
                ops.planar_to_nhwc(d_out.contiguous(), y0, x0, dz)
                acc(z, dz)
            elif kind == "conv":
                _, lname, src, out = e
                ent = G.pop(out.data_ptr(), None)
                if ent is None:
                    continue
                L, g = P.fwd[lname], ent[0].view(out.shape)
                if P.act[lname] == "relu":
                    dz = torch.empty_like(out)
                    ops.act_bwd(g, out, dz)
                else:
                    dz = g
                ws = self._ws("wgrad", ops.tapgemm_wgrad_workspace(L.table, [src], dz))
                dbl = db[L.b_off:L.b_off + L.bias_c]
                if not ops.tapgemm_wgrad(L.table, [src], dz, dw[L.w_off:L.w_off + L.w_numel], True, ws, db=dbl, db_period=L.bias_c):
                    rows = dz.numel() // L.bias_c
                    ops.colsum(dz, rows, L.bias_c, dbl, True, self._ws("colsum", ops.colsum_workspace(rows, L.bias_c)))
                Lb = P.bwd[lname]
                ds = torch.empty_like(src)
                ops.tapgemm(Lb.table, [dz], ds, st["bwd_w"][Lb.w_off:Lb.w_off + Lb.w_numel], epi=0)
                acc(src, ds)
            elif kind == "bn":
                _, pname, z, a, ss, mr, c = e
                ent = G.pop(a.data_ptr(), None)
                if ent is None:
                    continue
                pw = P.params[pname + ".weight"]                               # weight then bias: adjacent in the bucket
                rows = z.numel() // z.shape[-1]
                dz = torch.empty_like(z)
                # (a block's BatchNorm runs once per neighbour frame: its parameter gradients accumulate)
                gsl = torch.empty(2 * c, dtype=pd, device=dev)
                ops.bn_relu_bwd(ent[0].view(a.shape), z, 0, c, ss, mr, gsl, dz, 0, z.shape[-1], False,
                                self._ws("bnbwd", ops.bn_relu_bwd_workspace(rows, c)))
                dst = gflat[pw.offset:pw.offset + 2 * c]
                ops.axpby(dst, gsl, dst, 1.0, 1.0)
                acc(z, dz)
            elif kind == "flowadd":
                _, z, fup, scale, flow = e
                ent = G.pop(flow.data_ptr(), None)
                if ent is None:
                    continue
                g = ent[0].view(flow.shape)
                dz = torch.empty_like(z)
                ops.planar_to_nhwc(g, 0, 0, dz)
                acc(z, dz)
                gs = torch.empty_like(g)
                ops.axpby(g, None, gs, scale, 0.0)
                acc(fup, gs)
            elif kind == "warpcat":
                _, buf, c_w, nbr, flow, scale, c_flow = e
                ent = G.get(buf.data_ptr())                                    # (the output block's input has one entry per frame)
                if ent is None:
                    continue
                d = torch.empty_like(flow)
                ops.warp_cat_bwd(ent[0].view(buf.shape), c_w, nbr, flow, scale, c_flow, d)
                acc(flow, d)
            elif kind == "flowup":
                _, flow, fup = e
                ent = G.pop(fup.data_ptr(), None)
                if ent is None:
                    continue
                d = torch.empty_like(flow)
                ops.upsample_linear_bwd(ent[0].view(fup.shape), d, True)
                acc(flow, d)                                                   # (the first level's flow is the constant zero)
        for lo, idx in st["unpack"]:
            ops.gather_add(dw, idx, gflat[lo:lo + idx.numel()])
        lo, idx = st["bias_unpack"]
        ops.gather_add(db, idx, gflat[lo:lo + idx.numel()])
        return gflat

    def forward(self, inputs):
        inputs = list(inputs)
        if len(inputs) != self.num_frames:
            raise ValueError(f"expected {self.num_frames} frames, got {len(inputs)}")
        for f in inputs:
            if f.dim() != 4 or f.shape[1] != self.in_channels:
                raise ValueError(f"expected frames of shape [N,{self.in_channels},h,w], got {tuple(f.shape)}")
        if not self._is_flat():
            self._flatten()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _ToflowFunction.apply(self, self.num_frames, *inputs, *self.parameters())
        self._pack(False)
        out, _ = self._forward([f.contiguous() for f in inputs], False)
        return out
