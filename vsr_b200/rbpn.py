"""RBPNet (Recurrent Back-Projection Network) on the tap-GEMM kernels — reference: src/model/nets/rbp_net.py:8-285
("same k x k stride-r conv / deconv shapes as DRFNet, PReLU with default init 0.25").

Same constructor arguments, forward I/O (list of `num_frames` frames [N,C,h,w] -> one frame [N,C,r*h,r*w], the centre
frame popped as the target like rbp_net.py:66-67) and state_dict (keys, shapes, default initialisation) as the reference.
Feature maps live in the phase-blocked pixel-major layout of the DRF nets, so

  * DeconvBlock / ConvBlock with the (k, s = r, p = 2) projection kernels are the `up` / `down` tap tables of DrfPlan with
    separate widths on the two sides (base_filter <-> feat);
  * the 3x3 convolutions of the residual blocks that act at HIGH resolution (res_feat2 / res_feat3) produce the four phase
    slots of a 2x2 block together (nt = 4 * feat: one A tile feeds 256 output columns, 9 of 16 source positions carry a
    kernel tap - the same table shape as the data gradient of the DRF output block);
  * every residual sum / difference of the DBPN stages is an epilogue of the convolution producing one of its operands:
    `h1 + h0`, `l1 + l0`, `h0 + e` are OUT2, `l0 - x`, `h0 - x`, `h0 - h1` are OUT2 with OUT2_SUB, ResnetBlock's
    `conv2(..) + x` is RES_PRE (rbp_net.py:84-87,241,274-275,284-285).

Backward: the forward pass records one entry per launch; the backward pass walks the record in reverse (the network is a
DAG with the weights shared by the num_frames - 1 neighbour iterations, so weight gradients accumulate): PReLU' +
slope-gradient partials (`vsr_act_bwd`), weight / bias gradient (`vsr_tapgemm_wgrad_bias`), data gradients by the
transposed tap tables, gradient sums by `vsr_axpby`.  Nothing runs on the CPU and no torch arithmetic op is launched.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _flat
from ._lib import EPI_BIAS, EPI_OUT2, EPI_OUT2_SUB, EPI_PRELU, EPI_RES_PRE
from .drf_plan import MAX_NT, PROJ, DrfPlan, Layer, _split_nt, phase_table
from .nets import _PRECISIONS, _TC_LAYOUT, BaseNet, pack_weights, packed_weight_state
from .ops import TapTable


# ---- parameter containers with the reference's module / attribute names (rbp_net.py:142-285) --------------------------
class ConvBlock(nn.Module):
    def __init__(self, cin, cout, k=3, s=1, p=1, activation="prelu"):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, p)
        if activation == "prelu":
            self.act = nn.PReLU()


class DeconvBlock(nn.Module):
    def __init__(self, cin, cout, k, s, p):
        super().__init__()
        self.deconv = nn.ConvTranspose2d(cin, cout, k, s, p)
        self.act = nn.PReLU()


class ResnetBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1)
        self.act = nn.PReLU()


class UpBlock(nn.Module):
    def __init__(self, c, k, s, p):
        super().__init__()
        self.up_conv1 = DeconvBlock(c, c, k, s, p)
        self.up_conv2 = ConvBlock(c, c, k, s, p)
        self.up_conv3 = DeconvBlock(c, c, k, s, p)


class DownBlock(nn.Module):
    def __init__(self, c, k, s, p):
        super().__init__()
        self.down_conv1 = ConvBlock(c, c, k, s, p)
        self.down_conv2 = DeconvBlock(c, c, k, s, p)
        self.down_conv3 = ConvBlock(c, c, k, s, p)


class DBPNet(nn.Module):
    def __init__(self, base_filter, feat, num_stages, r):
        super().__init__()
        k, s, p = PROJ[r]
        self.feat1 = ConvBlock(base_filter, feat, 1, 1, 0)
        self.up1 = UpBlock(feat, k, s, p)
        self.down1 = DownBlock(feat, k, s, p)
        self.up2 = UpBlock(feat, k, s, p)
        self.down2 = DownBlock(feat, k, s, p)
        self.up3 = UpBlock(feat, k, s, p)
        self.output = ConvBlock(num_stages * feat, feat, 1, 1, 0, activation=None)


class RbpPlan(DrfPlan):
    """Tap tables and packing maps of RBPNet; reuses DrfPlan's packing machinery (slab index maps, un-pack passes)."""

    def __init__(self, named_shapes, B, Fe, r, R, bf16):
        self.variant, self.B, self.Fe, self.F, self.G, self.r, self.R, self.bf16 = "rbp", B, Fe, Fe, 0, r, R, bf16
        if bf16:
            if B % 64 or Fe % 64:
                raise ValueError(f"bf16/tcgen05 mode needs base_filter % 64 == 0 and feat % 64 == 0 (got {B}, {Fe}); "
                                 "use precision='fp32'")
            self.kc = 64
        else:
            self.kc = math.gcd(B, Fe)
        self.kb = Fe // self.kc
        self.k, self.s, self.p = PROJ[r]
        self.phases = phase_table(r)
        self.slot_of = {yx: i for i, yx in enumerate(self.phases)}
        self.params, self.n_params, self.fwd, self.bwd = {}, 0, {}, {}
        for name, shape in named_shapes:
            self._add_param(name, shape)
        self._build_layers()
        self._finalize()

    def _slope(self, prefix):
        return self.params[prefix + ".act.weight"]

    # LR map with lr_c channels (GEMM K) -> phase-blocked HR map with hr_c channels per slot: transposed convolution
    # forward, strided convolution data gradient.  widx(a, b, ky, kx): a = LR-side channel, b = HR-side channel.
    def _up_g(self, lname, widx, store, lr_c, hr_c, slope=None, bias_name=None):
        s, p, k, kc = self.s, self.p, self.k, self.kc
        sig = {h: tuple(d for d in range(-4, 5) if 0 <= h + p - s * d < k) for h in range(s)}
        by_sig = {}
        for slot, (hy, wx) in enumerate(self.phases):
            by_sig.setdefault((sig[hy], sig[wx]), []).append(slot)
        groups, max_slots = [], max(1, MAX_NT // hr_c)
        for (sy, sx), slots in sorted(by_sig.items()):
            slots = sorted(slots)
            runs, cur = [], [slots[0]]
            for sl in slots[1:]:
                if sl == cur[-1] + 1 and len(cur) < max_slots:
                    cur.append(sl)
                else:
                    runs.append(cur); cur = [sl]
            runs.append(cur)
            groups += [(run, sy, sx) for run in runs]
        run_len = min(len(g[0]) for g in groups)
        norm = [(run[i:i + run_len], sy, sx) for run, sy, sx in groups for i in range(0, len(run), run_len)]
        assert all(len(run) == run_len for run, _, _ in norm)
        nt = run_len * hr_c
        jj, kk = np.arange(nt).reshape(nt, 1), np.arange(kc).reshape(1, kc)
        table_groups, slabs = [], []
        for run, sy, sx in norm:
            hy = np.array([self.phases[sl][0] for sl in run])[jj // hr_c]
            wx = np.array([self.phases[sl][1] for sl in run])[jj // hr_c]
            taps = []
            for dY in sy:
                for dX in sx:
                    for b in range(lr_c // kc):
                        taps.append((0, dY, dX, b * kc))
                        slabs.append(widx(b * kc + kk, jj % hr_c, hy + p - s * dY, wx + p - s * dX))
            table_groups.append((run[0] * hr_c, taps))
        bias = self._bias_idx(bias_name, len(self.phases) * hr_c, perm=lambda q: q % hr_c) if bias_name else None
        store[lname] = Layer(lname, TapTable(kc, nt, table_groups), slabs, len(self.phases) * hr_c, bias, slope=slope)

    # phase-blocked HR map (hr_c channels per slot, GEMM K) -> LR map with lr_c channels: strided convolution forward,
    # transposed convolution data gradient.  widx(a, b, ky, kx): a = LR-side channel, b = HR-side channel.
    def _down_g(self, lname, widx, store, lr_c, hr_c, slope=None, bias_name=None):
        s, p, k, kc = self.s, self.p, self.k, self.kc
        groups, slabs = [], []
        for (o0, nt) in _split_nt(lr_c):
            j, kk = self._jk(nt)
            taps = []
            for ky in range(k):
                dY, hy = divmod(ky - p, s)
                for kx in range(k):
                    dX, wx = divmod(kx - p, s)
                    slot = self.slot_of[(hy, wx)]
                    for b in range(hr_c // kc):
                        taps.append((0, dY, dX, slot * hr_c + b * kc))
                        slabs.append(widx(o0 + j, b * kc + kk, ky, kx))
            groups.append((o0, taps))
        bias = self._bias_idx(bias_name) if bias_name else None
        store[lname] = Layer(lname, TapTable(kc, _split_nt(lr_c)[0][1], groups), slabs, lr_c, bias, slope=slope)

    def _deconv(self, lname, prefix, cin, cout):
        WT = self._W(prefix + ".deconv")                   # ConvTranspose2d weight [Cin, Cout, k, k]
        f = lambda a, b, ky, kx: WT.idx(a, b, ky, kx)
        self._up_g(lname, f, self.fwd, cin, cout, slope=self._slope(prefix), bias_name=prefix + ".deconv")
        self._down_g(lname, f, self.bwd, cin, cout)

    def _sconv(self, lname, prefix, cin, cout):
        Wc = self._W(prefix + ".conv")                     # Conv2d weight [Cout, Cin, k, k]
        f = lambda a, b, ky, kx: Wc.idx(a, b, ky, kx)
        self._down_g(lname, f, self.fwd, cout, cin, slope=self._slope(prefix), bias_name=prefix + ".conv")
        self._up_g(lname, f, self.bwd, cout, cin)

    # 3x3 convolution at LOW resolution, cin -> cout
    def _lr3(self, lname, wname, cin, cout, slope=None):
        W, kc = self._W(wname), self.kc
        groups, slabs = [], []
        for (o0, nt) in _split_nt(cout):
            j, k = self._jk(nt)
            taps = []
            for ky in range(3):
                for kx in range(3):
                    for b in range(cin // kc):
                        taps.append((0, ky - 1, kx - 1, b * kc))
                        slabs.append(W.idx(o0 + j, b * kc + k, ky, kx))
            groups.append((o0, taps))
        self.fwd[lname] = Layer(lname, TapTable(kc, _split_nt(cout)[0][1], groups), slabs, cout, self._bias_idx(wname), slope=slope)
        groups, slabs = [], []
        for (o0, nt) in _split_nt(cin):
            j, k = self._jk(nt)
            taps = []
            for ky in (2, 1, 0):        # row shifts ascending with the slab index (shared A boxes in the tensor-core kernel)
                for kx in (2, 1, 0):
                    for b in range(cout // kc):
                        taps.append((0, -(ky - 1), -(kx - 1), b * kc))
                        slabs.append(W.idx(b * kc + k, o0 + j, ky, kx))
            groups.append((o0, taps))
        self.bwd[lname] = Layer(lname, TapTable(kc, _split_nt(cin)[0][1], groups), slabs, cin)

    # 3x3 convolution at HIGH resolution on the phase-blocked map, C -> C
    def _hr3(self, lname, wname, C, slope=None):
        W, kc, r = self._W(wname), self.kc, self.r
        n_slots = r * r
        slot = self.slot_of
        bias = self._bias_idx(wname, n_slots * C, perm=lambda q: q % C)
        if r in (2, 4, 8) and 4 * C <= MAX_NT:
            # the four slots of a 2x2 block together (nt = 4C): their outputs read the same 4x4 neighbourhood, each
            # with its own kernel offset (structural-zero rows where the offset leaves the 3x3 window)
            jj, k2 = self._jk(4 * C)
            oi, oj, ch = (jj // C) // 2, (jj // C) % 2, jj % C
            for fwd in (True, False):
                groups, slabs = [], []
                for b, (Py, Px) in enumerate(phase_table(r // 2)):
                    assert [self.phases[4 * b + 2 * i + j] for i in (0, 1) for j in (0, 1)] == \
                        [(2 * Py + i, 2 * Px + j) for i in (0, 1) for j in (0, 1)]
                    taps = []
                    rows = sorted(range(2 * Py - 1, 2 * Py + 3), key=lambda R: (R % r, R // r))
                    for Cc in range(2 * Px - 1, 2 * Px + 3):
                        dX, qx = divmod(Cc, r)
                        for bk in range(C // kc):
                            for R in rows:
                                dY, qy = divmod(R, r)
                                taps.append((0, dY, dX, slot[(qy, qx)] * C + bk * kc))
                                if fwd:     # y[P] = sum x[P + (ky-1, kx-1)] w[ky, kx]: source R = out + ky - 1
                                    ky, kx = R - (2 * Py + oi) + 1, Cc - (2 * Px + oj) + 1
                                    idx = W.idx(ch, bk * kc + k2, np.clip(ky, 0, 2), np.clip(kx, 0, 2))
                                else:       # dx[Q] = sum dy[Q - (ky-1, kx-1)] w[ky, kx]: source R = out - ky + 1
                                    ky, kx = (2 * Py + oi) - R + 1, (2 * Px + oj) - Cc + 1
                                    idx = W.idx(bk * kc + k2, ch, np.clip(ky, 0, 2), np.clip(kx, 0, 2))
                                ok = (ky >= 0) & (ky <= 2) & (kx >= 0) & (kx <= 2)
                                slabs.append(np.where(ok, idx, -1))
                    groups.append((4 * b * C, taps))
                tab = TapTable(kc, 4 * C, groups, useful=9.0 / 16.0)
                if fwd:
                    self.fwd[lname] = Layer(lname, tab, slabs, n_slots * C, bias, slope=slope)
                else:
                    self.bwd[lname] = Layer(lname, tab, slabs, n_slots * C)
            return
        j, k = self._jk(C)
        for fwd in (True, False):
            groups, slabs = [], []
            for sl, (py, px) in enumerate(self.phases):
                taps = []
                for ky in ((0, 1, 2) if fwd else (2, 1, 0)):
                    for kx in ((0, 1, 2) if fwd else (2, 1, 0)):
                        dY, qy = divmod(py + (ky - 1 if fwd else 1 - ky), r)
                        dX, qx = divmod(px + (kx - 1 if fwd else 1 - kx), r)
                        for b in range(C // kc):
                            taps.append((0, dY, dX, slot[(qy, qx)] * C + b * kc))
                            slabs.append(W.idx(j, b * kc + k, ky, kx) if fwd else W.idx(b * kc + k, j, ky, kx))
                groups.append((sl * C, taps))
            if fwd:
                self.fwd[lname] = Layer(lname, TapTable(kc, C, groups), slabs, n_slots * C, bias, slope=slope)
            else:
                self.bwd[lname] = Layer(lname, TapTable(kc, C, groups), slabs, n_slots * C)

    def _build_layers(self):
        B, Fe, R = self.B, self.Fe, self.R
        self._conv1x1_cat("dbp_f1", "dbp_net.feat1.conv", 1, src_c=B, slope=self._slope("dbp_net.feat1"))
        self._dgrad1x1("dbp_f1", [("dbp_net.feat1.conv", 0)], out_c=B)
        for blk, kinds in (("up1", "dsd"), ("down1", "sds"), ("up2", "dsd"), ("down2", "sds"), ("up3", "dsd")):
            stem = "up_conv" if blk.startswith("up") else "down_conv"
            for i, kind in enumerate(kinds):
                prefix = f"dbp_net.{blk}.{stem}{i + 1}"
                (self._deconv if kind == "d" else self._sconv)(f"{blk}_{i + 1}", prefix, Fe, Fe)
        self._conv1x1_cat("dbp_out", "dbp_net.output.conv", 3, src_c=Fe)
        for i in range(3):
            self._dgrad1x1(f"dbp_out@{i}", [("dbp_net.output.conv", i * Fe)], out_c=Fe)
        for i in range(R):
            sl = self.params[f"res_feat1.{i}.act.weight"]
            self._lr3(f"rf1_{i}_c1", f"res_feat1.{i}.conv1", B, B, slope=sl)
            self._lr3(f"rf1_{i}_c2", f"res_feat1.{i}.conv2", B, B, slope=sl)
        self._deconv("rf1_dc", f"res_feat1.{R}", B, Fe)
        for blk in ("res_feat2", "res_feat3"):
            short = "rf2" if blk == "res_feat2" else "rf3"
            for i in range(R):
                sl = self.params[f"{blk}.{i}.act.weight"]
                self._hr3(f"{short}_{i}_c1", f"{blk}.{i}.conv1", Fe, slope=sl)
                self._hr3(f"{short}_{i}_c2", f"{blk}.{i}.conv2", Fe, slope=sl)
        self._hr3("rf2_c", f"res_feat2.{R}.conv", Fe, slope=self._slope(f"res_feat2.{R}"))
        self._sconv("rf3_sc", f"res_feat3.{R}", Fe, B)


class _RbpFunction(torch.autograd.Function):
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


class RBPNet(BaseNet):
    """Recurrent Back-Projection Network (reference: rbp_net.py:8-91).  Args as the reference: in_channels, out_channels,
    base_filter, feat, num_stages (3: DBPNet concatenates three stages), num_resblocks, num_frames, upscale_factor;
    precision 'fp32' (CUDA-core strict mode) | 'bf16' (tcgen05 mode, base_filter % 64 == 0 and feat % 64 == 0) | 'bf16x3' /
    'tf32' (strict accuracy on the tensor cores, same channel constraint).
    forward(list of num_frames tensors [N,C,h,w]) -> tensor [N,out_channels,r*h,r*w].  out_channels must be 1 (the last
    convolution runs on the N = 1 kernels)."""

    def __init__(self, in_channels, out_channels, base_filter, feat, num_stages, num_resblocks, num_frames, upscale_factor,
                 precision="fp32"):
        super().__init__()
        if upscale_factor not in PROJ:
            raise ValueError(f"The upscale factor should be 2, 3, 4 or 8. Got {upscale_factor}.")
        if precision not in _PRECISIONS:
            raise ValueError(f"precision should be one of {sorted(_PRECISIONS)}. Got {precision!r}.")
        if num_stages != 3:
            raise ValueError("DBPNet concatenates exactly three stages (rbp_net.py:137): num_stages must be 3")
        if out_channels != 1:
            raise NotImplementedError("RBPNet: out_channels must be 1 (single-channel cine MRI)")
        self.in_channels, self.out_channels, self.base_filter, self.feat = in_channels, out_channels, base_filter, feat
        self.num_resblocks, self.num_frames, self.upscale_factor, self.precision = num_resblocks, num_frames, upscale_factor, precision
        self.t = num_frames // 2 if num_frames % 2 == 1 else num_frames // 2 - 1
        k, s, p = PROJ[upscale_factor]
        B, Fe, R = base_filter, feat, num_resblocks
        self.feat0 = ConvBlock(in_channels, B, 3, 1, 1)
        self.feat1 = ConvBlock(in_channels * 2, B, 3, 1, 1)
        self.dbp_net = DBPNet(B, Fe, num_stages, upscale_factor)
        self.res_feat1 = nn.Sequential(*[ResnetBlock(B) for _ in range(R)], DeconvBlock(B, Fe, k, s, p))
        self.res_feat2 = nn.Sequential(*[ResnetBlock(Fe) for _ in range(R)], ConvBlock(Fe, Fe, 3, 1, 1))
        self.res_feat3 = nn.Sequential(*[ResnetBlock(Fe) for _ in range(R)], ConvBlock(Fe, B, k, s, p))
        self.output = ConvBlock((num_frames - 1) * Fe, out_channels, 3, 1, 1, activation=None)
        self._plan = RbpPlan([(n, tuple(q.shape)) for n, q in self.named_parameters()], B, Fe, upscale_factor, R,
                             precision in _TC_LAYOUT)
        self._ops = None
        self._dev_state = None
        self.flat = self.flat_grad = None
        self._flatten()

    # ---- flat parameter bucket (same scheme as the other nets) ----
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

    def enable_sync_bn(self, process_group=None):
        """(MISRTrainStep calls this on every MISR net under data parallelism; RBPNet has no BatchNorm)"""

    def _state(self):
        if self._dev_state is None:
            self._dev_state = self._make_state(self._plan, None)
        return self._dev_state

    def _make_state(self, P, split):
        """packed weight / bias buffers and un-packing maps of plan `P` (split: None = by the precision, True / False =
        tripled bf16 slabs of the bf16x3 mode / plain slabs - nets that mix both kinds of layers keep one state per plan)"""
        if True:
            dev = self.flat.device
            act = torch.float64 if self.flat.dtype == torch.float64 else _PRECISIONS[self.precision]
            st = {"act": act,
                  "fwd_b": torch.empty(P.fwd_b_numel, dtype=self.flat.dtype, device=dev),
                  "fwd_b_idx": torch.from_numpy(P.fwd_b_idx).to(dev),
                  "unpack": [(lo, torch.from_numpy(i).to(dev)) for lo, i in P.unpack_passes], "ws": {}}
            st.update(packed_weight_state(self, P, dev, act, split))
            b = P.bias_unpack_idx
            nz = (b >= 0).nonzero()[0]
            st["bias_unpack"] = (int(nz.min()), torch.from_numpy(b[nz.min():nz.max() + 1].copy()).to(dev))
        return st

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

    # ---- forward: every launch is recorded (when `save`) for the reverse walk of _backward ----
    def _forward(self, frames, save):
        P, ops, st = self._plan, self._backend(), self._state()
        frames = list(frames)
        if len(frames) != self.num_frames:
            raise ValueError(f"expected {self.num_frames} frames, got {len(frames)}")
        x = frames.pop(self.t)                                                         # rbp_net.py:66-67
        N, _, h, w = x.shape
        B, Fe, r, R = P.B, P.Fe, P.r, P.R
        r2 = r * r
        act, dev = st["act"], x.device
        tape = [] if save else None
        lr = lambda c: torch.empty(N, h, w, c, dtype=act, device=dev)
        hr = lambda: torch.empty(N, h, w, r2 * Fe, dtype=act, device=dev)
        hv = lambda z: z.view(N, h, w * r2, Fe)                                       # an HR map as [pixels][Fe]
        slope_of = lambda pref: self.flat[pref.offset:pref.offset + 1]

        def conv(lname, srcs, out, residual=None, out2=None, res2=None, sub=False, views=None):
            """out = act(conv(srcs) + bias [+ residual]); out2 = out +/- res2"""
            L = P.fwd[lname]
            epi, kw = EPI_BIAS, {}
            if L.slope is not None:
                epi |= EPI_PRELU
                kw["slope"] = slope_of(L.slope)
            if residual is not None:
                epi |= EPI_RES_PRE
                kw["residual"] = residual
            if out2 is not None:
                epi |= EPI_OUT2 | (EPI_OUT2_SUB if sub else 0)
                kw.update(out2=out2, res2=res2)
            v = views or (lambda z: z)
            ops.tapgemm(L.table, [v(s) for s in srcs], v(out), st["fwd_w"][L.w_off:L.w_off + L.w_numel],
                        bias=st["fwd_b"][L.b_off:L.b_off + L.out_c], epi=epi,
                        **{k: (v(t) if k in ("residual", "out2", "res2") else t) for k, t in kw.items()})
            if save:
                tape.append(("conv", lname, srcs, out, residual, out2, res2, sub, views is not None))
            return out

        def first(prefix, xin):
            y = lr(B)
            ops.conv3x3_first(xin, self._pview(self.flat, prefix + ".conv.weight"), self._pview(self.flat, prefix + ".conv.bias"),
                              slope_of(P.params[prefix + ".act.weight"]), y)
            if save:
                tape.append(("first", prefix, xin, y))
            return y

        def resblock(short, i, xin, new):
            t = conv(f"{short}_{i}_c1", [xin], new())                                  # rbp_net.py:229-235
            return conv(f"{short}_{i}_c2", [t], new(), residual=xin)                  # :237-246 (same PReLU twice)

        def up(blk, a):                                                                # UpBlock, :268-276
            h0 = conv(f"{blk}_1", [a], hr())
            d = lr(Fe)
            conv(f"{blk}_2", [h0], lr(Fe), out2=d, res2=a, sub=True)                   # l0, d = l0 - x
            H = hr()
            conv(f"{blk}_3", [d], hr(), out2=H, res2=h0)                               # h1, H = h1 + h0
            return H

        def down(blk, Hin):                                                            # DownBlock, :278-285
            l0 = conv(f"{blk}_1", [Hin], lr(Fe))
            d = hr()
            conv(f"{blk}_2", [l0], hr(), out2=d, res2=Hin, sub=True)                   # h0, d = h0 - x
            Lo = lr(Fe)
            conv(f"{blk}_3", [d], lr(Fe), out2=Lo, res2=l0)                            # l1, L = l1 + l0
            return Lo

        feat_input = first("feat0", x.contiguous())                                    # :70
        feat_frame = [first("feat1", torch.cat([x, nb], dim=1).contiguous()) for nb in frames]   # :71-73
        Ht = []
        for j in range(len(frames)):                                                   # :77-86
            # res_feat1 first: its output h1 is an epilogue operand of the last DBPN convolution (e = h0 - h1)
            z = feat_frame[j]
            for i in range(R):
                z = resblock("rf1", i, z, lambda: lr(B))
            h1 = conv("rf1_dc", [z], hr())
            a = conv("dbp_f1", [feat_input], lr(Fe))                                   # DBPNet.forward, :129-139
            H1 = up("up1", a)
            H2 = up("up2", down("down1", H1))
            H3 = up("up3", down("down2", H2))
            h0, e = hr(), hr()
            conv("dbp_out", [H3, H2, H1], h0, out2=e, res2=h1, sub=True, views=hv)     # h0, e = h0 - h1  (:82)
            for i in range(R):
                e = resblock("rf2", i, e, hr)
            hsum = hr()
            conv("rf2_c", [e], hr(), out2=hsum, res2=h0)                               # e = res_feat2(e); h = h0 + e  (:83-84)
            Ht.append(hsum)
            if j + 1 < len(frames):                                                    # (the last feat_input is unused)
                z = hsum
                for i in range(R):
                    z = resblock("rf3", i, z, hr)
                feat_input = conv("rf3_sc", [z], lr(B))                                # :86
        # reconstruction: 3x3 convolution of the concatenation of the Ht = sum over j of 3x3 convolutions (:89-90)
        wv = self._pview(self.flat, "output.conv.weight").view(-1)
        bias = self._pview(self.flat, "output.conv.bias")
        y = torch.empty(N, 1, h * r, w * r, dtype=self.flat.dtype, device=dev)
        tmp = torch.empty_like(y) if len(Ht) > 1 else None
        for j, Hj in enumerate(Ht):
            wj = wv[j * Fe * 9:(j + 1) * Fe * 9].view(1, Fe, 3, 3)
            if j == 0:
                ops.conv3x3_last(Hj, r, Fe, P.phases, wj, bias, y)
            else:
                ops.conv3x3_last(Hj, r, Fe, P.phases, wj, torch.zeros_like(bias), tmp)
                ops.axpby(y, tmp, y, 1.0, 1.0)
        return y, ((tape, Ht, (N, h, w)) if save else None)

    # ---- backward: reverse walk of the record ----
    def _backward(self, saved, dy):
        P, ops, st = self._plan, self._backend(), self._state()
        tape, Ht, (N, h, w) = saved
        B, Fe, r = P.B, P.Fe, P.r
        r2 = r * r
        dev, pd = dy.device, self.flat.dtype
        gflat = torch.zeros(P.n_params, dtype=pd, device=dev)
        dw = torch.zeros(P.fwd_w_numel, dtype=pd, device=dev)
        db = torch.zeros(P.fwd_b_numel, dtype=pd, device=dev)
        n_act = sum(1 for rec in tape if rec[0] == "first" or (rec[0] == "conv" and P.fwd[rec[1]].slope is not None))
        partials = torch.zeros(max(n_act, 1), ops.partials_len, dtype=pd, device=dev)
        row_dst = []
        hv = lambda z: z.view(N, h, w * r2, Fe)
        G = {}                                  # data_ptr of a map -> [gradient (flat), owned]

        def acc(t, g, sign=1.0):
            k = t.data_ptr()
            g = g.reshape(-1)
            if k not in G:
                if sign == 1.0:
                    G[k] = [g, False]
                else:
                    n = torch.empty_like(g)
                    ops.axpby(g, None, n, sign, 0.0)
                    G[k] = [n, True]
            else:
                cur, owned = G[k]
                dst = cur if owned else torch.empty_like(cur)
                ops.axpby(cur, g, dst, 1.0, sign)
                G[k] = [dst, True]

        def act_bwd(g, y, pref):
            dz = torch.empty_like(y)
            i = len(row_dst)
            row_dst.append(pref.offset)
            ops.act_bwd(g.view(y.shape), y, dz, slope=self.flat[pref.offset:pref.offset + 1], slope_partials=partials[i])
            return dz

        # reconstruction convolution: d(Ht[j]), d(weight slice j); the bias gradient once
        wv, gwv = self._pview(self.flat, "output.conv.weight").view(-1), self._pview(gflat, "output.conv.weight").view(-1)
        gb = self._pview(gflat, "output.conv.bias")
        for j, Hj in enumerate(Ht):
            dH = torch.empty_like(Hj)
            ws = self._ws("last", ops.conv3x3_last_bwd_workspace(Hj, r, Fe, 1))
            ops.conv3x3_last_bwd(Hj, r, Fe, P.phases, wv[j * Fe * 9:(j + 1) * Fe * 9].view(1, Fe, 3, 3), dy, dH,
                                 gwv[j * Fe * 9:(j + 1) * Fe * 9].view(1, Fe, 3, 3), gb if j == 0 else torch.zeros_like(gb), True, ws)
            acc(Hj, dH)
        for rec in reversed(tape):
            if rec[0] == "first":
                _, prefix, xin, y = rec
                ent = G.pop(y.data_ptr(), None)
                if ent is None:
                    continue
                dz = act_bwd(ent[0], y, P.params[prefix + ".act.weight"])
                ws = self._ws("first", ops.conv3x3_first_bwd_workspace(xin, B))
                ops.conv3x3_first_bwd(xin, dz, self._pview(gflat, prefix + ".conv.weight"), self._pview(gflat, prefix + ".conv.bias"),
                                      True, ws)
                continue
            _, lname, srcs, out, residual, out2, res2, sub, viewed = rec
            L = P.fwd[lname]
            e1 = G.pop(out.data_ptr(), None)
            e2 = G.pop(out2.data_ptr(), None) if out2 is not None else None
            if e2 is not None:
                acc(res2, e2[0], -1.0 if sub else 1.0)
            if e1 is None and e2 is None:
                continue                        # (a map nobody consumed: the last h1 / l1 operands always are consumed)
            if e1 is None:
                g = e2[0]
            elif e2 is None:
                g = e1[0]
            else:
                g = e1[0] if e1[1] else torch.empty_like(e1[0])
                ops.axpby(e1[0], e2[0], g, 1.0, 1.0)
            dz = act_bwd(g, out, L.slope) if L.slope is not None else g.view(out.shape)
            if residual is not None:
                acc(residual, dz)
            v = hv if viewed else (lambda z: z)
            vs, vdz = [v(s) for s in srcs], v(dz)
            ws = self._ws("wgrad", ops.tapgemm_wgrad_workspace(L.table, vs, vdz))
            dbl = db[L.b_off:L.b_off + L.bias_c]
            if not ops.tapgemm_wgrad(L.table, vs, vdz, dw[L.w_off:L.w_off + L.w_numel], True, ws, db=dbl, db_period=L.bias_c):
                rows = dz.numel() // L.bias_c
                ops.colsum(dz, rows, L.bias_c, dbl, True, self._ws("colsum", ops.colsum_workspace(rows, L.bias_c)))
            for i, s in enumerate(srcs):
                Lb = P.bwd[lname if len(srcs) == 1 else f"{lname}@{i}"]
                ds = torch.empty_like(s)
                ops.tapgemm(Lb.table, [vdz], v(ds), st["bwd_w"][Lb.w_off:Lb.w_off + Lb.w_numel], epi=0)
                acc(s, ds)
        for lo, idx in st["unpack"]:
            ops.gather_add(dw, idx, gflat[lo:lo + idx.numel()])
        lo, idx = st["bias_unpack"]
        ops.gather_add(db, idx, gflat[lo:lo + idx.numel()])
        if row_dst:
            # (the launch sequence is a function of the architecture only: the destination table is uploaded by the first,
            #  eager step and reused - a host -> device copy is not capturable in the step's CUDA graph)
            rd = st.get("row_dst")
            if rd is None or rd.numel() != len(row_dst) or rd.device != dev:
                rd = st["row_dst"] = torch.tensor(row_dst, dtype=torch.int32, device=dev)
            ops.reduce_partials(partials, len(row_dst), rd, gflat)
        return gflat

    def forward(self, inputs):
        inputs = list(inputs)
        for f in inputs:
            if f.dim() != 4 or f.shape[1] != self.in_channels:
                raise ValueError(f"expected frames of shape [N,{self.in_channels},h,w], got {tuple(f.shape)}")
        if not self._is_flat():
            self._flatten()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _RbpFunction.apply(self, len(inputs), *inputs, *self.parameters())
        self._pack(False)
        return self._forward([f.contiguous() for f in inputs], False)[0]
