"""Forward / backward schedule of DRFNet on the kernel set (reference: drf_net.py:38-49,118-133).

The engine owns the whole backward pass (it is one autograd node): data-gradients of the feature
lists are single tap-GEMMs over the *virtual concat of the consumers' gradients*, the PReLU
derivative and the scalar-slope gradient are fused into the epilogue of the kernel that produces
each gradient map, and weight gradients of all T frames accumulate into one packed fp32 buffer
that is un-packed into the flat gradient bucket once per step.

`ops` is vsr_b200.ops.CudaOps in the product (tests substitute the torch emulation to check this
schedule on a CPU-only box).
"""
import torch

from ._lib import EPI_BIAS, EPI_OUT2, EPI_PRELU, EPI_PRELU_BWD, EPI_RES_PRE
from .drf_plan import DrfPlan


def _stacked_view(ts):
    """[T, ...] view of T equally shaped tensors that are consecutive slices of one buffer (outputs of the batched
    output block, gradients written by the segmented loss kernel, batches of DeviceCineLoader), else None"""
    t0 = ts[0]
    step = t0.numel() * t0.element_size()
    if t0.is_contiguous() and all(t.shape == t0.shape and t.dtype == t0.dtype and t.is_contiguous() and
                                  t.data_ptr() == t0.data_ptr() + i * step for i, t in enumerate(ts)):
        room = (t0.untyped_storage().nbytes() - t0.storage_offset() * t0.element_size()) // max(step, 1)
        if room >= len(ts):
            return t0.as_strided((len(ts), *t0.shape), (t0.numel(), *t0.stride()))
    return None


def _as_stacked(ts):
    """[T, ...] tensor of T equally shaped tensors: a view when possible (_stacked_view), else one copy"""
    v = _stacked_view(ts)
    return v if v is not None else torch.stack([t.contiguous() for t in ts])


def shared_wgrad_sets(members, max_dz, max_srcs, max_acc):
    """members: [(layer name, ntaps)] of 1x1 convolutions on prefixes of ONE feature list, ntaps ascending (layer g reads the
    maps 0..ntaps-1).  Returns the contiguous partition into sets that one `wgrad_shared` launch can take (<= max_dz
    layers, <= max_srcs maps, sum of ceil(ntaps / 2) <= max_acc tensor-memory accumulators) reading the fewest maps:
    a set costs its largest ntaps (every source once) + one gradient map per layer."""
    n, best = len(members), None
    if n == 0:
        return []
    for cut in range(1 << (n - 1)):
        groups, cur = [], [members[0]]
        for i in range(1, n):
            if cut >> (i - 1) & 1:
                groups.append(cur)
                cur = []
            cur.append(members[i])
        groups.append(cur)
        if any(len(g_) > max_dz or g_[-1][1] > max_srcs or sum((nt + 1) // 2 for _, nt in g_) > max_acc for g_ in groups):
            continue
        cost = sum(g_[-1][1] + len(g_) for g_ in groups)
        if best is None or cost < best[0]:
            best = (cost, groups)
    return best[1] if best else []


class _Frame:
    __slots__ = ("x", "a1", "inn", "hidden", "lr", "u", "hr", "d", "f", "feat", "s", "y")


class DrfEngine:
    def __init__(self, plan: DrfPlan, ops, device, act_dtype, param_dtype=torch.float32):
        self.plan, self.ops, self.device = plan, ops, device
        self.act_dtype, self.param_dtype = act_dtype, param_dtype
        dev = device
        # precision='bf16x3' (ops.SplitOps): fp32 maps, bf16 weight slabs tripled as [wh | wh | wl] (DrfPlan.split_index)
        self.split = bool(getattr(ops, "split", False))
        self.wmul = 3 if self.split else 1
        w_dtype = torch.bfloat16 if self.split else act_dtype
        self.fwd_w = torch.empty(self.wmul * plan.fwd_w_numel, dtype=w_dtype, device=dev)
        self.bwd_w = torch.empty(self.wmul * plan.bwd_w_numel, dtype=w_dtype, device=dev)
        self.fwd_b = torch.empty(plan.fwd_b_numel, dtype=param_dtype, device=dev)
        self.fwd_w_idx = torch.from_numpy(plan.split_index("fwd") if self.split else plan.fwd_w_idx).to(dev)
        self.bwd_w_idx = torch.from_numpy(plan.split_index("bwd") if self.split else plan.bwd_w_idx).to(dev)
        self.fwd_b_idx = torch.from_numpy(plan.fwd_b_idx).to(dev)
        # un-packing maps (packed weight / bias gradients -> flat bucket), cut at the gradient-bucket boundaries
        self.buckets = []
        for blo, bhi, names in plan.grad_buckets(3):
            wparts = []
            for lo, idx in plan.unpack_passes:
                a, b_ = max(lo, blo), min(lo + len(idx), bhi)
                if a < b_ and (idx[a - lo:b_ - lo] >= 0).any():
                    wparts.append((a, torch.from_numpy(idx[a - lo:b_ - lo].copy()).to(dev)))
            b = plan.bias_unpack_idx[blo:bhi]
            nz = (b >= 0).nonzero()[0]
            bpart = (blo + int(nz.min()), torch.from_numpy(b[nz.min():nz.max() + 1].copy()).to(dev)) if len(nz) else None
            self.buckets.append((blo, bhi, names, wparts, bpart))
        self._ws = {}
        self._row_dst = {}
        self.flat = None

    # ---- buffers ---------------------------------------------------------------------------
    def _new(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=dtype or self.act_dtype, device=self.device)

    def _workspace(self, key, nbytes):
        nbytes = max(int(nbytes), 16)
        ws = self._ws.get(key)
        if ws is None or ws.numel() * 4 < nbytes:
            ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
            self._ws[key] = ws
        return ws

    def _pview(self, flat, name):
        p = self.plan.params[name]
        n = 1
        for d in p.shape:
            n *= d
        return flat[p.offset:p.offset + n].view(p.shape)

    def _slope(self, pref):
        return self.flat[pref.offset:pref.offset + 1]

    # ---- weights ---------------------------------------------------------------------------
    def pack(self, flat, need_bwd):
        """flat fp32 parameter bucket -> packed slabs (one gather kernel per buffer)."""
        self.flat = flat
        gather_w = self.ops.gather_split if self.split else self.ops.gather
        gather_w(flat, self.fwd_w_idx, self.fwd_w)
        self.ops.gather(flat, self.fwd_b_idx, self.fwd_b)
        if need_bwd:
            gather_w(flat, self.bwd_w_idx, self.bwd_w)

    def _fw(self, L):
        return self.fwd_w[self.wmul * L.w_off:self.wmul * (L.w_off + L.w_numel)]

    def _bw(self, L):
        return self.bwd_w[self.wmul * L.w_off:self.wmul * (L.w_off + L.w_numel)]

    def _fwd(self, lname, srcs, out, extra=0, **kw):
        L = self.plan.fwd[lname]
        epi = EPI_BIAS | extra
        slope = None
        if L.slope is not None:
            epi |= EPI_PRELU
            slope = self._slope(L.slope)
        self.ops.tapgemm(L.table, srcs, out, self._fw(L), bias=self.fwd_b[L.b_off:L.b_off + L.out_c],
                         epi=epi, slope=slope, **kw)

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, frames, save):
        """frames: list of T contiguous [N,Cin,h,w] tensors. Returns (outputs, saved frames)."""
        P, ops = self.plan, self.ops
        F, G, r = P.F, P.G, P.r
        r2 = r * r
        saved, outs, prev_f = [], [], None
        T = len(frames)
        # Saved activations of a layer live in ONE [T, N, h, w, c] buffer (frame t = slice t), so that the
        # weight gradient of the layer is a single launch over all T frames (backward()).
        uniform = T > 1 and all(f.shape == frames[0].shape for f in frames)
        stacked = save and uniform
        batch_io = uniform          # input / output blocks over all frames at once, also without saving (inference)
        bufs = {}

        def alloc(key, t, *shape):
            if not (stacked or (batch_io and key in ("a1", "inn", "s0"))):
                return self._new(*shape)
            b = bufs.get(key)
            if b is None:
                b = bufs[key] = self._new(T, *shape)
            return b[t]

        flat_tn = lambda b: b.view(b.shape[0] * b.shape[1], *b.shape[2:])
        batch_out = batch_io and P.variant == "drf"
        if batch_io:
            # the input block does not see the recurrence: one launch per layer for all T frames
            N, _, h, w = frames[0].shape
            x_all = bufs["x_all"] = torch.stack(frames)
            a1_all = bufs["a1"] = self._new(T, N, h, w, 4 * F)
            ops.conv3x3_first(flat_tn(x_all), self._pview(self.flat, f"{P.in_name}.conv1.weight"),
                              self._pview(self.flat, f"{P.in_name}.conv1.bias"),
                              self._slope(P.params[f"{P.in_name}.prelu1.weight"]), flat_tn(a1_all))
            inn_all = bufs["inn"] = self._new(T, N, h, w, F)
            self._fwd("in2", [flat_tn(a1_all)], flat_tn(inn_all))
        for t, x in enumerate(frames):
            N, _, h, w = x.shape
            S = _Frame()
            S.x = x
            S.a1 = alloc("a1", t, N, h, w, 4 * F)
            if not batch_io:
                ops.conv3x3_first(x, self._pview(self.flat, f"{P.in_name}.conv1.weight"),
                                  self._pview(self.flat, f"{P.in_name}.conv1.bias"),
                                  self._slope(P.params[f"{P.in_name}.prelu1.weight"]), S.a1)
            S.inn = alloc("inn", t, N, h, w, F)
            if not batch_io:
                self._fwd("in2", [S.a1], S.inn)
            S.hidden = S.inn if t == 0 else prev_f                      # drf_net.py:42-43
            S.lr, S.hr, S.u, S.d = [alloc("lr0", t, N, h, w, F)], [], [None], [None]
            self._fwd("fin", [S.inn, S.hidden], S.lr[0])
            hv = lambda z: z.view(N, h, w * r2, F)
            for g in range(G):
                if g == 0:
                    src = S.lr[0]
                else:
                    S.u.append(alloc(f"u{g}", t, N, h, w, F))
                    self._fwd(f"up{g}_c1", S.lr[:g + 1], S.u[g])
                    src = S.u[g]
                S.hr.append(alloc(f"hr{g}", t, N, h, w, r2 * F))
                self._fwd(f"up{g}_dc", [src], S.hr[g])
                if g == 0:
                    src = S.hr[0]
                else:
                    S.d.append(alloc(f"d{g}", t, N, h, w, r2 * F))
                    self._fwd(f"dn{g}_c1", [hv(z) for z in S.hr[:g + 1]], hv(S.d[g]))
                    src = S.d[g]
                S.lr.append(alloc(f"lr{g + 1}", t, N, h, w, F))
                self._fwd(f"dn{g}_sc", [src], S.lr[g + 1])
            if P.variant == "srfb":
                # srfb_net.py:44-48: residual = r_block(f);  output = bilinear(input) + residual
                S.f = alloc("f", t, N, h, w, F)
                S.feat = None
                self._fwd("fout", S.lr[1:], S.f)
                hr_out = alloc("s0", t, N, h, w, r2 * F)
                self._fwd("rdc", [S.f], hr_out)
                S.s = [hr_out]
                res = self._new(N, P.cout, h * r, w * r, dtype=self.param_dtype)
                ops.conv3x3_last(hr_out, r, F, P.phases, self._pview(self.flat, P.last_name + ".weight"),
                                 self._pview(self.flat, P.last_name + ".bias"), res)
                up = self._new(N, P.cin, h * r, w * r, dtype=self.param_dtype)
                ops.upsample_linear(x, up, False)
                y = self._new(N, P.cout, h * r, w * r, dtype=self.param_dtype)
                ops.add(up, res, y)
            else:
                S.f, S.feat = alloc("f", t, N, h, w, F), alloc("s0", t, N, h, w, F)
                self._fwd("fout", S.lr[1:], S.f, extra=EPI_OUT2, out2=S.feat, res2=S.inn)   # :46 global skip
                S.s = [S.feat]
                y = None
                if not batch_out:
                    for lv in range(P.out_levels):
                        L = P.fwd[f"out{lv + 1}"]
                        nxt = alloc(f"s{lv + 1}", t, N, h, w, L.out_c)
                        self._fwd(L.name, [S.s[-1]], nxt)
                        S.s.append(nxt)
                    y = self._new(N, P.cout, h * r, w * r, dtype=self.param_dtype)
                    ops.conv3x3_last(S.s[-1], r, F, P.phases, self._pview(self.flat, P.last_name + ".weight"),
                                     self._pview(self.flat, P.last_name + ".bias"), y)
            outs.append(y)
            prev_f = S.f
            if save:
                saved.append(S)
            else:
                S = None
        if batch_out:
            # neither does the output block (drf_net.py:136-147 acts on feat_t only): all T frames per launch
            N, _, h, w = frames[0].shape
            prev = bufs["s0"]
            for lv in range(P.out_levels):
                L = P.fwd[f"out{lv + 1}"]
                nxt = bufs[f"s{lv + 1}"] = self._new(T, N, h, w, L.out_c)
                self._fwd(L.name, [flat_tn(prev)], flat_tn(nxt))
                for t, S in enumerate(saved):          # (empty when nothing is saved: inference)
                    S.s.append(nxt[t])
                prev = nxt
            y_all = self._new(T * N, P.cout, h * r, w * r, dtype=self.param_dtype)
            ops.conv3x3_last(flat_tn(prev), r, F, P.phases, self._pview(self.flat, P.last_name + ".weight"),
                             self._pview(self.flat, P.last_name + ".bias"), y_all)
            outs = [y_all[t * N:(t + 1) * N] for t in range(T)]
        if stacked:
            # hidden state seen by frame t (drf_net.py:42-43) as one stacked tensor: [inn_0, f_0, ..., f_{T-2}]
            hid = self._new(T, *bufs["f"].shape[1:])
            hid[0].copy_(bufs["inn"][0])
            hid[1:].copy_(bufs["f"][:T - 1])
            bufs["hidden"] = hid
            saved[0].y = bufs          # the stacked buffers travel with the saved frames
        return outs, saved

    # ---- backward --------------------------------------------------------------------------
    def backward(self, saved, d_outs, on_bucket=None):
        """d_outs: list of T [N,Cout,rh,rw] gradients (None = zero). Returns the flat gradient.
        `on_bucket(gflat, lo, hi)` is called as soon as gflat[lo:hi] is final (ranges of DrfPlan.grad_buckets, in
        order): the data-parallel trainer starts that range's all-reduce while the next range is computed."""
        P, ops = self.plan, self.ops
        F, G, r = P.F, P.G, P.r
        r2 = r * r
        dev, pd = self.device, self.param_dtype
        gflat = torch.zeros(P.n_params, dtype=pd, device=dev)
        dw_packed = torch.zeros(P.fwd_w_numel, dtype=pd, device=dev)
        db_packed = torch.zeros(P.fwd_b_numel, dtype=pd, device=dev)
        T = len(saved)
        n_rows = T * (4 * G + 4)
        partials = torch.zeros(n_rows, ops.partials_len, dtype=pd, device=dev)
        row_dst, row = [], [0]

        def prow(pref):
            i = row[0]
            row[0] += 1
            row_dst.append(pref.offset)
            return partials[i]

        pending = {}     # layer -> [srcs, dz, workspace, slices used]: partials awaiting the per-step reduction
        bufs = getattr(saved[0], "y", None) if T > 1 else None     # stacked activations (forward())
        stacked = isinstance(bufs, dict)
        dzb = {}         # stacked gradient maps: key -> [T, N, h, w, c]
        deferred = {}    # layer -> (keys of the stacked sources, key of the stacked dz, view shape or None)

        def dz_alloc(key, t, *shape):
            """gradient map that a weight gradient consumes: frame t's slice of a [T, ...] buffer"""
            if not stacked:
                return self._new(*shape)
            b = dzb.get(key)
            if b is None:
                b = dzb[key] = self._new(T, *shape)
            return b[t]

        def wgrad(lname, srcs, dz, src_keys=None, dz_key=None, view=None):
            L = P.fwd[lname]
            if stacked and src_keys is not None:
                deferred[lname] = (src_keys, dz_key, view)      # one launch over all T frames after the loop
                return
            if T > 1:
                # same layer, same shapes every frame: accumulate split partials, reduce once (below)
                # every frame writes its own workspace slice; one reduction per layer per step (below)
                wsl = self._workspace("wg:" + lname, T * ops.tapgemm_wgrad_workspace(L.table, srcs, dz))
                used = pending[lname][3] if lname in pending else 0
                if ops.tapgemm_wgrad_partial(L.table, srcs, dz, wsl, used, T, L.bias_c):
                    pending[lname] = [srcs, dz, wsl, used + 1]
                    return
            ws = self._workspace("wgrad", ops.tapgemm_wgrad_workspace(L.table, srcs, dz))
            db = db_packed[L.b_off:L.b_off + L.bias_c]
            fused = ops.tapgemm_wgrad(L.table, srcs, dz, dw_packed[L.w_off:L.w_off + L.w_numel], True, ws,
                                      db=db, db_period=L.bias_c)
            if not fused:      # CUDA-core path / unsupported shape: separate column-sum kernels
                rows = dz.numel() // L.bias_c
                wsb = self._workspace("colsum", ops.colsum_workspace(rows, L.bias_c))
                ops.colsum(dz, rows, L.bias_c, db, True, wsb)

        def dgrad(lname, srcs, out, aux=None, slope_ref=None, residual=None):
            L = P.bwd[lname]
            epi, kw = 0, {}
            if residual is not None:
                epi |= EPI_RES_PRE
                kw["residual"] = residual
            if aux is not None:
                epi |= EPI_PRELU_BWD
                kw.update(aux_y=aux, slope=self._slope(slope_ref), slope_partials=prow(slope_ref))
            ops.tapgemm(L.table, srcs, out, self._bw(L), epi=epi, **kw)

        def act_bwd(dy, y, dz, slope_ref):
            ops.act_bwd(dy, y, dz, slope=self._slope(slope_ref), slope_partials=prow(slope_ref))

        flat_tn = lambda b: b.view(b.shape[0] * b.shape[1], *b.shape[2:])
        batch_io = stacked and "x_all" in bufs
        batch_out = batch_io and P.variant == "drf"
        if batch_out:
            # output block of all T frames at once (it does not see the recurrence): last conv, then the
            # data gradients of the out-level convolutions; their weight gradients are deferred like the rest
            S0 = saved[0]
            N, _, h, w = S0.x.shape
            n_lv = len(S0.s) - 1
            zero = None
            douts = []
            for d in d_outs:
                if d is None:
                    if zero is None:
                        zero = torch.zeros(N, P.cout, h * r, w * r, dtype=pd, device=dev)
                    d = zero
                douts.append(d)
            d_all = _as_stacked(douts).view(T * N, P.cout, h * r, w * r)
            d_s = dzb[f"ds{n_lv}"] = self._new(T, N, h, w, S0.s[-1].shape[-1])
            ws = self._workspace("last", ops.conv3x3_last_bwd_workspace(flat_tn(bufs[f"s{n_lv}"]), r, F, P.cout))
            ops.conv3x3_last_bwd(flat_tn(bufs[f"s{n_lv}"]), r, F, P.phases, self._pview(self.flat, P.last_name + ".weight"),
                                 d_all, flat_tn(d_s), self._pview(gflat, P.last_name + ".weight"),
                                 self._pview(gflat, P.last_name + ".bias"), True, ws)
            for lv in reversed(range(P.out_levels)):
                lname = f"out{lv + 1}"
                deferred[lname] = ([f"s{lv}"], f"ds{lv + 1}", None)
                d_prev = dzb[f"ds{lv}"] = self._new(T, N, h, w, S0.s[lv].shape[-1])
                dgrad(lname, [flat_tn(d_s)], flat_tn(d_prev))
                d_s = d_prev
        next_dz_lr0 = None
        for t in reversed(range(T)):
            S = saved[t]
            N, _, h, w = S.x.shape
            hv = lambda z: z.view(N, h, w * r2, F)
            new = lambda c=F: self._new(N, h, w, c)
            if batch_out:
                d_s = dzb["ds0"][t]
            else:
                d_out = d_outs[t]
                if d_out is None:
                    d_out = torch.zeros(N, P.cout, h * r, w * r, dtype=pd, device=dev)
                # ---- output block ----
                n_lv = len(S.s) - 1
                d_s = dz_alloc(f"ds{n_lv}", t, N, h, w, S.s[-1].shape[-1])
                ws = self._workspace("last", ops.conv3x3_last_bwd_workspace(S.s[-1], r, F, P.cout))
                ops.conv3x3_last_bwd(S.s[-1], r, F, P.phases, self._pview(self.flat, P.last_name + ".weight"),
                                     d_out.contiguous(), d_s, self._pview(gflat, P.last_name + ".weight"),
                                     self._pview(gflat, P.last_name + ".bias"), True, ws)
                for lv in reversed(range(P.out_levels)):
                    lname = f"out{lv + 1}"
                    wgrad(lname, [S.s[lv]], d_s, [f"s{lv}"], f"ds{lv + 1}")
                    d_prev = dz_alloc(f"ds{lv}", t, N, h, w, S.s[lv].shape[-1])
                    dgrad(lname, [d_s], d_prev)
                    d_s = d_prev
            if P.variant == "srfb":
                dz_r = dz_alloc("dz_r", t, N, h, w, r2 * F)
                act_bwd(d_s, S.s[0], dz_r, P.fwd["rdc"].slope)       # r_block.prelu1
                wgrad("rdc", [S.f], dz_r, ["f"], "dz_r")
                d_s = new()
                dgrad("rdc", [dz_r], d_s)
            d_feat = d_s
            # ---- feedback block output (+ hidden-state gradient of frame t+1: BPTT) ----
            dz_f = dz_alloc("dz_f", t, N, h, w, F)
            fout = P.fwd["fout"]
            if next_dz_lr0 is not None:
                dgrad("fin_hid", [next_dz_lr0], dz_f, aux=S.f, slope_ref=fout.slope, residual=d_feat)
            else:
                act_bwd(d_feat, S.f, dz_f, fout.slope)
            wgrad("fout", S.lr[1:], dz_f, [f"lr{j}" for j in range(1, G + 1)], "dz_f")
            dz_u, dz_d, p_hr0, p_lr0 = {}, {}, None, None
            for g in reversed(range(G)):
                j = g + 1
                srcs = [dz_f] + [dz_u[gg] for gg in range(max(j, 1), G)]
                dz_lr = dz_alloc(f"dz_lr{j}", t, N, h, w, F)
                dgrad(f"lr{j}", srcs, dz_lr, aux=S.lr[j], slope_ref=P.fwd[f"dn{g}_sc"].slope)
                # down-projection group g
                wgrad(f"dn{g}_sc", [S.hr[0] if g == 0 else S.d[g]], dz_lr, ["hr0" if g == 0 else f"d{g}"], f"dz_lr{j}")
                if g >= 1:
                    dz_d[g] = dz_alloc(f"dz_d{g}", t, N, h, w, r2 * F)
                    dgrad(f"dn{g}_sc", [dz_lr], dz_d[g], aux=S.d[g], slope_ref=P.fwd[f"dn{g}_c1"].slope)
                    wgrad(f"dn{g}_c1", [hv(z) for z in S.hr[:g + 1]], hv(dz_d[g]), [f"hr{gg}" for gg in range(g + 1)],
                          f"dz_d{g}", (h, w * r2, F))
                else:
                    p_hr0 = new(r2 * F)
                    dgrad("dn0_sc", [dz_lr], p_hr0)
                # gradient of hr_g
                cons = [dz_d[gg] for gg in range(max(g, 1), G)]
                dz_hr = dz_alloc(f"dz_hr{g}", t, N, h, w, r2 * F)
                up_slope = P.fwd[f"up{g}_dc"].slope
                if cons:
                    dgrad(f"hr{g}", [hv(c) for c in cons], hv(dz_hr), aux=hv(S.hr[g]), slope_ref=up_slope,
                          residual=hv(p_hr0) if g == 0 else None)
                else:
                    act_bwd(p_hr0, S.hr[g], dz_hr, up_slope)
                # up-projection group g
                wgrad(f"up{g}_dc", [S.lr[0] if g == 0 else S.u[g]], dz_hr, ["lr0" if g == 0 else f"u{g}"], f"dz_hr{g}")
                if g >= 1:
                    dz_u[g] = dz_alloc(f"dz_u{g}", t, N, h, w, F)
                    dgrad(f"up{g}_dc", [dz_hr], dz_u[g], aux=S.u[g], slope_ref=P.fwd[f"up{g}_c1"].slope)
                    wgrad(f"up{g}_c1", S.lr[:g + 1], dz_u[g], [f"lr{jj}" for jj in range(g + 1)], f"dz_u{g}")
                else:
                    p_lr0 = new()
                    dgrad("up0_dc", [dz_hr], p_lr0)
            # ---- feedback block input ----
            dz_lr0 = dz_alloc("dz_lr0", t, N, h, w, F)
            cons = [dz_u[gg] for gg in range(1, G)]
            fin = P.fwd["fin"]
            if cons:
                dgrad("lr0", cons, dz_lr0, aux=S.lr[0], slope_ref=fin.slope, residual=p_lr0)
            else:
                act_bwd(p_lr0, S.lr[0], dz_lr0, fin.slope)
            wgrad("fin", [S.inn, S.hidden], dz_lr0, ["inn", "hidden"], "dz_lr0")
            dz_in = dz_alloc("dz_in", t, N, h, w, F)
            in2 = P.fwd["in2"]
            skip = d_feat if P.variant == "drf" else None      # the feature skip exists in DRFNet only
            if t == 0:   # hidden == in_features at the first frame (drf_net.py:42-43)
                dgrad("fin_in0", [dz_lr0, dz_lr0], dz_in, aux=S.inn, slope_ref=in2.slope, residual=skip)
            else:
                dgrad("fin_in", [dz_lr0], dz_in, aux=S.inn, slope_ref=in2.slope, residual=skip)
            # ---- input block ----
            wgrad("in2", [S.a1], dz_in, ["a1"], "dz_in")
            if not batch_io:
                dz_a1 = new(4 * F)
                dgrad("in2", [dz_in], dz_a1, aux=S.a1, slope_ref=P.params[f"{P.in_name}.prelu1.weight"])
                ws = self._workspace("first", ops.conv3x3_first_bwd_workspace(S.x, 4 * F))
                ops.conv3x3_first_bwd(S.x, dz_a1, self._pview(gflat, f"{P.in_name}.conv1.weight"),
                                      self._pview(gflat, f"{P.in_name}.conv1.bias"), True, ws)
            next_dz_lr0 = dz_lr0
        if batch_io:
            # input block of all T frames at once: data gradient of conv2 (+ PReLU'), then the first conv
            x_all = flat_tn(bufs["x_all"])
            dz_a1 = self._new(*flat_tn(bufs["a1"]).shape)
            dgrad("in2", [flat_tn(dzb["dz_in"])], dz_a1, aux=flat_tn(bufs["a1"]),
                  slope_ref=P.params[f"{P.in_name}.prelu1.weight"])
            ws = self._workspace("first", ops.conv3x3_first_bwd_workspace(x_all, 4 * F))
            ops.conv3x3_first_bwd(x_all, dz_a1, self._pview(gflat, f"{P.in_name}.conv1.weight"),
                                  self._pview(gflat, f"{P.in_name}.conv1.bias"), True, ws)
        # PReLU slope gradients: every partial row exists by now (data gradients only)
        rd = self._row_dst.get(T)
        if rd is None or rd.numel() != len(row_dst):        # the launch sequence is a function of T only
            rd = torch.tensor(row_dst, dtype=torch.int32, device=dev)
            self._row_dst[T] = rd
        ops.reduce_partials(partials, row[0], rd, gflat)

        def deferred_wgrad(lname):
            # weight (+ bias) gradient of the layer over all T frames in one launch: [T, N, ...] -> [T*N, ...]
            src_keys, dz_key, view = deferred[lname]
            L = P.fwd[lname]

            def flat(b):
                b = b.view(b.shape[0] * b.shape[1], *b.shape[2:])
                return b if view is None else b.view(b.shape[0], *view)

            srcs, dz = [flat(bufs[k]) for k in src_keys], flat(dzb[dz_key])
            ws = self._workspace("wgrad", ops.tapgemm_wgrad_workspace(L.table, srcs, dz))
            db = db_packed[L.b_off:L.b_off + L.bias_c]
            fused = ops.tapgemm_wgrad(L.table, srcs, dz, dw_packed[L.w_off:L.w_off + L.w_numel], True, ws,
                                      db=db, db_period=L.bias_c)
            if not fused:
                rows = dz.numel() // L.bias_c
                wsb = self._workspace("colsum", ops.colsum_workspace(rows, L.bias_c))
                ops.colsum(dz, rows, L.bias_c, db, True, wsb)

        # Dense connections (drf_net.py:89-105): the 1x1 convolution of projection group g reads the concatenation of the
        # feature maps 0..g, so the separate weight gradients read map j once for every g >= j - 20 + 5 high-resolution
        # maps for the down-projection layers of a 6-group net.  Sets of those layers go through ONE pass over the maps
        # (ops.wgrad_shared: every source tile loaded once, <= 8 tensor-memory accumulators per launch); the sets are the
        # contiguous partition of g = 1..G-1 that reads the fewest maps.
        shared_done = {}

        shared_sets = lambda members: shared_wgrad_sets(members, ops.MAX_SHARED_DZ, ops.MAX_SHARED_SRCS, ops.MAX_SHARED_ACC)

        if stacked and hasattr(ops, "wgrad_shared"):
            for fam, src_key, dz_fmt, view in ((lambda g: f"dn{g}_c1", "hr", "dz_d{}", True), (lambda g: f"up{g}_c1", "lr", "dz_u{}", False)):
                members = [(fam(g), g + 1) for g in range(1, G) if fam(g) in deferred]
                if len(members) < 2:
                    continue
                for grp in shared_sets(members):
                    if len(grp) < 2:
                        continue
                    shp = None

                    def flat(b, shp_=None):
                        b = b.view(b.shape[0] * b.shape[1], *b.shape[2:])
                        return b.view(b.shape[0], b.shape[1], b.shape[2] * r2, F) if view else b

                    srcs = [flat(bufs[f"{src_key}{j}"]) for j in range(grp[-1][1])]
                    dzs = [flat(dzb[dz_fmt.format(nt - 1)]) for _, nt in grp]
                    ntaps = [nt for _, nt in grp]
                    if not ops.wgrad_shared_ok(srcs, dzs, ntaps):
                        continue
                    for lname, _ in grp:
                        shared_done[lname] = (grp, srcs, dzs, ntaps)

        def shared_wgrad(lname):
            grp, srcs, dzs, ntaps = shared_done[lname]
            layers = [P.fwd[n_] for n_, _ in grp]
            ops.wgrad_shared(srcs, dzs, ntaps, [dw_packed[L.w_off:L.w_off + L.w_numel] for L in layers],
                             [db_packed[L.b_off:L.b_off + L.bias_c] for L in layers], True,
                             lambda nbytes: self._workspace("wgshared", nbytes))
            for n_, _ in grp:
                shared_done[n_] = None               # the whole set is done

        for lname, (srcs, dz, wsl, used) in pending.items():
            L = P.fwd[lname]
            ops.tapgemm_wgrad_finish(L.table, srcs, dz, dw_packed[L.w_off:L.w_off + L.w_numel],
                                     db_packed[L.b_off:L.b_off + L.bias_c], L.bias_c, True, used, T, wsl)
        # ---- bucket by bucket: deferred weight gradients, un-pack (weights, biases), hand the range over ----
        for blo, bhi, names, wparts, bpart in self.buckets:
            for lname in names:
                if lname in shared_done:
                    if shared_done[lname] is not None:
                        shared_wgrad(lname)
                elif lname in deferred:
                    deferred_wgrad(lname)
            for lo, idx in wparts:
                ops.gather_add(dw_packed, idx, gflat[lo:lo + idx.numel()])
            if bpart is not None:
                ops.gather_add(db_packed, bpart[1], gflat[bpart[0]:bpart[0] + bpart[1].numel()])
            if on_bucket is not None:
                on_bucket(gflat, blo, bhi)
        return gflat
