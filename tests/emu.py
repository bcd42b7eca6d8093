"""TEST INFRASTRUCTURE — a plain-torch emulation of every op of vsr_b200.ops.CudaOps.

It exists so that (a) the host logic (tap tables, weight packing, forward/backward schedule) can
be verified on a CPU-only box against the oracle, and (b) each CUDA kernel can be checked against
an independent implementation on the GPU box.  Nothing in the product imports this file.
"""
import math

import torch
import torch.nn.functional as F

from vsr_b200 import _lib
from vsr_b200.ops import slab_index

E = _lib


def _shift(x, dy, dx):
    """y[n, i, j] = x[n, i+dy, j+dx], zero outside."""
    n, h, w, c = x.shape
    out = torch.zeros_like(x)
    ys0, ys1 = max(0, dy), min(h, h + dy)
    xs0, xs1 = max(0, dx), min(w, w + dx)
    if ys1 <= ys0 or xs1 <= xs0:
        return out
    out[:, ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx] = x[:, ys0:ys1, xs0:xs1]
    return out


def unswizzle_slabs(w, n_taps, nt):
    """bf16 swizzled slab image -> plain [n_taps, nt, 64]."""
    w = w.reshape(n_taps, nt * 64)
    j = torch.arange(nt).view(nt, 1)
    k = torch.arange(64).view(1, 64)
    idx = (j * 64 + (((k >> 3) ^ (j & 7)) << 3) + (k & 7)).reshape(-1).to(w.device)
    return w[:, idx].reshape(n_taps, nt, 64)


def swizzle_slabs(w_plain):
    """plain [n_taps, nt, 64] -> swizzled image (flat)."""
    n_taps, nt, kc = w_plain.shape
    assert kc == 64
    j = torch.arange(nt).view(nt, 1)
    k = torch.arange(64).view(1, 64)
    idx = (j * 64 + (((k >> 3) ^ (j & 7)) << 3) + (k & 7)).reshape(-1).to(w_plain.device)
    out = torch.empty(n_taps, nt * 64, dtype=w_plain.dtype, device=w_plain.device)
    out[:, idx] = w_plain.reshape(n_taps, nt * 64)
    return out.reshape(-1)


_INT = {torch.float32: torch.int32, torch.bfloat16: torch.int16, torch.float64: torch.int64}
PRELU_TINY = 2.0 ** -24


def prelu_fwd(v, slope, out_dtype):
    """PReLU as the kernels store it (csrc/common.cuh `Prelu`): slope 0 multiplies the non-positive side by 2^-24,
    a negative slope stores [x > 0] in the mantissa LSB of y.  Returns (stored y, y in v's precision)."""
    a = float(slope)
    fwd = a if a != 0.0 else PRELU_TINY
    y = torch.where(v > 0, v, fwd * v)
    st = y.to(out_dtype)
    if a < 0:
        bits = st.contiguous().view(_INT[out_dtype])
        st = ((bits & ~1) | (v > 0).to(bits.dtype)).view(out_dtype)
    return st, y


def prelu_bwd(g, y_stored, slope, ct):
    """(dL/dx, d(slope) contribution) from dL/dy and the stored y"""
    a = float(slope)
    fwd = a if a != 0.0 else PRELU_TINY
    if a < 0:
        pos = (y_stored.contiguous().view(_INT[y_stored.dtype]) & 1) != 0
    else:
        pos = y_stored > 0
    y = y_stored.to(ct)
    da = torch.where(pos, torch.zeros_like(g), g * (y * (1.0 / fwd))).sum()
    return torch.where(pos, g, a * g), da


class EmuOps:
    """Same interface as CudaOps; computes in fp32 (or fp64 if the tensors are fp64) with torch ops.
    `swizzled` says whether bf16 weight slabs are in the swizzled image (as the product packs them).
    """
    name = "emu"

    def __init__(self, partials_len=1024, swizzled_bf16=True):
        self.partials_len = partials_len
        self.swizzled_bf16 = swizzled_bf16
        self.launches = 0

    # ---- tap-GEMM ----------------------------------------------------------------------
    def _slabs(self, tab, w):
        if w.dtype == torch.bfloat16 and self.swizzled_bf16 and tab.kc == 64:
            return unswizzle_slabs(w, tab.n_taps_total, tab.nt)
        return w.reshape(tab.n_taps_total, tab.nt, tab.kc)

    def tapgemm(self, tab, srcs, out, w, bias=None, epi=0, out_scale=1.0, slope=None, residual=None,
                aux_y=None, out2=None, res2=None, slope_partials=None, force_simt=False):
        ct = torch.float64 if out.dtype == torch.float64 else torch.float32
        slabs = self._slabs(tab, w).to(ct)
        ti = 0
        for o0, taps in tab.groups:
            acc = torch.zeros(*out.shape[:3], tab.nt, dtype=ct, device=out.device)
            for (s, dy, dx, c0) in taps:
                x = _shift(srcs[s][..., c0:c0 + tab.kc].to(ct), dy, dx)
                acc += x @ slabs[ti].t()
                ti += 1
            sl = slice(o0, o0 + tab.nt)
            v = acc
            if epi & E.EPI_BIAS:
                v = v + bias[sl].to(ct)
            if epi & E.EPI_SCALE:
                v = v * out_scale
            if epi & E.EPI_RES_PRE:
                v = v + residual[..., sl].to(ct)
            if epi & E.EPI_RELU:
                v = v.clamp_min(0)
            if epi & E.EPI_PRELU_BWD:
                v, contrib = prelu_bwd(v, aux_y[..., sl], slope, ct)
                slope_partials.view(-1)[0] += contrib.to(slope_partials.dtype)
            elif epi & E.EPI_RELU_BWD:
                y = aux_y[..., sl].to(ct)
                v = torch.where(y > 0, v, torch.zeros_like(v))
            if epi & E.EPI_PRELU:
                stored, v = prelu_fwd(v, slope, out.dtype)
                out[..., sl] = stored
            else:
                out[..., sl] = v.to(out.dtype)
            if epi & E.EPI_OUT2:
                # the kernels add in fp32 from the unrounded value
                r2 = res2[..., sl].to(ct)
                out2[..., sl] = (v - r2 if epi & E.EPI_OUT2_SUB else v + r2).to(out2.dtype)
        self.launches += 1

    def tapgemm_wgrad_workspace(self, tab, srcs, dz):
        return 16

    def tapgemm_wgrad(self, tab, srcs, dz, dw, accumulate, workspace, db=None, db_period=0):
        ct = torch.float64 if dw.dtype == torch.float64 else torch.float32
        if db is not None:
            s = dz.to(ct).reshape(-1, dz.shape[-1] // db_period, db_period).sum((0, 1))
            if accumulate:
                db += s.to(db.dtype)
            else:
                db.copy_(s.to(db.dtype))
        res = torch.empty(tab.n_taps_total, tab.nt, tab.kc, dtype=ct, device=dz.device)
        ti = 0
        for o0, taps in tab.groups:
            g = dz[..., o0:o0 + tab.nt].to(ct).reshape(-1, tab.nt)
            for (s, dy, dx, c0) in taps:
                x = _shift(srcs[s][..., c0:c0 + tab.kc].to(ct), dy, dx).reshape(-1, tab.kc)
                res[ti] = g.t() @ x
                ti += 1
        flat = dw.view(-1)[:res.numel()]
        if accumulate:
            flat += res.reshape(-1).to(dw.dtype)
        else:
            flat.copy_(res.reshape(-1).to(dw.dtype))
        self.launches += 2
        return db is not None

    # ---- small kernels -----------------------------------------------------------------
    def colsum_workspace(self, rows, c):
        return 16

    def colsum(self, x, rows, c, db, accumulate, workspace):
        s = x.reshape(rows, c).to(db.dtype).sum(0)
        if accumulate:
            db += s
        else:
            db.copy_(s)
        self.launches += 2

    def conv3x3_first(self, x, w, bias, slope, y):
        ct = x.dtype
        z = F.conv2d(x, w.to(ct), bias.to(ct) if bias is not None else None, padding=1)
        if slope is not None:
            st, _ = prelu_fwd(z.permute(0, 2, 3, 1).contiguous(), slope, y.dtype)
            y.copy_(st)
            self.launches += 1
            return
        y.copy_(z.permute(0, 2, 3, 1).to(y.dtype))
        self.launches += 1

    def conv3x3_first_bwd_workspace(self, x, cout):
        return 16

    def conv3x3_first_bwd(self, x, dz, dw, db, accumulate, workspace):
        ct = dw.dtype
        g = dz.to(ct).permute(0, 3, 1, 2)
        gw = torch.nn.grad.conv2d_weight(x.to(ct), (dz.shape[-1], x.shape[1], 3, 3), g, padding=1)
        gb = g.sum((0, 2, 3))
        if accumulate:
            dw += gw
            db += gb
        else:
            dw.copy_(gw)
            db.copy_(gb)
        self.launches += 2

    @staticmethod
    def _unblock(x, r, c, phase_yx):
        """phase-blocked [n,h,w,r*r*c] -> NCHW [n,c,rh,rw]."""
        n, h, w, _ = x.shape
        xv = x.reshape(n, h, w, r * r, c)
        out = torch.empty(n, c, h * r, w * r, dtype=x.dtype, device=x.device)
        for s, (py, px) in enumerate(phase_yx):
            out[:, :, py::r, px::r] = xv[:, :, :, s].permute(0, 3, 1, 2)
        return out

    @staticmethod
    def _block(x, r, phase_yx):
        """NCHW [n,c,rh,rw] -> phase-blocked [n,h,w,r*r*c]."""
        n, c, H, W = x.shape
        h, w = H // r, W // r
        out = torch.empty(n, h, w, r * r, c, dtype=x.dtype, device=x.device)
        for s, (py, px) in enumerate(phase_yx):
            out[:, :, :, s] = x[:, :, py::r, px::r].permute(0, 2, 3, 1)
        return out.reshape(n, h, w, r * r * c)

    def conv3x3_last(self, x, r, c, phase_yx, w, bias, y):
        ct = y.dtype
        xi = self._unblock(x.to(ct), r, c, phase_yx)
        y.copy_(F.conv2d(xi, w.to(ct), bias.to(ct) if bias is not None else None, padding=1))
        self.launches += 1

    def conv3x3_last_bwd_workspace(self, x, r, c, cout):
        return 16

    def conv3x3_last_bwd(self, x, r, c, phase_yx, w, dy, dx, dw, db, accumulate, workspace):
        ct = dw.dtype
        xi = self._unblock(x.to(ct), r, c, phase_yx)
        gx = torch.nn.grad.conv2d_input(xi.shape, w.to(ct), dy.to(ct), padding=1)
        gw = torch.nn.grad.conv2d_weight(xi, w.shape, dy.to(ct), padding=1)
        gb = dy.to(ct).sum((0, 2, 3))
        dx.copy_(self._block(gx, r, phase_yx).to(dx.dtype))
        if accumulate:
            dw += gw
            db += gb
        else:
            dw.copy_(gw)
            db.copy_(gb)
        self.launches += 2

    def act_bwd(self, dy, y, dz, slope=None, slope_partials=None):
        ct = torch.float64 if dy.dtype == torch.float64 else torch.float32
        g = dy.to(ct)
        if slope is not None:
            gx, da = prelu_bwd(g, y, slope, ct)
            slope_partials.view(-1)[0] += da.to(slope_partials.dtype)
            dz.copy_(gx.to(dz.dtype))
        else:
            dz.copy_(torch.where(y.to(ct) > 0, g, torch.zeros_like(g)).to(dz.dtype))
        self.launches += 1

    def add(self, a, b, out):
        ct = torch.float64 if a.dtype == torch.float64 else torch.float32
        out.copy_((a.to(ct) + b.to(ct)).to(out.dtype))
        self.launches += 1

    def axpby(self, a, b, out, alpha, beta):
        ct = torch.float64 if a.dtype == torch.float64 else torch.float32
        v = alpha * a.to(ct)
        if b is not None and beta != 0:
            v = v + beta * b.to(ct)
        out.copy_(v.to(out.dtype))
        self.launches += 1

    def reduce_partials(self, partials, rows, row_dst, dst):
        sums = partials.reshape(-1, partials.shape[-1])[:rows].sum(1)
        for r in range(rows):
            dst[int(row_dst[r])] += sums[r].to(dst.dtype)
        self.launches += 1

    def gather(self, src, idx, dst):
        li = idx.long()
        vals = torch.where(li >= 0, src[li.clamp_min(0)], torch.zeros((), dtype=src.dtype, device=src.device))
        dst.view(-1)[:li.numel()].copy_(vals.to(dst.dtype))
        self.launches += 1

    def gather_add(self, src, idx, dst):
        li = idx.long()
        vals = torch.where(li >= 0, src[li.clamp_min(0)], torch.zeros((), dtype=src.dtype, device=src.device))
        dst.view(-1)[:li.numel()] += vals.to(dst.dtype)
        self.launches += 1

    def cast(self, src, dst):
        dst.copy_(src.to(dst.dtype))
        self.launches += 1

    def zero_(self, t):
        t.zero_()

    # ---- Conv3d network: BatchNorm3d + ReLU on channel windows, dynamic filter tail ---------
    def copy_window(self, src, c0_src, dst, c0_dst, c):
        dst[..., c0_dst:c0_dst + c] = src[..., c0_src:c0_src + c]
        self.launches += 1

    def bn_stats_workspace(self, frames, rows_per_frame, c):
        return 16

    def bn_stats(self, x, c0, c, frames, stats, s0, workspace):
        v = x[..., c0:c0 + c].reshape(frames, -1, c).double()
        stats[:, 0, s0:s0 + c] = v.sum(1)
        stats[:, 1, s0:s0 + c] = (v * v).sum(1)
        self.launches += 2

    def tshift_add(self, z, g, frames_in, t_pad, bias, out, c0_out, frames_out, stats, s0, workspace):
        ct = torch.float64 if z.dtype == torch.float64 else torch.float32
        zf = z.reshape(frames_in, -1, z.shape[-1]).to(ct)
        of = out.reshape(frames_out, -1, out.shape[-1])
        for f in range(frames_out):
            acc = torch.zeros(zf.shape[1], g, dtype=ct, device=z.device)
            if bias is not None:
                acc += bias.to(ct)
            for kt in range(3):
                fi = f + kt - t_pad
                if 0 <= fi < frames_in:
                    acc += zf[fi, :, kt * g:(kt + 1) * g]
            of[f, :, c0_out:c0_out + g] = acc.to(out.dtype)
        if stats is not None:
            self.bn_stats(out, c0_out, g, frames_out, stats, s0, None)
        self.launches += 2 if stats is not None else 1

    def tshift_gather(self, dy, c0, g, frames_out, t_pad, dz, frames_in):
        yf = dy.reshape(frames_out, -1, dy.shape[-1])
        zf = dz.reshape(frames_in, -1, dz.shape[-1])
        zf.zero_()
        for f in range(frames_in):
            for kt in range(3):
                fo = f - kt + t_pad
                if 0 <= fo < frames_out:
                    zf[f, :, kt * g:(kt + 1) * g] = yf[fo, :, c0:c0 + g]
        self.launches += 1

    def bn_finalize(self, stats, s0, frames, rows_per_frame, c, gamma, beta, eps, momentum, running_mean,
                    running_var, training, scale_shift, mean_rstd):
        if training:
            count = float(frames * rows_per_frame)
            s = stats[:frames, 0, s0:s0 + c].sum(0)
            q = stats[:frames, 1, s0:s0 + c].sum(0)
            mean = s / count
            var = (q / count - mean * mean).clamp_min(0)
            if running_mean is not None:
                unb = var * count / (count - 1) if count > 1 else var
                running_mean.copy_(((1 - momentum) * running_mean.double() + momentum * mean).to(running_mean.dtype))
                running_var.copy_(((1 - momentum) * running_var.double() + momentum * unb).to(running_var.dtype))
        else:
            mean, var = running_mean.double(), running_var.double()
        rstd = 1.0 / torch.sqrt(var + eps)
        sc = gamma.double() * rstd
        scale_shift.zero_()
        scale_shift[0, :c] = sc.to(scale_shift.dtype)
        scale_shift[1, :c] = (beta.double() - mean * sc).to(scale_shift.dtype)
        mean_rstd[0] = mean.to(mean_rstd.dtype)
        mean_rstd[1] = rstd.to(mean_rstd.dtype)
        self.launches += 1

    def bn_relu(self, x, c0, c, scale_shift, y):
        ct = torch.float64 if x.dtype == torch.float64 else torch.float32
        v = x[..., c0:c0 + c].to(ct) * scale_shift[0, :c].to(ct) + scale_shift[1, :c].to(ct)
        y.zero_()
        y[..., :c] = v.clamp_min(0).to(y.dtype)
        self.launches += 1

    def bn_relu_bwd_workspace(self, rows, c):
        return 16

    def bn_relu_bwd(self, dy, x, c0, c, scale_shift, mean_rstd, dgamma_dbeta, dx, c0_dx, cp_dx, accumulate, workspace,
                    phase=3, sums=None, count=0):
        ct = torch.float64 if x.dtype == torch.float64 else torch.float32
        xv = x[..., c0:c0 + c].to(ct)
        sc, sh = scale_shift[0, :c].to(ct), scale_shift[1, :c].to(ct)
        mu, rs = mean_rstd[0].to(ct), mean_rstd[1].to(ct)
        g = torch.where(xv * sc + sh > 0, dy[..., :c].to(ct), torch.zeros((), dtype=ct, device=x.device))
        xhat = (xv - mu) * rs
        red = tuple(range(xv.dim() - 1))
        if phase & 1:
            dgamma_dbeta.view(2, c)[0] = (g * xhat).sum(red).to(dgamma_dbeta.dtype)
            dgamma_dbeta.view(2, c)[1] = g.sum(red).to(dgamma_dbeta.dtype)
        if phase & 2:
            sv = (dgamma_dbeta if sums is None else sums).view(2, c).to(ct)
            n = count if count > 0 else xv.numel() // c
            d = sc * (g - sv[1] / n - xhat * (sv[0] / n))
            if accumulate:
                dx[..., c0_dx:c0_dx + c] += d.to(dx.dtype)
            else:
                dx[..., c0_dx:c0_dx + cp_dx] = 0
                dx[..., c0_dx:c0_dx + c] = d.to(dx.dtype)
        self.launches += (2 if phase & 1 else 0) + (1 if phase & 2 else 0)

    @staticmethod
    def _duf_apply(logits, res, x, sf, r):
        n, cin, h, w = x.shape
        ct = x.dtype
        rr, k2 = r * r, sf * sf
        f = torch.softmax(logits[..., :k2 * rr].to(ct).reshape(n, h, w, k2, rr), dim=3)
        nb = F.unfold(x.reshape(n * cin, 1, h, w), sf, padding=sf // 2).reshape(n, cin, k2, h, w)
        out = torch.einsum("nckhw,nhwkp->ncphw", nb, f)
        out = out + res[..., :cin * rr].to(ct).reshape(n, h, w, cin, rr).permute(0, 3, 4, 1, 2)
        return F.pixel_shuffle(out.reshape(n, cin * rr, h, w), r)

    def duf_filter(self, logits, res, x, size_filter, r, y):
        y.copy_(self._duf_apply(logits, res, x, size_filter, r))
        self.launches += 1

    def duf_filter_bwd(self, logits, x, dy, size_filter, r, dlogits, dres):
        ct = x.dtype
        lg = logits.detach().to(ct).requires_grad_(True)
        rs = torch.zeros(*dres.shape, dtype=ct, device=x.device, requires_grad=True)
        with torch.enable_grad():
            out = self._duf_apply(lg, rs, x, size_filter, r)
        gl, gr = torch.autograd.grad(out, (lg, rs), dy.to(ct))
        dlogits.copy_(gl.to(dlogits.dtype))
        dres.copy_(gr.to(dres.dtype))
        self.launches += 1

    # ---- device-side data front end --------------------------------------------------------
    def cine_gather(self, vol, tab, r, f_first, f_count, mean, std, out):
        H, W = vol.shape[-2:]
        ph, pw = out.shape[-2], out.shape[-1]
        for i, row in enumerate(tab.tolist()):
            seq, fx, fy, y0, x0 = row[:5]
            for f in range(f_count):
                img = vol[seq, row[5 + f_first + f]]
                if fx:
                    img = img.flip(1)
                if fy:
                    img = img.flip(0)
                crop = img[y0 * r:y0 * r + ph, x0 * r:x0 * r + pw]
                out[f, i, 0] = (crop - torch.tensor(mean, dtype=crop.dtype)) / torch.tensor(std, dtype=crop.dtype)
        self.launches += 1

    # ---- loss / metrics ----------------------------------------------------------------
    def loss_fwd_bwd(self, out, target, kind, param, grad_scale, partials, grad):
        d = out - target
        if kind == 0:
            val, g = d.abs().sum(), torch.sign(d)
        elif kind == 1:
            val, g = (d * d).sum(), 2 * d
        elif kind == 2:
            s = torch.sqrt(d * d + param)
            val, g = s.sum(), d / s
        else:
            ad = d.abs()
            q = ad.clamp_max(param)
            val = (0.5 * q * q + param * (ad - q)).sum()
            g = torch.where(ad < param, d, param * torch.sign(d))
        partials.view(-1)[0] += val.to(partials.dtype)
        if grad is not None:
            grad.copy_((g * grad_scale).to(grad.dtype))
        self.launches += 1

    def loss_fwd_bwd_seg(self, out, target, n_segments, kind, param, grad_scale, partials, grad):
        o, t = out.reshape(n_segments, -1), target.reshape(n_segments, -1)
        g = grad.view(n_segments, -1) if grad is not None else [None] * n_segments
        for s_ in range(n_segments):
            self.loss_fwd_bwd(o[s_], t[s_], kind, param, grad_scale, partials[s_], g[s_])
        self.launches -= n_segments - 1

    def metric_workspace(self, n, per_sample):
        return 16

    @staticmethod
    def _denorm(x, mean, std):
        if std <= 0:
            return x
        return (x * std + mean).round().clamp(0, 255)

    def psnr(self, out, target, mean, std, max_value, psnr_out, workspace):
        a, b = self._denorm(out, mean, std), self._denorm(target, mean, std)
        n = out.shape[0]
        mse = ((a - b) ** 2).reshape(n, -1).mean(1)
        psnr_out.copy_(10 * torch.log10(max_value ** 2 / (mse + 1e-10)))
        self.launches += 2

    def ssim(self, out, target, win11, mean, std, c1, c2, ssim_out, workspace):
        a, b = self._denorm(out, mean, std), self._denorm(target, mean, std)
        n = out.shape[0]
        a = a.reshape(n, 1, *out.shape[-2:])
        b = b.reshape(n, 1, *out.shape[-2:])
        k2 = torch.outer(win11, win11).reshape(1, 1, 11, 11).to(a.dtype)
        conv = lambda t: F.conv2d(t, k2)
        mu1, mu2 = conv(a), conv(b)
        s11 = conv(a * a) - mu1 ** 2
        s22 = conv(b * b) - mu2 ** 2
        s12 = conv(a * b) - mu1 * mu2
        m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s11 + s22 + c2))
        ssim_out.copy_(m.reshape(n, -1).mean(1))
        self.launches += 2

    # ---- standalone --------------------------------------------------------------------
    def pixel_shuffle(self, x, y, r, inverse=False):
        y.copy_(F.pixel_unshuffle(x, r) if inverse else F.pixel_shuffle(x, r))
        self.launches += 1

    def upsample_linear(self, x, y, align_corners):
        mode = "bilinear" if x.dim() == 4 else "trilinear"
        y.copy_(F.interpolate(x, size=y.shape[2:], mode=mode, align_corners=bool(align_corners)))
        self.launches += 1

    def upsample_linear_bwd(self, dy, dx, align_corners):
        mode = "bilinear" if dx.dim() == 4 else "trilinear"
        with torch.enable_grad():
            x = torch.zeros_like(dx, requires_grad=True)
            y = F.interpolate(x, size=dy.shape[2:], mode=mode, align_corners=bool(align_corners))
            dx.copy_(torch.autograd.grad(y, x, dy)[0])
        self.launches += 1

    # ---- flow-based recurrent net (csrc/flow.cu) ----
    def maxpool2x2(self, x, y, idx):
        v, i = F.max_pool2d(x.permute(0, 3, 1, 2), 2, return_indices=True)
        n, h, w, c = x.shape
        iy, ix = i // w, i % w
        y.copy_(v.permute(0, 2, 3, 1))
        idx.copy_(((iy % 2) * 2 + ix % 2).permute(0, 2, 3, 1).to(idx.dtype))
        self.launches += 1

    def maxpool2x2_bwd(self, dy, idx, dx):
        n, h, w, c = dx.shape
        d = dx.view(n, h // 2, 2, w // 2, 2, c)
        for k in range(4):
            d[:, :, k // 2, :, k % 2, :] = torch.where(idx == k, dy, torch.zeros_like(dy))
        self.launches += 1

    def upsample2x_nhwc(self, x, y):
        y.copy_(F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1))
        self.launches += 1

    def upsample2x_nhwc_bwd(self, dy, dx):
        with torch.enable_grad():
            x = torch.zeros_like(dx, requires_grad=True)
            y = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
            dx.copy_(torch.autograd.grad(y, x, dy)[0])
        self.launches += 1

    def flow_tanh(self, z, y0, x0, flow):
        h, w = flow.shape[2:]
        flow.copy_(torch.tanh(z[:, y0:y0 + h, x0:x0 + w, :2]).permute(0, 3, 1, 2))
        self.launches += 1

    def flow_tanh_bwd(self, dflow, flow, y0, x0, dz):
        h, w = flow.shape[2:]
        dz.zero_()
        dz[:, y0:y0 + h, x0:x0 + w, :2] = (dflow * (1 - flow * flow)).permute(0, 2, 3, 1)
        self.launches += 1

    @staticmethod
    def _stn_grid(flow):
        n, _, h, w = flow.shape
        ys = torch.linspace(-1, 1, h, dtype=torch.float64)
        xs = torch.linspace(-1, 1, w, dtype=torch.float64)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        mesh = torch.stack([gx, gy], dim=-1).to(flow.dtype).to(flow.device)
        return mesh.unsqueeze(0) + flow.permute(0, 2, 3, 1)

    def grid_warp(self, img, flow, out):
        out.copy_(F.grid_sample(img, self._stn_grid(flow), mode="bilinear", padding_mode="border", align_corners=False))
        self.launches += 1

    def grid_warp_bwd(self, img, flow, dout, dflow):
        with torch.enable_grad():
            f = flow.detach().clone().requires_grad_(True)
            o = F.grid_sample(img, self._stn_grid(f), mode="bilinear", padding_mode="border", align_corners=False)
            dflow.copy_(torch.autograd.grad(o, f, dout)[0])
        self.launches += 1

    def s2d_cat(self, hr, lr, r, out):
        out.zero_()
        out[..., :r * r] = F.pixel_unshuffle(hr, r).permute(0, 2, 3, 1)
        out[..., r * r] = lr[:, 0]
        self.launches += 1

    def s2d_cat_bwd(self, dout, r, dhr):
        dhr.copy_(F.pixel_shuffle(dout[..., :r * r].permute(0, 3, 1, 2), r))
        self.launches += 1

    # ---- weight gradients of several 1x1 convolutions over one feature list (csrc/wgrad_shared.cu) ----
    MAX_SHARED_ACC, MAX_SHARED_DZ, MAX_SHARED_SRCS = 8, 4, 8

    def wgrad_shared_ok(self, srcs, dzs, ntaps):
        c = dzs[0].shape[-1]
        return (all(t.shape[-1] == c for t in (*srcs, *dzs)) and len(srcs) <= self.MAX_SHARED_SRCS and
                len(dzs) <= self.MAX_SHARED_DZ and sum((n + 1) // 2 for n in ntaps) <= self.MAX_SHARED_ACC)

    def wgrad_shared(self, srcs, dzs, ntaps, dws, dbs, accumulate, workspace_of):
        workspace_of(16)
        for dz, nt, dw, db in zip(dzs, ntaps, dws, dbs):
            ct = torch.float64 if dz.dtype == torch.float64 else torch.float32
            c = dz.shape[-1]
            z = dz.reshape(-1, c).to(ct)
            g = torch.stack([z.t() @ srcs[t].reshape(-1, c).to(ct) for t in range(nt)]).reshape(-1)        # [nt][j][k]
            tgt = dw.view(-1)[:g.numel()]
            tgt.copy_((tgt.to(ct) + g if accumulate else g).to(dw.dtype))
            if db is not None:
                cs = z.sum(0)
                db.copy_((db.to(ct) + cs if accumulate else cs).to(db.dtype))
        self.launches += 2

    # ---- TOFlowNet (csrc/toflow.cu) ----
    def upsample_bicubic(self, x, r, y):
        y.copy_(F.interpolate(x.reshape(-1, 1, *x.shape[-2:]), scale_factor=r, mode="bicubic", align_corners=False).reshape(y.shape))
        self.launches += 1

    def min_partials(self, x, partials):
        partials.fill_(float("inf"))
        partials[0] = x.min()
        self.launches += 1

    def pad_fill(self, x, y0, x0, partials, out):
        h, w = x.shape[-2:]
        out.copy_(torch.full_like(out, float(partials.min())))
        out[..., y0:y0 + h, x0:x0 + w] = x
        self.launches += 1

    def avgpool2x2(self, x, y):
        y.copy_(F.avg_pool2d(x.reshape(-1, 1, *x.shape[-2:]), 2, 2).reshape(y.shape))
        self.launches += 1

    @staticmethod
    def _toflow_warp(nbr, flow):
        n, _, h, w = flow.shape
        gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        vx = gx.to(flow.dtype).to(flow.device) + flow[:, 0]
        vy = gy.to(flow.dtype).to(flow.device) + flow[:, 1]
        grid = torch.stack((2.0 * vx / max(w - 1, 1) - 1.0, 2.0 * vy / max(h - 1, 1) - 1.0), dim=3)
        return F.grid_sample(nbr.reshape(n, 1, h, w), grid, mode="bilinear", padding_mode="zeros", align_corners=False)[:, 0]

    def warp_cat(self, out, c_ref, ref, c_w, nbr, flow, scale, c_flow):
        n, h, w, _ = out.shape
        if ref is not None:
            out[..., c_ref] = ref.reshape(n, h, w)
        f = scale * flow if flow is not None else torch.zeros(n, 2, h, w, dtype=out.dtype, device=out.device)
        if nbr is not None:
            out[..., c_w] = self._toflow_warp(nbr, f)
        if c_flow >= 0:
            out[..., c_flow:c_flow + 2] = f.permute(0, 2, 3, 1)
        self.launches += 1

    def warp_cat_bwd(self, dout, c_w, nbr, flow, scale, c_flow, dflow):
        g = torch.zeros_like(flow)
        if nbr is not None:
            with torch.enable_grad():
                f = flow.detach().clone().requires_grad_(True)
                o = self._toflow_warp(nbr, scale * f)
                g = torch.autograd.grad(o, f, dout[..., c_w].contiguous())[0]
        if c_flow >= 0:
            g = g + scale * dout[..., c_flow:c_flow + 2].permute(0, 3, 1, 2)
        dflow.copy_(g)
        self.launches += 1

    def flow_add(self, z, flow_up, scale, flow):
        flow.copy_(scale * flow_up + z[..., :2].permute(0, 3, 1, 2))
        self.launches += 1

    def planar_to_nhwc(self, d, y0, x0, dz):
        n, kc, h, w = d.shape
        dz.zero_()
        dz[:, y0:y0 + h, x0:x0 + w, :kc] = d.permute(0, 2, 3, 1)
        self.launches += 1

    def head_add(self, z, xref, y0, x0, out):
        h, w = out.shape[-2:]
        n = z.shape[0]
        out.copy_((z[:, y0:y0 + h, x0:x0 + w, 0] + xref.reshape(n, *z.shape[1:3])[:, y0:y0 + h, x0:x0 + w]).reshape(out.shape))
        self.launches += 1

    def adam_flat(self, p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
        gi = g * grad_scale
        if weight_decay != 0:
            gi = gi + weight_decay * p
        m.mul_(beta1).add_(gi, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        denom = v.sqrt() / math.sqrt(bc2) + eps
        p.addcdiv_(m, denom, value=-lr / bc1)
        self.launches += 1

    def adam_flat_dev(self, p, g, m, v, hyper):
        lr, b1, b2, eps, wd, step, gs = [float(x) for x in hyper.tolist()]
        self.adam_flat(p, g, m, v, lr, b1, b2, eps, wd, int(round(step)), gs)

    def scale_(self, x, alpha):
        x.mul_(alpha)
        self.launches += 1

    # deferred wgrad reduction: the emulation keeps the running sums in a side table keyed by workspace
    def tapgemm_wgrad_partial(self, tab, srcs, dz, workspace, slice_, n_slices, db_period):
        ct = torch.float64 if dz.dtype == torch.float64 else torch.float32
        dw = torch.zeros(tab.n_taps_total * tab.nt * tab.kc, dtype=ct, device=dz.device)
        db = torch.zeros(db_period, dtype=ct, device=dz.device)
        self.tapgemm_wgrad(tab, srcs, dz, dw, False, workspace, db=db, db_period=db_period)
        store = self.__dict__.setdefault("_wg_store", {})
        key = workspace.data_ptr()
        if slice_ > 0 and key in store:
            store[key][0].add_(dw)
            store[key][1].add_(db)
        else:
            store[key] = [dw, db]
        return True

    def tapgemm_wgrad_finish(self, tab, srcs, dz, dw, db, db_period, accumulate, used, n_slices, workspace):
        pdw, pdb = self._wg_store.pop(workspace.data_ptr())
        flat = dw.view(-1)[:pdw.numel()]
        if accumulate:
            flat += pdw.to(dw.dtype)
            db += pdb.to(db.dtype)
        else:
            flat.copy_(pdw.to(dw.dtype))
            db.copy_(pdb.to(db.dtype))
        self.launches += 1
