"""Per-parameter gradient error of precision='bf16x3' (and 'fp32') against the float64 oracle on a golden fixture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restated                       # noqa: E402
from tests.test_oracle import _state              # noqa: E402
from vsr_b200.nets import DRFNet                  # noqa: E402

fx = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "drfnet_f64_g2_x4.pt"))
print("kwargs", fx["kwargs"], "inputs", len(fx["inputs"]), tuple(fx["inputs"][0].shape))
sd = {k: v.double().requires_grad_(True) for k, v in _state(fx).items()}
ref_outs = restated.drfnet_forward([t.double() for t in fx["inputs"]], sd, fx["kwargs"]["upscale_factor"])
torch.stack([torch.nn.MSELoss()(o, t.double()) for o, t in zip(ref_outs, fx["targets"])]).mean().backward()
ref = {k: v.grad for k, v in sd.items()}
gmax = max(float(g.abs().max()) for g in ref.values())
for prec in ("fp32", "bf16x3"):
    net = DRFNet(precision=prec, **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.cuda()
    outs = net([t.cuda() for t in fx["inputs"]])
    torch.stack([torch.nn.MSELoss()(o, t.cuda()) for o, t in zip(outs, fx["targets"])]).mean().backward()
    print(prec, "output err", max(float((o.detach().cpu().double() - r.detach()).abs().max() / r.detach().abs().max())
                                  for o, r in zip(outs, ref_outs)))
    rows = []
    for k, p in net.named_parameters():
        d = (p.grad.cpu().double() - ref[k])
        rows.append((float(d.abs().max()) / gmax, float(d.norm() / (ref[k].norm() + 1e-30)), k, float(ref[k].abs().max())))
    rows.sort(reverse=True)
    for r in rows[:8]:
        print("   %-44s worst/gmax %.2e  rel L2 %.2e  |g|max %.3e" % (r[2], r[0], r[1], r[3]))

# ---- per-launch check of the split tap-GEMM against float64 arithmetic on the same operands ----
from vsr_b200 import _lib                        # noqa: E402
from vsr_b200.ops import SplitOps, split_ops     # noqa: E402

ops = split_ops()
orig = SplitOps.tapgemm
jj, kk = torch.arange(256).view(256, 1), torch.arange(64).view(1, 64)


def plain_w(w, tab):
    """[wh | wh | wl] swizzled slabs of every group -> float64 [n_taps][nt][64] (wh + wl)"""
    nt = tab.nt
    pos = (jj[:nt] * 64 + (((kk >> 3) ^ (jj[:nt] & 7)) << 3) + (kk & 7)).reshape(-1).to(w.device)
    out, off = [], 0
    for _, taps in tab.groups:
        n = len(taps)
        blk = w[off:off + 3 * n * nt * 64].view(3, n, nt * 64).double()
        full = blk[0] + blk[2]
        assert torch.equal(blk[0], blk[1])
        out.append(full[:, pos].view(n, nt, 64))
        off += 3 * n * nt * 64
    return torch.cat(out)


def checked(self, tab, srcs, out, w, bias=None, epi=0, **kw):
    raw = torch.empty_like(out)
    orig(self, tab, srcs, raw, w)
    W = plain_w(w, tab)
    n, h, wd, _ = out.shape
    ref = torch.zeros(n, h, wd, out.shape[-1], dtype=torch.float64, device=out.device)
    ti = 0
    for o0, taps in tab.groups:
        for (s, dy, dx, c0) in taps:
            x = srcs[s][..., c0:c0 + 64].double()
            xs = torch.zeros_like(x)
            ys, ye = max(0, -dy), min(h, h - dy)
            xs_, xe = max(0, -dx), min(wd, wd - dx)
            if ys < ye and xs_ < xe:
                xs[:, ys:ye, xs_:xe] = x[:, ys + dy:ye + dy, xs_ + dx:xe + dx]
            ref[..., o0:o0 + tab.nt] += xs @ W[ti].t()
            ti += 1
    touched = torch.zeros(out.shape[-1], dtype=torch.bool, device=out.device)
    for o0, _ in tab.groups:
        touched[o0:o0 + tab.nt] = True
    err = float((raw.double() - ref)[..., touched].abs().max() / ref.abs().max())
    flag = "  <<<<" if err > 3e-5 else ""
    print(f"  taps{tab.n_taps_total:3d} nt{tab.nt:3d} g{tab.n_groups} srcs{len(srcs)} out{tuple(out.shape)} epi{epi}: raw err {err:.2e}{flag}")
    rc = orig(self, tab, srcs, out, w, bias=bias, epi=epi, **kw)
    # the epilogue in float64 on the float64 accumulators
    v = ref.clone()
    if epi & _lib.EPI_BIAS:
        v = v + bias.double()
    if epi & _lib.EPI_RES_PRE:
        v = v + kw["residual"].double()
    sl = float(kw["slope"]) if kw.get("slope") is not None else None
    if epi & _lib.EPI_PRELU_BWD:
        y = kw["aux_y"].double()
        v = torch.where(y > 0, v, sl * v)
    if epi & _lib.EPI_PRELU:
        v = torch.where(v > 0, v, sl * v)
    e2 = float((out.double() - v).abs().max() / v.abs().max())
    print(f"        after the epilogue: err {e2:.2e}" + ("  <<<<" if e2 > 3e-5 else ""))
    return rc


orig_wg = SplitOps.tapgemm_wgrad


def checked_wg(self, tab, srcs, dz, dw, accumulate, workspace, db=None, db_period=0):
    tmp = torch.zeros(tab.n_taps_total, tab.nt, 64, device=dz.device)
    orig_wg(self, tab, srcs, dz, tmp, False, workspace)
    n, h, wd, _ = dz.shape
    worst, ti = 0.0, 0
    scale = 0.0
    refs = []
    for o0, taps in tab.groups:
        for (s, dy, dx, c0) in taps:
            x = srcs[s][..., c0:c0 + 64].double()
            xs = torch.zeros_like(x)
            ys, ye = max(0, -dy), min(h, h - dy)
            xs_, xe = max(0, -dx), min(wd, wd - dx)
            if ys < ye and xs_ < xe:
                xs[:, ys:ye, xs_:xe] = x[:, ys + dy:ye + dy, xs_ + dx:xe + dx]
            refs.append(torch.einsum("nhwj,nhwk->jk", dz[..., o0:o0 + tab.nt].double(), xs))
    ref = torch.stack(refs)
    err = float((tmp.double() - ref).abs().max() / ref.abs().max())
    print(f"  WGRAD taps{tab.n_taps_total:3d} nt{tab.nt:3d} g{tab.n_groups} srcs{len(srcs)} dz{tuple(dz.shape)}: err {err:.2e}" + ("  <<<<" if err > 3e-5 else ""))
    return orig_wg(self, tab, srcs, dz, dw, accumulate, workspace, db=db, db_period=db_period)


SplitOps.tapgemm_wgrad = checked_wg


SplitOps.tapgemm = checked
net = DRFNet(precision="bf16x3", **fx["kwargs"])
net.load_state_dict(_state(fx))
net = net.cuda()
outs = net([t.cuda() for t in fx["inputs"]])
print("---- backward ----")
torch.stack([torch.nn.MSELoss()(o, t.cuda()) for o, t in zip(outs, fx["targets"])]).mean().backward()
