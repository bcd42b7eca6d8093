"""precision='bf16x3' (alias 'tf32'): the strict mode on tensor cores - fp32 maps, every product as three bf16 tcgen05
products (csrc/split.cu, ops.SplitOps).  The bar is the strict mode's: outputs and gradients within 1e-4 (tensor-normalised)
of the real reference's golden, and of the oracle on the config-2 model."""
import os

import pytest
import torch

from oracle import restated
from tests.test_oracle import _state
from vsr_b200 import _lib
from vsr_b200.nets import DRFNet
from vsr_b200.ops import SplitOps, TapTable, split_ops

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _grad_errors(got, ref):
    gmax = max(float(g.abs().max()) for g in ref.values())
    num = sum(float(((got[k].double() - g.double()) ** 2).sum()) for k, g in ref.items()) ** 0.5
    den = sum(float((g.double() ** 2).sum()) for g in ref.values()) ** 0.5
    worst = max(float((got[k].double() - g.double()).abs().max()) for k, g in ref.items()) / gmax
    return num / den, worst


def test_split_planes_and_raw_tapgemm_against_float64():
    """one 3x3 convolution (9 taps x 2 channel blocks, two groups) through split -> tcgen05 x3 -> raw accumulators, and
    every epilogue flag set of the schedule through vsr_tap_epilogue, against float64 arithmetic on the same operands"""
    ops = split_ops()
    g = torch.Generator().manual_seed(1)
    n, h, w, cin, cout = 3, 20, 24, 128, 128
    x = torch.randn(n, h, w, cin, generator=g).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    pl = ops._planes(x)
    back = pl[0].float() + pl[1].float()
    assert float((back - x).abs().max()) <= 2.0 ** -16 * float(x.abs().max())
    groups, slabs = [], []
    for o0 in (0, 64):
        taps = []
        for ky in range(3):
            for kx in range(3):
                for b in range(cin // 64):
                    taps.append((0, ky - 1, kx - 1, b * 64))
                    slabs.append(wt[o0:o0 + 64, b * 64:(b + 1) * 64, ky, kx])        # [nt, kc]
        groups.append((o0, taps))
    tab = TapTable(64, 64, groups)
    # packed slabs [wh | wh | wl] per group, each in the swizzled image
    jj, kk = torch.arange(64).view(64, 1), torch.arange(64).view(1, 64)
    pos = (jj * 64 + (((kk >> 3) ^ (jj & 7)) << 3) + (kk & 7)).reshape(-1).cuda()

    def image(s):
        out = torch.empty(64 * 64, dtype=torch.bfloat16, device="cuda")
        out[pos] = s.reshape(-1)
        return out

    per_group = len(groups[0][1])
    chunks = []
    for gi in range(2):
        sl = slabs[gi * per_group:(gi + 1) * per_group]
        hi = [s.bfloat16() for s in sl]
        lo = [(s - s.bfloat16().float()).bfloat16() for s in sl]
        chunks += [image(t) for t in hi] + [image(t) for t in hi] + [image(t) for t in lo]
    wbuf = torch.cat(chunks)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), padding=1).permute(0, 2, 3, 1)
    out = torch.empty(n, h, w, cout, device="cuda")
    ops.tapgemm(tab, [x], out, wbuf)
    torch.cuda.synchronize()
    err = _rel(out.double(), ref)
    print("raw bf16x3 tap-GEMM vs float64:", err)
    assert err <= 2e-5
    # epilogue flag sets of the schedule
    bias = torch.randn(cout, generator=g).cuda()
    slope = torch.tensor([0.2]).cuda()
    res = torch.randn(n, h, w, cout, generator=g).cuda()
    out2 = torch.empty_like(out)
    ops.tapgemm(tab, [x], out, wbuf, bias=bias, epi=_lib.EPI_BIAS | _lib.EPI_PRELU | _lib.EPI_OUT2, slope=slope, out2=out2, res2=res)
    z = ref + bias.double()
    y = torch.where(z > 0, z, 0.2 * z)
    assert _rel(out.double(), y) <= 2e-5 and _rel(out2.double(), y + res.double()) <= 2e-5
    partials = torch.zeros(ops.partials_len, device="cuda")
    aux = torch.randn(n, h, w, cout, generator=g).cuda()
    ops.tapgemm(tab, [x], out, wbuf, epi=_lib.EPI_PRELU_BWD | _lib.EPI_RES_PRE, slope=slope, residual=res, aux_y=aux,
                slope_partials=partials)
    v = ref + res.double()
    want = torch.where(aux.double() > 0, v, 0.2 * v)
    dslope = float((v * (aux.double() / 0.2) * (aux.double() <= 0)).sum())
    assert _rel(out.double(), want) <= 2e-5
    assert abs(float(partials.double().sum()) - dslope) <= 1e-4 * abs(dslope) + 1e-3


def test_split_weight_gradient_against_float64():
    ops = split_ops()
    g = torch.Generator().manual_seed(2)
    n, h, w, cin, cout = 4, 32, 32, 64, 128
    x = torch.randn(n, h, w, cin, generator=g).cuda()
    dz = torch.randn(n, h, w, cout, generator=g).cuda()
    taps = [(0, ky - 1, kx - 1, 0) for ky in range(3) for kx in range(3)]
    tab = TapTable(64, 128, [(0, taps)])
    dw = torch.zeros(9, 128, 64, device="cuda")
    fused = ops.tapgemm_wgrad(tab, [x], dz, dw, False, None)
    assert fused is False
    xp = torch.nn.functional.pad(x.double(), (0, 0, 1, 1, 1, 1))
    for t, (_, dy, dx, _) in enumerate(taps):
        sh = xp[:, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
        ref = torch.einsum("nhwj,nhwk->jk", dz.double(), sh)
        assert _rel(dw[t].double(), ref) <= 2e-5, t
    dw2 = dw.clone()
    ops.tapgemm_wgrad(tab, [x], dz, dw2, True, None)
    assert _rel(dw2, 2 * dw) <= 1e-6


def _branch_matched_oracle(net, outs, inputs, targets, sd64, loss="l1", forward=None):
    """Gradients of the float64 oracle with every PReLU on the branch OUR forward pass took, fed the loss gradient of OUR
    outputs.  PReLU' and d(L1) jump at zero, so two correct evaluations whose activations differ by round-off (here
    ~1e-5 of the map's range, the 16 significant bits of a bf16 pair) disagree on the branch of the few elements that
    close to zero, and each disagreement moves the gradient by O(1) of that element's share - a property of the net, not
    of the arithmetic.  Pinning the branches (and checking that every pinned element really is within round-off of zero)
    compares what the mode computes: the backward pass.  Returns (grads, disagreeing elements, their largest |x| / max|x|)."""
    from tests.emu import EmuOps
    P = net._plan
    F_, r, G = P.F, P.r, P.G
    nchw = lambda z: z.permute(0, 3, 1, 2)
    hr = lambda z: EmuOps._unblock(z, r, F_, P.phases)
    masks = []
    for S in outs[0].grad_fn.saved:                 # the frames saved by the engine, in the oracle's call order
        seq = [nchw(S.a1), nchw(S.inn), nchw(S.lr[0])]
        for g in range(G):
            if g >= 1:
                seq.append(nchw(S.u[g]))
            seq.append(hr(S.hr[g]))
            if g >= 1:
                seq.append(hr(S.d[g]))
            seq.append(nchw(S.lr[g + 1]))
        seq.append(nchw(S.f))
        if P.variant == "srfb":                     # r_block.prelu1 on the up-projected features (srfb_net.py:46)
            seq.append(hr(S.s[0]))
        masks += [(m > 0).cpu() for m in seq]
    stats = {"flips": 0, "worst": 0.0, "n": 0}
    real = restated._prelu

    def pinned(x, sd, key):
        m = masks.pop(0)
        assert m.shape == x.shape, (key, m.shape, x.shape)
        dis = m != (x.detach() > 0)
        stats["flips"] += int(dis.sum())
        stats["n"] += dis.numel()
        if dis.any():
            stats["worst"] = max(stats["worst"], float(x.detach().abs()[dis].max() / x.detach().abs().max()))
        return x * torch.where(m, torch.ones_like(x), sd[key + ".weight"].expand_as(x))

    restated._prelu = pinned
    try:
        ref_outs = forward(sd64) if forward is not None else restated.drfnet_forward([t.double() for t in inputs], sd64, r)
    finally:
        restated._prelu = real
    assert not masks
    T = len(outs)
    if loss == "l1":
        g = [torch.sign(o.detach().cpu().double() - t.double()) / (o.numel() * T) for o, t in zip(outs, targets)]
    else:
        g = [2.0 * (o.detach().cpu().double() - t.double()) / (o.numel() * T) for o, t in zip(outs, targets)]
    torch.autograd.backward(ref_outs, g)
    return {k: v.grad for k, v in sd64.items()}, ref_outs, stats


def _check_against_oracle(net, inputs, targets, what):
    sd64 = {k: v.detach().cpu().double().requires_grad_(True) for k, v in net.state_dict().items()}
    outs = net([t.cuda() for t in inputs])
    loss = torch.stack([torch.nn.L1Loss()(o, t.cuda()) for o, t in zip(outs, targets)]).mean()
    ref, ref_outs, st = _branch_matched_oracle(net, outs, inputs, targets, sd64)
    loss.backward()
    for o, ro in zip(outs, ref_outs):
        assert _rel(o.detach().cpu().double(), ro.detach()) <= 1e-4
    ref_loss = torch.stack([torch.nn.L1Loss()(o.detach(), t.double()) for o, t in zip(ref_outs, targets)]).mean()
    assert abs(loss.item() - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    l2, worst = _grad_errors({k: p.grad.cpu() for k, p in net.named_parameters()}, ref)
    print(f"{what}: gradient vs the float64 oracle rel L2 {l2:.2e}, worst element / max {worst:.2e}; "
          f"{st['flips']} of {st['n']} PReLU inputs on the other branch in float64, the largest {st['worst']:.1e} of its map's range")
    # every pinned element is within round-off of zero, and there are few of them
    assert st["worst"] <= 1e-4 and st["flips"] <= 1e-4 * st["n"]
    assert l2 <= 1e-4 and worst <= 1e-4
    return outs


@pytest.mark.parametrize("precision", ["bf16x3", "tf32"])
def test_bf16x3_mode_matches_reference_golden(precision):
    """the strict bar (north_star fp32 mode: <= 1e-4) on the tensor cores: outputs and loss against the real reference's
    golden, every gradient against the float64 oracle on the same PReLU branches (_branch_matched_oracle)"""
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f64_g2_x4.pt"))
    net = DRFNet(precision=precision, **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.to("cuda")
    assert isinstance(net._backend(), SplitOps)
    outs = _check_against_oracle(net, fx["inputs"], fx["targets"], "golden fixture")
    for o, ref in zip(outs, fx["outputs"]):
        assert (o.detach().cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    loss = torch.stack([torch.nn.L1Loss()(o.detach().cpu(), t) for o, t in zip(outs, fx["targets"])]).mean()
    assert abs(loss.item() - float(fx["loss_l1"])) <= 1e-5 * abs(float(fx["loss_l1"]))
    with torch.no_grad():
        again = net.eval()([t.cuda() for t in fx["inputs"]])
    for a, b in zip(again, outs):
        assert torch.equal(a, b.detach())


@pytest.mark.parametrize("r,hw", [(2, (24, 20)), (3, (16, 12)), (8, (12, 16))])
def test_bf16x3_other_upscale_factors_against_the_oracle(r, hw):
    """every supported upscale factor (6x6 s2, 7x7 s3, 12x12 s8 projection tables; maps wider than 1024 channels at x8
    take the term-by-term weight gradient)"""
    torch.manual_seed(r)
    net = DRFNet(precision="bf16x3", in_channels=1, out_channels=1, num_features=64, num_groups=2, upscale_factor=r).to("cuda")
    g = torch.Generator().manual_seed(10 + r)
    x = [torch.randn(3, 1, *hw, generator=g) for _ in range(3)]
    y = [torch.randn(3, 1, hw[0] * r, hw[1] * r, generator=g) for _ in range(3)]
    _check_against_oracle(net, x, y, f"x{r}")


def test_bf16x3_config2_model_against_the_oracle():
    """the config-2 model (DRFNet-L: F64, G6, x4, T5) on 2 patches of LR 32x32 directly against the oracle evaluated in
    float64: outputs, loss and every gradient within the strict bar"""
    from bench import MODEL, make_batches
    lrs, hrs = make_batches(1, 2, seed=5, pinned=False)[0]
    torch.manual_seed(0)
    net = DRFNet(precision="bf16x3", **MODEL).to("cuda")
    _check_against_oracle(net, lrs, hrs, "config-2 model")


# ---- the other 64-channel nets in the strict tensor-core mode -------------------------------------------------------------
def _float64_grads(fwd, sd, outs, targets, loss_of):
    """gradients of the float64 oracle fed the loss gradient of OUR outputs (d(L1) jumps where an output crosses its target)"""
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    ref_outs = fwd(sd64)
    g = [torch.sign(o.detach().cpu().double() - t.double()) / (o.numel() * len(outs)) for o, t in zip(outs, targets)]
    torch.autograd.backward(ref_outs, g)
    return {k: v.grad for k, v in sd64.items()}, ref_outs


@pytest.mark.parametrize("name", ["srfbnet_f64_g2_x4", "edsrnet_f64_b2_x4"])
def test_bf16x3_srfbnet_and_edsrnet_match_reference_golden(name):
    """outputs at the strict bar against the real reference's goldens; gradients against the float64 oracle: measured and
    printed; the bar is 1e-4 where no activation sits on its kink (these fixtures), see _branch_matched_oracle otherwise"""
    from tests import test_secondary_nets as S
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    net = S.CLS[fx["cls"]](precision="bf16x3", **fx["kwargs"])
    net.load_state_dict(S._state(fx))
    net = net.to("cuda")
    assert isinstance(net._backend(), SplitOps)
    out = net(fx["input"].cuda())
    outs = out if isinstance(out, list) else [out]
    for o, ref in zip(outs, fx["outputs"]):
        assert (o.detach().cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    y = fx["target"]
    loss = torch.stack([torch.nn.L1Loss()(o, y.cuda()) for o in outs]).mean()
    assert abs(float(loss) - float(fx["loss_l1"])) <= 1e-5 * float(fx["loss_l1"])
    loss.backward()
    kw = fx["kwargs"]
    got = {k: p.grad.cpu() for k, p in net.named_parameters()}
    if fx["cls"] == "SRFBNet":
        # the feedback net has PReLUs: the oracle on OUR branches (an element or two of this fixture sit on the kink: 1.03e-4
        # against the plain float64 oracle)
        fwd = lambda sd: restated.srfbnet_forward(fx["input"].double(), sd, kw["upscale_factor"], kw["num_steps"])
        sd64 = {k: v.double().requires_grad_(True) for k, v in S._state(fx).items()}
        net.zero_grad()
        out = net(fx["input"].cuda())
        ref, _, st = _branch_matched_oracle(net, out, None, [y] * len(out), sd64, forward=fwd)
        assert st["worst"] <= 1e-4 and st["flips"] <= 1e-4 * st["n"]
        print(f"   {st['flips']} of {st['n']} PReLU inputs pinned, the largest {st['worst']:.1e} of its map's range")
    else:
        fwd = lambda sd: [restated.edsrnet_forward(fx["input"].double(), sd, kw["upscale_factor"])]
        ref, _ = _float64_grads(fwd, S._state(fx), outs, [y] * len(outs), None)
    l2, worst = _grad_errors(got, ref)
    print(f"{name} bf16x3: gradient vs the float64 oracle rel L2 {l2:.2e}, worst element / max {worst:.2e}")
    assert l2 <= 1e-4 and worst <= 1e-4


def test_bf16x3_rbpnet_matches_reference_golden():
    """outputs against the real reference's golden at the strict bar; gradients against the float64 oracle with every PReLU
    slope set to 1 (the activations become the identity, so no element can sit on a kink - with the golden's slopes a single
    PReLU input within round-off of zero moves this small fixture's gradient by 5e-4, whichever branch is 'right'): every
    tap table, data gradient, weight gradient and epilogue of the net is exercised and compared as arithmetic"""
    from tests import test_rbpnet as R
    from vsr_b200.rbpn import RBPNet
    fx = torch.load(R.BIG)
    kw = fx["kwargs"]
    net = RBPNet(precision="bf16x3", **kw)
    net.load_state_dict(R._state(fx))
    net = net.cuda()
    assert isinstance(net._backend(), SplitOps)
    with torch.no_grad():
        out = net([t.cuda() for t in fx["inputs"]])
    assert (out.cpu() - fx["output"]).abs().max() <= 1e-4 * fx["output"].abs().max()
    sd = {k: (torch.ones_like(v) if k.endswith("act.weight") else v) for k, v in R._state(fx).items()}
    net.load_state_dict(sd)
    out = net([t.cuda() for t in fx["inputs"]])
    torch.nn.L1Loss()(out, fx["target"].cuda()).backward()
    fwd = lambda sd64: [restated.rbpnet_forward([t.double() for t in fx["inputs"]], sd64, kw["upscale_factor"], kw["num_frames"])]
    ref, ref_outs = _float64_grads(fwd, sd, [out], [fx["target"]], None)
    assert _rel(out.detach().cpu().double(), ref_outs[0].detach()) <= 1e-4
    l2, worst = _grad_errors({k: p.grad.cpu() for k, p in net.named_parameters()}, ref)
    print(f"rbpnet_b64_f64_x4 bf16x3 (slopes = 1): gradient vs the float64 oracle rel L2 {l2:.2e}, worst element / max {worst:.2e}")
    assert l2 <= 1e-4 and worst <= 1e-4
