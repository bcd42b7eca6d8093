"""FRVSRNet (after RBPNet): oracle restatement vs goldens made by the real reference, host logic of the
drop-in through the kernel emulation (CPU), GPU parity through the C-ABI."""
import glob
import os

import pytest
import torch

from oracle import restated
from oracle.make_golden import seeded_fill

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "frvsrnet_*.pt")))
ids = [os.path.basename(p)[:-3] for p in CASES]


def _state(fx):
    return seeded_fill({k: torch.zeros(s) for k, s in fx["state_shapes"].items()}, fx["state_seed"])


def _losses(sr_imgs, lr_imgs, fx):
    l1 = torch.nn.MSELoss()         # the fixtures use a smooth loss (oracle/make_golden_frvsr.py)
    flow = torch.stack([l1(a, b.to(a.device)) for a, b in zip(lr_imgs, fx["inputs"])]).mean()      # acdc_frvsr_trainer.py:86
    sr = torch.stack([l1(a, b.to(a.device)) for a, b in zip(sr_imgs, fx["targets"])]).mean()       # :87
    return flow, sr


def _oracle_grads(fx):
    sd = {k: v.clone().requires_grad_(True) for k, v in _state(fx).items()}
    sr, lr = restated.frvsrnet_forward(fx["inputs"], sd, fx["kwargs"]["upscale_factor"])
    flow_loss, sr_loss = _losses(sr, lr, fx)
    (flow_loss + sr_loss).backward()
    return sr, lr, flow_loss, sr_loss, {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("path", CASES, ids=ids)
def test_oracle_restatement_matches_reference_golden(path):
    fx = torch.load(path)
    sr, lr, flow_loss, sr_loss, grads = _oracle_grads(fx)
    for o, ref in zip(sr, fx["sr_imgs"]):
        assert (o.detach() - ref).abs().max() <= 1e-5 * ref.abs().max()
    for o, ref in zip(lr, fx["lr_imgs"]):
        assert (o.detach() - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert abs(float(flow_loss) - float(fx["flow_loss"])) <= 1e-6 and abs(float(sr_loss) - float(fx["sr_loss"])) <= 1e-6
    for k, dg in fx["grad_digest"].items():
        assert abs(float(grads[k].norm()) - float(dg["norm"])) <= 1e-4 * float(dg["norm"]) + 1e-9, k
        assert (grads[k].reshape(-1)[:16] - dg["head"]).abs().max() <= 1e-4 * float(grads[k].abs().max()) + 1e-9, k


def _oracle_grads64(fx):
    fx64 = dict(fx, inputs=[x.double() for x in fx["inputs"]], targets=[x.double() for x in fx["targets"]])
    sd = {k: v.double().requires_grad_(True) for k, v in _state(fx).items()}
    sr, lr = restated.frvsrnet_forward(fx64["inputs"], sd, fx["kwargs"]["upscale_factor"])
    flow_loss, sr_loss = _losses(sr, lr, fx64)
    (flow_loss + sr_loss).backward()
    return fx64, sr, lr, {k: v.grad for k, v in sd.items()}


def _grad_err(got, ref64, prefix=""):
    """(largest element error / largest gradient element, global relative L2) against the float64 oracle, over the
    parameters whose name starts with `prefix`"""
    keys = [k for k in ref64 if k.startswith(prefix)]
    gmax = max(float(ref64[k].abs().max()) for k in keys)
    worst = num = den = 0.0
    for k in keys:
        d = got[k].double().cpu() - ref64[k]
        worst = max(worst, float(d.abs().max()) / gmax)
        num += float((d ** 2).sum())
        den += float((ref64[k] ** 2).sum())
    return worst, (num / den) ** 0.5


def _check_net(net, fx, device, out_tol, grad_tol, srnet_floor=None):
    """outputs against the golden of the real reference; gradients against the oracle evaluated in FLOAT64.
    SRNet's parameters: `grad_tol` (1e-4).  The flow net's parameters: d(loss)/d(flow) is a difference of neighbouring
    pixels of the (detached) previous output times w / 2 and the net is full of kinks (max pooling, LeakyReLU, the warp's
    cell), so its fp32 gradients are ill-conditioned whoever computes them - the reference's own fp32 arithmetic (the fp32
    oracle) is 3e-6 .. 4e-3 away from float64 on these fixtures depending on the input AND the machine's rounding; the bar
    there is 5e-3 of the largest flow-net gradient (or four times the fp32 oracle's own error if larger).  On the GPU the
    same floor applies to SRNet's parameters (`srnet_floor`): the CUDA-core tap-GEMM accumulates K = 9 C terms sequentially in
    fp32, which leaves ~2e-6 (relative) in the flow after FNet's 14 convolutions - measured against the emulation, entry by
    entry - and the warp of the previous output carries it into SRNet's input of the next frame.
    test_host_logic_exact_in_float64 shows that the tables and the recorded backward are exact."""
    _, _, _, _, g32 = _oracle_grads(fx)
    _, _, _, g64 = _oracle_grads64(fx)
    sr, lr = net([x.to(device) for x in fx["inputs"]])
    for o, ref in zip(sr, fx["sr_imgs"]):
        assert o.shape == ref.shape
        assert (o.detach().cpu() - ref).abs().max() <= out_tol * ref.abs().max()
    for o, ref in zip(lr, fx["lr_imgs"]):
        assert (o.detach().cpu() - ref).abs().max() <= out_tol * ref.abs().max()
    flow_loss, sr_loss = _losses(sr, lr, fx)
    # (loss bar of DESIGN.md section 1: relative 1e-5; the fixtures' losses are 1.0 .. 1.8)
    assert abs(float(flow_loss) - float(fx["flow_loss"])) <= 1e-5 * max(1.0, float(fx["flow_loss"]))
    assert abs(float(sr_loss) - float(fx["sr_loss"])) <= 1e-5 * max(1.0, float(fx["sr_loss"]))
    (flow_loss + sr_loss).backward()
    got = {k: p.grad.detach() for k, p in net.named_parameters()}
    for prefix, floor in (("srnet.", srnet_floor or grad_tol), ("fnet.", 5e-3)):
        e_ref, e_got = _grad_err(g32, g64, prefix), _grad_err(got, g64, prefix)
        print(f"{prefix} gradient error vs the float64 oracle: ours {e_got[0]:.2e} / {e_got[1]:.2e}, the fp32 oracle {e_ref[0]:.2e} / {e_ref[1]:.2e}")
        assert e_got[0] <= max(floor, 4 * e_ref[0]) and e_got[1] <= max(floor, 4 * e_ref[1]), prefix


def test_host_logic_exact_in_float64():
    """tables, packing maps and the recorded backward in float64 through the emulation = the float64 oracle to round-off:
    whatever differs in fp32 is arithmetic, not logic"""
    from tests.emu import EmuOps
    from vsr_b200.frvsr import FRVSRNet
    for path in CASES:
        fx = torch.load(path)
        fx64, sr, lr, g64 = _oracle_grads64(fx)
        net = FRVSRNet(**fx["kwargs"])
        net.load_state_dict(_state(fx))
        net = net.double()
        net._ops = EmuOps()
        s2, l2 = net(fx64["inputs"])
        for a, b in zip(s2 + l2, sr + lr):
            assert (a.detach() - b.detach()).abs().max() <= 1e-10
        f2, ss2 = _losses(s2, l2, fx64)
        (f2 + ss2).backward()
        worst, l2err = _grad_err({k: p.grad.detach() for k, p in net.named_parameters()}, g64)
        assert worst <= 1e-10 and l2err <= 1e-10


def test_state_dict_contract():
    from vsr_b200.frvsr import FRVSRNet
    fx = torch.load(CASES[0])
    net = FRVSRNet(**fx["kwargs"])
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == fx["state_shapes"]
    assert list(net.state_dict()) == list(fx["state_shapes"])
    with pytest.raises(ValueError):
        FRVSRNet(1, 1, 2)
    with pytest.raises(ValueError):
        FRVSRNet(1, 1, 4, precision="bf16")


@pytest.mark.parametrize("path", CASES, ids=ids)
def test_plan_and_recorded_backward_through_the_emulation(path):
    from tests.emu import EmuOps
    from vsr_b200.frvsr import FRVSRNet
    fx = torch.load(path)
    net = FRVSRNet(**fx["kwargs"])
    net.load_state_dict(_state(fx))
    net._ops = EmuOps()
    _check_net(net, fx, "cpu", 1e-4, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=ids)
def test_gpu_matches_reference_golden(path):
    from vsr_b200.frvsr import FRVSRNet
    fx = torch.load(path)
    net = FRVSRNet(**fx["kwargs"])
    net.load_state_dict(_state(fx))
    _check_net(net.cuda(), fx, "cuda", 1e-4, 1e-4, srnet_floor=5e-3)


@pytest.mark.gpu
def test_flow_kernels_match_the_emulation():
    """csrc/flow.cu against the torch emulation: max pooling (+ indices), bilinear x2 of pixel-major maps, flow head with the
    crop, STN warp (in-range, clipped and border samples) and its flow gradient, space-to-depth + concatenation"""
    from tests.emu import EmuOps
    from vsr_b200.ops import cuda_ops
    ops, emu = cuda_ops(), EmuOps()
    g = torch.Generator(device="cuda").manual_seed(3)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)

    def both(fn, outs):
        res = []
        for o in (ops, emu):
            bufs = [torch.full_like(t, 7) if t.is_floating_point() else torch.zeros_like(t) for t in outs]
            fn(o, *bufs)
            res.append(bufs)
        return res

    x = rnd(2, 6, 10, 24)
    (y, i), (ye, ie) = both(lambda o, y, i: o.maxpool2x2(x, y, i), [torch.empty(2, 3, 5, 24, device="cuda"), torch.empty(2, 3, 5, 24, device="cuda", dtype=torch.uint8)])
    assert torch.equal(y, ye) and torch.equal(i, ie)
    dy = rnd(2, 3, 5, 24)
    (dx,), (dxe,) = both(lambda o, d: o.maxpool2x2_bwd(dy, i, d), [torch.empty_like(x)])
    assert torch.equal(dx, dxe)
    for shp in ((2, 5, 7, 8), (1, 1, 1, 4), (1, 1, 3, 40)):
        x = rnd(*shp)
        up = torch.empty(shp[0], 2 * shp[1], 2 * shp[2], shp[3], device="cuda")
        (y,), (ye,) = both(lambda o, y: o.upsample2x_nhwc(x, y), [up])
        assert (y - ye).abs().max() <= 1e-6
        dy = rnd(*up.shape)
        (dx,), (dxe,) = both(lambda o, d: o.upsample2x_nhwc_bwd(dy, d), [torch.empty_like(x)])
        assert (dx - dxe).abs().max() <= 1e-5
    z = rnd(2, 16, 24, 32)
    (f,), (fe,) = both(lambda o, f: o.flow_tanh(z, 2, 3, f), [torch.empty(2, 2, 12, 20, device="cuda")])
    assert (f - fe).abs().max() <= 1e-6
    df = rnd(2, 2, 12, 20)
    (dz,), (dze,) = both(lambda o, d: o.flow_tanh_bwd(df, fe, 2, 3, d), [torch.empty_like(z)])
    assert (dz - dze).abs().max() <= 1e-6
    for (h, w, scale) in ((16, 16, 0.3), (12, 20, 2.5), (5, 1, 0.5), (64, 64, 0.05)):
        img, flow, dout = rnd(2, 1, h, w), rnd(2, 2, h, w) * scale, rnd(2, 1, h, w)
        (o1,), (o2,) = both(lambda o, out: o.grid_warp(img, flow, out), [torch.empty_like(img)])
        assert (o1 - o2).abs().max() <= 2e-5 * max(1.0, float(img.abs().max()))
        (d1,), (d2,) = both(lambda o, d: o.grid_warp_bwd(img, flow, dout, d), [torch.empty_like(flow)])
        # a sample within round-off of a cell boundary may take the neighbouring cell: allow a few such elements
        bad = (d1 - d2).abs() > 1e-4 * max(1.0, float(d2.abs().max()))
        assert int(bad.sum()) <= 2, int(bad.sum())
    hr, lr = rnd(2, 1, 24, 32), rnd(2, 1, 6, 8)
    (s,), (se,) = both(lambda o, out: o.s2d_cat(hr, lr, 4, out), [torch.empty(2, 6, 8, 32, device="cuda")])
    assert torch.equal(s, se)
    ds = rnd(2, 6, 8, 32)
    (dh,), (dhe,) = both(lambda o, d: o.s2d_cat_bwd(ds, 4, d), [torch.empty_like(hr)])
    assert torch.equal(dh, dhe)


def _frvsr_step_vs_oracle(device, steps, use_graph, w_tol, l_tol):
    """FRVSRTrainStep against the oracle stepped with torch.optim.Adam (acdc_frvsr_trainer.py:41-50,85-88): both losses of
    every step, PSNR / SSIM of the SR frames, the weights after `steps` steps (eps = 1e-4: tests/test_trainstep_gpu.py)"""
    from tests.emu import EmuOps
    from vsr_b200.frvsr import FRVSRNet
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import FRVSRTrainStep
    kw = dict(in_channels=1, out_channels=1, upscale_factor=4, num_resblocks=1)
    shapes = {k: tuple(v.shape) for k, v in FRVSRNet(**kw).state_dict().items()}
    sd0 = seeded_fill({k: torch.zeros(s) for k, s in shapes.items()}, 11)
    g = torch.Generator().manual_seed(12)
    base = torch.randn(2, 1, 8, 8, generator=g)
    lrs = [base + 0.3 * torch.randn(2, 1, 8, 8, generator=g) for _ in range(2)]
    hrs = [torch.randn(2, 1, 32, 32, generator=g) for _ in range(2)]
    net = FRVSRNet(**kw)
    net.load_state_dict(sd0)
    if device == "cpu":
        net._ops = EmuOps()
    net = net.to(device)
    opt = FlatAdam(net.parameters(), lr=1e-3, eps=1e-4)
    step = FRVSRTrainStep(net, [torch.nn.L1Loss(), torch.nn.MSELoss()], [1.0, 0.5], [PSNR().to(device), SSIM().to(device)], opt,
                          "acdc", use_graph=use_graph)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    ref_opt = torch.optim.Adam(list(sd.values()), lr=1e-3, eps=1e-4)
    for _ in range(steps):
        acc = torch.zeros(5, device=device)
        lv, outs = step.train_step([x.to(device) for x in lrs], [y.to(device) for y in hrs], acc)
        sr, lr = restated.frvsrnet_forward(lrs, sd, 4)
        flow_loss = torch.stack([restated.l1_loss(a, b) for a, b in zip(lr, lrs)]).mean()
        sr_loss = torch.stack([restated.mse_loss(a, b) for a, b in zip(sr, hrs)]).mean()
        psnr, ssim = restated.vsr_metrics([o.detach() for o in sr], hrs)
        ref_opt.zero_grad()
        (flow_loss + 0.5 * sr_loss).backward()
        ref_opt.step()
        assert abs(float(lv[0]) - float(flow_loss)) <= l_tol * float(flow_loss)
        assert abs(float(lv[1]) - float(sr_loss)) <= l_tol * float(sr_loss)
        assert abs(float(acc[0]) - float(flow_loss + 0.5 * sr_loss)) <= l_tol * float(flow_loss + 0.5 * sr_loss)
        assert abs(float(acc[3]) - float(psnr)) <= 2e-3 and abs(float(acc[4]) - float(ssim)) <= 1e-4
    wmax = max(float(v.abs().max()) for v in sd.values())
    for k, p in net.named_parameters():
        assert (p.data.cpu() - sd[k].data).abs().max() <= w_tol * wmax, k


def test_frvsr_train_step_matches_reference_step_host_logic():
    _frvsr_step_vs_oracle("cpu", 2, False, 2e-5, 2e-5)


@pytest.mark.gpu
def test_frvsr_train_step_gpu_graphed():
    """through the C-ABI, the step replayed as a CUDA graph (2 eager + capture + replay).  The losses of every step check
    the weights the previous steps produced at 1e-4; the weights themselves after 4 steps at 1e-3 of the largest weight:
    Adam divides by sqrt(v) + eps, so the flow net's ill-conditioned fp32 gradients (_check_net) move elements with small
    gradients by a visible fraction of lr = 1e-3 per step"""
    _frvsr_step_vs_oracle("cuda", 4, True, 1e-3, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=ids)
def test_gpu_bf16x3_matches_reference_golden(path):
    """precision='bf16x3': SRNet's 64 -> 64 layers on the tcgen05 tap-GEMM as three bf16 products per product, the flow net
    and the 17-channel head on the CUDA cores, fp32 maps - the same bars as the strict fp32 mode"""
    from vsr_b200.frvsr import FRVSRNet
    from vsr_b200.ops import SplitOps
    fx = torch.load(path)
    net = FRVSRNet(precision="bf16x3", **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.cuda()
    assert isinstance(net._backend(), SplitOps) and {"s_b0_1", "s_d1", "s_d2"} <= set(net._planB.fwd) and "f4_2" in net._plan.fwd
    _check_net(net, fx, "cuda", 1e-4, 1e-4, srnet_floor=5e-3)


@pytest.mark.gpu
def test_frvsr_train_step_gpu_graphed_bf16x3():
    """the fused, CUDA-graphed FRVSR step in the mixed mode: graph replay = eager launch bit for bit"""
    from vsr_b200.frvsr import FRVSRNet
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import FRVSRTrainStep
    fx = torch.load(CASES[0])
    res = []
    for use_graph in (False, True):
        net = FRVSRNet(precision="bf16x3", **fx["kwargs"])
        net.load_state_dict(_state(fx))
        net = net.cuda()
        step = FRVSRTrainStep(net, [torch.nn.L1Loss(), torch.nn.MSELoss()], [1.0, 0.5], [PSNR().cuda(), SSIM().cuda()],
                              FlatAdam(net.parameters(), lr=1e-4), "acdc", use_graph=use_graph)
        log = []
        for _ in range(5):
            acc = torch.zeros(5, device="cuda")
            lv, _ = step.train_step([x.cuda() for x in fx["inputs"]], [y.cuda() for y in fx["targets"]], acc)
            log.append(torch.cat([lv.reshape(-1), acc]).clone())
        res.append((torch.stack(log), net.flat.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
