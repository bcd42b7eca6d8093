"""TOFlowNet: oracle restatement vs goldens made by the real reference, host logic of the drop-in
through the kernel emulation (CPU, fp32 and exact in float64), GPU parity through the C-ABI."""
import glob
import os

import pytest
import torch

from oracle import restated
from oracle.make_golden_toflow import fill

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "toflownet_*.pt")))
ids = [os.path.basename(p)[:-3] for p in CASES]


def _state(fx):
    return fill({k: torch.zeros(s, dtype=fx["state_dtypes"][k]) for k, s in fx["state_shapes"].items()}, fx["state_seed"])


def _oracle(fx, dtype=torch.float32):
    sd = {k: (v.to(dtype).requires_grad_(True) if v.dtype.is_floating_point and "running_" not in k else v.clone())
          for k, v in _state(fx).items()}
    buffers = {k: v.to(dtype).clone() for k, v in sd.items() if "running_" in k}
    out = restated.toflownet_forward([x.to(dtype) for x in fx["inputs"]], sd, fx["kwargs"]["upscale_factor"], True, buffers)
    loss = restated.mse_loss(out, fx["target"].to(dtype))
    loss.backward()
    return out, loss, {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point and v.requires_grad}, buffers


@pytest.mark.parametrize("path", CASES, ids=ids)
def test_oracle_restatement_matches_reference_golden(path):
    fx = torch.load(path)
    out, loss, grads, buffers = _oracle(fx)
    assert (out.detach() - fx["output"]).abs().max() <= 1e-5 * fx["output"].abs().max()
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * float(fx["loss"])
    for k, dg in fx["grad_digest"].items():
        assert abs(float(grads[k].norm()) - float(dg["norm"])) <= 1e-3 * float(dg["norm"]) + 1e-8, k
    for k, v in buffers.items():
        assert (v - fx["buffers_after"][k]).abs().max() <= 1e-5 * max(1.0, float(fx["buffers_after"][k].abs().max())), k


def _oracle64(fx):
    return _oracle(fx, torch.float64)


def _grad_err(got, ref64):
    gmax = max(float(g.abs().max()) for g in ref64.values())
    worst = num = den = 0.0
    for k, g in ref64.items():
        d = got[k].double().cpu() - g
        worst = max(worst, float(d.abs().max()) / gmax)
        num += float((d ** 2).sum())
        den += float((g ** 2).sum())
    return worst, (num / den) ** 0.5


def _load(fx, dtype=torch.float32):
    from vsr_b200.toflow import TOFlowNet
    net = TOFlowNet(**fx["kwargs"])
    net.load_state_dict(_state(fx))
    return net.to(dtype).train()


def test_state_dict_contract():
    from vsr_b200.toflow import TOFlowNet
    fx = torch.load(CASES[0])
    net = TOFlowNet(**fx["kwargs"])
    sd = net.state_dict()
    assert list(sd) == list(fx["state_shapes"])
    assert {k: tuple(v.shape) for k, v in sd.items()} == fx["state_shapes"]
    assert {k: v.dtype for k, v in sd.items()} == fx["state_dtypes"]
    assert net.ref_idx == 1 and TOFlowNet(1, 1, 4, 2).ref_idx == 1 and TOFlowNet(1, 1, 7, 4).ref_idx == 3     # toflow_net.py:21
    with pytest.raises(ValueError):
        TOFlowNet(1, 1, 3, 2, precision="bf16")
    with pytest.raises(RuntimeError):
        net([torch.zeros(1, 1, 8, 8)] * 3)           # no CPU fallback


@pytest.mark.parametrize("path", CASES, ids=ids)
def test_host_logic_exact_in_float64(path):
    """tables, packing maps, the recorded backward and the BatchNorm buffer updates in float64 through the emulation = the
    float64 oracle to round-off: whatever differs in fp32 is arithmetic, not logic"""
    from tests.emu import EmuOps
    fx = torch.load(path)
    out64, loss64, g64, buf64 = _oracle64(fx)
    net = _load(fx, torch.float64)
    net._ops = EmuOps()
    out = net([x.double() for x in fx["inputs"]])
    assert out.shape == fx["output"].shape
    assert (out.detach() - out64.detach()).abs().max() <= 1e-9 * float(out64.abs().max())
    restated.mse_loss(out, fx["target"].double()).backward()
    worst, l2 = _grad_err({k: p.grad.detach() for k, p in net.named_parameters()}, g64)
    assert worst <= 1e-9 and l2 <= 1e-9, (worst, l2)
    sd = net.state_dict()
    for k, v in buf64.items():
        assert (sd[k].double() - v).abs().max() <= 1e-9 * max(1.0, float(v.abs().max())), k
    for k, v in fx["buffers_after"].items():
        if "num_batches" in k:
            assert int(sd[k]) == int(v), k


def _pinned_oracle64(fx, tape):
    """gradients of the float64 oracle with every ReLU on the branch OUR forward pass took (masks from the recorded launches,
    which run in the oracle's call order: per neighbour frame, per pyramid level, the four BatchNorm + ReLU of the block,
    then the three ReLUs of the output block).  Returns (grads, number of pinned elements that the oracle itself puts on
    the other branch, the largest |x| / max|x| among them)."""
    import torch.nn.functional as F
    masks = []
    for e in tape:
        if e[0] == "bn":
            _, _, _, a, _, _, c = e
            masks.append((a[..., :c] > 0).permute(0, 3, 1, 2).cpu())
        elif e[0] == "conv" and e[1] in ("o0", "o1", "o2"):
            masks.append((e[3][..., :64] > 0).permute(0, 3, 1, 2).cpu())
    stats = {"flips": 0, "worst": 0.0}
    real = F.relu

    def relu(x, inplace=False):
        m = masks.pop(0)
        assert m.shape == x.shape, (m.shape, x.shape)
        dis = m != (x.detach() > 0)
        if dis.any():
            stats["flips"] += int(dis.sum())
            stats["worst"] = max(stats["worst"], float(x.detach().abs()[dis].max() / x.detach().abs().max()))
        return x * m.to(x.dtype)

    F.relu = relu
    try:
        _, _, g64, _ = _oracle(fx, torch.float64)
    finally:
        F.relu = real
    assert not masks
    return g64, stats["flips"], stats["worst"]


def _check_net(net, fx, device, grad_tol):
    """outputs, loss and running buffers against the golden of the real reference at 1e-4 / 1e-5; gradients against the
    oracle in FLOAT64 **on the same ReLU branches**.  SpyNet is full of kinks (a ReLU after every BatchNorm): ONE ReLU input
    within round-off of zero that takes the other branch moves the gradient of these small fixtures by 5e-3, and which
    elements do depends on the summation order (the same tables with 16- or 32-channel blocks: 3e-5 or 5e-3 against the plain
    float64 oracle through the fp32 emulation) - so the oracle is evaluated with its ReLUs pinned to OUR branches, every
    pinned element that it would have put on the other branch is checked to lie within round-off of zero, and the bar stays
    at `grad_tol` (or four times the fp32 oracle's own distance from float64: the warp's bilinear cell is a second, un-pinned
    kink).  That tables, packing maps and the recorded backward are right is shown exactly by
    test_host_logic_exact_in_float64 (1e-9)."""
    _, _, g32, _ = _oracle(fx)
    _, _, g64_plain, _ = _oracle64(fx)
    out = net([x.to(device) for x in fx["inputs"]])
    assert out.shape == fx["output"].shape
    assert (out.detach().cpu() - fx["output"]).abs().max() <= 1e-4 * fx["output"].abs().max()
    g64, flips, worst_x = _pinned_oracle64(fx, out.grad_fn.tape)
    loss = torch.nn.MSELoss()(out, fx["target"].to(device))
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * float(fx["loss"])
    loss.backward()
    sd = net.state_dict()
    for k, v in fx["buffers_after"].items():
        if "num_batches" in k:
            assert int(sd[k]) == int(v), k
        else:
            assert (sd[k].cpu() - v).abs().max() <= 1e-5 * max(1.0, float(v.abs().max())), k
    got = {k: p.grad.detach() for k, p in net.named_parameters()}
    e_ref = _grad_err(g32, g64_plain)
    e_got, e_plain = _grad_err(got, g64), _grad_err(got, g64_plain)
    print(f"TOFlowNet gradient error vs the float64 oracle (worst / L2): ours {e_got[0]:.2e} / {e_got[1]:.2e} on our ReLU branches "
          f"({flips} pinned elements, the largest {worst_x:.1e} of its map's range; {e_plain[0]:.2e} / {e_plain[1]:.2e} against the "
          f"plain oracle), the fp32 oracle {e_ref[0]:.2e} / {e_ref[1]:.2e}")
    assert flips <= 20 and worst_x <= 1e-4
    assert e_got[0] <= max(grad_tol, 4 * e_ref[0]) and e_got[1] <= max(grad_tol, 4 * e_ref[1])


@pytest.mark.parametrize("path", CASES, ids=ids)
def test_plan_and_recorded_backward_through_the_emulation(path):
    from tests.emu import EmuOps
    fx = torch.load(path)
    net = _load(fx)
    net._ops = EmuOps()
    _check_net(net, fx, "cpu", 1e-4)


def test_eval_mode_uses_running_buffers():
    from tests.emu import EmuOps
    fx = torch.load(CASES[0])
    net = _load(fx).eval()
    net._ops = EmuOps()
    sd = {k: v.clone() for k, v in _state(fx).items()}
    with torch.no_grad():
        out = net(fx["inputs"])
        ref = restated.toflownet_forward(fx["inputs"], sd, fx["kwargs"]["upscale_factor"], training=False)
    assert (out - ref).abs().max() <= 1e-4 * ref.abs().max()
    for k, v in net.state_dict().items():
        assert torch.equal(v, sd[k]), k              # buffers untouched


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=ids)
def test_gpu_matches_reference_golden(path):
    fx = torch.load(path)
    _check_net(_load(fx).cuda(), fx, "cuda", 1e-4)


@pytest.mark.gpu
def test_gpu_eval_mode_and_larger_frames():
    """eval mode (running buffers) and a batch of larger, non-padded frames on the device against the oracle"""
    fx = torch.load(CASES[0])
    net = _load(fx).cuda().eval()
    sd = _state(fx)
    g = torch.Generator().manual_seed(5)
    base = torch.randn(3, 1, 24, 32, generator=g)
    inputs = [base + 0.2 * torch.randn(3, 1, 24, 32, generator=g) for _ in range(3)]
    with torch.no_grad():
        out = net([x.cuda() for x in inputs])
        ref = restated.toflownet_forward(inputs, sd, fx["kwargs"]["upscale_factor"], training=False)
    assert (out.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()


@pytest.mark.gpu
def test_toflow_kernels_match_the_emulation():
    """csrc/toflow.cu against the torch emulation: bicubic up-sampling, minimum + padding, average pooling, warp + concat and
    its flow gradient, flow update, channel scatter, output head"""
    from tests.emu import EmuOps
    from vsr_b200.ops import cuda_ops
    ops, emu = cuda_ops(), EmuOps()
    g = torch.Generator(device="cuda").manual_seed(4)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)

    def both(fn, outs):
        res = []
        for o in (ops, emu):
            bufs = [torch.full_like(t, 7) for t in outs]
            fn(o, *bufs)
            res.append(bufs)
        return res

    for r, (h, w) in ((2, (16, 16)), (4, (6, 10)), (3, (5, 7))):
        x = rnd(3, 2, h, w)
        (y,), (ye,) = both(lambda o, y: o.upsample_bicubic(x, r, y), [torch.empty(3, 2, h * r, w * r, device="cuda")])
        assert (y - ye).abs().max() <= 2e-6 * float(ye.abs().max())
    x = rnd(4, 3, 24, 40)
    (p,), (pe,) = both(lambda o, p: o.min_partials(x, p), [torch.empty(ops.partials_len, device="cuda")])
    assert float(p.min()) == float(x.min()) == float(pe.min())
    (q,), (qe,) = both(lambda o, q: o.pad_fill(x, 4, 3, p, q), [torch.empty(4, 3, 32, 48, device="cuda")])
    assert torch.equal(q, qe)
    (a,), (ae,) = both(lambda o, a: o.avgpool2x2(x, a), [torch.empty(4, 3, 12, 20, device="cuda")])
    assert (a - ae).abs().max() <= 1e-6
    for (h, w, mag, scale) in ((16, 16, 0.7, 2.0), (12, 20, 3.0, 1.0), (2, 2, 0.5, 2.0), (48, 32, 0.1, 1.0)):
        ref, nbr, flow = rnd(2, h, w), rnd(2, h, w), rnd(2, 2, h, w) * mag
        (o1,), (o2,) = both(lambda o, out: (out.zero_(), o.warp_cat(out, 0, ref, 1, nbr, flow, scale, 2)), [torch.empty(2, h, w, 16, device="cuda")])
        assert (o1 - o2).abs().max() <= 2e-5 * max(1.0, float(nbr.abs().max()))
        dout = rnd(2, h, w, 16)
        (d1,), (d2,) = both(lambda o, d: o.warp_cat_bwd(dout, 1, nbr, flow, scale, 2, d), [torch.empty_like(flow)])
        bad = (d1 - d2).abs() > 1e-4 * max(1.0, float(d2.abs().max()))       # a sample on a cell boundary may take the next cell
        assert int(bad.sum()) <= 2, int(bad.sum())
        (d1,), (d2,) = both(lambda o, d: o.warp_cat_bwd(dout, 5, nbr, flow, scale, -1, d), [torch.empty_like(flow)])
        bad = (d1 - d2).abs() > 1e-4 * max(1.0, float(d2.abs().max()))
        assert int(bad.sum()) <= 2, int(bad.sum())
    z, fup = rnd(2, 12, 20, 16), rnd(2, 2, 12, 20)
    (f,), (fe,) = both(lambda o, f: o.flow_add(z, fup, 2.0, f), [torch.empty(2, 2, 12, 20, device="cuda")])
    assert (f - fe).abs().max() <= 1e-6
    d = rnd(2, 2, 8, 14)
    (dz,), (dze,) = both(lambda o, t: o.planar_to_nhwc(d, 3, 2, t), [torch.empty(2, 12, 20, 16, device="cuda")])
    assert torch.equal(dz, dze)
    xr = rnd(2, 12, 20)
    (o1,), (o2,) = both(lambda o, t: o.head_add(z, xr, 3, 2, t), [torch.empty(2, 1, 8, 14, device="cuda")])
    assert torch.equal(o1, o2)


def _misr_step_vs_oracle(device, steps, use_graph, w_tol, l_tol):
    """the fused MISR step (acdc_misr_trainer.py:8-50: net -> loss on the centre frame -> backward -> Adam -> PSNR / SSIM) with
    TOFlowNet against the oracle stepped with torch.optim.Adam: loss of every step, the BatchNorm running buffers (every
    SpyNet block runs once per neighbour frame and step) and the weights after `steps` steps (eps = 1e-4: see
    tests/test_trainstep_gpu.py on why)"""
    from tests.emu import EmuOps
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainStep
    fx = torch.load(CASES[0])
    r = fx["kwargs"]["upscale_factor"]
    net = _load(fx)
    if device == "cpu":
        net._ops = EmuOps()
    net = net.to(device)
    opt = FlatAdam(net.parameters(), lr=1e-4, eps=1e-4)
    step = MISRTrainStep(net, [torch.nn.MSELoss()], [1.0], [PSNR().to(device), SSIM().to(device)], opt, "acdc", use_graph=use_graph)
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running_" not in k else v.clone())
          for k, v in _state(fx).items()}
    buffers = {k: v for k, v in sd.items() if "running_" in k}
    ref_opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-4, eps=1e-4)
    for it in range(steps):
        acc = torch.zeros(4, device=device)
        lv, _ = step.train_step([x.to(device) for x in fx["inputs"]], [fx["target"].to(device)], acc)
        out = restated.toflownet_forward(fx["inputs"], sd, r, True, buffers)
        loss = restated.mse_loss(out, fx["target"])
        psnr, ssim = restated.vsr_metrics([out.detach()], [fx["target"]])
        ref_opt.zero_grad()
        loss.backward()
        ref_opt.step()
        print(f"   step loss {float(lv[0]):.6f} oracle {float(loss):.6f} rel {abs(float(lv[0]) - float(loss)) / float(loss):.1e}")
        # the first step checks the whole fused step at l_tol; from the second step on the weights carry whatever ReLU
        # branches differed in the previous backward pass (see _check_net): 2e-3
        tol = l_tol if it == 0 else 2e-3
        assert abs(float(lv[0]) - float(loss)) <= tol * float(loss)
        assert abs(float(acc[2]) - float(psnr)) <= 2e-3 + 10 * tol and abs(float(acc[3]) - float(ssim)) <= 1e-4 + tol
    wmax = max(float(v.abs().max()) for v in sd.values() if v.requires_grad)
    got = net.state_dict()
    worst = 0.0
    for k, v in sd.items():
        if "num_batches" in k:
            assert int(got[k]) == steps * (fx["kwargs"]["num_frames"] - 1), k
        elif "running_" in k:
            assert (got[k].cpu() - v).abs().max() <= (1e-4 if steps == 1 else 5e-3) * max(1.0, float(v.abs().max())), k
        else:
            worst = max(worst, float((got[k].cpu() - v.data).abs().max()) / wmax)
    print(f"TOFlowNet MISR step on {device}: weights after {steps} steps within {worst:.2e} of the largest weight")
    # (lr = 1e-4: at the fixture's random weights one Adam step of 1e-3 on every weight raises the largest gradient 15x -
    # the flow pyramid is chaotic there and round-off differences of 3e-5 grow to 1e-2 within a step, in the oracle's own
    # fp32 arithmetic as well.)  Adam moves every element by ~lr per step whatever the size of its gradient: what is left
    # are the ill-conditioned fp32 flow gradients (_check_net) as a fraction of lr in elements with small gradients
    assert worst <= w_tol


def test_misr_train_step_with_toflownet_host_logic():
    _misr_step_vs_oracle("cpu", 1, False, 2e-4, 2e-5)


@pytest.mark.gpu
def test_misr_train_step_with_toflownet_gpu():
    """through the C-ABI against the oracle's Adam steps.  Two steps: at the fixture's random weights the flow pyramid is
    chaotic under training (the per-step loss error grows 7e-8 -> 2e-6 -> 1.5e-3 from round-off alone, in eager and graphed
    runs alike), so longer runs are compared with themselves (next test), not with the oracle"""
    _misr_step_vs_oracle("cuda", 2, False, 1e-3, 1e-4)


@pytest.mark.gpu
def test_misr_train_step_with_toflownet_graph_replay_equals_eager():
    """2 eager steps + capture + 2 replays of the CUDA-graphed step = 5 eager steps, bit for bit: losses, weights, BatchNorm
    running buffers and batch counters (they advance inside the graph)"""
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainStep
    fx = torch.load(CASES[0])
    res = []
    for use_graph in (False, True):
        net = _load(fx).cuda()
        opt = FlatAdam(net.parameters(), lr=1e-4, eps=1e-4)
        step = MISRTrainStep(net, [torch.nn.MSELoss()], [1.0], [PSNR().cuda(), SSIM().cuda()], opt, "acdc", use_graph=use_graph)
        losses = []
        for _ in range(5):
            acc = torch.zeros(4, device="cuda")
            lv, _ = step.train_step([x.cuda() for x in fx["inputs"]], [fx["target"].cuda()], acc)
            losses.append(torch.cat([lv.reshape(-1), acc]).clone())
        res.append((torch.stack(losses), {k: v.clone() for k, v in net.state_dict().items()}))
    assert torch.equal(res[0][0], res[1][0])
    for k, v in res[0][1].items():
        assert torch.equal(v, res[1][1][k]), k
    assert int(res[1][1]["spy_net.blocks.0.block.1.num_batches_tracked"]) == 5 * (fx["kwargs"]["num_frames"] - 1)


def test_misr_trainer_epoch_on_cpu():
    """MISRTrainer = AcdcMISRTrainer's epoch loop (base_trainer.py:99-144; the reference trains TOFlowNet with it) over the
    synthetic MISR dataset through the emulation: the log has the reference's keys, the loss falls, validation runs in eval
    mode (running buffers) without touching them"""
    from tests.emu import EmuOps
    from vsr_b200.data import Dataloader, SyntheticCineDataset
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainer
    from vsr_b200.toflow import TOFlowNet
    ds = SyntheticCineDataset(2, num_frames=3, temporal_order="middle", type="train", num_sequences=1, patch_size=(8, 8),
                              misr=True)
    ds.data = ds.data[:4]
    loader = Dataloader(ds, batch_size=2, pin_memory=False)
    torch.manual_seed(0)
    net = TOFlowNet(1, 1, 3, 2)
    net._ops = EmuOps()
    tr = MISRTrainer("cpu", loader, loader, net, [torch.nn.L1Loss()], [1.0], [PSNR(), SSIM()],
                     FlatAdam(net.parameters(), lr=1e-3), None, None, None, 1)
    logs = [tr._run_epoch("training")[0] for _ in range(2)]
    assert list(logs[0]) == ["Loss", "L1Loss", "PSNR", "SSIM"]
    assert logs[1]["Loss"] < logs[0]["Loss"]
    before = {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k}
    vlog, _, out = tr._run_epoch("validation")
    assert out.shape == (2, 1, 16, 16) and vlog["Loss"] > 0
    after = net.state_dict()
    assert all(torch.equal(v, after[k]) for k, v in before.items())
    assert int(after["spy_net.blocks.0.block.1.num_batches_tracked"]) == 2 * 2 * 2        # epochs x batches x neighbours


def test_checkpoint_interchange_with_live_reference(tmp_path):
    """base_trainer.py:229-237 / base_predictor.py:135-136: a checkpoint written from the drop-in loads (strict) into the
    reference's TOFlowNet and back; both nets then agree in training mode (outputs, running buffers) and in eval mode"""
    from oracle import load_reference
    from tests.emu import EmuOps
    from vsr_b200.toflow import TOFlowNet
    if not load_reference.available():
        pytest.skip("/root/reference not mounted")
    Ref = load_reference.load().TOFlowNet
    kw = dict(in_channels=1, out_channels=1, num_frames=3, upscale_factor=2)
    torch.manual_seed(1)
    ours = TOFlowNet(**kw)
    ours._ops = EmuOps()
    torch.save({"net": ours.state_dict()}, tmp_path / "ck.pth")
    ref = Ref(**kw)
    ref.load_state_dict(torch.load(tmp_path / "ck.pth")["net"], strict=True)
    x = [torch.randn(2, 1, 9, 10) for _ in range(3)]                 # padded to 32 x 32 inside the nets
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    ours.train(), ref.train()
    with torch.no_grad():
        a, b = ours(x), ref(list(x))
    assert rel(a, b) <= 2e-5
    sa, sb = ours.state_dict(), ref.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        if "running" in k or "num_batches" in k:
            assert torch.allclose(sa[k].double(), sb[k].double(), rtol=1e-5, atol=1e-6), k
    ours2 = TOFlowNet(**kw)
    ours2.load_state_dict(ref.state_dict(), strict=True)
    ours2._ops = EmuOps()
    ours2.eval(), ref.eval()
    with torch.no_grad():
        assert rel(ours2(x), ref(list(x))) <= 2e-5
