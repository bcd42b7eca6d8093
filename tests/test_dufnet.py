"""DUFNet, the Conv3d network (reference duf_net.py:9-214): oracle restatement vs goldens made
by the real reference, host logic on CPU through the kernel emulation (fp32 tables and the padded bf16 tables),
and GPU parity through the C-ABI (fp32 strict mode, bf16 tcgen05 mode, each new kernel against the emulation)."""
import glob
import os

import pytest
import torch

from oracle import restated
from oracle.make_golden_duf import duf_fill
from tests.emu import EmuOps
from vsr_b200.duf import DUFNet

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "dufnet*.pt")))
ids = lambda ps: [os.path.basename(p)[:-3] for p in ps]
# bf16 storage of every activation of this BatchNorm-heavy net with random weights: the REFERENCE itself under
# torch.autocast(bfloat16) has a global relative L2 gradient error of 0.17 (L1 loss; 0.15 with MSE) and an output
# error of 1.5e-2 against its own fp32 run on these inputs (measured in the build container); the tcgen05 mode
# is held to the same scale: <= 0.25 global L2 against the fp32 mode, digests within 0.2 of the largest gradient.
BF16_GRAD_TOL = 0.2


def _state(fx):
    return duf_fill({k: torch.zeros(s, dtype=fx["state_dtypes"][k]) for k, s in fx["state_shapes"].items()},
                    fx["state_seed"])


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _check(net, fx, dev, out_tol, grad_tol, buf_tol):
    inputs = [f.to(dev) for f in fx["inputs"]]
    net.eval()
    with torch.no_grad():
        assert _rel(net(inputs).cpu(), fx["output_eval"]) <= out_tol        # running statistics
    net.train()
    out = net(inputs)
    assert out.shape == fx["output"].shape
    assert _rel(out.detach().cpu(), fx["output"]) <= out_tol                # batch statistics
    loss = torch.nn.L1Loss()(out, fx["target"].to(dev))
    loss.backward()
    # gradient digests, normalised by the largest gradient norm of the net (gradients of the convolution
    # biases in front of a BatchNorm are exactly zero in exact arithmetic)
    gmax = max(float(d["norm"]) for d in fx["grad_digest"].values())
    hmax = max(float(d["head"].abs().max()) for d in fx["grad_digest"].values())
    for k, p in net.named_parameters():
        d = fx["grad_digest"][k]
        g = p.grad.detach().cpu()
        assert abs(float(g.norm()) - float(d["norm"])) <= grad_tol * gmax, k
        assert float((g.reshape(-1)[:16] - d["head"]).abs().max()) <= grad_tol * hmax, k
    sd = net.state_dict()
    for k, v in fx["buffers_after"].items():                                 # momentum update, unbiased variance
        assert float((sd[k].cpu() - v).abs().max()) <= buf_tol, k
    assert int(sd["denseLayer.tail.bn.num_batches_tracked"]) == 1
    return float(loss.detach())


@pytest.mark.parametrize("path", CASES, ids=ids(CASES))
def test_restated_matches_reference_golden(path):
    fx = torch.load(path)
    kw, sd = fx["kwargs"], _state(fx)
    o = restated.dufnet_forward(fx["inputs"], sd, kw["size_filter"], kw["upscale_factor"], training=True)
    assert _rel(o, fx["output"]) <= 1e-6
    o = restated.dufnet_forward(fx["inputs"], sd, kw["size_filter"], kw["upscale_factor"], training=False)
    assert _rel(o, fx["output_eval"]) <= 1e-6


HOST_CASES = [p for p in CASES if "dufnet16" in p or "dufnet52" in p]     # (the 12-layer fixture is covered on the GPU)


@pytest.mark.parametrize("path", HOST_CASES, ids=ids(HOST_CASES))
def test_host_logic_fp32_matches_reference_golden(path):
    fx = torch.load(path)
    net = DUFNet(**fx["kwargs"])
    assert list(net.state_dict()) == list(fx["state_shapes"])
    net.load_state_dict(_state(fx))
    net._ops = EmuOps()
    loss = _check(net, fx, "cpu", 2e-5, 1e-4, 1e-5)
    assert abs(loss - float(fx["loss_l1"])) <= 1e-5 * float(fx["loss_l1"])


def test_host_logic_bf16_tables():
    """the padded (multiple-of-64) tables and 64-wide growth tiles of the tcgen05 mode, emulated with bf16 storage"""
    fx = torch.load(CASES[1])
    net = DUFNet(precision="bf16", **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net._ops = EmuOps()
    _check(net, fx, "cpu", 5e-2, BF16_GRAD_TOL, 5e-3)


def test_constructor_contract():
    with pytest.raises(AssertionError):
        DUFNet(1, 1, 7, 5, 4, "_DenseLayer99")
    with pytest.raises(ValueError):
        DUFNet(1, 1, 5, 5, 4, "_DenseLayer16")
    net = DUFNet(1, 1, 7, 5, 4, "_DenseLayer16")
    with pytest.raises(RuntimeError):
        net([torch.zeros(1, 1, 8, 8)] * 7)          # no CPU fallback


class _MaskedRelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return x * mask

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


def _oracle64_grads(fx, tol, flips):
    """fp64 oracle gradients (L1 loss, batch statistics) with an explicit ReLU: inputs with |x| < tol are listed (in call
    order) and the i-th of them takes the OTHER branch when i is in `flips`.  Returns (grads, outputs, near-zero inputs)."""
    import torch.nn.functional as F
    kw = fx["kwargs"]
    sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in _state(fx).items()}
    near, real = [], F.relu

    def relu(x, inplace=False):
        mask = (x > 0).to(x.dtype)
        for ix in (x.detach().abs() < tol).nonzero():
            ix = tuple(int(v) for v in ix)
            if len(near) in flips:
                mask[ix] = 1.0 - mask[ix]
            near.append(float(x[ix]))
        return _MaskedRelu.apply(x, mask)

    F.relu = relu
    try:
        out = restated.dufnet_forward([f.double() for f in fx["inputs"]], sd, kw["size_filter"], kw["upscale_factor"], training=True)
        restated.l1_loss(out, fx["target"].double()).backward()
    finally:
        F.relu = real
    return {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.requires_grad}, out.detach(), near


def _grad_err(got, want):
    gmax = max(float(g.abs().max()) for g in want.values())
    return max(float((got[k].double() - g).abs().max()) for k, g in want.items()) / gmax


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=ids(CASES))
def test_gpu_fp32_matches_reference_golden(path):
    """fp32 strict mode: outputs, running buffers and loss against the real reference's golden at 1e-4 / 1e-5; EVERY
    gradient element against the fp64 oracle at 1e-4 of the largest gradient.  A ReLU input within fp32 round-off of zero
    may legitimately take either branch in an fp32 evaluation (the reference's own fp32 run included): the fp64 oracle is
    evaluated with those few elements (|x| < 3e-6, listed by the oracle itself) on the branch that matches - found by
    flipping them one at a time - instead of widening the tolerance (round 1 held the 12- / 24-layer backbones to
    5e-4 / 2e-3)."""
    fx = torch.load(path)
    net = DUFNet(precision="fp32", **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.to("cuda")
    loss = _check(net, fx, "cuda", 1e-4, 1.0, 1e-5)        # (gradient digests vs the fp32 reference: checked below instead)
    assert abs(loss - float(fx["loss_l1"])) <= 1e-5 * float(fx["loss_l1"])
    got = {k: p.grad.detach().cpu() for k, p in net.named_parameters()}
    want, out64, near = _oracle64_grads(fx, 3e-6, set())
    err, flips = _grad_err(got, want), set()
    assert len(near) <= 40, f"{len(near)} ReLU inputs within 3e-6 of zero: the fixture is degenerate"
    for i in range(len(near)):
        if err <= 1e-4:
            break
        trial, _, _ = _oracle64_grads(fx, 3e-6, flips | {i})
        e = _grad_err(got, trial)
        if e < err:
            err, flips = e, flips | {i}
    print(f"{os.path.basename(path)}: {len(near)} ReLU inputs within 3e-6 of zero, branches flipped {[(i, '%.1e' % near[i]) for i in sorted(flips)]}, "
          f"gradient error vs fp64 oracle {err:.2e}")
    assert err <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=ids(CASES))
def test_gpu_bf16_close_to_reference(path):
    fx = torch.load(path)
    net = DUFNet(precision="bf16", **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.to("cuda")
    _check(net, fx, "cuda", 5e-2, BF16_GRAD_TOL, 5e-3)
    g16 = net.flat_grad.clone()
    ref = DUFNet(precision="fp32", **fx["kwargs"])
    ref.load_state_dict(_state(fx))
    ref = ref.to("cuda").train()
    torch.nn.L1Loss()(ref([f.cuda() for f in fx["inputs"]]), fx["target"].cuda()).backward()
    # The bar is the REFERENCE ALGORITHM's own sensitivity to bf16, measured here on the same fixture: the oracle under
    # torch.autocast(bfloat16) (convolutions in bf16, BatchNorm in fp32) against the oracle in fp32.  The tcgen05 mode
    # additionally STORES every activation and gradient map in bf16 and still has to stay within 1.1x that error (measured:
    # 0.095-0.19 against 0.107-0.23); round 1 quoted the reference-under-autocast figure from a one-off measurement.
    kw = fx["kwargs"]
    flat = {}
    for amp in (False, True):
        sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in _state(fx).items()}
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=amp):
            o = restated.dufnet_forward(fx["inputs"], sd, kw["size_filter"], kw["upscale_factor"], training=True)
            l = restated.l1_loss(o.float(), fx["target"])
        l.backward()
        flat[amp] = torch.cat([sd[k].grad.reshape(-1) for k, _ in net.named_parameters()])
    e_ref = float((flat[True] - flat[False]).norm() / flat[False].norm())
    e_got = float((g16 - ref.flat_grad).norm() / ref.flat_grad.norm())
    print(f"{os.path.basename(path)}: bf16 gradient error (global rel L2) {e_got:.3f}; the oracle under autocast(bf16): {e_ref:.3f}")
    assert e_got <= 1.1 * e_ref
    net.train()
    with torch.no_grad():
        out = net([f.cuda() for f in fx["inputs"]]).cpu()
    den = lambda t: restated.denormalize(t, "acdc")
    p_ref = restated.psnr(den(fx["output"]), den(fx["target"]))
    p_got = restated.psnr(den(out), den(fx["target"]))
    assert abs(float(p_ref) - float(p_got)) <= 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_gpu_dense3d_kernels_match_emulation(dtype):
    """vsr_copy_window / vsr_bn_stats / vsr_bn_finalize / vsr_bn_relu / vsr_bn_relu_bwd / vsr_duf_filter{,_bwd}"""
    from vsr_b200.ops import cuda_ops
    ops, emu, dev = cuda_ops(), EmuOps(), "cuda"
    g = torch.Generator(device="cpu").manual_seed(5)
    F, N, h, w, ld, c0, c, cp = 3, 2, 9, 7, 160, 32, 96, 128
    x = (torch.randn(F * N, h, w, ld, generator=g) * 1.5 + 0.3).to(dev).to(dtype)
    ws = lambda nbytes: torch.empty(max(nbytes, 16) // 8 + 1, dtype=torch.float64, device=dev)
    st_a = torch.zeros(F, 2, 200, dtype=torch.float64, device=dev)
    st_b = torch.zeros_like(st_a)
    ops.bn_stats(x, c0, c, F, st_a, 64, ws(ops.bn_stats_workspace(F, N * h * w, c)))
    emu.bn_stats(x, c0, c, F, st_b, 64, None)
    assert torch.allclose(st_a, st_b, rtol=1e-5, atol=1e-4)        # fp32 runs of 32 rows inside a thread, fp64 across
    gamma = (1 + 0.1 * torch.randn(c, generator=g)).to(dev)
    beta = (0.1 * torch.randn(c, generator=g)).to(dev)
    outs = []
    for o in (ops, emu):
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        ss, mr = torch.empty(2, cp, device=dev), torch.empty(2, c, device=dev)
        o.bn_finalize(st_b[1:], 64, 2, N * h * w, c, gamma, beta, 1e-5, 0.1, rm, rv, True, ss, mr)
        outs.append((ss, mr, rm, rv))
    for a, b in zip(*outs):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    ss, mr = outs[1][:2]
    xs = x[N:]                                                   # the two frames the statistics cover
    y_a, y_b = torch.empty(2 * N, h, w, cp, dtype=dtype, device=dev), torch.empty(2 * N, h, w, cp, dtype=dtype, device=dev)
    ops.bn_relu(xs, c0, c, ss, y_a)
    emu.bn_relu(xs, c0, c, ss, y_b)
    assert torch.equal(y_a[..., c:], torch.zeros_like(y_a[..., c:]))
    assert torch.allclose(y_a.float(), y_b.float(), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=1e-6)
    dy = torch.randn(2 * N, h, w, cp, generator=g).to(dev).to(dtype)
    tol = dict(rtol=2e-2, atol=2e-2) if dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-5)
    for acc in (False, True):
        res = []
        for o in (ops, emu):
            gb = torch.zeros(2 * c, device=dev)
            dx = torch.ones(2 * N, h, w, ld, dtype=dtype, device=dev)
            o.bn_relu_bwd(dy, xs, c0, c, ss, mr, gb, dx, 16, c if acc else 112, acc,
                          ws(ops.bn_relu_bwd_workspace(2 * N * h * w, c)))
            res.append((gb, dx.float()))
        assert torch.allclose(res[0][0], res[1][0], rtol=1e-3, atol=1e-3)
        assert torch.allclose(res[0][1], res[1][1], **tol)
        # the two phases apart (synchronised BatchNorm runs an all-reduce between them) = the fused call
        gb2 = torch.zeros(2 * c, device=dev)
        dx2 = torch.ones(2 * N, h, w, ld, dtype=dtype, device=dev)
        wsp = ws(ops.bn_relu_bwd_workspace(2 * N * h * w, c))
        ops.bn_relu_bwd(dy, xs, c0, c, ss, mr, gb2, None, 0, c, False, wsp, phase=1)
        ops.bn_relu_bwd(dy, xs, c0, c, ss, mr, gb2, dx2, 16, c if acc else 112, acc, wsp, phase=2, sums=gb2.clone(),
                        count=2 * N * h * w)
        assert torch.equal(gb2, res[0][0]) and torch.equal(dx2.float(), res[0][1])
    for t_pad, f_out in ((1, F), (0, F - 2)):                    # temporal shift-add (+ statistics of the new slice)
        zz = torch.randn(F * N, h, w, 128, generator=g).to(dev).to(dtype)
        bias = torch.randn(32, generator=g).to(dev)
        o_a, o_b = torch.zeros(f_out * N, h, w, ld, dtype=dtype, device=dev), torch.zeros(f_out * N, h, w, ld, dtype=dtype, device=dev)
        s_a, s_b = torch.zeros(f_out, 2, 200, dtype=torch.float64, device=dev), torch.zeros(f_out, 2, 200, dtype=torch.float64, device=dev)
        ops.tshift_add(zz, 32, F, t_pad, bias, o_a, 64, f_out, s_a, 96, ws(ops.bn_stats_workspace(f_out, N * h * w, 32)))
        emu.tshift_add(zz, 32, F, t_pad, bias, o_b, 64, f_out, s_b, 96, None)
        assert torch.allclose(o_a.float(), o_b.float(), rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=1e-5)
        assert float(o_a[..., :64].abs().max()) == 0 and float(o_a[..., 96:].abs().max()) == 0
        assert torch.allclose(s_a, s_b, rtol=2e-2 if dtype == torch.bfloat16 else 1e-5, atol=0.5 if dtype == torch.bfloat16 else 1e-4)
    for t_pad, f_out in ((1, F), (0, F - 2)):                    # and its inverse for the weight gradient
        gy = torch.randn(f_out * N, h, w, ld, generator=g).to(dev).to(dtype)
        z_a, z_b = torch.full((F * N, h, w, 128), 3.0, dtype=dtype, device=dev), torch.empty(F * N, h, w, 128, dtype=dtype, device=dev)
        ops.tshift_gather(gy, 64, 32, f_out, t_pad, z_a, F)
        emu.tshift_gather(gy, 64, 32, f_out, t_pad, z_b, F)
        assert torch.equal(z_a, z_b)
    d_a = torch.zeros(F * N, h, w, 64, dtype=dtype, device=dev)
    ops.copy_window(x, c0, d_a, 16, 32)
    assert torch.equal(d_a[..., 16:48], x[..., c0:c0 + 32]) and float(d_a[..., :16].abs().max()) == 0
    for cin, sf, r in ((1, 5, 4), (2, 3, 3), (1, 7, 2)):
        cf, cr = sf * sf * r * r, cin * r * r
        ldl, ldr = -(-cf // 64) * 64, 64
        lg = torch.randn(N, h, w, ldl, generator=g).to(dev).to(dtype)
        rs = torch.randn(N, h, w, ldr, generator=g).to(dev).to(dtype)
        img = torch.randn(N, cin, h, w, generator=g).to(dev)
        ya, yb = torch.empty(N, cin, h * r, w * r, device=dev), torch.empty(N, cin, h * r, w * r, device=dev)
        ops.duf_filter(lg, rs, img, sf, r, ya)
        emu.duf_filter(lg, rs, img, sf, r, yb)
        assert torch.allclose(ya, yb, rtol=1e-4, atol=1e-5)
        gy = torch.randn(N, cin, h * r, w * r, generator=g).to(dev)
        ga = [torch.full((N, h, w, ldl), 7.0, dtype=dtype, device=dev), torch.full((N, h, w, ldr), 7.0, dtype=dtype, device=dev)]
        gb_ = [torch.empty_like(ga[0]), torch.empty_like(ga[1])]
        ops.duf_filter_bwd(lg, img, gy, sf, r, ga[0], ga[1])
        emu.duf_filter_bwd(lg, img, gy, sf, r, gb_[0], gb_[1])
        for a, b in zip(ga, gb_):
            assert torch.allclose(a.float(), b.float(), **tol)


# ---- data parallelism with synchronised BatchNorm, two gloo ranks on CPU through the emulation ----
def _sync_worker(rank, world, port, path, ret):
    import torch.distributed as dist
    from vsr_b200.metrics import PSNR
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    fx = torch.load(path)
    net = DUFNet(**fx["kwargs"])
    net.load_state_dict(_state(fx))
    net._ops = EmuOps()
    net.train()
    opt = FlatAdam(net.parameters(), lr=1e-3)
    step = MISRTrainStep(net, [torch.nn.L1Loss()], [1.0], [PSNR()], opt, "acdc")
    n = fx["inputs"][0].shape[0] // world
    sl = slice(rank * n, (rank + 1) * n)
    lv, outs = step.train_step([f[sl] for f in fx["inputs"]], [fx["target"][sl]])
    if rank == 0:
        ret["out"], ret["grad"] = outs[0].detach().clone(), net.flat_grad.detach().clone() / world
        ret["rm"] = net.denseLayer.tail.bn.running_mean.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_sync_bn_equal_one_rank_on_the_whole_batch():
    import torch.multiprocessing as mp
    from vsr_b200.metrics import PSNR
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainStep
    path = [p for p in CASES if p.endswith("dufnet16_x4.pt")][0]          # batch of 2
    fx = torch.load(path)
    net = DUFNet(**fx["kwargs"])
    net.load_state_dict(_state(fx))
    net._ops = EmuOps()
    net.train()
    step = MISRTrainStep(net, [torch.nn.L1Loss()], [1.0], [PSNR()], FlatAdam(net.parameters(), lr=1e-3), "acdc")
    _, outs = step.train_step(fx["inputs"], [fx["target"]])
    ret = mp.Manager().dict()
    mp.spawn(_sync_worker, args=(2, 29500 + os.getpid() % 2000, path, ret), nprocs=2, join=True)
    assert _rel(ret["out"], outs[0][:1].detach()) <= 1e-5                 # rank 0's half of the batch
    # the all-reduced gradient (mean over ranks) equals the whole-batch gradient.  (Parameters after the Adam step are
    # not compared: the convolution biases in front of a BatchNorm have zero gradient up to round-off, and Adam's
    # first step moves them by lr * sign(noise).)
    assert float((ret["grad"] - net.flat_grad).abs().max()) <= 1e-5 * float(net.flat_grad.abs().max())
    assert float((ret["rm"] - net.denseLayer.tail.bn.running_mean).abs().max()) <= 1e-6


def test_misr_trainer_epoch_on_cpu():
    """MISRTrainer = AcdcMISRTrainer's epoch loop (base_trainer.py:99-144) over the synthetic MISR dataset
    (acdc_misr_dataset.py contract), through the emulation: the log has the reference's keys and the loss falls."""
    from vsr_b200.data import Dataloader, SyntheticCineDataset
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainer
    ds = SyntheticCineDataset(4, num_frames=7, temporal_order="middle", type="train", num_sequences=1, patch_size=(12, 12),
                              misr=True)
    item = ds[3]
    assert len(item["lr_imgs"]) == 7 and item["lr_imgs"][0].shape == (1, 12, 12) and item["hr_img"].shape == (1, 48, 48)
    ds.data = ds.data[:4]
    loader = Dataloader(ds, batch_size=2, pin_memory=False)
    net = DUFNet(1, 1, 7, 5, 4, "_DenseLayer16")
    net._ops = EmuOps()
    tr = MISRTrainer("cpu", loader, loader, net, [torch.nn.L1Loss()], [1.0], [PSNR(), SSIM()],
                     FlatAdam(net.parameters(), lr=2e-3), None, None, None, 1)
    logs = [tr._run_epoch("training")[0] for _ in range(3)]
    assert list(logs[0]) == ["Loss", "L1Loss", "PSNR", "SSIM"]
    assert logs[2]["Loss"] < logs[0]["Loss"]
    vlog, _, out = tr._run_epoch("validation")
    assert out.shape == (2, 1, 48, 48) and vlog["Loss"] > 0


def test_misr_predictor_on_cpu(tmp_path):
    """MISRPredictor = AcdcMISRPredictor's loop (acdc_misr_predictor.py:31-110) through the emulation: the PSNR it logs
    equals the oracle's PSNR of the oracle's (eval-mode) output, and results.csv has one row per item."""
    from vsr_b200.data import Dataloader, SyntheticCineDataset
    from vsr_b200.metrics import PSNR
    from vsr_b200.runner import MISRPredictor
    ds = SyntheticCineDataset(2, num_frames=7, temporal_order="middle", type="valid", num_sequences=1, patch_size=None,
                              misr=True)
    ds.data = ds.data[:3]
    crop = lambda it: {"lr_imgs": [f[:, :12, :12].contiguous() for f in it["lr_imgs"]],
                       "hr_img": it["hr_img"][:, :24, :24].contiguous(), "index": it["index"]}
    items = [crop(ds[i]) for i in range(3)]

    class Small(torch.utils.data.Dataset):
        data = ds.data

        def __len__(self):
            return 3

        def __getitem__(self, i):
            return items[i]

    net = DUFNet(1, 1, 7, 5, 2, "_DenseLayer16")
    net._ops = EmuOps()
    pred = MISRPredictor("cpu", Dataloader(Small(), batch_size=3, pin_memory=False), net, [torch.nn.L1Loss()], [1.0],
                         [PSNR()], saved_dir=str(tmp_path), exported=True)
    log = pred.predict()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    want = []
    for it in items:
        o = restated.dufnet_forward([f[None] for f in it["lr_imgs"]], sd, 5, 2, training=False)
        want.append(float(restated.psnr(restated.denormalize(o, "acdc"), restated.denormalize(it["hr_img"][None], "acdc"))))
    assert abs(log["PSNR"] - sum(want) / 3) <= 1e-3
    rows = (tmp_path / "results.csv").read_text().strip().splitlines()
    assert rows[0] == "name,PSNR,L1Loss" and len(rows) == 4 and rows[2].startswith("slice00000_frame02,")


def test_config_driven_construction():
    """getattr(src.model.nets, 'DUFNet')(**kwargs) of main.py:56,167-178 resolves to the drop-in"""
    from vsr_b200.config import build_net
    net = build_net({"net": {"name": "DUFNet", "kwargs": dict(in_channels=1, out_channels=1, num_frames=7, size_filter=5,
                                                               upscale_factor=4, backbone="_DenseLayer16")}})
    assert isinstance(net, DUFNet) and net.precision == "fp32"
    assert sum(p.numel() for p in net.parameters()) == net._plan.n_params == net.flat.numel()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 9, 11), (3, 7, 20)], ids=["1x9x11", "3x7x20"])
def test_gpu_ragged_sizes_match_emulation(shape):
    """frame sizes that are no multiple of any pixel box (partial TMA tiles, ragged row splits in the BatchNorm passes):
    the CUDA path in fp32 against the torch emulation on the same device, forward and every gradient; bf16 against fp32."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n, h, w = shape
    g = torch.Generator().manual_seed(11)
    frames = [torch.randn(n, 1, h, w, generator=g).cuda() for _ in range(7)]
    target = torch.randn(n, 1, 2 * h, 2 * w, generator=g).cuda()
    fx = torch.load([p for p in CASES if p.endswith("dufnet16_x2.pt")][0])
    res = {}
    for which in ("cuda", "emu", "bf16"):
        net = DUFNet(1, 1, 7, 5, 2, "_DenseLayer16", precision="bf16" if which == "bf16" else "fp32")
        net.load_state_dict(_state(fx))
        net = net.cuda().train()
        if which == "emu":
            net._ops = EmuOps()
        out = net(frames)
        torch.nn.MSELoss()(out, target).backward()
        res[which] = (out.detach(), net.flat_grad.clone())
    assert _rel(res["cuda"][0], res["emu"][0]) <= 1e-5
    # random inputs, no control of ReLU ties: with 99 rows per frame ONE flipped mask element moves a gradient by ~1e-3
    # of the largest one (see the fixture generator); an indexing bug shows as >= 1e-1
    assert float((res["cuda"][1] - res["emu"][1]).abs().max()) <= 3e-3 * float(res["emu"][1].abs().max())
    assert float((res["cuda"][1] - res["emu"][1]).norm() / res["emu"][1].norm()) <= 3e-3
    assert _rel(res["bf16"][0], res["emu"][0]) <= 5e-2
    assert float((res["bf16"][1] - res["emu"][1]).norm() / res["emu"][1].norm()) <= 0.25


@pytest.mark.gpu
def test_gpu_misr_trainer_with_device_loader():
    """the whole training path on the device: DeviceCineLoader -> MISRTrainer (CUDA-graphed fused step) -> log"""
    from vsr_b200.data import DeviceCineLoader, SyntheticCineDataset
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainer
    ds = SyntheticCineDataset(4, num_frames=7, temporal_order="middle", type="train", num_sequences=1, patch_size=(16, 16),
                              misr=True)
    ds.data = ds.data[:8]
    loader = DeviceCineLoader(ds, "cuda:0", batch_size=4)
    torch.manual_seed(0)
    net = DUFNet(1, 1, 7, 5, 4, "_DenseLayer16", precision="bf16")
    tr = MISRTrainer("cuda:0", loader, loader, net, [torch.nn.L1Loss()], [1.0], [PSNR(), SSIM()],
                     FlatAdam(net.parameters(), lr=1e-3), None, None, None, 1, use_graph=True)
    logs = [tr._run_epoch("training")[0] for _ in range(4)]          # 8 steps: eager, eager, capture, replays
    assert list(logs[0]) == ["Loss", "L1Loss", "PSNR", "SSIM"]
    assert all(torch.isfinite(torch.tensor(list(l.values()))).all() for l in logs)
    assert logs[-1]["Loss"] < logs[0]["Loss"]
    vlog, _, out = tr._run_epoch("validation")
    assert out.shape == (4, 1, 64, 64) and vlog["PSNR"] > 0


def test_checkpoint_interchange_with_live_reference(tmp_path):
    """base_trainer.py:229-237 / base_predictor.py:135-136: a checkpoint written from the drop-in loads (strict) into
    the reference's DUFNet and back, and both nets then agree (training mode: outputs, running buffers)."""
    from oracle import load_reference
    if not load_reference.available():
        pytest.skip("/root/reference not mounted")
    load_reference.load()
    Ref = load_reference._load("src.model.nets.duf_net", "src/model/nets/duf_net.py").DUFNet
    kw = dict(in_channels=1, out_channels=1, num_frames=7, size_filter=5, upscale_factor=3, backbone="_DenseLayer16")
    torch.manual_seed(1)
    ours = DUFNet(**kw)
    ours._ops = EmuOps()
    torch.save({"net": ours.state_dict()}, tmp_path / "ck.pth")
    ref = Ref(**kw)
    ref.load_state_dict(torch.load(tmp_path / "ck.pth")["net"], strict=True)
    x = [torch.randn(2, 1, 9, 10) for _ in range(7)]
    ours.train(), ref.train()
    with torch.no_grad():
        a, b = ours(x), ref(x)
    assert _rel(a, b) <= 2e-5
    sa, sb = ours.state_dict(), ref.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        if "running" in k or "num_batches" in k:
            assert torch.allclose(sa[k].double(), sb[k].double(), rtol=1e-5, atol=1e-6), k
    ours2 = DUFNet(**kw)
    ours2.load_state_dict(ref.state_dict(), strict=True)          # and back, buffers included
    ours2._ops = EmuOps()
    ours2.eval(), ref.eval()
    with torch.no_grad():
        assert _rel(ours2(x), ref(x)) <= 2e-5


@pytest.mark.parametrize("backbone", ["_DenseLayer16", "_DenseLayer28", "_DenseLayer52"])
@pytest.mark.parametrize("bf16", [False, True], ids=["fp32", "bf16"])
def test_plan_invariants(backbone, bf16):
    """structure of the packed tables: every convolution weight / bias is read by exactly one forward slab position
    (so one un-pack pass returns its gradient), every data-gradient table covers the same weights, taps stay inside
    their sources, and tensor-core tables have the shapes the tcgen05 kernels take."""
    import numpy as np
    from vsr_b200.duf import DufPlan
    P = DufPlan(1, 7, 5, 4, backbone, bf16)
    conv_w = [p for n, p in P.params.items() if n.endswith(".weight") and len(p.shape) >= 4 and not n.startswith("head")]
    fwd = np.concatenate([np.stack(L.slabs).reshape(-1) for L in P.fwd.values()])
    bwd = np.concatenate([np.stack(L.slabs).reshape(-1) for L in P.bwd.values()])
    for p in conv_w:
        n = int(np.prod(p.shape))
        for arr in (fwd, bwd):
            hit = arr[(arr >= p.offset) & (arr < p.offset + n)]
            assert len(hit) == n and len(np.unique(hit)) == n, p.name
    assert len(P.unpack_passes) == 1
    biases = [p for n, p in P.params.items() if n.endswith(".bias") and ".bn" not in n and not n.startswith("head")]
    packed_b = np.concatenate([np.asarray(L.bias_idx) for L in P.fwd.values() if L.bias_idx is not None])
    for p in biases:
        assert np.count_nonzero((packed_b >= p.offset) & (packed_b < p.offset + p.shape[0])) == p.shape[0], p.name
    for store in (P.fwd, P.bwd):
        for L in store.values():
            t = L.table
            assert all(c0 % 8 == 0 and c0 >= 0 for _, taps in t.groups for (_, _, _, c0) in taps)
            assert len({len(taps) for _, taps in t.groups}) == 1            # equal groups: one launch geometry
            if bf16:
                assert t.kc == 64 and t.nt % 64 == 0 and 64 <= t.nt <= 256, L.name
    assert P.ntz == (128 if bf16 and P.Gr == 32 else 64 if bf16 else 3 * P.Gr)
