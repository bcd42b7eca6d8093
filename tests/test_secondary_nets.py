"""SRFBNet / EDSRNet: oracle restatement vs the real reference's goldens,
host logic on CPU through the kernel emulation, and GPU parity through the C-ABI."""
import glob
import os

import pytest
import torch

from oracle import restated
from oracle.make_golden import seeded_fill
from tests.emu import EmuOps
from vsr_b200.edsr import EDSRNet
from vsr_b200.nets import SRFBNet

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "srfbnet_*.pt")) + glob.glob(os.path.join(GOLDEN, "edsrnet_*.pt")))
SMALL = [p for p in CASES if "_f8_" in p]
BIG = [p for p in CASES if "_f64_" in p]
CLS = {"SRFBNet": SRFBNet, "EDSRNet": EDSRNet}
ids = lambda ps: [os.path.basename(p)[:-3] for p in ps]


def _state(fx):
    if fx["state_dict"] is not None:
        return fx["state_dict"]
    return seeded_fill({k: torch.zeros(s) for k, s in fx["state_shapes"].items()}, fx["state_seed"])


def _oracle(fx, sd):
    kw = fx["kwargs"]
    if fx["cls"] == "SRFBNet":
        return restated.srfbnet_forward(fx["input"], sd, kw["upscale_factor"], kw["num_steps"])
    return [restated.edsrnet_forward(fx["input"], sd, kw["upscale_factor"])]


def _check(net, fx, x, y, out_tol, grad_tol):
    out = net(x)
    outs = out if isinstance(out, list) else [out]
    for o, ref in zip(outs, fx["outputs"]):
        assert o.shape == ref.shape
        assert (o.detach().cpu() - ref).abs().max() <= out_tol * ref.abs().max()
    loss = torch.stack([torch.nn.L1Loss()(o, y) for o in outs]).mean()
    loss.backward()
    got = {k: p.grad.detach().cpu() for k, p in net.named_parameters()}
    if fx["grads"] is not None:
        gmax = max(float(g.abs().max()) for g in fx["grads"].values())
        for k, g in fx["grads"].items():
            assert (got[k] - g).abs().max() <= grad_tol * gmax, k
    else:
        num = den = 0.0
        for k, dg in fx["grad_digest"].items():
            num += abs(float(got[k].norm()) - float(dg["norm"]))
            den += float(dg["norm"])
        assert num / den <= grad_tol
    return float(loss)


@pytest.mark.parametrize("path", CASES, ids=ids(CASES))
def test_restated_matches_reference_golden(path):
    fx = torch.load(path)
    outs = _oracle(fx, _state(fx))
    for o, ref in zip(outs, fx["outputs"]):
        assert (o - ref).abs().max() <= 1e-5 * ref.abs().max()


@pytest.mark.parametrize("path", SMALL, ids=ids(SMALL))
def test_host_logic_matches_reference_golden(path):
    fx = torch.load(path)
    net = CLS[fx["cls"]](**fx["kwargs"])
    net.load_state_dict(_state(fx))
    net._ops = EmuOps()
    loss = _check(net, fx, fx["input"], fx["target"], 2e-5, 1e-4)
    assert abs(loss - float(fx["loss_l1"])) <= 1e-5 * float(fx["loss_l1"])


def test_state_dict_keys_match_reference():
    for path in SMALL:
        fx = torch.load(path)
        net = CLS[fx["cls"]](**fx["kwargs"])
        assert list(net.state_dict()) == list(fx["state_dict"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", SMALL, ids=ids(SMALL))
def test_gpu_fp32_matches_reference_golden(path):
    fx = torch.load(path)
    net = CLS[fx["cls"]](precision="fp32", **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.to("cuda")
    _check(net, fx, fx["input"].cuda(), fx["target"].cuda(), 1e-4, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("path", BIG, ids=ids(BIG))
def test_gpu_bf16_close_to_reference(path):
    fx = torch.load(path)
    net = CLS[fx["cls"]](precision="bf16", **fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.to("cuda")
    _check(net, fx, fx["input"].cuda(), fx["target"].cuda(), 5e-2, 5e-2)
    with torch.no_grad():
        out = net(fx["input"].cuda())
    outs = out if isinstance(out, list) else [out]
    den = lambda t: restated.denormalize(t, "acdc")
    for o, ref in zip(outs, fx["outputs"]):
        p_ref = restated.psnr(den(ref), den(fx["target"]))
        p_got = restated.psnr(den(o.cpu()), den(fx["target"]))
        assert abs(float(p_ref) - float(p_got)) <= 0.05


# ---- fused single-image training steps (acdc_sisr_trainer.py, acdc_sisr_srfb_trainer.py) ---------------------------
def _sisr_step_vs_oracle(path, device, precision="fp32", steps=2, use_graph=False, w_tol=2e-5, l_tol=2e-5):
    """SISRTrainStep / SISRSRFBTrainStep against the oracle stepped with torch.optim.Adam on the same weights and batch:
    the loss of every step, the metrics of the last output, the weights after `steps` steps (eps = 1e-4: see
    tests/test_trainstep_gpu.py on why)"""
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import SISRSRFBTrainStep, SISRTrainStep
    fx = torch.load(path)
    kw = fx["kwargs"]
    net = CLS[fx["cls"]](precision=precision, **kw)
    net.load_state_dict(_state(fx))
    if device == "cpu":
        net._ops = EmuOps()
    net = net.to(device)
    opt = FlatAdam(net.parameters(), lr=1e-3, eps=1e-4)
    cls = SISRSRFBTrainStep if fx["cls"] == "SRFBNet" else SISRTrainStep
    step = cls(net, [torch.nn.L1Loss(), torch.nn.MSELoss()], [1.0, 0.5], [PSNR().to(device), SSIM().to(device)], opt, "acdc",
               use_graph=use_graph)
    sd = {k: v.clone().requires_grad_(True) for k, v in _state(fx).items()}
    ref_opt = torch.optim.Adam(list(sd.values()), lr=1e-3, eps=1e-4)
    x, y = fx["input"], fx["target"]
    if y.shape[-1] < 11:                                   # SSIM needs an 11 x 11 window: tile the small fixtures
        rep = -(-11 // y.shape[-1])
        x, y = x.repeat(1, 1, rep, rep), y.repeat(1, 1, rep, rep)
    for _ in range(steps):
        acc = torch.zeros(5, device=device)
        lv, _ = step.train_step([x.to(device)], [y.to(device)], acc)
        outs = _oracle(dict(fx, input=x), sd)
        l1 = torch.stack([restated.l1_loss(o, y) for o in outs]).mean()
        mse = torch.stack([restated.mse_loss(o, y) for o in outs]).mean()
        psnr, ssim = restated.vsr_metrics([outs[-1].detach()], [y])
        ref_opt.zero_grad()
        (l1 + 0.5 * mse).backward()
        ref_opt.step()
        assert abs(float(lv[0]) - float(l1)) <= l_tol * float(l1) and abs(float(lv[1]) - float(mse)) <= l_tol * float(mse)
        assert abs(float(acc[0]) - float(l1 + 0.5 * mse)) <= l_tol * float(l1 + 0.5 * mse)
        assert abs(float(acc[3]) - float(psnr)) <= (2e-3 if precision == "fp32" else 5e-2)
        assert abs(float(acc[4]) - float(ssim)) <= (1e-4 if precision == "fp32" else 5e-3)
    wmax = max(float(v.abs().max()) for v in sd.values())
    for k, p in net.named_parameters():
        assert (p.data.cpu() - sd[k].data).abs().max() <= w_tol * wmax, k


@pytest.mark.parametrize("path", SMALL, ids=ids(SMALL))
def test_sisr_train_steps_match_reference_step_host_logic(path):
    _sisr_step_vs_oracle(path, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("path", SMALL, ids=ids(SMALL))
def test_sisr_train_steps_gpu_graphed(path):
    """the same through the C-ABI with the step replayed as a CUDA graph (2 eager + capture + replay)"""
    _sisr_step_vs_oracle(path, "cuda", steps=4, use_graph=True, w_tol=1e-4, l_tol=1e-4)


def test_sisr_trainers_epoch_loop_and_batch_keys(tmp_path):
    """SISRTrainer / SISRSRFBTrainer: `lr_img` / `hr_img` batches, one epoch of training + validation, log keys"""
    from vsr_b200.metrics import PSNR
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import SISRSRFBTrainer, SISRTrainer

    class Loader(list):
        batch_size = 2

    g = torch.Generator().manual_seed(3)
    for path in (SMALL[0], SMALL[-1]):
        fx = torch.load(path)
        net = CLS[fx["cls"]](**fx["kwargs"])
        net._ops = EmuOps()
        r = fx["kwargs"]["upscale_factor"]
        batches = Loader({"lr_img": torch.randn(2, 1, 6, 6, generator=g), "hr_img": torch.randn(2, 1, 6 * r, 6 * r, generator=g)}
                         for _ in range(2))
        trainer_cls = SISRSRFBTrainer if fx["cls"] == "SRFBNet" else SISRTrainer
        tr = trainer_cls("cpu", batches, batches, net, [torch.nn.L1Loss()], [1.0], [PSNR()], FlatAdam(net.parameters(), lr=1e-3),
                         None, None, None, 1)
        log, _, outputs = tr._run_epoch("training")
        assert set(log) == {"Loss", "L1Loss", "PSNR"} and all(v == v for v in log.values())
        vlog, _, _ = tr._run_epoch("validation")
        assert vlog["Loss"] < log["Loss"] * 1.5
