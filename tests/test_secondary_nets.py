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
