"""Pins oracle/restated.py against (a) the golden vectors produced by the REAL reference
(tests/golden, oracle/make_golden.py) and (b) the live reference where /root/reference exists."""
import glob
import os

import pytest
import torch

from oracle import load_reference, restated
from oracle.make_golden import seeded_fill

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "drfnet_*.pt")))


def _state(fx):
    if fx["state_dict"] is not None:
        return fx["state_dict"]
    shapes = _shapes(fx["kwargs"])
    return seeded_fill(shapes, fx["state_seed"])


def _shapes(kw):
    """state_dict key -> zero tensor of the reference shape, built without the reference."""
    F, G, r, ci, co = kw["num_features"], kw["num_groups"], kw["upscale_factor"], kw["in_channels"], kw["out_channels"]
    k = restated.PROJ[r][0]
    sd = {}

    def conv(p, o, i, ks):
        sd[p + ".weight"] = torch.zeros(o, i, ks, ks)
        sd[p + ".bias"] = torch.zeros(o)

    def prelu(p):
        sd[p + ".weight"] = torch.zeros(1)

    conv("in_block.conv1", 4 * F, ci, 3); prelu("in_block.prelu1")
    conv("in_block.conv2", F, 4 * F, 1); prelu("in_block.prelu2")
    conv("f_block.in_block.conv", F, 2 * F, 1); prelu("f_block.in_block.prelu")
    for g in range(G):
        if g == 0:
            conv("f_block.up_blocks.0.deconv", F, F, k); prelu("f_block.up_blocks.0.prelu")
            conv("f_block.down_blocks.0.conv", F, F, k); prelu("f_block.down_blocks.0.prelu")
        else:
            conv(f"f_block.up_blocks.{g}.conv1", F, F * (g + 1), 1); prelu(f"f_block.up_blocks.{g}.prelu1")
            conv(f"f_block.up_blocks.{g}.deconv2", F, F, k); prelu(f"f_block.up_blocks.{g}.prelu2")
            conv(f"f_block.down_blocks.{g}.conv1", F, F * (g + 1), 1); prelu(f"f_block.down_blocks.{g}.prelu1")
            conv(f"f_block.down_blocks.{g}.conv2", F, F, k); prelu(f"f_block.down_blocks.{g}.prelu2")
    conv("f_block.out_block.conv", F, F * G, 1); prelu("f_block.out_block.prelu")
    if r == 3:
        conv("out_block.conv1", 9 * F, F, 3); conv("out_block.conv2", co, F, 3)
    else:
        n = {2: 1, 4: 2, 8: 3}[r]
        for i in range(n):
            conv(f"out_block.conv{i + 1}", 4 * F, F, 3)
        conv(f"out_block.conv{n + 1}", co, F, 3)
    return sd


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-3] for p in CASES])
def test_restated_matches_golden(path):
    fx = torch.load(path)
    sd = {k: v.clone().requires_grad_(True) for k, v in _state(fx).items()}
    r = fx["kwargs"]["upscale_factor"]
    outs = restated.drfnet_forward(fx["inputs"], sd, r)
    for o, ref in zip(outs, fx["outputs"]):
        assert (o - ref).abs().max() <= 1e-5 * ref.abs().max()
    loss = restated.vsr_loss(outs, fx["targets"], restated.l1_loss)
    assert abs(float(loss) - float(fx["loss_l1"])) <= 1e-6 * abs(float(fx["loss_l1"]))
    loss.backward()
    if fx["grads"] is not None:
        gmax = max(float(g.abs().max()) for g in fx["grads"].values())
        for k, g in fx["grads"].items():
            assert (sd[k].grad - g).abs().max() <= 2e-5 * gmax, k
    else:
        for k, dg in fx["grad_digest"].items():
            g = sd[k].grad.reshape(-1)
            assert abs(float(g.norm()) - float(dg["norm"])) <= 1e-4 * float(dg["norm"]) + 1e-7, k
            assert (g[:16] - dg["head"]).abs().max() <= 1e-4 * float(dg["norm"]) + 1e-7, k
    psnr, ssim = restated.vsr_metrics([o.detach() for o in outs], fx["targets"])
    assert abs(float(psnr) - float(fx["psnr"])) < 1e-4
    assert abs(float(ssim) - float(fx["ssim"])) < 1e-5


def test_shapes_helper_matches_golden_state_dict():
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g3_x4.pt"))
    sh = _shapes(fx["kwargs"])
    assert sorted(sh) == sorted(fx["state_dict"])
    for k in sh:
        assert sh[k].shape == fx["state_dict"][k].shape


def test_losses_metrics_known_answers():
    fx = torch.load(os.path.join(GOLDEN, "losses_metrics.pt"))
    a, b = fx["a"], fx["b"]
    assert torch.equal(restated.denormalize(a, "acdc"), fx["den_acdc_a"])
    assert torch.equal(restated.denormalize(b, "dsb15"), fx["den_dsb15_b"])
    close = lambda x, y, tol=1e-6: abs(float(x) - float(y)) <= tol * max(1.0, abs(float(y)))
    assert close(restated.l1_loss(a, b), fx["l1"])
    assert close(restated.mse_loss(a, b), fx["mse"])
    assert close(restated.huber_loss(a, b, 0.5), fx["huber_0.5"])
    assert close(restated.charbonnier_loss(a, b, 1e-6), fx["charbonnier_1e-6"])
    da, db = fx["den_acdc_a"], fx["den_dsb15_b"]
    assert close(restated.psnr(da, db), fx["psnr_mean"])
    assert torch.allclose(restated.psnr(da, db, size_average=False), fx["psnr_per"], rtol=1e-6)
    assert close(restated.ssim(da, db), fx["ssim_mean"], 1e-5)
    assert torch.allclose(restated.ssim(da, db, size_average=False), fx["ssim_per"], atol=1e-5)


def test_ssim3d_and_cardiac_match_reference_golden():
    # metrics.py:51-113 with dim=3 and the Cardiac* wrappers (:116-165), values from the real reference
    fx = torch.load(os.path.join(GOLDEN, "metrics3d.pt"))
    da, db = restated.denormalize(fx["a"], "acdc"), restated.denormalize(fx["b"], "acdc")
    assert abs(float(restated.ssim(da, db, dim=3)) - float(fx["ssim3_mean"])) <= 1e-5
    assert torch.allclose(restated.ssim(da, db, dim=3, size_average=False), fx["ssim3_per"], atol=1e-5)
    w1 = restated.ssim_window_1d()
    assert torch.allclose(torch.einsum("i,j,k->ijk", w1, w1, w1), fx["ssim3_window"][0, 0], atol=1e-9)
    box = fx["box"]["patient007"]
    assert abs(float(restated.cardiac(restated.psnr, fx["img_a"], fx["img_b"], box)) - float(fx["cardiac_psnr"])) <= 1e-4
    assert abs(float(restated.cardiac(restated.ssim, fx["img_a"], fx["img_b"], box)) - float(fx["cardiac_ssim"])) <= 1e-5


def test_ssim_window_is_the_reference_quirk():
    # metrics.py:74: exp(-((i-5)/(2 sigma))^2) — effective sigma = 1.5*sqrt(2), not 1.5
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g2_x2.pt"))
    w1 = restated.ssim_window_1d()
    assert torch.allclose(torch.outer(w1, w1), fx["ssim_window"][0, 0], atol=1e-8)


@pytest.mark.skipif(not load_reference.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("r,G", [(2, 2), (3, 1), (4, 3), (8, 1)])
def test_restated_matches_live_reference(r, G):
    ref = load_reference.load()
    torch.manual_seed(r)
    net = ref.DRFNet(1, 1, 12, G, r)
    x = [torch.randn(2, 1, 9, 8) for _ in range(3)]
    want = net(x)
    got = restated.drfnet_forward(x, dict(net.state_dict()), r)
    for a, b in zip(got, want):
        assert (a - b).abs().max() <= 1e-5 * b.abs().max()


@pytest.mark.skipif(not load_reference.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("backbone,r,cin", [("_DenseLayer16", 4, 1), ("_DenseLayer28", 3, 2), ("_DenseLayer52", 2, 1)])
def test_restated_dufnet_matches_live_reference(backbone, r, cin):
    """oracle.restated.dufnet_forward against the reference's own DUFNet (duf_net.py:9-99), default initialisation,
    training (batch statistics) and evaluation (running statistics) mode"""
    load_reference.load()
    DUFNet = load_reference._load("src.model.nets.duf_net", "src/model/nets/duf_net.py").DUFNet
    torch.manual_seed(r)
    net = DUFNet(cin, cin, 7, 5, r, backbone)
    x = [torch.randn(2, cin, 9, 8) for _ in range(7)]
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net.train()
    want = net(x)
    got = restated.dufnet_forward(x, sd, 5, r, training=True)
    assert (got - want).abs().max() <= 1e-5 * want.abs().max()
    net.eval()
    want = net(x)
    got = restated.dufnet_forward(x, dict(net.state_dict()), 5, r, training=False)
    assert (got - want).abs().max() <= 1e-5 * want.abs().max()


def test_host_downscale_restatement_equals_reference_downscale():
    """vsr_b200.data.downscale (the numpy / cv2 restatement the synthetic dataset uses on the host) against the REAL
    reference Downscale class (acdc_preprocess.py:102-180; tests/golden/downscale.pt, oracle/make_golden_downscale.py):
    every pixel of every case is the same integer"""
    import numpy as np
    from vsr_b200.data import downscale, lowpass_matrix
    for c in torch.load(os.path.join(GOLDEN, "downscale.pt")):
        hr, lr = c["hr"].float().numpy(), c["lr"].float().numpy()
        got = np.stack([downscale(f, c["r"]) for f in hr])
        assert np.array_equal(got, lr)
    P = lowpass_matrix(12, 4)
    assert P.shape == (12, 12, 2) and abs(P[..., 0].sum(axis=1) - 1.0).max() < 1e-12     # the DC component passes unchanged
