"""GPU parity of the bandwidth-bound kernels against the torch emulation / the oracle."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import restated
from tests.emu import EmuOps
from vsr_b200.drf_plan import phase_table

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _ops():
    from vsr_b200.ops import cuda_ops
    return cuda_ops()


def _g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cin", [1, 3])
def test_conv3x3_first_and_bwd(dtype, cin):
    ops, emu, g = _ops(), EmuOps(), _g(1)
    n, h, w, cout = 3, 9, 11, 64
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.3
    b = torch.randn(cout, device="cuda", generator=g)
    a = torch.tensor([0.2], device="cuda")
    ys = []
    for o in (ops, emu):
        y = torch.zeros(n, h, w, cout, device="cuda", dtype=dtype)
        o.conv3x3_first(x, wt, b, a, y)
        ys.append(y.float())
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert (ys[0] - ys[1]).abs().max() <= tol * ys[1].abs().max()
    dz = torch.randn(n, h, w, cout, device="cuda", generator=g).to(dtype)
    rs = []
    for o in (ops, emu):
        dw, db = torch.ones_like(wt), torch.ones_like(b)
        ws = torch.empty(max(16, o.conv3x3_first_bwd_workspace(x, cout)) // 4 + 4, device="cuda")
        o.conv3x3_first_bwd(x, dz, dw, db, True, ws)
        rs.append((dw, db))
    assert (rs[0][0] - rs[1][0]).abs().max() <= 1e-4 * rs[1][0].abs().max()
    assert (rs[0][1] - rs[1][1]).abs().max() <= 1e-4 * rs[1][1].abs().max()


@pytest.mark.parametrize("cout,n,h,w,slope", [(256, 5, 32, 32, 0.2), (256, 3, 19, 23, -0.3), (128, 2, 33, 17, 0.0), (64, 7, 9, 11, 0.25),
                                             (32, 1, 16, 17, None)])
def test_conv3x3_first_tensor_core_path(cout, n, h, w, slope):
    """firstconv_mma.cu (bf16 maps, split-bf16 operands on mma.sync) against an fp64 convolution, and against this library's
    CUDA-core kernels (VSR_FC_SIMT=1): forward within one bf16 rounding of the exact result, weight / bias gradient at fp32
    accuracy; several CTA tiles, ragged last tile, slope > 0 / < 0 / = 0 / no activation"""
    ops, g = _ops(), _g(5)
    x = torch.randn(n, 1, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, 1, 3, 3, device="cuda", generator=g) * 0.3
    b = torch.randn(cout, device="cuda", generator=g)
    a = None if slope is None else torch.tensor([slope], device="cuda")
    z = F.conv2d(x.double(), wt.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    want = z if slope is None else torch.where(z > 0, z, z * (slope if slope != 0.0 else 2.0 ** -24))
    dz = torch.randn(n, h, w, cout, device="cuda", generator=g).to(torch.bfloat16)
    res = {}
    for simt in ("0", "1"):
        os.environ["VSR_FC_SIMT"] = simt
        ops.lib.vsr_reload_tunables()
        try:
            y = torch.zeros(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
            ops.conv3x3_first(x, wt, b, a, y)
            dw, db = torch.zeros_like(wt), torch.zeros_like(b)
            ws = torch.empty(ops.conv3x3_first_bwd_workspace(x, cout) // 4 + 4, device="cuda")
            ops.conv3x3_first_bwd(x, dz, dw, db, False, ws)
            res[simt] = (y.double(), dw, db)
        finally:
            del os.environ["VSR_FC_SIMT"]
            ops.lib.vsr_reload_tunables()
    xd = F.unfold(x.double(), 3, padding=1).transpose(1, 2).reshape(-1, 9)                   # [pixels][9]
    dw_ref = (dz.double().reshape(-1, cout).t() @ xd).reshape(cout, 1, 3, 3)
    db_ref = dz.double().sum((0, 1, 2))
    for simt in ("0", "1"):
        y, dw, db = res[simt]
        rnd = 2.0 ** -8 if slope is None or slope >= 0 else 2.0 ** -6                         # one bf16 rounding (+ the sign tag in the LSB)
        # (+ the dropped xl * wl term of the split operands: 2^-16 of the sum of |x w|)
        assert ((y - want).abs() <= rnd * want.abs() + (1e-4 if simt == "0" else 1e-6)).all(), simt
        assert (dw.double() - dw_ref).abs().max() <= 2e-5 * dw_ref.abs().max(), simt
        assert (db.double() - db_ref).abs().max() <= 2e-5 * db_ref.abs().max(), simt


@pytest.mark.parametrize("r,c,h,w", [(4, 64, 11, 9), (3, 64, 13, 12), (2, 128, 20, 17), (8, 64, 5, 9)])
def test_conv3x3_last_tensor_core_path_several_tiles(r, c, h, w):
    """bf16 maps through lastconv_mma.cu with more than one 32x32 output tile per image and partial tiles on both axes"""
    test_conv3x3_last_and_bwd(torch.bfloat16, r, c, n=3, h=h, w=w)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("r,c", [(2, 64), (4, 64), (3, 8), (8, 16), (4, 128)])
def test_conv3x3_last_and_bwd(dtype, r, c, n=2, h=5, w=6):
    ops, emu, g = _ops(), EmuOps(), _g(2)
    ph = phase_table(r)
    x = torch.randn(n, h, w, r * r * c, device="cuda", generator=g).to(dtype)
    wt = torch.randn(1, c, 3, 3, device="cuda", generator=g) * 0.1
    b = torch.randn(1, device="cuda", generator=g)
    # bf16 maps with 64 / 128 channels take the tensor-core kernels (lastconv_mma.cu): like every tensor-core layer their
    # operands are bf16 (weights and dy are rounded inside the kernel, accumulation is fp32) - the emulation gets the same
    # rounded operands, so that what is compared is the arithmetic, at the fp32 tolerances
    mma = dtype == torch.bfloat16 and c in (64, 128)
    rnd = (lambda t: t.bfloat16().float()) if mma else (lambda t: t)
    ys = []
    for o in (ops, emu):
        y = torch.zeros(n, 1, h * r, w * r, device="cuda")
        o.conv3x3_last(x, r, c, ph, wt if o is ops else rnd(wt), b, y)
        ys.append(y)
    assert (ys[0] - ys[1]).abs().max() <= 1e-4 * ys[1].abs().max()
    dy = torch.randn(n, 1, h * r, w * r, device="cuda", generator=g)
    rs = []
    for o in (ops, emu):
        dx = torch.zeros_like(x)
        dw, db = torch.zeros_like(wt), torch.zeros_like(b)
        ws = torch.empty(max(16, o.conv3x3_last_bwd_workspace(x, r, c, 1)) // 4 + 4, device="cuda")
        if o is ops:
            o.conv3x3_last_bwd(x, r, c, ph, wt, dy, dx, dw, db, False, ws)
        else:
            o.conv3x3_last_bwd(x, r, c, ph, rnd(wt), rnd(dy), dx, dw, db, False, ws)
            if mma:
                db.copy_(dy.sum().view(1))          # the bias gradient sums the fp32 dy
        rs.append((dx.float(), dw, db))
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert (rs[0][0] - rs[1][0]).abs().max() <= tol * rs[1][0].abs().max()
    assert (rs[0][1] - rs[1][1]).abs().max() <= 1e-3 * rs[1][1].abs().max()
    assert (rs[0][2] - rs[1][2]).abs().max() <= 1e-3 * rs[1][2].abs().max()


@pytest.mark.parametrize("kind,param", [(0, 0.0), (1, 0.0), (2, 1e-6), (3, 0.5)])
def test_loss_fwd_bwd(kind, param):
    ops, g = _ops(), _g(3)
    fx = torch.load(os.path.join(GOLDEN, "losses_metrics.pt"))
    a, b = fx["a"].cuda(), fx["b"].cuda()
    part = torch.zeros(ops.partials_len, device="cuda")
    grad = torch.empty_like(a)
    ops.loss_fwd_bwd(a, b, kind, param, 1.0 / a.numel(), part, grad)
    want = [fx["l1"], fx["mse"], fx["charbonnier_1e-6"], fx["huber_0.5"]][kind]
    got = part.sum().item() / a.numel()
    assert abs(got - float(want)) <= 1e-5 * abs(float(want))
    ac = fx["a"].clone().requires_grad_(True)
    fn = [restated.l1_loss, restated.mse_loss, lambda o, t: restated.charbonnier_loss(o, t, 1e-6),
          lambda o, t: restated.huber_loss(o, t, 0.5)][kind]
    fn(ac, fx["b"]).backward()
    assert (grad.cpu() - ac.grad).abs().max() <= 1e-5 * ac.grad.abs().max()


def test_psnr_ssim_match_reference_known_answers():
    ops = _ops()
    fx = torch.load(os.path.join(GOLDEN, "losses_metrics.pt"))
    a, b = fx["a"].cuda(), fx["b"].cuda()
    n = a.shape[0]
    ws = torch.empty(ops.metric_workspace(n, a.numel() // n) // 4 + 4, device="cuda")
    # the golden uses acdc stats for `a` and dsb15 stats for `b`; check each denormalisation via
    # pre-denormalised inputs (std <= 0 disables it) and the fused path on same-stat inputs
    da, db = fx["den_acdc_a"].cuda(), fx["den_dsb15_b"].cuda()
    out = torch.empty(n, device="cuda")
    ops.psnr(da, db, 0.0, -1.0, 255.0, out, ws)
    assert torch.allclose(out.cpu(), fx["psnr_per"], rtol=1e-5)
    win = restated.ssim_window_1d().cuda()
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    ops.ssim(da, db, win, 0.0, -1.0, c1, c2, out, ws)
    assert torch.allclose(out.cpu(), fx["ssim_per"], atol=2e-5)
    # fused denormalize (bit-exact rounding): acdc on both
    ops.psnr(a, b, 54.089, 48.084, 255.0, out, ws)
    want = restated.psnr(restated.denormalize(fx["a"], "acdc"), restated.denormalize(fx["b"], "acdc"), size_average=False)
    assert torch.allclose(out.cpu(), want, rtol=1e-5)
    ops.ssim(a, b, win, 54.089, 48.084, c1, c2, out, ws)
    want = restated.ssim(restated.denormalize(fx["a"], "acdc"), restated.denormalize(fx["b"], "acdc"), size_average=False)
    assert torch.allclose(out.cpu(), want, atol=2e-5)


def test_ssim3d_and_cardiac_metrics_match_reference_golden(tmp_path):
    """SSIM(dim=3) (metrics.py:51-113) and CardiacPSNR / CardiacSSIM (:116-165) through the drop-in classes."""
    import pickle
    from vsr_b200.metrics import SSIM, CardiacPSNR, CardiacSSIM
    fx = torch.load(os.path.join(GOLDEN, "metrics3d.pt"))
    a, b = fx["a"].cuda(), fx["b"].cuda()
    m = SSIM(dim=3, dataset="acdc").cuda()                      # fused denormalize
    assert abs(float(m(a, b)) - float(fx["ssim3_mean"])) <= 2e-5
    per = SSIM(dim=3, size_average=False, dataset="acdc").cuda()(a, b)
    assert torch.allclose(per.cpu(), fx["ssim3_per"], atol=2e-5)
    assert torch.allclose(m.weight.cpu(), fx["ssim3_window"], atol=1e-9)
    with pytest.raises(ValueError):
        SSIM(dim=3).cuda()(a[:, :, 0], b[:, :, 0])
    path = tmp_path / "coords.pkl"
    with open(path, "wb") as f:
        pickle.dump(fx["box"], f)
    ia, ib = fx["img_a"].cuda(), fx["img_b"].cuda()
    assert abs(float(CardiacPSNR(str(path)).cuda()(ia, ib, "patient007")) - float(fx["cardiac_psnr"])) <= 1e-3
    assert abs(float(CardiacSSIM(str(path)).cuda()(ia, ib, "patient007")) - float(fx["cardiac_ssim"])) <= 2e-5


def test_small_helpers():
    ops, emu, g = _ops(), EmuOps(), _g(4)
    src = torch.randn(1000, device="cuda", generator=g)
    idx = torch.randint(-1, 1000, (4096,), device="cuda", generator=g, dtype=torch.int32)
    for dt in (torch.float32, torch.bfloat16):
        a, b = torch.zeros(4096, device="cuda", dtype=dt), torch.zeros(4096, device="cuda", dtype=dt)
        ops.gather(src, idx, a); emu.gather(src, idx, b)
        assert torch.equal(a, b)
    a, b = torch.ones(4096, device="cuda"), torch.ones(4096, device="cuda")
    ops.gather_add(src, idx, a); emu.gather_add(src, idx, b)
    assert torch.equal(a, b)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(777, 64, device="cuda", generator=g).to(dt)
        d1, d2 = torch.ones(64, device="cuda"), torch.ones(64, device="cuda")
        ws = torch.empty(ops.colsum_workspace(777, 64) // 4 + 4, device="cuda")
        ops.colsum(x, 777, 64, d1, True, ws); emu.colsum(x, 777, 64, d2, True, ws)
        assert (d1 - d2).abs().max() <= 1e-3
        y = torch.randn(8, 16, 64, device="cuda", generator=g).to(dt)
        dy = torch.randn(8, 16, 64, device="cuda", generator=g).to(dt)
        sl = torch.tensor([0.3], device="cuda")
        o1, o2 = torch.zeros_like(y), torch.zeros_like(y)
        p1, p2 = torch.zeros(ops.partials_len, device="cuda"), torch.zeros(ops.partials_len, device="cuda")
        ops.act_bwd(dy, y, o1, sl, p1); emu.act_bwd(dy, y, o2, sl, p2)
        assert (o1.float() - o2.float()).abs().max() <= 1e-2
        assert abs(p1.sum().item() - p2.sum().item()) <= 1e-2 * max(1.0, abs(p2.sum().item()))
        s1, s2 = torch.zeros_like(y), torch.zeros_like(y)
        ops.add(y, dy, s1); emu.add(y, dy, s2)
        assert torch.equal(s1, s2)
    parts = torch.randn(5, ops.partials_len, device="cuda", generator=g)
    rd = torch.tensor([0, 2, 2, 1, 0], device="cuda", dtype=torch.int32)
    d1, d2 = torch.zeros(3, device="cuda"), torch.zeros(3, device="cuda")
    ops.reduce_partials(parts, 5, rd, d1); emu.reduce_partials(parts, 5, rd, d2)
    assert (d1 - d2).abs().max() <= 1e-3


def test_adam_flat_matches_torch_adam():
    ops, g = _ops(), _g(5)
    p0 = torch.randn(10000, device="cuda", generator=g)
    grads = [torch.randn(10000, device="cuda", generator=g) for _ in range(3)]
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    q, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for i, gr in enumerate(grads):
        p.grad = gr.clone()
        opt.step()
        ops.adam_flat(q, gr, m, v, 1e-2, 0.9, 0.999, 1e-8, 0.0, i + 1)
    assert (q - p.data).abs().max() <= 1e-6


@pytest.mark.parametrize("r", [2, 3])
def test_pixel_shuffle(r):
    ops, g = _ops(), _g(6)
    x = torch.randn(2, 3 * r * r, 5, 7, device="cuda", generator=g)
    y = torch.empty(2, 3, 5 * r, 7 * r, device="cuda")
    ops.pixel_shuffle(x, y, r)
    assert torch.equal(y, F.pixel_shuffle(x, r))
    z = torch.empty_like(x)
    ops.pixel_shuffle(y, z, r, inverse=True)
    assert torch.equal(z, x)


@pytest.mark.parametrize("ac", [False, True])
@pytest.mark.parametrize("three_d", [False, True])
def test_upsample_linear_fwd_bwd(ac, three_d):
    ops, g = _ops(), _g(7)
    if three_d:
        x = torch.randn(2, 3, 4, 5, 6, device="cuda", generator=g)
        size, mode = (8, 15, 12), "trilinear"
    else:
        x = torch.randn(2, 3, 7, 9, device="cuda", generator=g)
        size, mode = (28, 36), "bilinear"
    xr = x.clone().requires_grad_(True)
    want = F.interpolate(xr, size=size, mode=mode, align_corners=ac)
    y = torch.empty_like(want)
    ops.upsample_linear(x, y, ac)
    assert (y - want).abs().max() <= 1e-5
    dy = torch.randn(want.shape, device="cuda", generator=g)
    want.backward(dy)
    dx = torch.empty_like(x)
    ops.upsample_linear_bwd(dy, dx, ac)
    assert (dx - xr.grad).abs().max() <= 1e-4


@pytest.mark.parametrize("ac", [False, True])
@pytest.mark.parametrize("shape,size", [((2, 2, 37, 300), (148, 1200)),      # x4, two x tiles, rows in several blocks
                                        ((1, 1, 9, 17), (72, 136)),          # x8
                                        ((1, 2, 10, 11), (23, 30)),          # non-integer ratio
                                        ((1, 1, 20, 20), (10, 7)),           # down-scaling: the generic gather kernels
                                        ((1, 2, 6, 20, 33), (12, 40, 66)),   # trilinear x2
                                        ((1, 1, 3, 5, 300), (7, 20, 1100))]) # trilinear, mixed ratios, two x tiles
def test_upsample_linear_staged_kernels(ac, shape, size):
    """the staged forward / transposed-stencil backward kernels (and their generic fallbacks) against torch on sizes that
    exercise several x tiles, row blocks, x8, non-integer ratios and both align modes"""
    ops, g = _ops(), _g(17)
    x = torch.randn(*shape, device="cuda", generator=g)
    mode = "trilinear" if len(shape) == 5 else "bilinear"
    xr = x.clone().requires_grad_(True)
    want = F.interpolate(xr, size=size, mode=mode, align_corners=ac)
    y = torch.empty_like(want)
    ops.upsample_linear(x, y, ac)
    assert (y - want).abs().max() <= 1e-5
    dy = torch.randn(want.shape, device="cuda", generator=g)
    want.backward(dy)
    dx = torch.empty_like(x)
    ops.upsample_linear_bwd(dy, dx, ac)
    assert (dx - xr.grad).abs().max() <= 2e-4 * max(1.0, float(xr.grad.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,r", [((3, 1, 32, 32), 4),          # the global skip of config 2
                                     ((2, 2, 1, 1), 4), ((1, 1, 1, 70), 2), ((1, 1, 70, 1), 8),   # degenerate axes, clamps everywhere
                                     ((1, 3, 67, 95), 2), ((2, 1, 129, 61), 4), ((1, 1, 40, 31), 8),   # several warps / strips, ragged
                                     ((1, 2, 1, 9, 11), 2), ((2, 1, 5, 33, 35), 2), ((1, 1, 3, 18, 31), 4)])   # trilinear
def test_upsample_integer_ratio_kernels(shape, r):
    """resample_int.cu (align_corners = False, ratio 2 / 4 / 8: periodic weights, one input column per thread) against torch,
    forward and backward, and against this library's generic kernels (VSR_UP_GENERIC=1)"""
    ops, g = _ops(), _g(23)
    x = torch.randn(*shape, device="cuda", generator=g)
    mode = "trilinear" if len(shape) == 5 else "bilinear"
    xr = x.clone().requires_grad_(True)
    want = F.interpolate(xr, scale_factor=r, mode=mode, align_corners=False)
    y = torch.full_like(want, float("nan"))
    ops.upsample_linear(x, y, False)
    assert (y - want).abs().max() <= 1e-5
    dy = torch.randn(want.shape, device="cuda", generator=g)
    want.backward(dy)
    dx = torch.full_like(x, float("nan"))
    ops.upsample_linear_bwd(dy, dx, False)
    assert (dx - xr.grad).abs().max() <= 2e-5 * max(1.0, float(xr.grad.abs().max()))
    os.environ["VSR_UP_GENERIC"] = "1"
    ops.lib.vsr_reload_tunables()
    try:
        y2, dx2 = torch.empty_like(y), torch.empty_like(dx)
        ops.upsample_linear(x, y2, False)
        ops.upsample_linear_bwd(dy, dx2, False)
    finally:
        del os.environ["VSR_UP_GENERIC"]
        ops.lib.vsr_reload_tunables()
    assert (y - y2).abs().max() <= 1e-5 and (dx - dx2).abs().max() <= 2e-4 * max(1.0, float(dx2.abs().max()))


@pytest.mark.gpu
def test_device_cine_loader_bit_exact_vs_host_loader():
    """vsr_cine_gather (window + flips + crop + Normalize + collate on the device) against the host loader: bit-exact"""
    import torch
    from vsr_b200.data import Dataloader, DeviceCineLoader, SyntheticCineDataset
    for misr, r in ((False, 4), (True, 2), (False, 3)):
        kw = dict(downscale_factor=r, num_frames=7 if misr else 5, temporal_order="middle" if misr else "last",
                  type="train", num_sequences=2, patch_size=(13, 10), seed=5, misr=misr)
        host_ds, dev_ds = SyntheticCineDataset(**kw), SyntheticCineDataset(**kw)
        host = iter(Dataloader(host_ds, batch_size=4, pin_memory=False))
        dev = iter(DeviceCineLoader(dev_ds, "cuda", batch_size=4))
        for _ in range(3):
            a, b = next(host), next(dev)
            assert all(torch.equal(x, y.cpu()) for x, y in zip(a["lr_imgs"], b["lr_imgs"]))
            if misr:
                assert torch.equal(a["hr_img"], b["hr_img"].cpu())
            else:
                assert all(torch.equal(x, y.cpu()) for x, y in zip(a["hr_imgs"], b["hr_imgs"]))


def test_device_downscale_bit_exact_vs_reference_downscale():
    """csrc/downscale.cu (k-space truncation as two complex FP64 GEMMs per frame + integer-ratio bicubic + round + clip)
    against the golden outputs of the REAL reference Downscale class (acdc_preprocess.py:102-180): the same integers, for
    x2 / x3 / x4, square and non-square frames; and a synthetic dataset built with device=... equals the host-built one."""
    import os
    from vsr_b200.data import SyntheticCineDataset, downscale_device
    for c in torch.load(os.path.join(os.path.dirname(__file__), "golden", "downscale.pt")):
        got = downscale_device(c["hr"].float().cuda(), c["r"])
        assert torch.equal(got.cpu(), c["lr"].float()), (c["r"], tuple(c["hr"].shape), int((got.cpu() != c["lr"].float()).sum()))
    kw = dict(downscale_factor=4, num_frames=3, type="train", num_sequences=1, patch_size=(16, 16), seed=9)
    host, dev = SyntheticCineDataset(**kw), SyntheticCineDataset(device="cuda", **kw)
    import numpy as np
    assert all(np.array_equal(a, b) for a, b in zip(host.lr, dev.lr))
