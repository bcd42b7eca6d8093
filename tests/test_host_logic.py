"""Host-logic parity on CPU: the product's plan + engine (tap tables, phase-blocked layout, weight
packing, fused backward schedule) driven through the torch emulation of the kernels
(tests/emu.py) must reproduce the real reference's golden outputs, loss and gradients."""
import glob
import os

import pytest
import torch

from oracle import restated
from tests.emu import EmuOps
from tests.test_oracle import _state
from vsr_b200.nets import DRFNet, DRFSISRNet

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "drfnet_*.pt")))


def _build(fx, dtype=torch.float32):
    net = DRFNet(**fx["kwargs"])
    net.load_state_dict(_state(fx))
    if dtype == torch.float64:
        net = net.double()
    net._ops = EmuOps()
    return net


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-3] for p in CASES])
def test_engine_matches_reference_golden(path):
    fx = torch.load(path)
    net = _build(fx)
    outs = net(fx["inputs"])
    for o, ref in zip(outs, fx["outputs"]):
        assert o.shape == ref.shape
        assert (o - ref).abs().max() <= 2e-5 * ref.abs().max()
    loss = torch.stack([torch.nn.L1Loss()(o, t) for o, t in zip(outs, fx["targets"])]).mean()
    assert abs(float(loss) - float(fx["loss_l1"])) <= 1e-5 * abs(float(fx["loss_l1"]))
    loss.backward()
    got = {k: p.grad for k, p in net.named_parameters()}
    if fx["grads"] is not None:
        gmax = max(float(g.abs().max()) for g in fx["grads"].values())
        for k, g in fx["grads"].items():
            assert got[k].shape == g.shape
            assert (got[k] - g).abs().max() <= 1e-4 * gmax, k
    else:
        for k, dg in fx["grad_digest"].items():
            g = got[k].reshape(-1)
            assert abs(float(g.norm()) - float(dg["norm"])) <= 2e-4 * float(dg["norm"]) + 1e-6, k


def test_engine_fp64_vs_restated_fp64():
    """tight check of every gradient against the fp64 oracle (removes fp32 noise)."""
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g3_x4.pt"))
    net = _build(fx, torch.float64)
    x = [t.double() for t in fx["inputs"]]
    y = [t.double() for t in fx["targets"]]
    outs = net(x)
    loss = torch.stack([((o - t) ** 2).mean() for o, t in zip(outs, y)]).mean()
    loss.backward()
    sd = {k: v.double().clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    ref_outs = restated.drfnet_forward(x, sd, 4)
    ref_loss = torch.stack([((o - t) ** 2).mean() for o, t in zip(ref_outs, y)]).mean()
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-12 * abs(float(ref_loss))
    for k, p in net.named_parameters():
        assert (p.grad - sd[k].grad).abs().max() <= 1e-10 * max(1e-30, float(sd[k].grad.abs().max())) + 1e-14, k


def test_default_init_equals_reference_init():
    """same construction order => same default initialisation under the same seed (goldens were
    made with torch.manual_seed(idx) before constructing the reference class)."""
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g2_x2.pt"))
    torch.manual_seed(0)
    net = DRFNet(**fx["kwargs"])
    sd = net.state_dict()
    assert list(sd) == list(fx["state_dict"])
    for k, v in fx["state_dict"].items():
        if "prelu" in k:
            continue   # the golden generator perturbed the PReLU slopes after construction
        assert torch.equal(sd[k], v), k


def test_state_dict_roundtrip_and_flat_bucket():
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g3_x4.pt"))
    net = DRFNet(**fx["kwargs"])
    net.load_state_dict(fx["state_dict"])
    for k, v in net.state_dict().items():
        assert torch.equal(v, fx["state_dict"][k])
    assert net._is_flat()
    assert net.flat.numel() == sum(p.numel() for p in net.parameters())
    # an optimizer built BEFORE .to()/.double() keeps working (main.py:73 vs base_trainer.py:30)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    net = net.double()
    assert net._is_flat()
    assert all(p.dtype == torch.float64 for g in opt.param_groups for p in g["params"])


def test_bad_upscale_raises_like_reference():
    with pytest.raises(ValueError, match="The upscale factor should be 2, 3, 4 or 8"):
        DRFNet(1, 1, 8, 2, 5)


def test_no_cpu_fallback():
    net = DRFNet(1, 1, 8, 1, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net([torch.zeros(1, 1, 8, 8)])


def test_sisr_net_is_the_vsr_net_on_a_repeated_frame():
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g2_x2.pt"))
    kw = dict(fx["kwargs"])
    sisr = DRFSISRNet(num_steps=3, **kw)
    sisr.load_state_dict(fx["state_dict"])
    sisr._ops = EmuOps()
    x = fx["inputs"][0]
    with torch.no_grad():
        outs = sisr(x)
    ref = restated.drfnet_forward([x, x, x], fx["state_dict"], kw["upscale_factor"])
    for a, b in zip(outs, ref):
        assert (a - b).abs().max() <= 2e-5 * b.abs().max()


def test_device_cine_loader_equals_host_loader():
    """DeviceCineLoader (volumes resident on the device, one gather kernel per batch) yields bit-identical batches to
    Dataloader(SyntheticCineDataset): same windows, flips, crops, normalisation (transforms.py:154-168,321-450)."""
    import torch
    from tests.emu import EmuOps
    from vsr_b200.data import Dataloader, DeviceCineLoader, SyntheticCineDataset
    for misr in (False, True):
        kw = dict(downscale_factor=4, num_frames=5 if not misr else 7, temporal_order="last" if not misr else "middle",
                  type="train", num_sequences=2, patch_size=(12, 10), seed=3, misr=misr)
        host_ds, dev_ds = SyntheticCineDataset(**kw), SyntheticCineDataset(**kw)
        host = iter(Dataloader(host_ds, batch_size=3, pin_memory=False))
        dev = iter(DeviceCineLoader(dev_ds, "cpu", batch_size=3, ops=EmuOps()))
        for _ in range(3):
            a, b = next(host), next(dev)
            for x, y in zip(a["lr_imgs"], b["lr_imgs"]):
                assert x.shape == y.shape == (3, 1, 12, 10) and torch.equal(x, y)
            if misr:
                assert torch.equal(a["hr_img"], b["hr_img"]) and b["hr_img"].shape == (3, 1, 48, 40)
            else:
                assert len(b["hr_imgs"]) == 5 and all(torch.equal(x, y) for x, y in zip(a["hr_imgs"], b["hr_imgs"]))
            assert torch.equal(a["index"], b["index"])


def test_predictor_cardiac_metrics_match_oracle(tmp_path):
    """CardiacPSNR / CardiacSSIM inside the predictor's fused loop (acdc_vsr_predictor.py:134-154, metrics.py:116-165):
    per-sample bounding boxes looked up by patient name, against the oracle on the cropped frames."""
    import pickle
    from pathlib import Path

    import torch
    from oracle import restated
    from tests.emu import EmuOps
    from vsr_b200.data import Dataloader
    from vsr_b200.metrics import PSNR, CardiacPSNR, CardiacSSIM
    from vsr_b200.nets import DRFNet
    from vsr_b200.runner import VSRPredictor
    boxes = {"patient001": (2, 30, 4, 32), "patient002": (0, 24, 8, 36)}
    with open(tmp_path / "boxes.pkl", "wb") as f:
        pickle.dump(boxes, f)
    g = torch.Generator().manual_seed(3)
    items = [{"lr_imgs": [torch.randn(1, 10, 10, generator=g) for _ in range(2)],
              "hr_imgs": [torch.randn(1, 40, 40, generator=g) for _ in range(2)], "index": i} for i in range(2)]

    class Two(torch.utils.data.Dataset):
        data = [(Path("patient001_2d+1d_sequence03.nii.gz"),), (Path("patient002_2d+1d_sequence01.nii.gz"),)]

        def __len__(self):
            return 2

        def __getitem__(self, i):
            return items[i]

    torch.manual_seed(0)
    net = DRFNet(1, 1, 8, 2, 4)
    net._ops = EmuOps()
    metrics = [PSNR(), CardiacPSNR(str(tmp_path / "boxes.pkl")), CardiacSSIM(str(tmp_path / "boxes.pkl"))]
    pred = VSRPredictor("cpu", Dataloader(Two(), batch_size=2, pin_memory=False), net, [torch.nn.L1Loss()], [1.0], metrics,
                        saved_dir=str(tmp_path), exported=True)
    log = pred.predict()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    want_p, want_s = [], []
    for it, who in zip(items, ("patient001", "patient002")):
        outs = restated.drfnet_forward([f[None] for f in it["lr_imgs"]], sd, 4)
        h0, hn, w0, wn = boxes[who]
        for o, y in zip(outs, it["hr_imgs"]):
            a = restated.denormalize(o, "acdc")[..., h0:hn, w0:wn]
            b = restated.denormalize(y[None], "acdc")[..., h0:hn, w0:wn]
            want_p.append(float(restated.psnr(a, b)))
            want_s.append(float(restated.ssim(a, b)))
    assert abs(log["CardiacPSNR"] - sum(want_p) / 4) <= 1e-3
    assert abs(log["CardiacSSIM"] - sum(want_s) / 4) <= 1e-4
    rows = (tmp_path / "results.csv").read_text().strip().splitlines()
    assert rows[0] == "name,PSNR,CardiacPSNR,CardiacSSIM,L1Loss" and len(rows) == 5
    assert rows[1].startswith("patient001_2d_slice03_frame01,")


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver times next to ours): exactly one JSON line on stdout with the
    contract's keys, for the headline workload and for the Conv3d path."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ([], ["--workload", "duf"]):
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra,
                           capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, VSR_REF_FULL="0", VSR_REF_SAMPLE="2"))
        assert r.returncode == 0, r.stderr[-500:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"] == "hr_voxels_per_s_train_step" and d["higher_is_better"] is True
        assert d["value"] > 0 and d["unit"] == "HR voxels/s" and d["vs_baseline"] is None
        assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
        assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert "workload" in d["config"] and "sample" in d["config"]


def _odd_slopes(sd):
    """every PReLU slope of the net replaced by a value from {-0.1, 0, 0.2, -0.35}: nn.PReLU's slope is an
    unconstrained parameter (drf_net.py:56) and may cross zero while training"""
    vals = [-0.1, 0.0, 0.2, -0.35]
    out, i = {}, 0
    for k, v in sd.items():
        if "prelu" in k:
            out[k] = torch.full_like(v, vals[i % len(vals)])
            i += 1
        else:
            out[k] = v.clone()
    return out


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 2e-6), (torch.float32, 1e-4)])
def test_prelu_slopes_zero_and_negative(dtype, tol):
    """outputs and every gradient (the slope gradients included) against the fp64 oracle when slopes are 0 or
    negative: the backward pass recovers the branch and the pre-activation from the stored post-activation
    (csrc/common.cuh `Prelu`)."""
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g3_x4.pt"))
    sd0 = _odd_slopes(fx["state_dict"])
    net = DRFNet(**fx["kwargs"])
    net.load_state_dict(sd0)
    if dtype == torch.float64:
        net = net.double()
    net._ops = EmuOps()
    x = [t.to(dtype) for t in fx["inputs"]]
    y = [t.to(dtype) for t in fx["targets"]]
    outs = net(x)
    loss = torch.stack([((o - t) ** 2).mean() for o, t in zip(outs, y)]).mean()
    loss.backward()
    sd = {k: v.double().clone().requires_grad_(True) for k, v in sd0.items()}
    ref_outs = restated.drfnet_forward([t.double() for t in x], sd, 4)
    ref_loss = torch.stack([((o - t.double()) ** 2).mean() for o, t in zip(ref_outs, y)]).mean()
    ref_loss.backward()
    for o, r in zip(outs, ref_outs):
        assert (o.double() - r).abs().max() <= tol * r.abs().max()
    gmax = max(float(v.grad.abs().max()) for v in sd.values())
    for k, p in net.named_parameters():
        assert (p.grad.double() - sd[k].grad).abs().max() <= tol * gmax, k
    slopes = [k for k in sd if "prelu" in k]
    assert any(float(sd[k].grad.abs()) > 1e-3 * gmax for k in slopes)      # the slope gradients are exercised


def test_shared_weight_gradient_sets():
    """the sets of dense 1x1 layers that go through one wgrad_shared launch: contiguous, within the kernel's limits
    (<= 4 gradient maps, <= 8 sources, <= 8 tensor-memory accumulators), fewest maps read"""
    from vsr_b200.drf_engine import shared_wgrad_sets
    mk = lambda G: [(f"dn{g}_c1", g + 1) for g in range(1, G)]
    names = lambda sets: [[n for n, _ in s] for s in sets]
    assert names(shared_wgrad_sets(mk(3), 4, 8, 8)) == [["dn1_c1", "dn2_c1"]]
    six = shared_wgrad_sets(mk(6), 4, 8, 8)
    assert names(six) == [["dn1_c1", "dn2_c1"], ["dn3_c1", "dn4_c1", "dn5_c1"]]          # 3 + 2 and 6 + 3 = 14 maps (25 apart)
    assert sum(s[-1][1] + len(s) for s in six) == 14
    for G in range(2, 8):
        sets = shared_wgrad_sets(mk(G), 4, 8, 8)
        assert [m for s in sets for m in s] == mk(G)
        for s in sets:
            assert len(s) <= 4 and s[-1][1] <= 8 and sum((nt + 1) // 2 for _, nt in s) <= 8
        assert sum(s[-1][1] + len(s) for s in sets) <= sum(nt + 1 for _, nt in mk(G))
    assert shared_wgrad_sets([], 4, 8, 8) == []


def test_split_ops_plane_cache_invalidation():
    """ops.SplitOps (precision='bf16x3'): a cached plane pair dies when a tap-GEMM output overlaps its fp32 map and ALL pairs
    die when any method that may write a map is fetched; methods that cannot write a map leave the cache alone"""
    from vsr_b200.ops import SplitOps
    ops = SplitOps()
    buf = torch.zeros(4, 8, 8, 64)
    key = lambda t: (t.data_ptr(), t.data_ptr() + t.numel() * 4)
    ops._cache[key(buf[0])] = "p0"
    ops._cache[key(buf[1])] = "p1"
    ops._cache[key(buf[2:4].reshape(-1))] = "p23"
    ops._invalidate(buf[1])                                   # exact range
    assert set(ops._cache.values()) == {"p0", "p23"}
    ops._invalidate(buf[3, 2:3])                              # a slice inside the stacked pair
    assert set(ops._cache.values()) == {"p0"}
    ops._invalidate(None)
    for name in ("tapgemm", "tapgemm_wgrad", "colsum", "wgrad_shared", "gather", "reduce_partials", "conv3x3_first_bwd"):
        getattr(ops, name)
        assert ops._cache, name                               # read-only with respect to maps
    assert ops.split and ops.partials_len > 0 and ops._cache  # plain attributes
    for name in ("act_bwd", "conv3x3_first", "conv3x3_last_bwd", "gather_split", "add", "upsample_linear", "loss_fwd_bwd_seg"):
        ops._cache[key(buf[0])] = "p0"
        getattr(ops, name)
        assert not ops._cache, name                           # may write a map: everything is re-split
    t3 = SplitOps.table3(__import__("vsr_b200.ops", fromlist=["TapTable"]).TapTable(64, 64, [(0, [(0, 0, 0, 0), (1, 1, -1, 64)])]))
    assert t3.groups == [(0, [(0, 0, 0, 0), (1, 1, -1, 64), (8, 0, 0, 0), (9, 1, -1, 64), (0, 0, 0, 0), (1, 1, -1, 64)])]


def test_flat_bucket_check_is_cached_and_invalidated():
    """`_is_flat` runs twice per training step: the walk over the module tree is repeated only after a parameter was
    registered somewhere; a re-pointed `.data`, a replaced Parameter object and `load_state_dict(assign=True)` are all
    detected, in-place loads and dtype moves keep the bucket."""
    import torch.nn as nn
    from vsr_b200.nets import DRFNet
    net = DRFNet(in_channels=1, out_channels=1, num_features=8, num_groups=2, upscale_factor=2)
    assert net._is_flat() and net.__dict__.get("_flat_cache") is not None
    cached = net.__dict__["_flat_cache"]
    assert net._is_flat() and net.__dict__["_flat_cache"] is cached           # second call: pointer comparison only
    p0 = next(net.parameters())
    p0.data = p0.data.clone()
    assert not net._is_flat()
    net._flatten()
    assert net._is_flat()
    net.in_block.conv1.weight = nn.Parameter(net.in_block.conv1.weight.detach().clone())
    assert not net._is_flat()
    net._flatten()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    assert net._is_flat()
    net.load_state_dict(sd, assign=True)
    assert not net._is_flat()
    net._flatten()
    net.double()
    assert net._is_flat() and net.flat.dtype == torch.float64
    for p, ref in zip(net.parameters(), net._plan.params.values()):
        assert p.data_ptr() == net.flat.data_ptr() + ref.offset * net.flat.element_size()
