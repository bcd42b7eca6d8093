"""GPU tests at the shapes BASELINE.json names besides the bench workload (configs[0] and configs[2])."""
import pytest
import torch

from oracle import restated
from vsr_b200.metrics import PSNR, SSIM
from vsr_b200.nets import DRFNet

pytestmark = pytest.mark.gpu


def _cine(n, t, h, r, seed):
    """ACDC-like synthetic sequence: integers in [0,255], box-downscaled LR, ACDC normalisation."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n, 1, h * r, h * r, generator=g) * 255
    lrs, hrs = [], []
    for i in range(t):
        f = (base * (0.75 + 0.02 * i)).round().clamp(0, 255)
        lrs.append(((torch.nn.functional.avg_pool2d(f, r).round() - 54.089) / 48.084).contiguous())
        hrs.append(((f - 54.089) / 48.084).contiguous())
    return lrs, hrs


def test_config1_acdc_crop_x2_training_window_matches_oracle():
    """configs[0]: the repo's default SR model on one ACDC-shaped cropped cine (128x128 x 10 slices, x2), the
    trainer's T=5 window (acdc_vsr_dataset.py:66-75): forward + L1 + backward in strict fp32 mode against the
    CPU oracle on the same weights and inputs - outputs / loss / gradients within the fp32 bar (1e-4)."""
    kw = dict(in_channels=1, out_channels=1, num_features=64, num_groups=6, upscale_factor=2)
    lrs, hrs = _cine(10, 5, 64, 2, seed=5)
    torch.manual_seed(1)
    net = DRFNet(precision="fp32", **kw)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    ref_outs = restated.drfnet_forward(lrs, sd, 2)
    ref_loss = restated.vsr_loss(ref_outs, hrs, restated.l1_loss)
    ref_loss.backward()
    net = net.to("cuda")
    outs = net([x.cuda() for x in lrs])
    loss = torch.stack([torch.nn.L1Loss()(o, y.cuda()) for o, y in zip(outs, hrs)]).mean()
    loss.backward()
    for o, r in zip(outs, ref_outs):
        assert (o.detach().cpu() - r.detach()).abs().max() <= 1e-4 * r.detach().abs().max()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    gmax = max(float(v.grad.abs().max()) for v in sd.values())
    num = sum(float(((p.grad.cpu() - sd[k].grad) ** 2).sum()) for k, p in net.named_parameters()) ** 0.5
    den = sum(float((v.grad ** 2).sum()) for v in sd.values()) ** 0.5
    assert num / den <= 1e-4
    for k, p in net.named_parameters():
        assert (p.grad.cpu() - sd[k].grad).abs().max() <= 1e-4 * gmax, k


def test_config3_dsb15_full_fov_inference_metrics_on_device():
    """configs[2]: full-FOV DSB15-shaped cine (256x256 x 12 slices x 30 frames, x4): the 12 slices are the
    batch ("tiles"), all 30 frames in one no-grad call, PSNR / SSIM of every frame on the device.  The bf16
    tcgen05 path against this library's strict fp32 mode: PSNR within 0.05 dB per frame; the device metrics
    against the CPU oracle's metrics of the same outputs."""
    kw = dict(in_channels=1, out_channels=1, num_features=64, num_groups=6, upscale_factor=4)
    lrs, hrs = _cine(12, 30, 64, 4, seed=6)
    x, y = [t.cuda() for t in lrs], [t.cuda() for t in hrs]
    torch.manual_seed(2)
    ref = DRFNet(precision="fp32", **kw).to("cuda").eval()
    net = DRFNet(precision="bf16", **kw)
    net.load_state_dict(ref.state_dict())
    net = net.to("cuda").eval()
    with torch.no_grad():
        o_ref, o_bf = ref(x), net(x)
    assert len(o_bf) == 30 and o_bf[0].shape == (12, 1, 256, 256)
    psnr, ssim = PSNR(dataset="dsb15").cuda(), SSIM(dataset="dsb15").cuda()
    p_ref = torch.stack([psnr(o, t) for o, t in zip(o_ref, y)])
    p_bf = torch.stack([psnr(o, t) for o, t in zip(o_bf, y)])
    s_bf = torch.stack([ssim(o, t) for o, t in zip(o_bf, y)])
    assert (p_ref - p_bf).abs().max().item() <= 0.05, (p_ref - p_bf).abs().max().item()
    for i in (0, 29):       # the fused device metrics against the oracle on the same outputs
        den_o, den_t = restated.denormalize(o_bf[i].cpu(), "dsb15"), restated.denormalize(hrs[i], "dsb15")
        assert abs(float(restated.psnr(den_o, den_t)) - float(p_bf[i])) <= 1e-3
        assert abs(float(restated.ssim(den_o, den_t)) - float(s_bf[i])) <= 2e-5


def test_predictor_on_device_batched(tmp_path):
    """VSRPredictor on the GPU with a batch of 3 synthetic validation sequences: the log equals the mean of the
    per-frame rows it exported, and the rows equal the stand-alone metric modules on the network outputs."""
    import csv
    from vsr_b200.data import Dataloader, SyntheticCineDataset
    from vsr_b200.runner import VSRPredictor
    ds = SyntheticCineDataset(4, type="valid", dataset="acdc", num_sequences=3, seed=3)
    loader = Dataloader(ds, batch_size=3, shuffle=False, num_workers=0)
    torch.manual_seed(4)
    net = DRFNet(1, 1, 64, 2, 4, precision="bf16")
    pred = VSRPredictor("cuda", loader, net, [torch.nn.L1Loss()], [1.0], [PSNR(), SSIM()], saved_dir=str(tmp_path),
                        exported=True, dataset="acdc")
    log = pred.predict()
    rows = list(csv.reader(open(tmp_path / "results.csv")))
    T = ds.T
    assert rows[0] == ["name", "PSNR", "SSIM", "L1Loss"] and len(rows) == 1 + 3 * T
    vals = torch.tensor([[float(v) for v in r[1:]] for r in rows[1:]])
    assert abs(float(vals[:, 0].mean()) - log["PSNR"]) <= 1e-3
    assert abs(float(vals[:, 1].mean()) - log["SSIM"]) <= 1e-5
    assert abs(float(vals[:, 2].mean()) - log["L1Loss"]) <= 1e-5 and abs(log["Loss"] - log["L1Loss"]) <= 1e-6
    batch = next(iter(loader))
    x, y = [t.cuda() for t in batch["lr_imgs"]], [t.cuda() for t in batch["hr_imgs"]]
    with torch.no_grad():
        outs = pred.net(x)
    p = PSNR(size_average=False, dataset="acdc").cuda()(outs[2], y[2])
    assert abs(float(p[1]) - float(vals[1 * T + 2, 0])) <= 1e-3      # sequence 1, frame 3
