"""Generate tests/golden/*.pt by running the REAL reference (stub-loaded from /root/reference) on
seeded inputs.  Run in the build container only:   python -m oracle.make_golden

Each fixture holds: ctor kwargs, the reference state_dict (default init under torch.manual_seed),
inputs/targets, the reference outputs, L1 loss and parameter gradients (fp32 run), plus the
reference's PSNR/SSIM of the denormalised outputs.  Kept tiny (< 1.5 MB total).
"""
import os

import torch

from oracle import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, ctor kwargs, N, T, h, w
    ("drfnet_f8_g2_x2", dict(in_channels=1, out_channels=1, num_features=8, num_groups=2, upscale_factor=2), 2, 3, 12, 10),
    ("drfnet_f8_g2_x3", dict(in_channels=1, out_channels=1, num_features=8, num_groups=2, upscale_factor=3), 1, 2, 12, 12),
    ("drfnet_f8_g3_x4", dict(in_channels=1, out_channels=1, num_features=8, num_groups=3, upscale_factor=4), 2, 2, 12, 12),
    ("drfnet_f8_g1_x8", dict(in_channels=1, out_channels=1, num_features=8, num_groups=1, upscale_factor=8), 1, 2, 11, 11),
    ("drfnet_f64_g2_x4", dict(in_channels=1, out_channels=1, num_features=64, num_groups=2, upscale_factor=4), 1, 2, 12, 12),
]


def synth(n, t, h, w, r, seed):
    g = torch.Generator().manual_seed(seed)
    hr = torch.rand(n, 1, h * r, w * r, generator=g) * 255
    frames_hr, frames_lr = [], []
    for i in range(t):
        f = (hr * (0.8 + 0.05 * i)).round().clamp(0, 255)
        lr = torch.nn.functional.avg_pool2d(f, r).round()
        frames_hr.append((f - 54.089) / 48.084)
        frames_lr.append((lr - 54.089) / 48.084)
    return frames_lr, frames_hr


def seeded_fill(sd, seed):
    """Deterministic weights that need not be stored: N(0, 1/fan_in) per tensor in key order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd):
        v = sd[k]
        if k.endswith("prelu.weight") or "prelu" in k.split(".")[-2]:
            out[k] = 0.2 + 0.05 * torch.randn(v.shape, generator=g)
        elif v.dim() > 1:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (1.5 / fan_in ** 0.5)
        else:
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
    return out


def grad_digest(g):
    f = g.reshape(-1)
    return {"norm": f.norm().clone(), "sum": f.sum().clone(), "head": f[:16].clone()}


def main():
    ref = load_reference.load()
    os.makedirs(OUT, exist_ok=True)
    for idx, (name, kw, n, t, h, w) in enumerate(CASES):
        torch.manual_seed(idx)
        net = ref.DRFNet(**kw)
        big = kw["num_features"] >= 32
        if big:
            net.load_state_dict(seeded_fill(net.state_dict(), seed=1000 + idx))
        # perturb PReLU slopes and biases so that every parameter matters
        if not big:
            with torch.no_grad():
                for k, p in net.named_parameters():
                    if "prelu" in k:
                        p.add_(0.05 * torch.randn_like(p))
        r = kw["upscale_factor"]
        lr, hr = synth(n, t, h, w, r, seed=100 + idx)
        outs = net(lr)
        loss = torch.stack([torch.nn.L1Loss()(o, y) for o, y in zip(outs, hr)]).mean()
        net.zero_grad()
        loss.backward()
        den_o = [ref.denormalize(o.detach(), "acdc") for o in outs]
        den_t = [ref.denormalize(y, "acdc") for y in hr]
        psnr = torch.stack([ref.PSNR()(a, b) for a, b in zip(den_o, den_t)]).mean()
        ssim = torch.stack([ref.SSIM()(a, b) for a, b in zip(den_o, den_t)]).mean()
        fx = {
            "kwargs": kw,
            "state_dict": None if big else {k: v.detach().clone() for k, v in net.state_dict().items()},
            "state_seed": 1000 + idx if big else None,
            "inputs": lr, "targets": hr,
            "outputs": [o.detach().clone() for o in outs],
            "loss_l1": loss.detach().clone(),
            "grads": None if big else {k: p.grad.detach().clone() for k, p in net.named_parameters()},
            "grad_digest": {k: grad_digest(p.grad.detach()) for k, p in net.named_parameters()} if big else None,
            "psnr": psnr, "ssim": ssim,
            "ssim_window": ref.SSIM().weight.detach().clone(),
        }
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB", "loss", float(loss), "psnr", float(psnr), "ssim", float(ssim))

    # the secondary nets on the same kernels: SRFBNet (srfb_net.py) and EDSRNet (edsr_net.py)
    for name, cls, kw, n, h, w in [
        ("srfbnet_f8_g2_x4", "SRFBNet", dict(in_channels=1, out_channels=1, num_steps=3, num_features=8, num_groups=2, upscale_factor=4), 2, 10, 12),
        ("srfbnet_f8_g2_x2", "SRFBNet", dict(in_channels=1, out_channels=1, num_steps=2, num_features=8, num_groups=2, upscale_factor=2), 1, 12, 9),
        ("edsrnet_f8_b3_x4", "EDSRNet", dict(in_channels=1, out_channels=1, num_resblocks=3, num_features=8, upscale_factor=4), 2, 10, 12),
        ("edsrnet_f8_b2_x3", "EDSRNet", dict(in_channels=1, out_channels=1, num_resblocks=2, num_features=8, upscale_factor=3), 1, 11, 9),
        ("srfbnet_f64_g2_x4", "SRFBNet", dict(in_channels=1, out_channels=1, num_steps=2, num_features=64, num_groups=2, upscale_factor=4), 1, 12, 12),
        ("edsrnet_f64_b2_x4", "EDSRNet", dict(in_channels=1, out_channels=1, num_resblocks=2, num_features=64, upscale_factor=4), 1, 12, 12),
    ]:
        torch.manual_seed(len(name))
        net = getattr(ref, cls)(**kw)
        big = kw["num_features"] >= 32
        if big:
            net.load_state_dict(seeded_fill(net.state_dict(), seed=2000 + len(name)))
        r = kw["upscale_factor"]
        lr, hr = synth(n, 1, h, w, r, seed=300 + len(name))
        x, y = lr[0], hr[0]
        out = net(x)
        outs = out if isinstance(out, list) else [out]
        loss = torch.stack([torch.nn.L1Loss()(o, y) for o in outs]).mean()
        net.zero_grad()
        loss.backward()
        fx = {"cls": cls, "kwargs": kw, "input": x, "target": y,
              "state_dict": None if big else {k: v.detach().clone() for k, v in net.state_dict().items()},
              "state_seed": 2000 + len(name) if big else None,
              "state_shapes": {k: tuple(v.shape) for k, v in net.state_dict().items()},
              "outputs": [o.detach().clone() for o in outs], "loss_l1": loss.detach().clone(),
              "grads": None if big else {k: p.grad.detach().clone() for k, p in net.named_parameters()},
              "grad_digest": {k: grad_digest(p.grad.detach()) for k, p in net.named_parameters()} if big else None}
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))

    # losses / metrics known-answer vectors from the reference classes
    g = torch.Generator().manual_seed(7)
    a = torch.randn(3, 1, 40, 36, generator=g)
    b = a + 0.3 * torch.randn(3, 1, 40, 36, generator=g)
    da, db = ref.denormalize(a, "acdc"), ref.denormalize(b, "dsb15")
    fx = {
        "a": a, "b": b, "den_acdc_a": da, "den_dsb15_b": db,
        "huber_0.5": ref.HuberLoss(0.5)(a, b), "charbonnier_1e-6": ref.CharbonnierLoss(1e-6)(a, b),
        "l1": torch.nn.L1Loss()(a, b), "mse": torch.nn.MSELoss()(a, b),
        "psnr_mean": ref.PSNR()(da, db), "psnr_per": ref.PSNR(size_average=False)(da, db),
        "ssim_mean": ref.SSIM()(da, db), "ssim_per": ref.SSIM(size_average=False)(da, db),
    }
    torch.save(fx, os.path.join(OUT, "losses_metrics.pt"))
    print("losses_metrics", {k: float(v) for k, v in fx.items() if v.dim() == 0})


def metrics3d():
    """SSIM(dim=3) and the Cardiac* wrappers (metrics.py:51-165) of the real reference on a seeded volume."""
    import pickle
    import tempfile
    ref = load_reference.load()
    g = torch.Generator().manual_seed(11)
    a = torch.randn(2, 1, 13, 20, 18, generator=g)
    b = a + 0.3 * torch.randn(2, 1, 13, 20, 18, generator=g)
    da, db = ref.denormalize(a, "acdc"), ref.denormalize(b, "acdc")
    box = {"patient007": (3, 31, 2, 30)}
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as f:
        pickle.dump(box, f)
    a2, b2 = da[:, :, 0].repeat(1, 1, 2, 2), db[:, :, 0].repeat(1, 1, 2, 2)       # [2,1,40,36] images
    fx = {
        "a": a, "b": b,
        "ssim3_mean": ref.SSIM(dim=3)(da, db), "ssim3_per": ref.SSIM(dim=3, size_average=False)(da, db),
        "ssim3_window": ref.SSIM(dim=3).weight.detach().clone(),
        "box": box, "img_a": a2, "img_b": b2,
        "cardiac_psnr": ref.CardiacPSNR(f.name)(a2, b2, "patient007"),
        "cardiac_ssim": ref.CardiacSSIM(f.name)(a2, b2, "patient007"),
    }
    os.unlink(f.name)
    torch.save(fx, os.path.join(OUT, "metrics3d.pt"))
    print("metrics3d", {k: float(v) for k, v in fx.items() if torch.is_tensor(v) and v.dim() == 0})


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "metrics3d":      # added later: leaves the other fixtures untouched
        metrics3d()
    else:
        main()
        metrics3d()
