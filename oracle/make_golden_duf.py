"""Generate tests/golden/dufnet_*.pt by running the REAL reference DUFNet (stub-loaded from /root/reference,
src/model/nets/duf_net.py) on seeded inputs.  Run in the build container only:  python -m oracle.make_golden_duf

The network has ~1.9 M parameters, so the fixtures hold a seed for the weights (duf_fill), the inputs, the
reference output, loss, per-parameter gradient digests and the BatchNorm running buffers after the step.
"""
import os

import torch

from oracle import load_reference
from oracle.make_golden import OUT, grad_digest

CASES = [
    # name, ctor kwargs, N, h, w
    ("dufnet16_x4", dict(in_channels=1, out_channels=1, num_frames=7, size_filter=5, upscale_factor=4, backbone="_DenseLayer16"), 2, 12, 16),
    ("dufnet16_x2", dict(in_channels=1, out_channels=1, num_frames=7, size_filter=5, upscale_factor=2, backbone="_DenseLayer16"), 1, 16, 12),
    ("dufnet16_x3_c2", dict(in_channels=2, out_channels=2, num_frames=7, size_filter=3, upscale_factor=3, backbone="_DenseLayer16"), 1, 12, 12),
    ("dufnet28_x2", dict(in_channels=1, out_channels=1, num_frames=7, size_filter=5, upscale_factor=2, backbone="_DenseLayer28"), 1, 12, 12),
    ("dufnet52_x2", dict(in_channels=1, out_channels=1, num_frames=7, size_filter=5, upscale_factor=2, backbone="_DenseLayer52"), 1, 10, 12),
]


def duf_fill(sd, seed):
    """Deterministic DUFNet weights / BatchNorm buffers that need not be stored."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd):
        v = sd[k]
        leaf = k.split(".")[-1]
        if leaf == "num_batches_tracked":
            out[k] = torch.zeros_like(v)
        elif leaf == "running_var":
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif leaf == "running_mean":
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif ".bn" in k and leaf == "weight":
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() > 1:
            out[k] = torch.randn(v.shape, generator=g) * (1.5 / v[0].numel() ** 0.5)
        else:
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
    return out


def synth_frames(n, t, h, w, r, cin, seed):
    g = torch.Generator().manual_seed(seed)
    hr = torch.rand(n, cin, h * r, w * r, generator=g) * 255
    lrs = []
    for i in range(t):
        f = (hr * (0.7 + 0.05 * i)).round().clamp(0, 255)
        lrs.append((torch.nn.functional.avg_pool2d(f, r).round() - 54.089) / 48.084)
    target = ((hr * (0.7 + 0.05 * (t // 2))).round().clamp(0, 255) - 54.089) / 48.084
    return lrs, target


def main():
    ref = load_reference.load()
    DUFNet = load_reference._load("src.model.nets.duf_net", "src/model/nets/duf_net.py").DUFNet
    import sys
    only = set(sys.argv[1:])
    for idx, (name, kw, n, h, w) in enumerate(CASES):
        if only and name not in only:
            continue
        torch.manual_seed(idx)
        net = DUFNet(**kw)
        net.load_state_dict(duf_fill(net.state_dict(), 3000 + idx))
        # A ReLU input within fp32 round-off of zero makes the gradients of two CORRECT fp32 implementations differ
        # by a whole element (measured: one such element among the 144 rows of the tail BatchNorm moved a slice of
        # the gradients by 2e-3).  Pick, among a few input seeds, the one with the widest margin.
        margins = {}
        hooks = [m.register_forward_pre_hook(lambda mod, inp, name=name_: margins.__setitem__(
            name, min(margins.get(name, 1e9), float(inp[0].abs().min())))) for name_, m in net.named_modules()
            if isinstance(m, torch.nn.ReLU)]
        best = None
        net.train()
        for k in range(16):
            seed = 500 + idx + 100 * k
            lrs, target = synth_frames(n, kw["num_frames"], h, w, kw["upscale_factor"], kw["in_channels"], seed)
            margins.clear()
            with torch.no_grad():
                net(lrs)
            small = min(v for k_, v in margins.items() if not k_.startswith("denseLayer.conv"))
            dense = min(v for k_, v in margins.items() if k_.startswith("denseLayer.conv"))
            score = min(small, 5 * dense)
            if best is None or score > best[0]:
                best = (score, seed, small, dense)
        for hk in hooks:
            hk.remove()
        net.load_state_dict(duf_fill(net.state_dict(), 3000 + idx))      # undo the running-statistics updates
        print(name, "input seed", best[1], "smallest |ReLU input|: tail/heads %.2e, dense layers %.2e" % best[2:])
        lrs, target = synth_frames(n, kw["num_frames"], h, w, kw["upscale_factor"], kw["in_channels"], best[1])
        net.eval()
        with torch.no_grad():
            out_eval = net(lrs)
        net.train()
        out = net(lrs)
        loss = torch.nn.L1Loss()(out, target)
        net.zero_grad()
        loss.backward()
        sd = net.state_dict()
        fx = {"kwargs": kw, "state_seed": 3000 + idx, "state_shapes": {k: tuple(v.shape) for k, v in sd.items()},
              "state_dtypes": {k: v.dtype for k, v in sd.items()},
              "inputs": lrs, "target": target, "output": out.detach().clone(), "output_eval": out_eval.clone(),
              "loss_l1": loss.detach().clone(),
              "grad_digest": {k: grad_digest(p.grad.detach()) for k, p in net.named_parameters()},
              "buffers_after": {k: v.clone() for k, v in sd.items() if "running" in k and ("conv0." in k or "tail" in k)},
              "psnr": ref.PSNR()(ref.denormalize(out.detach(), "acdc"), ref.denormalize(target, "acdc"))}
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(name, os.path.getsize(path) // 1024, "KiB", "loss", float(loss), "psnr", float(fx["psnr"]))


if __name__ == "__main__":
    main()
