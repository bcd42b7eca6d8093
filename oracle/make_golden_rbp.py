"""Generate tests/golden/rbpnet_*.pt by running the REAL reference RBPNet (stub-loaded from /root/reference,
src/model/nets/rbp_net.py) on seeded inputs: state_dict (default initialisation, PReLU slopes perturbed), inputs, target,
output, L1 loss and every parameter gradient.  Run in the build container only:   python -m oracle.make_golden_rbp"""
import os

import torch

from oracle import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [
    # name, kwargs, N, h, w
    ("rbpnet_b16_f8_x2", dict(in_channels=1, out_channels=1, base_filter=16, feat=8, num_stages=3, num_resblocks=2,
                              num_frames=3, upscale_factor=2), 2, 10, 12),
    ("rbpnet_b16_f8_x4", dict(in_channels=1, out_channels=1, base_filter=16, feat=8, num_stages=3, num_resblocks=1,
                              num_frames=4, upscale_factor=4), 1, 9, 8),
    ("rbpnet_b8_f8_x3", dict(in_channels=1, out_channels=1, base_filter=8, feat=8, num_stages=3, num_resblocks=1,
                             num_frames=3, upscale_factor=3), 1, 8, 8),
    ("rbpnet_b64_f64_x4", dict(in_channels=1, out_channels=1, base_filter=64, feat=64, num_stages=3, num_resblocks=1,
                               num_frames=3, upscale_factor=4), 1, 8, 8),
]


def main():
    ref = load_reference.load()
    for idx, (name, kw, n, h, w) in enumerate(CASES):
        torch.manual_seed(100 + idx)
        net = ref.RBPNet(**kw)
        g = torch.Generator().manual_seed(200 + idx)
        with torch.no_grad():
            for k, p in net.named_parameters():
                if k.endswith("act.weight"):
                    p.copy_(0.25 + 0.1 * torch.randn(p.shape, generator=g))
        r = kw["upscale_factor"]
        inputs = [torch.randn(n, 1, h, w, generator=g) for _ in range(kw["num_frames"])]
        target = torch.randn(n, 1, h * r, w * r, generator=g)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        out = net(list(inputs))
        loss = torch.nn.L1Loss()(out, target)
        loss.backward()
        fx = {"kwargs": kw, "state_dict": sd, "inputs": inputs, "target": target, "output": out.detach().clone(),
              "loss_l1": loss.detach().clone()}
        big = sum(v.numel() for v in sd.values()) > 400000
        if big:     # keep the fixture small: weights re-created from the seed, gradients as digests
            fx["state_dict"], fx["state_seed"] = None, 100 + idx
            fx["slopes"] = {k: v for k, v in sd.items() if k.endswith("act.weight")}
            fx["grads"] = None
            fx["grad_digest"] = {k: {"norm": p.grad.norm().clone(), "head": p.grad.reshape(-1)[:8].clone()} for k, p in net.named_parameters()}
        else:
            fx["grads"] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print(name, os.path.getsize(os.path.join(OUT, name + ".pt")), "bytes", "loss", float(loss))


if __name__ == "__main__":
    main()
