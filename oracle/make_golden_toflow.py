"""Generate tests/golden/toflownet_*.pt by running the REAL reference TOFlowNet (stub-loaded from /root/reference,
src/model/nets/toflow_net.py) in training mode on seeded inputs: the seed of the weights (oracle.make_golden.seeded_fill;
BatchNorm scales around 1, running variances positive), the inputs / target, the output, the MSE loss, a digest of every
parameter gradient and the BatchNorm running buffers after the forward pass (every SpyNet block runs once per neighbour).
Run in the build container only:   python -m oracle.make_golden_toflow"""
import os

import torch

from oracle import load_reference
from oracle.make_golden import grad_digest, seeded_fill

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [
    # name, kwargs, N, h, w
    ("toflownet_t3_x2", dict(in_channels=1, out_channels=1, num_frames=3, upscale_factor=2), 2, 16, 16),
    ("toflownet_t4_x4_pad", dict(in_channels=1, out_channels=1, num_frames=4, upscale_factor=4), 1, 6, 10),
]


def fill(sd, seed):
    out = seeded_fill({k: v for k, v in sd.items() if v.dtype.is_floating_point}, seed)
    g = torch.Generator().manual_seed(seed + 1)
    for k, v in sd.items():
        if not v.dtype.is_floating_point:
            out[k] = v.clone()                                   # num_batches_tracked
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif ".block." in k and k.endswith(".weight") and v.dim() == 1:
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)   # BatchNorm scale
    return out


def main():
    ref = load_reference.load()
    for idx, (name, kw, n, h, w) in enumerate(CASES):
        torch.manual_seed(500 + idx)
        net = ref.TOFlowNet(**kw)
        order = list(net.state_dict().keys())
        sd = fill(net.state_dict(), 500 + idx)
        net.load_state_dict(sd)
        net.train()
        g = torch.Generator().manual_seed(600 + idx)
        r = kw["upscale_factor"]
        base = torch.randn(n, 1, h, w, generator=g)
        inputs = [base + 0.3 * torch.randn(n, 1, h, w, generator=g) for _ in range(kw["num_frames"])]
        target = torch.randn(n, 1, h * r, w * r, generator=g)
        out = net(list(inputs))
        loss = torch.nn.MSELoss()(out, target)
        loss.backward()
        after = net.state_dict()
        fx = {"kwargs": kw, "state_seed": 500 + idx, "state_shapes": {k: tuple(sd[k].shape) for k in order},
              "state_dtypes": {k: sd[k].dtype for k in order}, "inputs": inputs, "target": target,
              "output": out.detach().clone(), "loss": loss.detach().clone(),
              "grad_digest": {k: grad_digest(p.grad) for k, p in net.named_parameters()},
              "buffers_after": {k: after[k].detach().clone() for k in order if "running_" in k or "num_batches" in k}}
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print(name, os.path.getsize(os.path.join(OUT, name + ".pt")), "bytes", float(loss))


if __name__ == "__main__":
    main()
