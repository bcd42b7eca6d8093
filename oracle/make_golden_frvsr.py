"""Generate tests/golden/frvsrnet_*.pt by running the REAL reference FRVSRNet (stub-loaded from /root/reference,
src/model/nets/frvsr_net.py) on seeded inputs.  The net has 1.9 M parameters (FNet is fixed at 32 .. 256 channels), so the
fixture stores the seed of the weights (oracle.make_golden.seeded_fill), the inputs / targets, the outputs (sr_imgs and
lr_imgs), the two losses of acdc_frvsr_trainer.py:85-88 (with nn.MSELoss as the configured loss) and a digest of every parameter gradient of flow_loss + sr_loss.
Run in the build container only:   python -m oracle.make_golden_frvsr"""
import os

import torch

from oracle import load_reference
from oracle.make_golden import grad_digest, seeded_fill

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [
    # name, kwargs, N, T, h, w
    ("frvsrnet_b2_x4", dict(in_channels=1, out_channels=1, upscale_factor=4, num_resblocks=2), 2, 3, 16, 16),
    ("frvsrnet_b1_x4_pad", dict(in_channels=1, out_channels=1, upscale_factor=4, num_resblocks=1), 1, 2, 12, 20),
]


def _fp32_vs_fp64(sd, inputs, targets, r):
    from oracle import restated
    l1 = torch.nn.MSELoss()      # a smooth loss: the sign of an L1 residual within round-off would be one more kink
    grads = []
    for dt in (torch.float32, torch.float64):
        p = {k: v.to(dt).clone().requires_grad_(True) for k, v in sd.items()}
        xs, ys = [x.to(dt) for x in inputs], [y.to(dt) for y in targets]
        sr, lr = restated.frvsrnet_forward(xs, p, r)
        loss = torch.stack([l1(a, b) for a, b in zip(lr, xs)]).mean() + torch.stack([l1(a, b) for a, b in zip(sr, ys)]).mean()
        loss.backward()
        grads.append({k: v.grad.double() for k, v in p.items()})
    gmax = max(float(g.abs().max()) for g in grads[1].values())
    return max(float((grads[0][k] - grads[1][k]).abs().max()) for k in sd) / gmax


def main():
    ref = load_reference.load()
    for idx, (name, kw, n, t, h, w) in enumerate(CASES):
        torch.manual_seed(300 + idx)
        net = ref.FRVSRNet(**kw)
        order = list(net.state_dict().keys())
        sd = seeded_fill({k: v for k, v in net.state_dict().items()}, 300 + idx)
        net.load_state_dict(sd)
        r = kw["upscale_factor"]
        g = torch.Generator().manual_seed(400 + idx)
        base = torch.randn(n, 1, h, w, generator=g)      # correlated neighbours for the flow net
        inputs = [base + 0.3 * torch.randn(n, 1, h, w, generator=g) for _ in range(t)]
        targets = [torch.randn(n, 1, h * r, w * r, generator=g) for _ in range(t)]
        # d(loss)/d(flow) is a difference of neighbouring pixels of the previous output times w / 2: the flow net's gradients
        # are ill-conditioned in fp32 whoever computes them (printed for information; tests/test_frvsrnet.py measures it)
        print(f"  {name}: fp32 vs float64 gradient error of the reference arithmetic {_fp32_vs_fp64(sd, inputs, targets, r):.2e}")
        sr_imgs, lr_imgs = net(list(inputs))
        l1 = torch.nn.MSELoss()      # a smooth loss: the sign of an L1 residual within round-off would be one more kink
        flow_loss = torch.stack([l1(a, b) for a, b in zip(lr_imgs, inputs)]).mean()     # acdc_frvsr_trainer.py:86
        sr_loss = torch.stack([l1(a, b) for a, b in zip(sr_imgs, targets)]).mean()      # :87
        (flow_loss + sr_loss).backward()
        fx = {"kwargs": kw, "state_seed": 300 + idx, "state_shapes": {k: tuple(sd[k].shape) for k in order},
              "inputs": inputs, "targets": targets, "sr_imgs": [o.detach().clone() for o in sr_imgs],
              "lr_imgs": [o.detach().clone() for o in lr_imgs], "loss": "MSELoss", "flow_loss": flow_loss.detach().clone(),
              "sr_loss": sr_loss.detach().clone(),
              "grad_digest": {k: grad_digest(p.grad) for k, p in net.named_parameters()}}
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print(name, os.path.getsize(os.path.join(OUT, name + ".pt")), "bytes", float(flow_loss), float(sr_loss))


if __name__ == "__main__":
    main()
