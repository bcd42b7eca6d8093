"""FRVSRNet (after RBPNet): oracle restatement vs goldens made by the real reference, host logic of the
drop-in through the kernel emulation (CPU), GPU parity through the C-ABI."""
import glob
import os

import pytest
import torch

from oracle import restated
from oracle.make_golden import seeded_fill

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "frvsrnet_*.pt")))
ids = [os.path.basename(p)[:-3] for p in CASES]


def _state(fx):
    return seeded_fill({k: torch.zeros(s) for k, s in fx["state_shapes"].items()}, fx["state_seed"])


def _losses(sr_imgs, lr_imgs, fx):
    l1 = torch.nn.L1Loss()
    flow = torch.stack([l1(a, b.to(a.device)) for a, b in zip(lr_imgs, fx["inputs"])]).mean()      # acdc_frvsr_trainer.py:86
    sr = torch.stack([l1(a, b.to(a.device)) for a, b in zip(sr_imgs, fx["targets"])]).mean()       # :87
    return flow, sr


def _oracle_grads(fx):
    sd = {k: v.clone().requires_grad_(True) for k, v in _state(fx).items()}
    sr, lr = restated.frvsrnet_forward(fx["inputs"], sd, fx["kwargs"]["upscale_factor"])
    flow_loss, sr_loss = _losses(sr, lr, fx)
    (flow_loss + sr_loss).backward()
    return sr, lr, flow_loss, sr_loss, {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("path", CASES, ids=ids)
def test_oracle_restatement_matches_reference_golden(path):
    fx = torch.load(path)
    sr, lr, flow_loss, sr_loss, grads = _oracle_grads(fx)
    for o, ref in zip(sr, fx["sr_imgs"]):
        assert (o.detach() - ref).abs().max() <= 1e-5 * ref.abs().max()
    for o, ref in zip(lr, fx["lr_imgs"]):
        assert (o.detach() - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert abs(float(flow_loss) - float(fx["flow_loss"])) <= 1e-6 and abs(float(sr_loss) - float(fx["sr_loss"])) <= 1e-6
    for k, dg in fx["grad_digest"].items():
        assert abs(float(grads[k].norm()) - float(dg["norm"])) <= 1e-4 * float(dg["norm"]) + 1e-9, k
        assert (grads[k].reshape(-1)[:16] - dg["head"]).abs().max() <= 1e-4 * float(grads[k].abs().max()) + 1e-9, k
