"""TOFlowNet: oracle restatement vs goldens made by the real reference, host logic of the drop-in
through the kernel emulation (CPU, fp32 and exact in float64), GPU parity through the C-ABI."""
import glob
import os

import pytest
import torch

from oracle import restated
from oracle.make_golden_toflow import fill

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "toflownet_*.pt")))
ids = [os.path.basename(p)[:-3] for p in CASES]


def _state(fx):
    return fill({k: torch.zeros(s, dtype=fx["state_dtypes"][k]) for k, s in fx["state_shapes"].items()}, fx["state_seed"])


def _oracle(fx, dtype=torch.float32):
    sd = {k: (v.to(dtype).requires_grad_(True) if v.dtype.is_floating_point and "running_" not in k else v.clone())
          for k, v in _state(fx).items()}
    buffers = {k: v.to(dtype).clone() for k, v in sd.items() if "running_" in k}
    out = restated.toflownet_forward([x.to(dtype) for x in fx["inputs"]], sd, fx["kwargs"]["upscale_factor"], True, buffers)
    loss = restated.mse_loss(out, fx["target"].to(dtype))
    loss.backward()
    return out, loss, {k: v.grad for k, v in sd.items() if v.dtype.is_floating_point and v.requires_grad}, buffers


@pytest.mark.parametrize("path", CASES, ids=ids)
def test_oracle_restatement_matches_reference_golden(path):
    fx = torch.load(path)
    out, loss, grads, buffers = _oracle(fx)
    assert (out.detach() - fx["output"]).abs().max() <= 1e-5 * fx["output"].abs().max()
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * float(fx["loss"])
    for k, dg in fx["grad_digest"].items():
        assert abs(float(grads[k].norm()) - float(dg["norm"])) <= 1e-3 * float(dg["norm"]) + 1e-8, k
    for k, v in buffers.items():
        assert (v - fx["buffers_after"][k]).abs().max() <= 1e-5 * max(1.0, float(fx["buffers_after"][k].abs().max())), k
