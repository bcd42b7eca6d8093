"""DUFNet (Conv3d path) training-step timing on one GPU: CUDA events around the whole
forward + L1 + backward, then one extra pass with events around every tap-GEMM / weight-gradient launch.

    python tools/duf_bench.py [--n 4] [--hw 32] [--r 4] [--precision bf16] [--steps 10] [--out file.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsr_b200.duf import DUFNet  # noqa: E402
from vsr_b200.ops import cuda_ops  # noqa: E402


def flops_fwd(net, n, h, w):
    """algorithmic forward FLOPs (2 * out pixels * Cout * Cin * taps; padding taps counted), duf_net.py layers"""
    P = net._plan
    px, total = n * h * w, 0.0
    total += 2.0 * px * P.T * 64 * P.cin * 9
    for i in range(P.L):
        _, tin, _, tout = P.frames_of(i)
        total += 2.0 * px * tin * P.C[i] * P.C[i] + 2.0 * px * tout * P.Gr * P.C[i] * 27
    total += 2.0 * px * (256 * P.ctot * 9 + 256 * 768 + 512 * P.cf + 256 * P.cr)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--hw", type=int, default=32)
    ap.add_argument("--r", type=int, default=4)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.manual_seed(0)
    dev = "cuda"
    net = DUFNet(1, 1, 7, 5, a.r, "_DenseLayer16", precision=a.precision).to(dev).train()
    g = torch.Generator(device="cpu").manual_seed(0)
    frames = [torch.randn(a.n, 1, a.hw, a.hw, generator=g).to(dev) for _ in range(7)]
    target = torch.randn(a.n, 1, a.hw * a.r, a.hw * a.r, generator=g).to(dev)
    lossf = torch.nn.L1Loss()
    ops = cuda_ops()

    def step():
        out = net(frames)
        loss = lossf(out, target)
        loss.backward()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    l0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (ops.launches - l0) // a.steps
    # forward only
    with torch.no_grad():
        for _ in range(2):
            net(frames)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            net(frames)
        e1.record()
        torch.cuda.synchronize()
    ms_fwd = e0.elapsed_time(e1) / a.steps
    # the full training step as the MISR trainer runs it (fused L1, FlatAdam, PSNR + SSIM), CUDA-graphed
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import MISRTrainStep
    opt = FlatAdam(net.parameters(), lr=1e-4)
    ts = MISRTrainStep(net, [torch.nn.L1Loss()], [1.0], [PSNR().to(dev), SSIM().to(dev)], opt, "acdc", use_graph=True)
    acc = torch.zeros(4, device=dev)
    for _ in range(max(a.warmup, 4)):
        ts.train_step(frames, [target], acc)
    torch.cuda.synchronize()
    l0 = ops.launches
    e0.record()
    for _ in range(a.steps):
        ts.train_step(frames, [target], acc)
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / a.steps
    ops.start_timing()
    step()
    torch.cuda.synchronize()
    detail = {}
    for kind, fl, s, e, sig, nbytes in ops.gemm_records(ops.stop_timing()):
        d = detail.setdefault(f"{kind}:{sig}", {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["n"] += 1; d["ms"] += s.elapsed_time(e); d["flops"] += fl; d["bytes"] += nbytes
    for d in detail.values():
        d["tflops_padded"] = d["flops"] / d["ms"] / 1e9
        d["gbs"] = d["bytes"] / d["ms"] / 1e6
    gemm_ms = sum(d["ms"] for d in detail.values())
    f = flops_fwd(net, a.n, a.hw, a.hw)
    hr_vox = a.n * (a.hw * a.r) ** 2
    res = {"net": "DUFNet _DenseLayer16", "precision": a.precision, "n": a.n, "lr": a.hw, "r": a.r, "frames": 7,
           "ms_per_step_graphed_full": ms_graph, "hr_voxels_per_s_train_graphed": hr_vox / ms_graph * 1e3,
           "algorithmic_tflops_step_graphed": 3 * f / ms_graph / 1e9,
           "ms_per_step": ms, "ms_forward": ms_fwd, "launches_per_step": launches,
           "hr_voxels_per_s_train": hr_vox / ms * 1e3, "hr_voxels_per_s_infer": hr_vox / ms_fwd * 1e3,
           "algorithmic_tflops_step": 3 * f / ms / 1e9, "algorithmic_tflops_forward": f / ms_fwd / 1e9,
           "gemm_ms_events": gemm_ms,
           "detail": dict(sorted(detail.items(), key=lambda kv: -kv[1]["ms"]))}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
