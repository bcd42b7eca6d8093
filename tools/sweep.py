"""BASELINE configs[4]: kernel shape sweep.  One un-graphed training step of the drop-in DRFNet per
configuration (channels F, patch = d slices of LR s x s folded into the batch, T frames, x4) with a CUDA-event
pair around every tensor-core launch; every distinct kernel shape is reported against its own bound
(tensor: sustained bf16 peak, hbm: measured copy bandwidth), plus the bandwidth kernels of tools/kbench.py.

    python tools/sweep.py --json profiles/r01_sweep.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import peaks  # noqa: E402
from vsr_b200.metrics import PSNR, SSIM  # noqa: E402
from vsr_b200.nets import DRFNet  # noqa: E402
from vsr_b200.ops import cuda_ops  # noqa: E402
from vsr_b200.optim import FlatAdam  # noqa: E402
from vsr_b200.runner import VSRTrainStep  # noqa: E402

CONFIGS = [  # F, d (slices = batch), s (LR side), T
    (64, 32, 32, 5), (64, 48, 48, 5), (64, 64, 64, 5), (64, 32, 32, 7),
    (128, 32, 32, 5), (128, 48, 48, 5), (128, 32, 32, 7),
]


def one(F, d, s, T, pk):
    dev = torch.device("cuda")
    torch.manual_seed(0)
    net = DRFNet(1, 1, F, 6, 4, precision="bf16").to(dev)
    opt = FlatAdam(net.parameters(), lr=1e-4)
    step = VSRTrainStep(net, [torch.nn.L1Loss()], [1.0], [PSNR().to(dev), SSIM().to(dev)], opt, "acdc", use_graph=False)
    ops = cuda_ops()
    x = [torch.randn(d, 1, s, s, device=dev) for _ in range(T)]
    y = [torch.randn(d, 1, 4 * s, 4 * s, device=dev) for _ in range(T)]
    acc = torch.zeros(4, device=dev)
    step.train_step(x, y, acc)
    torch.cuda.synchronize()
    ops.start_timing()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step.train_step(x, y, acc)
    b.record()
    torch.cuda.synchronize()
    timing = ops.gemm_records(ops.stop_timing())
    detail = {}
    for kind, flops, e0, e1, sig, nbytes in timing:
        q = detail.setdefault(f"{kind}:{sig}", [0.0, 0.0, 0, 0.0])
        q[0] += flops
        q[1] += e0.elapsed_time(e1)
        q[2] += 1
        q[3] += nbytes
    rows = {}
    ridge = pk["bf16_tflops_sustained"] * 1e3 / pk["hbm_gbs"]
    for k, v in sorted(detail.items(), key=lambda kv: -kv[1][1]):
        tf, gb = v[0] / (v[1] * 1e-3) / 1e12, v[3] / (v[1] * 1e-3) / 1e9
        tensor = v[0] / max(v[3], 1.0) > ridge
        rows[k] = {"us_per_launch": v[1] / v[2] * 1e3, "launches": v[2], "tflops": tf, "gbs": gb,
                   "bound": "tensor" if tensor else "hbm",
                   "frac": tf / pk["bf16_tflops_sustained"] if tensor else gb / pk["hbm_gbs"]}
    ms = a.elapsed_time(b)
    del step, net, opt, x, y
    torch.cuda.empty_cache()
    return {"config": {"F": F, "slices": d, "lr": s, "frames": T, "upscale": 4},
            "step_ms_ungraphed": ms, "hr_voxels_per_s_ungraphed": d * T * (4 * s) ** 2 / ms * 1e3, "shapes": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    pk = peaks()
    out = []
    for F, d, s, T in CONFIGS:
        r = one(F, d, s, T, pk)
        out.append(r)
        top = list(r["shapes"].items())[:6]
        print(f"F={F} slices={d} LR={s}x{s} T={T}: {r['step_ms_ungraphed']:.1f} ms/step (un-graphed)", flush=True)
        for k, v in top:
            print(f"    {k:46s} {v['us_per_launch']:8.1f} us  {v['bound']:6s} {v['frac']:.2f}", flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
