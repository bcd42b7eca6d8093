"""Inference throughput at BASELINE configs[2]: full-FOV DSB15-shaped cine (256x256 x 12 slices x 30 frames, x4),
DRFNet-L bf16, all frames in one no-grad call, PSNR / SSIM of every frame on the device.

    python tools/infer_bench.py [--iters 10] [--json out.json] [--net drf|duf]

--net duf: the same cine through DUFNet-16 (Conv3d path): every output frame is one 7-frame window with temporal
wrap-around (acdc_misr_dataset.py:55-66), i.e. 12 x 30 = 360 windows, run in batches of 60 (eval mode).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsr_b200.metrics import PSNR, SSIM  # noqa: E402
from vsr_b200.nets import DRFNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default=None)
    ap.add_argument("--net", default="drf", choices=["drf", "duf"])
    args = ap.parse_args()
    if args.net == "duf":
        return main_duf(args)
    n, t, h, r = 12, 30, 64, 4
    torch.manual_seed(0)
    net = DRFNet(1, 1, 64, 6, r, precision="bf16").to("cuda").eval()
    x = [torch.randn(n, 1, h, h, device="cuda") for _ in range(t)]
    y = [torch.randn(n, 1, h * r, h * r, device="cuda") for _ in range(t)]
    psnr, ssim = PSNR(dataset="dsb15").cuda(), SSIM(dataset="dsb15").cuda()

    def run(with_metrics):
        with torch.no_grad():
            outs = net(x)
            if with_metrics:
                return torch.stack([psnr(o, q) for o, q in zip(outs, y)]), torch.stack([ssim(o, q) for o, q in zip(outs, y)])
        return outs

    res = {}
    for name, wm in (("forward", False), ("forward+psnr+ssim", True)):
        for _ in range(3):
            run(wm)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            run(wm)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        vox = n * t * (h * r) ** 2
        res[name] = {"ms": ms, "hr_voxels_per_s": vox / ms * 1e3, "tflops_algorithmic": 10.673e6 * n * t * h * h / ms / 1e9}
        print(name, json.dumps(res[name]), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"workload": "C3: DRFNet-L x4 inference, 12 slices x 30 frames, LR 64x64 -> HR 256x256, bf16", **res}, f, indent=1)


def main_duf(args):
    from vsr_b200.duf import DUFNet
    n, t, h, r, nf, bs = 12, 30, 64, 4, 7, 60
    torch.manual_seed(0)
    net = DUFNet(1, 1, nf, 5, r, "_DenseLayer16", precision="bf16").to("cuda").eval()
    cine = torch.randn(t, n, 1, h, h, device="cuda")                       # [frame][slice]
    target = torch.randn(t * n, 1, h * r, h * r, device="cuda")
    psnr, ssim = PSNR(dataset="dsb15").cuda(), SSIM(dataset="dsb15").cuda()
    # window of output frame f: frames f-3 .. f+3 (mod t); the windows of `bs // n` consecutive output frames per call
    fpc = bs // n

    def run(with_metrics):
        vals = []
        with torch.no_grad():
            for f0 in range(0, t, fpc):
                frames = [torch.cat([cine[(f + k - nf // 2) % t] for f in range(f0, f0 + fpc)]) for k in range(nf)]
                out = net(frames)
                if with_metrics:
                    tg = target[f0 * n:(f0 + fpc) * n]
                    vals.append((psnr(out, tg), ssim(out, tg)))
        return vals

    res = {}
    for name, wm in (("forward", False), ("forward+psnr+ssim", True)):
        for _ in range(3):
            run(wm)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            run(wm)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        res[name] = {"ms": ms, "hr_voxels_per_s": n * t * (h * r) ** 2 / ms * 1e3}
        print(name, json.dumps(res[name]), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"workload": "C3 through DUFNet-16 x4 (Conv3d path): 360 windows of 7 frames, LR 64x64 -> HR 256x256, "
                                   "batches of 60 windows, bf16, eval mode", **res}, f, indent=1)


if __name__ == "__main__":
    main()
