"""The bandwidth-bound kernels of BASELINE.json north_star (b) alone, on maps far larger than the 126 MB L2 (about 1 GB of
traffic per launch) and at the BASELINE config-2 sizes: achieved GB/s of ALGORITHMIC bytes against the
measured copy bandwidth (MEASURED_PEAKS.json), CUDA events on the launching stream, SM clocks sampled.

    python tools/bw_bench.py [--iters 10] [--json gpurun_out/bw.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler, peaks  # noqa: E402
from vsr_b200.drf_plan import phase_table  # noqa: E402
from vsr_b200.ops import cuda_ops  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None, help="run only the rows whose name contains this substring")
    args = ap.parse_args()
    ops, pk = cuda_ops(), peaks()
    dev = "cuda"
    rows = {}

    def rec(name, size, nbytes, fn, flops=None):
        if args.only and args.only not in name:
            return
        ms = timed(fn, args.iters)
        r = {"size": size, "us": ms * 1e3, "algorithmic_mb": nbytes / 1e6, "gbs": nbytes / ms / 1e6,
             "frac_hbm": nbytes / ms / 1e6 / pk["hbm_gbs"]}
        if flops:
            r["gflops"] = flops / ms / 1e6
        rows[f"{name} [{size}]"] = r
        print(f"{name:34s} {size:34s} {ms * 1e3:9.1f} us {r['gbs']:8.0f} GB/s  {r['frac_hbm']:.2f} of HBM" +
              (f"  {r['gflops'] / 1e3:.1f} TFLOP/s fp32" if flops else ""), flush=True)

    sampler = ClockSampler(0)
    sampler.start()
    for label, (n, hh, ww) in (("large", (96, 1024, 1024)), ("config-2", (160, 128, 128))):
        numel = n * hh * ww
        a = torch.randn(n, 1, hh, ww, device=dev)
        b = torch.randn(n, 1, hh, ww, device=dev)
        g = torch.empty_like(a)
        part = torch.zeros(max(1, n // 32), ops.partials_len, device=dev)
        segs = max(1, n // 32)
        size = f"{label}: {n}x1x{hh}x{ww} fp32"
        for kind, nm in ((0, "loss L1 fwd+bwd"), (1, "loss MSE fwd+bwd")):
            rec(nm, size, 12 * numel, lambda kind=kind: ops.loss_fwd_bwd_seg(a, b, segs, kind, 0.0, 1.0 / numel, part, g))
        ws = torch.empty(ops.metric_workspace(n, hh * ww) // 4 + 4, device=dev)
        out = torch.empty(n, device=dev)
        rec("denormalize + PSNR", size, 8 * numel, lambda: ops.psnr(a, b, 54.089, 48.084, 255.0, out, ws))
        win = torch.ones(11, device=dev) / 11
        rec("denormalize + SSIM (11x11, 5 maps)", size, 8 * numel,
            lambda: ops.ssim(a.view(n, hh, ww), b.view(n, hh, ww), win, 54.089, 48.084, 6.5, 58.5, out, ws),
            flops=2.0 * 110 * numel)
        del a, b, g
    # up-sampling (fp32 NCHW, the global skip of srfb_net.py:47 and its trilinear sibling)
    for label, shp, size in (("large", (64, 1, 512, 512), (2048, 2048)), ("config-2", (160, 1, 32, 32), (128, 128))):
        x = torch.randn(*shp, device=dev)
        y = torch.empty(shp[0], 1, *size, device=dev)
        dx = torch.empty_like(x)
        nb = 4 * (x.numel() + y.numel())
        s = f"{label}: {shp[0]}x1x{shp[2]}x{shp[3]} -> x4"
        rec("bilinear x4 fwd", s, nb, lambda: ops.upsample_linear(x, y, False))
        rec("bilinear x4 bwd", s, nb, lambda: ops.upsample_linear_bwd(y, dx, False))
        del x, y, dx
    x = torch.randn(8, 1, 64, 256, 256, device=dev)
    y = torch.empty(8, 1, 128, 512, 512, device=dev)
    dx = torch.empty_like(x)
    nb = 4 * (x.numel() + y.numel())
    rec("trilinear x2 fwd", "large: 8x1x64x256x256 -> x2", nb, lambda: ops.upsample_linear(x, y, False))
    rec("trilinear x2 bwd", "large: 8x1x64x256x256 -> x2", nb, lambda: ops.upsample_linear_bwd(y, dx, False))
    del x, y, dx
    x = torch.randn(64, 16, 512, 512, device=dev)
    y = torch.empty(64, 4, 1024, 1024, device=dev)
    rec("pixel shuffle x2", "large: 64x16x512x512", 8 * x.numel(), lambda: ops.pixel_shuffle(x, y, 2))
    del x, y
    # first / last convolutions of the net (config 2: 160 frames x samples of LR 32x32, F = 64, x4) and a large case
    for label, n in (("large", 2560), ("config-2", 160)):
        h = w = 32
        F = 64
        xin = torch.randn(n, 1, h, w, device=dev)
        w1 = torch.randn(4 * F, 1, 3, 3, device=dev) * 0.1
        b1 = torch.zeros(4 * F, device=dev)
        slope = torch.tensor([0.2], device=dev)
        a1 = torch.empty(n, h, w, 4 * F, device=dev, dtype=torch.bfloat16)
        px = n * h * w
        s = f"{label}: {n}x1x{h}x{w} -> 256 ch bf16"
        rec("first conv 3x3 (K = 9) + PReLU", s, px * (4 + 2 * 4 * F), lambda: ops.conv3x3_first(xin, w1, b1, slope, a1))
        dw1, db1 = torch.zeros_like(w1), torch.zeros_like(b1)
        wsf = torch.empty(ops.conv3x3_first_bwd_workspace(xin, 4 * F) // 4 + 4, device=dev)
        rec("first conv weight gradient", s, px * (4 + 2 * 4 * F), lambda: ops.conv3x3_first_bwd(xin, a1, dw1, db1, False, wsf))
        del a1
        r = 4
        ph = phase_table(r)
        xs = torch.randn(n, h, w, r * r * F, device=dev).to(torch.bfloat16)
        wl = torch.randn(1, F, 3, 3, device=dev) * 0.1
        bl = torch.zeros(1, device=dev)
        yl = torch.empty(n, 1, h * r, w * r, device=dev)
        hp = n * h * w * r * r
        s = f"{label}: {n}x{h * r}x{w * r}x64 bf16 -> 1 ch"
        rec("last conv 3x3 (N = 1) fwd", s, hp * (2 * F + 4), lambda: ops.conv3x3_last(xs, r, F, ph, wl, bl, yl))
        dxs = torch.empty_like(xs)
        dwl, dbl = torch.zeros_like(wl), torch.zeros_like(bl)
        wsl = torch.empty(ops.conv3x3_last_bwd_workspace(xs, r, F, 1) // 4 + 4, device=dev)
        rec("last conv 3x3 dx + dw", s, hp * (2 * 2 * F + 2 * F + 2 * 4),
            lambda: ops.conv3x3_last_bwd(xs, r, F, ph, wl, yl, dxs, dwl, dbl, False, wsl))
        del xs, dxs, yl
    # flat-bucket kernels
    P = 64 * 1024 * 1024
    p, gr, m, v = (torch.randn(P, device=dev) for _ in range(4))
    v.abs_()
    hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-8, 0.0, 1.0, 1.0], device=dev)
    rec("Adam on the flat bucket", f"large: {P} parameters", 28 * P, lambda: ops.adam_flat_dev(p, gr, m, v, hyper))
    a16, b16 = torch.randn(P, device=dev).to(torch.bfloat16), torch.randn(P, device=dev).to(torch.bfloat16)
    o16 = torch.empty_like(a16)
    rec("add (bf16)", f"large: {P} elements", 6 * P, lambda: ops.add(a16, b16, o16))
    big16 = torch.randn(8 * P, device=dev, dtype=torch.bfloat16)      # 1 GB
    for c in (64, 256):                                  # bias gradients: column sums of a [rows][c] bf16 map
        nrow = 8 * P // c
        dbias = torch.zeros(c, device=dev)
        wsc = torch.empty(ops.colsum_workspace(nrow, c) // 4 + 4, device=dev)
        rec("column sums (bias gradient, bf16)", f"large: {nrow} x {c}", 2 * 8 * P,
            lambda nrow=nrow, c=c, dbias=dbias, wsc=wsc: ops.colsum(big16, nrow, c, dbias, False, wsc))
    del big16
    part = torch.zeros(ops.partials_len, device=dev)
    rec("PReLU backward (act_bwd, bf16)", f"large: {P} elements", 6 * P, lambda: ops.act_bwd(a16, b16, o16, slope, part))
    res = {"what": "bandwidth-bound kernels alone: us per launch (median), algorithmic bytes / time, fraction of the measured copy "
                   f"bandwidth ({pk['hbm_gbs']} GB/s, {pk['source']}); SSIM also as fp32 TFLOP/s (110 FMA per pixel: it is bound by "
                   "the FP32 pipes, 72 TFLOP/s nominal, not by HBM)",
           "clocks": sampler.finish(), "rows": rows}
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
