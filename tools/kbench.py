"""Per-kernel timings at the BASELINE config-2 shapes (DRFNet-L x4, N=32, LR 32x32) and the sweep
of BASELINE config 5.  CUDA events on the launching stream, L2 flushed between iterations.

    python tools/kbench.py [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from vsr_b200.ops import TapTable, cuda_ops  # noqa: E402

PEAKS = {"hbm_gbs": 6553.3, "bf16_tflops": 1648.6}
try:
    with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as f:
        PEAKS.update(json.load(f))
except OSError:
    pass


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)  # > L2 (126 MB): evicts the working set
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


CASES = None
EPI = 1 | 4


def tapgemm_case(name, tab, n, h, w, src_cs, out_c, dtype, iters, flush, results):
    if CASES and name not in CASES:
        return
    ops = cuda_ops()
    srcs = [torch.randn(n, h, w, c, device="cuda").to(dtype) for c in src_cs]
    out = torch.empty(n, h, w, out_c, device="cuda", dtype=dtype)
    wts = (torch.randn(tab.n_taps_total * tab.nt * tab.kc, device="cuda") * 0.05).to(dtype)
    bias = torch.zeros(out_c, device="cuda")
    slope = torch.tensor([0.2], device="cuda")
    fn = lambda: ops.tapgemm(tab, srcs, out, wts, bias=bias, epi=EPI, slope=slope)
    ms = timed(fn, iters, flush)
    pix = n * h * w
    flops = 2.0 * pix * tab.n_taps_total * tab.nt * tab.kc
    es = out.element_size()
    bytes_ = es * (sum(pix * c for c in src_cs) + pix * out_c) + wts.numel() * es
    r = {"kernel": name, "dtype": str(dtype).split(".")[-1], "ms": ms, "tflops": flops / ms / 1e9,
         "gbs": bytes_ / ms / 1e6, "flops": flops, "bytes": bytes_,
         "frac_tensor": flops / ms / 1e9 / PEAKS["bf16_tflops"], "frac_hbm": bytes_ / ms / 1e6 / PEAKS["hbm_gbs"]}
    results.append(r)
    print(json.dumps(r), flush=True)


def bw_case(name, fn, nbytes, iters, flush, results):
    ms = timed(fn, iters, flush)
    r = {"kernel": name, "ms": ms, "gbs": nbytes / ms / 1e6, "bytes": nbytes, "frac_hbm": nbytes / ms / 1e6 / PEAKS["hbm_gbs"]}
    results.append(r)
    print(json.dumps(r), flush=True)


def bandwidth_cases(iters, flush, results):
    """HBM-bound kernels at BASELINE config-2/3 sizes; bytes = algorithmic bytes."""
    from vsr_b200.drf_plan import phase_table
    ops = cuda_ops()
    dev = "cuda"
    n, H, W = 32 * 5, 128, 128                      # all T=5 frames of a config-2 batch in one call
    a, b = torch.randn(n, 1, H, W, device=dev), torch.randn(n, 1, H, W, device=dev)
    numel = a.numel()
    part = torch.zeros(ops.partials_len, device=dev)
    grad = torch.empty_like(a)
    for kind, nm in ((0, "l1"), (1, "mse"), (2, "charbonnier"), (3, "huber")):
        bw_case(f"loss_{nm}_fwd_bwd", lambda k=kind: ops.loss_fwd_bwd(a, b, k, 0.5, 1.0 / numel, part, grad), 12 * numel, iters, flush, results)
    ws = torch.empty(ops.metric_workspace(n, H * W) // 4 + 4, device=dev)
    out = torch.empty(n, device=dev)
    bw_case("psnr_denorm", lambda: ops.psnr(a, b, 54.089, 48.084, 255.0, out, ws), 8 * numel, iters, flush, results)
    win = torch.ones(11, device=dev) / 11
    bw_case("ssim_denorm", lambda: ops.ssim(a.view(n, H, W), b.view(n, H, W), win, 54.089, 48.084, 6.5, 58.5, out, ws), 8 * numel, iters, flush, results)
    # DSB15-shaped frame batch (config 3): 12 x 256 x 256
    a3, b3 = torch.randn(12 * 30, 1, 256, 256, device=dev), torch.randn(12 * 30, 1, 256, 256, device=dev)
    ws3 = torch.empty(ops.metric_workspace(360, 65536) // 4 + 4, device=dev)
    out3 = torch.empty(360, device=dev)
    bw_case("psnr_denorm_dsb15", lambda: ops.psnr(a3, b3, 51.193, 52.671, 255.0, out3, ws3), 8 * a3.numel(), iters, flush, results)
    bw_case("ssim_denorm_dsb15", lambda: ops.ssim(a3.view(360, 256, 256), b3.view(360, 256, 256), win, 51.193, 52.671, 6.5, 58.5, out3, ws3), 8 * a3.numel(), iters, flush, results)
    # pixel shuffle / upsample (standalone kernels of the sweep)
    x = torch.randn(32, 256, 64, 64, device=dev)
    y = torch.empty(32, 64, 128, 128, device=dev)
    bw_case("pixel_shuffle_r2", lambda: ops.pixel_shuffle(x, y, 2), 8 * x.numel(), iters, flush, results)
    bw_case("pixel_unshuffle_r2", lambda: ops.pixel_shuffle(y, x, 2, inverse=True), 8 * x.numel(), iters, flush, results)
    xi = torch.randn(32, 64, 64, 64, device=dev)
    yo = torch.empty(32, 64, 256, 256, device=dev)
    bw_case("bilinear_x4", lambda: ops.upsample_linear(xi, yo, False), 4 * (xi.numel() + yo.numel()), iters, flush, results)
    bw_case("bilinear_x4_bwd", lambda: ops.upsample_linear_bwd(yo, xi, False), 4 * (xi.numel() + yo.numel()), iters, flush, results)
    x3 = torch.randn(4, 32, 16, 64, 64, device=dev)
    y3 = torch.empty(4, 32, 32, 128, 128, device=dev)
    bw_case("trilinear_x2", lambda: ops.upsample_linear(x3, y3, False), 4 * (x3.numel() + y3.numel()), iters, flush, results)
    bw_case("trilinear_x2_bwd", lambda: ops.upsample_linear_bwd(y3, x3, False), 4 * (x3.numel() + y3.numel()), iters, flush, results)
    # Adam over a DRFNet-L sized and a large bucket
    for P in (3658907, 64 * 1024 * 1024):
        p_, g_, m_, v_ = (torch.randn(P, device=dev) for _ in range(4))
        v_.abs_()
        bw_case(f"adam_flat_{P}", lambda: ops.adam_flat(p_, g_, m_, v_, 1e-4, 0.9, 0.999, 1e-8, 0.0, 3), 28 * P, iters, flush, results)
    # boundary convolutions
    xin = torch.randn(32, 1, 32, 32, device=dev)
    w1, b1 = torch.randn(256, 1, 3, 3, device=dev), torch.randn(256, device=dev)
    sl = torch.tensor([0.2], device=dev)
    a1 = torch.empty(32, 32, 32, 256, device=dev, dtype=torch.bfloat16)
    bw_case("conv3x3_first_bf16", lambda: ops.conv3x3_first(xin, w1, b1, sl, a1), 2 * a1.numel() + 4 * xin.numel(), iters, flush, results)
    hr = torch.randn(32, 32, 32, 1024, device=dev).to(torch.bfloat16)
    wl, bl = torch.randn(1, 64, 3, 3, device=dev) * 0.1, torch.randn(1, device=dev)
    yl = torch.empty(32, 1, 128, 128, device=dev)
    ph = phase_table(4)
    bw_case("conv3x3_last_bf16", lambda: ops.conv3x3_last(hr, 4, 64, ph, wl, bl, yl), 2 * hr.numel() + 4 * yl.numel(), iters, flush, results)
    dyl, dxl = torch.randn_like(yl), torch.empty_like(hr)
    dwl, dbl = torch.zeros_like(wl), torch.zeros_like(bl)
    wsl = torch.empty(ops.conv3x3_last_bwd_workspace(hr, 4, 64, 1) // 4 + 4, device=dev)
    bw_case("conv3x3_last_bwd_bf16", lambda: ops.conv3x3_last_bwd(hr, 4, 64, ph, wl, dyl, dxl, dwl, dbl, False, wsl), 4 * hr.numel() + 4 * yl.numel(), iters, flush, results)
    z = torch.randn(524288, 64, device=dev).to(torch.bfloat16)
    dbz = torch.zeros(64, device=dev)
    wsz = torch.empty(ops.colsum_workspace(524288, 64) // 4 + 4, device=dev)
    bw_case("colsum_hr_bf16", lambda: ops.colsum(z, 524288, 64, dbz, False, wsz), 2 * z.numel(), iters, flush, results)
    z2 = torch.empty_like(z)
    pr = torch.zeros(ops.partials_len, device=dev)
    bw_case("prelu_bwd_hr_bf16", lambda: ops.act_bwd(z, z, z2, sl, pr), 6 * z.numel(), iters, flush, results)
    bw_case("add_hr_bf16", lambda: ops.add(z, z, z2), 6 * z.numel(), iters, flush, results)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--fp32", action="store_true", help="also time the CUDA-core fp32 kernels")
    ap.add_argument("--cases", default=None, help="comma-separated kernel names")
    ap.add_argument("--epi", type=int, default=5, help="epilogue flags (default bias+PReLU)")
    args = ap.parse_args()
    global CASES, EPI
    CASES = set(args.cases.split(",")) if args.cases else None
    EPI = args.epi
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")  # 256 MB
    results = []
    N, h, w, F = 32, 32, 32, 64
    dts = [torch.bfloat16] + ([torch.float32] if args.fp32 else [])
    for dt in dts:
        # LR 1x1 on a 3-way concat
        tapgemm_case("conv1x1_lr_cat3", TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(3)])]),
                     N, h, w, [F] * 3, F, dt, args.iters, flush, results)
        # HR 1x1 on a 3-way concat: phase-blocked maps viewed as [N, h, 16w, 64]
        tapgemm_case("conv1x1_hr_cat3", TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(3)])]),
                     N, h, w * 16, [F] * 3, F, dt, args.iters, flush, results)
        tapgemm_case("conv1x1_hr_cat6", TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(6)])]),
                     N, h, w * 16, [F] * 6, F, dt, args.iters, flush, results)
        # deconv 8x8 s4 p2: 4 groups x 4 taps, N = 256
        groups = []
        for g in range(4):
            gy, gx = g // 2, g % 2
            groups.append((g * 256, [(0, dy - 1 + gy, dx - 1 + gx, 0) for dy in (0, 1) for dx in (0, 1)]))
        tapgemm_case("deconv8x8s4", TapTable(64, 256, groups), N, h, w, [F], 16 * F, dt, args.iters, flush, results)
        # strided conv 8x8 s4 p2: 64 taps on the phase-blocked HR map
        taps = []
        for ky in range(8):
            for kx in range(8):
                dy, py = divmod(ky - 2, 4)
                dx, px = divmod(kx - 2, 4)
                taps.append((0, dy, dx, (py * 4 + px) * 64))
        tapgemm_case("conv8x8s4", TapTable(64, 64, [(0, taps)]), N, h, w, [16 * F], F, dt, args.iters, flush, results)
        # 3x3 F -> 4F at LR and at 2x
        t33 = [(0, dy, dx, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        tapgemm_case("conv3x3_n256_lr", TapTable(64, 256, [(0, t33)]), N, h, w, [F], 4 * F, dt, args.iters, flush, results)
        tapgemm_case("conv3x3_n256_2x", TapTable(64, 256, [(0, t33)]), N, 2 * h, 2 * w, [F], 4 * F, dt, args.iters, flush, results)
    if not args.cases or "bw" in args.cases:
        bandwidth_cases(args.iters, flush, results)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
