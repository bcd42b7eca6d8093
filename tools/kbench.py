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
    if args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
