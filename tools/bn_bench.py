"""HBM roofline of the bandwidth-bound kernels of the Conv3d path (vsr_b200/csrc/dense3d.cu), each alone on maps
far larger than the 126 MB L2: algorithmic bytes / CUDA-event time against MEASURED_PEAKS.json's copy bandwidth.

    python tools/bn_bench.py [--out file.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vsr_b200.ops import cuda_ops  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    peak = 6553.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    ops, dev = cuda_ops(), "cuda"
    res = []
    for dtype, es in ((torch.bfloat16, 2), (torch.float32, 4)):
        frames, rows_pf, ld = 7, 1 << 18, 288                       # 1.8 M rows x 288 channels (1.06 / 2.1 GB)
        rows = frames * rows_pf
        x = torch.randn(rows, 1, 1, ld, device=dev).to(dtype)
        ws = torch.empty(1 << 22, dtype=torch.float64, device=dev)
        for c in (64, 128, 256):
            cp = c
            stats = torch.zeros(frames, 2, ld, dtype=torch.float64, device=dev)
            ss, mr = torch.empty(2, cp, device=dev), torch.empty(2, c, device=dev)
            gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            y = torch.empty(rows, 1, 1, cp, dtype=dtype, device=dev)
            dy = torch.randn(rows, 1, 1, cp, device=dev).to(dtype)
            dx = torch.zeros(rows, 1, 1, ld, dtype=dtype, device=dev)
            gb = torch.zeros(2 * c, device=dev)
            ms = timed(lambda: ops.bn_stats(x, 0, c, frames, stats, 0, ws))
            res.append(("bn_stats", str(dtype), c, ms, es * rows * c))
            ops.bn_finalize(stats, 0, frames, rows_pf, c, gamma, beta, 1e-5, 0.1, None, None, True, ss, mr)
            ms = timed(lambda: ops.bn_relu(x, 0, c, ss, y))
            res.append(("bn_relu", str(dtype), c, ms, es * rows * (c + cp)))
            ms = timed(lambda: ops.bn_relu_bwd(dy, x, 0, c, ss, mr, gb, y, 0, cp, False, ws))
            res.append(("bn_relu_bwd(write)", str(dtype), c, ms, es * rows * (2 * c + 2 * c + cp)))   # two passes read dy, x
            ms = timed(lambda: ops.bn_relu_bwd(dy, x, 0, c, ss, mr, gb, dx, 0, c, True, ws))
            res.append(("bn_relu_bwd(accumulate)", str(dtype), c, ms, es * rows * (2 * c + 2 * c + 2 * c)))
            ms = timed(lambda: ops.copy_window(x, 0, y, 0, c))
            res.append(("copy_window", str(dtype), c, ms, es * rows * 2 * c))
            del y, dy, dx
    out = [{"kernel": k, "dtype": d, "channels": c, "ms": ms, "algorithmic_bytes": b, "gbs": b / ms / 1e6,
            "frac_of_hbm_peak": b / ms / 1e6 / peak} for k, d, c, ms, b in res]
    for o in out:
        print(f"{o['kernel']:26s} {o['dtype']:15s} c={o['channels']:3d} {o['ms']:8.3f} ms {o['gbs']:7.0f} GB/s  {o['frac_of_hbm_peak']:.2f}")
    if a.out:
        json.dump({"hbm_peak_gbs": peak, "rows": 7 << 18, "ld": 288, "results": out}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
