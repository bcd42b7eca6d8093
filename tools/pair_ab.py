"""A/B timing of single-CTA vs CTA-pair (tcgen05 cta_group::2) tap-GEMM launches on the exact layer tables of the
BASELINE config-2 net (DRFNet-L x4, batch 32, LR 32x32): every distinct convolution shape of the step, forward and data
gradient, CUDA events on the launching stream, L2 flushed between iterations, SM clocks sampled.

    python tools/pair_ab.py [--iters 20] [--json out.json] [--batch 32]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ClockSampler, peaks  # noqa: E402
from vsr_b200._lib import EPI_BIAS, EPI_PRELU, EPI_PRELU_BWD  # noqa: E402
from vsr_b200.drf_plan import DrfPlan  # noqa: E402
from vsr_b200.ops import cuda_ops  # noqa: E402


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
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
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--lr", type=int, default=32)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    ops, pk = cuda_ops(), peaks()
    P = DrfPlan(1, 1, 64, 6, 4, bf16=True)
    n, h, w, F, r2 = args.batch, args.lr, args.lr, 64, 16
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    bf = torch.bfloat16
    lrm = lambda c=F: torch.randn(n, h, w, c, device="cuda").to(bf)
    hrv = lambda: torch.randn(n, h, w * r2, F, device="cuda").to(bf)        # an HR map seen as [n][h][w*16][64]
    slope = torch.tensor([0.2], device="cuda")
    part = torch.zeros(ops.partials_len, device="cuda")
    cases = [
        # (name, layer, store, sources, out shape, epilogue)
        ("deconv 8x8s4 fwd (up0_dc)", P.fwd["up2_dc"], [lrm()], (n, h, w, r2 * F), EPI_BIAS | EPI_PRELU),
        ("strided conv 8x8s4 fwd (dn_sc)", P.fwd["dn2_sc"], [lrm(r2 * F)], (n, h, w, F), EPI_BIAS | EPI_PRELU),
        ("dgrad of strided conv + PReLU' (dn_sc bwd)", P.bwd["dn2_sc"], [lrm()], (n, h, w, r2 * F), EPI_PRELU_BWD),
        ("dgrad of deconv + PReLU' (up_dc bwd)", P.bwd["up2_dc"], [lrm(r2 * F)], (n, h, w, F), EPI_PRELU_BWD),
        ("3x3 conv F->4F at 1x (out1)", P.fwd["out1"], [lrm()], (n, h, w, 4 * F), EPI_BIAS),
        ("3x3 conv F->4F at 2x (out2)", P.fwd["out2"], [lrm(4 * F)], (n, h, w, r2 * F), EPI_BIAS),
        ("dgrad of out2", P.bwd["out2"], [lrm(r2 * F)], (n, h, w, 4 * F), 0),
        ("HR 1x1 on 6-way concat (dn5_c1)", P.fwd["dn5_c1"], [hrv() for _ in range(6)], (n, h, w * r2, F), EPI_BIAS | EPI_PRELU),
        ("HR 1x1 on 2-way concat (dn1_c1)", P.fwd["dn1_c1"], [hrv() for _ in range(2)], (n, h, w * r2, F), EPI_BIAS | EPI_PRELU),
        ("LR 1x1 on 6-way concat (fout-like, up5_c1)", P.fwd["up5_c1"], [lrm() for _ in range(6)], (n, h, w, F), EPI_BIAS | EPI_PRELU),
    ]
    sampler = ClockSampler(0)
    sampler.start()
    rows = {}
    for name, L, srcs, oshape, epi in cases:
        tab = L.table
        out = torch.empty(*oshape, device="cuda", dtype=bf)
        wts = (torch.randn(len(L.slabs) * tab.nt * tab.kc, device="cuda") * 0.05).to(bf)
        bias = torch.zeros(oshape[-1], device="cuda")
        aux = torch.randn(*oshape, device="cuda").to(bf) if epi & EPI_PRELU_BWD else None
        fn = lambda: ops.tapgemm(tab, srcs, out, wts, bias=bias if epi & EPI_BIAS else None, epi=epi, slope=slope, aux_y=aux,
                                 slope_partials=part if epi & EPI_PRELU_BWD else None)
        row = {}
        for mode in ("0", "1"):
            os.environ["VSR_TC_PAIR"] = mode
            ops.lib.vsr_reload_tunables()
            row["pair" if mode == "1" else "single"] = timed(fn, args.iters, flush) * 1e3
        pix = oshape[0] * oshape[1] * oshape[2]
        flops = 2.0 * pix * tab.n_taps_total * tab.nt * tab.kc
        nbytes = 2 * (sum(s.numel() for s in srcs) + out.numel() + (aux.numel() if aux is not None else 0))
        best = min(row.values())
        row.update({"tflops_best": flops / best / 1e6, "gbs_best": nbytes / best / 1e3,
                    "frac_tensor_best": flops / best / 1e6 / pk["bf16_tflops_sustained"], "frac_hbm_best": nbytes / best / 1e3 / pk["hbm_gbs"],
                    "taps": tab.n_taps_total, "nt": tab.nt, "groups": tab.n_groups})
        rows[name] = row
        print(f"{name:48s} single {row['single']:7.1f} us   pair {row['pair']:7.1f} us   best: {row['tflops_best']:6.0f} TF/s "
              f"({row['frac_tensor_best']:.2f} of sustained bf16), {row['gbs_best']:5.0f} GB/s ({row['frac_hbm_best']:.2f} of HBM)", flush=True)
    del os.environ["VSR_TC_PAIR"]
    ops.lib.vsr_reload_tunables()
    res = {"workload": f"tap-GEMM shapes of DRFNet-L x4, batch {n}, LR {h}x{w}, bf16; us per launch, median of {args.iters}, L2 flushed",
           "clocks": sampler.finish(), "rows": rows}
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
