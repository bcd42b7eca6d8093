"""HBM-bound 1x1 convolutions on concatenations (config-2 HR shapes): time vs taps, epilogue kind, env knobs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kbench  # noqa: E402
from vsr_b200.ops import TapTable, cuda_ops  # noqa: E402


def case(taps, epi, n_px_w, iters, flush):
    ops = cuda_ops()
    dt = torch.bfloat16
    N, h, F = 32, 32, 64
    w = n_px_w
    tab = TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(taps)])])
    srcs = [torch.randn(N, h, w, F, device="cuda").to(dt) for _ in range(taps)]
    out = torch.empty(N, h, w, F, device="cuda", dtype=dt)
    wts = (torch.randn(taps * 64 * 64, device="cuda") * 0.05).to(dt)
    bias = torch.zeros(F, device="cuda")
    slope = torch.tensor([0.2], device="cuda")
    aux = torch.randn(N, h, w, F, device="cuda").to(dt)
    res = torch.randn(N, h, w, F, device="cuda").to(dt)
    part = torch.zeros(ops.partials_len, device="cuda")
    kw = {}
    n_maps = taps + 1
    if epi & 16:
        kw.update(aux_y=aux, slope_partials=part)
        n_maps += 1
    if epi & 2:
        kw.update(residual=res)
        n_maps += 1
    fn = lambda: ops.tapgemm(tab, srcs, out, wts, bias=bias if epi & 1 else None, epi=epi, slope=slope, **kw)
    ms = kbench.timed(fn, iters, flush)
    nbytes = n_maps * out.numel() * 2
    return ms, nbytes / ms / 1e6


def main():
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    for taps in (1, 2, 3, 6):
        for epi in (5, 16, 18):
            ms, gbs = case(taps, epi, 512, 10, flush)
            print(f"hr1x1 taps={taps} epi={epi:2d}: {ms * 1e3:7.1f} us  {gbs:6.0f} GB/s ({gbs / 65.533:.0f}% of measured HBM)", flush=True)


if __name__ == "__main__":
    main()
