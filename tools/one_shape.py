"""Run ONE tap-GEMM shape a few times (for `ncu --set full --import-source on -k regex:tapgemm_tc2 -s 4 -c 1`).

    python tools/one_shape.py hr2 | hr6 | deconv | conv8 | deconv_bwd | dgrad16 | duf_c2z | duf_c2 | duf_dgrad

`dgrad16`: the data gradient of the strided 8x8 convolution with the PReLU' epilogue (16 taps, nt 256, 4 groups, epi 16:
saved activation by TMA, slope-gradient partial sums) - the shape whose epilogue was instruction-bound.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hr_sweep  # noqa: E402
import kbench  # noqa: E402
from vsr_b200.ops import TapTable  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "hr2"
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    if which.startswith("hr"):
        ms, gbs = hr_sweep.case(int(which[2:]), 5, 512, 3, flush)
        print(which, ms * 1e3, "us", gbs, "GB/s")
        return
    res = []
    kbench.CASES = None
    if which.startswith("duf"):
        # growth convolution of DUFNet-16, second dense layer (96 -> 32 channels, stored as 128), batch 32 x 7 frames
        n, h, w = 7 * 32, 32, 32
        kbench.EPI = 0
        sp = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
        if which == "duf_c2z":      # temporal taps as output columns: one source, 9 x 2 taps, N = 96 of 128
            tab = TapTable(64, 128, [(0, [(0, dy, dx, b * 64) for dy, dx in sp for b in range(2)])])
            kbench.tapgemm_case("duf_c2z", tab, n, h, w, [128], 128, torch.bfloat16, 3, flush, res)
        elif which == "duf_c2":     # direct form: three frame-shifted sources, 27 x 2 taps, N = 32 of 64
            kbench.EPI = 1
            tab = TapTable(64, 64, [(96, [(kt, dy, dx, b * 64) for kt in range(3) for dy, dx in sp for b in range(2)])])
            kbench.tapgemm_case("duf_c2", tab, n, h, w, [128, 128, 128], 288, torch.bfloat16, 3, flush, res)
        else:                       # its data gradient: 27 taps of the 64-wide gradient window, N = 128
            tab = TapTable(64, 128, [(0, [(kt, -dy, -dx, 96) for kt in range(3) for dy, dx in sp])])
            kbench.tapgemm_case("duf_dgrad", tab, n, h, w, [288, 288, 288], 128, torch.bfloat16, 3, flush, res)
        return
    N, h, w, F = 32, 32, 32, 64
    if which == "dgrad16":
        ops = kbench.cuda_ops()
        groups = []
        for g in range(4):
            gy, gx = g // 2, g % 2
            groups.append((g * 256, [(0, dy - 1 + gy, dx - 1 + gx, 0) for dy in (0, 1) for dx in (0, 1)]))
        tab = TapTable(64, 256, groups)
        dz = torch.randn(N, h, w, F, device="cuda").to(torch.bfloat16)
        y = torch.randn(N, h, w, 16 * F, device="cuda").to(torch.bfloat16)        # saved activation of the consumer
        out = torch.empty_like(y)
        wts = (torch.randn(tab.n_taps_total * tab.nt * tab.kc, device="cuda") * 0.05).to(torch.bfloat16)
        slope = torch.tensor([0.2], device="cuda")
        parts = torch.zeros(1024, device="cuda")
        fn = lambda: ops.tapgemm(tab, [dz], out, wts, epi=16, slope=slope, aux_y=y, slope_partials=parts)
        ms = kbench.timed(fn, 5, flush)
        print("dgrad16", ms * 1e3, "us")
        return
    if which in ("deconv", "deconv_bwd"):
        groups = []
        for g in range(4):
            gy, gx = g // 2, g % 2
            groups.append((g * 256, [(0, dy - 1 + gy, dx - 1 + gx, 0) for dy in (0, 1) for dx in (0, 1)]))
        kbench.EPI = 5
        kbench.tapgemm_case("deconv8x8s4", TapTable(64, 256, groups), N, h, w, [F], 16 * F, torch.bfloat16, 3, flush, res)
    else:
        taps = []
        for ky in range(8):
            for kx in range(8):
                dy, py = divmod(ky - 2, 4)
                dx, px = divmod(kx - 2, 4)
                taps.append((0, dy, dx, (py * 4 + px) * 64))
        kbench.tapgemm_case("conv8x8s4", TapTable(64, 64, [(0, taps)]), N, h, w, [16 * F], F, torch.bfloat16, 3, flush, res)


if __name__ == "__main__":
    main()
