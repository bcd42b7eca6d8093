"""Run ONE tap-GEMM shape a few times (for `ncu --set full --import-source on -k regex:tapgemm_tc2 -s 4 -c 1`).

    python tools/one_shape.py hr2 | hr6 | deconv | conv8 | deconv_bwd
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
    N, h, w, F = 32, 32, 32, 64
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
