"""Timing attribution of the tcgen05 tap-GEMM: re-times the config-2 convolution shapes with parts of
the kernel switched off (needs the attribution build, `python -m vsr_b200.build --attrib`; VSR_TC_DEBUG bits: 1 no epilogue, 16 no epilogue stores, 2 no A loads,
4 no B loads, 8 no MMAs).  Results are WRONG by construction; only the times mean anything.

    python tools/attrib.py --json gpurun_out/attrib.json
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kbench  # noqa: E402
from vsr_b200.ops import cuda_ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cases", default="deconv8x8s4,conv8x8s4,conv3x3_n256_2x,conv1x1_hr_cat6")
    ap.add_argument("--modes", default="0,16,1,2,4,6,8,7,15")
    args = ap.parse_args()
    out = []
    for case in args.cases.split(","):
        for res in ("0", "1"):
            for mode in args.modes.split(","):
                os.environ["VSR_TC_DEBUG"] = mode
                os.environ["VSR_TC_RESIDENT"] = res
                cuda_ops().lib.vsr_reload_tunables()
                sys.argv = ["kbench", "--cases", case, "--iters", str(args.iters)]
                results = []
                kbench.CASES = {case}
                flush = torch.zeros(64 * 1024 * 1024, device="cuda")
                import io
                import contextlib
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    run_cases(args.iters, flush, results)
                for r in results:
                    r["debug"] = int(mode)
                    r["resident"] = int(res)
                    out.append(r)
                    print(f"{r['kernel']:18s} res={res} dbg={int(mode):2d}  {r['ms'] * 1e3:8.1f} us  {r['tflops']:7.1f} TF/s", flush=True)
    os.environ["VSR_TC_DEBUG"] = "0"
    cuda_ops().lib.vsr_reload_tunables()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


def run_cases(iters, flush, results):
    from vsr_b200.ops import TapTable
    N, h, w, F = 32, 32, 32, 64
    dt = torch.bfloat16
    kbench.tapgemm_case("conv1x1_hr_cat6", TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(6)])]),
                        N, h, w * 16, [F] * 6, F, dt, iters, flush, results)
    groups = []
    for g in range(4):
        gy, gx = g // 2, g % 2
        groups.append((g * 256, [(0, dy - 1 + gy, dx - 1 + gx, 0) for dy in (0, 1) for dx in (0, 1)]))
    kbench.tapgemm_case("deconv8x8s4", TapTable(64, 256, groups), N, h, w, [F], 16 * F, dt, iters, flush, results)
    taps = []
    for ky in range(8):
        for kx in range(8):
            dy, py = divmod(ky - 2, 4)
            dx, px = divmod(kx - 2, 4)
            taps.append((0, dy, dx, (py * 4 + px) * 64))
    kbench.tapgemm_case("conv8x8s4", TapTable(64, 64, [(0, taps)]), N, h, w, [16 * F], F, dt, iters, flush, results)
    t33 = [(0, dy, dx, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    kbench.tapgemm_case("conv3x3_n256_2x", TapTable(64, 256, [(0, t33)]), N, 2 * h, 2 * w, [F], 4 * F, dt, iters, flush, results)


if __name__ == "__main__":
    main()
