"""Fixed per-launch cost of the tap-GEMM: the same layer at shrinking batch sizes (tiles per CTA -> 0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kbench  # noqa: E402
from vsr_b200.ops import TapTable, cuda_ops  # noqa: E402


def main():
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    ops = cuda_ops()
    dt = torch.bfloat16
    F = 64
    for n in (32, 8, 4, 1):
        res = []
        kbench.CASES = None
        kbench.tapgemm_case(f"conv1x1_lr_cat3_n{n}", TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(3)])]),
                            n, 32, 32, [F] * 3, F, dt, 20, flush, res)
    # back-to-back launches without flush: average per launch
    for n in (32, 1):
        tab = TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(3)])])
        srcs = [torch.randn(n, 32, 32, F, device="cuda").to(dt) for _ in range(3)]
        out = torch.empty(n, 32, 32, F, device="cuda", dtype=dt)
        wts = (torch.randn(3 * 64 * 64, device="cuda") * 0.05).to(dt)
        bias = torch.zeros(F, device="cuda")
        slope = torch.tensor([0.2], device="cuda")
        fn = lambda: ops.tapgemm(tab, srcs, out, wts, bias=bias, epi=5, slope=slope)
        for _ in range(3):
            fn()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(100):
                    fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        print(f"graph of 100 launches, n={n}: {a.elapsed_time(b) * 10:.2f} us per launch", flush=True)
    # an empty-ish elementwise kernel for reference
    z = torch.zeros(1024, device="cuda")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        z.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(100):
                z.add_(1.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"graph of 100 tiny torch kernels: {a.elapsed_time(b) * 10:.2f} us per launch", flush=True)





def alternating():
    """Does alternating kernels with different shared-memory carve-outs cost extra per launch?"""
    ops = cuda_ops()
    dt = torch.bfloat16
    F = 64
    n = 1
    tab = TapTable(64, 64, [(0, [(s, 0, 0, 0) for s in range(3)])])
    srcs = [torch.randn(n, 32, 32, F, device="cuda").to(dt) for _ in range(3)]
    out = torch.empty(n, 32, 32, F, device="cuda", dtype=dt)
    wts = (torch.randn(3 * 64 * 64, device="cuda") * 0.05).to(dt)
    bias = torch.zeros(F, device="cuda")
    slope = torch.tensor([0.2], device="cuda")
    z = torch.zeros(1024, device="cuda")
    z1, z2, z3 = (torch.zeros(4096, device="cuda", dtype=dt) for _ in range(3))
    fa = lambda: ops.tapgemm(tab, srcs, out, wts, bias=bias, epi=5, slope=slope)
    variants = {"tapgemm only": [fa], "tapgemm + torch add_": [fa, lambda: z.add_(1.0)],
                "tapgemm + vsr add": [fa, lambda: ops.add(z1, z2, z3)], "torch add_ only": [lambda: z.add_(1.0)],
                "vsr add only": [lambda: ops.add(z1, z2, z3)]}
    for name, fns in variants.items():
        for f in fns:
            f()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(100):
                    for f in fns:
                        f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        print(f"{name}: {a.elapsed_time(b) * 10:.2f} us per round of {len(fns)} launches", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "alt":
        alternating()
    else:
        main()
