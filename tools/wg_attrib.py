"""Per-role cycle counts of the tcgen05 weight-gradient kernel at the config-2 strided-conv shape
(needs the attribution build: `python -m vsr_b200.build --attrib`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsr_b200.ops import TapTable, cuda_ops  # noqa: E402


def main():
    ops = cuda_ops()
    N, h, w, F = 160, 32, 32, 64          # the T=5 frames of config 2 stacked (one launch per layer per step)
    taps = []
    for ky in range(8):
        for kx in range(8):
            dy, py = divmod(ky - 2, 4)
            dx, px = divmod(kx - 2, 4)
            taps.append((0, dy, dx, (py * 4 + px) * 64))
    tab = TapTable(64, 64, [(0, taps)])
    src = torch.randn(N, h, w, 16 * F, device="cuda").to(torch.bfloat16)
    dz = torch.randn(N, h, w, F, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(64 * 64 * 64, device="cuda")
    ws = torch.empty(ops.tapgemm_wgrad_workspace(tab, [src], dz) // 4 + 4, device="cuda")
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    for tall in ("0", "1"):
        os.environ["VSR_WG_TALL"] = tall
        os.environ["VSR_WG_DEBUG"] = "0"
        ops.lib.vsr_reload_tunables()
        for _ in range(3):
            ops.tapgemm_wgrad(tab, [src], dz, dw, False, ws)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.add_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.tapgemm_wgrad(tab, [src], dz, dw, False, ws)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print(f"wgrad conv8x8s4 tall={tall}: {ts[len(ts) // 2] * 1e3:.1f} us (kernel + split reduce)", flush=True)
        os.environ["VSR_WG_DEBUG"] = "32"
        ops.lib.vsr_reload_tunables()
        ops.tapgemm_wgrad(tab, [src], dz, dw, False, ws)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
