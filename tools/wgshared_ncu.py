"""Two launches of wgrad_shared at the config-2 shapes of the down-projection 1x1 layers (for an ncu --set full capture):
sets {1,2} and {3,4,5} of a 6-group net, all 5 frames stacked: [160, 32, 512, 64] bf16 maps of 335 MB."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vsr_b200.ops import cuda_ops   # noqa: E402

ops = cuda_ops()
g = torch.Generator(device="cuda").manual_seed(0)
srcs = [torch.randn(160, 32, 512, 64, device="cuda", generator=g).bfloat16() for _ in range(6)]
dzs = [torch.randn(160, 32, 512, 64, device="cuda", generator=g).bfloat16() for _ in range(3)]
ws = {}


def workspace_of(n):
    if "t" not in ws or ws["t"].numel() * 4 < n:
        ws["t"] = torch.empty((n + 3) // 4, device="cuda")
    return ws["t"]


for ntaps in ([2, 3], [4, 5, 6]):
    dws = [torch.zeros(nt, 64, 64, device="cuda") for nt in ntaps]
    dbs = [torch.zeros(64, device="cuda") for _ in ntaps]
    for _ in range(3):
        ops.wgrad_shared(srcs[:max(ntaps)], dzs[:len(ntaps)], ntaps, dws, dbs, False, workspace_of)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.wgrad_shared(srcs[:max(ntaps)], dzs[:len(ntaps)], ntaps, dws, dbs, False, workspace_of)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = 2 * 160 * 32 * 512 * 64 * (max(ntaps) + len(ntaps)) / 1e9
    print(f"ntaps {ntaps}: {ms * 1e3:.1f} us per call (kernel + reduce), {gb:.2f} GB -> {gb / ms * 1e3:.0f} GB/s")
