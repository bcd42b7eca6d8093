"""BASELINE configs[4]: conv / upsample kernel shape sweep (channels 64-128, patches 32^2-64^2 x time).  Every configuration
runs the SAME measurement as the headline bench (bench.train_bench: CUDA-graphed training step of the drop-in DRFNet timed
with CUDA events after warm-up, clocks sampled during the timed region, then the per-kernel pass whose table sums to the
graphed step) with the net / patch / frame counts swapped in; every distinct tap-GEMM / weight-gradient shape is reported
against its own bound (tensor: sustained bf16 peak, hbm: measured copy bandwidth).  The bandwidth kernels (up-sampling, loss,
metrics, Adam, ...) have their own sweep with clock records: tools/bw_bench.py -> profiles/r02_bw_kernels.json.

    python tools/sweep.py --json gpurun_out/sweep_config5.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vsr_b200.ops import cuda_ops  # noqa: E402

CONFIGS = [  # F, d (slices = batch), s (LR side), T
    (64, 32, 32, 5), (64, 48, 48, 5), (64, 64, 64, 5), (64, 32, 32, 7),
    (128, 32, 32, 5), (128, 48, 48, 5), (128, 32, 32, 7),
]


def fwd_flops_per_lr_pixel(F, G, r=4, k=8):
    """algorithmic MACs x 2 of one DRFNet forward per LR pixel and frame (drf_net.py:55-147; SURVEY 8d gives 10.673 MFLOP for
    F = 64, G = 6, x4): input block, feedback block (1x1 on concatenations, k x k stride-r projections), output block"""
    r2 = r * r
    macs = 9 * 4 * F + 4 * F * F + 2 * F * F                      # in_block conv1 (Cin = 1), conv2, f_block.in_block
    for g in range(G):
        if g:
            macs += (g + 1) * F * F                               # up conv1 on the LR concat
            macs += (g + 1) * F * F * r2                          # down conv1 on the HR concat
        macs += 2 * k * k * F * F                                 # transposed + strided projection
    macs += G * F * F                                             # f_block.out_block
    res = 1
    for _ in range({2: 1, 4: 2, 8: 3}[r]):
        macs += 9 * F * 4 * F * res                               # out_block conv + PixelShuffle(2)
        res *= 4
    macs += 9 * F * res                                           # last 3x3 convolution (Cout = 1)
    return 2.0 * macs


def one(F, d, s, T, steps, warmup):
    dev = torch.device("cuda", 0)
    bench.MODEL = dict(bench.MODEL, num_features=F)
    bench.BATCH, bench.T, bench.LR = d, T, s
    ns = argparse.Namespace(batch=d, no_graph=False, steps=steps, warmup=warmup, precision="bf16", gpus=1)
    barrier = torch.cuda.synchronize
    r = bench.train_bench(ns, "bf16", steps, warmup, 1, 0, dev, cuda_ops(), barrier, with_e2e=False, with_kernels=True)
    try:
        bench.build_line(ns, r, 1)                                # writes the full per-shape table (scratch file) ...
    except KeyError:
        pass                                                      # ... before it asks for the end-to-end numbers this run skips
    with open(os.path.join(ROOT, "gpurun_out", "kernel_detail_full.json")) as f:
        det = json.load(f)
    flops = 3.0 * fwd_flops_per_lr_pixel(F, bench.MODEL["num_groups"]) * d * T * s * s
    rows = [v for k, v in det["shapes"].items() if k.startswith("tapgemm")]
    t_all = sum(v["ms_per_step"] for v in rows)
    tens = [v for v in rows if v["bound"] == "tensor"]
    return {"config": {"F": F, "G": bench.MODEL["num_groups"], "slices": d, "lr": s, "frames": T, "upscale": 4, "precision": "bf16"},
            "ms_per_step": r["ms"], "ms_per_step_median": r["ms_median"], "cuda_graph": r["cuda_graph"],
            "hr_voxels_per_s": d * T * (4 * s) ** 2 / r["ms"] * 1e3, "step_tflops_algorithmic": flops / (r["ms"] * 1e-3) / 1e12,
            "clocks": r["clocks"], "kernels": det["kernels"],
            "tapgemm_time_share_hbm_bound_shapes": sum(v["ms_per_step"] for v in rows if v["bound"] == "hbm") / max(t_all, 1e-9),
            "tapgemm_frac_tensor_bound_shapes": (sum(v["ms_per_step"] * v["frac"] for v in tens) / max(sum(v["ms_per_step"] for v in tens), 1e-9)),
            "tapgemm_frac_time_weighted_own_bound": sum(v["ms_per_step"] * v["frac"] for v in rows) / max(t_all, 1e-9),
            "shapes": det["shapes"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    args = ap.parse_args()
    assert abs(fwd_flops_per_lr_pixel(64, 6) - bench.FWD_FLOPS_PER_LR_PIXEL) <= 0.01 * bench.FWD_FLOPS_PER_LR_PIXEL
    out = []
    for F, d, s, T in CONFIGS:
        r = one(F, d, s, T, args.steps, args.warmup)
        out.append(r)
        print(f"F={F} slices={d} LR={s}x{s} T={T}: {r['ms_per_step']:.2f} ms/step (graphed) = {r['hr_voxels_per_s'] / 1e6:.0f} M HR voxels/s, "
              f"{r['step_tflops_algorithmic']:.0f} TFLOP/s algorithmic, tensor-bound shapes {r['tapgemm_frac_tensor_bound_shapes']:.2f} of "
              f"the sustained bf16 peak, clocks {r['clocks']['sm_mhz']} MHz {r['clocks']['reasons']}", flush=True)
        for k, v in list(r["shapes"].items())[:6]:
            print(f"    {k:50s} {v['us_per_launch']:8.1f} us  {v['bound']:6s} {v['frac']:.2f}", flush=True)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
