"""Training-step throughput of the OTHER drop-in nets (EDSRNet, SRFBNet, DRFSISRNet, RBPNet,
FRVSRNet, TOFlowNet) with the measurement rules of bench.py: the fused, CUDA-graphed training step of each net (forward +
loss + backward + FlatAdam + PSNR / SSIM) on ACDC-shaped synthetic patches, inputs resident, W warm-up steps, K timed steps
between CUDA events, SM clocks and throttle reasons sampled during the timed region; beside each, the reference's CPU
algorithm (oracle/restated.py + torch.optim.Adam, the same step) on the box's host cores on a bounded sample.

    python tools/nets_bench.py --json gpurun_out/nets_bench.json
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import restated  # noqa: E402  (CPU baseline leg only)
from vsr_b200.metrics import PSNR, SSIM  # noqa: E402
from vsr_b200.optim import FlatAdam  # noqa: E402

LR, R = 32, 4


def frames(n, t, seed, hr=False):
    g = torch.Generator().manual_seed(seed)
    s = LR * R if hr else LR
    base = torch.randn(n, 1, s, s, generator=g)
    return [base + 0.2 * torch.randn(n, 1, s, s, generator=g) for _ in range(t)]


def cases():
    from vsr_b200.edsr import EDSRNet
    from vsr_b200.frvsr import FRVSRNet
    from vsr_b200.nets import DRFSISRNet, SRFBNet
    from vsr_b200.rbpn import RBPNet
    from vsr_b200.runner import FRVSRTrainStep, MISRTrainStep, SISRSRFBTrainStep, SISRTrainStep
    from vsr_b200.toflow import TOFlowNet
    l1 = torch.nn.L1Loss
    out = []
    # (name, batch, make net(precision), step class, losses, inputs(n) -> (inputs, targets), oracle forward(sd, inputs) -> outputs,
    #  output frames per sample, precisions)
    out.append(("EDSRNet(F64, 16 blocks) x4", 32, lambda p: EDSRNet(1, 1, 16, 64, 4, precision=p), SISRTrainStep, [l1()],
                lambda n: (frames(n, 1, 1), frames(n, 1, 2, True)),
                lambda sd, x: [restated.edsrnet_forward(x[0], sd, 4)], 1, ("bf16", "bf16x3", "fp32")))
    out.append(("SRFBNet(F64, G6, 4 steps) x4", 32, lambda p: SRFBNet(1, 1, 4, 64, 6, 4, precision=p), SISRSRFBTrainStep, [l1()],
                lambda n: (frames(n, 1, 3), frames(n, 1, 4, True)),
                lambda sd, x: restated.srfbnet_forward(x[0], sd, 4, 4), 1, ("bf16", "bf16x3")))
    out.append(("DRFSISRNet(F64, G6, 4 steps) x4", 32, lambda p: DRFSISRNet(1, 1, 4, 64, 6, 4, precision=p), SISRSRFBTrainStep, [l1()],
                lambda n: (frames(n, 1, 5), frames(n, 1, 6, True)),
                lambda sd, x: restated.drfnet_forward([x[0]] * 4, sd, 4), 1, ("bf16", "bf16x3")))
    out.append(("RBPNet(64, 64, 3 stages, 5 blocks, 7 frames) x4", 8,
                lambda p: RBPNet(1, 1, 64, 64, 3, 5, 7, 4, precision=p), MISRTrainStep, [l1()],
                lambda n: (frames(n, 7, 7), frames(n, 1, 8, True)),
                lambda sd, x: [restated.rbpnet_forward(x, sd, 4, 7)], 1, ("bf16", "bf16x3", "fp32")))
    out.append(("FRVSRNet(10 blocks) x4, 5 frames", 8, lambda p: FRVSRNet(1, 1, 4, num_resblocks=10, precision=p), FRVSRTrainStep,
                [l1(), l1()], lambda n: (frames(n, 5, 9), frames(n, 5, 10, True)),
                lambda sd, x: restated.frvsrnet_forward(x, sd, 4), 5, ("bf16x3", "fp32")))
    out.append(("TOFlowNet(7 frames) x4", 8, lambda p: TOFlowNet(1, 1, 7, 4), MISRTrainStep, [torch.nn.MSELoss()],
                lambda n: (frames(n, 7, 11), frames(n, 1, 12, True)),
                lambda sd, x: [restated.toflownet_forward(x, sd, 4, True, None)], 1, ("fp32",)))
    return out


def gpu_leg(make_net, Step, losses, data, batch, precision, steps, warmup, dev):
    torch.manual_seed(0)
    net = make_net(precision).to(dev).train()
    opt = FlatAdam(net.parameters(), lr=1e-4)
    weights = [1.0] * len(losses)
    step = Step(net, losses, weights, [PSNR().to(dev), SSIM().to(dev)], opt, "acdc", use_graph=True)
    x, y = data(batch)
    x, y = [t.to(dev) for t in x], [t.to(dev) for t in y]
    acc = torch.zeros(1 + len(losses) + 2, device=dev)
    for _ in range(warmup):
        step.train_step(x, y, acc)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(dev.index or 0)
    sampler.start()
    ms, med, _ = bench.time_steps(lambda i: step.train_step(x, y, acc), steps, torch.cuda.synchronize)
    clocks = sampler.finish()
    loss = float(acc[0]) / max(steps + warmup, 1)
    n_params = sum(p.numel() for p in net.parameters())
    del step, net, opt
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "ms_per_step_median": med, "cuda_graph": True, "clocks": clocks, "mean_loss": loss, "parameters": n_params}


def cpu_leg(make_net, losses, data, oracle, sample, budget_s=20.0):
    """the reference's algorithm on the host: oracle forward + loss + backward + torch.optim.Adam on `sample` patches"""
    torch.manual_seed(0)
    sd = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and "running_" not in k else v.clone())
          for k, v in make_net("fp32").state_dict().items()}
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-4)
    x, y = data(sample)

    def one():
        outs = oracle(sd, x)
        if isinstance(outs, tuple):                       # FRVSRNet: (sr frames, warped lr frames)
            sr, lr = outs
            loss = torch.stack([losses[0](a, b) for a, b in zip(lr, x)]).mean() + torch.stack([losses[1](a, b) for a, b in zip(sr, y)]).mean()
        else:
            tg = y if len(y) == len(outs) else [y[0]] * len(outs)
            loss = torch.stack([losses[0](a, b) for a, b in zip(outs, tg)]).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()

    one()
    times, t_start = [], time.perf_counter()
    while len(times) < 3 and time.perf_counter() - t_start < budget_s:
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    for name, batch, make_net, Step, losses, data, oracle, out_frames, precisions in cases():
        vox = batch * out_frames * (LR * R) ** 2
        row = {"net": name, "workload": f"training step, batch {batch} of LR {LR}x{LR} -> HR {LR * R}x{LR * R}, {out_frames} output frame(s) per sample",
               "hr_voxels_per_step": vox, "modes": {}}
        for p in precisions:
            r = gpu_leg(make_net, Step, losses, data, batch, p, args.steps, args.warmup, dev)
            r["hr_voxels_per_s"] = vox / r["ms_per_step"] * 1e3
            row["modes"][p] = r
            print(f"{name:52s} {p:7s} {r['ms_per_step']:9.2f} ms/step  {r['hr_voxels_per_s'] / 1e6:8.1f} M HR voxels/s  "
                  f"clocks {r['clocks']['sm_mhz']} MHz {r['clocks']['reasons']}", flush=True)
        if not args.no_cpu:
            sample = 2
            t = cpu_leg(make_net, losses, data, oracle, sample)
            row["cpu_baseline"] = {"value": sample * out_frames * (LR * R) ** 2 / t, "unit": "HR voxels/s", "cores": torch.get_num_threads(),
                                   "kind": "port", "sample": f"{sample} of {batch} patches, median of <= 3 steps, {t * 1e3:.0f} ms per step"}
            print(f"{'':52s} cpu     {row['cpu_baseline']['value'] / 1e6:8.3f} M HR voxels/s on {row['cpu_baseline']['cores']} cores", flush=True)
        rows.append(row)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
