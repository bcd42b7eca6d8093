#!/usr/bin/env python
"""bench.py — headline benchmark: HR voxels/s of the DRFNet-L training step (BASELINE.json
configs[1]: DRFNet-L F=64 G=6, x4, batch of 32 cropped 2D+t patches LR 32x32, T=5 frames, bf16).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arm (oracle port)

A step = forward over T frames + fused L1 loss + full BPTT backward + (NCCL all-reduce) + Adam +
PSNR/SSIM of the training outputs (what acdc_vsr_trainer.py:41-55 does per batch).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL writes its version banner / debug lines to stdout by default: send them to stderr so that rank 0's
# stdout is exactly one JSON line whatever NCCL_DEBUG the box sets
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

MODEL = dict(in_channels=1, out_channels=1, num_features=64, num_groups=6, upscale_factor=4)
BATCH, T, LR = 32, 5, 32
METRIC, UNIT = "hr_voxels_per_s_train_step", "HR voxels/s"
# algorithmic forward FLOPs per LR pixel per frame of DRFNet-L x4; fwd+bwd = 3x
FWD_FLOPS_PER_LR_PIXEL = 10.673e6


# second workload (--workload duf): the Conv3d network of SURVEY §8 row a15, same contract
DUF_MODEL = dict(in_channels=1, out_channels=1, num_frames=7, size_filter=5, upscale_factor=4, backbone="_DenseLayer16")
WORKLOAD = "drf"


def duf_fwd_flops_per_sample():
    """algorithmic forward FLOPs of DUFNet-16 x4 per sample (7 frames of LR x LR): 2 * out pixels * Cout * Cin * taps"""
    px, G, C = LR * LR, 32, [64 + 32 * i for i in range(7)]
    total = 2.0 * px * 7 * 64 * 1 * 9
    for i in range(6):
        tin = 7 if i < 3 else 7 - 2 * (i - 3)
        tout = tin if i < 3 else tin - 2
        total += 2.0 * px * tin * C[i] * C[i] + 2.0 * px * tout * G * C[i] * 27
    return total + 2.0 * px * (256 * 256 * 9 + 256 * 768 + 512 * 400 + 256 * 16)


def workload_name(batch):
    if WORKLOAD == "duf":
        return (f"DUFNet-16 (Conv3d path) x4 train step, batch {batch} x 7 frames, LR {LR}x{LR} -> HR {4 * LR}x{4 * LR}, "
                "L1 + Adam + PSNR/SSIM on training outputs")
    return (f"C2: DRFNet-L(F64,G6) x4 train step, batch {batch} x T{T}, LR {LR}x{LR} -> HR {4 * LR}x{4 * LR}, "
            "L1 + Adam + PSNR/SSIM on training outputs")


def frames_in():
    return 7 if WORKLOAD == "duf" else T


def hr_voxels(batch):
    """HR voxels produced per step: T frames per sequence (DRFNet), one centre frame per sequence (DUFNet)"""
    return batch * (4 * LR) ** 2 * (1 if WORKLOAD == "duf" else T)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p.update(json.load(f))
            p["source"] = "measured"
    except OSError:
        pass
    return p


# Whatever a library writes to the process's stdout (NCCL prints its version banner there on some boxes, even
# with NCCL_DEBUG_FILE set) goes to stderr: file descriptor 1 is pointed at stderr for the whole run and the ONE
# JSON line is written to the saved original descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


class ClockSampler(threading.Thread):
    """samples SM clocks and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_batches(n_batches, batch, seed, pinned):
    """synthetic normalised LR/HR cine patches, [T][N,1,h,w] / [T][N,1,4h,4w] fp32, on the host."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        hr = torch.rand(batch, 1, 4 * LR, 4 * LR, generator=g) * 255
        lrs, hrs = [], []
        for t in range(frames_in()):
            f = (hr * (0.7 + 0.06 * t)).round().clamp(0, 255)
            l = torch.nn.functional.avg_pool2d(f, 4).round()
            hrs.append(((f - 54.089) / 48.084).contiguous())
            lrs.append(((l - 54.089) / 48.084).contiguous())
        if WORKLOAD == "duf":
            hrs = [hrs[3]]                       # the MISR target: the centre HR frame (acdc_misr_trainer.py:24)
        if pinned:
            lrs, hrs = [x.pin_memory() for x in lrs], [x.pin_memory() for x in hrs]
        out.append((lrs, hrs))
    return out


def cpu_port(sample):
    """the reference's algorithm (oracle/restated.py) as one training step on the host: returns (step fn, voxels)"""
    from oracle import restated
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1: use every host core
    torch.manual_seed(0)
    lrs, hrs = make_batches(1, sample, 0, False)[0]
    if WORKLOAD == "duf":
        from vsr_b200.duf import DUFNet
        net = DUFNet(**DUF_MODEL)            # parameter container only (same init as the reference class)
        sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
              for k, v in net.state_dict().items()}
        opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-4)

        def step():
            out = restated.dufnet_forward(lrs, sd, 5, 4, training=True)
            loss = restated.l1_loss(out, hrs[0])
            opt.zero_grad()
            loss.backward()
            opt.step()
            restated.vsr_metrics([out.detach()], hrs)
            return float(loss.detach())
    else:
        from vsr_b200.nets import DRFNet
        net = DRFNet(**MODEL)            # parameter container only (same init as the reference class)
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
        opt = torch.optim.Adam(list(sd.values()), lr=1e-4)

        def step():
            outs = restated.drfnet_forward(lrs, sd, 4)
            loss = restated.vsr_loss(outs, hrs, restated.l1_loss)
            opt.zero_grad()
            loss.backward()
            opt.step()
            restated.vsr_metrics([o.detach() for o in outs], hrs)
            return float(loss.detach())
    return step, hr_voxels(sample)


def run_reference(args):
    """Reference arm: the reference's algorithm on the host CPU (oracle/restated.py port — the real
    reference needs /root/reference, which does not exist on the GPU box), all host threads, on a
    bounded sample of the same workload (2 of the 32 patches per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    step, vox = cpu_port(sample)
    cores = torch.get_num_threads()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = vox / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(BATCH), "sample": f"{sample} of {BATCH} patches per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} of {BATCH} patches x T{frames_in()} per step, torch {torch.__version__} CPU fp32"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline(budget_s=20.0):
    """oracle port timed on this box's host cores on a bounded sample (rank 0, N=1 only)."""
    sample = 2
    step, vox = cpu_port(sample)
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 and (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    dt = min(times)
    return {"value": vox / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample} of {BATCH} patches x T{frames_in()}, best of {len(times)} steps, torch {torch.__version__} CPU fp32"}


def psnr_delta(precision, dev):
    """BASELINE.json's "PSNR delta vs CPU ref": PSNR (after denormalize, src/utils.py:1-20) of the GPU path's output and
    of the CPU reference port's output, both against the synthetic target, on the same inputs (2 patches) and the same
    initial weights (seed 0, the reference's default initialisation); plus the largest tensor-normalised output error."""
    from oracle import restated
    lrs, hrs = make_batches(1, 2, 0, False)[0]
    torch.manual_seed(0)
    with torch.no_grad():
        if WORKLOAD == "duf":
            from vsr_b200.duf import DUFNet
            net = DUFNet(precision=precision, **DUF_MODEL)
            sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
            refs = [restated.dufnet_forward(lrs, sd, 5, 4, training=True)]
            net = net.to(dev).train()
            outs = [net([x.to(dev) for x in lrs]).float().cpu()]
        else:
            from vsr_b200.nets import DRFNet
            net = DRFNet(precision=precision, **MODEL)
            sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
            refs = restated.drfnet_forward(lrs, sd, 4)
            net = net.to(dev)
            outs = [o.float().cpu() for o in net([x.to(dev) for x in lrs])]
        den = lambda t: restated.denormalize(t, "acdc")
        p_ref = sum(float(restated.psnr(den(r), den(h))) for r, h in zip(refs, hrs)) / len(refs)
        p_gpu = sum(float(restated.psnr(den(o), den(h))) for o, h in zip(outs, hrs)) / len(outs)
        err = max(float((o - r).abs().max() / r.abs().max()) for o, r in zip(outs, refs))
    return {"value": p_gpu - p_ref, "unit": "dB", "gpu_db": p_gpu, "cpu_ref_db": p_ref, "max_rel_output_error": err,
            "sample": "2 patches, all frames, initial weights (seed 0), oracle port on the host"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="drf", choices=["drf", "duf"],
                    help="drf = the headline (BASELINE configs[1], DRFNet-L); duf = the Conv3d network (DUFNet-16)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.nets import DRFNet
    from vsr_b200.ops import cuda_ops
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import VSRTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)                               # identical initial weights on every rank
    if WORKLOAD == "duf":
        from vsr_b200.duf import DUFNet
        from vsr_b200.runner import MISRTrainStep as StepCls
        net = DUFNet(precision=args.precision, **DUF_MODEL).to(dev).train()
    else:
        StepCls = VSRTrainStep
        net = DRFNet(precision=args.precision, **MODEL).to(dev)
    opt = FlatAdam(net.parameters(), lr=1e-4)
    step = StepCls(net, [torch.nn.L1Loss()], [1.0], [PSNR().to(dev), SSIM().to(dev)], opt, "acdc",
                   use_graph=not args.no_graph)
    ops = cuda_ops()

    n_host = 4
    host = make_batches(n_host, args.batch, seed=1234 + rank, pinned=True)   # weak scaling: own shard per rank
    dev_batches = [([x.to(dev) for x in l], [y.to(dev) for y in h]) for l, h in host]
    acc = torch.zeros(4, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    for i in range(args.warmup):
        step.train_step(*dev_batches[i % n_host], acc)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step.train_step(*dev_batches[i % n_host], acc)
    e1.record()
    barrier()
    launches = ops.launches - l0                       # 0 when the step is replayed as a CUDA graph
    ms = e0.elapsed_time(e1) / args.steps
    # per-kernel pass: the same steps again with a CUDA-event pair around every tap-GEMM / wgrad
    # launch (kept out of the timed region above so that `value` carries no event overhead)
    prof_steps = min(args.steps, 3)
    use_graph, step.use_graph = step.use_graph, False
    ops.timing = []
    l1 = ops.launches
    for i in range(prof_steps):
        step.train_step(*dev_batches[i % n_host], acc)
    launches_per_step = (ops.launches - l1) // prof_steps
    if launches == 0:                                  # graph replay: same kernels, launched by the graph
        launches = launches_per_step * args.steps
    step.use_graph = use_graph
    barrier()
    timing, ops.timing = ops.timing, None

    # ---- end to end: pinned host inputs in, loss value out, every step ------------------------
    stage = [([torch.empty_like(x, device=dev) for x in host[0][0]], [torch.empty_like(y, device=dev) for y in host[0][1]])
             for _ in range(2)]
    loss_host = torch.zeros(1).pin_memory()
    h2d = sum(x.numel() * 4 for x in host[0][0]) + sum(y.numel() * 4 for y in host[0][1])

    def e2e_step(i):
        lrs, hrs = host[i % n_host]
        dl, dh = stage[i % 2]
        for d, s in zip(dl, lrs):
            d.copy_(s, non_blocking=True)
        for d, s in zip(dh, hrs):
            d.copy_(s, non_blocking=True)
        lv, _ = step.train_step(dl, dh, acc)
        loss_host.copy_(lv[:1], non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the user reads the loss value
        return float(loss_host[0])

    for i in range(2):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        last_loss = e2e_step(i)
    f1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_e2e = f0.elapsed_time(f1) / args.steps

    t = torch.tensor([ms, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    vox = world * hr_voxels(args.batch)
    value, value_e2e = vox / (ms * 1e-3), vox / (ms_e2e * 1e-3)

    if rank == 0:
        pk = peaks()
        # dominant kernel family: the tcgen05 tap-GEMM (all forward / data-gradient convolutions)
        agg, detail = {}, {}
        for kind, flops, a, b, sig, nbytes in timing:
            t_ms = a.elapsed_time(b)
            d = agg.setdefault(kind, [0.0, 0.0, 0])
            d[0] += flops
            d[1] += t_ms
            d[2] += 1
            q = detail.setdefault(f"{kind}:{sig}", [0.0, 0.0, 0, 0.0])
            q[0] += flops
            q[1] += t_ms
            q[2] += 1
            q[3] += nbytes
        dom = max(agg, key=lambda k: agg[k][1])
        traffic = None
        try:       # per-launch dram__bytes_read+write of the same step under ncu (tools/ncu_summary.py)
            with open(os.path.join(ROOT, "profiles", "r01_duf_by_kernel_v3.json" if WORKLOAD == "duf" else "r01_dram_by_kernel_v9.json")) as f:
                traffic = json.load(f).get({"tapgemm": "tapgemm_tc2_kernel", "wgrad": "wgrad_tc_kernel"}.get(dom, dom), {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        flops, kms, cnt = agg[dom]
        achieved = flops / (kms * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        kernel_share = {k: {"ms_per_step": v[1] / prof_steps, "launches_per_step": v[2] / prof_steps,
                            "tflops": v[0] / (v[1] * 1e-3) / 1e12} for k, v in agg.items()}
        top = sorted(detail.items(), key=lambda kv: -kv[1][1])[:14]

        def shape_row(v):
            # every shape against ITS OWN bound: tensor if its algorithmic intensity is above the ridge of the
            # measured peaks (FLOP/B), else HBM
            tf, gb = v[0] / (v[1] * 1e-3) / 1e12, v[3] / (v[1] * 1e-3) / 1e9
            tensor = v[0] / max(v[3], 1.0) > pk["bf16_tflops_sustained"] * 1e3 / pk["hbm_gbs"]
            return {"ms_per_step": v[1] / prof_steps, "n_per_step": v[2] / prof_steps, "tflops": tf, "gbs": gb,
                    "bound": "tensor" if tensor else "hbm",
                    "frac": tf / pk["bf16_tflops_sustained"] if tensor else gb / pk["hbm_gbs"]}

        kernel_detail = {k: shape_row(v) for k, v in top}
        all_rows = [shape_row(v) for k, v in detail.items() if k.startswith(dom)]
        t_all = sum(r["ms_per_step"] for r in all_rows)
        # the dominant kernel runs tensor-bound and HBM-bound shapes: time-weighted mean of every shape's
        # fraction of ITS OWN bound, and the split of its time between the two bounds
        frac_own = sum(r["ms_per_step"] * r["frac"] for r in all_rows) / max(t_all, 1e-9)
        hbm_share = sum(r["ms_per_step"] for r in all_rows if r["bound"] == "hbm") / max(t_all, 1e-9)
        step_flops = 3.0 * FWD_FLOPS_PER_LR_PIXEL * args.batch * T * LR * LR
        if WORKLOAD == "duf":
            step_flops = 3.0 * duf_fwd_flops_per_sample() * args.batch
        try:      # full per-shape table for the profile notes (scratch; the JSON line keeps the top 14)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "kernel_detail_full.json"), "w") as f:
                json.dump({k: shape_row(v) for k, v in sorted(detail.items(), key=lambda kv: -kv[1][1])}, f, indent=1)
        except OSError:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.batch), "per_gpu_batch": args.batch, "frames": frames_in(),
                       "parallelism": f"dp{world}", "cuda_graph": bool(step.use_graph),
                       "l2": ("per-step working set (~5 GB of activations) exceeds the 126 MB L2; inputs rotate over 4 batches" if WORKLOAD == "drf"
                              else "per-step working set (~1.5 GB of activations and gradients) exceeds the 126 MB L2; inputs rotate over 4 batches")},
            "clocks": sampler.result(),
            "e2e": {"value": value_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e, "last_loss": last_loss},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": sum(v[3] for k, v in detail.items() if k.startswith(dom)) / max(cnt, 1),
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['source']})",
                         "frac_time_weighted_own_bound": frac_own, "time_share_hbm_bound_shapes": hbm_share,
                         "launches": cnt, "kernels": kernel_share, "kernel_detail": kernel_detail,
                         "timing_pass": f"{prof_steps} extra steps of the same workload, CUDA events around each launch",
                         "step_tflops_algorithmic": step_flops / (ms * 1e-3) / 1e12 * world},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            try:
                line["psnr_delta"] = psnr_delta(args.precision, dev)
            except Exception as e:      # a reporting extra must never cost the bench line
                line["psnr_delta"] = {"error": f"{type(e).__name__}: {e}"}
        emit(line)
    # hang-proof exit: synchronise, then leave without tearing NCCL / CUDA graphs down
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
