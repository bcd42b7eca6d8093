#!/usr/bin/env python
"""bench.py — headline benchmark: HR voxels/s of the DRFNet-L training step (BASELINE.json
configs[1]: DRFNet-L F=64 G=6, x4, batch of 32 cropped 2D+t patches LR 32x32, T=5 frames, bf16).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arm

A step = forward over T frames + fused L1 loss + full BPTT backward + (NCCL all-reduce) + Adam +
PSNR/SSIM of the training outputs (what acdc_vsr_trainer.py:41-55 does per batch).
Prints ONE JSON line on rank 0.  At N=1 the line also carries `infer` (BASELINE configs[2]: full-FOV DSB15-shaped
cine inference with PSNR/SSIM on the device), `fp32` (the same training step in the strict fp32 mode on CUDA cores) and
`bf16x3` (the strict mode on tensor cores).
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
# BASELINE configs[2]: full-FOV DSB15-shaped cine, 256x256 x 12 slices x 30 frames, x4
INFER = dict(slices=12, frames=30, lr=64)


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
            time.sleep(0.02)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.result()

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


# ---- the reference's CPU implementation of the step ---------------------------------------------------------
def cpu_step_fn(sample):
    """One training step of the reference on the host, all cores: returns (step fn, HR voxels per step, kind).
    kind 'reference' = the UNMODIFIED reference classes (DRFNet, PSNR, SSIM, denormalize stub-loaded from
    /root/reference by oracle/load_reference.py, stepped exactly as acdc_vsr_trainer.py:41-55 does) when the reference
    tree exists (this container); 'port' = oracle/restated.py (the GPU box has no /root/reference).  Nothing of the
    product package is imported here."""
    from oracle import load_reference, restated
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
        return step, hr_voxels(sample), "port"
    if load_reference.available():
        ns = load_reference.load()
        net = ns.DRFNet(**MODEL)
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)
        loss_fn, psnr, ssim = torch.nn.L1Loss(), ns.PSNR(), ns.SSIM()

        def step():
            outs = net(lrs)                                                                       # :41
            loss = torch.stack([loss_fn(o, t) for o, t in zip(outs, hrs)]).mean()                 # :42-43, :86
            opt.zero_grad()
            loss.backward()
            opt.step()                                                                            # :44-46
            with torch.no_grad():                                                                 # :52, :90-107
                o = [ns.denormalize(x, "acdc") for x in outs]
                t = [ns.denormalize(x, "acdc") for x in hrs]
                torch.stack([psnr(a, b) for a, b in zip(o, t)]).mean()
                torch.stack([ssim(a, b) for a, b in zip(o, t)]).mean()
            return float(loss.detach())
        return step, hr_voxels(sample), "reference"
    sd = {k: v.requires_grad_(True) for k, v in restated.drfnet_init(**MODEL).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4)

    def step():
        outs = restated.drfnet_forward(lrs, sd, 4)
        loss = restated.vsr_loss(outs, hrs, restated.l1_loss)
        opt.zero_grad()
        loss.backward()
        opt.step()
        restated.vsr_metrics([o.detach() for o in outs], hrs)
        return float(loss.detach())
    return step, hr_voxels(sample), "port"


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the step on the box's host cores, all threads, on
    the arm's own config.  The FULL batch of 32 patches is timed (one step) and, when the whole --steps/--warmup run on
    the full batch would not end within a few minutes, the timed steps use a bounded sample of 8 patches with the
    measured per-patch scaling reported next to it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("VSR_REF_BUDGET_S", "230"))
    small = int(os.environ.get("VSR_REF_SAMPLE", "8"))
    scaling, sample, step, vox, kind = {}, small, None, None, None
    done_warm = 0
    if os.environ.get("VSR_REF_FULL", "1") != "0":        # (tests switch the full-batch probe off to stay fast)
        full, vox_full, kind = cpu_step_fn(BATCH)
        t0 = time.perf_counter()
        full()                                          # also the first warm-up step
        t_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        full()
        t_full = time.perf_counter() - t0
        scaling = {"full_batch_patches": BATCH, "full_batch_ms_per_step": t_full * 1e3,
                   "full_batch_value": vox_full / t_full, "first_step_ms": t_first * 1e3}
        if (args.steps + max(args.warmup - 2, 0)) * t_full <= budget:
            step, vox, sample, done_warm = full, vox_full, BATCH, 2
        else:
            del full
    if step is None:
        step, vox, kind = cpu_step_fn(sample)
    for _ in range(max(args.warmup - done_warm, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = vox / dt
    if sample != BATCH and scaling:
        scaling.update({"sample_patches": sample, "sample_ms_per_step": dt * 1e3,
                        "ms_per_patch_full": scaling["full_batch_ms_per_step"] / BATCH, "ms_per_patch_sample": dt * 1e3 / sample})
    cores = torch.get_num_threads()
    what = f"{sample} of {BATCH} patches x T{frames_in()} per step, torch {torch.__version__} CPU fp32"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(BATCH), "sample": f"{sample} of {BATCH} patches per step",
                       "per_patch_scaling": scaling},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline(budget_s=25.0):
    """the reference's CPU step timed on this box's host cores on a bounded sample (rank 0, N=1 only)."""
    sample = 8
    step, vox, kind = cpu_step_fn(sample)
    times = []
    step()                                                    # warm-up (allocator, thread pool)
    t_start = time.perf_counter()
    while len(times) < 3 and (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    dt = sorted(times)[len(times) // 2]
    return {"value": vox / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{sample} of {BATCH} patches x T{frames_in()}, 1 warm-up + median of {len(times)} steps, "
                      f"torch {torch.__version__} CPU fp32"}


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


# ---- device timing helpers ------------------------------------------------------------------------------------
def time_steps(fn, n, barrier):
    """n calls of fn(i) between barriers: (total ms / n, median of the per-call ms) from CUDA events on the current
    stream, one event after every call"""
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    evs[0].record()
    for i in range(n):
        fn(i)
        evs[i + 1].record()
    barrier()
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n))
    return evs[0].elapsed_time(evs[n]) / n, per[n // 2], per


def kernel_pass(step_fn, ops, n_steps, head_start_cycles=40_000_000):
    """Per-kernel device times of the step: the same steps launched kernel by kernel with a CUDA-event pair around
    EVERY C-ABI call (vsr_b200.ops._TimedLib).  A spin kernel in front of every step keeps the device busy while
    the host enqueues, so an event pair spans the kernel(s) of that call only - never the host's launch gaps (round
    1's numbers included them and summed to more than the step)."""
    torch.cuda.synchronize()
    ops.start_timing()
    l0 = ops.launches
    for i in range(n_steps):
        torch.cuda._sleep(head_start_cycles)
        step_fn(i)
    launches = (ops.launches - l0) // n_steps
    torch.cuda.synchronize()
    return ops.stop_timing(), launches


def infer_bench(dev, precision, iters=5):
    """BASELINE configs[2]: DRFNet-L x4 on a full-FOV DSB15-shaped cine, all 30 frames of all 12 slices in one no-grad
    call (slices are independent: tile = slice), L1 + PSNR + SSIM of every frame on the device (the predictor's loop
    body, acdc_vsr_predictor.py:53-66): device-resident and end-to-end (pinned host frames in, metrics out)."""
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.nets import DRFNet
    from vsr_b200.runner import VSRTrainStep
    n, t, h, r = INFER["slices"], INFER["frames"], INFER["lr"], 4
    torch.manual_seed(0)
    net = DRFNet(precision=precision, **MODEL).to(dev).eval()
    ev = VSRTrainStep(net, [torch.nn.L1Loss()], [1.0], [PSNR().to(dev), SSIM().to(dev)], None, "dsb15")
    g = torch.Generator().manual_seed(5)
    hx = torch.randn(t, n, 1, h, h, generator=g).pin_memory()
    hy = torch.randn(t, n, 1, h * r, h * r, generator=g).pin_memory()
    dx, dy = hx.to(dev), hy.to(dev)
    res_host = torch.zeros(3).pin_memory()
    sync = lambda: torch.cuda.synchronize()

    def resident(i):
        ev.eval_frames(list(dx.unbind(0)), list(dy.unbind(0)))

    def e2e(i):
        dx.copy_(hx, non_blocking=True)
        dy.copy_(hy, non_blocking=True)
        _, losses, metrics = ev.eval_frames(list(dx.unbind(0)), list(dy.unbind(0)))
        res_host.copy_(torch.cat([losses.mean().view(1), metrics.mean(dim=(1, 2))]), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def forward_only(i):
        with torch.no_grad():
            net(list(dx.unbind(0)))

    out = {}
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    for name, fn in (("forward", forward_only), ("forward+l1+psnr+ssim", resident), ("e2e", e2e)):
        for i in range(3):
            fn(i)
        ms, med, _ = time_steps(fn, iters, sync)
        out[name] = {"ms": ms, "ms_median": med}
    vox = n * t * (h * r) ** 2
    ms = out["forward+l1+psnr+ssim"]["ms"]
    line = {"workload": f"C3: DRFNet-L x4 inference, {n} slices x {t} frames, LR {h}x{h} -> HR {h * r}x{h * r}, "
                        f"{precision}, L1 + PSNR + SSIM per frame on the device",
            "metric": "hr_voxels_per_s_inference", "unit": UNIT, "value": vox / ms * 1e3, "ms": ms,
            "ms_forward_only": out["forward"]["ms"], "tflops_algorithmic": FWD_FLOPS_PER_LR_PIXEL * n * t * h * h / ms / 1e9,
            "e2e": {"value": vox / out["e2e"]["ms"] * 1e3, "unit": UNIT, "ms": out["e2e"]["ms"],
                    "h2d_bytes": hx.numel() * 4 + hy.numel() * 4, "d2h_bytes": 12},
            "iters": iters, "clocks": sampler.finish()}
    del net, ev
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16x3"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `infer` and `fp32` sub-lines")
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
    from vsr_b200.ops import cuda_ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops = cuda_ops()
    if args.precision == "bf16x3":
        from vsr_b200.ops import split_ops
        ops = split_ops()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    main_run = train_bench(args, args.precision, args.steps, args.warmup, world, rank, dev, ops, barrier,
                           with_e2e=True, with_kernels=True)
    line = None
    if rank == 0:
        line = build_line(args, main_run, world)
        if world == 1 and not args.no_extras and WORKLOAD == "drf":
            try:
                line["infer"] = infer_bench(dev, args.precision)
            except Exception as e:      # a reporting extra must never cost the bench line
                line["infer"] = {"error": f"{type(e).__name__}: {e}"}
            if args.precision == "bf16":
                try:
                    r32 = train_bench(args, "fp32", 3, 3, 1, 0, dev, ops, barrier, with_e2e=False, with_kernels=False)
                    line["fp32"] = {"workload": workload_name(args.batch) + ", precision='fp32' (strict mode: <= 1e-4 of the reference)",
                                    "metric": METRIC, "unit": UNIT, "value": r32["value"], "ms_per_step": r32["ms"],
                                    "ms_per_step_median": r32["ms_median"], "steps": 3, "warmup": 3,
                                    "cuda_graph": r32["cuda_graph"], "gpu_launches_per_step": r32["launches_per_step"],
                                    "step_tflops_algorithmic": r32["step_flops"] / (r32["ms"] * 1e-3) / 1e12,
                                    "clocks": r32["clocks"]}
                except Exception as e:
                    line["fp32"] = {"error": f"{type(e).__name__}: {e}"}
                try:
                    # the same strict bar on the tensor cores: fp32 maps, three bf16 tcgen05 products per product
                    from vsr_b200.ops import split_ops
                    rx3 = train_bench(args, "bf16x3", 5, 3, 1, 0, dev, split_ops(), barrier, with_e2e=False, with_kernels=False)
                    line["bf16x3"] = {"workload": workload_name(args.batch) + ", precision='bf16x3' (strict mode on tensor cores: "
                                      "<= 1e-4 of the reference; fp32 maps, xh*wh + xl*wh + xh*wl on tcgen05)",
                                      "metric": METRIC, "unit": UNIT, "value": rx3["value"], "ms_per_step": rx3["ms"],
                                      "ms_per_step_median": rx3["ms_median"], "steps": 5, "warmup": 3,
                                      "cuda_graph": rx3["cuda_graph"], "gpu_launches_per_step": rx3["launches_per_step"],
                                      "step_tflops_algorithmic": rx3["step_flops"] / (rx3["ms"] * 1e-3) / 1e12,
                                      "speedup_vs_fp32_mode": (line["fp32"]["ms_per_step"] / rx3["ms"]
                                                               if "ms_per_step" in line.get("fp32", {}) else None),
                                      "clocks": rx3["clocks"]}
                except Exception as e:
                    line["bf16x3"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            try:
                line["psnr_delta"] = psnr_delta(args.precision, dev)
            except Exception as e:
                line["psnr_delta"] = {"error": f"{type(e).__name__}: {e}"}
        emit(line)
    # orderly teardown: graphs first (they hold NCCL work), then the process group; the interpreter exits normally so
    # that at-exit hooks (the driver's record of the loaded native libraries) run
    del main_run
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    sys.stdout.flush()
    sys.stderr.flush()


def train_bench(args, precision, steps, warmup, world, rank, dev, ops, barrier, with_e2e, with_kernels):
    """the training-step benchmark on this rank: device-resident timing, per-kernel pass, end-to-end timing"""
    import torch.distributed as dist
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.nets import DRFNet
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import VSRTrainStep

    torch.manual_seed(0)                               # identical initial weights on every rank
    if WORKLOAD == "duf":
        from vsr_b200.duf import DUFNet
        from vsr_b200.runner import MISRTrainStep as StepCls
        net = DUFNet(precision=precision, **DUF_MODEL).to(dev).train()
    else:
        StepCls = VSRTrainStep
        net = DRFNet(precision=precision, **MODEL).to(dev)
    opt = FlatAdam(net.parameters(), lr=1e-4)
    step = StepCls(net, [torch.nn.L1Loss()], [1.0], [PSNR().to(dev), SSIM().to(dev)], opt, "acdc",
                   use_graph=not args.no_graph)

    n_host = 4
    host = make_batches(n_host, args.batch, seed=1234 + rank, pinned=True)   # weak scaling: own shard per rank
    dev_batches = [([x.to(dev) for x in l], [y.to(dev) for y in h]) for l, h in host]
    acc = torch.zeros(4, device=dev)

    # ---- device-resident timing -------------------------------------------------------------
    for i in range(warmup):
        step.train_step(*dev_batches[i % n_host], acc)
    barrier()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = ops.launches
    ms, ms_median, per_step = time_steps(lambda i: step.train_step(*dev_batches[i % n_host], acc), steps, barrier)
    launches = ops.launches - l0                       # 0 when the step is replayed as a CUDA graph
    res = {"ms": ms, "ms_median": ms_median, "per_step_ms": per_step, "cuda_graph": bool(step.use_graph),
           "precision": precision, "steps": steps, "warmup": warmup}
    # ---- per-kernel pass (outside the timed region) --------------------------------------------
    prof_steps = min(steps, 3)
    use_graph, step.use_graph = step.use_graph, False
    if with_kernels:
        res["records"], res["launches_per_step"] = kernel_pass(
            lambda i: step.train_step(*dev_batches[i % n_host], acc), ops, prof_steps)
    else:
        l1 = ops.launches
        step.train_step(*dev_batches[0], acc)
        res["launches_per_step"] = ops.launches - l1
    res["prof_steps"] = prof_steps
    if launches == 0:                                  # graph replay: same kernels, launched by the graph
        launches = res["launches_per_step"] * steps
    res["launches"] = launches
    step.use_graph = use_graph
    barrier()

    # ---- end to end: pinned host inputs in, loss value out, every step ------------------------
    # The trainer's own loader path (vsr_b200.data.DeviceStager, what VSRTrainer wraps its loaders in): the batch of step
    # k + 1 is copied from pinned host memory on a copy stream while step k computes; every step reads its loss back.
    # K timed steps issue exactly K batch copies inside the timed region (the copy of the first timed batch was issued by
    # the step before it, the last timed step issues one more; the closing barrier waits for it).
    if with_e2e:
        from vsr_b200.data import DeviceStager
        loss_host = torch.zeros(1).pin_memory()
        res["h2d"] = sum(x.numel() * 4 for x in host[0][0]) + sum(y.numel() * 4 for y in host[0][1])
        last = [0.0]

        def host_batches():
            i = 0
            while True:
                yield {"lr": host[i % n_host][0], "hr": host[i % n_host][1]}
                i += 1

        staged = iter(DeviceStager(host_batches(), dev))
        cur = [next(staged)]

        def e2e_step(i):
            b = cur[0]
            lv, _ = step.train_step(b["lr"], b["hr"], acc)
            loss_host.copy_(lv[:1], non_blocking=True)
            cur[0] = next(staged)                          # issues the copy of the batch after next; the device is busy with this step
            torch.cuda.current_stream().synchronize()      # the user reads the loss value
            last[0] = float(loss_host[0])

        for i in range(2):
            e2e_step(i)
        res["ms_e2e"], res["ms_e2e_median"], _ = time_steps(e2e_step, steps, barrier)
        res["last_loss"] = last[0]
    res["clocks"] = sampler.finish()

    t = torch.tensor([res["ms"], res.get("ms_e2e", 0.0), res["ms_median"], res.get("ms_e2e_median", 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["ms"], res["ms_e2e"], res["ms_median"], res["ms_e2e_median"] = t.tolist()
    vox = world * hr_voxels(args.batch)
    res["value"] = vox / (res["ms"] * 1e-3)
    res["value_e2e"] = vox / (res["ms_e2e"] * 1e-3) if with_e2e else None
    res["step_flops"] = (3.0 * duf_fwd_flops_per_sample() * args.batch if WORKLOAD == "duf"
                         else 3.0 * FWD_FLOPS_PER_LR_PIXEL * args.batch * T * LR * LR)
    del step, net, opt
    torch.cuda.empty_cache()
    return res


FAMILY = {"vsr_tapgemm": "tapgemm", "vsr_tapgemm_wgrad": "wgrad", "vsr_tapgemm_wgrad_bias": "wgrad",
          "vsr_tapgemm_wgrad_partial": "wgrad", "vsr_tapgemm_wgrad_finish": "wgrad"}
KERNEL_OF = {"tapgemm": "tapgemm_tc2_kernel", "wgrad": "wgrad_tc_kernel"}


def build_line(args, r, world):
    pk = peaks()
    ms = r["ms"]
    prof_steps = r["prof_steps"]
    # every C-ABI call of the profiled steps: family = tap-GEMM / weight gradient / the entry point's own name
    fam, detail = {}, {}
    recs = [(name, e0.elapsed_time(e1), meta) for name, e0, e1, meta, rc in r["records"]
            if not (name == "vsr_tapgemm_wgrad_partial" and rc != 1)]
    # Consistency with the timed (graph-replayed) step.  An un-graphed, event-bracketed call pays a launch latency that the
    # graph does not: the event record keeps the next kernel from being launched early (programmatic dependent launch
    # hides its prologue - barrier init, TMEM allocation, table staging - behind the predecessor's tail) and the records
    # themselves take time.  That cost is per CALL, not per byte, so when the sum of the event times exceeds the step the
    # same constant c = (sum - step) / calls is taken off every call (never more than 80 % of a call): the table then sums
    # to at most the measured step.
    raw_sum = sum(t for _, t, _ in recs) / prof_steps
    n_calls = max(len(recs), 1)
    c_call = max(0.0, (raw_sum - ms) * prof_steps / n_calls)
    for name, t_raw, meta in recs:
        t_ms = max(t_raw - c_call, 0.2 * t_raw)
        f = fam.setdefault(FAMILY.get(name, name[4:]), [0.0, 0.0, 0])
        f[1] += t_ms
        f[2] += 1
        if meta is not None:
            kind, flops, sig, nbytes = meta
            f[0] += flops
            q = detail.setdefault(f"{kind}:{sig}", [0.0, 0.0, 0, 0.0])
            q[0] += flops
            q[1] += t_ms
            q[2] += 1
            q[3] += nbytes
    sum_ms = sum(v[1] for v in fam.values()) / prof_steps
    scale = min(1.0, ms / sum_ms) if sum_ms > 0 else 1.0        # (only the 80 % clamp can leave the sum above the step)
    dom = max((k for k in fam if k in KERNEL_OF), key=lambda k: fam[k][1])
    traffic = None
    for fn in (("r02_duf_by_kernel.json", "r01_duf_by_kernel_v3.json") if WORKLOAD == "duf"
               else ("r02_dram_by_kernel.json", "r01_dram_by_kernel_v9.json")):
        try:       # per-launch dram__bytes_read+write of the same step under ncu (tools/ncu_summary.py)
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                traffic = json.load(f).get(KERNEL_OF[dom], {}).get("dram_bytes_per_launch")
            break
        except (OSError, ValueError):
            pass
    flops, kms, cnt = fam[dom]
    kms *= scale
    achieved = flops / (kms * 1e-3) / 1e12
    peak = pk["bf16_tflops_sustained"]
    fam_bytes = sum(v[3] for k, v in detail.items() if k.startswith(dom))
    gbs_fam = fam_bytes / (kms * 1e-3) / 1e9
    intensity = flops / max(fam_bytes, 1.0)
    ridge = pk["bf16_tflops_sustained"] * 1e3 / pk["hbm_gbs"]
    kernel_share = {k: {"ms_per_step": v[1] * scale / prof_steps, "calls_per_step": v[2] / prof_steps,
                        **({"tflops": v[0] / (v[1] * scale * 1e-3) / 1e12} if v[0] else {})}
                    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])}

    def shape_row(v):
        # every shape against ITS OWN bound: tensor if its algorithmic intensity is above the ridge of the
        # measured peaks (FLOP/B), else HBM
        t = v[1] * scale
        tf, gb = v[0] / (t * 1e-3) / 1e12, v[3] / (t * 1e-3) / 1e9
        tensor = v[0] / max(v[3], 1.0) > pk["bf16_tflops_sustained"] * 1e3 / pk["hbm_gbs"]
        return {"ms_per_step": t / prof_steps, "n_per_step": v[2] / prof_steps, "us_per_launch": 1e3 * t / v[2],
                "tflops": tf, "gbs": gb, "bound": "tensor" if tensor else "hbm",
                "frac": tf / pk["bf16_tflops_sustained"] if tensor else gb / pk["hbm_gbs"]}

    top = sorted(detail.items(), key=lambda kv: -kv[1][1])[:14]
    kernel_detail = {k: shape_row(v) for k, v in top}
    all_rows = [shape_row(v) for k, v in detail.items() if k.startswith(dom)]
    t_all = sum(x["ms_per_step"] for x in all_rows)
    # the dominant kernel runs tensor-bound and HBM-bound shapes: time-weighted mean of every shape's
    # fraction of ITS OWN bound, and the split of its time between the two bounds
    frac_own = sum(x["ms_per_step"] * x["frac"] for x in all_rows) / max(t_all, 1e-9)
    hbm_share = sum(x["ms_per_step"] for x in all_rows if x["bound"] == "hbm") / max(t_all, 1e-9)
    tensor_rows = [x for x in all_rows if x["bound"] == "tensor"]
    t_tensor = sum(x["ms_per_step"] for x in tensor_rows)
    frac_tensor = sum(x["ms_per_step"] * x["frac"] for x in tensor_rows) / max(t_tensor, 1e-9)
    try:      # full per-shape table for the profile notes (scratch; the JSON line keeps the top 14)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "kernel_detail_full.json"), "w") as f:
            json.dump({"kernels": kernel_share, "shapes": {k: shape_row(v) for k, v in sorted(detail.items(), key=lambda kv: -kv[1][1])},
                       "ms_per_step": ms, "kernel_ms_sum_events": raw_sum, "per_call_latency_removed_us": c_call * 1e3,
                       "scale": scale, "clocks": r["clocks"]}, f, indent=1)
    except OSError:
        pass
    return {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": r["steps"],
        "warmup": r["warmup"], "ms_per_step": ms, "ms_per_step_median": r["ms_median"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if r["precision"] == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "per_gpu_batch": args.batch, "frames": frames_in(),
                   "parallelism": f"dp{world}", "cuda_graph": r["cuda_graph"],
                   "collective": ("none" if world == 1 else "NCCL all-reduce of 3 ranges of the flat fp32 gradient bucket, on a "
                                  "communication stream, captured inside the step's CUDA graph"),
                   "l2": ("per-step working set (~5 GB of activations) exceeds the 126 MB L2; inputs rotate over 4 batches" if WORKLOAD == "drf"
                          else "per-step working set (~1.5 GB of activations and gradients) exceeds the 126 MB L2; inputs rotate over 4 batches")},
        "clocks": r["clocks"],
        "e2e": {"value": r["value_e2e"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                "ms_per_step": r["ms_e2e"], "ms_per_step_median": r["ms_e2e_median"], "last_loss": r["last_loss"],
                "path": "DeviceStager (the trainer's loader wrapper): pinned-host batch of step k+1 copied on a copy stream "
                        "while step k computes; loss read back every step"},
        "gpu_launches": r["launches"],
        # The roofline of the dominant kernel as a whole: its launches move `fam_bytes` algorithmic bytes for `flops`
        # algorithmic FLOPs; below the ridge of the measured peaks (FLOP/B) the HBM roof is the binding one, above it the
        # tensor roof.  Both views are always reported.
        "roofline": {**({"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak}
                        if intensity > ridge else
                        {"bound": "hbm", "achieved": gbs_fam, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs_fam / pk["hbm_gbs"]}),
                     "kernel": dom, "traffic": traffic,
                     **({"note": "FLOPs and bytes are those of the channel-padded tap tables (64 + 32 i channels padded to multiples "
                                 "of 64 with structural-zero weights, the (1,3,3) column form computes 128 columns for 96): about "
                                 "3/4 of them are useful work (DESIGN.md section 3b)"} if WORKLOAD == "duf" else {}),
                     "algorithmic_bytes_per_launch": fam_bytes / max(cnt, 1),
                     "algorithmic_flops_per_launch": flops / max(cnt, 1),
                     "intensity_flop_per_byte": intensity, "ridge_flop_per_byte": ridge,
                     "tensor_view": {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak},
                     "hbm_view": {"achieved": gbs_fam, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs_fam / pk["hbm_gbs"]},
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained / hbm_gbs ({pk['source']})",
                     "frac_time_weighted_own_bound": frac_own, "time_share_hbm_bound_shapes": hbm_share,
                     "frac_tensor_bound_shapes": frac_tensor,
                     "launches": cnt, "kernels": kernel_share, "kernel_detail": kernel_detail,
                     "kernel_time_method": (f"{prof_steps} extra un-graphed steps of the same workload, a CUDA-event pair around every C-ABI "
                                            "call, a spin kernel in front of every step so that the host stays ahead (no launch gaps "
                                            "inside a pair); a constant per-call launch latency c = (sum - step) / calls is taken off "
                                            "every call when the sum exceeds the graph-replayed step, so that the table sums to at "
                                            "most that step"),
                     "kernel_ms_sum_events": raw_sum, "per_call_latency_removed_us": c_call * 1e3,
                     "calls_per_step": n_calls / prof_steps, "kernel_time_scale": scale,
                     "kernel_ms_sum": sum_ms * scale,
                     "step_tflops_algorithmic": r["step_flops"] / (ms * 1e-3) / 1e12 * world},
    }


if __name__ == "__main__":
    main()
