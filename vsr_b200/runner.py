"""Data-parallel VSR trainer (reference: src/runner/trainers/base_trainer.py:46-144,
acdc_vsr_trainer.py:16-123).

Same constructor arguments and epoch loop contract as the reference's AcdcVSRTrainer (so the
reference's Monitor / loggers keep working); the per-step body is re-designed for B200:

  H2D (pinned, side stream)  ->  engine forward (T frames)  ->  fused loss fwd+bwd kernels (they
  write dL/dout directly)  ->  engine backward into ONE flat gradient bucket  ->  one NCCL
  all-reduce of that bucket over NVLink (world_size > 1)  ->  one fused Adam kernel  ->  fused
  denormalize+PSNR/SSIM kernels;  loss / metric sums stay on the device and are read back once per
  epoch instead of 1 + #loss + #metric `.item()` syncs per step (acdc_vsr_trainer.py:119-123).
"""
import logging
import os
import random

import numpy as np
import torch
import torch.distributed as dist

from .drf_engine import _as_stacked, _stacked_view
from .losses import LOSS_KINDS
from .optim import FlatAdam
from .utils import DATASET_STATS


def _loss_kind(fn):
    name = fn.__class__.__name__
    if name not in LOSS_KINDS:
        raise NotImplementedError(f"loss {name} has no fused kernel (supported: {sorted(LOSS_KINDS)})")
    param = float(getattr(fn, "epsilon", getattr(fn, "delta", 0.0)))
    return LOSS_KINDS[name], param


class VSRTrainStep:
    """The fused training / evaluation step on one rank."""

    def __init__(self, net, loss_fns, loss_weights, metric_fns, optimizer, dataset="acdc", process_group=None,
                 use_graph=False):
        self.net, self.optimizer = net, optimizer
        # CUDA-graph the whole step (needs FlatAdam: its hyper-parameters live on the device);
        # the first two calls per input shape run eagerly, the third captures, later ones replay.
        self.use_graph = use_graph and isinstance(optimizer, FlatAdam)
        self._graphs, self._calls = {}, {}
        self.losses = [_loss_kind(f) for f in loss_fns]
        self.loss_names = [f.__class__.__name__ for f in loss_fns]
        self.loss_weights = [float(w) for w in loss_weights]
        self.metric_names = [m.__class__.__name__ for m in metric_fns]
        for m in self.metric_names:
            if m not in ("PSNR", "SSIM", "CardiacPSNR", "CardiacSSIM"):
                raise NotImplementedError(f"metric {m} has no fused kernel (supported: PSNR, SSIM, CardiacPSNR, CardiacSSIM)")
        self.metric_fns = metric_fns
        self.mean, self.std = DATASET_STATS[dataset]
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self._comm_stream, self._comm_pending = None, False
        # "overlap": range k is all-reduced on a side stream while the weight gradients of range k + 1 run;
        # "serial": ONE all-reduce of the whole bucket on the compute stream after backward.  (NCCL's CTAs take SMs
        # away from the persistent one-CTA-per-SM weight-gradient kernels running beside them, whose displaced CTAs
        # then run as a second wave - the overlap can cost as much as the exchange it hides.)
        # Measured equal on 2 GPUs (10.97 ms per step both ways, profiles/r02c_2gpu_{overlap,serial}.json).
        self.comm_mode = os.environ.get("VSR_COMM_MODE", "overlap")
        if self.comm_mode not in ("overlap", "serial"):
            raise ValueError(f"VSR_COMM_MODE must be 'overlap' or 'serial', not {self.comm_mode!r}")
        if isinstance(optimizer, FlatAdam):
            optimizer.bind(net)
            optimizer.grad_scale = 1.0 / self.world
        self._bufs = {}

    def _buf(self, key, shape, dtype=torch.float32):
        b = self._bufs.get(key)
        if b is None or b.shape != torch.Size(shape):
            b = torch.empty(shape, dtype=dtype, device=self.net.flat.device)
            self._bufs[key] = b
        return b

    def _ops(self):
        return self.net._backend()

    def _metric_values(self, outs, targets):
        """[n_metric, T, N*C] per-frame, per-image metric values (PSNR fills [:, :, :N]): ONE launch pair per metric
        over all T frames (they are consecutive slices of one buffer, or stacked into one here)."""
        ops = self._ops()
        T, n, c = len(outs), outs[0].shape[0], outs[0].shape[1]
        o_all, y_all = _as_stacked(outs), _as_stacked(targets)          # [T, N, C, H, W]
        per = outs[0].numel() // n
        ws = self._buf("mws", (ops.metric_workspace(T * n * c, per) // 4 + 4,))
        vals = self._buf("mvals", (len(self.metric_names), T, n * c))
        for i, (name, fn) in enumerate(zip(self.metric_names, self.metric_fns)):
            if name == "PSNR":
                flat = self._buf("mpsnr", (T * n,))
                ops.psnr(o_all.view(T * n, -1), y_all.view(T * n, -1), self.mean, self.std, float(fn.max_value), flat, ws)
                vals[i, :, :n].copy_(flat.view(T, n))
            else:
                hw = outs[0].shape[2:]
                ops.ssim(o_all.view(T * n * c, *hw), y_all.view(T * n * c, *hw), fn.window, self.mean, self.std,
                         fn.c1, fn.c2, vals[i].view(-1), ws)
        return vals

    def _metrics(self, outs, targets, acc):
        """acc[1 + n_loss + i] += mean over frames of metric i (fused denormalize)."""
        for name in self.metric_names:
            if name.startswith("Cardiac"):
                raise NotImplementedError("Cardiac* metrics need the patient of every sample: predictors only "
                                          "(acdc_vsr_predictor.py:134-154); the reference's trainers do not use them")
        n = outs[0].shape[0]
        vals = self._metric_values(outs, targets)
        k = 1 + len(self.losses)
        for i, name in enumerate(self.metric_names):
            cnt = vals.shape[2] if name == "SSIM" else n
            acc[k + i] += vals[i, :, :cnt].mean()

    def _loss_partials(self, outs, targets, want_grad):
        """fused losses, ONE launch per loss over all T frames: (partials [L*T, partials_len], dL/dout frames)"""
        ops = self._ops()
        T, L = len(outs), len(self.losses)
        o_all, y_all = _as_stacked(outs), _as_stacked(targets)
        numel = outs[0].numel()
        partials = self._buf("lpart", (L * T, ops.partials_len))
        partials.zero_()
        g_all = torch.empty_like(o_all) if want_grad else None
        for li, (kind, param) in enumerate(self.losses):
            g = g_all
            if want_grad and li > 0:
                g = self._buf("gtmp", o_all.shape)
            ops.loss_fwd_bwd_seg(o_all, y_all, T, kind, param, self.loss_weights[li] / (numel * T),
                                 partials[li * T:(li + 1) * T], g)
            if want_grad and li > 0:
                ops.add(g_all, g, g_all)
        return partials, (list(g_all.unbind(0)) if want_grad else [None] * T)

    def _loss(self, outs, targets, want_grad):
        """returns ([L] loss values on device: mean over frames of the per-frame means, list of dL/dout)."""
        ops = self._ops()
        T, L = len(outs), len(self.losses)
        partials, grads = self._loss_partials(outs, targets, want_grad)
        sums = self._buf("lsum", (L,))
        sums.zero_()
        rd = self._bufs.get("lrd")
        if rd is None or rd.numel() != L * T:
            rd = torch.tensor([li for li in range(L) for _ in range(T)], dtype=torch.int32,
                              device=self.net.flat.device)
            self._bufs["lrd"] = rd
        ops.reduce_partials(partials, L * T, rd, sums)
        return sums / (outs[0].numel() * T), grads     # all frames share a shape

    def _device_fwd_bwd(self, inputs, targets):
        """pack + forward + fused loss + backward: everything before the gradient exchange."""
        net = self.net
        eng = self._engine()
        eng.pack(net.flat, need_bwd=True)
        outs, saved = eng.forward(inputs, save=True)
        lvals, grads = self._loss(outs, targets, True)
        overlap = self.world > 1 and self.comm_mode == "overlap"
        gflat = eng.backward(saved, grads, on_bucket=self._reduce_bucket if overlap else None)
        if self.world > 1 and not overlap:
            dist.all_reduce(gflat, group=self.pg)
        net.flat_grad = gflat
        return lvals, outs, gflat

    def _reduce_bucket(self, gflat, lo, hi):
        """all-reduce (sum; the 1/world mean is folded into Adam) of a finished range of the flat gradient bucket on
        a side stream, so that NCCL over NVLink runs while the weight gradients of the next range are computed; the
        fork / join are stream dependencies, hence captured with the rest of the step in its CUDA graph."""
        if not gflat.is_cuda:
            dist.all_reduce(gflat[lo:hi], group=self.pg)
            return
        cur = torch.cuda.current_stream(gflat.device)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=gflat.device)
        self._comm_stream.wait_stream(cur)
        with torch.cuda.stream(self._comm_stream):
            dist.all_reduce(gflat[lo:hi], group=self.pg)
        self._comm_pending = True

    def _join_comm(self, gflat):
        if self._comm_pending:
            torch.cuda.current_stream(gflat.device).wait_stream(self._comm_stream)
            self._comm_pending = False

    def _device_update(self, lvals, outs, targets, gflat, acc, with_metrics):
        """(join the gradient all-reduces) + fused Adam + logging / metrics."""
        self._join_comm(gflat)
        if isinstance(self.optimizer, FlatAdam):
            self.optimizer.launch(gflat)
        if acc is not None:
            self._log(acc, lvals)
            if with_metrics and self.metric_names:
                self._metrics(outs, targets, acc)

    def _device_step(self, inputs, targets, acc, with_metrics):
        """everything of a step that runs on the device."""
        lvals, outs, gflat = self._device_fwd_bwd(inputs, targets)
        self._device_update(lvals, outs, targets, gflat, acc, with_metrics)
        return lvals, outs, gflat

    def train_step(self, inputs, targets, acc=None, with_metrics=True):
        """One optimisation step. `acc` ([1 + n_loss + n_metric] device tensor) accumulates
        Loss, each loss and each metric for logging. Returns the loss values (device tensor)."""
        net = self.net
        if not net._is_flat():
            net._flatten()
        inputs = [x.contiguous() for x in inputs]
        self._loss_weights_dev(inputs[0].device)
        flat_adam = isinstance(self.optimizer, FlatAdam)
        if flat_adam:
            self.optimizer.prepare_step()
        if self.use_graph and net.flat.is_cuda:
            return self._graphed_step(inputs, targets, acc, with_metrics)
        targets = list(_as_stacked(targets).unbind(0))      # one buffer: loss / metric kernels take all frames at once
        lvals, outs, gflat = self._device_step(inputs, targets, acc, with_metrics)
        if not flat_adam:
            if self.world > 1:
                gflat.mul_(1.0 / self.world)
            for p, ref in zip(net.parameters(), net._plan.params.values()):
                p.grad = gflat[ref.offset:ref.offset + p.numel()].view(ref.shape)
            self.optimizer.step()
        return lvals, outs

    def _graphed_step(self, inputs, targets, acc, with_metrics):
        key = (len(inputs), tuple(inputs[0].shape), tuple(targets[0].shape), acc is not None, with_metrics)
        n = self._calls.get(key, 0)
        self._calls[key] = n + 1
        if n < 2:                                   # eager: fills workspaces / descriptor caches
            lvals, outs, _ = self._device_step(inputs, targets, acc, with_metrics)
            return lvals, outs
        g = self._graphs.get(key)
        if g is None:
            dev = self.net.flat.device
            T = len(inputs)
            # static inputs / targets as ONE [T, N, ...] buffer each: the loss and metric kernels take all frames
            # in one launch
            tin = torch.empty(T, *inputs[0].shape, dtype=inputs[0].dtype, device=dev)
            ttg = torch.empty(len(targets), *targets[0].shape, dtype=targets[0].dtype, device=dev)
            st = {"in_all": tin, "tg_all": ttg, "in": list(tin.unbind(0)), "tg": list(ttg.unbind(0)),
                  "acc": torch.zeros(1 + len(self.losses) + len(self.metric_names), device=dev)}
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                # the whole step is one graph, on one rank or many: forward, fused loss, backward, the NCCL
                # all-reduces of the gradient ranges (forked onto a communication stream inside the capture),
                # Adam, metrics
                with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
                    st["lvals"], st["outs"], st["gflat"] = self._device_fwd_bwd(st["in"], st["tg"])
                    st["acc"].zero_()
                    self._device_update(st["lvals"], st["outs"], st["tg"], st["gflat"], st["acc"], with_metrics)
            torch.cuda.current_stream(dev).wait_stream(side)
            st["graph"] = graph
            self._graphs[key] = g = st
        for dst_all, dst, src in ((g["in_all"], g["in"], inputs), (g["tg_all"], g["tg"], targets)):
            v = _stacked_view(src)
            if v is not None:
                dst_all.copy_(v, non_blocking=True)           # one copy for all frames
            else:
                for d, s_ in zip(dst, src):
                    d.copy_(s_, non_blocking=True)
        g["graph"].replay()
        self.net.flat_grad = g["gflat"]
        if acc is not None:
            acc += g["acc"]
        return g["lvals"], g["outs"]

    @torch.no_grad()
    def eval_step(self, inputs, targets, acc=None):
        net = self.net
        if not net._is_flat():
            net._flatten()
        eng = self._engine()
        eng.pack(net.flat, need_bwd=False)
        outs, _ = eng.forward([x.contiguous() for x in inputs], save=False)
        targets = list(_as_stacked(targets).unbind(0))
        lvals, _ = self._loss(outs, targets, False)
        if acc is not None:
            self._log(acc, lvals)
            if self.metric_names:
                self._metrics(outs, targets, acc)
        return lvals, outs

    @torch.no_grad()
    def eval_frames(self, inputs, targets, patients=None):
        """The predictor's loop body (acdc_vsr_predictor.py:53-66) for a whole batch of sequences: forward under
        no_grad, then on the device per-frame losses [T, n_loss] (batch means) and per-frame, per-sample metrics
        [n_metric, T, N] with the denormalisation fused.  CardiacPSNR / CardiacSSIM (metrics.py:116-165) run the same
        kernels on every sample's bounding box (`patients`: one name per sample, :134-154).
        Returns (outputs, losses, metrics); nothing syncs."""
        net = self.net
        if not net._is_flat():
            net._flatten()
        ops = self._ops()
        outs = self._infer(inputs)
        targets = list(_as_stacked(targets).unbind(0))
        T, L, n, c = len(outs), len(self.losses), outs[0].shape[0], outs[0].shape[1]
        partials, _ = self._loss_partials(outs, targets, False)                 # one launch per loss, all frames
        losses = (partials.sum(dim=1) / outs[0].numel()).view(L, T).t().contiguous()
        cardiac = [i for i, name in enumerate(self.metric_names) if name.startswith("Cardiac")]
        plain = [i for i in range(len(self.metric_names)) if i not in cardiac]
        vals = torch.zeros(len(self.metric_names), T, n * c, device=outs[0].device)
        if plain:
            names, fns = self.metric_names, self.metric_fns
            self.metric_names, self.metric_fns = [names[i] for i in plain], [fns[i] for i in plain]
            try:
                vals[plain] = self._metric_values(outs, targets)                # one launch pair per metric
            finally:
                self.metric_names, self.metric_fns = names, fns
        if cardiac:
            if patients is None or len(patients) != n:
                raise ValueError(f"{self.metric_names[cardiac[0]]} needs the patient name of every sample of the batch")
            o_all, y_all = _as_stacked(outs), _as_stacked(targets)              # [T, N, C, H, W]
            per = outs[0].numel() // n
            ws = self._buf("mws", (ops.metric_workspace(T * n * c, per) // 4 + 4,))
            for i in cardiac:
                fn = self.metric_fns[i]
                m = fn.metric
                for s_, who in enumerate(patients):
                    # one crop (all frames of the sample) and one launch pair per sample: boxes differ per patient
                    h0, hn, w0, wn = fn.coordinates[who]
                    o = o_all[:, s_, :, h0:hn, w0:wn].contiguous()              # [T, C, bh, bw]
                    y = y_all[:, s_, :, h0:hn, w0:wn].contiguous()
                    if self.metric_names[i] == "CardiacPSNR":
                        tmp = self._buf("cpsnr", (T,))
                        ops.psnr(o.view(T, -1), y.view(T, -1), self.mean, self.std, float(m.max_value), tmp, ws)
                        vals[i, :, s_].copy_(tmp)
                    else:
                        tmp = self._buf("cssim", (T * c,))
                        ops.ssim(o.view(T * c, *o.shape[2:]), y.view(T * c, *y.shape[2:]), m.window, self.mean, self.std,
                                 m.c1, m.c2, tmp, ws)
                        vals[i, :, s_ * c:(s_ + 1) * c].copy_(tmp.view(T, c))
        metrics = torch.stack([vals[i, :, :n] if name.endswith("PSNR") else vals[i].view(T, n, c).mean(dim=2)
                               for i, name in enumerate(self.metric_names)]) if self.metric_names else vals
        return outs, losses, metrics

    def _infer(self, inputs):
        """forward without saving activations -> list of output frames"""
        eng = self._engine()
        eng.pack(self.net.flat, need_bwd=False)
        outs, _ = eng.forward([x.contiguous() for x in inputs], save=False)
        return outs

    def _loss_weights_dev(self, device):
        """the loss weights on the device (made outside any graph capture: a host -> device copy is not capturable)"""
        w = self._bufs.get("lw")
        if w is None or w.device != device:
            w = torch.tensor(self.loss_weights, device=device)
            self._bufs["lw"] = w
        return w

    def _log(self, acc, lvals):
        w = self._loss_weights_dev(lvals.device)
        # the kernels already applied the weights to the gradients; the logged values follow the
        # reference: Loss = sum_i w_i * loss_i, then each unweighted loss_i (acdc_vsr_trainer.py:43)
        acc[0] += (lvals * w).sum()
        acc[1:1 + lvals.numel()] += lvals

    def _engine(self):
        net = self.net
        if net._engine is None or net._engine.ops is not net._backend():
            from .drf_engine import DrfEngine
            from .nets import _PRECISIONS
            net._engine = DrfEngine(net._plan, net._backend(), net.flat.device, _PRECISIONS[net.precision],
                                    net.flat.dtype)
        return net._engine


class MISRTrainStep(VSRTrainStep):
    """The same fused step for the multi-image SR nets (reference: acdc_misr_trainer.py:8-50 — inputs are the
    `num_frames` LR frames, the target ONE HR frame): DUFNet forward, fused loss, DUFNet backward into the flat
    bucket, all-reduce, fused Adam, fused PSNR / SSIM; CUDA-graphed like the VSR step.  `targets` is a one-element
    list.  Data parallel: `sync_bn=True` (default) synchronises every BatchNorm over the ranks (DUFNet.enable_sync_bn),
    so G ranks with batch B reproduce one device with batch G*B; `sync_bn=False` = rank-local statistics (torch DDP's
    default), which keeps the step inside one CUDA graph."""

    def __init__(self, net, loss_fns, loss_weights, metric_fns, optimizer, dataset="acdc", process_group=None,
                 use_graph=False, sync_bn=True):
        super().__init__(net, loss_fns, loss_weights, metric_fns, optimizer, dataset, process_group, use_graph)
        if self.world > 1 and sync_bn:
            net.enable_sync_bn(process_group)
            # the BatchNorm all-reduces sit inside forward / backward and are captured with them in the step's CUDA
            # graph (NCCL is capturable; measured 7.1 ms eager -> 5.5 ms graphed per step on 2 GPUs);
            # VSR_SYNCBN_GRAPH=0 launches them eagerly instead
            if os.environ.get("VSR_SYNCBN_GRAPH", "1") == "0":
                self.use_graph = False

    def _device_fwd_bwd(self, inputs, targets):
        net = self.net
        net._pack(True)
        y, saved = net._forward(inputs, True)
        lvals, grads = self._loss([y], targets, True)
        gflat = net._backward(saved, grads[0])
        if self.world > 1:
            self._reduce_bucket(gflat, 0, gflat.numel())
        net.flat_grad = gflat
        return lvals, [y], gflat

    def _infer(self, inputs):
        self.net._pack(False)
        return [self.net._forward([x.contiguous() for x in inputs], False)[0]]

    @torch.no_grad()
    def eval_step(self, inputs, targets, acc=None):
        net = self.net
        if not net._is_flat():
            net._flatten()
        y = self._infer(inputs)[0]
        targets = [t.contiguous() for t in targets]
        lvals, _ = self._loss([y], targets, False)
        if acc is not None:
            self._log(acc, lvals)
            if self.metric_names:
                self._metrics([y], targets, acc)
        return lvals, [y]


class SISRTrainStep(MISRTrainStep):
    """The fused step for the single-image nets with ONE output (reference: acdc_sisr_trainer.py:8-48 on
    base_trainer.py:99-144; EDSRNet): `inputs` / `targets` are one-element lists [lr_img] / [hr_img]."""

    def __init__(self, net, loss_fns, loss_weights, metric_fns, optimizer, dataset="acdc", process_group=None,
                 use_graph=False):
        super().__init__(net, loss_fns, loss_weights, metric_fns, optimizer, dataset, process_group, use_graph,
                         sync_bn=False)

    def _device_fwd_bwd(self, inputs, targets):
        net = self.net
        net._pack(True)
        y, saved = net._forward(inputs[0], True)
        lvals, grads = self._loss([y], targets, True)
        gflat = net._backward(saved, grads[0])
        if self.world > 1:
            self._reduce_bucket(gflat, 0, gflat.numel())
        net.flat_grad = gflat
        return lvals, [y], gflat

    def _infer(self, inputs):
        self.net._pack(False)
        return [self.net._forward(inputs[0].contiguous(), False)[0]]


class SISRSRFBTrainStep(VSRTrainStep):
    """The fused step for the single-image nets that return one output per feedback step (reference:
    acdc_sisr_srfb_trainer.py:8-38; SRFBNet, DRFSISRNet): every loss is the mean over the steps against the SAME target
    (:22-24), the metrics see the last step's output only (:36).  `inputs` / `targets` are [lr_img] / [hr_img]; the
    recurrent engine runs `net.num_steps` iterations on the one image, exactly as the nets' own forward does."""

    def _expand(self, inputs, targets):
        S = self.net.num_steps
        return [inputs[0]] * S, [targets[0]] * S

    def train_step(self, inputs, targets, acc=None, with_metrics=True):
        return super().train_step(*self._expand(inputs, targets), acc, with_metrics)

    def eval_step(self, inputs, targets, acc=None):
        return super().eval_step(*self._expand(inputs, targets), acc)

    def _metrics(self, outs, targets, acc):
        super()._metrics(outs[-1:], targets[-1:], acc)


class FRVSRTrainStep(VSRTrainStep):
    """The fused step for FRVSRNet (reference: acdc_frvsr_trainer.py:8-120): the net returns (sr_imgs, lr_imgs);
    losses = [flow_loss, sr_loss] with flow_loss = mean over frames of loss_fns[0](warped previous LR frame, LR frame)
    and sr_loss = mean over frames of loss_fns[1](SR frame, HR frame) (:85-88); the metrics see the SR frames (:99-107).
    Forward with the launch record, the two fused losses (one launch each for all frames), the recorded backward,
    all-reduce, fused Adam, fused PSNR / SSIM; CUDA-graphed when the frame size is a multiple of 8 (FNet's padding reads
    the minimum of the frames on the host, frvsr_net.py:149-156)."""

    def __init__(self, net, loss_fns, loss_weights, metric_fns, optimizer, dataset="acdc", process_group=None,
                 use_graph=False):
        if len(loss_fns) != 2:
            raise ValueError("FRVSR takes two losses: [flow loss, SR loss] (acdc_frvsr_trainer.py:85-88)")
        super().__init__(net, loss_fns, loss_weights, metric_fns, optimizer, dataset, process_group, use_graph)

    def _pair_loss(self, li, outs, refs, want_grad):
        """loss li between two lists of frames: (its value [1] on the device, weight * dL/d(outs) per frame)"""
        ops = self._ops()
        T = len(outs)
        o_all, y_all = _as_stacked(outs), _as_stacked(refs)
        numel = outs[0].numel()
        partials = self._buf(f"lpart{li}", (T, ops.partials_len))
        partials.zero_()
        g_all = torch.empty_like(o_all) if want_grad else None
        kind, param = self.losses[li]
        ops.loss_fwd_bwd_seg(o_all, y_all, T, kind, param, self.loss_weights[li] / (numel * T), partials, g_all)
        total = self._buf(f"lsum{li}", (1,))
        total.zero_()
        rd = self._bufs.get(f"lrd{li}")
        if rd is None or rd.numel() != T:
            rd = torch.zeros(T, dtype=torch.int32, device=self.net.flat.device)
            self._bufs[f"lrd{li}"] = rd
        ops.reduce_partials(partials, T, rd, total)
        return total / (numel * T), (list(g_all.unbind(0)) if want_grad else None)

    def _device_fwd_bwd(self, inputs, targets):
        net = self.net
        net._pack(True)
        sr, lr, tape = net._forward(inputs, True)
        l_flow, g_lr = self._pair_loss(0, lr, inputs, True)
        l_sr, g_sr = self._pair_loss(1, sr, targets, True)
        gflat = net._backward(tape, g_sr, g_lr)
        if self.world > 1:
            self._reduce_bucket(gflat, 0, gflat.numel())
        net.flat_grad = gflat
        return torch.cat([l_flow, l_sr]), sr, gflat

    def train_step(self, inputs, targets, acc=None, with_metrics=True):
        graph = self.use_graph
        if inputs[0].shape[-2] % 8 or inputs[0].shape[-1] % 8:
            self.use_graph = False
        try:
            return super().train_step(inputs, targets, acc, with_metrics)
        finally:
            self.use_graph = graph

    @torch.no_grad()
    def eval_step(self, inputs, targets, acc=None):
        net = self.net
        if not net._is_flat():
            net._flatten()
        inputs = [x.contiguous() for x in inputs]
        self._loss_weights_dev(inputs[0].device)
        net._pack(False)
        sr, lr, _ = net._forward(inputs, False)
        targets = list(_as_stacked(targets).unbind(0))
        lvals = torch.cat([self._pair_loss(0, lr, inputs, False)[0], self._pair_loss(1, sr, targets, False)[0]])
        if acc is not None:
            self._log(acc, lvals)
            if self.metric_names:
                self._metrics(sr, targets, acc)
        return lvals, sr


class VSRTrainer:
    """Drop-in for AcdcVSRTrainer / Dsb15VSRTrainer (same constructor keywords; `dataset` selects
    the denormalisation constants, default 'acdc').  Under torchrun (an initialised process group) every rank runs
    this class on its own shard of the samples: the loaders are re-built around a rank-aware sampler that is
    deterministic under the reference's per-epoch reseed (data.ShardedSampler, base_trainer.py:54), the step is the
    CUDA-graphed fused step with the gradient all-reduce inside it (`use_graph`, `process_group`), every rank takes
    the same early-stopping decision, and rank 0 alone logs and checkpoints."""

    _step_cls = None        # set below (VSRTrainStep; MISRTrainer: MISRTrainStep)

    def __init__(self, device, train_dataloader, valid_dataloader, net, loss_fns, loss_weights, metric_fns,
                 optimizer, lr_scheduler, logger, monitor, num_epochs, dataset="acdc", use_graph=True,
                 process_group=None):
        from .data import shard_loader
        self.device = torch.device(device)
        self.rank = dist.get_rank(process_group) if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.pg = process_group
        self.train_dataloader, self._train_sampler = shard_loader(train_dataloader, self.rank, self.world)
        self.valid_dataloader, self._valid_sampler = shard_loader(valid_dataloader, self.rank, self.world)
        self.net = net.to(self.device)
        self.loss_fns, self.metric_fns = list(loss_fns), [m.to(self.device) for m in metric_fns]
        self.loss_weights = list(loss_weights)
        self.optimizer, self.lr_scheduler = optimizer, lr_scheduler
        self.logger, self.monitor, self.num_epochs = logger, monitor, num_epochs
        self.epoch, self.np_random_seeds = 1, None
        self.step = self._make_step(dataset, use_graph and self.device.type == "cuda")

    def _make_step(self, dataset, use_graph):
        return VSRTrainStep(self.net, self.loss_fns, self.loss_weights, self.metric_fns, self.optimizer, dataset,
                            process_group=self.pg, use_graph=use_graph)

    def _check_batch_counts(self, loader):
        """every rank must run the same number of steps (each step holds a collective)"""
        if self.world > 1 and hasattr(loader, "__len__"):
            n = torch.tensor([len(loader), -len(loader)], dtype=torch.int64, device=self.device)
            dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.pg)
            if int(n[0]) != -int(n[1]):
                raise RuntimeError(f"rank {self.rank}: {len(loader)} batches per epoch, but the ranks disagree "
                                   f"(max {int(n[0])}, min {-int(n[1])}): shard the loader with equal lengths "
                                   "(vsr_b200.data.ShardedSampler)")

    def _keys(self):
        return ["Loss"] + [f.__class__.__name__ for f in self.loss_fns] + [m.__class__.__name__ for m in self.metric_fns]

    def _run_epoch(self, mode):
        from .data import DeviceStager
        training = mode == "training"
        self.net.train(training)
        loader = self.train_dataloader if training else self.valid_dataloader
        self._check_batch_counts(loader)
        keys = self._keys()
        acc = torch.zeros(len(keys), device=self.device)
        count, batch, outputs = 0, None, None
        for batch in (DeviceStager(loader, self.device) if self.device.type == "cuda" else loader):
            inputs, targets = batch["lr_imgs"], batch["hr_imgs"]
            bs, T = inputs[0].shape[0], len(inputs)
            step_acc = torch.zeros_like(acc)
            if training:
                _, outputs = self.step.train_step(inputs, targets, step_acc)
            else:
                _, outputs = self.step.eval_step(inputs, targets, step_acc)
            acc += step_acc * (bs * T)                 # acdc_vsr_trainer.py:119-123 weighting
            count += bs * T
        if self.step.world > 1:
            cnt = torch.tensor([float(count)], device=self.device)
            dist.all_reduce(acc, group=self.pg)
            dist.all_reduce(cnt, group=self.pg)
            count = cnt.item()
        vals = (acc / max(count, 1)).tolist()          # the one host sync of the epoch
        return dict(zip(keys, vals)), batch, outputs

    def _sync_seeds(self):
        """one seed list for all ranks (rank 0's): the per-epoch permutation must be the same everywhere"""
        if self.world > 1:
            t = torch.tensor(self.np_random_seeds, dtype=torch.int64, device=self.device)
            dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
            self.np_random_seeds = t.tolist()

    def train(self):
        if self.np_random_seeds is None:
            self.np_random_seeds = random.sample(range(10000000), k=self.num_epochs)
        self._sync_seeds()
        while self.epoch <= self.num_epochs:
            seed = self.np_random_seeds[self.epoch - 1]
            # base_trainer.py:54; ranks > 0 offset the numpy stream so that their host-side augmentation draws differ
            np.random.seed((seed + self.rank) % (2 ** 32))
            for smp in (self._train_sampler, self._valid_sampler):
                if smp is not None:
                    smp.set_epoch_seed(seed)
            logging.info(f"Epoch {self.epoch}.")
            train_log, train_batch, train_outputs = self._run_epoch("training")
            logging.info(f"Train log: {train_log}.")
            valid_log, valid_batch, valid_outputs = self._run_epoch("validation")
            logging.info(f"Valid log: {valid_log}.")
            if self.lr_scheduler is not None:
                if isinstance(self.lr_scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau):
                    self.lr_scheduler.step(valid_log["Loss"])
                else:
                    self.lr_scheduler.step()
            if self.rank == 0 and self.logger is not None:
                self.logger.write(self.epoch, train_log, train_batch, train_outputs, valid_log, valid_batch,
                                  valid_outputs)
            if self.monitor is not None:
                # Monitor.is_best updates the early-stopping counter (monitor.py:38-63): EVERY rank calls it with the
                # same all-reduced valid_log, so every rank leaves the loop at the same epoch (a rank that kept going
                # would block in the next all-reduce); only rank 0 writes the checkpoints
                saved_path = self.monitor.is_saved(self.epoch)
                if saved_path and self.rank == 0:
                    self.save(saved_path)
                saved_path = self.monitor.is_best(valid_log)
                if saved_path and self.rank == 0:
                    self.save(saved_path)
                if self.monitor.is_early_stopped():
                    break
            self.epoch += 1

    def save(self, path):
        """same checkpoint dictionary as base_trainer.py:224-237."""
        torch.save({"net": self.net.state_dict(), "optimizer": self.optimizer.state_dict(),
                    "lr_scheduler": self.lr_scheduler.state_dict() if self.lr_scheduler else None,
                    "monitor": self.monitor, "epoch": self.epoch, "random_state": random.getstate(),
                    "np_random_seeds": self.np_random_seeds}, path)

    def load(self, path):
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.net.load_state_dict(ck["net"])
        self.optimizer.load_state_dict(ck["optimizer"])
        if ck["lr_scheduler"]:
            self.lr_scheduler.load_state_dict(ck["lr_scheduler"])
        self.monitor = ck["monitor"]
        self.epoch = ck["epoch"] + 1
        random.setstate(ck["random_state"])
        self.np_random_seeds = ck["np_random_seeds"]


AcdcVSRTrainer = VSRTrainer


class MISRTrainer(VSRTrainer):
    """Drop-in for AcdcMISRTrainer (acdc_misr_trainer.py:8-50 on base_trainer.py:99-144): batches carry
    `lr_imgs` (list of num_frames tensors) and ONE `hr_img`; the log weights every batch by the loader's batch
    size.  The step is MISRTrainStep (DUFNet, synchronised BatchNorm across ranks)."""

    def __init__(self, device, train_dataloader, valid_dataloader, net, loss_fns, loss_weights, metric_fns,
                 optimizer, lr_scheduler, logger, monitor, num_epochs, dataset="acdc", use_graph=True,
                 process_group=None):
        super().__init__(device, train_dataloader, valid_dataloader, net, loss_fns, loss_weights, metric_fns,
                         optimizer, lr_scheduler, logger, monitor, num_epochs, dataset, use_graph, process_group)

    def _make_step(self, dataset, use_graph):
        return MISRTrainStep(self.net, self.loss_fns, self.loss_weights, self.metric_fns, self.optimizer, dataset,
                             process_group=self.pg, use_graph=use_graph)

    def _get_inputs_targets(self, batch):
        return batch["lr_imgs"], [batch["hr_img"]]                         # acdc_misr_trainer.py:16-25

    def _run_epoch(self, mode):
        from .data import DeviceStager
        training = mode == "training"
        self.net.train(training)
        loader = self.train_dataloader if training else self.valid_dataloader
        self._check_batch_counts(loader)
        keys = self._keys()
        acc = torch.zeros(len(keys), device=self.device)
        count, batch, outputs = 0, None, None
        batches = DeviceStager(loader, self.device) if self.device.type == "cuda" else loader
        for batch in batches:
            inputs, targets = self._get_inputs_targets(batch)
            step_acc = torch.zeros_like(acc)
            if training:
                _, outs = self.step.train_step(inputs, targets, step_acc)
            else:
                _, outs = self.step.eval_step(inputs, targets, step_acc)
            outputs = outs[0] if len(outs) == 1 else list(outs)
            bs = loader.batch_size or inputs[0].shape[0]                   # base_trainer.py:139-141
            acc += step_acc * bs
            count += bs
        if self.step.world > 1:
            cnt = torch.tensor([float(count)], device=self.device)
            dist.all_reduce(acc, group=self.pg)
            dist.all_reduce(cnt, group=self.pg)
            count = cnt.item()
        vals = (acc / max(count, 1)).tolist()          # the one host sync of the epoch
        return dict(zip(keys, vals)), batch, outputs


AcdcMISRTrainer = MISRTrainer


class SISRTrainer(MISRTrainer):
    """Drop-in for AcdcSISRTrainer (acdc_sisr_trainer.py:8-48): batches carry ONE `lr_img` and ONE `hr_img`; the net
    returns one image (EDSRNet).  Same epoch loop, log weighting, sharding and checkpoints as the other trainers; the
    step is the fused, CUDA-graphed SISRTrainStep."""

    def _make_step(self, dataset, use_graph):
        return SISRTrainStep(self.net, self.loss_fns, self.loss_weights, self.metric_fns, self.optimizer, dataset,
                             process_group=self.pg, use_graph=use_graph)

    def _get_inputs_targets(self, batch):
        return [batch["lr_img"]], [batch["hr_img"]]                        # acdc_sisr_trainer.py:15-24


class SISRSRFBTrainer(SISRTrainer):
    """Drop-in for AcdcSISRSRFBTrainer (acdc_sisr_srfb_trainer.py:8-38): nets with one output per feedback step
    (SRFBNet, DRFSISRNet) - losses averaged over the steps, metrics on the last step."""

    def _make_step(self, dataset, use_graph):
        return SISRSRFBTrainStep(self.net, self.loss_fns, self.loss_weights, self.metric_fns, self.optimizer, dataset,
                                 process_group=self.pg, use_graph=use_graph)


AcdcSISRTrainer = SISRTrainer
AcdcSISRSRFBTrainer = SISRSRFBTrainer


class FRVSRTrainer(VSRTrainer):
    """Drop-in for AcdcFRVSRTrainer / Dsb15FRVSRTrainer (acdc_frvsr_trainer.py:8-120): `lr_imgs` / `hr_imgs` batches, the log
    weights every batch by batch_size * T like the VSR trainer; the step is FRVSRTrainStep."""

    def _make_step(self, dataset, use_graph):
        return FRVSRTrainStep(self.net, self.loss_fns, self.loss_weights, self.metric_fns, self.optimizer, dataset,
                              process_group=self.pg, use_graph=use_graph)


AcdcFRVSRTrainer = FRVSRTrainer


class VSRPredictor:
    """Drop-in for AcdcVSRPredictor / Dsb15VSRPredictor (acdc_vsr_predictor.py:15-110; base_predictor.py:6-23):
    same constructor keywords (`dataset` selects the denormalisation constants) and the same `predict()` log
    (Loss, each loss, each metric; every sequence weighted by batch_size * T, :96-98,160-165).

    Differences, all on purpose: any batch size (the reference insists on 1 because its metrics loop is
    per-sequence; here losses and metrics of all frames of all sequences of a batch come from the fused device
    kernels with one host read-back per batch), and `exported=True` writes `results.csv` only (one row per
    sequence and frame: name_frameNN, metrics, losses - :70-73,101-104); PNG / GIF export stays with the
    reference tooling (out of the hot path).  For batch_size > 1 the loss columns are batch
    means."""

    def __init__(self, device, test_dataloader, net, loss_fns, loss_weights, metric_fns, saved_dir=None,
                 exported=False, dataset="acdc"):
        self.device = torch.device(device)
        self.test_dataloader = test_dataloader
        self.net = net.to(self.device)
        self.loss_fns, self.metric_fns = list(loss_fns), [m.to(self.device) for m in metric_fns]
        self.loss_weights = torch.tensor(loss_weights, dtype=torch.float, device=self.device)
        self.exported = exported
        if exported:
            from pathlib import Path
            self.saved_dir = Path(saved_dir)
        self.step = VSRTrainStep(self.net, self.loss_fns, list(loss_weights), self.metric_fns, None, dataset)

    def _name(self, index):
        data = getattr(self.test_dataloader.dataset, "data", None)
        try:
            entry = data[index][0]
            if isinstance(entry, (int, np.integer)):      # synthetic loader: (sequence number, frame)
                return f"sequence{int(entry):05d}"
            return entry.parts[-1].split(".")[0]          # the reference's lr_path stem (:59-60)
        except (AttributeError, TypeError, IndexError):
            return f"sequence{int(index):05d}"

    def _patients(self, index):
        """patient name of every sample (`patient_..._sid` file stems, acdc_vsr_predictor.py:59-61) when a Cardiac*
        metric asks for bounding boxes"""
        if not any(m.__class__.__name__.startswith("Cardiac") for m in self.metric_fns):
            return None
        idx = index.tolist() if torch.is_tensor(index) else list(index)
        return [self._name(i).split("_")[0] for i in idx]

    def _batches(self):
        if self.device.type == "cuda":
            from .data import DeviceStager
            return DeviceStager(self.test_dataloader, self.device)
        return self.test_dataloader

    def predict(self):
        import csv
        self.net.eval()
        keys = ["Loss"] + [f.__class__.__name__ for f in self.loss_fns] + [m.__class__.__name__ for m in self.metric_fns]
        log = dict.fromkeys(keys, 0.0)
        rows = [["name"] + keys[1 + len(self.loss_fns):] + keys[1:1 + len(self.loss_fns)]]
        count = 0
        for batch in self._batches():
            inputs, targets, index = batch["lr_imgs"], batch["hr_imgs"], batch["index"]
            bs, T = inputs[0].shape[0], len(inputs)
            _, losses, metrics = self.step.eval_frames(inputs, targets, self._patients(index))
            loss = (losses.mean(dim=0) * self.loss_weights).sum()
            host = torch.cat([loss.view(1), losses.mean(dim=0), metrics.mean(dim=(1, 2)) if metrics.numel() else metrics.view(0)])
            vals = host.tolist()                                   # the one host read-back of the batch
            for k, v in zip(keys, vals):
                log[k] += v * bs * T
            count += bs * T
            if self.exported:
                lt, mt = losses.tolist(), metrics.tolist()
                idx = index.tolist() if torch.is_tensor(index) else list(index)
                for i in range(bs):
                    name = self._name(idx[i]).replace("2d+1d", "2d").replace("sequence", "slice")
                    for t in range(T):
                        rows.append([f"{name}_frame{t + 1:0>2d}"] + [m[t][i] for m in mt] + lt[t])
        if self.exported:
            self.saved_dir.mkdir(parents=True, exist_ok=True)
            with open(self.saved_dir / "results.csv", "w", newline="") as f:
                csv.writer(f).writerows(rows)
        for k in log:
            log[k] /= max(count, 1)
        logging.info(f"Test log: {log}.")
        return log


AcdcVSRPredictor = Dsb15VSRPredictor = VSRPredictor


class MISRPredictor(VSRPredictor):
    """Drop-in for AcdcMISRPredictor / Dsb15MISRPredictor (acdc_misr_predictor.py:15-110): every item is a window of
    `num_frames` LR frames and one target frame; the log weights items by the batch size (:96-98); `exported=True`
    writes `results.csv` with one row per item, `<sequence>_frame<t+1>` (:66-73).  Any batch size (the reference
    insists on 1); losses and metrics come from the fused device kernels, one host read-back per batch."""

    def __init__(self, device, test_dataloader, net, loss_fns, loss_weights, metric_fns, saved_dir=None,
                 exported=False, dataset="acdc"):
        super().__init__(device, test_dataloader, net, loss_fns, loss_weights, metric_fns, saved_dir, exported, dataset)
        self.step = MISRTrainStep(self.net, self.loss_fns, list(loss_weights), self.metric_fns, None, dataset)

    def predict(self):
        import csv
        self.net.eval()
        keys = ["Loss"] + [f.__class__.__name__ for f in self.loss_fns] + [m.__class__.__name__ for m in self.metric_fns]
        log = dict.fromkeys(keys, 0.0)
        rows = [["name"] + keys[1 + len(self.loss_fns):] + keys[1:1 + len(self.loss_fns)]]
        data = getattr(self.test_dataloader.dataset, "data", None)
        count = 0
        for batch in self._batches():
            inputs, target, index = batch["lr_imgs"], batch["hr_img"], batch["index"]
            bs = inputs[0].shape[0]
            _, losses, metrics = self.step.eval_frames(inputs, [target], self._patients(index))   # [1, L], [M, 1, bs]
            loss = (losses[0] * self.loss_weights).sum()
            host = torch.cat([loss.view(1), losses[0], metrics.mean(dim=(1, 2)) if metrics.numel() else metrics.view(0)])
            vals = host.tolist()                                               # the one host read-back of the batch
            for k, v in zip(keys, vals):
                log[k] += v * bs
            count += bs
            if self.exported:
                lt, mt = losses[0].tolist(), metrics.tolist()
                idx = index.tolist() if torch.is_tensor(index) else list(index)
                for i in range(bs):
                    name = self._name(idx[i]).replace("2d+1d", "2d").replace("sequence", "slice")
                    try:
                        t = int(data[idx[i]][-1])
                    except (TypeError, IndexError, ValueError):
                        t = 0
                    rows.append([f"{name}_frame{t + 1:0>2d}"] + [m[0][i] for m in mt] + lt)
        if self.exported:
            self.saved_dir.mkdir(parents=True, exist_ok=True)
            with open(self.saved_dir / "results.csv", "w", newline="") as f:
                csv.writer(f).writerows(rows)
        for k in log:
            log[k] /= max(count, 1)
        logging.info(f"Test log: {log}.")
        return log


AcdcMISRPredictor = Dsb15MISRPredictor = MISRPredictor
