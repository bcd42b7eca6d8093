"""PSNR / SSIM with fused kernels (reference: src/model/metrics.py:9-113).

Same class names and constructor arguments as the reference; an optional `dataset` argument
fuses the trainer's `denormalize` (src/utils.py) into the metric kernel so the three extra
passes per tensor disappear (acdc_vsr_trainer.py:100-101)."""
import math

import torch
import torch.nn as nn

from .utils import DATASET_STATS


def _check(output, target):
    if not output.is_cuda:
        raise RuntimeError("vsr_b200 metrics run on CUDA only (there is no CPU fallback)")
    if output.shape != target.shape:
        raise ValueError("output and target must have the same shape")


class PSNR(nn.Module):
    """The PSNR score (metrics.py:9-36). Args: size_average (default True), max_value (default 255),
    dataset (optional: fuse denormalize for 'acdc' / 'dsb15')."""

    def __init__(self, size_average=True, max_value=255, dataset=None):
        super().__init__()
        self.size_average, self.max_value = size_average, max_value
        self.mean, self.std = DATASET_STATS[dataset] if dataset else (0.0, -1.0)

    @torch.no_grad()
    def forward(self, output, target):
        from .ops import cuda_ops
        _check(output, target)
        ops = cuda_ops()
        o, t = output.detach().float().contiguous(), target.detach().float().contiguous()
        n = o.shape[0]
        ws = torch.empty(ops.metric_workspace(n, o.numel() // n) // 4 + 4, dtype=torch.float32, device=o.device)
        out = torch.empty(n, dtype=torch.float32, device=o.device)
        ops.psnr(o, t, self.mean, self.std, float(self.max_value), out, ws)
        return out.mean() if self.size_average else out


class SSIM(nn.Module):
    """The SSIM score (metrics.py:39-113), dim=2 (images [N,C,H,W]) or dim=3 (volumes [N,C,D,H,W]). The window
    reproduces the reference's exp(-((i-5)/(2*sigma))^2), sigma=1.5, normalised (metrics.py:74-77); the
    product window is applied separably (one fused pass for dim=2, three passes for dim=3)."""

    def __init__(self, dim=2, channels=1, size_average=True, value_range=255, dataset=None):
        super().__init__()
        if dim not in (2, 3):
            raise ValueError(f"Only dim=2, 3 are supported. Received dim={dim}.")
        self.dim, self.channels = dim, channels
        self.size_average, self.value_range = size_average, value_range
        self.c1, self.c2 = (0.01 * value_range) ** 2, (0.03 * value_range) ** 2
        i = torch.arange(11, dtype=torch.float32)
        g = 1 / (1.5 * math.sqrt(2 * math.pi)) * torch.exp(-((i - 5) / (2 * 1.5)) ** 2)
        self.register_buffer("window", g / g.sum(), persistent=False)   # not in the reference state_dict
        # the reference registers the full 2-D kernel as `weight`; keep the buffer for state parity
        k = torch.outer(g, g) if dim == 2 else torch.einsum("i,j,k->ijk", g, g, g)
        self.register_buffer("weight", (k / k.sum()).view(1, 1, *k.shape).repeat(channels, *[1] * (dim + 1)))
        self.groups = channels
        self.mean, self.std = DATASET_STATS[dataset] if dataset else (0.0, -1.0)

    @torch.no_grad()
    def forward(self, output, target):
        from .ops import cuda_ops
        _check(output, target)
        ops = cuda_ops()
        o, t = output.detach().float().contiguous(), target.detach().float().contiguous()
        if o.dim() != self.dim + 2:
            raise ValueError(f"SSIM(dim={self.dim}) expects a {self.dim + 2}-D tensor, got {o.dim()}-D")
        if self.dim == 3:
            n, c, d, h, w = o.shape
            o, t = o.view(n * c, d, h, w), t.view(n * c, d, h, w)
            ws = torch.empty(ops.ssim3d_workspace(n * c, d, h, w) // 4 + 4, dtype=torch.float32, device=o.device)
            out = torch.empty(n * c, dtype=torch.float32, device=o.device)
            ops.ssim3d(o, t, self.window, self.mean, self.std, self.c1, self.c2, out, ws)
            out = out.view(n, c).mean(1)
            return out.mean() if self.size_average else out
        n, c, h, w = o.shape
        # depthwise (groups == channels): every channel is an independent image
        o, t = o.view(n * c, h, w), t.view(n * c, h, w)
        ws = torch.empty(ops.metric_workspace(n * c, h * w) // 4 + 4, dtype=torch.float32, device=o.device)
        out = torch.empty(n * c, dtype=torch.float32, device=o.device)
        ops.ssim(o, t, self.window, self.mean, self.std, self.c1, self.c2, out, ws)
        out = out.view(n, c).mean(1)
        return out.mean() if self.size_average else out


class _Cardiac(nn.Module):
    """Metric restricted to the cardiac bounding box of a patient (metrics.py:116-165): the pickle at
    `coordinates_path` maps patient name -> (h0, hn, w0, wn); forward(output, target, name)."""
    metric_cls = None

    def __init__(self, coordinates_path, **kwargs):
        super().__init__()
        import pickle
        self.metric = self.metric_cls(**kwargs)
        with open(coordinates_path, "rb") as f:
            self.coordinates = pickle.load(f)

    def forward(self, output, target, name):
        h0, hn, w0, wn = self.coordinates[name]
        return self.metric(output[..., h0:hn, w0:wn], target[..., h0:hn, w0:wn])


class CardiacPSNR(_Cardiac):
    """The cardiac PSNR score (metrics.py:116-139)."""
    metric_cls = PSNR


class CardiacSSIM(_Cardiac):
    """The cardiac SSIM score (metrics.py:142-165)."""
    metric_cls = SSIM
