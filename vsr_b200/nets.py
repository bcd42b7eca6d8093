"""Drop-in SR nets: same class names, constructor kwargs, forward I/O and state_dict layout as the
reference (src/model/nets/drf_net.py:8-49, drf_sisr_net.py:8-50, base_net.py:5-13), executed by
hand-written sm_100a kernels through libvsr_sm100.so.

The sub-modules below (`in_block`, `f_block`, `out_block`) only *hold parameters* — they give the
reference's state_dict keys, shapes, dtypes and default initialisation (same construction order,
hence the same values under the same torch seed).  They are never called: forward and backward
run in vsr_b200.drf_engine.  All parameters are views into one flat fp32 bucket (`net.flat`), and
all gradients of a step are views into one flat bucket (`net.flat_grad`) — the unit the
data-parallel trainer all-reduces.
"""
import math

import torch
import torch.nn as nn

from . import _flat
from .drf_engine import DrfEngine
from .drf_plan import PROJ, DrfPlan

# 'fp32': CUDA-core strict mode; 'bf16': bf16 storage + tcgen05; 'bf16x3' (alias 'tf32': the reference configs' name for
# "tensor cores, fp32 storage"): fp32 storage, every product as three bf16 tcgen05 products (ops.SplitOps) - strict
# mode accuracy (<= 1e-4 of the reference) on tensor cores
_PRECISIONS = {"fp32": torch.float32, "bf16": torch.bfloat16, "bf16x3": torch.float32, "tf32": torch.float32}
_TC_LAYOUT = ("bf16", "bf16x3", "tf32")      # plans with 64-channel taps and swizzled bf16 weight slabs


class TripledSlabs:
    """A weight buffer of the bf16x3 mode addressed with the plan's own offsets: every layer's slabs are stored three times
    over ([wh | wh | wl] per group, DrfPlan.split_index), so the plan's [w_off, w_off + w_numel) is [3 w_off, 3 (w_off + w_numel))"""

    def __init__(self, buf):
        self.buf = buf

    def __getitem__(self, s):
        return self.buf[3 * s.start:3 * s.stop]


def split_mode(net):
    """True when `net` runs the strict mode on tensor cores (precision 'bf16x3' / 'tf32'; the float64 emulation runs of the
    tests keep the plain layout)"""
    return net.precision in ("bf16x3", "tf32") and net.flat.dtype != torch.float64 and getattr(net._backend(), "split", False)


def packed_weight_state(net, P, dev, act, split=None):
    """the packed weight / bias buffers and packing maps shared by the nets that keep their own state dict (RBPNet,
    EDSRNet, ...): plain slabs, or the tripled bf16 slabs of the bf16x3 mode (`split`: None = by the net's precision)"""
    if split_mode(net) if split is None else split:
        return {"fwd_w": TripledSlabs(torch.empty(3 * P.fwd_w_numel, dtype=torch.bfloat16, device=dev)),
                "bwd_w": TripledSlabs(torch.empty(3 * P.bwd_w_numel, dtype=torch.bfloat16, device=dev)),
                "fwd_w_idx": torch.from_numpy(P.split_index("fwd")).to(dev), "bwd_w_idx": torch.from_numpy(P.split_index("bwd")).to(dev),
                "split": True}
    return {"fwd_w": torch.empty(P.fwd_w_numel, dtype=act, device=dev), "bwd_w": torch.empty(P.bwd_w_numel, dtype=act, device=dev),
            "fwd_w_idx": torch.from_numpy(P.fwd_w_idx).to(dev), "bwd_w_idx": torch.from_numpy(P.bwd_w_idx).to(dev), "split": False}


def pack_weights(net, st, need_bwd):
    ops = net._backend()
    if st["split"]:
        ops.gather_split(net.flat, st["fwd_w_idx"], st["fwd_w"].buf)
        if need_bwd:
            ops.gather_split(net.flat, st["bwd_w_idx"], st["bwd_w"].buf)
    else:
        ops.gather(net.flat, st["fwd_w_idx"], st["fwd_w"])
        if need_bwd:
            ops.gather(net.flat, st["bwd_w_idx"], st["bwd_w"])
    ops.gather(net.flat, st["fwd_b_idx"], st["fwd_b"])


class BaseNet(nn.Module):
    """The base class for all nets (reference: base_net.py:5-13)."""

    def __init__(self):
        super().__init__()

    def __repr__(self):
        n = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return super().__repr__() + f"\nTrainable parameters: {n / 1e6} M\nMemory usage: {(n * 4) / (1 << 20)} MB"


def _seq(**mods):
    s = nn.Sequential()
    for k, m in mods.items():
        s.add_module(k, m)
    return s


def _prelu():
    return nn.PReLU(num_parameters=1, init=0.2)


class _FBlockParams(nn.Module):
    """Parameter container with the reference's _FBlock naming (drf_net.py:61-116)."""

    def __init__(self, F, G, r):
        super().__init__()
        k, s, p = PROJ[r]
        self.in_block = _seq(conv=nn.Conv2d(2 * F, F, 1), prelu=_prelu())
        self.up_blocks, self.down_blocks = nn.ModuleList(), nn.ModuleList()
        for g in range(G):
            if g == 0:
                self.up_blocks.append(_seq(deconv=nn.ConvTranspose2d(F, F, k, s, p), prelu=_prelu()))
                self.down_blocks.append(_seq(conv=nn.Conv2d(F, F, k, s, p), prelu=_prelu()))
            else:
                self.up_blocks.append(_seq(conv1=nn.Conv2d(F * (g + 1), F, 1), prelu1=_prelu(),
                                           deconv2=nn.ConvTranspose2d(F, F, k, s, p), prelu2=_prelu()))
                self.down_blocks.append(_seq(conv1=nn.Conv2d(F * (g + 1), F, 1), prelu1=_prelu(),
                                             conv2=nn.Conv2d(F, F, k, s, p), prelu2=_prelu()))
        self.out_block = _seq(conv=nn.Conv2d(F * G, F, 1), prelu=_prelu())
        self._hidden_state = None

    @property
    def hidden_state(self):
        return self._hidden_state

    @hidden_state.setter
    def hidden_state(self, state):
        self._hidden_state = state


def _out_block(F, cout, r):
    s = nn.Sequential()
    if math.log(r, 2) % 1 == 0:
        n = int(math.log(r, 2))
        for i in range(n):
            s.add_module(f"conv{i + 1}", nn.Conv2d(F, 4 * F, 3, padding=1))
            s.add_module(f"pixelshuffle{i + 1}", nn.PixelShuffle(2))
        s.add_module(f"conv{n + 1}", nn.Conv2d(F, cout, 3, padding=1))
    else:
        s.add_module("conv1", nn.Conv2d(F, 9 * F, 3, padding=1))
        s.add_module("pixelshuffle1", nn.PixelShuffle(3))
        s.add_module("conv2", nn.Conv2d(F, cout, 3, padding=1))
    return s


class _DRFFunction(torch.autograd.Function):
    """One autograd node for the whole T-frame forward; backward is the engine's own schedule."""

    @staticmethod
    def forward(ctx, net, T, *args):
        eng = net._engine
        eng.pack(net.flat, need_bwd=True)
        outs, saved = eng.forward([f.contiguous() for f in args[:T]], save=True)
        ctx.net, ctx.saved, ctx.T = net, saved, T
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        net = ctx.net
        gflat = net._engine.backward(ctx.saved, list(grads))
        ctx.saved = None
        net.flat_grad = gflat
        pg = []
        for p in net._plan.params.values():
            n = 1
            for d in p.shape:
                n *= d
            pg.append(gflat[p.offset:p.offset + n].view(p.shape))
        return (None, None) + (None,) * ctx.T + tuple(pg)


class _DRFBase(BaseNet):
    _variant = "drf"

    def _setup(self, in_channels, out_channels, num_features, num_groups, upscale_factor, precision):
        if upscale_factor not in [2, 3, 4, 8]:
            raise ValueError(f"The upscale factor should be 2, 3, 4 or 8. Got {upscale_factor}.")
        if precision not in _PRECISIONS:
            raise ValueError(f"precision should be one of {sorted(_PRECISIONS)}. Got {precision!r}.")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_features, self.num_groups = num_features, num_groups
        self.upscale_factor, self.precision = upscale_factor, precision
        F = num_features
        first = _seq(conv1=nn.Conv2d(in_channels, 4 * F, 3, padding=1), prelu1=_prelu(),
                     conv2=nn.Conv2d(4 * F, F, 1), prelu2=_prelu())
        if self._variant == "drf":
            self.in_block = first
            self.f_block = _FBlockParams(F, num_groups, upscale_factor)
            self.out_block = _out_block(F, out_channels, upscale_factor)
        else:                       # SRFBNet naming (srfb_net.py:34-36,137-151)
            k, s_, p_ = PROJ[upscale_factor]
            self.lrf_block = first
            self.f_block = _FBlockParams(F, num_groups, upscale_factor)
            self.r_block = _seq(deconv1=nn.ConvTranspose2d(F, F, k, s_, p_), prelu1=_prelu(),
                                conv2=nn.Conv2d(F, out_channels, 3, padding=1))
        self._plan = DrfPlan(in_channels, out_channels, F, num_groups, upscale_factor,
                             bf16=(precision in _TC_LAYOUT), variant=self._variant)
        names = [n for n, _ in self.named_parameters()]
        assert names == list(self._plan.params), "parameter order differs from the plan"
        self._engine = None
        self._ops = None          # tests may set an emulated backend here; product uses CudaOps
        self.flat = None
        self.flat_grad = None
        self._flatten()

    # ---- flat parameter bucket -------------------------------------------------------------
    def _flatten(self):
        params = list(self.parameters())
        dev, dt = params[0].device, params[0].dtype
        flat = torch.empty(self._plan.n_params, dtype=dt, device=dev)
        for p, ref in zip(params, self._plan.params.values()):
            n = p.numel()
            flat[ref.offset:ref.offset + n].copy_(p.data.reshape(-1))
            p.data = flat[ref.offset:ref.offset + n].view(ref.shape)
        self.flat = flat
        if self._engine is not None and (self._engine.device != dev or self._engine.param_dtype != dt):
            self._engine = None

    def _is_flat(self):
        return _flat.is_flat(self)      # cached: the module tree is walked only after a parameter registration

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)     # .to(device) / .double(): params moved in place
        self._flatten()
        return out

    def _backend(self):
        if self._ops is not None:
            return self._ops
        if self.flat.device.type != "cuda":
            raise RuntimeError("vsr_b200 nets run on CUDA only (there is no CPU fallback); call .to('cuda')")
        from .ops import cuda_ops, split_ops
        return split_ops() if self.precision in ("bf16x3", "tf32") else cuda_ops()

    def _run(self, frames):
        if not self._is_flat():
            self._flatten()
        ops = self._backend()
        if self._engine is None or self._engine.ops is not ops:
            act = _PRECISIONS[self.precision]
            if self.flat.dtype == torch.float64:     # emulated high-precision oracle runs (tests)
                act = torch.float64
            self._engine = DrfEngine(self._plan, ops, self.flat.device, act, self.flat.dtype)
        frames = list(frames)
        for f in frames:
            if f.dim() != 4 or f.shape[1] != self.in_channels:
                raise ValueError(f"expected frames of shape [N,{self.in_channels},h,w], got {tuple(f.shape)}")
        needs_grad = torch.is_grad_enabled() and (
            any(p.requires_grad for p in self.parameters()) or any(f.requires_grad for f in frames))
        if needs_grad:
            outs = _DRFFunction.apply(self, len(frames), *frames, *self.parameters())
        else:
            self._engine.pack(self.flat, need_bwd=False)
            outs, _ = self._engine.forward([f.contiguous() for f in frames], save=False)
        return list(outs)


class DRFNet(_DRFBase):
    """Deep Recurrent Feedback Network for video SR (reference: drf_net.py:8-49).

    Args: in_channels, out_channels, num_features, num_groups, upscale_factor (2, 3, 4 or 8) — as the
    reference; precision ('fp32' = CUDA-core strict mode, 'bf16' = tcgen05 tensor-core mode, 'bf16x3' / 'tf32' = strict
    accuracy on tensor cores: fp32 maps, three bf16 tcgen05 products per product; num_features % 64 == 0).
    forward(list of T tensors [N,C,h,w]) -> list of T tensors [N,C,r*h,r*w].
    """

    def __init__(self, in_channels, out_channels, num_features, num_groups, upscale_factor, precision="fp32"):
        super().__init__()
        self._setup(in_channels, out_channels, num_features, num_groups, upscale_factor, precision)

    def forward(self, inputs):
        return self._run(inputs)


class DRFSISRNet(_DRFBase):
    """DRFN for single-image SR (reference: drf_sisr_net.py:8-50): the same blocks iterated
    `num_steps` times on one image; returns the list of the per-step outputs."""

    def __init__(self, in_channels, out_channels, num_steps, num_features, num_groups, upscale_factor,
                 precision="fp32"):
        super().__init__()
        self.num_steps = num_steps
        self._setup(in_channels, out_channels, num_features, num_groups, upscale_factor, precision)

    def forward(self, input):
        return self._run([input] * self.num_steps)


class SRFBNet(_DRFBase):
    """Super-Resolution FeedBack Network (reference: srfb_net.py:8-50): `num_steps` iterations of the
    feedback block on one image; each step's output = bilinear(input) + r_block(features).
    forward(tensor [N,C,h,w]) -> list of num_steps tensors [N,C,r*h,r*w]."""
    _variant = "srfb"

    def __init__(self, in_channels, out_channels, num_steps, num_features, num_groups, upscale_factor,
                 precision="fp32"):
        super().__init__()
        self.num_steps = num_steps
        if in_channels != out_channels:
            raise ValueError("SRFBNet adds the up-sampled input to the output: in_channels must equal out_channels")
        self._setup(in_channels, out_channels, num_features, num_groups, upscale_factor, precision)

    def forward(self, input):
        return self._run([input] * self.num_steps)
