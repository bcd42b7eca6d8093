"""Where do the CUDA kernels and the torch emulation (tests/emu.py, run on the same GPU tensors) part ways?
Runs one DUFNet golden through the CUDA back end with ONE op family at a time replaced by the emulation and
prints the largest gradient difference against the all-emulation run.

    python tools/duf_diag.py [golden name] [fp32|bf16]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden_duf import duf_fill  # noqa: E402
from tests.emu import EmuOps  # noqa: E402
from vsr_b200.duf import DUFNet  # noqa: E402
from vsr_b200.ops import CudaOps  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "dufnet28_x2"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
fx = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"))
sd = duf_fill({k: torch.zeros(s, dtype=fx["state_dtypes"][k]) for k, s in fx["state_shapes"].items()}, fx["state_seed"])
FAMILIES = {
    "none": [], "all": None,
    "tapgemm": ["tapgemm"], "wgrad": ["tapgemm_wgrad", "tapgemm_wgrad_workspace"], "colsum": ["colsum"],
    "first": ["conv3x3_first", "conv3x3_first_bwd"], "bn_stats": ["bn_stats", "bn_finalize"], "bn_relu": ["bn_relu"],
    "bn_relu_bwd": ["bn_relu_bwd"], "duf_filter": ["duf_filter"], "duf_filter_bwd": ["duf_filter_bwd"],
    "copy_gather": ["copy_window", "gather", "gather_add"],
    "none2": [], "bn_stats_only": ["bn_stats"], "bn_finalize_only": ["bn_finalize"], "first_fwd": ["conv3x3_first"],
    "first_bwd": ["conv3x3_first_bwd"], "tapgemm_fwd": ["tapgemm:fwd"], "tapgemm_dgrad": ["tapgemm:bwd"],
}
WATCH = ["denseLayer.tail.bn.bias", "denseLayer.tail.bn.weight", "denseLayer.conv11.bn2.bias", "denseLayer.conv9.conv2.weight", "head.weight"]


def run(family):
    net = DUFNet(precision=prec, **fx["kwargs"])
    net.load_state_dict(sd)
    net = net.to("cuda").train()
    emu = EmuOps()
    if FAMILIES[family] is None:
        net._ops = emu
    else:
        ops = CudaOps()
        for m in FAMILIES[family]:
            if ":" in m:
                want_bias = m.endswith("fwd")
                cu = ops.tapgemm
                setattr(ops, "tapgemm", lambda *a, _cu=cu, _w=want_bias, **kw: (emu.tapgemm if (kw.get("bias") is not None) == _w else _cu)(*a, **kw))
            else:
                setattr(ops, m, getattr(emu, m))
        net._ops = ops
    out = net([f.cuda() for f in fx["inputs"]])
    torch.nn.L1Loss()(out, fx["target"].cuda()).backward()
    return out.detach(), {k: p.grad.clone() for k, p in net.named_parameters()}


ref_out, ref = run("all")
gmax = max(float(g.abs().max()) for g in ref.values())
print("emu out vs golden", float((ref_out.cpu() - fx["output"]).abs().max() / fx["output"].abs().max()))
for fam in FAMILIES:
    if fam == "all":
        continue
    out, g = run(fam)
    worst = max(((float((g[k] - ref[k]).abs().max()) / gmax, k) for k in ref))
    print(f"emulated: {fam:16s} out diff {float((out - ref_out).abs().max() / ref_out.abs().max()):.2e}  "
          f"worst grad diff/gmax {worst[0]:.2e} ({worst[1]})  " + " ".join(f"{float((g[k] - ref[k]).abs().max()) / gmax:.1e}" for k in WATCH))
