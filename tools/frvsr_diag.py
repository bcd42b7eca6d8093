"""FRVSRNet on the GPU against the float64 oracle, per parameter (diagnostic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_frvsrnet import CASES, _losses, _oracle_grads, _oracle_grads64, _state  # noqa: E402
from vsr_b200.frvsr import FRVSRNet  # noqa: E402

for path in CASES:
    fx = torch.load(path)
    _, sr64, lr64, g64 = _oracle_grads64(fx)
    _, _, _, _, g32 = _oracle_grads(fx)
    net = FRVSRNet(**fx["kwargs"])
    net.load_state_dict(_state(fx))
    net = net.cuda()
    sr, lr = net([x.cuda() for x in fx["inputs"]])
    for t in range(len(sr)):
        print(os.path.basename(path), t, "sr err", float((sr[t].double().cpu() - sr64[t]).abs().max()), "lr err",
              float((lr[t].double().cpu() - lr64[t]).abs().max()))
    a, b = _losses(sr, lr, fx)
    (a + b).backward()
    gmax = max(float(v.abs().max()) for v in g64.values())
    rows = sorted(((float((p.grad.double().cpu() - g64[k]).abs().max()) / gmax, float((g32[k].double() - g64[k]).abs().max()) / gmax, k)
                   for k, p in net.named_parameters()), reverse=True)
    for r in rows[:6]:
        print("   ours %.2e  fp32 oracle %.2e  %s" % r)
    smax = max(float(v.abs().max()) for k, v in g64.items() if k.startswith("srnet."))
    print("  srnet only, relative to the largest srnet gradient (", smax, "), and each parameter's own max:")
    for k, p in net.named_parameters():
        if k.startswith("srnet."):
            d = float((p.grad.double().cpu() - g64[k]).abs().max())
            print("   %-34s ours %.2e (own %.2e)  fp32 oracle %.2e" % (k, d / smax, d / float(g64[k].abs().max()),
                                                                     float((g32[k].double() - g64[k]).abs().max()) / smax))
