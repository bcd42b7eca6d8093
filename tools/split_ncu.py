"""One eager bf16x3 training step of the config-2 model (for ncu captures of the split-mode kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import MODEL, make_batches           # noqa: E402
from vsr_b200.nets import DRFNet                # noqa: E402

lrs, hrs = make_batches(1, 32, seed=3, pinned=False)[0]
x, y = [t.cuda() for t in lrs], [t.cuda() for t in hrs]
torch.manual_seed(0)
net = DRFNet(precision="bf16x3", **MODEL).cuda()
for _ in range(2):
    net.zero_grad()
    outs = net(x)
    torch.stack([torch.nn.L1Loss()(o, t) for o, t in zip(outs, y)]).mean().backward()
torch.cuda.synchronize()
print("ok")
