"""Worker of tests/test_trainstep_gpu.py::test_two_rank_nccl_graphed_step_equals_one_rank_on_the_whole_batch:
launched by torchrun with 2 ranks; every rank trains on its half of a 4-patch batch through the CUDA-graphed
VSRTrainStep (NCCL all-reduces of the gradient ranges captured inside the graph); rank 0 saves the result."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import MODEL, make_batches
    from oracle import restated
    from vsr_b200.metrics import PSNR, SSIM
    from vsr_b200.nets import DRFNet
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import VSRTrainStep
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    sd0 = restated.drfnet_init(**MODEL)
    net = DRFNet(precision="bf16", **MODEL)
    net.load_state_dict(sd0)
    net = net.to(dev)
    opt = FlatAdam(net.parameters(), lr=1e-3, eps=1e-4)
    step = VSRTrainStep(net, [torch.nn.L1Loss()], [1.0], [PSNR().to(dev), SSIM().to(dev)], opt, "acdc", use_graph=True)
    lrs, hrs = make_batches(1, 4, seed=21, pinned=False)[0]
    per = 4 // world
    x = [t[rank * per:(rank + 1) * per].to(dev) for t in lrs]
    y = [t[rank * per:(rank + 1) * per].to(dev) for t in hrs]
    losses = []
    for _ in range(4):
        lv, _ = step.train_step(x, y)
        l = lv[:1].clone()
        dist.all_reduce(l)
        losses.append(float(l[0]) / world)
    torch.cuda.synchronize()
    flats = [torch.empty_like(net.flat) for _ in range(world)]
    dist.all_gather(flats, net.flat.detach().contiguous())
    if rank == 0:
        torch.save({"flat": flats[0].cpu(), "flat_rank1": flats[1].cpu(), "losses": losses, "graphed": bool(step._graphs)},
                   sys.argv[1])
    step._graphs.clear()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
