"""Data-parallel epoch loop on CPU (gloo, world_size 2, kernels emulated by tests/emu.py):
 - the rank-aware sampler: same permutation on every rank from the epoch's seed, disjoint shards that cover the
   dataset, equal lengths (the reference reseeds per epoch at base_trainer.py:54);
 - VSRTrainer: ranks shard the samples, every rank takes the same early-stopping decision (the reference's Monitor
   updates its counter inside is_best, monitor.py:38-63) and leaves the loop at the same epoch, rank 0 alone
   writes checkpoints, and the ranks end with identical weights."""
import math
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vsr_b200.data import Dataloader, ShardedSampler, SyntheticCineDataset, shard_loader


def test_sharded_sampler_partitions_and_is_seed_deterministic():
    n, world = 23, 4
    shards = []
    for rank in range(world):
        s = ShardedSampler(n, rank, world, shuffle=True)
        s.set_epoch_seed(1234)
        shards.append(list(s))
    assert len({len(x) for x in shards}) == 1 and len(shards[0]) == math.ceil(n / world)
    flat = [i for x in shards for i in x]
    assert set(flat) == set(range(n))                       # every sample is seen
    assert len(flat) - len(set(flat)) == world * math.ceil(n / world) - n      # only the wrap-around padding repeats
    again = ShardedSampler(n, 2, world, shuffle=True)
    again.set_epoch_seed(1234)
    assert list(again) == shards[2]                         # same seed -> same shard
    again.set_epoch_seed(1235)
    assert list(again) != shards[2]                         # next epoch -> another permutation
    plain = ShardedSampler(n, 1, world, shuffle=False)
    assert list(plain) == [1, 5, 9, 13, 17, 21]
    assert list(ShardedSampler(n, 3, world, shuffle=False)) == [3, 7, 11, 15, 19, 0]      # wrap-around padding


def test_shard_loader_rebuilds_torch_dataloaders():
    ds = SyntheticCineDataset(2, num_frames=3, type="train", num_sequences=1, patch_size=(16, 16), seed=3)
    loader = Dataloader(ds, batch_size=4, shuffle=True, num_workers=0)
    sharded, sampler = shard_loader(loader, 1, 2)
    assert isinstance(sampler, ShardedSampler) and sharded.batch_size == 4 and sharded.dataset is ds
    assert len(sharded) == math.ceil(math.ceil(len(ds) / 2) / 4)
    same, none = shard_loader(loader, 0, 1)
    assert same is loader and none is None


class _Monitor:
    """the reference's Monitor logic (monitor.py:13-63) without the directory handling"""

    def __init__(self, root, early_stop):
        self.root, self.early_stop, self.best, self.not_improved_count = root, early_stop, -math.inf, 0

    def is_saved(self, epoch):
        return os.path.join(self.root, f"model_{epoch}.pth")

    def is_best(self, valid_log):
        if valid_log["Loss"] > self.best + 1e9:         # mode 'max' on a loss that cannot rise this much: never improves
            self.best = valid_log["Loss"]
            self.not_improved_count = 0
            return os.path.join(self.root, "model_best.pth")
        self.not_improved_count += 1
        return None

    def is_early_stopped(self):
        return self.not_improved_count == self.early_stop


def _worker(rank, world, port, root, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import random
    from tests.emu import EmuOps
    from vsr_b200.metrics import PSNR
    from vsr_b200.nets import DRFNet
    from vsr_b200.optim import FlatAdam
    from vsr_b200.runner import VSRTrainer
    random.seed(7 + rank)            # ranks disagree on purpose: the trainer must broadcast rank 0's epoch seeds
    torch.manual_seed(0)
    net = DRFNet(1, 1, 8, 1, 2)
    net._ops = EmuOps()
    train = SyntheticCineDataset(2, num_frames=2, type="train", num_sequences=1, patch_size=(12, 12), seed=5)
    train.data = train.data[:6]
    valid = SyntheticCineDataset(2, num_frames=2, type="train", num_sequences=1, patch_size=(12, 12), seed=6)
    valid.data = valid.data[:4]
    opt = FlatAdam(net.parameters(), lr=1e-3)
    seen = []

    class Spy(Dataloader):
        def __iter__(self):
            for b in super().__iter__():
                seen.append(b["index"].tolist())
                yield b

    tr = VSRTrainer("cpu", Spy(train, batch_size=1, shuffle=True), Dataloader(valid, batch_size=1), net,
                    [torch.nn.L1Loss()], [1.0], [PSNR()], opt, None, None, _Monitor(root, early_stop=2), num_epochs=5)
    tr.train()
    q.put((rank, tr.epoch, tr.np_random_seeds, seen, net.flat.detach().numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_trainer_shards_and_stops_together(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, e0, s0, seen0, w0), (r1, e1, s1, seen1, w1) = res
    assert e0 == e1 == 3                      # best at epoch 1, then 2 epochs without improvement: both ranks stop together
    assert s0 == s1                           # one seed list (rank 0's)
    # epoch 1: three batches of one sample per rank, disjoint, together the six training samples
    a, b = [i for x in seen0[:3] for i in x], [i for x in seen1[:3] for i in x]
    assert sorted(a + b) == list(range(6))
    assert np.array_equal(w0, w1)             # identical weights on both ranks (same all-reduced gradients)
    saved = sorted(os.listdir(tmp_path))
    assert saved == ["model_1.pth", "model_2.pth", "model_3.pth", "model_best.pth"]
