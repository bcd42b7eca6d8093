"""Host logic of the fused training step, on CPU through the kernel emulation:
 - one rank: the fused step (loss kernels -> engine backward -> FlatAdam) equals the reference's
   step (net -> L1 -> backward -> torch.optim.Adam) restated with the oracle;
 - two ranks (gloo): sharding the batch + all-reducing the flat gradient equals the one-rank step
   on the whole batch (DRFNet has no cross-sample op: SURVEY.md §8e)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restated
from tests.emu import EmuOps
from vsr_b200.metrics import PSNR, SSIM
from vsr_b200.nets import DRFNet
from vsr_b200.optim import FlatAdam
from vsr_b200.runner import VSRTrainStep

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _setup(fx, world=None):
    net = DRFNet(**fx["kwargs"])
    net.load_state_dict(fx["state_dict"])
    net._ops = EmuOps()
    opt = FlatAdam(net.parameters(), lr=1e-3)
    step = VSRTrainStep(net, [torch.nn.L1Loss()], [1.0], [PSNR(), SSIM()], opt, "acdc")
    return net, opt, step


def _oracle_steps(fx, n_steps, lr=1e-3):
    sd = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    opt = torch.optim.Adam(list(sd.values()), lr=lr)
    r = fx["kwargs"]["upscale_factor"]
    losses = []
    for _ in range(n_steps):
        outs = restated.drfnet_forward(fx["inputs"], sd, r)
        loss = restated.vsr_loss(outs, fx["targets"], restated.l1_loss)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return sd, losses, outs


def test_fused_step_equals_reference_step():
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g2_x2.pt"))
    net, opt, step = _setup(fx)
    acc = torch.zeros(4)
    got_losses = []
    for _ in range(3):
        acc.zero_()
        lv, outs = step.train_step(fx["inputs"], fx["targets"], acc)
        got_losses.append(float(lv[0]))
    sd, want_losses, ref_outs = _oracle_steps(fx, 3)
    for a, b in zip(got_losses, want_losses):
        assert abs(a - b) <= 1e-5 * abs(b)
    for k, p in net.named_parameters():
        assert (p.data - sd[k].data).abs().max() <= 2e-5, k
    # logged values: Loss, L1Loss, PSNR, SSIM of the LAST step's training outputs
    psnr, ssim = restated.vsr_metrics([o.detach() for o in ref_outs], fx["targets"])
    assert abs(float(acc[0]) - want_losses[-1]) <= 1e-5 and abs(float(acc[1]) - want_losses[-1]) <= 1e-5
    assert abs(float(acc[2]) - float(psnr)) <= 1e-3
    assert abs(float(acc[3]) - float(ssim)) <= 1e-4


def test_eval_step_matches_oracle():
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g3_x4.pt"))
    net, opt, step = _setup(fx)
    acc = torch.zeros(4)
    lv, outs = step.eval_step(fx["inputs"], fx["targets"], acc)
    assert abs(float(lv[0]) - float(fx["loss_l1"])) <= 1e-5
    assert abs(float(acc[2]) - float(fx["psnr"])) <= 1e-3
    assert abs(float(acc[3]) - float(fx["ssim"])) <= 1e-4


def test_flat_adam_state_dict_roundtrip():
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g2_x2.pt"))
    net, opt, step = _setup(fx)
    step.train_step(fx["inputs"], fx["targets"])
    sd = opt.state_dict()
    assert len(sd["state"]) == len(list(net.parameters()))
    net2, opt2, step2 = _setup(fx)
    net2.load_state_dict(net.state_dict())
    opt2.load_state_dict(sd)
    step.train_step(fx["inputs"], fx["targets"])
    step2.train_step(fx["inputs"], fx["targets"])
    assert torch.equal(net.flat, net2.flat)


def _dp_worker(rank, world, port, path, q, comm_mode="overlap"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VSR_COMM_MODE=comm_mode)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    fx = torch.load(path)
    net, opt, step = _setup(fx)
    n = fx["inputs"][0].shape[0]
    per = n // world
    xs = [x[rank * per:(rank + 1) * per] for x in fx["inputs"]]
    ys = [y[rank * per:(rank + 1) * per] for y in fx["targets"]]
    for _ in range(2):
        step.train_step(xs, ys)
    if rank == 0:
        q.put(net.flat.detach().numpy().copy())   # by value (a tensor would be passed by fd)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("comm_mode", ["overlap", "serial"])
def test_two_rank_data_parallel_equals_single_rank(comm_mode):
    # overlap: bucket-by-bucket all-reduce from the engine's callback; serial: one all-reduce after backward
    path = os.path.join(GOLDEN, "drfnet_f8_g2_x2.pt")   # batch of 2 -> one sample per rank
    fx = torch.load(path)
    net, opt, step = _setup(fx)
    for _ in range(2):
        step.train_step(fx["inputs"], fx["targets"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + (7 if comm_mode == "serial" else 0)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, path, q, comm_mode)) for r in range(2)]
    for p in procs:
        p.start()
    flat = torch.from_numpy(q.get(timeout=180))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # mean of the per-rank mean-losses == whole-batch mean loss (equal shard sizes)
    assert (flat - net.flat).abs().max() <= 1e-6


def test_predictor_log_and_csv_match_reference_loop(tmp_path):
    """VSRPredictor.predict() (acdc_vsr_predictor.py:30-110) on CPU through the kernel emulation: the log and
    the per-frame rows of results.csv equal the reference's loop body restated with the oracle."""
    import csv
    from vsr_b200.runner import VSRPredictor
    fx = torch.load(os.path.join(GOLDEN, "drfnet_f8_g2_x2.pt"))
    net = DRFNet(**fx["kwargs"])
    net.load_state_dict(fx["state_dict"])
    net._ops = EmuOps()
    n = fx["inputs"][0].shape[0]

    class OneBatch:            # a "dataloader" that yields the golden sequences as one batch
        batch_size = n
        dataset = type("D", (), {"data": None})()

        def __iter__(self):
            yield {"lr_imgs": fx["inputs"], "hr_imgs": fx["targets"], "index": torch.arange(n)}

        def __len__(self):
            return 1

    pred = VSRPredictor("cpu", OneBatch(), net, [torch.nn.L1Loss()], [1.0], [PSNR(), SSIM()], saved_dir=str(tmp_path),
                        exported=True, dataset="acdc")
    log = pred.predict()
    outs = restated.drfnet_forward(fx["inputs"], dict(fx["state_dict"]), fx["kwargs"]["upscale_factor"])
    T = len(outs)
    l1 = [float(restated.l1_loss(o, t)) for o, t in zip(outs, fx["targets"])]
    den = lambda x: restated.denormalize(x.detach(), "acdc")
    ps = [restated.psnr(den(o), den(t), size_average=False) for o, t in zip(outs, fx["targets"])]
    ss = [restated.ssim(den(o), den(t), size_average=False) for o, t in zip(outs, fx["targets"])]
    assert abs(log["L1Loss"] - sum(l1) / T) <= 1e-5 and abs(log["Loss"] - sum(l1) / T) <= 1e-5
    assert abs(log["PSNR"] - float(torch.stack(ps).mean())) <= 1e-3
    assert abs(log["SSIM"] - float(torch.stack(ss).mean())) <= 1e-4
    rows = list(csv.reader(open(tmp_path / "results.csv")))
    assert rows[0] == ["name", "PSNR", "SSIM", "L1Loss"] and len(rows) == 1 + n * T
    assert rows[1][0] == "slice00000_frame01"
    for i in range(n):
        for t in range(T):
            r = rows[1 + i * T + t]
            assert abs(float(r[1]) - float(ps[t][i])) <= 1e-3 and abs(float(r[2]) - float(ss[t][i])) <= 1e-4
            assert abs(float(r[3]) - l1[t]) <= 1e-5
