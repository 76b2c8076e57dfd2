"""Two-rank NCCL test on hardware (skipped with fewer than two GPUs): the data-parallel step of idee_b200.trainer.Trainer must equal
the mean of the single-rank ORACLE steps on the two batch shards, and both ranks must hold bit-identical
parameters afterwards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from idee_b200 import _lib
        from idee_b200.config import default_config
        from idee_b200.models.build import VQ_model
        from idee_b200.trainer import Trainer, shard_batch
        from oracle import idee_oracle as O                      # the checker
        _lib.set_precision("fp32")
        cfg = O.OracleConfig(in_vars=2, in_chans=1)
        sd = O.make_state_dict(cfg, seed=5, kind="reference")
        x, m_ext, m_loss = O.make_inputs(cfg, 2 * world, 8, 12, 16, seed=5)
        # oracle: one single-rank step per shard, gradients averaged over the shards
        want, want_loss = None, []
        for r in range(world):
            ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            total, _ = O.train_step_loss(ref, *(shard_batch(t, r, world) for t in (x, m_ext, m_loss)), cfg)
            total.backward()
            g = {k: v.grad / world for k, v in ref.items()}
            want = g if want is None else {k: want[k] + g[k] for k in g}
            want_loss.append(float(total))
        torch.manual_seed(100 + rank)                            # different initial weights per rank: the broadcast must fix that
        model = VQ_model(default_config(in_channels_dynamic=2))
        if rank == 0:
            model.load_state_dict(sd, strict=False)
        model = model.cuda().train()
        tr = Trainer(model, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.003)        # broadcasts rank 0's parameters
        xs, es, ls = (shard_batch(t, rank, world).cuda() for t in (x, m_ext, m_loss))
        total, _ = tr.forward_backward(xs, es, ls)
        tr.reduce_gradients()
        named = dict(model.named_parameters())
        worst = 0.0
        for k, g in want.items():
            if float(g.abs().max()) > 1e-7:
                got = named[k].grad.detach().double().cpu()
                worst = max(worst, float((got - g.double()).abs().max() / g.double().abs().max()))
        loss_ok = abs(float(total) - want_loss[rank]) / abs(want_loss[rank]) < 1e-4
        tr.optimizer_step()                                      # (all-reduces again: harmless, gradients are already equal)
        for _ in range(2):
            tr.step(xs, es, ls)
        flat = tr.flat_params.clone()
        ref0 = flat.clone()
        dist.broadcast(ref0, src=0)
        out[rank] = (worst, loss_ok, bool(torch.equal(flat, ref0)))
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_step_equals_mean_of_shard_oracle_steps():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world
        for rank in range(world):
            worst, loss_ok, same = out[rank]
            assert worst < 5e-4, (rank, worst)
            assert loss_ok and same, (rank, out[rank])
