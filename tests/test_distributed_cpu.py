"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: flat parameter/gradient buffers, rank-0 weight broadcast,
the single gradient all-reduce (average) and batch sharding.  The CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from idee_b200.config import default_config
        from idee_b200.models.build import VQ_model
        from idee_b200.trainer import Trainer, shard_batch
        torch.manual_seed(rank)                      # deliberately different initial weights per rank
        model = VQ_model(default_config(in_channels_dynamic=2))
        tr = Trainer(model)                          # broadcasts rank 0's flat parameters
        flat0 = tr.flat_params.clone()
        gathered = [torch.empty_like(flat0) for _ in range(world)]
        dist.all_gather(gathered, flat0)
        same_weights = all(torch.equal(gathered[0], g) for g in gathered)
        # every parameter and gradient is a view of the flat buffers, state_dict is unchanged
        n_params = sum(p.numel() for p in model.parameters())
        views_ok = all(p.grad is not None and p.grad.shape == p.shape for p in model.parameters())
        tr._check_flat()
        # rank-dependent gradients -> one all-reduce -> mean over ranks, visible through p.grad
        tr.flat_grads.copy_(torch.arange(n_params, dtype=torch.float32) * (rank + 1))
        tr.reduce_gradients()
        expect = torch.arange(n_params, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        reduce_ok = torch.allclose(tr.flat_grads, expect)
        first = next(iter(model.parameters()))
        off = (first.data_ptr() - tr.flat_params.data_ptr()) // 4
        grad_view_ok = torch.allclose(first.grad.reshape(-1), expect[off:off + first.numel()])
        x = torch.arange(8 * 3).view(8, 3)
        shard = shard_batch(x, rank, world)
        shard_ok = shard.shape[0] == 8 // world and int(shard[0, 0]) == rank * (8 // world) * 3
        out[rank] = (same_weights, views_ok, reduce_ok, grad_view_ok, shard_ok, tr.flat_params.numel() == n_params)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_flat_allreduce_and_broadcast():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world
        for rank in range(world):
            assert all(out[rank]), (rank, out[rank])


def test_shard_batch_rejects_uneven():
    from idee_b200.trainer import shard_batch
    with pytest.raises(ValueError):
        shard_batch(torch.zeros(7, 2), 0, 2)
