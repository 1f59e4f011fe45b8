"""N > 1 host path on CPU: member sharding and the only collective of the design (all-gather of the
per-step loss scalars), with world_size-2 gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jsrl_corl_b200.ensemble import shard_members


def test_shard_members_partitions_exactly():
    for n, w in ((64, 8), (256, 8), (10, 4), (3, 8), (1, 1), (65, 2)):
        seen = []
        for r in range(w):
            seen += list(shard_members(n, w, r))
        assert seen == list(range(n))
        sizes = [len(shard_members(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_members, k_steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_members(n_members, world, rank)
    # stand-in for the engine output: loss[m, k, j] is a pure function of the GLOBAL member id, as it is on
    # the GPU (members are independent, so a member's trajectory does not depend on which rank holds it)
    local = torch.tensor([[[m * 1000.0 + k * 10.0 + j for j in range(3)] for k in range(k_steps)] for m in mine])
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction of bench.py
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.cat(gathered).numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_loss_allgather_world2_gloo(tmp_path):
    n_members, k_steps, world = 6, 4, 2
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, n_members, k_steps, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    want = np.array([[[m * 1000.0 + k * 10.0 + j for j in range(3)] for k in range(k_steps)] for m in range(n_members)])
    assert np.array_equal(got, want)
    assert np.load(tmp_path / "tmax.npy")[0] == world
