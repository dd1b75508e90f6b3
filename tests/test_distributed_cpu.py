"""CPU / gloo, world_size 2: the host logic of the data-parallel path (group sharding + gradient mean all-reduce)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from da_sac_b200 import synth
    from da_sac_b200.trainer import shard_batch, shard_groups, allreduce_mean_
    G, K = 4, 3
    batch = synth.make_target_batch(G, K, (32, 32), seed=0)
    mine = shard_batch(batch, K, world, rank)
    ok = mine[0].shape[0] == G * K // world and shard_groups(G, world, rank) == list(range(rank * 2, rank * 2 + 2))
    ok = ok and torch.equal(mine[3], batch[3][rank * 6:(rank + 1) * 6])
    g = torch.full((1000,), float(rank + 1))
    allreduce_mean_(g)                                   # DDP semantics: sum / world
    ok = ok and torch.allclose(g, torch.full((1000,), 1.5))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_groups_rejects_uneven_split():
    import pytest
    from da_sac_b200.trainer import shard_groups
    with pytest.raises(ValueError):
        shard_groups(3, 2, 0)
