"""Multi-process host logic of the sharded-chains path on CPU: world_size 2, gloo backend."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_chains, q):
    import torch.distributed as dist
    from particlesmc_b200.sharding import allreduce_sum, gather_chain_values, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = shard_range(n_chains, rank, world)
    # each rank "computes" a per-chain value that depends only on the GLOBAL chain id
    local = np.stack([np.arange(off, off + cnt, dtype=np.float64) * 1.5, np.full(cnt, float(rank))], axis=1)
    full = gather_chain_values(local, n_chains)
    # timing protocol of bench.py: max over ranks
    import torch
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hist = allreduce_sum(np.arange(5, dtype=np.float64) * (rank + 1))  # observable reduction (g(r) counts)
    assert np.array_equal(hist, np.arange(5) * 3.0)
    q.put((rank, off, cnt, full, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world, n_chains = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chains, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(o[1], o[2]) for o in out] == [(0, 4), (4, 3)]
    for rank, off, cnt, full, tmax in out:
        assert full.shape == (n_chains, 2)
        assert np.array_equal(full[:, 0], np.arange(n_chains) * 1.5)
        assert np.array_equal(full[:, 1], [0, 0, 0, 0, 1, 1, 1])
        assert tmax == 11.0
