"""Multi-rank single box: W ranks, each sweeping its slab of the cell grid (1/W of every colour's active
cells) and pushing accepted moves into the peers' memory through CUDA-IPC peer mappings, must reproduce the
single-rank run BIT FOR BIT (same-colour cells never interact, RNG is keyed by cell).  Runs with 2 processes; on a
single-GPU box both ranks share cuda:0 (peer memory over IPC works the same, only slower)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N, SWEEPS = 8192, 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(ctx, n):
    ctx.init_energy()
    ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
    ctx.seed(7)
    ctx.run(n)
    pos, sp = ctx.download()
    calls, acc = ctx.counters()
    return pos[0], float(ctx.energy()[0]), int(acc[0, 0])


def _make_ctx(device):
    from particlesmc_b200 import _lib as L
    from particlesmc_b200 import models as M
    from particlesmc_b200.device import DeviceContext
    from particlesmc_b200.synthetic import ka_lattice

    pos, sp, box = ka_lattice(N, 1.2, seed=9)
    ctx = DeviceContext(1, N, 3, 2, M.MODEL_LJ, mode=L.MODE_BOX, device=device)
    ctx.set_model(M.flatten_model_matrix(M.KobAndersen()))
    ctx.upload(pos, sp, box, 1.0)
    return ctx


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from particlesmc_b200.sharding import attach_box_peers

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    device = rank % torch.cuda.device_count()
    ctx = _make_ctx(device)
    attach_box_peers(ctx)
    pos, e, acc = _run(ctx, SWEEPS * N)
    dist.barrier()
    q.put((rank, pos, e, acc))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_two_ranks_reproduce_one_rank_bitwise():
    ctx = _make_ctx(0)
    ref_pos, ref_e, ref_acc = _run(ctx, SWEEPS * N)
    ctx.close()
    world = 2
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, pos, e, acc in out:
        assert np.array_equal(pos, ref_pos), f"rank {rank}: positions differ from the single-rank run"
        assert e == ref_e and acc == ref_acc
    assert 0.1 < ref_acc / (SWEEPS * N) < 0.9
