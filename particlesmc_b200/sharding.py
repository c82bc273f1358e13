"""Chain sharding across GPUs: independent chains split by contiguous global index ranges.

Rank r of W holds chains [offset, offset + count) and creates its device context with ``chain_offset = offset``:
every chain's Philox stream is keyed by its GLOBAL index, so the union of all shards is bit-identical to a
single-GPU run (tests/test_gpu_parity.py::test_sharded_chains_equal_unsharded).  There is no data-path
collective; ``gather_chain_values`` collects per-chain scalars (energies, counters) at schedule points.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_chains: int, rank: int, world: int) -> Tuple[int, int]:
    """(offset, count) of rank's shard: sizes differ by at most one, lower ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_chains, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def gather_chain_values(local: np.ndarray, n_chains: int):
    """All-gather per-chain values of every rank's shard into the global order (torch.distributed, any backend)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    counts = [shard_range(n_chains, r, world)[1] for r in range(world)]
    width = max(counts)
    local = np.asarray(local, dtype=np.float64)
    buf = torch.zeros((width,) + local.shape[1:], dtype=torch.float64, device=dev)
    buf[: counts[rank]] = torch.from_numpy(local).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def attach_box_peers(ctx) -> None:
    """Multi-GPU single box: all-gather the 64-byte CUDA IPC handles of every rank's shared block (host channel:
    torch.distributed, any backend) and attach them.  Call on every rank after ``ctx.upload``; collective."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.from_numpy(ctx.box_peer_handle().copy())
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        out = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(out, mine.to(dev))
        handles = torch.stack(out).cpu().numpy()
    else:
        out = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(out, mine)
        handles = torch.stack(out).numpy()
    ctx.box_peer_attach(rank, world, handles)
    dist.barrier()


def allreduce_sum(values: np.ndarray) -> np.ndarray:
    """Sum a small numpy array over all ranks (observable reductions: histograms, acceptance counters)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(values)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
