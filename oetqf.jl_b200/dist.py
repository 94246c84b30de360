"""Host-side plumbing for one-rank-per-GPU runs: shard arithmetic and the handle exchange.

torch.distributed is used only as the rendezvous/all-gather runtime (any backend: nccl on the GPU box,
gloo in the CPU tests); the data path itself is peer memory inside the kernels (csrc/comm.cu)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of n units for `rank`, in rank order, sizes rounded up to `align`
    (fault rows use 4 = the matvec's row-block size).  Trailing ranks may be empty."""
    assert world >= 1 and 0 <= rank < world and align >= 1
    per = -(-n // world)
    per = -(-per // align) * align
    return min(n, rank * per), min(n, (rank + 1) * per)


def all_shards(n: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    return [shard_range(n, world, r, align) for r in range(world)]


def exchange_handles(mine: bytes, group=None) -> List[bytes]:
    """All-gather the opaque window handles (oq_comm_export) in rank order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out: List[bytes] = [b""] * world
    dist.all_gather_object(out, mine, group=group)
    return out


def connect(problem, group=None) -> None:
    """Export this rank's window, exchange, map the peers (DeviceProblem.comm_export / comm_connect)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    problem.comm_connect(exchange_handles(problem.comm_export(rank, world), group))
    dist.barrier(group)


def local_state(parts: Sequence, fault_rows: Tuple[int, int], mantle_elems: Tuple[int, int] = (0, 0), kind="fault"):
    """This rank's slices of the reference-layout state partitions.
    kind = "fault": (v, θ, δ[, 𝓅]); "viscoelastic": (v, θ, ϵ, σ, δ)."""
    import numpy as np
    f0, f1 = fault_rows
    e0, e1 = mantle_elems

    def fl(x):
        return np.ascontiguousarray(np.asarray(x).reshape(-1, order="F")[f0:f1])

    def ml(x):
        return np.ascontiguousarray(np.asarray(x)[e0:e1, :].reshape(-1, order="F"))

    if kind == "viscoelastic":
        v, th, eps, sg, dl = parts
        return [fl(v), fl(th), ml(eps), ml(sg), fl(dl)]
    return [fl(x) for x in parts]
