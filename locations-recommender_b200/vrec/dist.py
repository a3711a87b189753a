"""Host-side helpers of the one-process-per-GPU layout (torch.distributed is plumbing only).

Region-sharded work needs no collective: `shard_range` / `assign_units` say which units a rank owns and
`gather_rows` puts the per-rank result rows back in order on rank 0.  The NCCL unique id of the
row-partitioned graph path travels over the existing process group with `broadcast_bytes`."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of n units owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Row block of a device-generated partitioned graph (vrec_sg_generate): equal slices of ceil(n / world)
    rows -- every generated row has the same number of in-edges.  Loaded graphs are cut by in-edges instead
    (engine.host_sg_partition / vrec_sg_row_range)."""
    s = (int(n) + world - 1) // world
    lo = min(n, s * rank)
    return lo, min(n, lo + s)


def assign_units(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of independent units (region-sets, graphs) to ranks."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(u) for u in out]


def broadcast_bytes(data: bytes | None, n: int, src: int = 0) -> bytes:
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def gather_rows(local: np.ndarray, dst: int = 0):
    """Concatenates per-rank row blocks (rank order) on `dst`; other ranks get None."""
    import torch.distributed as dist
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, parts, dst=dst)
    if parts is None:
        return None
    return np.concatenate(parts, axis=0)


def init_comm(ctx) -> None:
    """Sets up the library's NCCL communicator over an initialised torch.distributed group."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = broadcast_bytes(ctx.unique_id() if rank == 0 else None, 128, 0)
    ctx.init_comm(rank, world, uid)
