"""world_size-2 gloo tests of the host-side multi-GPU logic (no GPU needed)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))


def test_shard_ranges_cover():
    from vrec.dist import assign_units, shard_range, slice_rows
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
            s = [slice_rows(n, k, world) for k in range(world)]
            assert s[0][0] == 0 and max(b for _, b in s) == n
            assert all(b - a <= (n + world - 1) // world for a, b in s)
    units = assign_units([5, 1, 1, 1, 4, 4], 2)       # 3 single + 3 pairwise graphs over 2 ranks
    assert sorted(units[0] + units[1]) == list(range(6))
    assert abs(sum([5, 1, 1, 1, 4, 4][i] for i in units[0]) - 8) <= 1


def test_partition_bounds_balance_in_edges():
    """Row ranges of a row-partitioned graph: contiguous, covering, balanced by in-edges."""
    import numpy as np
    from vrec.engine import host_sg_partition
    rng = np.random.default_rng(11)
    for n, world in ((1, 2), (10, 3), (5000, 2), (5000, 8), (40000, 7)):
        deg = rng.integers(0, 40, size=n)
        deg[rng.integers(0, n, size=max(1, n // 100))] += 2000          # hubs
        rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        b = host_sg_partition(rowptr, world)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        nnz = int(rowptr[-1])
        per = np.diff(rowptr[b].astype(np.int64))
        if nnz and n >= 50 * world:
            assert per.max() <= nnz / world + deg.max() + 1                # no rank exceeds its share by a row
    # no in-edges at all: every rank still gets a well-formed (possibly empty) range
    b = host_sg_partition(np.zeros(6, dtype=np.int32), 4)
    assert b[0] == 0 and b[-1] == 5 and np.all(np.diff(b) >= 0)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vrec.dist import broadcast_bytes, gather_rows, shard_range
    uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, 0)
    lo, hi = shard_range(11, rank, world)
    local = np.arange(lo, hi)[:, None] * np.ones((1, 3), dtype=np.int64)
    full = gather_rows(local, 0)
    q.put((rank, uid == bytes(range(128)), None if full is None else full[:, 0].tolist()))
    dist.destroy_process_group()


def test_gloo_world2_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, True, list(range(11)))
    assert res[1] == (1, True, None)
