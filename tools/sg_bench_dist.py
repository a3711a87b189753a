#!/usr/bin/env python
"""Row-partitioned SG power iteration on the device-generated graph, one process per GPU (torchrun):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sg_bench_dist.py
Prints us/iteration (max over ranks) and the algorithmic GB/s of the whole graph."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vrec  # noqa: E402
from vrec import dist as vdist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--deg", type=int, default=100)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = vrec.Context(local)
vdist.init_comm(ctx)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
g = vrec.StochasticGraph.generate(a.n, a.deg, seed=5, rank=rank, world=world, ctx=ctx)
ctx.synchronize()
bytes_it = 12 * a.n * a.deg + 20 * a.n
g.iterate_device(a.iters)
ctx.synchronize()
for s in range(a.steps):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.iterate_device(a.iters)
    e1.record(stream)
    ctx.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(f"world={world} N={a.n} nnz={a.n * a.deg}: {ms / a.iters * 1e3:.0f} us/iter  "
              f"{bytes_it * a.iters / ms / 1e6:.0f} GB/s algorithmic over all ranks", flush=True)
g.close()
dist.destroy_process_group()
