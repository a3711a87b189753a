#!/usr/bin/env python
"""torchrun-launched check of the row-partitioned SG path: every rank's result must equal the
single-GPU engine and the CPU oracle bit for bit.  Usage:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_check_2gpu.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "locations-recommender_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vrec  # noqa: E402
from oracle import oracle  # noqa: E402
from vrec import dist as vdist, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = vrec.Context(local)
vdist.init_comm(ctx)
assert (ctx.rank, ctx.world) == (rank, world)
ok = True
for (n, deg, hub, seed) in [(5000, 5, 0.5, 3), (30001, 20, 0.2, 4)]:
    s, t, w = synth.random_stochastic_graph(n, deg, seed=seed, hub_fraction=hub)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx, partitioned=True)
    og = oracle.SgGraph(s, t, w)
    for vertex, eps, max_it in [(int(og.ids[0]), 0.0, 6), (int(og.ids[n // 2]), 1e-3, 20)]:
        rec = vrec.StochasticRecommender(g, eps, max_it)
        x = rec.stationary(vertex)
        rc, ox, oit, oconv, _ = og.run(vertex, eps, max_it)
        good = np.array_equal(x, ox) and (rec.last_iterations, rec.last_converged) == (oit, oconv)
        ids, pr, cnt, *_ = rec.recommend([vertex], og.ids[::3], 10)
        rc, wi, wp, _, _ = og.query(vertex, eps, max_it, og.ids[::3], 10)
        good = good and ids[0, :cnt[0]].tolist() == wi.tolist() and pr[0, :cnt[0]].tolist() == wp.tolist()
        print(f"[rank {rank}] N={n} vertex={vertex} eps={eps}: {'ok' if good else 'MISMATCH'} "
              f"(iterations {rec.last_iterations}, converged {rec.last_converged})", flush=True)
        ok = ok and good
    lo, hi = g.row_range()
    print(f"[rank {rank}] N={n}: rows [{lo}, {hi}) with {g.nnz} of {len(s)} in-edges", flush=True)
    g.close()
# generated graphs (one and two source blocks): every rank's partitioned result equals the unpartitioned
# engine on the same device, which the GPU suite checks against the oracle
for (n, deg, its) in [(40000, 16, 5), (7_000_000, 8, 3)]:
    gg = vrec.StochasticGraph.generate(n, deg, seed=5, rank=rank, world=world, ctx=ctx)
    rec = vrec.StochasticRecommender(gg, 1e-7, its)
    x = rec.stationary(0)
    g1 = vrec.StochasticGraph.generate(n, deg, seed=5, ctx=ctx)
    rec1 = vrec.StochasticRecommender(g1, 1e-7, its)
    x1 = rec1.stationary(0)
    same = np.array_equal(x, x1) and (rec.last_iterations, rec.last_converged) == (rec1.last_iterations,
                                                                                  rec1.last_converged)
    print(f"[rank {rank}] generated graph N={n}: partitioned == unpartitioned: {same}", flush=True)
    ok = ok and same
    gg.close()
    g1.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
