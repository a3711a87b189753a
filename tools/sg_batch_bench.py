"""Throughput of the SG batch kernel on a per-region graph of the default-data shape (config 4).

usage: python tools/sg_batch_bench.py [n_persons] [n_places] [n_queries] [tpc]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
sys.path.insert(0, ROOT)
import vrec  # noqa: E402
from vrec import synth  # noqa: E402

n_persons = int(sys.argv[1]) if len(sys.argv) > 1 else 770_000
n_places = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
n_q = int(sys.argv[3]) if len(sys.argv) > 3 else 20_000
tpc = int(sys.argv[4]) if len(sys.argv) > 4 else 0

t0 = time.time()
s, t, w, persons, places, cats = synth.random_layered_graph(
    20, n_places, n_persons, seed=4, places_per_person=2, cats_per_person=2, similar_per_place=50,
    hub_places=20, hub_fraction=0.1, duplicate_fraction=0.0)
print(f"graph: {len(s)} edges built in {time.time() - t0:.1f}s", flush=True)
ctx = vrec.Context(0)
t0 = time.time()
g = vrec.StochasticGraph(s, t, w, ctx=ctx)
print(f"load {time.time() - t0:.2f}s  N={g.N} nnz={g.nnz} batch_ok={g.batch_info(1)} n_active={g.batch_info(2)} "
      f"r_nnz={g.batch_info(3)} sell_nnz={g.batch_info(4)}", flush=True)
g.set_option("batch_targets_per_cta", tpc)
rng = np.random.default_rng(0)
q = rng.choice(persons, n_q, replace=False)
for eps, max_it in [(0.01, 1), (0.01, 20), (0.0, 20), (1e-4, 20)]:
    rec = vrec.StochasticRecommender(g, eps, max_it)
    rec.recommend(q[:2000], places, 10)      # warm-up (x1, allocations)
    t0 = time.time()
    oi, op, cnt, its, conv, st = rec.recommend(q, places, 10)
    dt = time.time() - t0
    assert g.batch_info(0) == n_q
    tot_it = int((its - 1 + conv).sum())      # SpMV passes after the shared first one
    print(f"eps={eps} max_it={max_it}: {n_q / dt:.0f} persons/s  ({dt * 1e3:.1f} ms, mean iterations {its.mean():.2f}, "
          f"{tot_it * g.batch_info(3) / dt / 1e9:.1f} G edge-terms/s; us prepare/kernel/results "
          f"{g.batch_info(5)}/{g.batch_info(6)}/{g.batch_info(7)})", flush=True)
rec = vrec.StochasticRecommender(g, 0.01, 1)
for nq in (100, 2000, 20000):
    t0 = time.time()
    rec.recommend(q[:nq], places, 10)
    print(f"max_it=1 n_q={nq}: {(time.time() - t0) * 1e3:.2f} ms  (us prepare/kernel/results: "
          f"{g.batch_info(5)}/{g.batch_info(6)}/{g.batch_info(7)})", flush=True)
# per-query path for comparison
g.set_option("batch", 0)
rec = vrec.StochasticRecommender(g, 0.01, 20)
t0 = time.time()
rec.recommend(q[:200], places, 10)
dt = time.time() - t0
print(f"per-query kernels: {200 / dt:.0f} persons/s", flush=True)
