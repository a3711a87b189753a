#!/usr/bin/env python
"""Latency of ONE stochastic-recommender query (BASELINE config 2) with and without the CUDA-graph while loop."""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import numpy as np  # noqa: E402

import vrec  # noqa: E402
from vrec import synth  # noqa: E402

persons = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pl = synth.sample_places(30000, seed=0)
v = synth.sample_place_visits(pl, 0, persons_per_region=persons, person_count_total=3 * persons, seed=0)
s, t, w = synth.build_stochastic_graph(v)
ctx = vrec.Context(0)
g = vrec.StochasticGraph(s, t, w, ctx=ctx)
flt = pl.of_region(0)
qs = np.unique(v.person_id)[::997][:40]
for eps, max_it in ((0.01, 20), (1e-9, 20)):
    for graph in (1, 0):
        g.set_option("graph", graph)
        rec = vrec.StochasticRecommender(g, eps, max_it)
        rec.recommend([int(qs[0])], flt, 10)
        lat, its = [], 0
        l0 = ctx.launch_count
        for q in qs:
            t0 = time.perf_counter()
            out = rec.recommend([int(q)], flt, 10)
            lat.append((time.perf_counter() - t0) * 1e3)
            its = int(out[3][0])
        print(f"N={g.N} nnz={g.nnz} eps={eps} maxIt={max_it} graph={graph}: median {statistics.median(lat):.3f} ms, "
              f"{its} iterations, {1e3 * statistics.median(lat) / (its + 1):.0f} us per sweep, "
              f"{(ctx.launch_count - l0) / len(qs):.1f} launches per query", flush=True)
