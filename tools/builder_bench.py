"""Timing of vrec_build_rating_vectors on the KNN bench workload's visits (one row per visit)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import vrec  # noqa: E402
from vrec import builders, synth  # noqa: E402

ctx = vrec.Context(0)
v, _ = synth.g2_place_visits(1_000_000, 100_000)
rng = np.random.default_rng(9)
perm = rng.permutation(int(v.count.sum()))
pe = np.repeat(v.person_id, v.count)[perm]
pl = np.repeat(v.place_id, v.count)[perm]
builders.build_rating_vectors(pe[:100000], pl[:100000], 100, ctx=ctx)
for i in range(6):
    t0 = time.perf_counter()
    out = builders.build_rating_vectors(pe, pl, 100, ctx=ctx)
    dt = time.perf_counter() - t0
    print(f"run {i}: {dt * 1e3:.1f} ms  {len(pe) / dt / 1e6:.1f} M visits/s", flush=True)
