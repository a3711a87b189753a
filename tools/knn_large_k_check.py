#!/usr/bin/env python
"""Quick GPU check of the fused top-K path for the largest K it serves (K <= 1024): engine == oracle, bit for bit.
No torch import (fast start).  python tools/knn_large_k_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "locations-recommender_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import vrec  # noqa: E402
from oracle import oracle  # noqa: E402
from vrec import synth  # noqa: E402
from helpers import oracle_knn_data  # noqa: E402

v, places = synth.g2_place_visits(20000, 2000, seed=20181231)
inp = synth.build_rating_vectors(v)
ctx = vrec.Context(0)
rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
d = oracle_knn_data(oracle, inp)
targets = inp.person_id[::700]
bad = 0
for K in (923, 1000, 1024):
    rc, opl, ort, ocnt, ost = oracle.knn_query_batch(d, targets, 0.5, 0.5, K, places.id, 10)
    pl, rt, cnt, st = vrec.KnnRecommender(rs, 0.5, 0.5, K).recommend(targets, places.id, 10)
    ok = (rc == 0 and np.array_equal(pl, opl) and np.array_equal(rt.view(np.int64), ort.view(np.int64))
          and cnt.tolist() == ocnt.tolist() and st.tolist() == ost.tolist())
    print(f"K={K}: {len(targets)} targets {'bit-exact' if ok else 'MISMATCH'}", flush=True)
    bad += not ok
sys.exit(1 if bad else 0)
