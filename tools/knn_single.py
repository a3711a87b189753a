"""One-person KNN queries (the launcher's REPL case): python tools/knn_single.py [K]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import vrec  # noqa: E402
from vrec import synth  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
v, places = synth.g2_place_visits(1_000_000, 100_000)
inp = synth.build_rating_vectors(v)
ctx = vrec.Context(0)
rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
if len(sys.argv) > 2:
    rs.set_option("knn_kernel", int(sys.argv[2]))
rec = vrec.KnnRecommender(rs, 0.5, 0.5, K)
flt = np.ascontiguousarray(places.id)
for i in range(6):
    t0 = time.perf_counter()
    rec.recommend([int(inp.person_id[5 + 7 * i])], flt, 10)
    print(f"query {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
