"""bench.py's parity helpers on the CPU: the checker must report bit-exact for identical answers and count
every target whose ids, rating bits, count or status differ (the engine is replaced by the oracle itself here)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "locations-recommender_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def test_knn_parity_helpers(oracle):
    import bench
    from vrec import synth
    v, places = synth.g2_place_visits(4000, 400, seed=20181231, region=0)
    inp = synth.build_rating_vectors(v)
    args = types.SimpleNamespace(k_nearest=50, max_recs=10)
    d = bench.oracle_knn_data(oracle, inp)

    class OracleAsEngine:
        def recommend(self, targets, flt, m):
            rc, *a = oracle.knn_query_batch(d, targets, 0.5, 0.5, 50, flt, m)
            assert rc == 0
            return a

    ids = inp.person_id[::40][:64]
    rc, *g = oracle.knn_query_batch(d, ids, 0.5, 0.5, 50, places.id, 10)
    assert rc == 0
    par = {"rec": OracleAsEngine(), "step": (ids, tuple(g))}
    cb, _ = bench.cpu_knn(args, inp, places, 32, check=par, d=d)
    assert cb["kind"] == "port" and cb["value"] > 0
    assert par["bit_exact"] and par["mismatching_targets"] == 0 and par["targets"] == 32
    assert par["timed_step"] == {"targets": len(ids), "mismatching_targets": 0, "bit_exact": True}
    # one rating off by one ulp, one place id changed, one count changed: three targets flagged
    g2 = [a.copy() for a in g]
    g2[1][3, 0] = np.nextafter(g2[1][3, 0], 0.0)
    g2[0][5, 1] += 1
    g2[2][7] -= 1
    assert bench.knn_count_bad(tuple(g2), g) == 3
    # no timed-step sample: the line still carries the main verdict and says why the other is missing
    par = {"rec": OracleAsEngine()}
    bench.cpu_knn(args, inp, places, 8, check=par, d=d)
    assert par["bit_exact"] and "error" in par["timed_step"]
