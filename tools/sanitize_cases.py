#!/usr/bin/env python
"""Reduced-size runs of every kernel family for compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
  compute-sanitizer --tool racecheck python tools/sanitize_cases.py
Every case is also compared with the oracle, so a run under the sanitizer is a parity run as well.  One tool per
gpurun call (B200_PROFILING.md)."""
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

small = "--small" in sys.argv
ctx = vrec.Context(0)
bad = 0


def check(name, ok):
    global bad
    print(f"{name}: {'ok' if ok else 'MISMATCH'}", flush=True)
    bad += not ok


# ---- KNN: every kernel variant, K regimes, a grid of more than one wave of CTAs
v, places = synth.g2_place_visits(6000 if small else 20000, 800 if small else 2000, seed=20181231)
inp = synth.build_rating_vectors(v)
d = oracle_knn_data(oracle, inp)
rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
flt = np.ascontiguousarray(places.id, dtype=np.int64)
for kernel, name in ((1, "exact scan"), (2, "fp32 tile"), (3, "tcgen05"), (4, "tcgen05 warp-specialised")):
    rs.set_option("knn_kernel", kernel)
    for K, nt in ((7, 40), (50, 300), (1024 if kernel <= 2 else 56, 24)):
        targets = inp.person_id[:: max(1, len(inp.person_id) // nt)][:nt]
        rc, opl, ort, ocnt, ost = oracle.knn_query_batch(d, targets, 0.5, 0.5, K, flt, 10)
        pl, rt, cnt, st = vrec.KnnRecommender(rs, 0.5, 0.5, K).recommend(targets, flt, 10)
        check(f"knn {name} K={K} targets={nt}",
              rc == 0 and np.array_equal(pl, opl) and np.array_equal(rt.view(np.int64), ort.view(np.int64))
              and cnt.tolist() == ocnt.tolist() and st.tolist() == ost.tolist())
rs.set_option("knn_kernel", 0)
# more than one wave of CTAs of the warp-specialised kernel: 8 target tiles x 32 candidate splits = 256 CTAs
targets = inp.person_id[:1024]
rs.set_option("splits", 32)
rc, opl, ort, ocnt, ost = oracle.knn_query_batch(d, targets, 0.5, 0.5, 50, flt, 10)
pl, rt, cnt, st = vrec.KnnRecommender(rs, 0.5, 0.5, 50).recommend(targets, flt, 10)
check("knn multi-wave grid (1024 targets x 32 splits)", np.array_equal(pl, opl) and np.array_equal(rt.view(np.int64), ort.view(np.int64)))
rs.set_option("splits", 0)
rs.set_option("compact_records", 0)
pl, rt, cnt, st = vrec.KnnRecommender(rs, 0.5, 0.5, 50).recommend(targets[:256], flt, 10)
check("knn full-size records", np.array_equal(pl, opl[:256]) and np.array_equal(rt.view(np.int64), ort[:256].view(np.int64)))
rs.set_option("compact_records", 1)
for K in (3000, 2_000_000):                    # similarity rows + radix select + column scan
    targets = inp.person_id[::3000][:6]
    rc, opl, ort, ocnt, ost = oracle.knn_query_batch(d, targets, 0.5, 0.5, K, flt, 10)
    pl, rt, cnt, st = vrec.KnnRecommender(rs, 0.5, 0.5, K).recommend(targets, flt, 10)
    check(f"knn large K={K}", np.array_equal(pl, opl) and np.array_equal(rt.view(np.int64), ort.view(np.int64)))
rs.close()

# ---- SG: per-query sweeps (long rows, CUDA-graph loop and queued loop), batch kernel, row-partitioned group
s, t, w, persons, pls, cats = synth.random_layered_graph(20, 300 if small else 1000, 3000 if small else 20000, seed=4,
                                                         places_per_person=2, cats_per_person=2, similar_per_place=20,
                                                         hub_places=10, hub_fraction=0.2, duplicate_fraction=0.05)
og = oracle.SgGraph(s, t, w)
g = vrec.StochasticGraph(s, t, w, ctx=ctx)
for graph in (1, 0):
    g.set_option("graph", graph)
    rec = vrec.StochasticRecommender(g, 1e-3, 20)
    ok = True
    for vtx in (int(persons[0]), int(pls[3]), int(cats[1])):
        x = rec.stationary(vtx)
        rc, ox, oit, oconv, _ = og.run(vtx, 1e-3, 20)
        ok = ok and np.array_equal(x, ox) and (rec.last_iterations, rec.last_converged) == (oit, oconv)
    check(f"sg per-query sweeps graph={graph}", ok)
g.set_option("batch", 2)
rec = vrec.StochasticRecommender(g, 0.01, 20)
q = persons[:: max(1, len(persons) // 200)][:200]
oi, op, cnt, its, conv, st = rec.recommend(q, pls, 10)
ok = g.batch_info(0) == len(q)
for r in range(0, len(q), 20):
    rc, wi, wp, oit, oconv = og.query(int(q[r]), 0.01, 20, pls, 10)
    ok = ok and oi[r, :cnt[r]].tolist() == wi.tolist() and op[r, :cnt[r]].tolist() == wp.tolist() and int(its[r]) == oit
check("sg batch kernel (200 persons)", ok)
g.close()
grp = vrec.StochasticGraphGroup(s, t, w, 3, ctx=ctx)
xs, its, conv, _ = grp.stationary(int(persons[5]), 1e-3, 20)
rc, ox, oit, oconv, _ = og.run(int(persons[5]), 1e-3, 20)
check("sg row-partitioned group (3 parts)", all(np.array_equal(xs[r], ox) and (its[r], conv[r]) == (oit, oconv) for r in range(3)))
grp.close()
gg = vrec.StochasticGraph.generate(200_000 if small else 7_000_000, 4, seed=5, ctx=ctx)     # two source blocks at 7 M
gg.iterate_device(2)
ctx.synchronize()
check("sg generated graph sweeps", True)
gg.close()
print("SANITIZE CASES:", "all ok" if bad == 0 else f"{bad} MISMATCH", flush=True)
sys.exit(1 if bad else 0)
