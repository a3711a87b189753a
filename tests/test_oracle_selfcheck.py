"""The C oracle against an independent pure-Python restatement of the Scala source (small cases),
and structural checks of the canonical summation order."""
import sys

import numpy as np
import pytest

from tests.helpers import oracle_knn_data, py_knn

sys.path.insert(0, "locations-recommender_b200")
from vrec import synth  # noqa: E402


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("k", [1, 5, 1000])
def test_knn_oracle_vs_python(oracle, seed, k):
    inp = synth.random_knn_inputs(60, 15, 6, seed=seed, separate_ratings=(seed % 2 == 0))
    d = oracle_knn_data(oracle, inp)
    flt = np.arange(0, 15, 2)
    for target in list(inp.person_id[:12]) + [5]:
        want = py_knn(inp, int(target), 0.3, 0.7, k, flt, 4)
        rc, ids, sims = oracle.knn_neighbours(d, target, 0.3, 0.7, k)
        if want is None:
            assert rc == oracle.ENOENT
            continue
        nb, est, recs = want
        assert rc == 0
        got = sorted(zip(ids.tolist(), sims.tolist()))
        assert got == sorted((int(inp.person_id[i]), s) for s, i in nb)
        rc, pl, rt = oracle.knn_estimates(d, target, 0.3, 0.7, k)
        assert rc == 0 and dict(zip(pl.tolist(), rt.tolist())) == est
        rc, opl, ort, ocnt, ost = oracle.knn_query_batch(d, [target], 0.3, 0.7, k, flt, 4)
        assert rc == 0 and ost[0] == 0
        assert list(zip(ort[0, :ocnt[0]].tolist(), opl[0, :ocnt[0]].tolist())) == recs


def test_knn_oracle_requires(oracle):
    inp = synth.random_knn_inputs(10, 5, 3, seed=9)
    d = oracle_knn_data(oracle, inp)
    t = inp.person_id[0]
    assert oracle.knn_neighbours(d, t, 0.0, 1.0, 3)[0] == oracle.EINVAL
    assert oracle.knn_neighbours(d, t, 0.6, 0.5, 3)[0] == oracle.EINVAL
    assert oracle.knn_neighbours(d, t, 0.5, 0.5, 0)[0] == oracle.EINVAL


def test_sg_oracle_vs_sequential_python(oracle):
    # rows with <= 3 in-edges: canonical order == left-to-right; longer rows: within 1e-12
    s, t, w = synth.random_stochastic_graph(40, 3, seed=4)
    g = oracle.SgGraph(s, t, w)
    ids = g.ids
    idx = {int(v): i for i, v in enumerate(ids)}
    order = sorted(range(len(s)), key=lambda e: (idx[int(t[e])], idx[int(s[e])], e))
    x = np.full(g.N, 1.0 / g.N)
    v = idx[int(ids[3])]
    for _ in range(5):
        sig = np.zeros(g.N)
        for e in order:
            sig[idx[int(t[e])]] += x[idx[int(s[e])]] * w[e]
        x = np.array([(1.0 if i == v else 0.0) * 0.15 + sig[i] * (1 - 0.15) for i in range(g.N)])
    rc, ox, it, conv, res = g.run(int(ids[3]), 0.0, 5)
    assert rc == 0 and it == 5 and conv == 0
    np.testing.assert_allclose(ox, x, rtol=1e-12, atol=0)


def test_sg_long_rows_segmenting(oracle):
    # a hub row longer than the 1024-term segment; mass must be conserved up to rounding
    s, t, w = synth.random_stochastic_graph(3000, 4, seed=5, hub_fraction=0.4)
    g = oracle.SgGraph(s, t, w)
    rc, x, it, conv, res = g.run(int(g.ids[10]), 1e-9, 50)
    assert rc == 0
    assert abs(x.sum() - 1.0) < 1e-9          # every vertex has out-edges: no mass is lost


def test_rating_vectors_builder_oracle_vs_numpy_restatement(oracle):
    """oracle/vro_build_rating_vectors against the numpy restatement of RatingsBuilder /
    RatingVectorsBuilder in vrec/synth.py (rank() keeps ties; indices ascending; size = max id + 1)."""
    v, places = synth.g2_place_visits(3000, 400, seed=7, mean_places=9.0)
    want = synth.build_rating_vectors(v, max_rated_places=5, max_rated_categories=3)
    for ent, top_n, rp, ci, cv_, dim in ((v.place_id, 5, want.place_rowptr, want.place_col, want.place_val, want.place_dim),
                                         (v.category_id, 3, want.cat_rowptr, want.cat_col, want.cat_val, want.cat_dim)):
        # pre-aggregated rows with weights, and the same visits expanded to one row per visit, shuffled
        rc, persons, rowptr, col, val, d = oracle.build_rating_vectors(v.person_id, ent, v.count, top_n)
        assert rc == 0 and d == dim
        assert np.array_equal(persons, want.person_id) and np.array_equal(rowptr, rp)
        assert np.array_equal(col, ci) and np.array_equal(val, cv_)
        rng = np.random.default_rng(1)
        perm = rng.permutation(int(v.count.sum()))
        pe = np.repeat(v.person_id, v.count)[perm]
        ee = np.repeat(ent, v.count)[perm]
        rc, persons2, rowptr2, col2, val2, d2 = oracle.build_rating_vectors(pe, ee, None, top_n)
        assert rc == 0 and d2 == dim and np.array_equal(persons2, persons) and np.array_equal(rowptr2, rowptr)
        assert np.array_equal(col2, col) and np.array_equal(val2, val)
    # ties at the cut all stay (rank(), knn/RatingsBuilder.scala:44-46); ids outside Int are an error (:36-41)
    rc, persons, rowptr, col, val, d = oracle.build_rating_vectors([5, 5, 5, 5, 9], [1, 2, 3, 4, 7], [3, 2, 2, 1, 1], 2)
    assert rc == 0 and persons.tolist() == [5, 9] and rowptr.tolist() == [0, 3, 4] and col.tolist() == [1, 2, 3, 7]
    assert val.tolist() == [3.0, 2.0, 2.0, 1.0] and d == 8
    rc, *_ = oracle.build_rating_vectors([1], [2 ** 31], None, 10)
    assert rc == oracle.EINVAL
    rc, persons, rowptr, *_ = oracle.build_rating_vectors([], [], None, 10)
    assert rc == 0 and len(persons) == 0 and rowptr.tolist() == [0]


def _expand_visits(v, rng, week_ms=6 * 24 * 3600 * 1000):
    """One row per visit (PlaceVisitCounts -> place_visits rows), shuffled, timestamps inside one window."""
    perm = rng.permutation(int(v.count.sum()))
    pe = np.repeat(v.person_id, v.count)[perm]
    pl = np.repeat(v.place_id, v.count)[perm]
    ca = np.repeat(v.category_id, v.count)[perm]
    ts = 1_546_300_800_000 + rng.integers(0, week_ms, len(pe))
    return pe, pl, ca, ts


def test_stochastic_graph_builder_oracle_vs_numpy_restatement(oracle):
    """vro_build_stochastic_graph against the numpy restatement of the four edge calculators in vrec/synth.py
    (every visit inside one 7-day window, so every pair of visits of a person counts)."""
    v, places = synth.g2_place_visits(400, 60, seed=5, mean_places=4.0)
    rng = np.random.default_rng(2)
    pe, pl, ca, ts = _expand_visits(v, rng)
    rc, s, t, w = oracle.build_stochastic_graph(pe, pl, ca, ts, 0.5, 0.5)
    assert rc == 0
    ws, wt, ww = synth.build_stochastic_graph(v, 0.5, 0.5)
    got = sorted(zip(s.tolist(), t.tolist(), w.tolist()))
    want = sorted(zip(ws.tolist(), wt.tolist(), ww.tolist()))
    assert [g[:2] for g in got] == [x[:2] for x in want]
    assert np.allclose([g[2] for g in got], [x[2] for x in want], rtol=1e-12, atol=0)
    # one family by hand: counts 3 / 2 / 2 / 1 of source 5, top 2 keeps the tie; weights = count / kept total * beta
    rc, s, t, w = oracle.build_edge_family([5, 5, 5, 5, 9], [1, 2, 3, 4, 7], [3, 2, 2, 1, 1], 2, 0.5)
    assert rc == 0 and s.tolist() == [5, 5, 5, 9] and t.tolist() == [1, 2, 3, 7]
    assert w.tolist() == [3 / 7 * 0.5, 2 / 7 * 0.5, 2 / 7 * 0.5, 1 / 1 * 0.5]
    # visits more than 7 days apart do not make places similar (PlaceSimilarPlace.scala:29-36)
    day = 24 * 3600 * 1000
    rc, s, t, w = oracle.build_stochastic_graph([1, 1, 1], [10, 11, 12], [0, 0, 1], [0, 3 * day, 11 * day], 0.5, 0.5)
    pairs = [(a, b) for a, b in zip(s.tolist(), t.tolist()) if a >= 10 and b >= 10 and a < 100]
    assert sorted(pairs) == [(10, 11), (11, 10)]


def _location_visits(rng, n_regions=2, places_per_region=400, n_visits=3000):
    """Places on a jittered grid (~150 m pitch) per region, visits scattered around random places of their region."""
    centres = [(48.85, 2.35), (59.93, 30.33), (-33.86, 151.2)][:n_regions]
    pid, plat, plon, pcat, preg = [], [], [], [], []
    for r, (la, lo) in enumerate(centres):
        side = int(np.sqrt(places_per_region))
        gy, gx = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
        plat.append(la + gy.ravel() * 0.00135 + rng.normal(0, 0.0002, side * side))
        plon.append(lo + gx.ravel() * 0.0021 + rng.normal(0, 0.0003, side * side))
        pid.append(40 + r * 100_000 + np.arange(side * side))
        pcat.append(rng.integers(0, 20, side * side))
        preg.append(np.full(side * side, r))
    pid, plat, plon, pcat, preg = map(np.concatenate, (pid, plat, plon, pcat, preg))
    pick = rng.integers(0, len(pid), n_visits)
    lat = plat[pick] + rng.normal(0, 0.0007, n_visits)
    lon = plon[pick] + rng.normal(0, 0.001, n_visits)
    person = 1_000_000 + rng.integers(0, 300, n_visits)
    ts = 1_546_300_800_000 + rng.integers(0, 30 * 24 * 3600 * 1000, n_visits)
    return (person, lat, lon, ts, preg[pick]), (pid, plat, plon, pcat, preg)


def test_place_visits_oracle(oracle):
    """vro_build_place_visits: the reference's cross-join row by row -- checked against numpy on the same formula."""
    rng = np.random.default_rng(6)
    (person, lat, lon, ts, reg), (pid, plat, plon, pcat, preg) = _location_visits(rng)
    rc, (op, ot, opl, org, oc), margin, miss = oracle.build_place_visits(person, lat, lon, ts, reg, pid, plat, plon, pcat, preg,
                                                                        7, 100.0)
    assert rc == 0 and len(op) > 100 and margin.min() >= 0 and miss > 0
    k = np.pi / 180
    ts_from = ts.max() - 7 * 86400000
    want = []
    for i in np.nonzero(ts >= ts_from)[0]:
        same = np.nonzero(preg == reg[i])[0]
        s1 = np.sin((plat[same] * k - lat[i] * k) / 2)
        s2 = np.sin((plon[same] * k - lon[i] * k) / 2)
        d = 6371000.0 * 2 * np.arcsin(np.sqrt(s1 * s1 + np.cos(lat[i] * k) * np.cos(plat[same] * k) * (s2 * s2)))
        for p in same[d <= 100.0]:
            want.append((int(person[i]), int(ts[i]), int(pid[p]), int(reg[i]), int(pcat[p])))
    got = list(zip(op.tolist(), ot.tolist(), opl.tolist(), org.tolist(), oc.tolist()))
    assert sorted(got) == sorted(want)
    assert oracle.lib().vro_distance_meters  # exported
