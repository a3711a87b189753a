"""GPU parity: the CUDA engine, called through the C ABI, against the CPU oracle on the same
seeded inputs.  fp64 values are compared bit-for-bit (the engine shares the oracle's canonical
summation orders); north_star's tolerance is 1e-6 relative, so exact equality is the stricter
statement.  Integer / index outputs (ids, counts, order, status) are always exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))

from tests.golden import reference_kats as K  # noqa: E402
from tests.helpers import oracle_knn_data  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vrec():
    import vrec as v
    return v


@pytest.fixture(scope="module")
def ctx(vrec):
    c = vrec.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def synth():
    from vrec import synth as s
    return s


def test_tcgen05_selftest(vrec, ctx):
    import ctypes as C
    err = C.c_double(-1.0)
    assert ctx.lib.vrec_debug_tc_selftest(ctx._h, C.byref(err)) == 0
    # fp16 inputs are exact, products exact in fp32, 128-term fp32 accumulation: error << 1e-3
    assert 0.0 <= err.value < 1e-3, err.value


# ------------------------------------------------------------------ SG
def _kat_graph(vrec, ctx):
    s, t, w = zip(*K.SG_EDGES)
    return vrec.StochasticGraph(s, t, w, ctx=ctx)


def test_sg_reference_kats_exact(vrec, ctx):
    g = _kat_graph(vrec, ctx)
    assert g.vertex_ids().tolist() == [1, 2, 3, 4, 5]
    for (vertex, eps, max_it), want in K.SG_CASES:
        rec = vrec.StochasticRecommender(g, eps, max_it)
        ids, pr = rec.makeRecommendations(vertex)
        got = sorted(zip(ids.tolist(), pr.tolist()), key=lambda r: -r[1])
        assert got == want                      # exact doubles, as StochasticRecommenderTest does
        oi, op, cnt, its, conv, st = rec.recommend([vertex], None, 10)
        assert list(zip(oi[0, :cnt[0]].tolist(), op[0, :cnt[0]].tolist())) == want
    rec = vrec.StochasticRecommender(g, 0.05, 1000)
    rec.stationary(1)
    assert (rec.last_iterations, rec.last_converged) == (3, 1)
    rec = vrec.StochasticRecommender(g, 0.01, 1)
    rec.stationary(1)
    assert (rec.last_iterations, rec.last_converged) == (1, 0)


def test_sg_errors_and_corner_cases(vrec, ctx):
    g = _kat_graph(vrec, ctx)
    with pytest.raises(vrec.NoSuchElement, match="No such vertex in the graph: 100"):
        vrec.StochasticRecommender(g, 0.05, 1000).makeRecommendations(K.SG_MISSING_VERTEX)
    with pytest.raises(ValueError):
        vrec.StochasticRecommender(g, -0.1, 10)
    with pytest.raises(ValueError):
        vrec.StochasticRecommender(g, 0.1, -1)
    x = vrec.StochasticRecommender(g, 0.05, 0).stationary(1)        # maxIterations = 0 -> x0
    assert np.all(x == 0.2)
    rec = vrec.StochasticRecommender(g, 0.0, 7)                      # epsilon = 0 runs to the limit
    rec.stationary(1)
    assert (rec.last_iterations, rec.last_converged) == (7, 0)
    oi, op, cnt, its, conv, st = vrec.StochasticRecommender(g, 0.01, 20).recommend([1, 100, 2], [2, 4, 99, 4], 10)
    assert st.tolist() == [0, -2, 0]
    assert sorted(oi[0, :cnt[0]].tolist()) == [2, 4] and cnt[1] == 0
    oi, op, cnt, *_ = vrec.StochasticRecommender(g, 0.01, 20).recommend([1], None, 0)
    assert cnt[0] == 0


def test_sg_degenerate_graphs(vrec, ctx, oracle):
    # no edges at all, one self-loop, one edge, duplicate edges and a dangling vertex
    g = vrec.StochasticGraph([], [], [], ctx=ctx)
    assert g.N == 0
    oi, op, cnt, its, conv, st = vrec.StochasticRecommender(g, 0.01, 20).recommend([1, 2], None, 5)
    assert st.tolist() == [-2, -2] and cnt.tolist() == [0, 0]
    for s, t, w in (([7], [7], [1.0]), ([1], [2], [1.0]), ([1, 1, 1, 2], [2, 2, 3, 3], [0.25, 0.25, 0.5, 1.0])):
        g = vrec.StochasticGraph(s, t, w, ctx=ctx)
        og = oracle.SgGraph(s, t, w)
        for v in sorted(set(s) | set(t)):
            for eps, mi in ((0.0, 6), (0.01, 20), (0.5, 0)):
                rec = vrec.StochasticRecommender(g, eps, mi)
                x = rec.stationary(v)
                rc, ox, oit, oconv, _ = og.run(v, eps, mi)
                assert rc == 0 and np.array_equal(x, ox) and (rec.last_iterations, rec.last_converged) == (oit, oconv)
                _check_sg_batch(rec, og, [v], None, 3, eps, mi)


@pytest.mark.parametrize("n,deg,hub,seed", [(50, 3, 0.0, 1), (2000, 6, 0.0, 2), (5000, 5, 0.5, 3),
                                            (30000, 40, 0.3, 4)])
def test_sg_random_graphs_bit_exact(vrec, ctx, synth, oracle, n, deg, hub, seed):
    s, t, w = synth.random_stochastic_graph(n, deg, seed=seed, hub_fraction=hub)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx)
    og = oracle.SgGraph(s, t, w)
    assert np.array_equal(g.vertex_ids(), og.ids)
    flt = og.ids[::3]
    for vertex, eps, max_it in [(int(og.ids[0]), 0.0, 5), (int(og.ids[n // 2]), 1e-3, 20),
                                (int(og.ids[-1]), 1e-6, 12)]:
        rec = vrec.StochasticRecommender(g, eps, max_it)
        x = rec.stationary(vertex)
        rc, ox, oit, oconv, ores = og.run(vertex, eps, max_it)
        assert rc == 0
        assert (rec.last_iterations, rec.last_converged) == (oit, oconv)
        assert np.array_equal(x, ox)                                   # bit-exact
        assert abs(rec.last_residual - ores) <= 1e-12 * max(ores, 1e-300)
        oi, op, cnt, its, conv, st = rec.recommend([vertex], flt, 10)
        rc, wi, wp, _, _ = og.query(vertex, eps, max_it, flt, 10)
        assert oi[0, :cnt[0]].tolist() == wi.tolist() and op[0, :cnt[0]].tolist() == wp.tolist()


def test_sg_default_sample_graph(vrec, ctx, synth, oracle):
    # config 2 shape at reduced size: per-region graph of sample-generator data (giant place rows)
    pl = synth.sample_places(30000, seed=0)
    v = synth.sample_place_visits(pl, 0, persons_per_region=30000, person_count_total=3_000_000, seed=0)
    s, t, w = synth.build_stochastic_graph(v)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx)
    og = oracle.SgGraph(s, t, w)
    person = int(v.person_id[0])
    rec = vrec.StochasticRecommender(g, 0.01, 20)                     # bin/stochastic_recommender.sh:32-35
    x = rec.stationary(person)
    rc, ox, oit, oconv, _ = og.run(person, 0.01, 20)
    assert (rec.last_iterations, rec.last_converged) == (oit, oconv)
    assert np.array_equal(x, ox)
    flt = pl.of_region(0)
    oi, op, cnt, *_ = rec.recommend([person], flt, 10)
    rc, wi, wp, _, _ = og.query(person, 0.01, 20, flt, 10)
    assert oi[0, :cnt[0]].tolist() == wi.tolist() and op[0, :cnt[0]].tolist() == wp.tolist()


def test_sg_generated_graph_matches_oracle(vrec, ctx, oracle):
    g = vrec.StochasticGraph.generate(20000, 16, seed=5, ctx=ctx)
    rowptr, src, w = g.export_csr()
    assert rowptr[-1] == g.nnz == 20000 * 16
    tgt = np.repeat(np.arange(g.N), np.diff(rowptr))
    for r in (0, 1, 777):
        assert np.all(np.diff(src[rowptr[r]:rowptr[r + 1]]) >= 0)
    og = oracle.SgGraph(src.astype(np.int64), tgt.astype(np.int64), w)
    if og.N == g.N:       # every vertex appears (overwhelmingly likely at this density)
        x = vrec.StochasticRecommender(g, 0.0, 6).stationary(0)
        rc, ox, *_ = og.run(0, 0.0, 6)
        assert np.array_equal(x, ox)


def test_sg_source_blocked_order(vrec, ctx, synth, oracle):
    # more than 3 * 2^21 vertices: the engine sweeps the graph once per block of 3 * 2^21 sources and the
    # canonical order adds the block sums left to right (oracle: CANON_SRC_BLOCK)
    g = vrec.StochasticGraph.generate(7_000_000, 6, seed=7, ctx=ctx)
    rowptr, src, w = g.export_csr()
    og = oracle.SgGraph.from_csr(rowptr.astype(np.int64), src, w)
    rec = vrec.StochasticRecommender(g, 0.0, 3)
    x = rec.stationary(0)
    rc, ox, oit, oconv, ores = og.run(0, 0.0, 3)
    assert rc == 0 and np.array_equal(x, ox)
    assert abs(rec.last_residual - ores) <= 1e-9 * ores     # two different fixed summation orders
    g.close()
    # host-loaded graph whose hub rows are long in BOTH source blocks (per-block segment tables)
    s, t, w = synth.random_stochastic_graph(6_400_000, 2, seed=8, hub_fraction=0.3)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx)
    og = oracle.SgGraph(s, t, w)
    v = int(og.ids[12345])
    rec = vrec.StochasticRecommender(g, 1e-4, 6)
    x = rec.stationary(v)
    rc, ox, oit, oconv, _ = og.run(v, 1e-4, 6)
    assert (rec.last_iterations, rec.last_converged) == (oit, oconv)
    assert np.array_equal(x, ox)
    g.close()
    # the same graph row-partitioned over 3 parts: two source blocks per part, peer stores, residual slots
    grp = vrec.StochasticGraphGroup(s, t, w, 3, ctx=ctx)
    xs, its, conv, _ = grp.stationary(v, 1e-4, 6)
    for r in range(3):
        assert (int(its[r]), int(conv[r])) == (oit, oconv), r
        assert np.array_equal(xs[r], ox), r
    grp.close()


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_sg_row_partitioned_group_bit_exact(vrec, ctx, synth, oracle, world):
    """SURVEY 8(e) row 3 on one GPU: the parts of a row-partitioned graph (rows balanced by in-edges), every
    part's sweep storing its rows of x' into all parts' buffers, the residual partials added in rank order --
    every part's copy of the result equals the oracle bit for bit, with the same iteration count."""
    from vrec.engine import host_sg_csr, host_sg_partition
    for (n, deg, hub, seed) in [(5000, 5, 0.5, 3), (30001, 20, 0.2, 4), (64, 2, 0.0, 9)]:
        s, t, w = synth.random_stochastic_graph(n, deg, seed=seed, hub_fraction=hub)
        og = oracle.SgGraph(s, t, w)
        grp = vrec.StochasticGraphGroup(s, t, w, world, ctx=ctx)
        ids, rowptr, _, _ = host_sg_csr(s, t, w)
        bounds = host_sg_partition(rowptr, world)
        assert [p.row_range() for p in grp.parts] == [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]
        assert sum(p.nnz for p in grp.parts) == len(s)
        for vertex, eps, max_it in [(int(og.ids[0]), 0.0, 6), (int(og.ids[n // 2]), 1e-3, 20),
                                    (int(og.ids[-1]), 0.05, 0)]:
            xs, its, conv, res = grp.stationary(vertex, eps, max_it)
            rc, ox, oit, oconv, ores = og.run(vertex, eps, max_it)
            assert rc == 0
            for r in range(world):
                assert (int(its[r]), int(conv[r])) == (oit, oconv), (world, r, vertex)
                assert np.array_equal(xs[r], ox), (world, r, vertex)
            if max_it > 0:
                assert np.all(res == res[0])                                  # the same sum on every rank
                assert abs(res[0] - ores) <= 1e-12 * max(ores, 1e-300)
        with pytest.raises(vrec.NoSuchElement):
            grp.stationary(-12345, 0.01, 5)
        grp.close()


def _check_sg_batch(rec, og, vertices, flt, max_recs, eps, max_it):
    oi, op, cnt, its, conv, st = rec.recommend(vertices, flt, max_recs)
    for q, v in enumerate(vertices):
        rc, wi, wp, oit, oconv = og.query(int(v), eps, max_it, flt, max_recs)
        assert st[q] == rc
        if rc:
            assert cnt[q] == 0
            continue
        assert (int(its[q]), int(conv[q])) == (oit, oconv), (q, v)
        assert oi[q, :cnt[q]].tolist() == wi.tolist(), (q, v)
        assert op[q, :cnt[q]].tolist() == wp.tolist(), (q, v)             # bit-exact


@pytest.mark.parametrize("tpc", [0, 1, 2])
def test_sg_batch_kernel_bit_exact(vrec, ctx, synth, oracle, tpc):
    # config 4 shape: many person start vertices on one graph; hub rows longer than the canonical
    # segment in both the place->place prefix and the person part, duplicate edges
    s, t, w, persons, places, cats = synth.random_layered_graph(5, 1500, 4000, seed=11)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx)
    og = oracle.SgGraph(s, t, w)
    assert g.batch_info(1) == 1 and g.batch_info(2) == 1505
    g.set_option("batch", 2)
    g.set_option("batch_targets_per_cta", tpc)
    rng = np.random.default_rng(tpc)
    some = rng.choice(persons, 37, replace=False)
    flt = np.concatenate([places[::2], [places[3], 10 ** 9]])          # a duplicate and an unknown id
    for eps, max_it, f, m in [(0.01, 20, flt, 10), (0.0, 4, None, 7), (1e-6, 20, flt, 3), (0.01, 1, flt, 10),
                              (0.01, 2, places, 2000), (1e-3, 20, flt, 0)]:
        rec = vrec.StochasticRecommender(g, eps, max_it)
        _check_sg_batch(rec, og, some, f, m, eps, max_it)
        assert g.batch_info(0) == len(some)
    # a call that mixes persons, a place, a category and an unknown id
    mixed = np.array([persons[0], places[0], 10 ** 9 + 1, persons[5], cats[1], persons[0]])
    rec = vrec.StochasticRecommender(g, 0.01, 20)
    _check_sg_batch(rec, og, mixed, flt, 10, 0.01, 20)
    assert g.batch_info(0) == 3
    # parameters for which the shared first iteration is already the answer use the per-query path
    rec = vrec.StochasticRecommender(g, 10.0, 20)
    _check_sg_batch(rec, og, some[:5], flt, 10, 10.0, 20)
    assert g.batch_info(0) == 0
    rec = vrec.StochasticRecommender(g, 0.01, 0)
    _check_sg_batch(rec, og, some[:5], flt, 10, 0.01, 0)
    assert g.batch_info(0) == 0
    g.set_option("batch", 0)
    rec = vrec.StochasticRecommender(g, 0.01, 20)
    _check_sg_batch(rec, og, some[:5], flt, 10, 0.01, 20)
    assert g.batch_info(0) == 0
    with pytest.raises(ValueError):
        g.set_option("batch", 7)


def test_sg_batch_default_sample_graph(vrec, ctx, synth, oracle):
    # per-region graph of sample-generator data through the four edge calculators
    pl = synth.sample_places(30000, seed=0)
    v = synth.sample_place_visits(pl, 0, persons_per_region=30000, person_count_total=3_000_000, seed=0)
    s, t, w = synth.build_stochastic_graph(v)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx)
    og = oracle.SgGraph(s, t, w)
    assert g.batch_info(1) == 1
    persons = np.unique(v.person_id)[:300]
    rec = vrec.StochasticRecommender(g, 0.01, 20)                     # bin/stochastic_recommender.sh:32-35
    _check_sg_batch(rec, og, persons, pl.of_region(0), 10, 0.01, 20)
    assert g.batch_info(0) == len(persons)


def test_sg_batch_not_applicable(vrec, ctx, synth, oracle):
    # every vertex has in-edges: no batch path, same results through the per-query kernels
    s, t, w = synth.random_stochastic_graph(3000, 6, seed=2)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx)
    og = oracle.SgGraph(s, t, w)
    g.set_option("batch", 2)
    rec = vrec.StochasticRecommender(g, 1e-3, 20)
    _check_sg_batch(rec, og, og.ids[:6], og.ids[::3], 10, 1e-3, 20)
    assert g.batch_info(0) == 0
    # active sources that follow in-degree-0 sources in a row (person ids below place ids)
    s, t, w, persons, places, cats = synth.random_layered_graph(4, 300, 500, seed=3)
    remap = {int(p): -int(p) for p in persons}
    s2 = np.array([remap.get(int(x), int(x)) for x in s], dtype=np.int64)
    g = vrec.StochasticGraph(s2, t, w, ctx=ctx)
    og = oracle.SgGraph(s2, t, w)
    assert g.batch_info(1) == 0
    rec = vrec.StochasticRecommender(g, 0.01, 20)
    _check_sg_batch(rec, og, [-int(persons[0]), -int(persons[7]), -int(persons[9]), -int(persons[11])], places, 10,
                    0.01, 20)


# ------------------------------------------------------------------ KNN
def _check_knn(vrec, oracle, rs, inp, pw, cw, k, targets, flt, max_recs):
    rec = vrec.KnnRecommender(rs, pw, cw, k)
    d = oracle_knn_data(oracle, inp)
    pl, rt, cnt, st = rec.recommend(targets, flt, max_recs)
    rc, opl, ort, ocnt, ost = oracle.knn_query_batch(d, targets, pw, cw, k, flt, max_recs)
    assert rc == 0
    assert st.tolist() == ost.tolist()
    assert cnt.tolist() == ocnt.tolist()
    for q in range(len(targets)):
        assert pl[q, :cnt[q]].tolist() == opl[q, :cnt[q]].tolist(), (q, targets[q])
        assert rt[q, :cnt[q]].tolist() == ort[q, :cnt[q]].tolist(), (q, targets[q])
    return rec, d


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("k", [1, 7, 50, 5000])
@pytest.mark.parametrize("path,kernel", [(0, 0), (1, 1), (1, 2), (1, 3), (1, 4), (2, 0)])
def test_knn_random_bit_exact(vrec, ctx, synth, oracle, seed, k, path, kernel):
    # path: rating reduction (0 auto, 1 neighbour-row gather, 2 column scan)
    # kernel: similarity + top-K (0 auto, 1 per-target exact scan, 2 tiled fp32 filter + exact survivors,
    #         3 the first tcgen05 filter kernel when the library was built with -DVREC_WITH_TC_BASELINE=1, else
    #         the same as 4: tcgen05 fp16 filter, warp-specialised with TMA bulk copies, + exact survivors)
    if path == 1 and k > 1024:
        pytest.skip("gather path needs k <= 1024")
    inp = synth.random_knn_inputs(700, 60, 9, seed=seed, separate_ratings=(seed == 2))
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    rs.set_option("rating_path", path)
    rs.set_option("knn_kernel", kernel)
    targets = np.concatenate([inp.person_id[:40], [1, 999999]])       # two unknown persons
    flt = np.arange(0, 60, 2)
    rec, d = _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, k, targets, flt, 10)
    _check_knn(vrec, oracle, rs, inp, 0.25, 0.75, k, targets[:8], None, 60)
    # neighbours and raw estimates of one target
    t = int(inp.person_id[5])
    rc, oids, osims = oracle.knn_neighbours(d, t, 0.5, 0.5, k)
    if rc == 0:
        sims = rec.similarities(t)
        want = np.zeros(len(inp.person_id))
        want[np.searchsorted(inp.person_id, oids)] = osims
        assert np.array_equal(sims, want)
        if k <= 1024:
            ids, s2 = rec.findSimilarPersons(t)
            assert ids.tolist() == oids.tolist() and s2.tolist() == osims.tolist()
        epl, ert = rec.makeRecommendations(t)
        rc, wpl, wrt = oracle.knn_estimates(d, t, 0.5, 0.5, k)
        assert epl.tolist() == wpl.tolist() and ert.tolist() == wrt.tolist()
    rs.close()


@pytest.mark.parametrize("k", [5, 50])
def test_knn_tail_places_postings_pass(vrec, ctx, synth, oracle, k):
    # far more places than filter dimensions: most pairs that matter share a TAIL place and come from the
    # postings pass; a target listed twice and an unknown person in the batch
    inp = synth.random_knn_inputs(4000, 900, 9, seed=21)
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    targets = np.concatenate([inp.person_id[:300], inp.person_id[:3], [999999]])
    _check_knn(vrec, oracle, rs, inp, 0.7, 0.3, k, targets, np.arange(0, 900, 3), 10)
    rs.close()


def test_knn_fallback_kernels_for_unusual_data(vrec, ctx, synth, oracle):
    # data outside the tensor-core path's preconditions must silently take the exact kernels:
    # (a) negative vector values (the dense-row / signature shortcuts assume non-negative data),
    # (b) more categories than filter dimensions, (c) long rows (hundreds of places per person), (d) K > 56
    import dataclasses
    inp = synth.random_knn_inputs(1500, 400, 9, seed=31)
    neg = dataclasses.replace(inp, place_val=np.where(np.arange(len(inp.place_val)) % 7 == 0, -inp.place_val,
                                                      inp.place_val))
    rs = vrec.KnnRegionSet(*neg.load_args(), ctx=ctx)
    _check_knn(vrec, oracle, rs, neg, 0.5, 0.5, 20, neg.person_id[:30], np.arange(0, 400, 2), 10)
    rs.close()
    # values that are not exact in fp32: the packed records must then keep doubles (visit counts pack as floats)
    frac = dataclasses.replace(inp, place_val=inp.place_val * 0.1 + 1e-9, cat_val=inp.cat_val / 3.0)
    rs = vrec.KnnRegionSet(*frac.load_args(), ctx=ctx)
    _check_knn(vrec, oracle, rs, frac, 0.5, 0.5, 20, frac.person_id[:60], np.arange(0, 400, 2), 10)
    rs.close()
    wide = synth.random_knn_inputs(1200, 300, 90, seed=32)
    rs = vrec.KnnRegionSet(*wide.load_args(), ctx=ctx)
    _check_knn(vrec, oracle, rs, wide, 0.4, 0.6, 20, wide.person_id[:30], None, 10)
    rs.close()
    long_rows = synth.random_knn_inputs(600, 2500, 9, seed=33, max_places=300)
    rs = vrec.KnnRegionSet(*long_rows.load_args(), ctx=ctx)
    _check_knn(vrec, oracle, rs, long_rows, 0.5, 0.5, 20, long_rows.person_id[:20], np.arange(0, 2500, 5), 10)
    _check_knn(vrec, oracle, rs, long_rows, 0.5, 0.5, 200, long_rows.person_id[:20], np.arange(0, 2500, 5), 10)
    rs.close()


def test_knn_errors(vrec, ctx, synth):
    inp = synth.random_knn_inputs(50, 10, 4, seed=3)
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    for pw, cw, k in [(0.0, 1.0, 5), (1.0, 0.0, 5), (0.6, 0.5, 5), (0.5, 0.5, 0)]:
        with pytest.raises(ValueError, match="requirement failed"):
            vrec.KnnRecommender(rs, pw, cw, k)
    rec = vrec.KnnRecommender(rs, 0.5, 0.5, 5)
    with pytest.raises(vrec.NoSuchElement, match="No such person: 12345678"):
        rec.makeRecommendations(12345678)
    pl, rt, cnt, st = rec.recommend([int(inp.person_id[0])], None, 0)
    assert cnt[0] == 0
    pl, rt, cnt, st = rec.recommend([], None, 5)
    assert len(cnt) == 0
    # unsorted persons are accepted (sorted internally)
    perm = np.random.default_rng(0).permutation(len(inp.person_id))
    rows = [(inp.place_col[inp.place_rowptr[i]:inp.place_rowptr[i + 1]],
             inp.place_val[inp.place_rowptr[i]:inp.place_rowptr[i + 1]],
             inp.cat_col[inp.cat_rowptr[i]:inp.cat_rowptr[i + 1]],
             inp.cat_val[inp.cat_rowptr[i]:inp.cat_rowptr[i + 1]]) for i in perm]
    prp = np.concatenate([[0], np.cumsum([len(r[0]) for r in rows])])
    crp = np.concatenate([[0], np.cumsum([len(r[2]) for r in rows])])
    rs2 = vrec.KnnRegionSet(inp.person_id[perm], prp, np.concatenate([r[0] for r in rows]),
                            np.concatenate([r[1] for r in rows]), inp.place_dim, crp,
                            np.concatenate([r[2] for r in rows]), np.concatenate([r[3] for r in rows]),
                            inp.cat_dim, inp.rating_person, inp.rating_place, inp.rating_value, ctx=ctx)
    t = inp.person_id[:10]
    a = vrec.KnnRecommender(rs, 0.5, 0.5, 5).recommend(t, None, 5)
    b = vrec.KnnRecommender(rs2, 0.5, 0.5, 5).recommend(t, None, 5)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    bad_col = inp.place_col.copy()
    if len(bad_col) > 1:
        bad_col[:] = 0
        with pytest.raises(ValueError):
            vrec.KnnRegionSet(inp.person_id, inp.place_rowptr, bad_col, inp.place_val, inp.place_dim,
                              inp.cat_rowptr, inp.cat_col, inp.cat_val, inp.cat_dim, ctx=ctx)


def test_knn_distance_kats(vrec, ctx):
    # DistanceTest.scala:34-60 through the engine: target = person 1, candidate = person 2
    for ((i1, v1), (i2, v2)), want in K.COSINE:
        pid = [1, 2]
        prp = [0, len(i1), len(i1) + len(i2)]
        rs = vrec.KnnRegionSet(pid, prp, list(i1) + list(i2), list(v1) + list(v2), 2,
                               [0, 1, 2], [0, 0], [1.0, 1.0], 1, ctx=ctx)
        sims = vrec.KnnRecommender(rs, 0.5, 0.5, 10).similarities(1)
        place_sim = want if want > 0 else 0.0
        assert sims[1] == place_sim * 0.5 + 1.0 * 0.5
        rs.close()


def test_knn_g2_shape(vrec, ctx, synth, oracle):
    # config 3 shape at reduced size (20k persons / 2k places), K = 50, both weights 0.5
    v, places = synth.g2_place_visits(20000, 2000, seed=20181231)
    inp = synth.build_rating_vectors(v)
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    targets = inp.person_id[::400]
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 50, targets, places.id, 10)
    rs.set_option("splits", 1)
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 50, targets[:10], places.id, 10)
    # one CTA walks all 157 candidate tiles -> the dense kernel bootstraps its thresholds from 32 tiles.
    # Regression: the bootstrap scratch (64 floats per target) lived in the heap area, 8 K bytes per target,
    # and overran it for K < 32 (illegal address at K = 7 on the full-size region-set)
    for k_small in (7, 31):
        _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, k_small, targets[:10], places.id, 10)
    rs.set_option("splits", 0)
    # the largest K of the fused top-K kernels: the postings kernel needs 8 x (12 K + 1728) bytes of shared
    # memory, more than the 100 KB it used to opt in to when K > 922 (launch failed with "invalid argument")
    for k_large in (923, 1024):
        _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, k_large, targets[:10], places.id, 10)
    for kern in (1, 2, 3, 4):               # exact scan, CUDA-core tile, tensor-core tiles: same answers
        rs.set_option("knn_kernel", kern)
        _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 50, targets[:10], places.id, 10)
    rs.set_option("knn_kernel", 0)
    _check_knn(vrec, oracle, rs, inp, 0.3, 0.7, 200, targets[:20], places.id, 10)
    # launcher configuration: K = 2 000 000 (every positive candidate), default sample data shape
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 2_000_000, targets[:3], places.id, 10)
    rs.close()


def test_knn_multi_wave_grid(vrec, ctx, synth, oracle):
    # 1024 targets = 8 target tiles x 32 candidate splits = 256 CTAs of the dense kernel on 148 SMs: the CTAs of
    # a tile start at different times (second wave).  Regression: the targets' dense category rows were
    # zero-filled and re-scattered by every CTA of the tile, so a late CTA wiped rows others were reading
    # (found by the full-size parity check of bench.py / tools/knn_fullsize_parity.py: ~2 % of the targets).
    v, places = synth.g2_place_visits(200000, 20000, seed=20181231)
    inp = synth.build_rating_vectors(v)
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    rng = np.random.default_rng(7)
    targets = inp.person_id[rng.choice(len(inp.person_id), 1024, replace=False)]
    rc, opl, ort, ocnt, ost = oracle.knn_query_batch(oracle_knn_data(oracle, inp), targets, 0.5, 0.5, 50,
                                                     places.id, 10)
    assert rc == 0
    rec = vrec.KnnRecommender(rs, 0.5, 0.5, 50)
    for post_first in (0, 0, 0, 1):
        rs.set_option("post_first", post_first)
        pl, rt, cnt, st = rec.recommend(targets, places.id, 10)
        assert st.tolist() == ost.tolist() and cnt.tolist() == ocnt.tolist()
        assert np.array_equal(pl, opl)
        assert np.array_equal(rt.view(np.int64), ort.view(np.int64))            # bit-exact
    rs.set_option("post_first", 0)
    rs.close()


@pytest.mark.parametrize("k", [5, 50, 300])
def test_knn_rating_scratch_paths(vrec, ctx, synth, oracle, k):
    # more than 4096 places: the rating reduction of a target whose neighbours hold <= 512 ratings goes through the
    # shared-memory place table (K = 5, most of K = 50), the others through the dense scratch rows (K = 300, the
    # heavy tail of K = 50) -- in one launch, the same warp alternating between the two
    v, places = synth.g2_place_visits(50000, 8000, seed=11)
    inp = synth.build_rating_vectors(v)
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    rng = np.random.default_rng(k)
    targets = inp.person_id[rng.choice(len(inp.person_id), 600, replace=False)]
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, k, targets, places.id[::2], 10)
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, k, targets[:50], None, 25)
    rs.close()


def test_knn_default_sample_data(vrec, ctx, synth, oracle):
    # config 1 shape at reduced size: degenerate diagonal data, huge tie groups
    pl = synth.sample_places(30000, seed=0)
    v = synth.sample_place_visits(pl, 0, persons_per_region=20000, person_count_total=3_000_000, seed=0)
    inp = synth.build_rating_vectors(v)
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    targets = inp.person_id[:3]
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 2_000_000, targets, pl.of_region(0), 10)
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 50, targets, pl.of_region(0), 10)
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 3000, targets[:2], pl.of_region(0), 10)   # radix-select path
    rs.close()


# ------------------------------------------------------------------ rating vectors builder
def test_rating_vectors_builder_bit_exact(vrec, ctx, synth, oracle):
    from vrec import builders
    v, places = synth.g2_place_visits(20000, 3000, seed=11, mean_places=9.0)
    rng = np.random.default_rng(3)
    perm = rng.permutation(int(v.count.sum()))
    pe = np.repeat(v.person_id, v.count)[perm]          # one row per visit, shuffled
    pl = np.repeat(v.place_id, v.count)[perm]
    ca = np.repeat(v.category_id, v.count)[perm]
    for ent, top_n in ((pl, 5), (ca, 3), (pl, 100)):
        rc, persons, rowptr, col, val, dim = oracle.build_rating_vectors(pe, ent, None, top_n)
        g = builders.build_rating_vectors(pe, ent, top_n, ctx=ctx)
        assert rc == 0 and g[4] == dim
        for a, b in zip(g[:4], (persons, rowptr, col, val)):
            assert np.array_equal(a, b)
    # pre-aggregated rows with weights give the same tables
    rc, persons, rowptr, col, val, dim = oracle.build_rating_vectors(v.person_id, v.place_id, v.count, 5)
    g = builders.build_rating_vectors(v.person_id, v.place_id, 5, weight=v.count, ctx=ctx)
    for a, b in zip(g[:4], (persons, rowptr, col, val)):
        assert np.array_equal(a, b)
    # the whole step, then the recommender on its output == the recommender on the numpy restatement's tables
    inp = builders.rating_vectors_builder(pe, pl, ca, max_rated_places=5, max_rated_categories=3, ctx=ctx)
    want = synth.build_rating_vectors(v, max_rated_places=5, max_rated_categories=3)
    for a, b in zip(inp.load_args(), want.load_args()):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    _check_knn(vrec, oracle, rs, inp, 0.5, 0.5, 7, inp.person_id[:40], places.id, 10)
    # corner cases: ties at the cut stay, out-of-range ids, empty input
    g = builders.build_rating_vectors([5, 5, 5, 5, 9], [1, 2, 3, 4, 7], 2, weight=[3, 2, 2, 1, 1], ctx=ctx)
    assert g[0].tolist() == [5, 9] and g[1].tolist() == [0, 3, 4] and g[2].tolist() == [1, 2, 3, 7] and g[4] == 8
    with pytest.raises(ValueError, match="Index out of Int range"):
        builders.build_rating_vectors([1], [2 ** 31], 10, ctx=ctx)
    g = builders.build_rating_vectors([], [], 10, ctx=ctx)
    assert len(g[0]) == 0 and g[1].tolist() == [0]


# ------------------------------------------------------------------ stochastic graph builder
def test_stochastic_graph_builder_bit_exact(vrec, ctx, synth, oracle):
    from vrec import builders
    v, places = synth.g2_place_visits(5000, 700, seed=13, mean_places=5.0)
    rng = np.random.default_rng(4)
    perm = rng.permutation(int(v.count.sum()))
    pe = np.repeat(v.person_id, v.count)[perm]
    pl = np.repeat(v.place_id, v.count)[perm]
    ca = np.repeat(v.category_id, v.count)[perm]
    ts = 1_546_300_800_000 + rng.integers(0, 20 * 24 * 3600 * 1000, len(pe))      # three weeks: the 7-day window matters
    rc, ws, wt, ww = oracle.build_stochastic_graph(pe, pl, ca, ts, 0.3, 0.7)
    assert rc == 0
    s, t, w = builders.stochastic_graph_builder(pe, pl, ca, ts, 0.3, 0.7, ctx=ctx)
    assert np.array_equal(s, ws) and np.array_equal(t, wt) and np.array_equal(w, ww)      # same order, same doubles
    # single families, with and without weights
    for src, dst, wgt, top_n, beta in ((pe, pl, None, 3, 0.5), (ca, pl, None, 100, 1.0),
                                       (v.person_id, v.category_id, v.count, 2, 0.25)):
        rc, ws, wt, ww = oracle.build_edge_family(src, dst, wgt, top_n, beta)
        s, t, w = builders.build_edge_family(src, dst, top_n, beta, weight=wgt, ctx=ctx)
        assert rc == 0 and np.array_equal(s, ws) and np.array_equal(t, wt) and np.array_equal(w, ww)
    # the built graph goes straight into the recommender
    s, t, w = builders.stochastic_graph_builder(pe, pl, ca, ts, 0.5, 0.5, ctx=ctx)
    g = vrec.StochasticGraph(s, t, w, ctx=ctx)
    og = oracle.SgGraph(s, t, w)
    rec = vrec.StochasticRecommender(g, 0.01, 20)
    _check_sg_batch(rec, og, np.unique(pe)[:12], places.id, 10, 0.01, 20)
    assert builders.stochastic_graph_builder([], [], [], [], ctx=ctx)[0].size == 0


# ------------------------------------------------------------------ place visits builder
def test_place_visits_builder_matches_oracle(vrec, ctx, oracle):
    from tests.test_oracle_selfcheck import _location_visits
    from vrec import builders
    rng = np.random.default_rng(8)
    (person, lat, lon, ts, reg), (pid, plat, plon, pcat, preg) = _location_visits(rng, n_regions=3, places_per_region=900,
                                                                                  n_visits=20000)
    reg = reg.copy()
    reg[:5] = 77                                                        # a region without places: no rows
    for days, acc in ((7, 100.0), (30, 250.0), (0, 100.0)):
        rc, want, margin, miss = oracle.build_place_visits(person, lat, lon, ts, reg, pid, plat, plon, pcat, preg, days, acc)
        got = builders.place_visits_builder(person, lat, lon, ts, reg, pid, plat, plon, pcat, preg, days, acc, ctx=ctx)
        assert rc == 0
        # floating point: a pair within 1e-6 m of the radius may legitimately fall on either side (the reference
        # itself uses FastMath, the oracle libm, the engine CUDA's libm); everything else must agree exactly,
        # in the same order (visit order, ascending place id)
        tol = 1e-6
        if margin.size and margin.min() > tol and miss > tol:
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
        else:
            gs = set(zip(*[x.tolist() for x in got]))
            ws = set(zip(*[x.tolist() for x in want]))
            safe = {r for r, m in zip(zip(*[x.tolist() for x in want]), margin.tolist()) if m > tol}
            assert safe <= gs and len(gs - ws) <= np.sum(margin <= tol) + 8
    # straight into the two builders behind it
    got = builders.place_visits_builder(person, lat, lon, ts, reg, pid, plat, plon, pcat, preg, 30, 100.0, ctx=ctx)
    inp = builders.rating_vectors_builder(got[0], got[2], got[4], ctx=ctx)
    assert len(inp.person_id) > 0 and inp.place_rowptr[-1] == len(inp.place_col)
    s_, t_, w_ = builders.stochastic_graph_builder(got[0], got[2], got[4], got[1], ctx=ctx)
    assert len(s_) > 0 and np.all(w_ > 0)
    assert builders.place_visits_builder([], [], [], [], [], pid, plat, plon, pcat, preg, ctx=ctx)[0].size == 0


def test_builder_reference_kats(vrec, ctx):
    """The reference's own builder tests through the engine: LocationTest.scala:8-29 (via the place visits
    builder's radius test) and StochasticGraphBuilderTest.scala:11-66 (balanced weights sum to exactly 1)."""
    from vrec import builders
    (a, b) = K.LOCATION_PAIR
    for radius, n_rows in ((K.LOCATION_PAIR_DISTANCE + K.LOCATION_PAIR_TOLERANCE, 1),
                           (K.LOCATION_PAIR_DISTANCE - K.LOCATION_PAIR_TOLERANCE, 0)):
        for (v, p) in ((a, b), (b, a)):                                  # commutative
            rows = builders.place_visits_builder([1], [v[0]], [v[1]], [0], [0], [40], [p[0]], [p[1]], [3], [0], 7,
                                                 radius, ctx=ctx)
            assert len(rows[0]) == n_rows
    rows = builders.place_visits_builder([1], [K.LOCATION_SAME[0]], [K.LOCATION_SAME[1]], [0], [0], [40],
                                         [K.LOCATION_SAME[0]], [K.LOCATION_SAME[1]], [3], [0], 7, 1e-9, ctx=ctx)
    assert rows[2].tolist() == [40]                                      # distance 0
    sums = {}
    for beta, edges in K.SGB_FAMILIES:
        cnt = [int(round(e[2] * 5)) for e in edges]
        s_, t_, w_ = builders.build_edge_family([e[0] for e in edges], [e[1] for e in edges], 100, beta, weight=cnt,
                                                ctx=ctx)
        assert {(x, y): z for x, y, z in zip(s_.tolist(), t_.tolist(), w_.tolist())} == {(e[0], e[1]): e[2] * beta
                                                                                        for e in edges}
        for x, z in zip(s_.tolist(), w_.tolist()):
            sums[x] = sums.get(x, 0.0) + z
    assert all(v == 1.0 for v in sums.values())
