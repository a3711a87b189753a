"""Pins the CPU oracle against the reference's own known-answer tests."""
import numpy as np

from tests.golden import reference_kats as K


def test_vector_length_kats(oracle):
    for (idx, vals), want in K.VECTOR_LENGTH:
        assert oracle.vector_length(vals) == want


def test_cosine_kats(oracle):
    for ((i1, v1), (i2, v2)), want in K.COSINE:
        assert oracle.cosine(i1, v1, i2, v2) == want


def _graph(oracle):
    s, t, w = zip(*K.SG_EDGES)
    return oracle.SgGraph(s, t, w)


def test_sg_kats_exact(oracle):
    g = _graph(oracle)
    assert list(g.ids) == [1, 2, 3, 4, 5]
    for (vertex, eps, max_it), want in K.SG_CASES:
        rc, x, it, conv, res = g.run(vertex, eps, max_it)
        assert rc == 0
        got = [(int(i), float(p)) for i, p in zip(g.ids, x) if i != vertex and p > 0]
        got.sort(key=lambda r: -r[1])
        assert got == want          # exact double equality, as in the reference
        rc, ids, pr, it2, conv2 = g.query(vertex, eps, max_it, None, 10)
        assert rc == 0 and list(zip(ids.tolist(), pr.tolist())) == want


def test_sg_converge_message_values(oracle):
    # the "converge" case stops with iteration = 3, residual 0.00234025
    g = _graph(oracle)
    rc, x, it, conv, res = g.run(1, 0.05, 1000)
    assert (rc, it, conv) == (0, 3, 1)
    assert abs(res - 0.00234025) < 1e-8
    rc, x, it, conv, res = g.run(1, 0.01, 1)
    assert (rc, it, conv) == (0, 1, 0)


def test_sg_missing_vertex(oracle):
    g = _graph(oracle)
    rc, *_ = g.run(K.SG_MISSING_VERTEX, 0.05, 1000)
    assert rc == oracle.ENOENT


def test_sg_corner_cases(oracle):
    g = _graph(oracle)
    rc, x, it, conv, _ = g.run(1, 0.05, 0)          # maxIterations = 0 returns x0
    assert rc == 0 and conv == 0 and it == 0 and np.all(x == 1.0 / 5)
    rc, *_ = g.run(1, -1.0, 5)
    assert rc == oracle.EINVAL
    rc, ids, pr, it, conv = g.query(1, 0.01, 20, [2, 4, 99], 10)   # filter keeps only listed ids
    assert rc == 0 and sorted(ids.tolist()) == [2, 4]
    rc, ids, pr, it, conv = g.query(1, 0.01, 20, None, 0)
    assert rc == 0 and len(ids) == 0


def test_location_distance_kats(oracle):
    """LocationTest.scala:8-29 against the oracle's haversine (Location.scala:30-43)."""
    import ctypes as C
    f = oracle.lib().vro_distance_meters
    f.restype = C.c_double
    f.argtypes = [C.c_double] * 4
    assert f(*K.LOCATION_SAME, *K.LOCATION_SAME) == 0
    (a, b) = K.LOCATION_PAIR
    d = f(*a, *b)
    assert abs(d - K.LOCATION_PAIR_DISTANCE) <= K.LOCATION_PAIR_TOLERANCE
    assert d == f(*b, *a)


def test_balanced_weights_kat(oracle):
    """StochasticGraphBuilderTest.scala:11-66 through vro_build_edge_family: every family's weights are
    reproduced from visit counts (0.4 = 2/5, 0.6 = 3/5, 1.0 = n/n), multiplied by the family's beta, and the
    balanced weights of every source sum to exactly 1.0."""
    sums = {}
    for beta, edges in K.SGB_FAMILIES:
        src = [e[0] for e in edges]
        dst = [e[1] for e in edges]
        cnt = [int(round(e[2] * 5)) for e in edges]                  # 1.0 -> 5, 0.4 -> 2, 0.6 -> 3
        rc, s, t, w = oracle.build_edge_family(src, dst, cnt, 100, beta)
        assert rc == 0
        want = {(e[0], e[1]): e[2] * beta for e in edges}
        assert {(a, b): x for a, b, x in zip(s.tolist(), t.tolist(), w.tolist())} == want     # exact doubles
        for a, x in zip(s.tolist(), w.tolist()):
            sums[a] = sums.get(a, 0.0) + x
    assert sums and all(v == 1.0 for v in sums.values()), sums
