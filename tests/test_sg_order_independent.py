"""Independent check of the SG oracle's summation order (parity item (v)).

The oracle and the CUDA engine share a *canonical* order for the in-edge sum of a vertex (32 lane-strided partial
sums + xor-butterfly, 1024-term segments, source blocks of 3 * 2^21 vertices -- oracle/vrec_oracle.c `warp_sum`,
`sg_sigma`), so "engine == oracle bit for bit" says nothing about how far that order is from the reference's.  The
reference sums in whatever order Spark's partial aggregation produces (stochastic/StochasticRecommender.scala:113-114);
north_star's tolerance against it is 1e-6 relative.  Here the same iteration is evaluated by code that shares nothing
with the oracle:

  * plain left-to-right fp64 over ascending sources (scipy's CSR mat-vec: `sum += Ax[jj] * Xx[Aj[jj]]`),
  * the same in 80-bit extended precision (numpy longdouble products, `np.add.reduceat`),

on a graph of BASELINE config 2's size (the default per-region graph of the sample generator, 1 M persons per
region) and on a graph with more than 3 * 2^21 vertices (two source blocks, hub rows that are long in both).
Asserted: every entry within 1e-6 relative of both, the same number of iterations, the same convergence flag and
the same top-10 ids in the same order.  CPU only."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))

ALPHA = 0.15                      # stochastic/StochasticRecommender.scala:38


def _independent_run(src, tgt, w, n, uidx, eps, max_it, extended):
    """step() (:92-106) with calcNextX (:108-128) as a plain CSR mat-vec, isConverged (:130-141) left to right."""
    if extended:
        order = np.lexsort((src, tgt))                   # rows ascending, sources ascending inside a row
        s_sorted, t_sorted = src[order], tgt[order]
        w_sorted = w[order].astype(np.longdouble)
        starts = np.flatnonzero(np.r_[True, t_sorted[1:] != t_sorted[:-1]])
        rows = t_sorted[starts]
    else:
        a = sp.csr_matrix((w, (tgt, src)), shape=(n, n))
        a.sort_indices()
    x = np.full(n, 1.0 / n)                              # :53-54
    u = np.zeros(n)
    u[uidx] = 1.0
    if max_it <= 0:
        return x, 0, 0
    for it in range(max_it):
        if extended:
            prod = x[s_sorted].astype(np.longdouble) * w_sorted
            sigma = np.zeros(n, dtype=np.longdouble)
            sigma[rows] = np.add.reduceat(prod, starts)
            nx = (u * ALPHA + (sigma * np.longdouble(1 - ALPHA))).astype(np.float64)
        else:
            nx = u * ALPHA + (a @ x) * (1 - ALPHA)
        d = nx - x
        res = float(np.sum((d * d).astype(np.longdouble))) if extended else float(np.cumsum(d * d)[-1])
        if res <= eps * eps:
            return nx, it, 1                             # "Converged in $iteration iterations" returns nextX
        x = nx
    return x, max_it, 0


def _top10(ids, x, vertex, flt):
    keep = (ids != vertex) & (x > 0)
    if flt is not None:
        keep &= np.isin(ids, flt)
    i, v = ids[keep], x[keep]
    order = np.lexsort((i, -v))[:10]
    return i[order].tolist()


def _compare(oracle, s, t, w, vertex, eps, max_it, flt):
    og = oracle.SgGraph(s, t, w)
    n = og.N
    rc, ox, oit, oconv, _ = og.run(vertex, eps, max_it)
    assert rc == 0
    si = np.searchsorted(og.ids, s)
    ti = np.searchsorted(og.ids, t)
    uidx = int(np.searchsorted(og.ids, vertex))
    want_top = _top10(og.ids, ox, vertex, flt)
    worst = 0.0
    for extended in (False, True):
        x, it, conv = _independent_run(si, ti, w, n, uidx, eps, max_it, extended)
        assert (it, conv) == (oit, oconv), (extended, it, conv, oit, oconv)
        big = np.abs(x) > 1e-300
        rel = np.abs(ox[big] - x[big]) / np.abs(x[big])
        worst = max(worst, float(rel.max()))
        assert rel.max() <= 1e-6, (extended, float(rel.max()))            # north_star's tolerance
        assert np.array_equal(ox == 0, x == 0)
        assert _top10(og.ids, x, vertex, flt) == want_top, extended
    return worst


@pytest.mark.timeout(600)
def test_canonical_order_vs_left_to_right_config2_size(oracle):
    from vrec import synth
    pl = synth.sample_places(30000, seed=0)
    v = synth.sample_place_visits(pl, 0, persons_per_region=1_000_000, person_count_total=3_000_000, seed=0)
    s, t, w = synth.build_stochastic_graph(v)
    assert len(s) > 2_000_000                                             # config 2: ~2.4 M edges
    flt = pl.of_region(0)
    persons = np.unique(v.person_id)
    worst = 0.0
    for person in (int(persons[0]), int(persons[len(persons) // 2])):
        # launcher values (bin/stochastic_recommender.sh:32-35), then a tight epsilon that runs every iteration
        for eps, max_it in ((0.01, 20), (1e-9, 20)):
            worst = max(worst, _compare(oracle, s, t, w, person, eps, max_it, flt))
    assert worst < 1e-9           # measured 3.7e-12: rows with 10^5 terms, left-to-right vs the canonical tree


@pytest.mark.timeout(600)
def test_canonical_order_vs_left_to_right_two_source_blocks(oracle):
    from vrec import synth
    s, t, w = synth.random_stochastic_graph(6_400_000, 2, seed=8, hub_fraction=0.3)
    ids = np.unique(np.concatenate([s, t]))
    assert len(ids) > 3 * 2 ** 21                                         # more than one canonical source block
    for vertex, eps, max_it in ((int(ids[12345]), 1e-4, 8), (int(ids[-7]), 0.0, 4)):
        _compare(oracle, s, t, w, vertex, eps, max_it, None)
