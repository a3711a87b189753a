"""GPU parity at BASELINE config 3's full size (P = 10^6 persons, 10^5 places): the engine through the C ABI
against the oracle on a sample of targets, in the batch shapes the reduced-size tests cannot reach — more than
one wave of CTAs, CTAs that walk >= 64 candidate tiles (threshold bootstrap), the bench's own 18 944-target
batch.  Every defect of round 1 lived in one of these.  Runs last (file name) and takes
about 25 s, most of it the host-side generator and the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))

from tests.helpers import oracle_knn_data  # noqa: E402

pytestmark = pytest.mark.gpu


def _bad(got, want):
    pl, rt, cnt, st = got
    opl, ort, ocnt, ost = want
    out = []
    for q in range(len(ocnt)):
        c = int(ocnt[q])
        if not (int(cnt[q]) == c and int(st[q]) == int(ost[q]) and np.array_equal(pl[q, :c], opl[q, :c])
                and np.array_equal(np.ascontiguousarray(rt[q, :c]).view(np.int64),
                                   np.ascontiguousarray(ort[q, :c]).view(np.int64))):
            out.append(q)
    return out


def test_knn_full_size_config3(oracle):
    import vrec
    from vrec import synth
    v, places = synth.g2_place_visits(1_000_000, 100_000, seed=20181231, region=0)     # bench.py's region-set
    inp = synth.build_rating_vectors(v)
    d = oracle_knn_data(oracle, inp)
    ctx = vrec.Context(0)
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    try:
        rng = np.random.default_rng(7)
        targets = inp.person_id[rng.choice(len(inp.person_id), 256, replace=False)]
        flt = np.ascontiguousarray(places.id, dtype=np.int64)
        threads = os.cpu_count() or 1
        rc, *want = oracle.knn_query_batch(d, targets, 0.5, 0.5, 50, flt, 10, n_threads=threads)
        assert rc == 0
        rec = vrec.KnnRecommender(rs, 0.5, 0.5, 50)
        # 256 targets = 2 target tiles x 32 candidate splits; 1024 = 8 x 32 = 256 CTAs on 148 SMs (two waves);
        # 18 944 = 148 tiles x 1 split (bench.py's batch, 512 bootstrap tiles).  The checked targets come first.
        rest = np.setdiff1d(inp.person_id, targets)
        for n_batch in (256, 1024, 18944):
            batch = np.concatenate([targets, rest[:n_batch - len(targets)]])
            for _ in range(2):                                  # the round-1 race showed up on some runs only
                got = rec.recommend(batch, flt, 10)
                assert _bad([g[:len(targets)] for g in got], want) == [], n_batch
        # small K: the bootstrap scratch must not depend on the heap size (K < 32 overran it)
        for k_small, nt in ((7, 64), (31, 64), (200, 32)):
            rc, *w2 = oracle.knn_query_batch(d, targets[:nt], 0.5, 0.5, k_small, flt, 10, n_threads=threads)
            assert rc == 0
            got = vrec.KnnRecommender(rs, 0.5, 0.5, k_small).recommend(targets[:nt], flt, 10)
            assert _bad(got, w2) == [], k_small
        # large K: K = 1 000 (fused top-K, fp32 tile filter), 3 000 and the launcher's
        # K = 2 000 000 (bin/knn_recommender.sh:32-36: every positive candidate is a neighbour) go through
        # the similarity-row / radix-select / column-scan kernels; few targets, the oracle pays ~0.1 s each
        for k_large, nt in ((1000, 8), (3000, 6), (2_000_000, 6)):
            rc, *w3 = oracle.knn_query_batch(d, targets[:nt], 0.5, 0.5, k_large, flt, 10, n_threads=threads)
            assert rc == 0
            got = vrec.KnnRecommender(rs, 0.5, 0.5, k_large).recommend(targets[:nt], flt, 10)
            assert _bad(got, w3) == [], k_large
    finally:
        rs.close()
        ctx.close()
