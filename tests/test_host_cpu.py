"""CPU-only checks of the product library: it loads, exports every symbol include/vrec.h declares,
fails loudly without a GPU, and its host-side preprocessing matches the oracle's."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "locations-recommender_b200", "libvrec.so")):
        g.build()
    from vrec import _lib
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "vrec.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vrec_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    from vrec import _lib
    for name in sorted(declared):
        assert hasattr(lib, name), f"libvrec.so does not export {name}"
        assert name in _lib.SIGNATURES, f"ctypes binding lacks {name}"
    assert lib.vrec_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import vrec
    with pytest.raises(vrec.VrecError) as e:
        vrec.Context()
    assert e.value.code == -19 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "locations-recommender_b200")
    for d, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(d, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("oracle's", "").replace(
                    "oracle (", "").replace("oracle/vrec_oracle.c", ""), f"{f} references the oracle"


def test_host_sg_csr_matches_oracle(lib, oracle):
    from vrec import synth
    from vrec.engine import host_sg_csr
    s, t, w = synth.random_stochastic_graph(200, 5, seed=11, hub_fraction=0.2)
    # duplicate edges must stay separate terms in file order
    s, t, w = np.concatenate([s, s[:50]]), np.concatenate([t, t[:50]]), np.concatenate([w, w[:50] * 0.5])
    ids, rowptr, src, ww = host_sg_csr(s, t, w)
    g = oracle.SgGraph(s, t, w)
    assert np.array_equal(ids, g.ids)
    assert rowptr[-1] == len(s) and np.all(np.diff(rowptr) >= 0)
    for r in range(len(ids)):
        seg = src[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(seg) >= 0)
    # same multiset of (target, source, weight) and stable duplicates
    idx = {int(v): i for i, v in enumerate(ids)}
    want = sorted(((idx[int(b)], idx[int(a)], e) for e, (a, b) in enumerate(zip(s, t))))
    got_w = ww
    assert np.array_equal(got_w, np.array([w[e] for (_, _, e) in want]))


def test_synth_shapes():
    from vrec import synth
    pl = synth.sample_places(300, seed=0)
    assert len(pl.id) == 300 and pl.id.min() == 40 and np.all(np.diff(pl.id) == 1)
    v = synth.sample_place_visits(pl, 0, persons_per_region=2000, correlated=False, seed=1)
    k = synth.build_rating_vectors(v)
    assert np.all(np.diff(k.person_id) > 0)
    assert k.person_id.min() >= synth.min_person_id(300)
    for r in range(len(k.person_id)):
        seg = k.place_col[k.place_rowptr[r]:k.place_rowptr[r + 1]]
        assert np.all(np.diff(seg) > 0)
    s, t, w = synth.build_stochastic_graph(v)
    u, inv = np.unique(s, return_inverse=True)
    np.testing.assert_allclose(np.bincount(inv, weights=w), 1.0, rtol=1e-12)
    # rank() keeps ties
    mask = synth._rank_filter(np.array([1, 1, 1, 1, 2]), np.array([5, 3, 3, 1, 9]), 2)
    assert mask.tolist() == [True, True, True, False, True]
