"""Device builders in front of the recommenders.

rating_vectors_builder: RatingVectorsBuilderMain.generateRegionRatingVectors
(knn/RatingVectorsBuilderMain.scala:41-73) for one region-set of place visits, through
vrec_build_rating_vectors (RatingsBuilder.calcRatings + RatingVectorsBuilder.calcRatingVectors on the GPU).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from .engine import Context, _check, _f64, _i64, _ptr, default_context
from .synth import KnnInputs


def build_rating_vectors(person_id, entity_id, top_n: int, weight=None, ctx: Optional[Context] = None):
    """-> persons (ascending), rowptr, col, val, dim  -- one entity column ("place_id" or "category_id")."""
    ctx = ctx or default_context()
    p, e = _i64(person_id), _i64(entity_id)
    if len(p) != len(e):
        raise ValueError("person_id and entity_id differ in length")
    w = _i64(weight) if weight is not None else None
    n = len(p)
    persons = np.zeros(max(1, n), dtype=np.int64)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    col = np.zeros(max(1, n), dtype=np.int32)
    val = np.zeros(max(1, n), dtype=np.float64)
    P, nnz, dim = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    _check(ctx.lib.vrec_build_rating_vectors(ctx._h, n, _ptr(p, L.i64p), _ptr(e, L.i64p), _ptr(w, L.i64p), int(top_n),
                                             C.byref(P), C.byref(nnz), _ptr(persons, L.i64p), _ptr(rowptr, L.i64p),
                                             _ptr(col, L.i32p), _ptr(val, L.f64p), C.byref(dim)))
    return persons[:P.value], rowptr[:P.value + 1], col[:nnz.value], val[:nnz.value], int(dim.value)


def rating_vectors_builder(person_id, place_id, category_id, weight=None, max_rated_places: int = 100,
                           max_rated_categories: int = 10, ctx: Optional[Context] = None) -> KnnInputs:
    """Place visits of one region-set -> the three tables the KNN recommender loads
    (bin/rating_vectors_builder.sh:32-34 defaults; knn/RatingVectorsBuilderMain.scala:41-73)."""
    persons, prp, pci, pv, place_dim = build_rating_vectors(person_id, place_id, max_rated_places, weight, ctx)
    persons_c, crp, cci, cv, cat_dim = build_rating_vectors(person_id, category_id, max_rated_categories, weight, ctx)
    assert np.array_equal(persons, persons_c)          # every visit row has a place and a category
    rat_person = np.repeat(persons, np.diff(prp))      # place_ratings = the non-zeros of the place vectors (:44-49,:71)
    return KnnInputs(persons, prp, pci, pv, place_dim, crp, cci, cv, cat_dim,
                     rat_person.astype(np.int64), pci.astype(np.int64), pv.astype(np.int64))


def build_edge_family(source_id, target_id, top_n: int, beta: float, weight=None, ctx: Optional[Context] = None):
    """One balanced edge family (source_id, target_id, balanced_weight), sorted by (source, target)."""
    ctx = ctx or default_context()
    s_, t_ = _i64(source_id), _i64(target_id)
    w = _i64(weight) if weight is not None else None
    n = len(s_)
    cap = max(1, n)
    os_, ot, ow = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.float64)
    ne = C.c_int64(0)
    _check(ctx.lib.vrec_build_edge_family(ctx._h, n, _ptr(s_, L.i64p), _ptr(t_, L.i64p), _ptr(w, L.i64p), int(top_n),
                                          float(beta), cap, C.byref(ne), _ptr(os_, L.i64p), _ptr(ot, L.i64p),
                                          _ptr(ow, L.f64p)))
    return os_[:ne.value], ot[:ne.value], ow[:ne.value]


def stochastic_graph_builder(person_id, place_id, category_id, timestamp_ms, beta_person_place: float = 0.5,
                             beta_person_category: float = 0.5, ctx: Optional[Context] = None):
    """Place visits of one region-set -> the stochastic graph's edges (source_id, target_id, balanced_weight)
    (bin/stochastic_graph_builder.sh:32-34 defaults; stochastic/StochasticGraphBuilderMain.scala:47-66)."""
    ctx = ctx or default_context()
    pe, pl, ca, ts = _i64(person_id), _i64(place_id), _i64(category_id), _i64(timestamp_ms)
    n = len(pe)
    if not (len(pl) == len(ca) == len(ts) == n):
        raise ValueError("visit columns differ in length")
    ne = C.c_int64(0)
    args = (ctx._h, n, _ptr(pe, L.i64p), _ptr(pl, L.i64p), _ptr(ca, L.i64p), _ptr(ts, L.i64p),
            float(beta_person_place), float(beta_person_category))
    rc = ctx.lib.vrec_build_stochastic_graph(*args, 0, C.byref(ne), None, None, None)      # sizes the outputs
    if rc not in (L.OK, L.ENOMEM):
        _check(rc)
    cap = max(1, ne.value)
    os_, ot, ow = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.float64)
    _check(ctx.lib.vrec_build_stochastic_graph(*args, cap, C.byref(ne), _ptr(os_, L.i64p), _ptr(ot, L.i64p),
                                               _ptr(ow, L.f64p)))
    return os_[:ne.value], ot[:ne.value], ow[:ne.value]


def place_visits_builder(person_id, latitude, longitude, timestamp_ms, region_id, place_id, place_latitude,
                         place_longitude, place_category, place_region, last_days_count: int = 7,
                         accuracy_m: float = 100.0, ctx: Optional[Context] = None):
    """Location visits x places -> place visits (person_id, timestamp_ms, place_id, region_id, category_id)
    (PlaceVisits.calcPlaceVisits, PlaceVisits.scala:11-48; --last-days-count 7, 100 m)."""
    ctx = ctx or default_context()
    pe, ts, rg = _i64(person_id), _i64(timestamp_ms), _i64(region_id)
    la, lo = _f64(latitude), _f64(longitude)
    pi, pc, pr = _i64(place_id), _i64(place_category), _i64(place_region)
    pla, plo = _f64(place_latitude), _f64(place_longitude)
    n, m = len(pe), len(pi)
    cnt = C.c_int64(0)
    args = (ctx._h, n, _ptr(pe, L.i64p), _ptr(la, L.f64p), _ptr(lo, L.f64p), _ptr(ts, L.i64p), _ptr(rg, L.i64p), m,
            _ptr(pi, L.i64p), _ptr(pla, L.f64p), _ptr(plo, L.f64p), _ptr(pc, L.i64p), _ptr(pr, L.i64p),
            int(last_days_count), float(accuracy_m))
    rc = ctx.lib.vrec_build_place_visits(*args, 0, C.byref(cnt), None, None, None, None, None)
    if rc not in (L.OK, L.ENOMEM):
        _check(rc)
    cap = max(1, cnt.value)
    out = [np.zeros(cap, dtype=np.int64) for _ in range(5)]
    _check(ctx.lib.vrec_build_place_visits(*args, cap, C.byref(cnt), *[_ptr(o, L.i64p) for o in out]))
    return tuple(o[:cnt.value] for o in out)
