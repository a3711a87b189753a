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
from .engine import Context, _check, _i64, _ptr, default_context
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
