"""ctypes front-end of the CPU oracle (oracle/vrec_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the
product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvrec_oracle.so")

OK, ENOENT, EINVAL = 0, -2, -22


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vrec_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libvrec_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.vro_vector_length.restype = C.c_double
        _lib.vro_sparse_dot.restype = C.c_double
        _lib.vro_cosine.restype = C.c_double
        _lib.vro_sg_vertex_count.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- Distance
def vector_length(values) -> float:
    v = _f64(values)
    return lib().vro_vector_length(_p(v, C.c_double), C.c_int32(len(v)))


def cosine(xi, xv, yi, yv) -> float:
    xi, xv, yi, yv = _i32(xi), _f64(xv), _i32(yi), _f64(yv)
    return lib().vro_cosine(_p(xi, C.c_int32), _p(xv, C.c_double), C.c_int32(len(xi)),
                            _p(yi, C.c_int32), _p(yv, C.c_double), C.c_int32(len(yi)))


def sparse_dot(xi, xv, yi, yv) -> float:
    xi, xv, yi, yv = _i32(xi), _f64(xv), _i32(yi), _f64(yv)
    return lib().vro_sparse_dot(_p(xi, C.c_int32), _p(xv, C.c_double), C.c_int32(len(xi)),
                                _p(yi, C.c_int32), _p(yv, C.c_double), C.c_int32(len(yi)))


# ---------------------------------------------------------------- KNN
class KnnData:
    """Region-set in CSR form.  persons ascending; ratings grouped by person."""

    def __init__(self, person_id, place_rowptr, place_col, place_val,
                 cat_rowptr, cat_col, cat_val,
                 rat_rowptr=None, rat_place=None, rat_val=None, place_dim=None):
        self.person_id = _i64(person_id)
        self.prp, self.pci, self.pv = _i64(place_rowptr), _i32(place_col), _f64(place_val)
        self.crp, self.cci, self.cv = _i64(cat_rowptr), _i32(cat_col), _f64(cat_val)
        if rat_rowptr is None:  # place_ratings == the non-zeros of the place vectors
            rat_rowptr, rat_place, rat_val = self.prp, self.pci.astype(np.int64), self.pv.astype(np.int64)
        self.rrp, self.rpl, self.rv = _i64(rat_rowptr), _i64(rat_place), _i64(rat_val)
        if place_dim is None:
            place_dim = int(max(self.rpl.max(initial=-1), self.pci.max(initial=-1))) + 1
        self.place_dim = int(place_dim)
        self.P = len(self.person_id)

    def _vec_args(self):
        return (C.c_int64(self.P), _p(self.person_id, C.c_int64),
                _p(self.prp, C.c_int64), _p(self.pci, C.c_int32), _p(self.pv, C.c_double),
                _p(self.crp, C.c_int64), _p(self.cci, C.c_int32), _p(self.cv, C.c_double))

    def _rat_args(self):
        return (_p(self.rrp, C.c_int64), _p(self.rpl, C.c_int64), _p(self.rv, C.c_int64),
                C.c_int64(self.place_dim))


def knn_similarities(d: KnnData, target, pw, cw):
    out = np.zeros(d.P, dtype=np.float64)
    rc = lib().vro_knn_similarities(*d._vec_args(), C.c_double(pw), C.c_double(cw),
                                    C.c_int64(int(target)), _p(out, C.c_double))
    return rc, out


def knn_neighbours(d: KnnData, target, pw, cw, k):
    n = max(1, min(int(k), d.P))
    ids = np.zeros(n, dtype=np.int64)
    sims = np.zeros(n, dtype=np.float64)
    cnt = C.c_int32(0)
    rc = lib().vro_knn_neighbours(*d._vec_args(), C.c_double(pw), C.c_double(cw), C.c_int32(int(k)),
                                  C.c_int64(int(target)), _p(ids, C.c_int64), _p(sims, C.c_double),
                                  C.byref(cnt))
    return rc, ids[:cnt.value], sims[:cnt.value]


def knn_estimates(d: KnnData, target, pw, cw, k):
    """Raw KnnRecommender.makeRecommendations rows, sorted by place id."""
    pl = np.zeros(max(1, d.place_dim), dtype=np.int64)
    rt = np.zeros(max(1, d.place_dim), dtype=np.float64)
    cnt = C.c_int64(0)
    rc = lib().vro_knn_estimates(*d._vec_args(), *d._rat_args(), C.c_int64(int(target)),
                                 C.c_double(pw), C.c_double(cw), C.c_int32(int(k)),
                                 _p(pl, C.c_int64), _p(rt, C.c_double), C.byref(cnt))
    return rc, pl[:cnt.value], rt[:cnt.value]


def knn_query_batch(d: KnnData, targets, pw, cw, k, place_filter, max_recs, n_threads=0):
    targets = _i64(targets)
    nt = len(targets)
    m = max(1, int(max_recs))
    out_place = np.full((nt, m), -1, dtype=np.int64)
    out_rating = np.zeros((nt, m), dtype=np.float64)
    out_count = np.zeros(nt, dtype=np.int32)
    out_status = np.zeros(nt, dtype=np.int32)
    if place_filter is None:
        fp, nf = None, 0
    else:
        place_filter = _i64(place_filter)
        fp, nf = _p(place_filter, C.c_int64), len(place_filter)
    rc = lib().vro_knn_query_batch(*d._vec_args(), *d._rat_args(),
                                   _p(targets, C.c_int64), C.c_int64(nt),
                                   C.c_double(pw), C.c_double(cw), C.c_int32(int(k)),
                                   fp, C.c_int64(nf), C.c_int32(int(max_recs)),
                                   _p(out_place, C.c_int64), _p(out_rating, C.c_double),
                                   _p(out_count, C.c_int32), _p(out_status, C.c_int32),
                                   C.c_int32(int(n_threads)))
    return rc, out_place, out_rating, out_count, out_status


# ---------------------------------------------------------------- SG
class SgGraph:
    @classmethod
    def from_csr(cls, rowptr, src, weight):
        """CSR of P^T over vertex ids 0..N-1 (rows = targets, ascending sources)."""
        rowptr, src, weight = _i64(rowptr), _i32(src), _f64(weight)
        self = cls.__new__(cls)
        self._h = C.c_void_p(0)
        rc = lib().vro_sg_from_csr(C.c_int64(len(rowptr) - 1), _p(rowptr, C.c_int64), _p(src, C.c_int32),
                                   _p(weight, C.c_double), C.byref(self._h))
        if rc:
            raise MemoryError(rc)
        self.N = len(rowptr) - 1
        self.ids = np.arange(self.N, dtype=np.int64)
        self.nnz = len(src)
        return self

    def __init__(self, source, target, weight):
        source, target, weight = _i64(source), _i64(target), _f64(weight)
        assert len(source) == len(target) == len(weight)
        self._h = C.c_void_p(0)
        rc = lib().vro_sg_build(C.c_int64(len(source)), _p(source, C.c_int64), _p(target, C.c_int64),
                                _p(weight, C.c_double), C.byref(self._h))
        if rc:
            raise MemoryError(rc)
        self.N = lib().vro_sg_vertex_count(self._h)
        self.ids = np.zeros(self.N, dtype=np.int64)
        lib().vro_sg_vertex_ids(self._h, _p(self.ids, C.c_int64))
        self.nnz = len(source)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value and _lib is not None:
                _lib.vro_sg_free(self._h)
                self._h = C.c_void_p(0)
        except Exception:
            pass

    def run(self, vertex, epsilon, max_iterations):
        """-> rc, x[N], iterations, converged, last_residual"""
        x = np.zeros(max(1, self.N), dtype=np.float64)
        it, cv, res = C.c_int32(0), C.c_int32(0), C.c_double(0)
        rc = lib().vro_sg_run(self._h, C.c_int64(int(vertex)), C.c_double(epsilon),
                              C.c_int32(int(max_iterations)), _p(x, C.c_double),
                              C.byref(it), C.byref(cv), C.byref(res))
        return rc, x[:self.N], it.value, cv.value, res.value

    def query(self, vertex, epsilon, max_iterations, place_filter, max_recs):
        m = max(1, int(max_recs))
        ids = np.zeros(m, dtype=np.int64)
        pr = np.zeros(m, dtype=np.float64)
        cnt, it, cv = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        if place_filter is None:
            fp, nf = None, 0
        else:
            place_filter = _i64(place_filter)
            fp, nf = _p(place_filter, C.c_int64), len(place_filter)
        rc = lib().vro_sg_query(self._h, C.c_int64(int(vertex)), C.c_double(epsilon),
                                C.c_int32(int(max_iterations)), fp, C.c_int64(nf),
                                C.c_int32(int(max_recs)), _p(ids, C.c_int64), _p(pr, C.c_double),
                                C.byref(cnt), C.byref(it), C.byref(cv))
        return rc, ids[:cnt.value], pr[:cnt.value], it.value, cv.value


def set_threads(n: int) -> None:
    lib().vro_set_threads(C.c_int32(int(n)))


def num_threads() -> int:
    return int(lib().vro_num_threads())


# ---------------------------------------------------------------- rating vectors builder
def build_rating_vectors(person_id, entity_id, weight, top_n):
    """-> rc, persons, rowptr, col, val, dim  (RatingsBuilder.calcRatings + RatingVectorsBuilder.calcRatingVectors)"""
    person_id, entity_id = _i64(person_id), _i64(entity_id)
    n = len(person_id)
    w = _i64(weight) if weight is not None else None
    persons = np.zeros(max(1, n), dtype=np.int64)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    col = np.zeros(max(1, n), dtype=np.int32)
    val = np.zeros(max(1, n), dtype=np.float64)
    P, nnz, dim = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    rc = lib().vro_build_rating_vectors(C.c_int64(n), _p(person_id, C.c_int64), _p(entity_id, C.c_int64),
                                        _p(w, C.c_int64) if w is not None else None, C.c_int32(int(top_n)),
                                        C.byref(P), C.byref(nnz), _p(persons, C.c_int64), _p(rowptr, C.c_int64),
                                        _p(col, C.c_int32), _p(val, C.c_double), C.byref(dim))
    return rc, persons[:P.value], rowptr[:P.value + 1], col[:nnz.value], val[:nnz.value], dim.value


def build_edge_family(source_id, target_id, weight, top_n, beta):
    s, t = _i64(source_id), _i64(target_id)
    n = len(s)
    w = _i64(weight) if weight is not None else None
    cap = max(1, n)
    os_, ot, ow = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.float64)
    ne = C.c_int64(0)
    rc = lib().vro_build_edge_family(C.c_int64(n), _p(s, C.c_int64), _p(t, C.c_int64),
                                     _p(w, C.c_int64) if w is not None else None, C.c_int32(int(top_n)),
                                     C.c_double(beta), C.c_int64(cap), C.byref(ne), _p(os_, C.c_int64),
                                     _p(ot, C.c_int64), _p(ow, C.c_double))
    return rc, os_[:ne.value], ot[:ne.value], ow[:ne.value]


def build_stochastic_graph(person_id, place_id, category_id, timestamp_ms, beta_person_place=0.5,
                           beta_person_category=0.5):
    pe, pl, ca, ts = _i64(person_id), _i64(place_id), _i64(category_id), _i64(timestamp_ms)
    n = len(pe)
    ne = C.c_int64(0)
    args = (C.c_int64(n), _p(pe, C.c_int64), _p(pl, C.c_int64), _p(ca, C.c_int64), _p(ts, C.c_int64),
            C.c_double(beta_person_place), C.c_double(beta_person_category))
    rc = lib().vro_build_stochastic_graph(*args, C.c_int64(0), C.byref(ne), None, None, None)
    cap = max(1, ne.value)
    os_, ot, ow = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.float64)
    rc = lib().vro_build_stochastic_graph(*args, C.c_int64(cap), C.byref(ne), _p(os_, C.c_int64), _p(ot, C.c_int64),
                                          _p(ow, C.c_double))
    return rc, os_[:ne.value], ot[:ne.value], ow[:ne.value]


def build_place_visits(person_id, latitude, longitude, timestamp_ms, region_id, place_id, place_latitude,
                       place_longitude, place_category, place_region, last_days_count=7, accuracy_m=100.0):
    """-> rc, (person, timestamp_ms, place, region, category), margins, closest_miss"""
    pe, ts, rg = _i64(person_id), _i64(timestamp_ms), _i64(region_id)
    la, lo = _f64(latitude), _f64(longitude)
    pi, pc, pr = _i64(place_id), _i64(place_category), _i64(place_region)
    pla, plo = _f64(place_latitude), _f64(place_longitude)
    n, m = len(pe), len(pi)
    cnt, miss = C.c_int64(0), C.c_double(0)
    fn = lib().vro_build_place_visits
    args = (C.c_int64(n), _p(pe, C.c_int64), _p(la, C.c_double), _p(lo, C.c_double), _p(ts, C.c_int64), _p(rg, C.c_int64),
            C.c_int64(m), _p(pi, C.c_int64), _p(pla, C.c_double), _p(plo, C.c_double), _p(pc, C.c_int64), _p(pr, C.c_int64),
            C.c_int32(int(last_days_count)), C.c_double(accuracy_m))
    fn(*args, C.c_int64(0), C.byref(cnt), None, None, None, None, None, None, C.byref(miss))
    cap = max(1, cnt.value)
    out = [np.zeros(cap, dtype=np.int64) for _ in range(5)]
    margin = np.zeros(cap, dtype=np.float64)
    rc = fn(*args, C.c_int64(cap), C.byref(cnt), *[_p(o, C.c_int64) for o in out], _p(margin, C.c_double), C.byref(miss))
    k = cnt.value
    return rc, tuple(o[:k] for o in out), margin[:k], miss.value
