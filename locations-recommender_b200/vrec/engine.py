"""Host-side mirror of the reference recommender classes (Python over the C ABI)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


class VrecError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class NoSuchElement(ValueError):
    """IllegalArgumentException("No such person: ..") / ("No such vertex in the graph: ..")."""


def _check(rc: int) -> None:
    if rc == L.OK:
        return
    msg = L.last_error()
    if rc == L.EINVAL:
        raise ValueError(msg)               # Scala: IllegalArgumentException("requirement failed: ...")
    if rc == L.ENOENT:
        raise NoSuchElement(msg)
    raise VrecError(rc, msg)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Context:
    """One per process and GPU; owns the CUDA stream all kernels run on."""

    def __init__(self, device: int = -1):
        self.lib = L.load()
        self._h = L.vp()
        _check(self.lib.vrec_init(int(device), C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            self.lib.vrec_shutdown(self._h)
            self._h = L.vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self.lib.vrec_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.vrec_launch_count(self._h))

    def synchronize(self) -> None:
        _check(self.lib.vrec_synchronize(self._h))

    # ---- multi-GPU: one process per GPU
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        _check(self.lib.vrec_comm_unique_id(buf))
        return buf.raw

    def init_comm(self, rank: int, world: int, unique_id: bytes) -> None:
        if len(unique_id) != 128:
            raise ValueError("unique id must be 128 bytes")
        _check(self.lib.vrec_comm_init(self._h, int(rank), int(world), C.create_string_buffer(unique_id, 128)))

    @property
    def rank(self) -> int:
        return int(self.lib.vrec_comm_rank(self._h))

    @property
    def world(self) -> int:
        return int(self.lib.vrec_comm_world(self._h))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class KnnRegionSet:
    """Device-resident region-set: place / category rating vectors + place ratings.

    Mirrors the three DataFrames of KnnRecommenderMain.makeRecommendations
    (knn/KnnRecommenderMain.scala:53-57)."""

    def __init__(self, person_id, place_rowptr, place_col, place_val, place_dim,
                 cat_rowptr, cat_col, cat_val, cat_dim,
                 rating_person=None, rating_place=None, rating_value=None, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        lib = self.ctx.lib
        person_id = _i64(person_id)
        prp, pci, pv = _i64(place_rowptr), _i32(place_col), _f64(place_val)
        crp, cci, cv = _i64(cat_rowptr), _i32(cat_col), _f64(cat_val)
        if len(prp) != len(person_id) + 1 or len(crp) != len(person_id) + 1:
            raise ValueError("rowptr arrays must have P+1 entries")
        if rating_person is None:
            rp = rl = rv = None
            nr = 0
        else:
            rp, rl, rv = _i64(rating_person), _i64(rating_place), _i64(rating_value)
            nr = len(rp)
        self._h = L.vp()
        _check(lib.vrec_knn_load(self.ctx._h, len(person_id), _ptr(person_id, L.i64p),
                                 _ptr(prp, L.i64p), _ptr(pci, L.i32p), _ptr(pv, L.f64p), int(place_dim),
                                 _ptr(crp, L.i64p), _ptr(cci, L.i32p), _ptr(cv, L.f64p), int(cat_dim),
                                 nr, _ptr(rp, L.i64p), _ptr(rl, L.i64p), _ptr(rv, L.i64p), C.byref(self._h)))
        self.P = len(person_id)
        self.place_dim = int(place_dim)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.ctx.lib.vrec_knn_free(self._h)
            self._h = L.vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name: str, value: int) -> None:
        _check(self.ctx.lib.vrec_knn_set_option(self._h, name.encode(), int(value)))

    @property
    def resident_bytes(self) -> int:
        return int(self.ctx.lib.vrec_knn_resident_bytes(self._h))

    def person_ids(self) -> np.ndarray:
        out = np.zeros(max(1, self.P), dtype=np.int64)
        _check(self.ctx.lib.vrec_knn_person_ids(self._h, _ptr(out, L.i64p)))
        return out[:self.P]


class KnnRecommender:
    """knn/KnnRecommender.scala:9-25.  `require`s raise ValueError at construction."""

    def __init__(self, region_set: KnnRegionSet, placeWeight: float, categoryWeight: float, kNearest: int):
        if not (placeWeight > 0 and placeWeight < 1.0):
            raise ValueError(f"requirement failed: Place weight must be in the interval (0; 1): {placeWeight}")
        if not (categoryWeight > 0 and categoryWeight < 1.0):
            raise ValueError(f"requirement failed: Category weight must be in the interval (0; 1): {categoryWeight}")
        if not (placeWeight + categoryWeight == 1.0):
            raise ValueError(f"requirement failed: Sum of weights must be 1.0: place: {placeWeight}, "
                             f"category: {categoryWeight}")
        if not kNearest > 0:
            raise ValueError("requirement failed: K nearest must be positive")
        self.rs = region_set
        self.pw, self.cw, self.k = float(placeWeight), float(categoryWeight), int(kNearest)

    # --- the reference method: DataFrame(place_id, estimated_rating), unordered
    def makeRecommendations(self, personId: int):
        rs = self.rs
        cap = max(1, rs.place_dim + 1)
        cap = max(cap, 1 << 16)
        while True:
            pl = np.zeros(cap, dtype=np.int64)
            rt = np.zeros(cap, dtype=np.float64)
            cnt = C.c_int64(0)
            _check(rs.ctx.lib.vrec_knn_estimates(rs._h, int(personId), self.pw, self.cw, self.k,
                                                 _ptr(pl, L.i64p), _ptr(rt, L.f64p), cap, C.byref(cnt)))
            if cnt.value < cap:
                return pl[:cnt.value], rt[:cnt.value]
            cap *= 4

    # --- makeRecommendations + printRecommendations' region filter and top-N, batched
    def recommend(self, person_ids: Sequence[int], place_filter=None, max_recommendations: int = 10):
        rs = self.rs
        t = _i64(person_ids)
        n = len(t)
        m = max(1, int(max_recommendations))
        out_place = np.full((n, m), -1, dtype=np.int64)
        out_rating = np.zeros((n, m), dtype=np.float64)
        out_count = np.zeros(max(1, n), dtype=np.int32)
        out_status = np.zeros(max(1, n), dtype=np.int32)
        f = _i64(place_filter) if place_filter is not None else None
        _check(rs.ctx.lib.vrec_knn_query(rs._h, _ptr(t, L.i64p), n, self.pw, self.cw, self.k,
                                         _ptr(f, L.i64p), 0 if f is None else len(f), int(max_recommendations),
                                         _ptr(out_place, L.i64p), _ptr(out_rating, L.f64p),
                                         _ptr(out_count, L.i32p), _ptr(out_status, L.i32p)))
        return out_place, out_rating, out_count[:n], out_status[:n]

    def last_neighbours(self, t: int):
        """Debug: (person ids ascending, similarities) of target index t of the last recommend() pass."""
        ids = np.zeros(self.k, dtype=np.int64)
        sims = np.zeros(self.k, dtype=np.float64)
        cnt = C.c_int32(0)
        _check(self.rs.ctx.lib.vrec_knn_debug_last_neighbours(self.rs._h, int(t), self.k, _ptr(ids, L.i64p),
                                                              _ptr(sims, L.f64p), C.byref(cnt)))
        return ids[:cnt.value], sims[:cnt.value]

    def findSimilarPersons(self, personId: int):
        cap = self.k
        ids = np.zeros(cap, dtype=np.int64)
        sims = np.zeros(cap, dtype=np.float64)
        cnt = C.c_int32(0)
        _check(self.rs.ctx.lib.vrec_knn_neighbours(self.rs._h, int(personId), self.pw, self.cw, self.k,
                                                   _ptr(ids, L.i64p), _ptr(sims, L.f64p), cap, C.byref(cnt)))
        return ids[:cnt.value], sims[:cnt.value]

    def similarities(self, personId: int) -> np.ndarray:
        out = np.zeros(max(1, self.rs.P), dtype=np.float64)
        _check(self.rs.ctx.lib.vrec_knn_similarities(self.rs._h, int(personId), self.pw, self.cw, self.k,
                                                     _ptr(out, L.f64p)))
        return out[:self.rs.P]


class StochasticGraph:
    """Device-resident stochastic graph (CSR of P^T) built from (source_id, target_id, balanced_weight)."""

    def __init__(self, source_id, target_id, balanced_weight, ctx: Optional[Context] = None, _handle=None,
                 partitioned: bool = False):
        self.ctx = ctx or default_context()
        self._h = L.vp()
        if _handle is not None:
            self._h = _handle
        else:
            s, t, w = _i64(source_id), _i64(target_id), _f64(balanced_weight)
            if not (len(s) == len(t) == len(w)):
                raise ValueError("edge arrays differ in length")
            load = self.ctx.lib.vrec_sg_load_partitioned if partitioned else self.ctx.lib.vrec_sg_load
            _check(load(self.ctx._h, len(s), _ptr(s, L.i64p), _ptr(t, L.i64p), _ptr(w, L.f64p),
                        C.byref(self._h)))
        self.N = int(self.ctx.lib.vrec_sg_vertex_count(self._h))
        self.nnz = int(self.ctx.lib.vrec_sg_edge_count(self._h))

    @classmethod
    def generate(cls, n_vertices: int, out_degree: int, seed: int = 5, rank: int = 0, world: int = 1,
                 ctx: Optional[Context] = None) -> "StochasticGraph":
        ctx = ctx or default_context()
        h = L.vp()
        _check(ctx.lib.vrec_sg_generate(ctx._h, int(n_vertices), int(out_degree), int(seed), int(rank),
                                        int(world), C.byref(h)))
        return cls(None, None, None, ctx=ctx, _handle=h)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.ctx.lib.vrec_sg_free(self._h)
            self._h = L.vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def vertex_ids(self) -> np.ndarray:
        out = np.zeros(max(1, self.N), dtype=np.int64)
        _check(self.ctx.lib.vrec_sg_vertex_ids(self._h, _ptr(out, L.i64p)))
        return out[:self.N]

    def export_csr(self):
        rowptr = np.zeros(self.N + 1, dtype=np.int32)
        src = np.zeros(max(1, self.nnz), dtype=np.int32)
        w = np.zeros(max(1, self.nnz), dtype=np.float64)
        _check(self.ctx.lib.vrec_sg_export_csr(self._h, _ptr(rowptr, L.i32p), _ptr(src, L.i32p), _ptr(w, L.f64p)))
        return rowptr, src[:self.nnz], w[:self.nnz]

    def row_range(self):
        """Rows of P^T this process owns (the whole graph unless row-partitioned)."""
        lo, hi = C.c_int64(0), C.c_int64(0)
        _check(self.ctx.lib.vrec_sg_row_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def iterate_device(self, iterations: int) -> None:
        _check(self.ctx.lib.vrec_sg_iterate_device(self._h, int(iterations)))

    def set_option(self, name: str, value: int) -> None:
        """"batch": 0 / 1 (default) / 2; "batch_targets_per_cta": 0 (auto) / 1 / 2 -- see include/vrec.h."""
        _check(self.ctx.lib.vrec_sg_set_option(self._h, name.encode(), int(value)))

    def batch_info(self, what: int) -> int:
        return int(self.ctx.lib.vrec_sg_batch_info(self._h, int(what)))

    @property
    def resident_bytes(self) -> int:
        return int(self.ctx.lib.vrec_sg_resident_bytes(self._h))


class StochasticGraphGroup:
    """The `world` parts of ONE row-partitioned graph held by this process (vrec_sg_group_load): same sweep
    kernel, peer stores and residual slots as the one-process-per-GPU path, driven from one stream."""

    def __init__(self, source_id, target_id, balanced_weight, world: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        s, t, w = _i64(source_id), _i64(target_id), _f64(balanced_weight)
        self.world = int(world)
        self._hs = (L.vp * self.world)()
        _check(self.ctx.lib.vrec_sg_group_load(self.ctx._h, self.world, len(s), _ptr(s, L.i64p), _ptr(t, L.i64p),
                                               _ptr(w, L.f64p), self._hs))
        self.parts = [StochasticGraph(None, None, None, ctx=self.ctx, _handle=L.vp(self._hs[r]))
                      for r in range(self.world)]
        self.N = self.parts[0].N

    def stationary(self, vertexId: int, epsilon: float, maxIterations: int):
        """-> (x[world, N], iterations[world], converged[world], residual[world])"""
        x = np.zeros((self.world, max(1, self.N)), dtype=np.float64)
        it = np.zeros(self.world, dtype=np.int32)
        cv = np.zeros(self.world, dtype=np.int32)
        res = np.zeros(self.world, dtype=np.float64)
        _check(self.ctx.lib.vrec_sg_group_stationary(self._hs, self.world, int(vertexId), float(epsilon),
                                                     int(maxIterations), _ptr(x, L.f64p), _ptr(it, L.i32p),
                                                     _ptr(cv, L.i32p), _ptr(res, L.f64p)))
        return x[:, :self.N], it, cv, res

    def close(self) -> None:
        for p in self.parts:
            p.close()
        self.parts = []


class StochasticRecommender:
    """stochastic/StochasticRecommender.scala:28-71."""

    def __init__(self, graph: StochasticGraph, epsilon: float, maxIterations: int, verbose: bool = False):
        if not epsilon >= 0:
            raise ValueError("requirement failed: epsilon must be non-negative")
        if not maxIterations >= 0:
            raise ValueError("requirement failed: max iterations number must be non-negative")
        self.g = graph
        self.epsilon, self.max_it = float(epsilon), int(maxIterations)
        self.verbose = verbose
        self.last_iterations = 0
        self.last_converged = 0
        self.last_residual = float("nan")

    def _message(self, it: int, conv: int) -> None:
        if self.verbose:   # the println's of stochastic/StochasticRecommender.scala:94,100
            if conv:
                print(f"Converged in {it} iterations")
            else:
                print(f"Number of iterations {it} reached the maximum {self.max_it}")

    def stationary(self, vertexId: int) -> np.ndarray:
        g = self.g
        x = np.zeros(max(1, g.N), dtype=np.float64)
        it, cv, res = C.c_int32(0), C.c_int32(0), C.c_double(0)
        _check(g.ctx.lib.vrec_sg_stationary(g._h, int(vertexId), self.epsilon, self.max_it, _ptr(x, L.f64p),
                                            C.byref(it), C.byref(cv), C.byref(res)))
        self.last_iterations, self.last_converged, self.last_residual = it.value, cv.value, res.value
        self._message(it.value, cv.value)
        return x[:g.N]

    # --- the reference method: DataFrame(id, probability) with id != vertex and probability > 0
    def makeRecommendations(self, vertexId: int):
        x = self.stationary(vertexId)
        ids = self.g.vertex_ids()
        keep = (ids != int(vertexId)) & (x > 0)
        return ids[keep], x[keep]

    # --- makeRecommendations + printRecommendations' join with the region's places and top-N, batched
    def recommend(self, vertex_ids: Sequence[int], place_filter=None, max_recommendations: int = 10):
        g = self.g
        v = _i64(vertex_ids)
        n = len(v)
        m = max(1, int(max_recommendations))
        out_id = np.full((n, m), -1, dtype=np.int64)
        out_prob = np.zeros((n, m), dtype=np.float64)
        cnt = np.zeros(max(1, n), dtype=np.int32)
        its = np.zeros(max(1, n), dtype=np.int32)
        conv = np.zeros(max(1, n), dtype=np.int32)
        status = np.zeros(max(1, n), dtype=np.int32)
        f = _i64(place_filter) if place_filter is not None else None
        _check(g.ctx.lib.vrec_sg_query(g._h, _ptr(v, L.i64p), n, self.epsilon, self.max_it,
                                       _ptr(f, L.i64p), 0 if f is None else len(f), int(max_recommendations),
                                       _ptr(out_id, L.i64p), _ptr(out_prob, L.f64p), _ptr(cnt, L.i32p),
                                       _ptr(its, L.i32p), _ptr(conv, L.i32p), _ptr(status, L.i32p)))
        if self.verbose:
            for q in range(n):
                if status[q] == L.OK:
                    self._message(int(its[q]), int(conv[q]))
        return out_id, out_prob, cnt[:n], its[:n], conv[:n], status[:n]


def host_sg_partition(rowptr, world: int) -> np.ndarray:
    """Row ranges of a row-partitioned graph (balanced by in-edges): bounds[world + 1]."""
    lib = L.load()
    rp = _i32(rowptr)
    out = np.zeros(world + 1, dtype=np.int64)
    _check(lib.vrec_host_sg_partition(len(rp) - 1, _ptr(rp, L.i32p), int(world), _ptr(out, L.i64p)))
    return out


def host_sg_csr(source_id, target_id, balanced_weight):
    """The CSR of P^T as vrec_sg_load builds it, computed on the host (no device needed)."""
    lib = L.load()
    s, t, w = _i64(source_id), _i64(target_id), _f64(balanced_weight)
    nnz = len(s)
    n = C.c_int64(0)
    ids = np.zeros(max(1, 2 * nnz), dtype=np.int64)
    rowptr = np.zeros(2 * nnz + 1, dtype=np.int32)
    src = np.zeros(max(1, nnz), dtype=np.int32)
    ww = np.zeros(max(1, nnz), dtype=np.float64)
    _check(lib.vrec_host_sg_csr(nnz, _ptr(s, L.i64p), _ptr(t, L.i64p), _ptr(w, L.f64p), C.byref(n),
                                _ptr(ids, L.i64p), _ptr(rowptr, L.i32p), _ptr(src, L.i32p), _ptr(ww, L.f64p)))
    N = n.value
    return ids[:N], rowptr[:N + 1], src[:nnz], ww[:nnz]
