"""Synthetic inputs of the two hot paths, in the shapes the reference's tools produce.

Nothing here is on the measured path: these are numpy restatements of
  * sample-generator (ids, regions, grid of places, visits/person, 7-day window):
      sample-generator/.../SampleGeneratorMain.scala:7-37, PlacesSampleGenerator.scala:50-71,
      LocationVisitsSampleGenerator.scala:23-31,56-68,84-128, PlaceVisits.scala:11-61
  * the KNN input builders (rank() <= topN with ties, SparseVector layout):
      knn/RatingsBuilder.scala:32-48, knn/RatingVectorsBuilder.scala:26-83
  * the SG input builders (four edge families, row-normalised, beta-balanced):
      stochastic/{PersonLikesPlace,PersonLikesCategory,PlaceSimilarPlace,CategorySelectedPlace}.scala,
      stochastic/StochasticGraphBuilder.scala:8-28, StochasticGraphBuilderMain.scala:47-66
Spark's rand(seed) stream cannot be reproduced without a JVM, so values differ from a real
sample_generator run; ids, shapes, distributions and file formats are the same.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# SampleGeneratorMain.scala:7-11
REGIONS = [
    (0, "Moscow", 55.623920, 55.823685, 37.404277, 37.795022),
    (1, "Peterburg", 59.857032, 60.006462, 30.196832, 30.490272),
    (2, "Kazan", 55.744243, 55.835127, 49.024581, 49.231314),
]
N_CATEGORIES = 20                       # SampleGeneratorMain.scala:13-34
MIN_CATEGORY_ID = 0
MIN_PLACE_ID = MIN_CATEGORY_ID + N_CATEGORIES * 2          # :36-37
EARTH_RADIUS_M = 6371 * 1000.0          # Location.scala:29


def min_person_id(place_count: int) -> int:
    return MIN_PLACE_ID + place_count * 2                   # SampleGeneratorMain.scala:54


@dataclass
class Places:
    id: np.ndarray
    latitude: np.ndarray
    longitude: np.ndarray
    category_id: np.ndarray
    region_id: np.ndarray

    def of_region(self, region: int) -> np.ndarray:
        return self.id[self.region_id == region]


def sample_places(place_count: int = 30000, seed: int = 0) -> Places:
    """PlacesSampleGenerator.generatePlaces: floor(sqrt(n))^2 grid per region, category = floor(U*20)."""
    rng = np.random.default_rng([seed, 101])
    per_region = place_count // len(REGIONS)
    side = int(np.floor(np.sqrt(per_region)))
    ids, lat, lon, reg = [], [], [], []
    for (rid, _name, lat0, lat1, lon0, lon1) in REGIONS:
        lat_step, lon_step = (lat1 - lat0) / side, (lon1 - lon0) / side
        li, lj = np.meshgrid(np.arange(1, side + 1), np.arange(1, side + 1), indexing="ij")
        idx = np.arange(side * side)
        ids.append(MIN_PLACE_ID + rid * side * side + idx)
        lat.append((lat0 + lat_step * li).ravel())
        lon.append((lon0 + lon_step * lj).ravel())
        reg.append(np.full(side * side, rid))
    ids = np.concatenate(ids).astype(np.int64)
    cat = (MIN_CATEGORY_ID + np.floor(rng.random(len(ids)) * N_CATEGORIES)).astype(np.int64)
    return Places(ids, np.concatenate(lat), np.concatenate(lon), cat, np.concatenate(reg).astype(np.int64))


def _haversine_m(lat1, lon1, lat2, lon2):
    lat1, lon1, lat2, lon2 = map(np.radians, (lat1, lon1, lat2, lon2))
    hav = np.sin((lat2 - lat1) / 2) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) / 2) ** 2
    return EARTH_RADIUS_M * 2 * np.arcsin(np.sqrt(hav))


@dataclass
class PlaceVisitCounts:
    """count(*) per (person_id, place_id) inside the last-days window, plus place categories."""
    person_id: np.ndarray
    place_id: np.ndarray
    count: np.ndarray
    category_id: np.ndarray     # category of place_id, per row


def sample_place_visits(places: Places, region: int, persons_per_region: int = 1_000_000,
                        person_count_total: int | None = None, place_count_total: int | None = None,
                        correlated: bool = True, last_days: int = 7, seed: int = 0) -> PlaceVisitCounts:
    """Visits of one region's persons that fall in the last `last_days` days and within 100 m of a place.

    Per person: floor(U*365)+1 visits over 8736 h (LocationVisitsSampleGenerator.scala:23-31,88-90);
    only the window fraction is materialised (binomial thinning; statistically the same as
    generating all visits and filtering, PlaceVisits.scala:50-61).  correlated=True reproduces the
    reference's use of rand(0) for latitude, longitude AND time (one factor per visit, so visits
    lie on the region's diagonal); correlated=False draws them independently."""
    rng = np.random.default_rng([seed, 202, region])
    n_regions = len(REGIONS)
    place_count_total = place_count_total or len(places.id)
    person_count_total = person_count_total or persons_per_region * n_regions
    pid0 = min_person_id(place_count_total) + region * (person_count_total // n_regions)
    hours = 8736
    window_lo = (hours - 1 - 24 * last_days) / hours            # factor whose offset is max - 7 days
    p_window = 1.0 - window_lo
    n_total = np.floor(rng.random(persons_per_region) * 365).astype(np.int64) + 1
    n_win = rng.binomial(n_total, p_window)
    person = np.repeat(np.arange(persons_per_region, dtype=np.int64) + pid0, n_win)
    nv = len(person)
    f_time = window_lo + rng.random(nv) * p_window
    if correlated:
        f_lat = f_lon = f_time
    else:
        f_lat, f_lon = rng.random(nv), rng.random(nv)
    (_rid, _n, lat0, lat1, lon0, lon1) = REGIONS[region]
    vlat = lat0 + (lat1 - lat0) * f_lat
    vlon = lon0 + (lon1 - lon0) * f_lon
    sel = places.region_id == region
    p_ids, p_lat, p_lon, p_cat = places.id[sel], places.latitude[sel], places.longitude[sel], places.category_id[sel]
    side = int(round(np.sqrt(len(p_ids))))
    gi = np.rint(f_lat * side).astype(np.int64)
    gj = np.rint(f_lon * side).astype(np.int64)
    out_person, out_place = [], []
    for di in (-1, 0, 1):           # 3x3 neighbourhood of the nearest grid node
        for dj in (-1, 0, 1):
            i, j = gi + di, gj + dj
            ok = (i >= 1) & (i <= side) & (j >= 1) & (j <= side)
            k = (i - 1) * side + (j - 1)
            k = np.where(ok, k, 0)
            d = _haversine_m(vlat, vlon, p_lat[k], p_lon[k])
            hit = ok & (d <= 100.0)                                # PlaceVisits.scala:22,127
            out_person.append(person[hit])
            out_place.append(k[hit])
    person = np.concatenate(out_person)
    pk = np.concatenate(out_place)
    key = person * np.int64(len(p_ids)) + pk
    uniq, cnt = np.unique(key, return_counts=True)
    person_u = uniq // len(p_ids)
    k_u = uniq % len(p_ids)
    return PlaceVisitCounts(person_u, p_ids[k_u], cnt.astype(np.int64), p_cat[k_u])


def merge_visits(parts) -> PlaceVisitCounts:
    """Region-set of several regions (PlaceVisits.extractRegionsPlaceVisits, PlaceVisits.scala:63-88)."""
    return PlaceVisitCounts(*(np.concatenate([getattr(p, f) for p in parts])
                              for f in ("person_id", "place_id", "count", "category_id")))


def _rank_filter(group: np.ndarray, count: np.ndarray, top_n: int) -> np.ndarray:
    """rank() over (partition by group order by count desc) <= top_n, ties kept
    (knn/RatingsBuilder.scala:43-47).  Returns a boolean mask."""
    order = np.lexsort((-count, group))
    g, c = group[order], count[order]
    idx = np.arange(len(g))
    new_group = np.ones(len(g), dtype=bool)
    new_group[1:] = g[1:] != g[:-1]
    new_val = new_group.copy()
    new_val[1:] |= c[1:] != c[:-1]
    group_start = np.maximum.accumulate(np.where(new_group, idx, 0))
    first_same = np.maximum.accumulate(np.where(new_val, idx, 0))
    rank = first_same - group_start + 1
    mask = np.zeros(len(g), dtype=bool)
    mask[order] = rank <= top_n
    return mask


def _to_csr(person: np.ndarray, entity: np.ndarray, value: np.ndarray, persons_sorted: np.ndarray):
    order = np.lexsort((entity, person))
    p, e, v = person[order], entity[order], value[order]
    row = np.searchsorted(persons_sorted, p)
    rowptr = np.zeros(len(persons_sorted) + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    np.cumsum(rowptr, out=rowptr)
    return rowptr, e.astype(np.int32), v.astype(np.float64)


@dataclass
class KnnInputs:
    person_id: np.ndarray
    place_rowptr: np.ndarray
    place_col: np.ndarray
    place_val: np.ndarray
    place_dim: int
    cat_rowptr: np.ndarray
    cat_col: np.ndarray
    cat_val: np.ndarray
    cat_dim: int
    rating_person: np.ndarray
    rating_place: np.ndarray
    rating_value: np.ndarray

    def load_args(self):
        return (self.person_id, self.place_rowptr, self.place_col, self.place_val, self.place_dim,
                self.cat_rowptr, self.cat_col, self.cat_val, self.cat_dim,
                self.rating_person, self.rating_place, self.rating_value)

    @property
    def algorithmic_bytes(self) -> int:
        """B_region = 2*4*(P+1) + 12*nnz_place + 12*nnz_cat + 2*8*P + 8*P."""
        P = len(self.person_id)
        return 8 * (P + 1) + 12 * len(self.place_col) + 12 * len(self.cat_col) + 24 * P


def build_rating_vectors(v: PlaceVisitCounts, max_rated_places: int = 100,
                         max_rated_categories: int = 10) -> KnnInputs:
    """RatingVectorsBuilderMain.generateRegionRatingVectors (knn/RatingVectorsBuilderMain.scala:41-73)."""
    keep = _rank_filter(v.person_id, v.count, max_rated_places)
    pp, pl, pc = v.person_id[keep], v.place_id[keep], v.count[keep]
    ckey = v.person_id * np.int64(1 << 20) + v.category_id
    cu, inv = np.unique(ckey, return_inverse=True)
    ccount = np.bincount(inv, weights=v.count.astype(np.float64)).astype(np.int64)
    cperson, ccat = cu >> 20, cu & ((1 << 20) - 1)
    ckeep = _rank_filter(cperson, ccount, max_rated_categories)
    cperson, ccat, ccount = cperson[ckeep], ccat[ckeep], ccount[ckeep]
    persons = np.unique(np.concatenate([pp, cperson]))
    prp, pci, pv = _to_csr(pp, pl, pc, persons)
    crp, cci, cv = _to_csr(cperson, ccat, ccount, persons)
    place_dim = int(pl.max()) + 1 if len(pl) else 1            # knn/RatingVectorsBuilder.scala:26-34
    cat_dim = int(ccat.max()) + 1 if len(ccat) else 1
    return KnnInputs(persons, prp, pci, pv, place_dim, crp, cci, cv, cat_dim,
                     pp.astype(np.int64), pl.astype(np.int64), pc.astype(np.int64))


def g2_place_visits(n_persons: int = 1_000_000, n_places: int = 100_000, seed: int = 20181231,
                    region: int = 0, mean_places: float = 7.0) -> tuple[PlaceVisitCounts, Places]:
    """Generator G2 of SURVEY.md §8(d) (rating level): per person 1+min(99, Poisson(7)) places drawn
    from a Zipf(0.8) popularity over a random permutation, visit count Geometric(0.5) >= 1."""
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, region])))
    place_ids = (MIN_PLACE_ID + np.arange(n_places)).astype(np.int64)
    cat = np.floor(rng.random(n_places) * N_CATEGORIES).astype(np.int64)
    perm = rng.permutation(n_places)
    pop = (np.arange(n_places) + 1.0) ** -0.8
    cdf = np.cumsum(pop / pop.sum())
    pid0 = min_person_id(n_places)
    d = 1 + np.minimum(99, rng.poisson(mean_places, n_persons))
    person = np.repeat(np.arange(n_persons, dtype=np.int64) + pid0, d)
    draw = perm[np.minimum(np.searchsorted(cdf, rng.random(len(person))), n_places - 1)]
    key = np.unique(person * np.int64(n_places) + draw)          # distinct places per person
    person_u, k_u = key // n_places, key % n_places
    count = rng.geometric(0.5, len(key)).astype(np.int64)
    places = Places(place_ids, np.zeros(n_places), np.zeros(n_places), cat, np.full(n_places, region, np.int64))
    return PlaceVisitCounts(person_u, place_ids[k_u], count, cat[k_u]), places


def build_stochastic_graph(v: PlaceVisitCounts, beta_person_place: float = 0.5,
                           beta_person_category: float = 0.5):
    """StochasticGraphBuilderMain.generateStochasticGraph (stochastic/StochasticGraphBuilderMain.scala:47-66).
    All visits of `v` lie inside one 7-day window, so PlaceSimilarPlace's |dt| <= 7 days
    condition (stochastic/PlaceSimilarPlace.scala:29-36) holds for every pair.
    Returns (source_id, target_id, balanced_weight) in the union order of the reference."""
    import scipy.sparse as sp

    def normalise(src, dst, cnt, top_n):
        keep = _rank_filter(src, cnt, top_n)
        src, dst, cnt = src[keep], dst[keep], cnt[keep]
        u, inv = np.unique(src, return_inverse=True)
        tot = np.bincount(inv, weights=cnt.astype(np.float64))
        return src, dst, cnt.astype(np.float64) / tot[inv]

    persons, prow = np.unique(v.person_id, return_inverse=True)
    places, pcol = np.unique(v.place_id, return_inverse=True)
    C = sp.csr_matrix((v.count.astype(np.float64), (prow, pcol)), shape=(len(persons), len(places)))
    # place -> place: count of (visit, that_visit) pairs of one person at different places
    co = (C.T @ C).tocoo()
    off = co.row != co.col
    pp_s, pp_t, pp_w = normalise(places[co.row[off]], places[co.col[off]], np.rint(co.data[off]).astype(np.int64), 50)
    # category -> place
    ckey = v.category_id * np.int64(1 << 40) + v.place_id
    cu, inv = np.unique(ckey, return_inverse=True)
    ccnt = np.bincount(inv, weights=v.count.astype(np.float64)).astype(np.int64)
    cp_s, cp_t, cp_w = normalise(cu >> 40, cu & ((1 << 40) - 1), ccnt, 100)
    # person -> place
    lp_s, lp_t, lp_w = normalise(v.person_id, v.place_id, v.count, 100)
    # person -> category
    pkey = v.person_id * np.int64(1 << 20) + v.category_id
    pu, inv = np.unique(pkey, return_inverse=True)
    pcnt = np.bincount(inv, weights=v.count.astype(np.float64)).astype(np.int64)
    lc_s, lc_t, lc_w = normalise(pu >> 20, pu & ((1 << 20) - 1), pcnt, 100)
    src = np.concatenate([pp_s, cp_s, lp_s, lc_s]).astype(np.int64)
    dst = np.concatenate([pp_t, cp_t, lp_t, lc_t]).astype(np.int64)
    w = np.concatenate([pp_w * 1.0, cp_w * 1.0, lp_w * beta_person_place, lc_w * beta_person_category])
    return src, dst, w


def random_knn_inputs(n_persons: int, n_places: int, n_cat: int, seed: int, max_places: int = 6,
                      max_count: int = 5, separate_ratings: bool = True, gaps: bool = True) -> KnnInputs:
    """Small random region-set for parity tests (ties are frequent on purpose)."""
    rng = np.random.default_rng(seed)
    pid = np.sort(rng.choice(np.arange(1000, 1000 + 4 * n_persons), n_persons, replace=False)).astype(np.int64)
    prow, pcol, pval, crow, ccol, cval = [], [], [], [], [], []
    prp, crp = [0], [0]
    for i in range(n_persons):
        npl = rng.integers(0 if gaps else 1, max_places + 1)
        cols = np.sort(rng.choice(n_places, min(npl, n_places), replace=False))
        pcol += cols.tolist()
        pval += rng.integers(1, max_count + 1, len(cols)).astype(float).tolist()
        prp.append(len(pcol))
        ncat = rng.integers(0 if gaps else 1, min(n_cat, 4) + 1)
        cc = np.sort(rng.choice(n_cat, ncat, replace=False))
        ccol += cc.tolist()
        cval += rng.integers(1, max_count + 1, len(cc)).astype(float).tolist()
        crp.append(len(ccol))
    prp, crp = np.array(prp, np.int64), np.array(crp, np.int64)
    pcol, ccol = np.array(pcol, np.int32), np.array(ccol, np.int32)
    pval, cval = np.array(pval, np.float64), np.array(cval, np.float64)
    rows = np.repeat(np.arange(n_persons), np.diff(prp))
    if separate_ratings:
        rp, rl, rv = pid[rows], pcol.astype(np.int64), pval.astype(np.int64)
        perm = rng.permutation(len(rp))
        rp, rl, rv = rp[perm], rl[perm], rv[perm]
    else:
        rp = rl = rv = None
    return KnnInputs(pid, prp, pcol, pval, n_places, crp, ccol, cval, n_cat, rp, rl, rv)


def random_stochastic_graph(n_vertices: int, out_degree: int, seed: int, id_stride: int = 3,
                            hub_fraction: float = 0.0):
    """Random row-stochastic graph with sparse ids; a hub_fraction of edges point at vertex 0 .. 2
    so that some rows of P^T are much longer than the canonical segment."""
    rng = np.random.default_rng(seed)
    ids = (np.arange(n_vertices) * id_stride + 7).astype(np.int64)
    src = np.repeat(np.arange(n_vertices), out_degree)
    dst = rng.integers(0, n_vertices, len(src))
    if hub_fraction > 0:
        hub = rng.random(len(src)) < hub_fraction
        dst = np.where(hub, rng.integers(0, 3, len(src)), dst)
    w = rng.random(len(src)) + 0.05
    tot = np.bincount(src, weights=w, minlength=n_vertices)
    w = w / tot[src]
    return ids[src], ids[dst], w


def random_layered_graph(n_categories: int, n_places: int, n_persons: int, seed: int,
                         places_per_person: int = 6, cats_per_person: int = 3, similar_per_place: int = 8,
                         hub_places: int = 2, hub_fraction: float = 0.3, duplicate_fraction: float = 0.02,
                         beta_place: float = 0.5, beta_category: float = 0.5):
    """Random graph with the edge families of stochastic/StochasticGraphBuilder.scala:8-28 and the
    id layout of the sample data (categories < places < persons, SampleGeneratorMain.scala:36-37,54):
    person->place (beta_place), person->category (beta_category), place->place, category->place, every
    family row-normalised.  A hub_fraction of the person->place and place->place edges point at the
    first hub_places places (rows of P^T far longer than the canonical segment); a
    duplicate_fraction of the person->place edges is repeated (duplicate (s, t) rows are separate
    terms, stochastic/StochasticRecommender.scala:108-114)."""
    rng = np.random.default_rng(seed)
    cat_ids = np.arange(n_categories, dtype=np.int64)
    place_ids = 2 * n_categories + np.arange(n_places, dtype=np.int64) * 2        # sparse ids
    person_ids = place_ids[-1] + 5 + np.arange(n_persons, dtype=np.int64) * 3

    def family(src_ids, n_dst, per_src, hub, beta, dst_ids):
        s = np.repeat(np.arange(len(src_ids)), per_src)
        d = rng.integers(0, n_dst, len(s))
        if hub and hub_places > 0:
            h = rng.random(len(s)) < hub_fraction
            d = np.where(h, rng.integers(0, min(hub_places, n_dst), len(s)), d)
        w = rng.random(len(s)) + 0.05
        tot = np.bincount(s, weights=w, minlength=len(src_ids))
        return src_ids[s], dst_ids[d], beta * w / tot[s]

    pp = family(person_ids, n_places, places_per_person, True, beta_place, place_ids)
    pc = family(person_ids, n_categories, cats_per_person, False, beta_category, cat_ids)
    ll = family(place_ids, n_places, similar_per_place, True, 1.0, place_ids)
    cl = family(cat_ids, n_places, min(100, n_places), False, 1.0, place_ids)
    src = np.concatenate([pp[0], pc[0], ll[0], cl[0]])
    dst = np.concatenate([pp[1], pc[1], ll[1], cl[1]])
    w = np.concatenate([pp[2], pc[2], ll[2], cl[2]])
    if duplicate_fraction > 0:
        n_dup = int(len(pp[0]) * duplicate_fraction)
        pick = rng.integers(0, len(pp[0]), n_dup)
        src = np.concatenate([src, pp[0][pick]])
        dst = np.concatenate([dst, pp[1][pick]])
        w = np.concatenate([w, pp[2][pick]])
    perm = rng.permutation(len(src))
    return src[perm], dst[perm], w[perm], person_ids, place_ids, cat_ids
