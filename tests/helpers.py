"""Shared helpers of the parity tests."""
import numpy as np


def ratings_csr(inp):
    """place_ratings COO -> rows grouped by person (stable), for the oracle."""
    if inp.rating_person is None:
        return None, None, None
    row = np.searchsorted(inp.person_id, inp.rating_person)
    order = np.argsort(row, kind="stable")
    rowptr = np.zeros(len(inp.person_id) + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    return np.cumsum(rowptr), inp.rating_place[order], inp.rating_value[order]


def oracle_knn_data(oracle, inp):
    return oracle.KnnData(inp.person_id, inp.place_rowptr, inp.place_col, inp.place_val,
                          inp.cat_rowptr, inp.cat_col, inp.cat_val,
                          *ratings_csr(inp), place_dim=max(inp.place_dim, 1))


# ---- independent pure-Python restatement (small cases only) -----------------------------
def py_cosine(ai, av, bi, bv):
    d = 0.0
    b = dict(zip(bi, bv))
    for i, v in zip(ai, av):          # ascending index of the row, as the mllib merge visits matches
        if i in b:
            d = d + v * b[i]
    la = 0.0
    for v in av:
        la = la + v * v
    lb = 0.0
    for v in bv:
        lb = lb + v * v
    return d / (np.sqrt(la) * np.sqrt(lb))


def py_knn(inp, target, pw, cw, k, place_filter, max_recs):
    """KnnRecommender.makeRecommendations + top-N, straight from the Scala source."""
    P = len(inp.person_id)
    where = {int(p): i for i, p in enumerate(inp.person_id)}
    if target not in where:
        return None
    t = where[target]

    def row(rp, ci, v, i):
        return ci[rp[i]:rp[i + 1]].tolist(), v[rp[i]:rp[i + 1]].tolist()
    tp = row(inp.place_rowptr, inp.place_col, inp.place_val, t)
    tc = row(inp.cat_rowptr, inp.cat_col, inp.cat_val, t)
    if not tp[0] or not tc[0]:
        return None
    sims = []
    for i in range(P):
        if i == t:
            continue
        ps = cs = 0.0
        keep = False
        r = row(inp.place_rowptr, inp.place_col, inp.place_val, i)
        if r[0]:
            c = py_cosine(r[0], r[1], tp[0], tp[1])
            if c > 0:
                ps, keep = c, True
        r = row(inp.cat_rowptr, inp.cat_col, inp.cat_val, i)
        if r[0]:
            c = py_cosine(r[0], r[1], tc[0], tc[1])
            if c > 0:
                cs, keep = c, True
        if keep:
            sims.append((ps * pw + cs * cw, i))
    sims.sort(key=lambda s: (-s[0], s[1]))
    nb = sorted(sims[:k], key=lambda s: s[1])
    simof = {i: s for s, i in nb}
    if inp.rating_person is None:
        rows = np.repeat(np.arange(P), np.diff(inp.place_rowptr))
        rp, rl, rv = inp.person_id[rows], inp.place_col.astype(np.int64), inp.place_val.astype(np.int64)
    else:
        rp, rl, rv = inp.rating_person, inp.rating_place, inp.rating_value
    num, den = {}, {}
    per_person = {}
    for p, l, v in zip(rp.tolist(), rl.tolist(), rv.tolist()):
        per_person.setdefault(where.get(p, -1), []).append((l, v))
    for s, i in nb:                     # ascending person index
        for l, v in per_person.get(i, []):
            num[l] = num.get(l, 0.0) + float(v) * s
            den[l] = den.get(l, 0.0) + s
    est = {l: num[l] / den[l] for l in num}
    ok = set(int(x) for x in place_filter) if place_filter is not None else None
    recs = [(e, l) for l, e in est.items() if ok is None or l in ok]
    recs.sort(key=lambda r: (-r[0], r[1]))
    return nb, est, recs[:max_recs]
