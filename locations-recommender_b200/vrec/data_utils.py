"""File naming and Parquet I/O of the recommenders' inputs (host shim, pyarrow).

Mirrors DataUtils.scala:9-60 (file names: "<dir>/<prefix>_region<a>[_region<b>]" over sorted distinct
region ids) and the on-disk schemas of SURVEY.md Appendix A.  Spark's VectorUDT column is the plain
struct<type: int8, size: int32, indices: list<int32>, values: list<double>>.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import pyarrow as pa
import pyarrow.dataset as pads
import pyarrow.parquet as pq


def generate_file_name(region_ids: Sequence[int], dir_path: str, prefix: str) -> str:
    # DataUtils.generateFileName (DataUtils.scala:54-60)
    regs = sorted(set(int(r) for r in region_ids))
    return f"{dir_path}/{prefix}_" + "_".join(f"region{r}" for r in regs)


def graph_file_name(region_ids, d):                     # DataUtils.scala:33-35
    return generate_file_name(region_ids, d, "stochastic_graph")


def place_rating_vectors_file_name(region_ids, d):      # DataUtils.scala:42-44
    return generate_file_name(region_ids, d, "place_rating_vectors")


def category_rating_vectors_file_name(region_ids, d):   # DataUtils.scala:46-48
    return generate_file_name(region_ids, d, "category_rating_vectors")


def place_ratings_file_name(region_ids, d):             # DataUtils.scala:50-52
    return generate_file_name(region_ids, d, "place_ratings")


VECTOR_UDT = pa.struct([("type", pa.int8()), ("size", pa.int32()),
                        ("indices", pa.list_(pa.int32())), ("values", pa.list_(pa.float64()))])


def _read(path: str) -> pa.Table:
    return pads.dataset(path, format="parquet", partitioning="hive").to_table()


def load_persons(data_dir: str):
    """persons_sample: id + home_region_id (directory partition) -- DataUtils.loadPersons, :9-15."""
    t = _read(f"{data_dir}/persons_sample")
    return (t["id"].to_numpy().astype(np.int64), t["home_region_id"].to_numpy().astype(np.int64))


def load_places(data_dir: str) -> pa.Table:
    """places_sample with region_id cast to long -- DataUtils.loadPlaces, :17-23."""
    return _read(f"{data_dir}/places_sample")


class Places:
    """places_sample as plain columns (Arrow buffers -> numpy, no pandas hop): the region filter and the
    final join of printRecommendations (knn/KnnRecommenderMain.scala:96-102)."""

    COLUMNS = ("id", "latitude", "longitude", "category_id", "name", "description", "region_id")

    def __init__(self, table: pa.Table):
        self.id = table["id"].to_numpy().astype(np.int64)
        self.region_id = table["region_id"].to_numpy().astype(np.int64)
        self._cols = {c: table[c] for c in self.COLUMNS if c not in ("id", "region_id")}
        self._order = np.argsort(self.id, kind="stable")
        self._sorted = self.id[self._order]

    def of_region(self, region: int) -> np.ndarray:
        return np.ascontiguousarray(self.id[self.region_id == int(region)])

    def row(self, place_id: int) -> list:
        k = int(np.searchsorted(self._sorted, place_id))
        i = int(self._order[k])
        assert self._sorted[k] == place_id
        return [int(place_id)] + [self._cols[c][i].as_py() for c in self.COLUMNS[1:-1]] + [int(self.region_id[i])]


def write_partitioned(table: pa.Table, path: str, partition_cols) -> None:
    pq.write_to_dataset(table, path, partition_cols=list(partition_cols))


# ---- rating vectors (VectorUDT) <-> CSR
def vectors_to_table(person_id, rowptr, col, val, size: int) -> pa.Table:
    n = len(person_id)
    offs = pa.array(np.asarray(rowptr, dtype=np.int32))
    indices = pa.ListArray.from_arrays(offs, pa.array(np.asarray(col, dtype=np.int32)))
    values = pa.ListArray.from_arrays(offs, pa.array(np.asarray(val, dtype=np.float64)))
    vec = pa.StructArray.from_arrays(
        [pa.array(np.zeros(n, dtype=np.int8)), pa.array(np.full(n, size, dtype=np.int32)), indices, values],
        fields=list(VECTOR_UDT))
    return pa.table({"person_id": pa.array(np.asarray(person_id, dtype=np.int64)), "rating_vector": vec})


def table_to_vectors(t: pa.Table):
    """-> person_id, rowptr, col, val, size.  The Arrow list offsets / values of the VectorUDT column ARE the CSR.
    Dense vectors (type == 1: `values` only, no `indices`) are never written by RatingVectorsBuilder
    (knn/RatingVectorsBuilder.scala:74-83 builds SparseVectors) and are rejected rather than misread."""
    pid = t["person_id"].to_numpy().astype(np.int64)
    vec = t["rating_vector"].combine_chunks()
    kinds = vec.field("type").to_numpy(zero_copy_only=False)
    if len(kinds) and np.any(kinds != 0):
        raise ValueError(f"rating_vector holds {int(np.count_nonzero(kinds != 0))} dense vectors (VectorUDT type 1); "
                         "the recommender inputs are sparse vectors")
    idx, vals = vec.field("indices"), vec.field("values")
    rowptr = idx.offsets.to_numpy().astype(np.int64)
    rowptr = rowptr - rowptr[0]
    col = idx.values.to_numpy().astype(np.int32)[int(idx.offsets[0].as_py()):][:rowptr[-1]]
    val = vals.values.to_numpy().astype(np.float64)[int(vals.offsets[0].as_py()):][:rowptr[-1]]
    size = int(vec.field("size")[0].as_py()) if len(pid) else 1
    return pid, rowptr, col, val, size


def write_knn_inputs(inp, region_ids, data_dir: str) -> None:
    """The three directories RatingVectorsBuilderMain writes (knn/RatingVectorsBuilderMain.scala:67-73)."""
    os.makedirs(data_dir, exist_ok=True)
    has_p = np.diff(inp.place_rowptr) > 0       # a person is a row of a table only if it has entries
    has_c = np.diff(inp.cat_rowptr) > 0

    def sub(rowptr, col, val, keep):
        lens = np.diff(rowptr)[keep]
        rp = np.concatenate([[0], np.cumsum(lens)])
        sel = np.repeat(keep, np.diff(rowptr))
        return rp, col[sel], val[sel]

    rp, c, v = sub(inp.place_rowptr, inp.place_col, inp.place_val, has_p)
    pq.write_table(vectors_to_table(inp.person_id[has_p], rp, c, v, inp.place_dim),
                   _single(place_rating_vectors_file_name(region_ids, data_dir)))
    rp, c, v = sub(inp.cat_rowptr, inp.cat_col, inp.cat_val, has_c)
    pq.write_table(vectors_to_table(inp.person_id[has_c], rp, c, v, inp.cat_dim),
                   _single(category_rating_vectors_file_name(region_ids, data_dir)))
    pq.write_table(pa.table({"person_id": pa.array(inp.rating_person), "place_id": pa.array(inp.rating_place),
                             "rating": pa.array(inp.rating_value)}),
                   _single(place_ratings_file_name(region_ids, data_dir)))


def _single(dir_path: str) -> str:
    os.makedirs(dir_path, exist_ok=True)
    return os.path.join(dir_path, "part-00000.parquet")


def load_knn_inputs(region_ids, data_dir: str, verbose: bool = True):
    """The three reads of KnnRecommenderMain (knn/KnnRecommenderMain.scala:69-88) -> vrec_knn_load arguments."""
    f1 = place_rating_vectors_file_name(region_ids, data_dir)
    f2 = category_rating_vectors_file_name(region_ids, data_dir)
    f3 = place_ratings_file_name(region_ids, data_dir)
    if verbose:
        print(f"Loading place rating vectors from {f1}")
        print(f"Loading category rating vectors from {f2}")
        print(f"Loading place ratings from {f3}")
    ppid, prp, pc, pv, pdim = table_to_vectors(_read(f1))
    cpid, crp, cc, cv, cdim = table_to_vectors(_read(f2))
    r = _read(f3)
    persons = np.union1d(ppid, cpid)

    def align(pid, rowptr, col, val):
        # rows of one table -> rows of the union of persons (persons missing from a table get an empty row);
        # vectorised: no per-person Python work at 10^6 persons
        order = np.argsort(pid, kind="stable")
        src_len = np.diff(rowptr)[order]
        lens = np.zeros(len(persons), dtype=np.int64)
        lens[np.searchsorted(persons, pid[order])] = src_len
        out_rp = np.concatenate([[0], np.cumsum(lens)])
        dst_start = np.concatenate([[0], np.cumsum(src_len)[:-1]]) if len(order) else np.zeros(0, dtype=np.int64)
        take = np.repeat(rowptr[:-1][order] - dst_start, src_len) + np.arange(int(src_len.sum()), dtype=np.int64)
        return out_rp, col[take], val[take]

    prp2, pc2, pv2 = align(ppid, prp, pc, pv)
    crp2, cc2, cv2 = align(cpid, crp, cc, cv)
    return (persons, prp2, pc2, pv2, pdim, crp2, cc2, cv2, cdim,
            r["person_id"].to_numpy().astype(np.int64), r["place_id"].to_numpy().astype(np.int64),
            r["rating"].to_numpy().astype(np.int64))


def write_graph(src, dst, w, region_ids, data_dir: str) -> None:
    """stochastic_graph_region<..> (stochastic/StochasticGraphBuilderMain.scala:68-73)."""
    pq.write_table(pa.table({"source_id": pa.array(np.asarray(src, np.int64)),
                             "target_id": pa.array(np.asarray(dst, np.int64)),
                             "balanced_weight": pa.array(np.asarray(w, np.float64))}),
                   _single(graph_file_name(region_ids, data_dir)))


def load_graph(region_ids, data_dir: str, verbose: bool = True):
    f = graph_file_name(region_ids, data_dir)
    if verbose:     # stochastic/StochasticRecommenderMain.scala:85
        print(f"Loading stochastic graph of visited places from {f}")
    t = _read(f)
    return (t["source_id"].to_numpy().astype(np.int64), t["target_id"].to_numpy().astype(np.int64),
            t["balanced_weight"].to_numpy().astype(np.float64))
