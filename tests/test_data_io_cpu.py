"""Host shim: file naming (DataUtils.scala:54-60), VectorUDT Parquet round trip, REPL input parsing
(RecommenderMainCommon.scala:16-25)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))

from vrec import data_utils as du  # noqa: E402
from vrec import synth  # noqa: E402
from vrec.main_common import calc_recommender_target, parse_input  # noqa: E402


def test_file_names():
    assert du.graph_file_name([2, 0], "data") == "data/stochastic_graph_region0_region2"
    assert du.graph_file_name([1, 1], "d") == "d/stochastic_graph_region1"
    assert du.place_rating_vectors_file_name([0], "d") == "d/place_rating_vectors_region0"
    assert du.category_rating_vectors_file_name([2, 1], "d") == "d/category_rating_vectors_region1_region2"
    assert du.place_ratings_file_name([0, 1], "d") == "d/place_ratings_region0_region1"


def test_parse_input():
    assert parse_input("123") == (123, None)
    assert parse_input("123 2") == (123, 2)
    assert parse_input("123    45") == (123, 45)
    for bad in ["", "abc", "12 x", "-3"]:
        with pytest.raises(ValueError, match="Failed to parse input"):
            parse_input(bad)
    persons = (np.array([10, 11, 12]), np.array([0, 1, 2]))
    t = calc_recommender_target(persons, 11, None)
    assert (t.personId, t.homeRegionId, t.targetRegionId) == (11, 1, 1)
    t = calc_recommender_target(persons, 12, 0)
    assert (t.homeRegionId, t.targetRegionId) == (2, 0)
    with pytest.raises(LookupError, match="Person not found: 99"):
        calc_recommender_target(persons, 99, None)


def test_knn_parquet_round_trip(tmp_path):
    inp = synth.random_knn_inputs(80, 30, 7, seed=5)
    du.write_knn_inputs(inp, (1, 0), str(tmp_path))
    assert os.path.isdir(tmp_path / "place_rating_vectors_region0_region1")
    got = du.load_knn_inputs((0, 1), str(tmp_path), verbose=False)
    # persons with no row in either table are not written; everything else is identical
    keep = (np.diff(inp.place_rowptr) > 0) | (np.diff(inp.cat_rowptr) > 0)
    assert np.array_equal(got[0], inp.person_id[keep])
    assert np.array_equal(got[2], inp.place_col) and np.array_equal(got[3], inp.place_val)
    assert np.array_equal(got[6], inp.cat_col) and np.array_equal(got[7], inp.cat_val)
    assert np.array_equal(np.diff(got[1]), np.diff(inp.place_rowptr)[keep])
    assert (got[4], got[8]) == (inp.place_dim, inp.cat_dim)
    assert sorted(zip(got[9], got[10], got[11])) == sorted(zip(inp.rating_person, inp.rating_place, inp.rating_value))


def test_graph_parquet_round_trip(tmp_path):
    s, t, w = synth.random_stochastic_graph(50, 3, seed=2)
    du.write_graph(s, t, w, (2,), str(tmp_path))
    s2, t2, w2 = du.load_graph((2,), str(tmp_path), verbose=False)
    assert np.array_equal(s, s2) and np.array_equal(t, t2) and np.array_equal(w, w2)


def test_knn_parquet_unordered_rows_and_dense_vectors(tmp_path):
    """Spark writes the rows of a table in no particular order, and in several part files: the loader must put
    every person's vector in its place (vectorised `align`); a dense VectorUDT row is rejected, not misread."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    inp = synth.random_knn_inputs(200, 40, 6, seed=9, gaps=False)
    rng = np.random.default_rng(1)

    def shuffled(rowptr, col, val, dim, name):
        perm = rng.permutation(len(inp.person_id))
        lens = np.diff(rowptr)[perm]
        take = np.concatenate([np.arange(rowptr[i], rowptr[i + 1]) for i in perm])
        t = du.vectors_to_table(inp.person_id[perm], np.concatenate([[0], np.cumsum(lens)]), col[take], val[take], dim)
        d = tmp_path / name
        d.mkdir()
        half = len(perm) // 2
        pq.write_table(t.slice(0, half), str(d / "part-00000.parquet"))
        pq.write_table(t.slice(half), str(d / "part-00001.parquet"))

    shuffled(inp.place_rowptr, inp.place_col, inp.place_val, inp.place_dim, "place_rating_vectors_region3")
    shuffled(inp.cat_rowptr, inp.cat_col, inp.cat_val, inp.cat_dim, "category_rating_vectors_region3")
    d = tmp_path / "place_ratings_region3"
    d.mkdir()
    pq.write_table(pa.table({"person_id": pa.array(inp.rating_person), "place_id": pa.array(inp.rating_place),
                             "rating": pa.array(inp.rating_value)}), str(d / "part-00000.parquet"))
    got = du.load_knn_inputs((3,), str(tmp_path), verbose=False)
    assert np.array_equal(got[0], inp.person_id)
    assert np.array_equal(got[1], inp.place_rowptr) and np.array_equal(got[2], inp.place_col)
    assert np.array_equal(got[3], inp.place_val)
    assert np.array_equal(got[5], inp.cat_rowptr) and np.array_equal(got[6], inp.cat_col)
    assert np.array_equal(got[7], inp.cat_val)
    # a dense vector (type 1) in the column
    t = du.vectors_to_table(inp.person_id[:3], np.array([0, 1, 2, 3]), np.array([0, 1, 2], np.int32),
                            np.array([1.0, 2.0, 3.0]), 5)
    vec = t["rating_vector"].combine_chunks()
    dense = pa.StructArray.from_arrays([pa.array(np.array([0, 1, 0], np.int8)), vec.field("size"),
                                        vec.field("indices"), vec.field("values")], fields=list(du.VECTOR_UDT))
    with pytest.raises(ValueError, match="dense vectors"):
        du.table_to_vectors(pa.table({"person_id": t["person_id"], "rating_vector": dense}))


def test_places_without_pandas(tmp_path):
    import pyarrow as pa
    t = pa.table({"id": pa.array([45, 41, 43], pa.int64()), "latitude": [1.0, 2.0, 3.0], "longitude": [4.0, 5.0, 6.0],
                  "category_id": pa.array([7, 8, 9], pa.int64()), "name": ["a-45", "b-41", "c-43"],
                  "description": ["x", "y", "z"], "region_id": pa.array([0, 1, 0], pa.int32())})
    p = du.Places(t)
    assert p.of_region(0).tolist() == [45, 43] and p.of_region(2).tolist() == []
    assert p.row(41) == [41, 2.0, 5.0, 8, "b-41", "y", 1]
