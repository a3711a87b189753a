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
