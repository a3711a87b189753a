"""The drop-in launchers end to end on a small --data-dir: same stdin protocol and messages as the
reference's REPL mains; recommended ids checked against the oracle."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))

from tests.helpers import oracle_knn_data  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("data"))
    env = dict(os.environ, DATA_DIR=d)
    subprocess.check_call([os.path.join(ROOT, "bin", "sample_data.sh"), "--person-count", "9000",
                           "--place-count", "300", "--uncorrelated"], env=env)
    return d


def _run(script, data_dir, stdin):
    env = dict(os.environ, DATA_DIR=data_dir)
    p = subprocess.run([os.path.join(ROOT, "bin", script)], input=stdin, capture_output=True, text=True, env=env,
                       timeout=300)
    return p.stdout, p.stderr


def _table_ids(out):
    return [int(m.group(1)) for m in re.finditer(r"^\|(\d+)\s*\|", out, flags=re.M)]


def test_knn_launcher(data_dir, oracle):
    from vrec import data_utils as du
    args = du.load_knn_inputs((0,), data_dir, verbose=False)
    person = int(args[0][0])
    out, err = _run("knn_recommender.sh", data_dir, f"{person}\n{person} 1\nabc\n1\n")
    assert "Loading place rating vectors from" in out and f"Person {person} might want to visit in region 0:" in out
    assert "Target region ID is not provided - falling back to the person's home region" in out
    assert "Done in " in out
    assert "IllegalArgumentException: Failed to parse input: abc" in err
    assert "NoSuchElementException: Person not found: 1" in err

    class Inp:
        pass
    inp = Inp()
    (inp.person_id, inp.place_rowptr, inp.place_col, inp.place_val, inp.place_dim, inp.cat_rowptr, inp.cat_col,
     inp.cat_val, inp.cat_dim, inp.rating_person, inp.rating_place, inp.rating_value) = args
    d = oracle_knn_data(oracle, inp)
    places = du.load_places(data_dir).to_pandas()
    flt = places[places["region_id"].astype(int) == 0]["id"].to_numpy()
    rc, opl, ort, ocnt, ost = oracle.knn_query_batch(d, [person], 0.5, 0.5, 2_000_000, flt, 10)
    first = out.split(f"Person {person} might want to visit in region 0:")[1].split("Done in")[0]
    assert _table_ids(first) == opl[0, :ocnt[0]].tolist()


def test_sg_launcher(data_dir, oracle):
    from vrec import data_utils as du
    s, t, w = du.load_graph((0,), data_dir, verbose=False)
    persons, home = du.load_persons(data_dir)
    person = int(np.intersect1d(persons[home == 0], s)[0])
    out, err = _run("stochastic_recommender.sh", data_dir, f"{person}\n424242\n")
    assert "Loading stochastic graph of visited places from" in out
    assert re.search(r"(Converged in \d+ iterations|Number of iterations \d+ reached the maximum 20)", out)
    assert "NoSuchElementException: Person not found: 424242" in err
    g = oracle.SgGraph(s, t, w)
    places = du.load_places(data_dir).to_pandas()
    flt = places[places["region_id"].astype(int) == 0]["id"].to_numpy()
    rc, ids, pr, it, conv = g.query(person, 0.01, 20, flt, 10)
    assert _table_ids(out.split("might want to visit")[1]) == ids.tolist()
