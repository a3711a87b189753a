"""One short run of the SG batch kernel (for ncu): python tools/sg_batch_probe.py [n_q] [max_it] [tpc]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import vrec  # noqa: E402
from vrec import synth  # noqa: E402

n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
max_it = int(sys.argv[2]) if len(sys.argv) > 2 else 6
tpc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
s, t, w, persons, places, cats = synth.random_layered_graph(
    20, 10_000, 200_000, seed=4, places_per_person=2, cats_per_person=2, similar_per_place=50,
    hub_places=20, hub_fraction=0.1, duplicate_fraction=0.0)
g = vrec.StochasticGraph(s, t, w, ctx=vrec.Context(0))
g.set_option("batch", 2)
g.set_option("batch_targets_per_cta", tpc)
rec = vrec.StochasticRecommender(g, 0.0, max_it)
oi, op, cnt, its, conv, st = rec.recommend(persons[:n_q], places, 10)
print("ok", g.batch_info(0), its[:4], cnt[:4])
