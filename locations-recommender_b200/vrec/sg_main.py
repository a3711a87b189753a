"""Drop-in for com.github.tashoyan.recommender.stochastic.StochasticRecommenderMain
(stochastic/StochasticRecommenderMain.scala): same flags
(stochastic/StochasticRecommenderArgParser.scala:11-55), REPL protocol and messages."""
from __future__ import annotations

import argparse
import sys
import time

from . import data_utils as du
from .engine import Context, StochasticGraph, StochasticRecommender
from .main_common import calc_recommender_target, parse_input, repl, show


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="recommender", description="Recommender")
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--epsilon", type=float, default=0.05)
    ap.add_argument("--max-iterations", type=int, default=20)
    ap.add_argument("--max-recommendations", type=int, default=10)
    a = ap.parse_args(argv)
    if not a.data_dir:
        ap.error("Data directory must be non-empty path")
    if a.epsilon < 0:
        ap.error("Epsilon must be non-negative")
    if a.max_iterations < 0:
        ap.error("Maximum iterations number must be non-negative")
    if a.max_recommendations < 0:
        ap.error("Maximum recommendations number must be non-negative")
    return a


def main(argv=None) -> None:
    cfg = parse_args(sys.argv[1:] if argv is None else argv)
    print(f"Actual configuration: StochasticRecommenderConfig({cfg.data_dir},{cfg.epsilon},{cfg.max_iterations},"
          f"{cfg.max_recommendations})")
    ctx = Context()
    print(f"Loading persons from {cfg.data_dir}/persons_sample")
    persons = du.load_persons(cfg.data_dir)
    print(f"Loading places from {cfg.data_dir}/places_sample")
    places = du.Places(du.load_places(cfg.data_dir))
    cache = {}

    def query(line: str) -> None:
        person, region = parse_input(line)
        tgt = calc_recommender_target(persons, person, region)
        key = tuple(sorted({tgt.homeRegionId, tgt.targetRegionId}))
        if key not in cache:
            cache[key] = StochasticGraph(*du.load_graph(key, cfg.data_dir), ctx=ctx)
        rec = StochasticRecommender(cache[key], cfg.epsilon, cfg.max_iterations, verbose=True)
        region_places = places.of_region(tgt.targetRegionId)
        print(f"Person {tgt.personId} might want to visit in region {tgt.targetRegionId}:")
        t0 = time.time()
        ids, pr, cnt, its, conv, st = rec.recommend([tgt.personId], region_places,
                                                    cfg.max_recommendations)
        if st[0] != 0:
            raise ValueError(f"No such vertex in the graph: {tgt.personId}")
        cols = ["id", "latitude", "longitude", "category_id", "name", "description", "region_id", "probability"]
        rows = [places.row(int(p)) + [r] for p, r in zip(ids[0, :cnt[0]], pr[0, :cnt[0]])]
        show(rows, cols)
        print(f"Done in {int((time.time() - t0) * 1000)} milliseconds")

    repl(query)


if __name__ == "__main__":
    main()
