"""Drop-in for com.github.tashoyan.recommender.knn.KnnRecommenderMain (knn/KnnRecommenderMain.scala):
same flags (knn/KnnRecommenderArgParser.scala:11-71), same REPL protocol, same messages."""
from __future__ import annotations

import argparse
import sys
import time

from . import data_utils as du
from .engine import Context, KnnRecommender, KnnRegionSet
from .main_common import calc_recommender_target, parse_input, repl, show


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="recommender", description="Recommender")
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--place-weight", type=float, required=True)
    ap.add_argument("--category-weight", type=float, required=True)
    ap.add_argument("--k-nearest", type=int, required=True)
    ap.add_argument("--max-recommendations", type=int, default=10)
    a = ap.parse_args(argv)
    if not a.data_dir:
        ap.error("Data directory must be non-empty path")
    for w in (a.place_weight, a.category_weight):
        if w <= 0 or w >= 1:
            ap.error("Weight must be in the interval (0; 1)")
    if a.k_nearest <= 0:
        ap.error("K nearest must be positive")
    if a.max_recommendations < 0:
        ap.error("Maximum recommendations number must be non-negative")
    if a.place_weight + a.category_weight != 1.0:
        ap.error(f"Sum of weights must be 1.0: for place-based similarity: {a.place_weight}, "
                 f"for category-based similarity: {a.category_weight}")
    return a


def main(argv=None) -> None:
    cfg = parse_args(sys.argv[1:] if argv is None else argv)
    print(f"Actual configuration: KnnRecommenderConfig({cfg.data_dir},{cfg.place_weight},{cfg.category_weight},"
          f"{cfg.k_nearest},{cfg.max_recommendations})")
    ctx = Context()
    print(f"Loading persons from {cfg.data_dir}/persons_sample")
    persons = du.load_persons(cfg.data_dir)
    print(f"Loading places from {cfg.data_dir}/places_sample")
    places = du.Places(du.load_places(cfg.data_dir))
    cache = {}          # region-sets stay resident on the device between queries

    def query(line: str) -> None:
        person, region = parse_input(line)
        tgt = calc_recommender_target(persons, person, region)
        key = tuple(sorted({tgt.homeRegionId, tgt.targetRegionId}))
        if key not in cache:
            cache[key] = KnnRegionSet(*du.load_knn_inputs(key, cfg.data_dir), ctx=ctx)
        rec = KnnRecommender(cache[key], cfg.place_weight, cfg.category_weight, cfg.k_nearest)
        region_places = places.of_region(tgt.targetRegionId)
        print(f"Person {tgt.personId} might want to visit in region {tgt.targetRegionId}:")
        t0 = time.time()
        pl, rt, cnt, st = rec.recommend([tgt.personId], region_places, cfg.max_recommendations)
        if st[0] != 0:
            raise ValueError(f"No such person: {tgt.personId}")
        cols = ["id", "latitude", "longitude", "category_id", "name", "description", "region_id", "place_id",
                "estimated_rating"]
        rows = [places.row(int(p)) + [int(p), r] for p, r in zip(pl[0, :cnt[0]], rt[0, :cnt[0]])]
        show(rows, cols)
        print(f"Done in {int((time.time() - t0) * 1000)} milliseconds")

    repl(query)


if __name__ == "__main__":
    main()
