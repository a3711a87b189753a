"""Writes a --data-dir in the reference's on-disk layout from the numpy
generators of vrec.synth -- what sample_generator.sh + rating_vectors_builder.sh +
stochastic_graph_builder.sh leave behind.  Host-side tooling, not on the measured path."""
from __future__ import annotations

import argparse
import itertools
import os

import numpy as np
import pyarrow as pa

from . import data_utils as du
from . import synth


def main(argv=None) -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--person-count", type=int, default=3_000_000)
    ap.add_argument("--place-count", type=int, default=30_000)
    ap.add_argument("--uncorrelated", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--gpu-builder", action="store_true",
                    help="build the rating vectors with vrec_build_rating_vectors (needs a B200) instead of numpy")
    a = ap.parse_args(argv)
    os.makedirs(a.data_dir, exist_ok=True)
    places = synth.sample_places(a.place_count, a.seed)
    n_reg = len(synth.REGIONS)
    per_region = a.person_count // n_reg
    pid0 = synth.min_person_id(a.place_count)
    ids = np.arange(per_region * n_reg, dtype=np.int64) + pid0
    du.write_partitioned(pa.table({"id": ids, "home_region_id": (ids - pid0) // per_region}),
                         f"{a.data_dir}/persons_sample", ["home_region_id"])
    names = [f"category{c}-{i}" for c, i in zip(places.category_id, places.id)]
    du.write_partitioned(pa.table({"id": places.id, "latitude": places.latitude, "longitude": places.longitude,
                                   "category_id": places.category_id, "name": names, "description": names,
                                   "region_id": places.region_id}), f"{a.data_dir}/places_sample", ["region_id"])
    visits = [synth.sample_place_visits(places, r, per_region, a.person_count, a.place_count,
                                        correlated=not a.uncorrelated, seed=a.seed) for r in range(n_reg)]
    sets = [(r,) for r in range(n_reg)] + list(itertools.combinations(range(n_reg), 2))   # PlaceVisits.scala:63-67
    for regs in sets:
        v = synth.merge_visits([visits[r] for r in regs])
        if a.gpu_builder:
            from . import builders
            inp = builders.rating_vectors_builder(v.person_id, v.place_id, v.category_id, weight=v.count)
        else:
            inp = synth.build_rating_vectors(v)
        du.write_knn_inputs(inp, regs, a.data_dir)
        du.write_graph(*synth.build_stochastic_graph(v), regs, a.data_dir)
        print(f"wrote region set {regs}: {len(np.unique(v.person_id))} persons")


if __name__ == "__main__":
    main()
