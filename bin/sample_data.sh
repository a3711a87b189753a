#!/bin/sh
# Writes synthetic inputs in the reference's --data-dir layout (stands in for sample_generator.sh +
# rating_vectors_builder.sh + stochastic_graph_builder.sh, which need Spark).
set -o nounset
set -o errexit
data_dir="${DATA_DIR:-data}"
base_dir="$(cd "$(dirname -- "$0")" ; cd .. ; pwd)"
PYTHONPATH="$base_dir/locations-recommender_b200${PYTHONPATH:+:$PYTHONPATH}" \
exec python -m vrec.sample_data --data-dir "$data_dir" "$@"
