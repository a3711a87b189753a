#!/bin/sh
# Drop-in for the reference's bin/knn_recommender.sh: same flags and values
# (bin/knn_recommender.sh:32-36), same REPL on stdin, engine = libvrec.so on the local B200.
set -o nounset
set -o errexit

data_dir="${DATA_DIR:-data}"
base_dir="$(cd "$(dirname -- "$0")" ; cd .. ; pwd)"
PYTHONPATH="$base_dir/locations-recommender_b200${PYTHONPATH:+:$PYTHONPATH}" \
exec python -m vrec.knn_main \
--data-dir "$data_dir" \
--place-weight 0.5 \
--category-weight 0.5 \
--k-nearest 2000000 \
--max-recommendations 10
