#!/bin/sh
# Drop-in for the reference's bin/stochastic_recommender.sh: same flags and values
# (bin/stochastic_recommender.sh:32-35), same REPL on stdin, engine = libvrec.so on the local B200.
set -o nounset
set -o errexit

data_dir="${DATA_DIR:-data}"
base_dir="$(cd "$(dirname -- "$0")" ; cd .. ; pwd)"
PYTHONPATH="$base_dir/locations-recommender_b200${PYTHONPATH:+:$PYTHONPATH}" \
exec python -m vrec.sg_main \
--data-dir "$data_dir" \
--epsilon 0.01 \
--max-iterations 20 \
--max-recommendations 10
