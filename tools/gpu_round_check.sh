#!/usr/bin/env bash
# One GPU-box call that re-validates everything measured in a round (≈ 5 min of box time):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/gpu_round_check.sh r2a'
# Outputs under gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
tag="${1:-check}"
out=gpurun_out
mkdir -p "$out"
echo "== GPU tests";            timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke";                timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full-size KNN parity (all variants, all K regimes)"
timeout 200 python tools/knn_fullsize_parity.py > "$out/${tag}_knn_fullsize_parity.log" 2>&1
echo "rc=$?"; grep -v '^{' "$out/${tag}_knn_fullsize_parity.log" | cut -c1-200 | tail -30
echo "== bench";                timeout 300 python bench.py > "$out/${tag}_bench.json" 2> "$out/${tag}_bench.err"
echo "rc=$?"; grep -i "parity\|knn device\|knn e2e\|sg device" "$out/${tag}_bench.err" | cut -c1-220
echo "== ncu launch list of one KNN step (after the bench above exited)"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$out/${tag}_knn_launches.csv" python tools/knn_bench.py --steps 1 --warmup 1 --batch 18944 \
    > "$out/${tag}_ncu.log" 2>&1
echo "rc=$?"
