#!/usr/bin/env bash
# Builds locations-recommender_b200/libvrec_<tag>.so with extra -D flags for vrec_knn.cu (A/B measurements):
#   tools/build_variant.sh v0 -DVREC_POST_VARIANT=0
#   VREC_LIB_PATH=locations-recommender_b200/libvrec_v0.so python tools/knn_bench.py ...
set -euo pipefail
tag="$1"; shift
cd "$(dirname "$0")/../locations-recommender_b200/csrc"
make -s -j4
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC,-Wall,-Wno-unused-function -I../../include "$@" -Xptxas -v -dc -o /tmp/vrec_knn_$tag.o vrec_knn.cu \
    2> /tmp/vrec_knn_$tag.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../libvrec_$tag.so vrec_api.o vrec_sg.o \
    vrec_sg_batch.o vrec_build.o /tmp/vrec_knn_$tag.o vrec_comm.o vrec_tc_selftest.o -lcudart_static -ldl -lrt -lpthread
grep -A3 "knn_postings_kernel" /tmp/vrec_knn_$tag.ptxas.log | grep -E "registers|spill" | tr "\n" " "; echo
