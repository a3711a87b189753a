#!/usr/bin/env python
"""Micro-benchmark of the batch KNN query on G2 data (tuning + the short command profiled under ncu)."""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import vrec  # noqa: E402
from vrec import _lib as L, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--persons", type=int, default=1_000_000)
ap.add_argument("--places", type=int, default=100_000)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--k", type=int, default=50)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--splits", type=int, default=0)
ap.add_argument("--skip-postings", type=int, default=0)
ap.add_argument("--tc-seed", type=int, default=0)
ap.add_argument("--compact", type=int, default=1, help="0 = exact evaluations on the full-size records (A/B)")
ap.add_argument("--post-first", type=int, default=0, help="1 = postings kernel before the dense filter (its K-th best seeds the thresholds)")
a = ap.parse_args()
v, places = synth.g2_place_visits(a.persons, a.places)
inp = synth.build_rating_vectors(v)
ctx = vrec.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
rs.set_option("knn_kernel", a.kernel)
rs.set_option("splits", a.splits)
rs.set_option("debug_skip_postings", a.skip_postings)
rs.set_option("tc_seed", a.tc_seed)
rs.set_option("compact_records", a.compact)
rs.set_option("post_first", a.post_first)
lib = ctx.lib
flt = np.ascontiguousarray(places.id)
lib.vrec_knn_set_filter(rs._h, flt.ctypes.data_as(L.i64p), len(flt))
B, m = a.batch, 10
P = len(inp.person_id)
d_place = torch.empty((B, m), dtype=torch.int64, device="cuda")
d_rating = torch.empty((B, m), dtype=torch.float64, device="cuda")
d_count = torch.empty(B, dtype=torch.int32, device="cuda")
d_status = torch.empty(B, dtype=torch.int32, device="cuda")
stats = (C.c_uint64 * 4)()
cyc = (C.c_uint64 * 12)()
for s in range(a.warmup + a.steps):
    t = torch.from_numpy(inp.person_id[(np.arange(B) + s * B) % P]).cuda()
    lib.vrec_knn_debug_stats(rs._h, stats)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rc = lib.vrec_knn_query_device(rs._h, t.data_ptr(), B, 0.5, 0.5, a.k, m, d_place.data_ptr(),
                                   d_rating.data_ptr(), d_count.data_ptr(), d_status.data_ptr())
    assert rc == 0, L.last_error()
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1)
    lib.vrec_knn_debug_stats(rs._h, stats)
    st = [int(x) for x in stats]
    lib.vrec_knn_debug_tc_cycles(rs._h, cyc)
    cy = [int(x) for x in cyc]
    if cy[11] and a.kernel == 3:
        print(f"   postings warp-iteration (block0 warp0): candidate id load {cy[9] / cy[11]:.0f} cycles, exact eval + insert "
              f"{cy[10] / cy[11]:.0f} cycles, iterations {cy[11]}")
    pb = (C.c_uint64 * 8)()
    lib.vrec_knn_debug_probe(rs._h, pb)
    if pb[3]:
        n = pb[3]
        print(f"   survivor eval (block0 thread0), cycles per eval over {n}: meta {pb[0] / n:.0f}, headers {pb[1] / n:.0f}, "
              f"place section {pb[6] / n:.0f}, place matching {pb[2] / n:.0f}, category {pb[7] / n:.0f}; heap insert "
              f"{pb[4] / max(1, pb[5]):.0f} over {pb[5]} inserts")
    nb = min(1024, (B + 127) // 128)
    bc = (C.c_uint64 * (2 * nb))()
    lib.vrec_knn_debug_tc_block_cycles(rs._h, bc, nb)
    bcs = np.array([int(x) for x in bc], dtype=np.float64).reshape(2, nb) / 1e6
    if cy[5] and a.kernel in (0, 4):
        n = cy[5]
        print(f"   ws block0 cycles/tile: producer [load issue+stage wait {cy[3] / n:.0f}, B-tile wait {cy[0] / n:.0f}, "
              f"accumulator wait {cy[7] / n:.0f}, mma issue {cy[8] / n:.0f}] consumer warp0 [vote+drain {cy[6] / n:.0f} (vote barrier {cy[9] / n:.0f}, own drain work {cy[10] / n:.0f}, drains {cy[11]}), "
              f"accumulator wait {cy[1] / n:.0f}, epilogue {cy[2] / n:.0f}] tiles {n}")
    elif cy[5]:
        print(f"   per-block Mcycles: dense mean {bcs[0].mean():.1f} max {bcs[0].max():.1f} | postings mean "
              f"{bcs[1].mean():.1f} max {bcs[1].max():.1f} | total max {(bcs[0] + bcs[1]).max():.1f}")
    if cy[5] and a.kernel == 3:
        print(f"   tc block0 cycles/tile: load-wait {cy[0] / cy[5]:.0f}, mma {cy[1] / cy[5]:.0f}, epilogue {cy[2] / cy[5]:.0f}, "
              f"mma issue {cy[8] / cy[5]:.0f}, load issue {cy[3] / cy[5]:.0f} [cp.async wait {cy[6] / cy[5]:.0f}, barrier {cy[7] / cy[5]:.0f}]; postings pass total {cy[4] / 1e6:.2f} Mcycles; tiles {cy[5]}")
    if cy[6] and a.kernel in (0, 4):
        print(f"   ws block0 first evaluator warp: {cy[6]} batches, {cy[4] / cy[6]:.0f} cycles evaluating and {cy[3] / cy[6]:.0f} waiting per batch")
    dms = C.c_double(0.0)
    lib.vrec_knn_last_dense_ms(rs._h, C.byref(dms))
    print(f"   dense filter kernel {dms.value:.2f} ms")
    print(f"kernel={a.kernel} step {s}: {ms:.1f} ms  {B / ms * 1e3:,.0f} persons/s   per target: postings evals {st[0] / B:.0f}, "
          f"filter survivors {st[1] / B:.0f}, heap inserts {st[2] / B:.0f}, queue overflow {st[3] / B:.0f}", flush=True)
