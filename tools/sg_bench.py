#!/usr/bin/env python
"""Micro-benchmark of the SG power iteration on the device-generated graph (used for tuning
and as the short command profiled under ncu)."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "locations-recommender_b200"))
import torch  # noqa: E402

import vrec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--deg", type=int, default=100)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--minb", type=int, default=0)
ap.add_argument("--rows-kernel", type=int, default=0, help="1 = round-1 half-warp-per-row kernel (A/B)")
a = ap.parse_args()
ctx = vrec.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
g = vrec.StochasticGraph.generate(a.n, a.deg, seed=5, ctx=ctx)
if a.minb:
    g.set_option("flat_variant", a.minb)
if a.rows_kernel:
    g.set_option("rows_kernel", 1)
ctx.synchronize()
bytes_it = 12 * g.nnz + 20 * g.N
for _ in range(a.warmup):
    g.iterate_device(a.iters)
ctx.synchronize()
for s in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.iterate_device(a.iters)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"N={g.N} nnz={g.nnz} {ms / a.iters * 1e3:.0f} us/iter  {bytes_it * a.iters / ms / 1e6:.0f} GB/s "
          f"({bytes_it * a.iters / ms / 1e6 / 6449.1:.3f} of measured HBM peak)", flush=True)
