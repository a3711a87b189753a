#!/usr/bin/env python
"""bench.py -- throughput of the two hot paths on B200, one JSON line on stdout.

Headline (`metric`, `value`): KNN target persons/second on BASELINE.json config 3 (batch KNN,
1 M persons / 100 K places, K = 50, place+category weights 0.5/0.5, top-10 places), data resident
in HBM.  `e2e` is the same through the host-buffer C-ABI call (vrec_knn_query).  The `sg` object
carries the second half of BASELINE.json's metric: SG power-iteration HBM GB/s on the oversized
synthetic graph (config 5 shape at the size that fits one GPU), with its own roofline / e2e /
cpu_baseline.  `--impl reference` times the CPU oracle port (the reference is Scala/Spark and
cannot run here: no JVM) on bounded samples of the same workloads.

A "step" = `--knn-calls` batches of `--knn-batch` targets per GPU (KNN; 8 x 18 944 by default, so that the 20
timed steps the driver asks for last ~2.5 s) / one 20-iteration power iteration (SG).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "locations-recommender_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries (NCCL's version banner, ...) write to the process's stdout; the contract is ONE JSON
# line there.  Keep the real stdout aside and point fd 1 at stderr for everything else.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peaks_all() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region (nvidia-smi, 100 ms)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except ValueError:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- workloads
def build_knn_inputs(args):
    from vrec import synth
    t0 = time.time()
    v, places = synth.g2_place_visits(args.knn_persons, args.knn_places, seed=20181231, region=0)
    inp = synth.build_rating_vectors(v)
    log(f"[bench] G2 region-set: P={len(inp.person_id)} nnz_place={len(inp.place_col)} "
        f"nnz_cat={len(inp.cat_col)} B_region={inp.algorithmic_bytes / 1e6:.1f} MB ({time.time() - t0:.1f}s)")
    return inp, places


def knn_workload_name(args):
    return (f"knn_batch P={args.knn_persons} places={args.knn_places} K={args.k_nearest} pw=cw=0.5 "
            f"top{args.max_recs} (BASELINE config 3, generator G2)")


def knn_config(args):
    """The `config` object: the workload only, the same keys and values in both arms; how an arm runs it
    (targets per step, sharding, L2 note) is that arm's `execution` object."""
    return {"workload": knn_workload_name(args), "persons": args.knn_persons, "places": args.knn_places,
            "k_nearest": args.k_nearest, "place_weight": 0.5, "category_weight": 0.5, "max_recs": args.max_recs,
            "generator": "G2 seed 20181231 region 0"}


def sg_workload_name(args):
    return (f"sg_power_iteration N={args.sg_vertices} in_degree={args.sg_degree} "
            f"nnz={args.sg_vertices * args.sg_degree} eps=0 {args.sg_iterations} iterations/step "
            f"(BASELINE config 5 shape, device generator)")


def ncu_summary(kernel: str):
    """DRAM traffic (and other counters) of one launch of `kernel` on the default workload, from the committed
    summary of an `ncu --set full` capture (tools/ncu_summary.py writes profiles/r2_ncu_summary.json).  They are NOT
    measured in this run -- a profiler cannot wrap a timed region -- so the line names the file, and reports null
    when the file has no entry for the kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def tensor_peak(clocks):
    """MEASURED_PEAKS.json has a burst figure (a kernel timed alone at full clocks) and a sustained one (a long
    loop under the power cap).  The regime is read off the SM clocks sampled during the timed region."""
    peaks = measured_peaks_all()
    burst, sust = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    if not burst:
        return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s burst)"
    sm, mx = (clocks or {}).get("sm_mhz"), (clocks or {}).get("sm_max_mhz")
    if sust and sm and mx and sm < 0.9 * mx:
        return float(sust), f"measured (MEASURED_PEAKS.json bf16_tflops_sustained: median SM clock {sm:.0f} of {mx:.0f} MHz)"
    return float(burst), (f"measured (MEASURED_PEAKS.json bf16_tflops, burst: median SM clock "
                          f"{sm if sm else float('nan'):.0f} of {mx if mx else float('nan'):.0f} MHz in the timed region)")


def sg_bytes_per_iteration(n, nnz):
    return 12 * nnz + 20 * n


def oracle_knn_data(oracle, inp):
    row = np.searchsorted(inp.person_id, inp.rating_person)
    order = np.argsort(row, kind="stable")
    rowptr = np.zeros(len(inp.person_id) + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    return oracle.KnnData(inp.person_id, inp.place_rowptr, inp.place_col, inp.place_val,
                          inp.cat_rowptr, inp.cat_col, inp.cat_val,
                          np.cumsum(rowptr), inp.rating_place[order], inp.rating_value[order],
                          place_dim=inp.place_dim)


def knn_count_bad(got, answers):
    """Targets whose recommended place ids / order, estimated-rating bits, count or status differ from the oracle's."""
    opl, ort, ocnt, ost = answers[:4]
    pl, rt, cnt, st = got
    bad = 0
    for q in range(len(ocnt)):
        c = int(ocnt[q])
        if not (int(cnt[q]) == c and int(st[q]) == int(ost[q]) and np.array_equal(pl[q, :c], opl[q, :c])
                and np.array_equal(np.ascontiguousarray(rt[q, :c]).view(np.int64),
                                   np.ascontiguousarray(ort[q, :c]).view(np.int64))):
            bad += 1
    return bad


def knn_parity(rec, targets, places, max_recs, answers):
    """Engine (vrec_knn_query, host buffers) against the oracle's answers for the same targets of the full-size
    region-set: recommended place ids and their order, estimated ratings bit for bit, counts and statuses."""
    bad = knn_count_bad(rec.recommend(targets, np.ascontiguousarray(places.id, dtype=np.int64), max_recs), answers)
    return {"targets": int(len(targets)), "mismatching_targets": bad, "bit_exact": bad == 0,
            "checked": "place ids + order, estimated_rating bits, counts, statuses against oracle/vrec_oracle.c "
                       "on the full-size region-set (the cpu_baseline sample)"}


def cpu_knn(args, inp, places, n_targets, repeats=1, check=None, d=None):
    """CPU oracle (port of the Scala path) on a bounded sample of the same workload."""
    from oracle import oracle
    oracle.build()
    if d is None:
        d = oracle_knn_data(oracle, inp)
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(7)
    targets = inp.person_id[rng.choice(len(inp.person_id), n_targets, replace=False)]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rc, *answers = oracle.knn_query_batch(d, targets, 0.5, 0.5, args.k_nearest, places.id, args.max_recs,
                                              n_threads=threads)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    if check is not None:           # the oracle as the checker: same targets through the engine, full size
        step = check.pop("step", None)
        check.update(knn_parity(check.pop("rec"), targets, places, args.max_recs, answers))
        try:        # and targets spread over the last timed end-to-end step (a full batch per call)
            ids, got = step
            rc, *ans = oracle.knn_query_batch(d, ids, 0.5, 0.5, args.k_nearest, places.id, args.max_recs,
                                              n_threads=threads)
            bad = knn_count_bad(got, ans) if rc == 0 else len(ids)
            check["timed_step"] = {"targets": int(len(ids)), "mismatching_targets": bad, "bit_exact": bad == 0}
        except Exception as e:          # the check must never cost the bench line
            check["timed_step"] = {"error": f"{type(e).__name__}: {e}"}
    return {"value": n_targets / best, "unit": "persons/s", "cores": threads, "kind": "port",
            "sample": f"{n_targets} random targets of the same region-set, {best:.2f}s, oracle/vrec_oracle.c "
                      f"with OpenMP over targets"}, best


def cpu_sg(args, n_vertices, iterations):
    """CPU oracle SpMV on a smaller graph of the same generator (host RAM bound)."""
    from oracle import oracle
    oracle.build()
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(5)
    deg = args.sg_degree
    # same shape as sg_gen_rows_kernel: every row has `deg` in-edges, half uniform / half skewed sources
    u = rng.random((n_vertices, deg))
    src = np.where((np.arange(deg) % 2 == 0)[None, :],
                   ((u ** 3 * n_vertices).astype(np.int64) * 2654435761 + 12345) % n_vertices,
                   (u * n_vertices).astype(np.int64)).astype(np.int32)
    del u
    src.sort(axis=1)
    rowptr = np.arange(n_vertices + 1, dtype=np.int64) * deg
    g = oracle.SgGraph.from_csr(rowptr, src.ravel(), np.full(n_vertices * deg, 1.0 / deg))
    del src
    oracle.set_threads(threads)
    t0 = time.perf_counter()
    rc, x, it, conv, res = g.run(int(g.ids[0]), 0.0, iterations)
    dt = time.perf_counter() - t0
    assert rc == 0 and it == iterations
    gbs = sg_bytes_per_iteration(g.N, g.nnz) * iterations / dt / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"N={g.N} nnz={g.nnz} same generator shape, {iterations} iterations, {dt:.2f}s, "
                      f"oracle/vrec_oracle.c with OpenMP over rows"}, dt


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp, places = build_knn_inputs(args)
    # a step = a bounded sample of the workload, sized so that the whole run ends within a few minutes
    per_step = min(args.ref_knn_targets, max(64, 8192 // max(1, args.warmup + args.steps)))
    from oracle import oracle
    oracle.build()
    d = oracle_knn_data(oracle, inp)
    times = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_knn(args, inp, places, per_step, d=d)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = per_step * len(times) / total
    sg_cb, _ = cpu_sg(args, args.ref_sg_vertices, 10)
    cb["value"] = value
    cb["sample"] = (f"{per_step} random targets of the same region-set per step, {len(times)} timed steps, "
                    f"oracle/vrec_oracle.c with OpenMP over targets")
    line = {
        "impl": "reference", "metric": "KNN target persons/sec (sim+top-K+rating)", "value": value,
        "unit": "persons/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": knn_config(args),
        "execution": {"step": f"{per_step} targets per step on {cb['cores']} host threads (a bounded sample: the "
                              f"CPU path needs ~9 ms per target)", "gpus_used": 0},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "persons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sg": {"metric": "SG power-iteration GB/s (algorithmic bytes)", "value": sg_cb["value"],
               "unit": "GB/s", "cpu_baseline": sg_cb},
        "note": "reference = Scala/Spark, not runnable here (no JVM); this arm is the C oracle port on all host cores",
    }
    emit(line)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import vrec
    from vrec import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = vrec.Context(local_rank)
    if world > 1:
        from vrec import dist as vdist
        vdist.init_comm(ctx)          # NCCL communicator of the library (row-partitioned SG path)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    lib = ctx.lib
    peak, peak_src = measured_peaks()

    # ---------------- KNN (headline)
    inp, places = build_knn_inputs(args)
    t0 = time.time()
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    load_s = time.time() - t0
    rec = vrec.KnnRecommender(rs, 0.5, 0.5, args.k_nearest)
    B, calls = args.knn_batch, max(1, args.knn_calls)
    n_steps_total = args.warmup + args.steps
    P = len(inp.person_id)
    # each rank walks its own contiguous range of targets (weak scaling: calls x B targets per GPU per step)
    tgt_ids = np.stack([inp.person_id[(np.arange(B) + (c * world + rank) * B) % P]
                        for c in range(n_steps_total * calls)]).reshape(n_steps_total, calls, B)
    d_targets = torch.from_numpy(tgt_ids).cuda()
    m = args.max_recs
    d_place = torch.empty((B, m), dtype=torch.int64, device="cuda")
    d_rating = torch.empty((B, m), dtype=torch.float64, device="cuda")
    d_count = torch.empty(B, dtype=torch.int32, device="cuda")
    d_status = torch.empty(B, dtype=torch.int32, device="cuda")
    flt = np.ascontiguousarray(places.id, dtype=np.int64)
    assert lib.vrec_knn_set_filter(rs._h, flt.ctypes.data_as(L.i64p), len(flt)) == 0
    torch.cuda.synchronize()

    def knn_step_device(s):
        for c in range(calls):
            rc = lib.vrec_knn_query_device(rs._h, d_targets[s, c].data_ptr(), B, 0.5, 0.5, args.k_nearest, m,
                                           d_place.data_ptr(), d_rating.data_ptr(), d_count.data_ptr(),
                                           d_status.data_ptr())
            if rc != 0:
                raise RuntimeError(L.last_error())

    for s in range(args.warmup):
        knn_step_device(s)
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count
    barrier()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        ev[i][0].record(stream)
        knn_step_device(args.warmup + i)
        ev[i][1].record(stream)
    ctx.synchronize()
    barrier()
    knn_ms = [a.elapsed_time(b) for a, b in ev]
    knn_total_ms = max_over_ranks(sum(knn_ms))
    knn_launches = ctx.launch_count - launches0
    clocks_knn = sampler.stop()
    status_ok = int((d_status == 0).sum().item())
    value = world * B * calls * args.steps / (knn_total_ms / 1e3)
    log(f"[bench] knn device: {value:,.0f} persons/s over {world} GPU(s); per-step ms {['%.1f' % x for x in knn_ms[:8]]}; "
        f"status ok {status_ok}/{B}")
    # the dense filter's duration: CUDA events the library records around that launch on its own stream (last call)
    import ctypes as C
    dense_ms = C.c_double(0.0)
    lib.vrec_knn_last_dense_ms(rs._h, C.byref(dense_ms))

    # end to end through the host-buffer ABI call, pinned host buffers
    h_targets = torch.from_numpy(tgt_ids).pin_memory()
    h_place = torch.empty((calls, B, m), dtype=torch.int64).pin_memory()
    h_rating = torch.empty((calls, B, m), dtype=torch.float64).pin_memory()
    h_count = torch.empty((calls, B), dtype=torch.int32).pin_memory()
    h_status = torch.empty((calls, B), dtype=torch.int32).pin_memory()

    def C_i64(p):
        return C.cast(p, L.i64p)

    def C_f64(p):
        return C.cast(p, L.f64p)

    def C_i32(p):
        return C.cast(p, L.i32p)

    def knn_step_host(s):
        for c in range(calls):
            rc = lib.vrec_knn_query(rs._h, C_i64(h_targets[s, c].data_ptr()), B, 0.5, 0.5, args.k_nearest,
                                    flt.ctypes.data_as(L.i64p), len(flt), m, C_i64(h_place[c].data_ptr()),
                                    C_f64(h_rating[c].data_ptr()), C_i32(h_count[c].data_ptr()),
                                    C_i32(h_status[c].data_ptr()))
            if rc != 0:
                raise RuntimeError(L.last_error())

    knn_step_host(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        knn_step_host(args.warmup + i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * calls * args.steps / e2e_s
    # bytes the host-buffer call moves per step, from the arrays it is given: target ids and the region's place ids
    # in, (place, rating) rows + counts + statuses out
    h2d = calls * (int(h_targets[0, 0].numel()) * 8 + flt.nbytes)
    d2h = calls * (int(h_place[0].numel()) * 8 + int(h_rating[0].numel()) * 8 + int(h_count[0].numel()) * 4 +
                   int(h_status[0].numel()) * 4)
    log(f"[bench] knn e2e: {e2e_value:,.0f} persons/s")

    # Roofline of the dominant KNN kernel, knn_tc_ws_kernel: a [B x 128] x [128 x P] fp16 product on tcgen05 with
    # the threshold filter fused into its TMEM epilogue.  (SURVEY.md 8(d)'s per-target byte figure B_region / T,
    # T = targets per pass, is reported alongside; with T = 18944 it is ~11 KB per target.)
    b_region = inp.algorithmic_bytes
    knn_call_ms = statistics.mean(knn_ms) / calls
    P_ = len(inp.person_id)
    knn_roof = None
    if dense_ms.value > 0:
        tc_ms = dense_ms.value
        tc_flops = 2.0 * B * P_ * 128                      # [B x 128] x [128 x P] fp16, fp32 accumulate
        tc_bytes = ((B + 127) // 128) * P_ * 256           # every CTA streams the feature matrix once (from L2)
        tf_peak, tf_src = tensor_peak(clocks_knn)
        prof = ncu_summary("knn_tc_ws_kernel") or {}
        knn_roof = {"bound": "tensor", "kernel": "knn_tc_ws_kernel", "achieved": tc_flops / (tc_ms / 1e3) / 1e12,
                    "peak": tf_peak, "unit": "TFLOP/s", "frac": tc_flops / (tc_ms / 1e3) / 1e12 / tf_peak,
                    "traffic": prof.get("dram_bytes"), "traffic_source": prof.get("source"),
                    "peak_source": tf_src,
                    "kernel_ms": tc_ms, "kernel_share_of_step": tc_ms / knn_call_ms,
                    "tensor_pipe_active_pct": prof.get("tensor_pipe_active_pct"),
                    "l2_to_sm_gbs": tc_bytes / (tc_ms / 1e3) / 1e9,
                    "survey_bytes_per_target": b_region / B,
                    "note": f"flops/launch = 2 x {B} targets x {P_} candidates x 128 dims, one launch per call of "
                            f"{B} targets ({calls} calls per step); kernel_ms = CUDA events around the launch on the "
                            f"library's stream, last call of the timed region.  The operand stream is {tc_bytes} B per "
                            "launch from L2 (neighbouring CTAs walk the tiles one apart, so HBM delivers each 32 KB tile "
                            "image once).  The kernel is paced by its MMA-issue / accumulator hand-over chain and the "
                            "consumer warps' exact evaluations, not by HBM: see DESIGN.md"}
    else:
        log("[bench] the dense filter kernel did not run in the last call: no KNN roofline")

    cpu_knn_base = knn_par = None
    if rank == 0 and not args.no_cpu_baseline:
        knn_par = {"rec": rec}
        try:        # 128 targets spread over the last end-to-end step, whose answers are still in the host buffers
            rows = np.arange(0, B, max(1, B // 128))[:128]
            lc = calls - 1
            knn_par["step"] = (tgt_ids[args.warmup + args.steps - 1][lc][rows],
                               (h_place.numpy()[lc][rows], h_rating.numpy()[lc][rows], h_count.numpy()[lc][rows],
                                h_status.numpy()[lc][rows]))
        except Exception as e:
            log(f"[bench] timed-step parity sample unavailable: {e}")
        cpu_knn_base, _ = cpu_knn(args, inp, places, args.cpu_knn_targets, check=knn_par)
        log(f"[bench] knn cpu baseline: {cpu_knn_base['value']:.1f} persons/s on {cpu_knn_base['cores']} threads; "
            f"full-size parity on its {knn_par['targets']} targets: "
            f"{'bit-exact' if knn_par['bit_exact'] else 'MISMATCH on %d targets' % knn_par['mismatching_targets']}; "
            f"timed step: {knn_par.get('timed_step')}")
    # one person at a time (the launchers' REPL, BASELINE configs 1-2): latency of a single query through the ABI
    single = None
    if rank == 0:
        def one(rec_, tgt):
            t1 = time.perf_counter()
            rec_.recommend([int(tgt)], flt, m)
            return (time.perf_counter() - t1) * 1e3
        rec_k = vrec.KnnRecommender(rs, 0.5, 0.5, args.k_nearest)
        rec_all = vrec.KnnRecommender(rs, 0.5, 0.5, 2_000_000)        # bin/knn_recommender.sh:32-35 default K
        pers = inp.person_id
        one(rec_k, pers[0]); one(rec_all, pers[0])
        single = {"knn_ms_k50": statistics.median(one(rec_k, pers[7 + 13 * i]) for i in range(9)),
                  "knn_ms_k2000000": statistics.median(one(rec_all, pers[11 + 17 * i]) for i in range(9)),
                  "note": "median wall time of vrec_knn_query for ONE person (ids in, top-10 out), region-set resident"}
        log(f"[bench] single query: K=50 {single['knn_ms_k50']:.2f} ms, K=2000000 {single['knn_ms_k2000000']:.2f} ms")
    rs_resident = rs.resident_bytes
    rs.close()
    del d_targets

    # ---------------- SG (second half of the metric)
    sg = None
    if not args.no_sg:
        sg = run_sg(args, vrec, ctx, stream, world, rank, barrier, max_over_ranks, peak, peak_src)
        sg["batch"] = run_sg_batch(args, vrec, ctx, world, rank, barrier, max_over_ranks)
    builder = run_builder(args, vrec, ctx, rank) if rank == 0 and not args.no_sg else None
    default_data = None
    if rank == 0 and not args.no_sg:
        try:
            default_data = run_default_data(args, vrec, ctx)
        except Exception as e:          # must never cost the headline line
            default_data = {"error": f"{type(e).__name__}: {e}"}
            log(f"[bench] default-data configs failed: {default_data['error']}")

    if rank == 0:
        line = {
            "metric": "KNN target persons/sec (sim+top-K+rating)", "value": value, "unit": "persons/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": knn_total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": knn_config(args),
            "execution": {"step": f"{calls} calls x {B} targets per GPU per step (vrec_knn_query_device / vrec_knn_query)",
                          "parallelism": f"targets sharded over {world} GPU(s), region-set replicated, no collective",
                          "l2": f"inputs larger than L2 (region-set {b_region / 1e6:.0f} MB + {rs_resident / 1e6:.0f} MB of "
                                "derived tables > 126 MB) and new targets every call",
                          "load_seconds": round(load_s, 2)},
            "e2e": {"value": e2e_value, "unit": "persons/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(knn_launches),
            "roofline": knn_roof,
            "cpu_baseline": cpu_knn_base,
            "parity": knn_par,
            "clocks": clocks_knn,
            "sg": sg,
            "single_query": single,
            "default_data": default_data,
            "builder": builder,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def sg_parity_config5(args, vrec, ctx, world, rank):
    """Config 5 parity (row-partitioned graph at N = 10^7): a graph of the same generator and vertex count with a
    smaller degree, so that the oracle finishes in seconds.  Rank 0 also builds the unpartitioned graph, exports its
    CSR and runs the oracle on it; EVERY rank's result of the partitioned query must equal the oracle's x bit for
    bit, with the same iteration count and flag."""
    import torch
    import torch.distributed as dist
    N, deg, eps, max_it = args.sg_vertices, args.sg_parity_degree, 1e-5, 6
    g = vrec.StochasticGraph.generate(N, deg, seed=5, rank=rank, world=world, ctx=ctx)
    rec = vrec.StochasticRecommender(g, eps, max_it)
    x = rec.stationary(0)
    mine = (rec.last_iterations, rec.last_converged)
    g.close()
    want = torch.zeros(N, dtype=torch.float64, device="cuda")
    meta = torch.zeros(3, dtype=torch.int64, device="cuda")
    if rank == 0:
        from oracle import oracle
        oracle.build()
        oracle.set_threads(os.cpu_count() or 1)
        g1 = vrec.StochasticGraph.generate(N, deg, seed=5, ctx=ctx) if world > 1 else \
            vrec.StochasticGraph.generate(N, deg, seed=5, rank=0, world=1, ctx=ctx)
        rowptr, src, w = g1.export_csr()
        g1.close()
        og = oracle.SgGraph.from_csr(rowptr.astype(np.int64), src, w)
        t0 = time.perf_counter()
        rc, ox, oit, oconv, _ = og.run(0, eps, max_it)
        odt = time.perf_counter() - t0
        del og, rowptr, src, w
        want.copy_(torch.from_numpy(ox))
        meta.copy_(torch.tensor([rc, oit, oconv]))
    if world > 1:
        dist.broadcast(want, 0)
        dist.broadcast(meta, 0)
    rc, oit, oconv = (int(v) for v in meta.tolist())
    ok = rc == 0 and mine == (oit, oconv) and bool(np.array_equal(x, want.cpu().numpy()))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out = {"workload": f"N={N} in_degree={deg} eps={eps} maxIt={max_it}, same device generator as the timed graph",
           "world": world, "ranks_bit_exact": bool(int(flag.item())), "iterations": mine[0], "converged": mine[1],
           "checked": "every rank's full stationary vector (bits), iteration count and convergence flag against "
                      "oracle/vrec_oracle.c on the exported CSR"}
    if rank == 0:
        out["oracle_seconds"] = round(odt, 2)
    return out


def run_sg(args, vrec, ctx, stream, world, rank, barrier, max_over_ranks, peak, peak_src):
    import torch
    N, deg, iters = args.sg_vertices, args.sg_degree, args.sg_iterations
    t0 = time.time()
    # one oversized graph: rows of P^T partitioned over the ranks; the sweep kernel stores x' into the peers' buffers
    g = vrec.StochasticGraph.generate(N, deg, seed=5, rank=rank, world=world, ctx=ctx)
    log(f"[bench] sg graph generated on device: N={g.N} nnz={g.nnz} ({time.time() - t0:.1f}s, "
        f"{g.resident_bytes / 1e9:.2f} GB resident)")
    bytes_it = sg_bytes_per_iteration(g.N, N * deg)          # whole graph, all ranks
    rows_mine = g.nnz // deg
    bytes_rank = 12 * g.nnz + 4 * rows_mine + 8 * g.N + 8 * rows_mine   # this rank's rows + the whole x it gathers from
    for _ in range(args.warmup):
        g.iterate_device(iters)
    ctx.synchronize()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    l0 = ctx.launch_count
    barrier()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        ev[i][0].record(stream)
        g.iterate_device(iters)
        ev[i][1].record(stream)
    ctx.synchronize()
    barrier()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = max_over_ranks(sum(ms))
    clocks = sampler.stop()
    launches = ctx.launch_count - l0
    gbs = bytes_it * iters * args.steps / (total_ms / 1e3) / 1e9
    per_it_ms = (total_ms / args.steps) / iters
    log(f"[bench] sg device: {gbs:,.0f} GB/s algorithmic, {per_it_ms * 1e3:.0f} us/iteration (max over ranks)")
    # end to end: one query through the host ABI (vertex id in, top-10 out), graph resident; random start vertices,
    # a tolerance no query meets, so that every query runs `iters` iterations
    rng = np.random.default_rng(17)
    qv = rng.integers(0, N, size=args.steps + 1)
    rec = vrec.StochasticRecommender(g, 0.0, iters)
    rec.recommend([int(qv[0])], None, 10)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        rec.recommend([int(qv[i + 1])], None, 10)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_gbs = bytes_it * iters * args.steps / e2e_s / 1e9
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu, _ = cpu_sg(args, args.cpu_sg_vertices, 10)
        log(f"[bench] sg cpu baseline: {cpu['value']:.2f} GB/s on {cpu['cores']} threads")
    prof = (ncu_summary("sg_spmv_kernel") or {}) if world == 1 and N == 10_000_000 and deg == 100 else {}
    blocks = (N + 6291456 - 1) // 6291456
    out = {
        "metric": "SG power-iteration HBM GB/s (algorithmic bytes 12*nnz + 20*N per iteration)",
        "value": gbs, "unit": "GB/s", "ms_per_step": total_ms / args.steps, "us_per_iteration": per_it_ms * 1e3,
        "n_gpus": world, "scaling": "strong",
        "config": {"workload": sg_workload_name(args), "l2": f"inputs larger than L2 ({bytes_it / 1e9:.1f} GB/iteration)",
                   "parallelism": "single GPU" if world == 1 else
                   f"rows of P^T partitioned over {world} GPUs; every rank's sweep stores its {8 * N / world / 1e6:.0f} MB "
                   f"of x' into the {world - 1} peers' buffers over NVLink (CUDA IPC mappings), residual partials in a "
                   "slot per rank, one flag barrier per iteration -- no collective inside the iteration"},
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 10 * 16 + 24,
                "note": "vrec_sg_query: one random vertex id in, ranked top-10 of all vertices out"},
        "roofline": {"bound": "hbm", "kernel": "sg_spmv_kernel",
                     "achieved": bytes_rank / (per_it_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_rank / (per_it_ms / 1e3) / 1e9 / peak,
                     "traffic": prof.get("dram_bytes"), "traffic_source": prof.get("source"),
                     "peak_source": peak_src,
                     "note": f"per GPU: this rank's algorithmic bytes / (step time / iterations); one iteration = one "
                             f"sg_spmv_kernel launch per source block ({blocks} at N = {N})"
                             + ("" if world == 1 else " + the exchange barrier; the peer stores of x' ride inside the sweep")
                             + ".  The sweep is bound by the L1TEX pipe (one wavefront per scattered 8-byte gather of "
                             "x): a kernel that only streams (src, w) and gathers x tops out at 0.48 of the HBM "
                             "roofline on this GPU (tools/gather_probe.cu, profiles/r2_gather_probe_flat_rows.log)"},
        "cpu_baseline": cpu, "clocks": clocks,
    }
    g.close()
    try:
        out["parity"] = sg_parity_config5(args, vrec, ctx, world, rank)
        if rank == 0:
            log(f"[bench] sg config-5 parity (world {world}): "
                f"{'bit-exact on every rank' if out['parity']['ranks_bit_exact'] else 'MISMATCH'}")
    except Exception as e:          # the check must never cost the bench line
        out["parity"] = {"error": f"{type(e).__name__}: {e}"}
    return out


REGION_SETS = [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2)]      # singles ++ 2-combinations, PlaceVisits.scala:63-67


def visit_rows(v, seed):
    """PlaceVisitCounts -> one row per visit (person, place, category, timestamp in ms inside the 7-day window): the
    input of the device builders (the reference's builders read place_visits rows, PlaceVisits.scala:116-121)."""
    rng = np.random.default_rng(seed)
    pe = np.repeat(v.person_id, v.count)
    pl = np.repeat(v.place_id, v.count)
    ca = np.repeat(v.category_id, v.count)
    ts = 1_546_300_800_000 - rng.integers(0, 7 * 24 * 3600 * 1000, len(pe))
    return pe, pl, ca, ts


def run_sg_batch(args, vrec, ctx, world, rank, barrier, max_over_ranks):
    """BASELINE config 4: a recommendation for the persons of EVERY per-region and pairwise graph (3 + 3,
    PlaceVisits.scala:63-67; graph selection of StochasticRecommenderMain.scala:53-62,83-89) of sample-generator
    shaped data (G1, independent coordinates).  Unit of work = (graph, range of its persons); units are assigned to
    the ranks by vrec.dist.assign_units (no collective: SURVEY 8(e)); every step each unit answers `--sgb-batch`
    persons of its range through vrec_sg_query with host buffers (ids in, ranked top-10 places out)."""
    import torch.distributed as dist
    from vrec import builders, synth
    from vrec import dist as vdist
    t0 = time.time()
    places = synth.sample_places(3 * args.sgb_places, seed=0)
    per_region = args.sgb_persons
    visits = [synth.sample_place_visits(places, r, persons_per_region=per_region, person_count_total=3 * per_region,
                                        correlated=False, seed=0) for r in range(3)]
    persons_of = [np.unique(v.person_id) for v in visits]
    graph_persons = [np.concatenate([persons_of[r] for r in rs_]) for rs_ in REGION_SETS]
    units = []                                                     # (graph, lo, hi) over graph_persons[graph]
    for gi, pp in enumerate(graph_persons):
        for lo in range(0, len(pp), args.sgb_unit):
            units.append((gi, lo, min(len(pp), lo + args.sgb_unit)))
    mine = [units[i] for i in vdist.assign_units([hi - lo for _, lo, hi in units], world)[rank]]
    graphs, edges = {}, {}
    for gi in sorted({u[0] for u in mine}):
        v = synth.merge_visits([visits[r] for r in REGION_SETS[gi]])
        s_, t_, w_ = builders.stochastic_graph_builder(*visit_rows(v, 40 + gi), ctx=ctx)     # the four edge families
        graphs[gi] = vrec.StochasticGraph(s_, t_, w_, ctx=ctx)
        edges[gi] = (s_, t_, w_)
    log(f"[bench] sg batch (config 4): {len(units)} units over {len(REGION_SETS)} graphs, {len(mine)} on rank {rank}; graphs "
        + ", ".join(f"{REGION_SETS[gi]}: N={g.N} nnz={g.nnz} active={g.batch_info(2)}" for gi, g in graphs.items())
        + f" ({time.time() - t0:.1f}s)")
    flt = {gi: np.ascontiguousarray(np.concatenate([places.of_region(r) for r in REGION_SETS[gi]])) for gi in graphs}
    # the target region of a pairwise query is one of the two; recommend among the places of the whole region-set
    recs = {gi: vrec.StochasticRecommender(g, 0.01, 20) for gi, g in graphs.items()}     # stochastic_recommender.sh:32-35
    rng = np.random.default_rng(100 + rank)
    n_q = args.sgb_batch
    # the batch kernel (one launch for any number of persons) must be what serves these graphs: a graph it cannot
    # take (see csrc/vrec_sg_batch.cu) would answer one person per launch, ~300x slower -- keep its share tiny then
    slow = set()
    for gi, g in graphs.items():
        recs[gi].recommend(graph_persons[gi][:8], flt[gi], 10)
        if g.batch_info(0) != 8:
            slow.add(gi)
            log(f"[bench] sg batch: graph {REGION_SETS[gi]} is not served by the batch kernel; 16 persons per unit per step")

    def step(keep=None):
        n = 0
        its_sum = 0
        for (gi, lo, hi) in mine:
            k = 16 if gi in slow else n_q
            q = graph_persons[gi][lo + rng.choice(hi - lo, min(k, hi - lo), replace=False)]
            out = recs[gi].recommend(q, flt[gi], 10)
            n += len(q)
            its_sum += int(out[3].sum())
            if keep is not None:
                keep[gi] = (q, out)
        return n, its_sum

    for _ in range(min(2, args.warmup)):
        step()
    l0 = ctx.launch_count
    barrier()
    t0 = time.perf_counter()
    n_local = its_total = 0
    last = {}
    for i in range(args.steps):
        n_, it_ = step(last if i == args.steps - 1 else None)
        n_local += n_
        its_total += it_
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    launches = ctx.launch_count - l0
    n_total = n_local
    if world > 1:
        import torch
        t = torch.tensor([n_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_total = int(t.item())
    value = n_total / dt
    # parity: every rank checks the graphs it owns against the oracle -- 4 persons of each graph through a fresh
    # call, and 4 spread over the last timed step's answers
    par_rows, cpu_time, cpu_n = [], 0.0, 0
    if not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        oracle.set_threads(os.cpu_count() or 1)
        for gi, g in graphs.items():
            og = oracle.SgGraph(*edges[gi])

            def count_bad(got, rows, wanted):
                g_id, g_pr, g_cnt, g_its, g_conv, g_st = got
                n_bad = 0
                for r, (rc, wi, wp, oit, oconv) in zip(rows, wanted):
                    c = len(wi)
                    if not (int(g_st[r]) == rc and int(g_cnt[r]) == c and (int(g_its[r]), int(g_conv[r])) == (oit, oconv)
                            and np.array_equal(g_id[r, :c], wi)
                            and np.array_equal(np.ascontiguousarray(g_pr[r, :c]).view(np.int64),
                                               np.asarray(wp, dtype=np.float64).view(np.int64))):
                        n_bad += 1
                return n_bad
            sample = graph_persons[gi][:: max(1, len(graph_persons[gi]) // 4)][:4]
            t1 = time.perf_counter()
            answers = [og.query(int(v), 0.01, 20, flt[gi], 10) for v in sample]
            cpu_time += time.perf_counter() - t1
            cpu_n += len(sample)
            bad = count_bad(recs[gi].recommend(sample, flt[gi], 10), range(len(sample)), answers)
            row = {"graph": list(REGION_SETS[gi]), "targets": len(sample), "mismatching_targets": bad}
            if gi in last:
                lq, lgot = last[gi]
                rows = list(range(0, len(lq), max(1, len(lq) // 4)))[:4]
                bad_step = count_bad(lgot, rows, [og.query(int(lq[r]), 0.01, 20, flt[gi], 10) for r in rows])
                row["timed_step"] = {"targets": len(rows), "mismatching_targets": bad_step}
                bad += bad_step
            row["bit_exact"] = bad == 0
            par_rows.append(row)
            del og
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, par_rows)
        par_rows = sorted({tuple(r["graph"]): r for rr in allp for r in rr}.values(), key=lambda r: r["graph"])
    parity = cpu = None
    if par_rows:
        parity = {"graphs": par_rows, "bit_exact": all(r["bit_exact"] for r in par_rows),
                  "checked": "place ids + order, probability bits, iterations, converged flags against "
                             "oracle/vrec_oracle.c on every full-size graph: a fresh sample and rows of the last timed step"}
        if rank == 0 and cpu_n:
            cpu = {"value": cpu_n / cpu_time, "unit": "persons/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"{cpu_n} persons over the graphs rank 0 owns, one query at a time, {cpu_time:.2f}s, "
                             f"oracle/vrec_oracle.c with OpenMP over rows"}
            log(f"[bench] sg batch: {value:,.0f} persons/s; cpu baseline {cpu['value']:.1f} persons/s; parity "
                f"{'bit-exact on all %d graphs' % len(par_rows) if parity['bit_exact'] else 'MISMATCH'}")
    out = {
        "metric": "SG recommendations, persons/s (power iteration to eps=0.01 + ranked top-10 per person)",
        "value": value, "unit": "persons/s", "ms_per_step": 1e3 * dt / args.steps, "n_gpus": world, "scaling": "strong",
        "config": {"workload": f"sg_batch: 3 regions x {per_region} persons, {args.sgb_places} places per region, 20 "
                               f"categories (sample-generator shape, independent coordinates); the 3 per-region + 3 "
                               f"pairwise graphs of the device builder; eps=0.01 maxIt=20 top10 (BASELINE config 4)",
                   "units": len(units), "persons_per_unit_per_step": n_q,
                   "parallelism": f"{len(units)} (graph, person-range) units assigned to {world} GPU(s) by cost, a graph "
                                  "is loaded where a unit needs it, no collective"},
        "e2e": {"value": value, "unit": "persons/s",
                "h2d_bytes_per_step": int(sum(8 * min(n_q, hi - lo) + flt[gi].nbytes for gi, lo, hi in mine)),
                "d2h_bytes_per_step": int(sum(min(n_q, hi - lo) * (10 * 16 + 16) for gi, lo, hi in mine))},
        "mean_iterations": its_total / max(1, n_local),
        "graphs_without_batch_kernel": sorted(list(REGION_SETS[gi]) for gi in slow),
        "gpu_launches": int(launches), "cpu_baseline": cpu, "parity": parity,
        "note": "value == e2e: every step goes through vrec_sg_query with host buffers; h2d / d2h bytes are rank 0's",
    }
    for g in graphs.values():
        g.close()
    return out


def run_default_data(args, vrec, ctx):
    """BASELINE configs 1 and 2: the sample generator's DEFAULT data (bin/sample_generator.sh:32-33: 10 000 places and
    1 M persons per region; generator G1 with the reference's correlated coordinates), region 0, through the device
    builders, then ONE person at a time as the launchers' REPL does -- KNN with the launcher's K = 2 000 000
    (bin/knn_recommender.sh:32-36), SG with eps = 0.01 / maxIt = 20 (bin/stochastic_recommender.sh:32-35) -- next
    to the CPU port on the same queries, results compared bit for bit."""
    import tempfile
    from oracle import oracle
    from vrec import builders, synth
    from vrec import data_utils as du
    oracle.build()
    threads = os.cpu_count() or 1
    oracle.set_threads(threads)
    t0 = time.time()
    places = synth.sample_places(30000, seed=0)
    v = synth.sample_place_visits(places, 0, persons_per_region=args.g1_persons, person_count_total=3 * args.g1_persons,
                                  seed=0)
    pe, pl, ca, ts = visit_rows(v, 3)
    flt = np.ascontiguousarray(places.of_region(0))
    gen_s = time.time() - t0
    out = {"workload": f"sample-generator defaults, region 0: {args.g1_persons} persons, 10 000 places, 20 categories, "
                       f"7-day window -> {len(pe)} place visits of {len(np.unique(pe))} persons (G1, correlated coordinates)",
           "generate_seconds": round(gen_s, 2)}
    # ---- config 1: KNN, one person, K = 2 000 000
    t0 = time.perf_counter()
    inp = builders.rating_vectors_builder(pe, pl, ca, ctx=ctx)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    rs = vrec.KnnRegionSet(*inp.load_args(), ctx=ctx)
    load_s = time.perf_counter() - t0
    K = 2_000_000
    rec = vrec.KnnRecommender(rs, 0.5, 0.5, K)
    d = oracle_knn_data(oracle, inp)
    qs = inp.person_id[:: max(1, len(inp.person_id) // 9)][:9]
    rec.recommend([int(qs[0])], flt, 10)
    lat, cpu_lat, bad = [], [], 0
    for q in qs:
        t1 = time.perf_counter()
        got = rec.recommend([int(q)], flt, 10)
        lat.append((time.perf_counter() - t1) * 1e3)
        t1 = time.perf_counter()
        rc, *want = oracle.knn_query_batch(d, np.array([q]), 0.5, 0.5, K, flt, 10, n_threads=threads)
        cpu_lat.append((time.perf_counter() - t1) * 1e3)
        bad += knn_count_bad(got, want) if rc == 0 else 1
    # Parquet -> first answer: the three directories as RatingVectorsBuilderMain writes them
    with tempfile.TemporaryDirectory() as tmp:
        du.write_knn_inputs(inp, (0,), tmp)
        t1 = time.perf_counter()
        loaded = du.load_knn_inputs((0,), tmp, verbose=False)
        read_s = time.perf_counter() - t1
        rs2 = vrec.KnnRegionSet(*loaded, ctx=ctx)
        first = vrec.KnnRecommender(rs2, 0.5, 0.5, K).recommend([int(qs[0])], flt, 10)
        first_s = time.perf_counter() - t1
        rs2.close()
    out["knn_config1"] = {
        "persons_in_region_set": int(len(inp.person_id)), "k_nearest": K,
        "query_ms": statistics.median(lat), "cpu_query_ms": statistics.median(cpu_lat), "cpu_threads": threads,
        "parity": {"targets": len(qs), "mismatching_targets": bad, "bit_exact": bad == 0},
        "device_builder_seconds": round(build_s, 3), "load_seconds": round(load_s, 3),
        "parquet_to_first_answer_seconds": round(first_s, 3), "parquet_read_seconds": round(read_s, 3),
        "note": "query_ms = median wall time of vrec_knn_query for one person (id in, top-10 out) with the region-set "
                "resident; cpu_query_ms = the oracle port on the same person with all host threads; "
                "parquet_to_first_answer = pyarrow read of the three directories + vrec_knn_load + the first query"}
    rs.close()
    # ---- config 2: SG, one person, the default per-region graph
    t0 = time.perf_counter()
    s_, t_, w_ = builders.stochastic_graph_builder(pe, pl, ca, ts, ctx=ctx)
    gbuild_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = vrec.StochasticGraph(s_, t_, w_, ctx=ctx)
    gload_s = time.perf_counter() - t0
    og = oracle.SgGraph(s_, t_, w_)
    srec = vrec.StochasticRecommender(g, 0.01, 20)
    srec.recommend([int(qs[0])], flt, 10)
    lat, cpu_lat, bad, its = [], [], 0, 0
    for q in qs:
        t1 = time.perf_counter()
        oi, op, cnt, it_, conv, st = srec.recommend([int(q)], flt, 10)
        lat.append((time.perf_counter() - t1) * 1e3)
        t1 = time.perf_counter()
        rc, wi, wp, oit, oconv = og.query(int(q), 0.01, 20, flt, 10)
        cpu_lat.append((time.perf_counter() - t1) * 1e3)
        its = int(it_[0])
        c = len(wi)
        ok = (int(st[0]) == rc and int(cnt[0]) == c and (int(it_[0]), int(conv[0])) == (oit, oconv)
              and np.array_equal(oi[0, :c], wi) and np.array_equal(np.ascontiguousarray(op[0, :c]).view(np.int64),
                                                                   np.asarray(wp, dtype=np.float64).view(np.int64)))
        bad += 0 if ok else 1
    ms = statistics.median(lat)
    bytes_it = sg_bytes_per_iteration(g.N, g.nnz)
    peak, _src = measured_peaks()
    out["sg_config2"] = {
        "vertices": int(g.N), "edges": int(g.nnz), "epsilon": 0.01, "max_iterations": 20,
        "query_ms": ms, "iterations": its, "us_per_spmv": 1e3 * ms / (its + 1),
        "cpu_query_ms": statistics.median(cpu_lat), "cpu_threads": threads,
        "roofline": {"bound": "latency", "algorithmic_bytes_per_iteration": bytes_it,
                     "hbm_time_us_per_iteration": bytes_it / (peak * 1e9) * 1e6,
                     "note": "a per-region graph moves ~40 MB per iteration (6 us of HBM time): the query is bound by "
                             "launch / synchronisation latency, reported as us_per_spmv"},
        "parity": {"targets": len(qs), "mismatching_targets": bad, "bit_exact": bad == 0},
        "device_builder_seconds": round(gbuild_s, 3), "load_seconds": round(gload_s, 3),
        "note": "query_ms = median wall time of vrec_sg_query for one person (id in, ranked top-10 places of region 0 "
                "out), graph resident"}
    g.close()
    log(f"[bench] default data: knn one person K=2000000 {out['knn_config1']['query_ms']:.2f} ms (cpu "
        f"{out['knn_config1']['cpu_query_ms']:.0f} ms), parquet -> first answer {first_s:.2f}s; sg one person {ms:.3f} ms "
        f"/ {its} iterations (cpu {out['sg_config2']['cpu_query_ms']:.0f} ms); parity "
        f"{out['knn_config1']['parity']['bit_exact']} / {out['sg_config2']['parity']['bit_exact']}")
    return out


def run_builder(args, vrec, ctx, rank):
    """SURVEY 8(f) rank 2, the step in front of the KNN path: RatingsBuilder + RatingVectorsBuilder for the place
    column of the KNN workload's visits (one row per visit), through vrec_build_rating_vectors with host buffers."""
    from vrec import builders, synth
    v, _ = synth.g2_place_visits(args.knn_persons, args.knn_places)
    rng = np.random.default_rng(9)
    perm = rng.permutation(int(v.count.sum()))
    pe = np.repeat(v.person_id, v.count)[perm]
    pl = np.repeat(v.place_id, v.count)[perm]
    builders.build_rating_vectors(pe[:100000], pl[:100000], 100, ctx=ctx)          # warm-up (module load)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        out = builders.build_rating_vectors(pe, pl, 100, ctx=ctx)
    dt = (time.perf_counter() - t0) / reps
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        ns = min(len(pe), 4_000_000)
        t1 = time.perf_counter()
        rc, *_ = oracle.build_rating_vectors(pe[:ns], pl[:ns], None, 100)
        cdt = time.perf_counter() - t1
        cpu = {"value": ns / cdt, "unit": "visits/s", "cores": 1, "kind": "port",
               "sample": f"first {ns} visit rows, {cdt:.2f}s, oracle/vrec_oracle.c (qsort, one thread)"}
    # SURVEY 8(f) rank 3, the step in front of the SG path: the four edge families from place visits
    ng_persons = min(args.knn_persons, 200_000)
    vg, _ = synth.g2_place_visits(ng_persons, max(1000, args.knn_places // 5), seed=77)
    permg = rng.permutation(int(vg.count.sum()))
    gpe = np.repeat(vg.person_id, vg.count)[permg]
    gpl = np.repeat(vg.place_id, vg.count)[permg]
    gca = np.repeat(vg.category_id, vg.count)[permg]
    gts = 1_546_300_800_000 + rng.integers(0, 14 * 24 * 3600 * 1000, len(gpe))
    builders.stochastic_graph_builder(gpe[:50000], gpl[:50000], gca[:50000], gts[:50000], ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(reps):
        gs, gt, gw = builders.stochastic_graph_builder(gpe, gpl, gca, gts, ctx=ctx)
    gdt = (time.perf_counter() - t0) / reps
    gcpu = None
    if rank == 0 and not args.no_cpu_baseline:
        nsamp = min(len(gpe), 400_000)
        keep = gpe < np.sort(np.unique(gpe))[min(len(np.unique(gpe)) - 1, 25_000)]
        keep &= np.cumsum(keep) <= nsamp
        t1 = time.perf_counter()
        oracle.build_stochastic_graph(gpe[keep], gpl[keep], gca[keep], gts[keep], 0.5, 0.5)
        gcdt = time.perf_counter() - t1
        gcpu = {"value": int(keep.sum()) / gcdt, "unit": "visits/s", "cores": 1, "kind": "port",
                "sample": f"the visits of the first 25000 persons ({int(keep.sum())} rows), {gcdt:.2f}s, "
                          f"oracle/vrec_oracle.c (qsort, one thread)"}
    graph = {"metric": "stochastic graph builder, visit rows/s (4 edge families incl. the co-visit self-join)",
             "value": len(gpe) / gdt, "unit": "visits/s", "ms_per_step": gdt * 1e3,
             "config": {"workload": f"{len(gpe)} visit rows of {ng_persons} persons over 14 days -> {len(gs)} edges"},
             "cpu_baseline": gcpu}
    # location visits x places -> place visits (haversine <= 100 m), spatial grid vs cross-join
    nlv, side = 2_000_000, 100                                   # 3 regions x 10 000 places (default place count)
    centres = [(48.85, 2.35), (59.93, 30.33), (-33.86, 151.2)]
    gy, gx = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    plat = np.concatenate([la + gy.ravel() * 0.00135 for la, lo in centres])
    plon = np.concatenate([lo + gx.ravel() * 0.0021 for la, lo in centres])
    pid = 40 + np.arange(len(plat), dtype=np.int64)
    pcat = rng.integers(0, 20, len(plat))
    preg = np.repeat(np.arange(3), side * side)
    pick = rng.integers(0, len(pid), nlv)
    vlat = plat[pick] + rng.normal(0, 0.0007, nlv)
    vlon = plon[pick] + rng.normal(0, 0.001, nlv)
    vper = 1_000_000 + rng.integers(0, 300_000, nlv)
    vts = 1_546_300_800_000 + rng.integers(0, 7 * 24 * 3600 * 1000, nlv)
    vreg = preg[pick]
    builders.place_visits_builder(vper[:10000], vlat[:10000], vlon[:10000], vts[:10000], vreg[:10000], pid, plat, plon, pcat,
                                  preg, 7, 100.0, ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(reps):
        pv = builders.place_visits_builder(vper, vlat, vlon, vts, vreg, pid, plat, plon, pcat, preg, 7, 100.0, ctx=ctx)
    pdt = (time.perf_counter() - t0) / reps
    pcpu = None
    if rank == 0 and not args.no_cpu_baseline:
        ns = 3000
        t1 = time.perf_counter()
        oracle.build_place_visits(vper[:ns], vlat[:ns], vlon[:ns], vts[:ns], vreg[:ns], pid, plat, plon, pcat, preg, 7, 100.0)
        pcdt = (time.perf_counter() - t1) / 2                    # the wrapper runs the join twice (size, then fill)
        pcpu = {"value": ns / pcdt, "unit": "location visits/s", "cores": 1, "kind": "port",
                "sample": f"first {ns} location visits x {len(pid)} places, {pcdt:.2f}s, the reference's cross-join "
                          f"row by row (oracle/vrec_oracle.c, one thread)"}
    place_visits = {"metric": "place visits builder, location visits/s (haversine <= 100 m against the region's places)",
                    "value": nlv / pdt, "unit": "location visits/s", "ms_per_step": pdt * 1e3,
                    "config": {"workload": f"{nlv} location visits x {len(pid)} places in 3 regions -> {len(pv[0])} place visits"},
                    "cpu_baseline": pcpu}
    return {"graph": graph, "place_visits": place_visits,
            "metric": "rating vectors builder, visit rows/s (count per (person, place), rank <= 100, CSR)",
            "value": len(pe) / dt, "unit": "visits/s", "ms_per_step": dt * 1e3,
            "config": {"workload": f"{len(pe)} visit rows of {args.knn_persons} persons x {args.knn_places} places "
                                   f"(generator G2, shuffled), {len(out[2])} ratings kept"},
            "e2e": {"value": len(pe) / dt, "unit": "visits/s", "h2d_bytes_per_step": 16 * len(pe),
                    "d2h_bytes_per_step": 8 * len(out[0]) + 12 * len(out[2])},
            "cpu_baseline": cpu,
            "note": "value == e2e (host arrays in, host CSR out); the two radix sorts are cub::DeviceRadixSort"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)     # x 8 calls x 15.5 ms: a 2.5 s timed region
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--knn-persons", type=int, default=1_000_000)
    ap.add_argument("--knn-places", type=int, default=100_000)
    ap.add_argument("--knn-batch", type=int, default=18944)   # 148 SMs x 128 targets: one full wave
    ap.add_argument("--knn-calls", type=int, default=8)       # calls per step: 20 steps x 8 x 15.5 ms = 2.5 s timed
    ap.add_argument("--k-nearest", type=int, default=50)
    ap.add_argument("--max-recs", type=int, default=10)
    ap.add_argument("--sg-vertices", type=int, default=10_000_000)
    ap.add_argument("--sg-degree", type=int, default=100)
    ap.add_argument("--sg-iterations", type=int, default=20)
    ap.add_argument("--cpu-knn-targets", type=int, default=1024)
    ap.add_argument("--cpu-sg-vertices", type=int, default=1_000_000)
    ap.add_argument("--ref-knn-targets", type=int, default=512)
    ap.add_argument("--ref-sg-vertices", type=int, default=1_000_000)
    ap.add_argument("--sg-parity-degree", type=int, default=10)
    ap.add_argument("--sgb-persons", type=int, default=400_000)     # persons per region of the config-4 graphs
    ap.add_argument("--sgb-places", type=int, default=10_000)
    ap.add_argument("--sgb-unit", type=int, default=100_000)        # persons per (graph, range) unit
    ap.add_argument("--sgb-batch", type=int, default=4096)          # persons per unit per step
    ap.add_argument("--g1-persons", type=int, default=1_000_000)    # persons per region of the default data (configs 1-2)
    ap.add_argument("--no-sg", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: W < 3 warm-up steps breaks the timing rules; use only for smoke runs")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
