#!/usr/bin/env python
"""Condenses `ncu --set full` captures into profiles/r2_ncu_summary.json, the file bench.py reads `roofline.traffic`
from (a profiler cannot wrap a timed region, so the bench line cites this file instead of pasting numbers).
  python tools/ncu_summary.py <kernel-name> <report.ncu-rep> "<workload / command>" [more triples ...]
Also writes the raw metric page of each report to profiles/<report>_raw.csv."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_ncu_summary.json")


def main():
    args = sys.argv[1:]
    assert args and len(args) % 3 == 0, __doc__
    try:
        summary = json.load(open(OUT))
    except Exception:
        summary = {}
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    for kernel, rep, workload in zip(args[0::3], args[1::3], args[2::3]):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        name = os.path.splitext(os.path.basename(rep))[0] + "_raw.csv"
        with open(os.path.join(ROOT, "profiles", name), "w") as f:
            f.write(raw)
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        picked = [r for r in rows[2:] if kernel in r[col["Kernel Name"]]]
        assert picked, f"{kernel} not in {rep}"
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
                 "s": 1e3, "%": 1.0, "": 1.0}

        def get(r, key):
            i = col[key]
            return float(r[i]) * scale.get(units[i], 1.0)

        launches = []
        for r in picked:
            launches.append({"duration_ms": get(r, "gpu__time_duration.sum"),
                             "dram_bytes": get(r, "dram__bytes_read.sum") + get(r, "dram__bytes_write.sum"),
                             "tensor_pipe_active_pct": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                             "l1tex_throughput_pct": get(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"),
                             "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct"),
                             "registers": get(r, "launch__registers_per_thread")})
        summary[kernel] = {
            "dram_bytes": sum(x["dram_bytes"] for x in launches),      # all launches of one iteration / one call
            "duration_ms": sum(x["duration_ms"] for x in launches),
            "tensor_pipe_active_pct": launches[0]["tensor_pipe_active_pct"],
            "launches": launches,
            "source": f"profiles/{name}: ncu --set full --clock-control none, {workload}, build {head}",
        }
    with open(OUT, "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "launches"} for k, v in summary.items()}, indent=1))


if __name__ == "__main__":
    main()
