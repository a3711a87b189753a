#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an `ncu --set full` report: joins the
report's SASS page with `nvdisasm -g` line info of the library that was profiled (ncu on the GPU box cannot import
the sources).   python tools/ncu_lines.py <report.ncu-rep> <kernel substring> [libvrec.so] [extra function substring]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kernel = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "locations-recommender_b200", "libvrec.so")
extra = sys.argv[4:]                   # device functions called by the kernel that ncu lists as separate segments
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
secs = []
for i, l in enumerate(sass):
    if l.startswith(".text.") and any(k in l for k in [kernel] + extra):
        cur, instr = None, []
        for l2 in sass[i + 1:]:
            if l2.startswith("//--------------------- .text"):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l2)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l2):
                instr.append(cur)
        secs.append(instr)
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                                      capture_output=True, text=True).stdout.splitlines()))
segs, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Address":
        hdr, cur = r, []
        segs.append(cur)
    elif cur is not None and len(r) > 10:
        cur.append(r)
ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
by, bys, byt, tot, used = collections.Counter(), collections.Counter(), collections.Counter(), 0, set()
for seg in segs:
    match = [k for k, ins in enumerate(secs) if len(ins) == len(seg) and k not in used]
    if not match:
        continue
    used.add(match[0])
    for line, r in zip(secs[match[0]], seg):
        c, s_, t = int(r[ia]), int(r[isamp]), int(r[it])
        by[line] += c
        bys[line] += s_
        byt[line] += t
        tot += c
ts = max(1, sum(bys.values()))
print(f"{kernel}: {tot} warp instructions, {ts} stall samples; lines by instructions executed")
for line, c in by.most_common(30):
    print(f"  {line[0]}:{line[1]:<5} {c / tot * 100:5.1f}% instr  {bys[line] / ts * 100:5.1f}% samples  "
          f"{byt[line] / max(1, c):4.1f} of 32 lanes active")
if os.environ.get("NCU_LINES_BY_SAMPLES"):          # second list: where the warps wait
    print("lines by stall samples")
    for line, s_ in bys.most_common(int(os.environ["NCU_LINES_BY_SAMPLES"])):
        c = by[line]
        print(f"  {line[0]}:{line[1]:<5} {c / tot * 100:5.1f}% instr  {s_ / ts * 100:5.1f}% samples  "
              f"{byt[line] / max(1, c):4.1f} of 32 lanes active")
