#!/usr/bin/env python
"""Rewrite profiles/traffic.json from the ncu summaries of a round (profiles/<prefix>_<workload>_dominant_kernel.txt, written
by tools/ncu_summary.py from `ncu --set full` captures): per workload the DRAM bytes of ONE launch of the dominant kernel,
the kernel's name and duration under ncu, and -- so that a stale figure can never ride along with a changed kernel -- the
hashes of the kernel sources the capture was taken from: of everything, and per kernel family (the files that define the
fused scan's resp. the probe's device code, mplan2vdl_b200.build.KERNEL_FAMILIES).  bench.py reports `roofline.traffic` only
while the hash of the workload's kernel family still equals the recorded one.

    python tools/update_traffic.py r02a          # after tools/gpu/profile_round.sh r02a + tools/ncu_summary.py
"""
import glob
import json
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from mplan2vdl_b200.build import source_hash  # noqa: E402

prefix = sys.argv[1]
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel per workload, from the ncu --set full "
                   f"captures summarised in profiles/{prefix}_*_dominant_kernel.txt; written by tools/update_traffic.py; bench.py reports "
                   "roofline.traffic only while the kernel sources still hash to source_hash",
       "source_hash": source_hash(), "family_hashes": {f: source_hash(f) for f in ("scan", "probe")}, "prefix": prefix, "workloads": {}}
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"{prefix}_*_dominant_kernel.txt"))):
    tag = os.path.basename(path)[len(prefix) + 1:-len("_dominant_kernel.txt")]
    text = open(path).read()
    m = re.search(r"traffic \(dram read \+ write\) per launch: (\d+) bytes", text)
    k = re.search(r"^kernel: (\S+)", text, re.M)
    t = re.search(r"gpu__time_duration.sum\s+([0-9.]+) (\w+)", text)
    if not m:
        continue
    us = float(t.group(1)) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(t.group(2), 1) if t else None
    out["workloads"][tag] = {"dram_bytes": int(m.group(1)), "kernel": k.group(1) if k else None, "ncu_duration_us": us,
                             "family": "scan" if k and "scan" in k.group(1) else "probe"}
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
    f.write("\n")
print(json.dumps(out, indent=1))
