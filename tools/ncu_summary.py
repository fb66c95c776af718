#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the text summary committed under profiles/:
headline metrics, stall breakdown, instruction mix by region.  Usage: ncu_summary.py rep.ncu-rep rows > out.txt"""
import csv
import io
import subprocess
import sys

rep, rows = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = r[0], r[1], r[-1]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
print(f"report: {rep}")
print(f"kernel: {m['Kernel Name'][0]}   grid {m.get('Grid Size', ('?',))[0]} block {m.get('Block Size', ('?',))[0]}")
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        print(f"  {k:78s} {m[k][0]} {m[k][1]}")
try:
    rd, wr = float(m["dram__bytes_read.sum"][0]), float(m["dram__bytes_write.sum"][0])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    tb = rd * scale.get(m["dram__bytes_read.sum"][1], 1) + wr * scale.get(m["dram__bytes_write.sum"][1], 1)
    print(f"  traffic (dram read + write) per launch: {tb:.0f} bytes")
except Exception as e:
    print("  traffic: n/a", e)
print("stalls (warp cycles per issued instruction):")
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = float(m[h][0] or 0)
        if v >= 0.05:
            print(f"  {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
d = list(csv.reader(io.StringIO(src)))
h2, data = d[1], d[2:]
iex, ith, isrc = h2.index("Instructions Executed"), h2.index("Thread Instructions Executed"), h2.index("Source")
tot, tth = sum(int(x[iex]) for x in data), sum(int(x[ith]) for x in data)
print(f"SASS: {len(data)} static instructions, {tot} warp-instructions, {tth} thread-instructions executed")
if rows:
    print(f"  thread-instructions per row: {tth / rows:.1f}")
ops = {}
for x in data:
    op = x[isrc].split()[1 if x[isrc].lstrip().startswith("@") else 0].split(".")[0] if x[isrc].strip() else "?"
    ops[op] = ops.get(op, 0) + int(x[iex])
print("  top opcodes by executed warp-instructions: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
tma = sum(v for k, v in ops.items() if k in ("UBLKCP", "UTMALDG"))
print(f"  TMA bulk copies (UBLKCP) executed: {tma}; mbarrier ops (SYNCS): {ops.get('SYNCS', 0)}")
