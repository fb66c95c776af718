#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + the hottest SASS instructions (needs ncu on PATH)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for i, h in enumerate(hdr):
    if h in want or 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.2:
        print(f"{h:90s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iex]) for r in data)
print('total warp instr', tot, 'static instr', len(data))
for n, r in enumerate(data):
    e = int(r[iex])
    if e > tot * thr:
        print(n, r[ia][-5:], e, r[ismp], r[isrc][:100])
