#!/usr/bin/env python
"""Group the SASS instructions of an .ncu-rep by execution count: shows which code regions the warp-instructions go to.
Usage: ncu_buckets.py rep.ncu-rep [min_share]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iex]) for r in data)
tots = sum(int(r[ismp]) for r in data)
# contiguous regions of (roughly) equal execution count
regions = []
for n, r in enumerate(data):
    e = int(r[iex]); sm = int(r[ismp])
    if regions and abs(e - regions[-1]['e']) <= 0.02 * max(e, regions[-1]['e'], 1):
        g = regions[-1]; g['n'] += 1; g['sum'] += e; g['smp'] += sm; g['last'] = n
    else:
        regions.append({'e': e, 'n': 1, 'sum': e, 'smp': sm, 'first': n, 'last': n})
print(f"total warp-instr {tot}, samples {tots}")
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
for g in regions:
    if g['sum'] > thr * tot:
        ops = {}
        for r in data[g['first']:g['last'] + 1]:
            t = r[isrc].split()
            op = (t[1] if t and t[0].startswith('@') else t[0] if t else '?').split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        top = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
        print(f"  sass[{g['first']:5d}..{g['last']:5d}] {g['n']:4d} instr x {g['e']:>10d} = {100*g['sum']/tot:5.1f}% of instr, {100*g['smp']/max(tots,1):5.1f}% of samples   {top}")
