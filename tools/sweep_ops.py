#!/usr/bin/env python
"""Per-op Voodoo bandwidth sweep (BASELINE.json config 5; SURVEY.md section 8 d): every op of the vocabulary through its
C entry point (vdl_op_*), N in {1M, 10M, 100M, 1B} rows, timed with CUDA events on the library's stream.

bytes = inputs read + outputs written per op (int64 vectors: 8 B; ranges are virtual: 0 B), GB/s = bytes / time.
One JSON line per (op, N) on stdout.  Usage: python tools/sweep_ops.py [--sizes 1e6,1e7,1e8,1e9] [--reps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1e6,1e7,1e8,1e9")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from mplan2vdl_b200 import synth
    from mplan2vdl_b200.executor import Context

    ctx = Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=0)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]

    def col(name, n, kind, vmin, p0, width=8, stride=1, p1=0):
        spec = synth.ColumnSpec(name, width, kind, vmin, stride, p0, p1, hash(name) & 0xFFFF)
        try:
            ctx.drop_column(name)
        except Exception:
            pass
        return ctx.fill_synthetic(name, spec, n, 0x5EED, 0)

    def timed(fn, reps):
        outs = fn()                       # warm-up (pool growth, scratch)
        for v in outs:
            ctx.free(v)
        ctx.synchronize()
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            outs = fn()
            e1.record(ext)
            ctx.synchronize()
            ms.append(e0.elapsed_time(e1))
            for v in outs:
                ctx.free(v)
        return min(ms), sorted(ms)[len(ms) // 2]

    for n in [int(float(x)) for x in args.sizes.split(",")]:
        a = col("s.a", n, synth.UNIFORM, 0, 1000)
        b = col("s.b", n, synth.UNIFORM, 1, 1000)
        flag = col("s.flag", n, synth.UNIFORM, 0, 2)               # 50 % selectivity predicate
        key = col("s.key", n, synth.UNIFORM, 0, 32)                # 32 buckets
        runs = col("s.runs", n, synth.FKDENSE, 0, max(1, n // 4), p1=n)   # nondecreasing, runs of ~4
        wide = col("s.wide", n, synth.UNIFORM, 0, 1 << 38)         # unsorted 38-bit keys (Q3's composite key width)
        perm = ctx.op_partition(key, 0, 1, 32)                     # a permutation to scatter / gather by
        ctx.synchronize()
        cases = {
            "Add (binary, 2 cols)": (lambda: [ctx.op_binary("Add", a, b)], 24 * n),
            "Greater (col vs RangeV const)": (lambda: [ctx.op_binary("Greater", a, ctx_const)], 16 * n),
            "FoldSelect (50%)": (lambda: [ctx.op_fold_select(flag)], 8 * n + 8 * (n // 2)),
            "Gather (by permutation)": (lambda: [ctx.op_gather(a, perm)], 24 * n),
            "Gather (by pos_: sequential)": (lambda: [ctx.op_gather(a, ctx_pos)], 16 * n),
            "Scatter (by permutation)": (lambda: [ctx.op_scatter(a, perm, n)], 24 * n),
            "Partition (32 buckets, unsorted)": (lambda: [ctx.op_partition(key, 0, 1, 32)], 16 * n),
            "Partition (unsorted 38-bit keys: LSD radix, 8 bits per pass)": (lambda: [ctx.op_partition(wide, 0, 1, (1 << 38) - 1)], 16 * n),
            "Partition (sorted keys -> identity)": (lambda: [ctx.op_partition(runs, 0, 1, max(1, n // 4))], 8 * n),
            "FoldSum (runs of ~4)": (lambda: [ctx.op_fold("FoldSum", runs, a)], 16 * n + 8 * (n // 4)),
            "FoldSum (one run)": (lambda: [ctx.op_fold("FoldSum", ctx_zero, a)], 8 * n),
        }
        ctx_const = ctx.op_range(500, 0, n)
        ctx_pos = ctx.op_range(0, 1, n)
        ctx_zero = ctx.op_range(0, 0, n)
        for name, (fn, nbytes) in cases.items():
            best, med = timed(fn, args.reps)
            print(json.dumps({"op": name, "rows": n, "ms_best": round(best, 4), "ms_median": round(med, 4), "algorithmic_bytes": nbytes,
                              "gbs": round(nbytes / best / 1e6, 1), "frac_of_measured_hbm_peak": round(nbytes / best / 1e6 / peak, 3)}), flush=True)
        for v in (perm, ctx_const, ctx_pos, ctx_zero):
            ctx.free(v)
        for c in ("s.a", "s.b", "s.flag", "s.key", "s.runs", "s.wide"):
            ctx.drop_column(c)
    ctx.close()


if __name__ == "__main__":
    main()
