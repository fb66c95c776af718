#!/usr/bin/env python
"""BASELINE config 5: every plan of the reference's tests/tpch10noorder suite that the restated translator prints
(plans/*.vdl, 15 of 22; the other 7 stop where the reference's own `error` calls do) over synthetic TPC-H columns at one
scale factor, on one GPU or row-range sharded over the ranks of a torchrun launch.  One JSON line per plan: ms per step
(CUDA events on the library's stream, max over ranks), lineitem rows/s, what the planner made of it (fused scans, probe
passes, map clusters, launches) and, with --parity, a bit-exact check against the CPU oracle on host-generated columns
(small scale factors only: the oracle materialises every intermediate).  Not a bench.py line: a suite table for DESIGN.md."""
import argparse
import glob
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10, help="the catalogue (bounds.csv, key widths) is the reference's SF10 one: larger scale "
                    "factors are valid only for plans whose keys do not depend on SF-specific bounds (Q1, Q5, Q6, Q12, Q19)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--plans", default="")
    ap.add_argument("--parity", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import __graft_entry__ as ge
    ge.build()
    from mplan2vdl_b200 import synth, tpch
    from mplan2vdl_b200.dist import ShardedPlan
    from mplan2vdl_b200.executor import Context
    from mplan2vdl_b200.meta import builtin_catalog
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cat = builtin_catalog()
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "plans", "q*.vdl")))
    if args.plans:
        names = [n for n in names if n in args.plans.split(",")]
    rows_total = synth.table_rows(cat, "lineitem", args.sf)
    for q in names:
        line = {"plan": q, "sf": args.sf, "n_gpus": world}
        ctx = None
        try:
            text = tpch.plan_text(q + ".vdl")
            cols = tpch.plan_columns(text)
            ctx = Context(local)
            info = tpch.load_synthetic(ctx, cat, cols, args.sf, rank=rank, world=world)
            plan = ctx.plan(text)
            st = plan.stats()
            # N > 1: only the plan shapes whose sharded execution is covered by tests -- partial tables that combine (fused scans,
            # probe folds) or an emit plan with a mergeable tail.  (The all-gather of survivors of several emit groups is
            # exercised on emulated ranks only; a data-dependent error on ONE rank would leave the others in a collective.)
            shardable = world == 1 or ((st["fused_scans"] or plan.num_partials) and plan.num_emits == 0) or (plan.num_emits > 0 and plan.tail_info() is not None)
            if not shardable:
                raise RuntimeError("not run sharded: no combinable partial table and no mergeable tail (runs on one GPU)")
            sp = ShardedPlan(ctx, plan, rank, world, info["row_base"])
            ext = torch.cuda.ExternalStream(ctx.stream, device=local)
            for _ in range(args.warmup):
                sp.step(copy=False, fetch=False)
            ctx.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ctx.launch_count
            t0 = time.perf_counter()
            e0.record(ext)
            for _ in range(args.steps):
                sp.step(copy=False, fetch=False)
            e1.record(ext)
            ctx.synchronize()
            wall = (time.perf_counter() - t0) * 1e3 / args.steps
            ms = e0.elapsed_time(e1) / args.steps
            if world > 1:
                t = torch.tensor([ms, wall], dtype=torch.float64, device=f"cuda:{local}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms, wall = float(t[0]), float(t[1])
            res = sp.global_result() if world > 1 else {k: np.array(v, copy=True) for k, v in plan.outputs(False).items()}
            st = plan.stats()
            line.update(ms_per_step=round(ms, 4), wall_ms_per_step=round(wall, 4), lineitem_rows_per_s=rows_total / (ms / 1e3),
                        launches_per_step=(ctx.launch_count - l0) // args.steps, fused_scans=st["fused_scans"],
                        probe_fold_groups=st.get("probe_folds"), probe_emit_groups=st.get("probe_emits"),
                        map_clusters=st.get("map_clusters"), outputs=len(res), result_rows=int(len(next(iter(res.values())))),
                        combine=("sharded tail" if sp.tail_mode else ("peer memory" if sp.peer_mode else "all-gather")) if world > 1 else None)
            if args.parity and rank == 0 and world == 1:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from util import host_columns, run_oracle
                rows = {t: synth.table_rows(cat, t, args.sf) for t in cat.tables}
                want = run_oracle(text, host_columns(cat, cols, rows, sf=args.sf))
                line["parity"] = "exact" if list(want) == list(res) and all(np.array_equal(want[k], res[k]) for k in want) else "MISMATCH"
            plan.close()
        except Exception as e:                                  # a plan that cannot run is a row of the table too
            line["error"] = f"{type(e).__name__}: {str(e)[:200]}"
        finally:
            if ctx is not None:
                ctx.close()
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
