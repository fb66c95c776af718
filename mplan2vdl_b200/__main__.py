"""`mplan2vdl ... | python -m mplan2vdl_b200 [--sf 1] [--csv]` -- the place of the HTTP Voodoo server in the
reference pipeline (eval_query.sh:18-26): read the Voodoo program mplan2vdl printed (stdin or a file), execute it on
GPU 0 over synthetic TPC-H-shaped tables generated to the reference's bounds metadata, and print the server's JSON
(resolve.py:8-32) -- or, with --csv, what `resolve.py dictionary.csv` would print from it."""
from __future__ import annotations

import argparse
import re
import sys
import time


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m mplan2vdl_b200")
    ap.add_argument("plan", nargs="?", default="-", help="Voodoo program text (VdlFormat); '-' = stdin")
    ap.add_argument("--sf", type=float, default=1.0, help="scale factor of the synthetic tables")
    ap.add_argument("--csv", action="store_true", help="decode dictionary-coded outputs and print CSV (resolve.py)")
    ap.add_argument("--no-fuse", action="store_true", help="op-at-a-time execution only")
    ap.add_argument("--explain", action="store_true", help="print what the planner makes of the program (JSON) and stop: needs no GPU")
    args = ap.parse_args(argv)
    text = sys.stdin.read() if args.plan == "-" else open(args.plan).read()
    text = re.sub(r" ;;.*", "", text)          # the pipeline strips the --metadata suffix with sed (eval_query.sh:20)
    if args.explain:
        import json
        from .executor import explain
        sys.stdout.write(json.dumps(explain(text, fuse=not args.no_fuse), indent=1) + "\n")
        return 0

    from . import resolve, tpch
    from .executor import Context
    from .meta import builtin_catalog
    cat = builtin_catalog()
    ctx = Context(0)
    tpch.load_synthetic(ctx, cat, tpch.plan_columns(text), args.sf)
    plan = ctx.plan(text, fuse=not args.no_fuse)
    t0 = time.perf_counter()
    out = plan.run()
    us = 1e6 * (time.perf_counter() - t0)
    doc = resolve.to_server_json(out, {"timeInMicrosecondsForPlan": us})
    sys.stdout.write(resolve.to_csv(resolve.resolve(doc, cat)) if args.csv else doc + "\n")
    plan.close()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
