#!/usr/bin/env python
"""Ad-hoc sweep of extra fuzz seeds on a GPU box: CUDA path (fused; and with the probe pass off) against the CPU oracle.
usage: python tools/gpu/fuzz_sweep.py FIRST LAST [sf]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np                                   # noqa: E402
import fuzz_plans                                    # noqa: E402
from mplan2vdl_b200 import synth, tpch, vlite        # noqa: E402
from mplan2vdl_b200.executor import Context          # noqa: E402
from mplan2vdl_b200.meta import builtin_catalog      # noqa: E402
from util import host_columns, run_gpu, run_oracle   # noqa: E402

first, last = int(sys.argv[1]), int(sys.argv[2])
sf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
cat = builtin_catalog()
ctx = Context(0)
bad = n = 0
for seed in range(first, last):
    for kind in ("single", "join"):
        q = fuzz_plans.single_table(seed) if kind == "single" else fuzz_plans.join_query(seed, cat)
        text = vlite.translate(cat, q)
        rows = {t: synth.table_rows(cat, t, sf) for t in cat.tables}
        cols = host_columns(cat, tpch.plan_columns(text), rows, sf=sf)
        want = run_oracle(text, cols)
        for mode in ("fused", "noprobe"):
            os.environ.pop("VDL_NO_PROBE", None)
            if mode == "noprobe":
                os.environ["VDL_NO_PROBE"] = "1"
            n += 1
            try:
                got, _ = run_gpu(text, cols, fuse=True, ctx=ctx)
                assert list(got) == list(want)
                for k in want:
                    np.testing.assert_array_equal(got[k], want[k], err_msg=k)
            except Exception as e:
                bad += 1
                print(f"MISMATCH {kind} seed {seed} {mode}: {type(e).__name__} {str(e)[:600]}".replace("\n", " "), flush=True)
                ctx.close()                      # the failed run left its columns registered: start clean
                ctx = Context(0)
os.environ.pop("VDL_NO_PROBE", None)
print(f"{n} runs, {bad} failures")
