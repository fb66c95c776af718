#!/usr/bin/env python
"""Known answers for the six TPC-H plans in scope, committed under tests/golden/.

The reference holds no result vectors (it ships no executor; tests/Tests.hs:17-18), so these come from the one evaluator
in this repo that shares nothing with the plan interpreters: oracle/sqlref.py, the SQL text of each query restated in
numpy over the base columns.  Inputs: the synthetic recipe at SF 0.01 (mplan2vdl_b200/synth.py, seed_for(0.01)); Q19
with its string-coded columns redrawn (tests/util.q19_columns).  The tests compare BOTH the CPU oracle and the CUDA
path with these files, so a change that moves the two in lockstep still shows.

    python tools/make_golden_results.py        # rewrites tests/golden/tpch_sf0.01_answers.json
"""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mplan2vdl_b200 import synth, tpch            # noqa: E402
from mplan2vdl_b200.meta import builtin_catalog    # noqa: E402
from oracle import sqlref                          # noqa: E402
import util                                        # noqa: E402

SF = 0.01


def columns_for(cat, q):
    if q == "q19":
        return util.q19_columns(cat, sf=SF)[1]
    text = util.plan_text(q + ".vdl")
    rows = {t: synth.table_rows(cat, t, SF) for t in cat.tables}
    return util.host_columns(cat, tpch.plan_columns(text), rows, sf=SF)


def main():
    cat = builtin_catalog()
    refs = {"q01": sqlref.q1, "q03": sqlref.q3, "q05": sqlref.q5, "q06": sqlref.q6, "q12": sqlref.q12,
            "q19": lambda c: sqlref.q19(c, cat.dictionary)}
    out = {"_comment": "answers of oracle/sqlref.py (numpy restatement of the SQL) on the synthetic SF 0.01 tables; "
                       "written by tools/make_golden_results.py", "sf": SF, "seed": synth.seed_for(SF), "answers": {}}
    for q, ref in refs.items():
        res = ref(columns_for(cat, q))
        out["answers"][q] = {k: [int(x) for x in v] for k, v in res.items()}
        print(q, {k: len(v) for k, v in res.items()})
    dst = os.path.join(ROOT, "tests", "golden", "tpch_sf0.01_answers.json")
    with open(dst, "w") as f:
        json.dump(out, f, separators=(",", ":"))
        f.write("\n")
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
