#!/usr/bin/env python
"""Regenerate plans/*.vdl with the Vlite/Vdl restatement (mplan2vdl_b200/vlite.py) from the hand-built relational IRs
in mplan2vdl_b200/tpch_queries.py.  q06.vdl must come out identical to the checked-in golden plan."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mplan2vdl_b200 import tpch_queries, vlite  # noqa: E402
from mplan2vdl_b200.meta import builtin_catalog  # noqa: E402

cat = builtin_catalog()
for q in sys.argv[1:] or ["q03", "q05"]:
    text = vlite.translate(cat, tpch_queries.QUERIES[q](cat))
    path = os.path.join(os.path.dirname(__file__), "..", "plans", f"{q}.vdl")
    with open(path, "w") as f:
        f.write(text)
    print(f"{path}: {len(text.splitlines())} statements")
