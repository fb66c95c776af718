#!/usr/bin/env python
"""Derive mplan2vdl_b200/catalog/tpch10noorder.json from the reference's metadata fixtures.

Run in the build container (needs /root/reference): python tools/make_catalog.py
The JSON is committed because the GPU box has no /root/reference.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mplan2vdl_b200.meta import load_metadata  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tests/tpch10noorder"
dst = os.path.join(os.path.dirname(__file__), "..", "mplan2vdl_b200", "catalog", "tpch10noorder.json")
cat = load_metadata(src)
with open(dst, "w") as f:
    json.dump(cat.to_json(), f, indent=1, sort_keys=True)
print(f"wrote {dst}: {len(cat.tables)} tables, {sum(len(t.columns) for t in cat.tables.values())} columns, "
      f"{sum(len(v) for v in cat.dictionary.values())} dictionary entries")
