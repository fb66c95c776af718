"""Synthetic TPC-H-shaped tables resident in HBM, sharded the way the executor scales out:
the fact table (lineitem) by contiguous row range, dimension tables replicated (SURVEY.md section 8 e)."""
from __future__ import annotations

import os
import re

from . import synth
from .meta import Catalog

PLANS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plans")
FACT_TABLE = "lineitem"
SHARD_ALIGN = 4096      # shard boundaries are multiples of the largest scan tile so every shard starts tile-aligned


def plan_text(name: str) -> str:
    with open(os.path.join(PLANS_DIR, name)) as f:
        return f.read()


def plan_columns(text: str) -> list:
    """Qualified names of the columns a plan Loads (Vdl.hs:419-420 `Load,<table>.<col>`)."""
    seen = []
    for m in re.finditer(r"^\d+,Load,([\w.]+)", text, re.M):
        if m.group(1) not in seen:
            seen.append(m.group(1))
    return seen


def shard_range(rows: int, rank: int, world: int) -> tuple:
    """(first global row, row count) of `rank`'s contiguous shard of a `rows`-row fact table."""
    per = -(-rows // world)
    per = -(-per // SHARD_ALIGN) * SHARD_ALIGN
    start = min(rows, rank * per)
    return start, max(0, min(rows, start + per) - start)


def load_synthetic(ctx, cat: Catalog, columns, sf: float, rank: int = 0, world: int = 1, rows_override: dict | None = None) -> dict:
    """Generate the named columns in place on ctx's GPU.  Returns {"row_base": int, "rows": {table: rows on this GPU}}."""
    seed = synth.seed_for(sf)
    rows_here, row_base = {}, 0
    heaps = {q[:-len(".heap")] for q in columns if synth.is_heap(q)}     # string columns whose heap a Like reads
    for q in columns:
        table = q.split(".")[0]
        if synth.is_heap(q):                               # the bytes of the heap: host-built (small), uploaded once
            heap, _ = synth.string_heap(cat, q[:-len(".heap")])
            ctx.upload_column(q, heap)
            continue
        total = (rows_override or {}).get(table, synth.table_rows(cat, table, sf))
        if table == FACT_TABLE:
            start, n = shard_range(total, rank, world)
            row_base = start
        else:
            start, n = 0, total
        if q in heaps:                                     # ... and its offsets column: pool entries drawn per global row
            ctx.upload_column(q, synth.string_offsets(cat, q, n, start, seed))
            rows_here[table] = n
            continue
        spec = synth.column_spec(cat, q, sf)
        if table in (rows_override or {}) and spec.kind == synth.FKDENSE:
            spec = synth.ColumnSpec(spec.name, spec.width, spec.kind, spec.vmin, spec.stride, spec.p0, total, spec.stream)
        ctx.fill_synthetic(q, spec, n, seed, start)
        rows_here[table] = n
    return {"row_base": row_base, "rows": rows_here}
