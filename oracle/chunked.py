"""The CPU oracle over a fact table that does not fit the host at once: SURVEY.md section 8 (d), "at SF100 the oracle
should generate/consume columns in chunks".

TEST INFRASTRUCTURE ONLY (see vdl_oracle.c): used by tests/ and by bench.py's in-run parity check.

The plan is cut at its Folds.  Per row-range chunk of the fact table the unmodified oracle interpreter runs the plan's
statements with every Fold result (and, per Fold, the key of each run: FoldMin(groups, groups)) added as outputs; the
per-chunk runs are merged BY KEY in chunk (= row) order -- FoldSum adds mod 2^64, FoldMin / FoldMax combine, FoldChoose
keeps the value of the earliest chunk that has the key (first of the run, G6); a key no chunk produced yields no run
(G14) -- which is what the whole-table evaluation computes, because the groups vector of every Fold the translator emits
is the Partition-sorted key (Vlite.hs:1057-1060, 1172: one run per key, ascending) or a constant (one run).  The
statements ABOVE the Folds (AVG's Divide, Vlite.hs:1038-1041) are then run once, again by the oracle interpreter, over
the merged vectors bound as columns.  Nothing is re-implemented here except the merge rule.
"""
from __future__ import annotations

import re

import numpy as np

from .oracle import Oracle

FOLDS = ("FoldSum", "FoldMin", "FoldMax", "FoldChoose", "FoldCount")
_REF = re.compile(r"Id (\d+)")


class NotChunkable(ValueError):
    pass


def _parse(text: str):
    stmts = []
    for line in text.splitlines():
        line = line.split(" ;; ")[0].rstrip()
        if not line:
            continue
        f = line.split(",")
        stmts.append((int(f[0]), f[1], f[2:]))
    assert [s[0] for s in stmts] == list(range(1, len(stmts) + 1))
    return stmts


def _refs(fields):
    return [int(m.group(1)) for x in fields for m in [_REF.fullmatch(x)] if m]


def split_plan(text: str, fact_table: str):
    """-> (per-chunk plan text, tail plan text, [(fold id, op, groups id)], {groups id: output name of its run keys})."""
    stmts = _parse(text)
    n = len(stmts)
    fact = [False] * (n + 1)          # depends on a fact-table column
    cut = []                          # first-level Folds over fact rows
    below_fold = [False] * (n + 1)    # depends on a cut Fold
    for sid, op, fields in stmts:
        r = _refs(fields)
        if op == "Load":
            fact[sid] = fields[0].split(".")[0] == fact_table
        else:
            fact[sid] = any(fact[a] for a in r)
        below_fold[sid] = any(below_fold[a] for a in r)
        if op in FOLDS and fact[sid] and not below_fold[sid]:
            g, d = r
            cut.append((sid, op, g))
            below_fold[sid] = True
            fact[sid] = False         # from here on the value is per key, not per fact row
        elif op == "MaterializeCompact" and fact[sid]:
            raise NotChunkable(f"statement {sid}: an output depends on fact rows without a Fold in between")
    if not cut:
        raise NotChunkable("no Fold over the fact table")
    # per-chunk plan: everything up to and including the cut Folds (ids must stay consecutive: renumber), outputs = the
    # cut Folds + their run keys
    keep = [s for s in stmts if s[1] != "MaterializeCompact" and (not below_fold[s[0]] or s[0] in {c[0] for c in cut})]
    ren = {}
    lines = []
    for sid, op, fields in keep:
        ren[sid] = len(lines) + 1
        lines.append(f"{ren[sid]},{op}," + ",".join(_REF.sub(lambda m: f"Id {ren[int(m.group(1))]}", x) if _REF.fullmatch(x) else x for x in fields))
    key_out = {}

    def add_output(node_new_id, name):
        lines.append(f"{len(lines) + 1},Project,{name},Id {node_new_id},val")
        lines.append(f"{len(lines) + 1},MaterializeCompact,Id {len(lines)}")

    for sid, op, g in cut:
        add_output(ren[sid], f"__fold_{sid}")
        if g not in key_out:
            lines.append(f"{len(lines) + 1},FoldMin,val,Id {ren[g]},val,Id {ren[g]},val")
            key_out[g] = f"__keys_{g}"
            add_output(len(lines), key_out[g])
    chunk_plan = "\n".join(lines) + "\n"
    # tail plan: the outputs' cones above the cut, Folds replaced by Loads of the merged vectors
    cutids = {c[0] for c in cut}
    need = set()
    stack = [sid for sid, op, _ in stmts if op == "MaterializeCompact"]
    by_id = {s[0]: s for s in stmts}
    while stack:
        sid = stack.pop()
        if sid in need:
            continue
        need.add(sid)
        if sid not in cutids:
            stack.extend(_refs(by_id[sid][2]))
    ren, lines = {}, []
    for sid, op, fields in stmts:
        if sid not in need:
            continue
        ren[sid] = len(lines) + 1
        if sid in cutids:
            lines.append(f"{ren[sid]},Load,__fold.f{sid}")
        else:
            lines.append(f"{ren[sid]},{op}," + ",".join(_REF.sub(lambda m: f"Id {ren[int(m.group(1))]}", x) if _REF.fullmatch(x) else x for x in fields))
    tail_plan = "\n".join(lines) + "\n"
    return chunk_plan, tail_plan, cut, key_out


class ChunkedOracle:
    """Feed the fact table chunk by chunk (in row order), then `finish()`."""

    def __init__(self, plan_text: str, fact_table: str = "lineitem", threads: int = 0):
        self.chunk_plan, self.tail_plan, self.cut, self.key_out = split_plan(plan_text, fact_table)
        self.fact_table, self.threads = fact_table, threads
        self.dims = {}
        self.keys = {g: np.zeros(0, np.int64) for g in self.key_out}          # merged ascending keys per groups node
        self.acc = {sid: np.zeros(0, np.int64) for sid, _, _ in self.cut}
        self.seconds = 0.0
        self.rows = 0

    def bind_dimension(self, name: str, arr: np.ndarray):
        self.dims[name] = arr

    def add_chunk(self, fact_cols: dict):
        """fact_cols: {qualified name: array} for one contiguous row range of the fact table, ranges fed in ascending order.
        FK index columns hold GLOBAL positions into the (fully bound) dimension tables, so nothing shifts."""
        o = Oracle()
        for k, v in self.dims.items():
            o.bind(k, v)
        for k, v in fact_cols.items():
            o.bind(k, v)
        out = o.run(self.chunk_plan, self.threads)
        self.seconds += o.seconds
        self.rows += len(next(iter(fact_cols.values())))
        merged_idx = {}
        for g, name in self.key_out.items():
            kc = out[name]
            if len(kc) > 1 and not np.all(kc[1:] > kc[:-1]):
                raise NotChunkable(f"groups of statement {g} are not one ascending run per key")
            old = self.keys[g]
            allk = np.union1d(old, kc)
            merged_idx[g] = (allk, np.searchsorted(allk, old), np.searchsorted(allk, kc), len(old))
            self.keys[g] = allk
        with np.errstate(over="ignore"):
            for sid, op, g in self.cut:
                allk, iold, inew, nold = merged_idx[g]
                vals = out[f"__fold_{sid}"]
                seen = np.zeros(len(allk), bool)
                seen[iold] = True
                acc = np.zeros(len(allk), np.int64)
                acc[iold] = self.acc[sid]
                fresh = ~seen[inew]
                if op in ("FoldSum", "FoldCount"):
                    acc[inew] = acc[inew] + vals                      # wraps mod 2^64 (G5)
                elif op == "FoldMin":
                    acc[inew] = np.where(fresh, vals, np.minimum(acc[inew], vals))
                elif op == "FoldMax":
                    acc[inew] = np.where(fresh, vals, np.maximum(acc[inew], vals))
                else:                                                 # FoldChoose: first of the run = earliest chunk
                    acc[inew] = np.where(fresh, vals, acc[inew])
                self.acc[sid] = acc

    def finish(self) -> dict:
        o = Oracle()
        for k, v in self.dims.items():
            o.bind(k, v)
        for sid, _, _ in self.cut:
            o.bind(f"__fold.f{sid}", np.ascontiguousarray(self.acc[sid]))
        out = o.run(self.tail_plan, self.threads)
        self.seconds += o.seconds
        return out


def run_chunked(plan_text: str, cat, sf: float, chunk_rows: int = 50_000_000, threads: int = 0, rows_override: dict | None = None,
                fact_table: str = "lineitem") -> tuple:
    """The plan over the synthetic tables of scale factor `sf` (mplan2vdl_b200/synth.py recipe), the fact table generated
    and consumed `chunk_rows` at a time.  -> ({output: array}, ChunkedOracle)."""
    from mplan2vdl_b200 import synth, tpch
    from .oracle import gen_column
    names = tpch.plan_columns(plan_text)
    seed = synth.seed_for(sf)
    co = ChunkedOracle(plan_text, fact_table, threads)
    total = {}
    for q in names:
        t = q.split(".")[0]
        total[t] = (rows_override or {}).get(t, synth.table_rows(cat, t, sf))
    specs = {}
    for q in names:
        spec = synth.column_spec(cat, q, sf)
        t = q.split(".")[0]
        if t in (rows_override or {}) and spec.kind == synth.FKDENSE:
            spec = synth.ColumnSpec(spec.name, spec.width, spec.kind, spec.vmin, spec.stride, spec.p0, total[t], spec.stream)
        specs[q] = spec
        if t != fact_table:
            co.bind_dimension(q, gen_column(spec, total[t], 0, seed, threads))
    fact_names = [q for q in names if q.split(".")[0] == fact_table]
    n = total[fact_table]
    for start in range(0, max(n, 1), chunk_rows):
        m = min(chunk_rows, n - start)
        co.add_chunk({q: gen_column(specs[q], m, start, seed, threads) for q in fact_names})
    return co.finish(), co
