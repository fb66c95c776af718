"""Several ranks driven from ONE process through the C ABI alone (vdl_comm_*; no torch, no IPC handles): what a single
Haskell host would bind (SURVEY.md section 8 b/e).  Two or three ranks are emulated on GPU 0 (the same device listed
several times); every rank must end with the global answer of the whole table, bit-exact against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from mplan2vdl_b200 import lib as L_, synth, tpch
from util import Q1_COLS, Q6_COLS, assert_same, host_columns, plan_text, run_oracle

pytestmark = pytest.mark.gpu


def outputs(L, plan):
    out = {}
    for i in range(L.vdl_plan_num_outputs(plan)):
        name, data, n = C.c_char_p(), C.POINTER(C.c_int64)(), C.c_int64()
        assert L.vdl_plan_output(plan, i, C.byref(name), C.byref(data), C.byref(n)) == 0
        out[name.value.decode()] = np.ctypeslib.as_array(data, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int64)
    return out


def upload(L, ctx, name, arr):
    v = C.c_int32()
    assert L.vdl_column_alloc(ctx, name.encode(), arr.dtype.itemsize, len(arr), C.byref(v)) == 0, L.vdl_last_error(ctx)
    assert L.vdl_column_upload(ctx, v, arr.ctypes.data, len(arr)) == 0, L.vdl_last_error(ctx)


def run_comm(text, shard_cols, row_bases, steps=3):
    L = L_.load()
    n = len(shard_cols)
    comm = C.c_void_p()
    devs = (C.c_int * n)(*([0] * n))
    assert L.vdl_comm_init_all(n, devs, C.byref(comm)) == 0
    assert L.vdl_comm_size(comm) == n
    for r in range(n):
        ctx = C.c_void_p(L.vdl_comm_ctx(comm, r))
        for k, v in shard_cols[r].items():
            upload(L, ctx, k, np.ascontiguousarray(v))
    cp = C.c_void_p()
    bases = (C.c_int64 * n)(*row_bases)
    rc = L.vdl_comm_plan_load(comm, text.encode(), L_.VDL_PLAN_FUSE, bases, C.byref(cp))
    assert rc == 0, L.vdl_comm_last_error(comm)
    results = []
    for _ in range(steps):                     # the first step wires the exchange buffers; then both epoch parities
        rc = L.vdl_comm_plan_run(cp)
        assert rc == 0, L.vdl_comm_last_error(comm)
        results.append([outputs(L, C.c_void_p(L.vdl_comm_plan_rank(cp, r))) for r in range(n)])
    L.vdl_comm_plan_destroy(cp)
    L.vdl_comm_destroy(comm)
    return results


@pytest.mark.parametrize("query,colnames", [("q06.vdl", Q6_COLS), ("q01.vdl", Q1_COLS)])
@pytest.mark.parametrize("world", [2, 3])
def test_single_process_ranks_scan_plans(catalog, query, colnames, world):
    rows, text = 70_001, plan_text(query)
    names = ["lineitem." + c for c in colnames]
    want = run_oracle(text, host_columns(catalog, names, {"lineitem": rows}))
    shards, bases = [], []
    for r in range(world):
        start, n = tpch.shard_range(rows, r, world)
        shards.append(host_columns(catalog, names, {"lineitem": n}, row_offset=start))
        bases.append(start)
    for step in run_comm(text, shards, bases):
        for got in step:
            assert_same(got, want)


def test_single_process_ranks_fk_join_plan(catalog):
    """Q5: the probe kernel's fold groups are combined the same way (dimension tables replicated on every rank)."""
    text, sf, world = plan_text("q05.vdl"), 0.02, 2
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    names = tpch.plan_columns(text)
    cols = host_columns(catalog, names, rows, sf=sf)
    want = run_oracle(text, cols)
    shards, bases = [], []
    for r in range(world):
        start, n = tpch.shard_range(rows["lineitem"], r, world)
        shards.append({k: (v[start:start + n] if k.startswith("lineitem.") else v) for k, v in cols.items()})
        bases.append(start)
    for step in run_comm(text, shards, bases):
        for got in step:
            assert_same(got, want)


def test_emit_plans_are_refused_not_miscomputed(catalog):
    text, sf = plan_text("q03.vdl"), 0.01
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    half = rows["lineitem"] // 2 // 4096 * 4096
    shards = [{k: (v[:half] if k.startswith("lineitem.") else v) for k, v in cols.items()},
              {k: (v[half:] if k.startswith("lineitem.") else v) for k, v in cols.items()}]
    with pytest.raises(AssertionError, match="emit"):
        run_comm(text, shards, [0, half], steps=1)
