"""Typed result columns (include/vdl_cuda.h vdl_plan_set_typed_outputs): the reference's result is text, so the width a
column crosses PCIe in is the executor's choice.  An op-at-a-time output whose every value is a value of a 4-byte column --
by provenance: emitted by a probe pass as a plain column, then Gather / Scatter / FoldChoose / FoldMin / FoldMax -- travels
as int32; sums and anything computed stay int64.  Values must not change."""
import ctypes as C

import numpy as np
import pytest

from mplan2vdl_b200 import synth, tpch
from util import assert_same, host_columns, plan_text, run_oracle

pytestmark = pytest.mark.gpu


def q3_case(catalog, sf=0.01):
    text = plan_text("q03.vdl")
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    return text, host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)


def test_q3_key_columns_travel_as_int32_with_the_same_values(catalog):
    from mplan2vdl_b200.executor import Context
    text, cols = q3_case(catalog)
    want = run_oracle(text, cols)
    ctx = Context(0)
    for k, v in cols.items():
        ctx.upload_column(k, v)
    plan = ctx.plan(text)
    wide = plan.run()
    assert all(v.dtype == np.int64 for v in wide.values())
    assert_same(wide, want)
    plan.set_typed_outputs(True)
    typed = plan.run()
    assert_same(typed, want)
    dt = {k: v.dtype for k, v in typed.items()}
    assert dt["revenue"] == np.int64                                    # a sum of products
    narrow = [k for k, d in dt.items() if d == np.int32]
    stored32 = {c for c in ("l_orderkey", "o_orderdate", "o_shippriority") if catalog.column({"l": "lineitem.", "o": "orders."}[c[0]] + c).width == 4}
    assert {k.split("__")[0] for k in narrow} == stored32 and stored32     # exactly the FoldChoose outputs of 4-byte columns
    # the int64 accessor refuses a column that was delivered as int32 instead of handing out half-width data
    name, data, n = C.c_char_p(), C.POINTER(C.c_int64)(), C.c_int64()
    i = list(typed).index(narrow[0])
    assert plan.L.vdl_plan_output(plan.h, i, C.byref(name), C.byref(data), C.byref(n)) != 0
    assert "vdl_plan_output_typed" in ctx.L.vdl_last_error(ctx.h).decode()
    plan.set_typed_outputs(False)
    assert all(v.dtype == np.int64 for v in plan.run().values())
    plan.close()
    ctx.close()


@pytest.mark.parametrize("world", [2, 3])
def test_typed_outputs_with_the_sharded_tail(catalog, world):
    from mplan2vdl_b200.dist import merge_tail_boundaries
    from mplan2vdl_b200.executor import Context
    text, cols = q3_case(catalog)
    want = run_oracle(text, cols)
    nli = len(cols["lineitem.lineitem_l_orderkey_l_linenumber_pkey"])
    ctxs, plans, recs = [], [], []
    for rank in range(world):
        start, n = tpch.shard_range(nli, rank, world)
        ctx = Context(0)
        for k, v in cols.items():
            ctx.upload_column(k, v[start:start + n] if k.startswith("lineitem.") else v)
        plan = ctx.plan(text)
        ops = plan.tail_info()
        plan.tail_enable(True)
        plan.set_typed_outputs(True)
        plan.run()
        recs.append(plan.tail_boundary())
        ctxs.append(ctx)
        plans.append(plan)
    got = {k: [] for k in want}
    for rank in range(world):
        plans[rank].tail_apply(*merge_tail_boundaries(recs, ops, rank))
        for k, v in plans[rank].outputs().items():
            got[k].append(v)
    assert any(v.dtype == np.int32 for v in got[next(iter(want))] + got[list(want)[2]])
    assert_same({k: np.concatenate(v) for k, v in got.items()}, want)
    for p, c in zip(plans, ctxs):
        p.close()
        c.close()


def test_computed_columns_stay_wide(catalog):
    """Q1's outputs come from a fused scan (its result buffer is int64) and Q19's single SUM is computed: nothing narrows."""
    from mplan2vdl_b200.executor import Context
    from util import Q1_COLS
    cols = host_columns(catalog, ["lineitem." + c for c in Q1_COLS], {"lineitem": 20_000})
    ctx = Context(0)
    for k, v in cols.items():
        ctx.upload_column(k, v)
    plan = ctx.plan(plan_text("q01.vdl"))
    plan.set_typed_outputs(True)
    out = plan.run()
    assert all(v.dtype == np.int64 for v in out.values())
    assert_same(out, run_oracle(plan_text("q01.vdl"), cols))
    plan.close()
    ctx.close()
