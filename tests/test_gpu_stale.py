"""A prepared plan must follow the DATA, not the handle: column statistics (the executor's stand-in for bounds.csv /
inferBounds, Vlite.hs:417-467) are baked into the fused scan as int32-narrowing proofs, 32-bit accumulator proofs
(RK_N32 / RK_MADW) and the static shape.  Rewriting a column under a plan that has already run -- vdl_column_upload into
the same handle, drop + re-upload (handles are recycled), or writing into caller-owned bound memory + vdl_column_touch
-- has to re-prove them.  Every case: run, rewrite, re-run the SAME plan object, compare with the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from mplan2vdl_b200 import synth, tpch
from mplan2vdl_b200.lib import VdlError
from util import Q1_COLS, Q6_COLS, assert_same, host_columns, plan_text, run_oracle

pytestmark = pytest.mark.gpu


def _upload(ctx, cols):
    return {k: ctx.upload_column(k, v) for k, v in cols.items()}


def test_q6_same_handle_wider_values(catalog):
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": 400_003})
    h = _upload(ctx, cols)
    text = plan_text("q06.vdl")
    plan = ctx.plan(text)
    assert_same(plan.run(), run_oracle(text, cols))
    assert plan.shape(0) == "sel3_sum2"
    wide = dict(cols)
    wide["lineitem.l_extendedprice"] = cols["lineitem.l_extendedprice"] * 1_000_003       # ~1e13: not an int32 any more
    g0 = ctx.generation(h["lineitem.l_extendedprice"])
    ctx.upload_into(h["lineitem.l_extendedprice"], wide["lineitem.l_extendedprice"].ctypes.data, 400_003)
    assert ctx.generation(h["lineitem.l_extendedprice"]) > g0
    assert_same(plan.run(), run_oracle(text, wide))
    assert plan.shape(0) != "sel3_sum2"           # the narrowing assumption of the static shape is gone
    # ... and back: narrow data again re-qualifies
    ctx.upload_into(h["lineitem.l_extendedprice"], cols["lineitem.l_extendedprice"].ctypes.data, 400_003)
    assert_same(plan.run(), run_oracle(text, cols))
    # a predicate column too (mode-2 low-word compares): dates shifted out of int32 on an int64 quantity column
    wide2 = dict(cols)
    wide2["lineitem.l_quantity"] = cols["lineitem.l_quantity"] + (np.int64(1) << 33) * (np.arange(400_003) % 2)
    ctx.upload_into(h["lineitem.l_quantity"], wide2["lineitem.l_quantity"].ctypes.data, 400_003)
    assert_same(plan.run(), run_oracle(text, wide2))
    plan.close()
    ctx.close()


def test_q1_accumulator_proofs_follow_the_data(catalog):
    """Q1's register-slot kernel keeps sum(l_quantity) in 32 bits per thread (RK_N32) and multiplies with mad.wide
    (RK_MADW) on the strength of the column maxima: larger values must re-prepare."""
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    rows = 600_011
    cols = host_columns(catalog, ["lineitem." + c for c in Q1_COLS], {"lineitem": rows})
    h = _upload(ctx, cols)
    text = plan_text("q01.vdl")
    plan = ctx.plan(text)
    assert_same(plan.run(), run_oracle(text, cols))
    assert plan.shape(0) == "sel1_key2_sum5"
    big = dict(cols)
    big["lineitem.l_quantity"] = cols["lineitem.l_quantity"] * 900_001                     # per-thread sums leave int32
    big["lineitem.l_extendedprice"] = cols["lineitem.l_extendedprice"] * 70_001           # products leave 32x32
    for k in ("lineitem.l_quantity", "lineitem.l_extendedprice"):
        ctx.upload_into(h[k], big[k].ctypes.data, rows)
    assert_same(plan.run(), run_oracle(text, big))
    neg = dict(cols)
    neg["lineitem.l_extendedprice"] = -cols["lineitem.l_extendedprice"]                    # the unsigned multiply-add needs factors >= 0
    for k in ("lineitem.l_quantity", "lineitem.l_extendedprice"):
        ctx.upload_into(h[k], neg[k].ctypes.data, rows)
    assert_same(plan.run(), run_oracle(text, neg))
    plan.close()
    ctx.close()


def test_drop_and_reupload_between_runs(catalog):
    """vdl_column_drop + a new column of the same name and length usually gets the SAME handle back (lowest free slot):
    the plan must notice through the write generation, not keep the freed pointer and the old proofs."""
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    rows = 250_000
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": rows})
    h = _upload(ctx, cols)
    text = plan_text("q06.vdl")
    plan = ctx.plan(text)
    assert_same(plan.run(), run_oracle(text, cols))
    other = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": rows}, seed=777)
    other["lineitem.l_extendedprice"] = other["lineitem.l_extendedprice"] * 1_000_003
    recycled = 0
    for k in cols:
        ctx.drop_column(k)
        h2 = ctx.upload_column(k, other[k])
        recycled += h2 == h[k]
    assert_same(plan.run(), run_oracle(text, other))
    assert recycled >= 1, "the scenario under test is a recycled handle"
    plan.close()
    ctx.close()


def test_bound_tensor_touch(catalog):
    import torch
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    rows = 200_000
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": rows})
    tens, h = {}, {}
    for k, v in cols.items():
        tens[k] = torch.from_numpy(v).cuda()
        h[k] = ctx.bind_tensor(k, tens[k])
    text = plan_text("q06.vdl")
    plan = ctx.plan(text)
    assert_same(plan.run(), run_oracle(text, cols))
    wide = dict(cols)
    wide["lineitem.l_extendedprice"] = cols["lineitem.l_extendedprice"] * 1_000_003
    tens["lineitem.l_extendedprice"].copy_(torch.from_numpy(wide["lineitem.l_extendedprice"]))
    torch.cuda.synchronize()
    ctx.touch(h["lineitem.l_extendedprice"])
    assert_same(plan.run(), run_oracle(text, wide))
    plan.close()
    ctx.close()


def test_probe_plan_follows_rewritten_columns(catalog):
    """FK-join plan on the probe kernel: rewrite a dimension column and a fact column under the same handles, and
    drop + re-create one (the probe descriptor holds device pointers and lengths)."""
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    text = plan_text("q05.vdl")
    sf = 0.02
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    h = _upload(ctx, cols)
    plan = ctx.plan(text)
    assert_same(plan.run(), run_oracle(text, cols))
    assert plan.stats()["probe_folds"] >= 1
    new = dict(cols)
    rng = np.random.default_rng(5)
    new["orders.o_orderdate"] = rng.permutation(cols["orders.o_orderdate"])
    new["lineitem.l_extendedprice"] = cols["lineitem.l_extendedprice"] * 1_000_003
    for k in ("orders.o_orderdate", "lineitem.l_extendedprice"):
        ctx.upload_into(h[k], new[k].ctypes.data, len(new[k]))
    assert_same(plan.run(), run_oracle(text, new))
    new["lineitem.l_discount"] = rng.permutation(cols["lineitem.l_discount"])
    ctx.drop_column("lineitem.l_discount")
    ctx.upload_column("lineitem.l_discount", new["lineitem.l_discount"])
    assert_same(plan.run(), run_oracle(text, new))
    plan.close()
    ctx.close()


def test_direct_fused_launch_on_stale_columns_fails_loudly(catalog):
    """Below the plan layer nobody can re-prepare on the caller's behalf: vdl_fused_launch on a scan whose column was
    rewritten returns VDL_ESTALE instead of computing with old proofs."""
    from mplan2vdl_b200 import lib as L_
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    L = ctx.L
    rows = 10_000
    x = np.arange(rows, dtype=np.int64) % 1000
    v = ctx.upload_column("t.x", x)
    d = L_.FusedDesc()
    d.rows, d.row_base, d.ncolumns = rows, 0, 1
    d.column[0] = v
    d.key_mask, d.domain, d.nfolds = -1, 1, 1
    d.fold[0].op, d.fold[0].nfactors = 0, 1
    d.fold[0].factor[0] = L_.Affine(0, 0, 0, 1)
    f = C.c_void_p()
    ctx.check(L.vdl_fused_prepare(ctx.h, C.byref(d), C.byref(f)))
    ctx.check(L.vdl_fused_launch_ex(f, 1))
    data, n = C.POINTER(C.c_int64)(), C.c_int64()
    ctx.check(L.vdl_fused_result_host(f, 0, C.byref(data), C.byref(n)))
    assert n.value == 1 and data[0] == int(x.sum())
    y = x * (1 << 40)
    ctx.upload_into(v, y.ctypes.data, rows)
    with pytest.raises(VdlError) as e:
        ctx.check(L.vdl_fused_launch_ex(f, 1))
    assert e.value.code == 7
    L.vdl_fused_destroy(f)
    ctx.close()
