"""Sharded tail (include/vdl_cuda.h vdl_plan_tail_*, mplan2vdl_b200/dist.py merge_tail_boundaries): plans whose outputs are
all Folds by runs of one sorted (or constant) groups vector keep their result sharded; only the groups that straddle shard
boundaries are merged, from one small record per rank.  The merge rule is tested on the host against a whole-vector fold;
the GPU tests run Q3 / Q19 on emulated ranks through the C ABI and compare the concatenated slices with the CPU oracle."""
import numpy as np
import pytest

from mplan2vdl_b200 import tpch
from mplan2vdl_b200.dist import FOLD_CHOOSE, FOLD_COUNT, FOLD_MAX, FOLD_MIN, FOLD_SUM, merge_tail_boundaries
from util import assert_same, host_columns, plan_text, run_oracle

OPS = [FOLD_CHOOSE, FOLD_SUM, FOLD_MIN, FOLD_MAX, FOLD_COUNT]


def fold_by_runs(keys, vals, op):
    if len(keys) == 0:
        return np.zeros(0, np.int64)
    heads = np.flatnonzero(np.r_[True, keys[1:] != keys[:-1]])
    if op == FOLD_CHOOSE:
        return vals[heads]
    if op == FOLD_COUNT:
        return np.diff(np.r_[heads, len(keys)]).astype(np.int64)
    f = {FOLD_SUM: np.add, FOLD_MIN: np.minimum, FOLD_MAX: np.maximum}[op]
    return f.reduceat(vals, heads)


def sharded(keys, vals, cuts):
    """Local folds per shard, boundary records, the merge of every rank, the slices concatenated."""
    bounds = [0] + list(cuts) + [len(keys)]
    world = len(bounds) - 1
    local, recs = [], []
    for r in range(world):
        k, v = keys[bounds[r]:bounds[r + 1]], vals[bounds[r]:bounds[r + 1]]
        outs = [fold_by_runs(k, v, op) for op in OPS]
        runs = len(outs[0])
        recs.append([1, runs, int(k[0]) if runs else 0, int(k[-1]) if runs else 0] +
                    [int(o[0]) if runs else 0 for o in outs] + [int(o[-1]) if runs else 0 for o in outs])
        local.append(outs)
    slices = []
    for r in range(world):
        m = merge_tail_boundaries(recs, OPS, r)
        if m is None:
            return None
        drop, last = m
        outs = [o.copy() for o in local[r]]
        if last is not None:
            for o, x in zip(outs, last):
                o[-1] = x
        slices.append([o[1:] if drop else o for o in outs])
    return [np.concatenate([s[i] for s in slices]) for i in range(len(OPS))]


@pytest.mark.parametrize("seed", range(12))
def test_merged_slices_equal_the_whole_vector_fold(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    keys = np.sort(rng.integers(0, max(2, n // int(rng.integers(1, 40))), n)).astype(np.int64)
    vals = rng.integers(-1000, 1000, n).astype(np.int64)
    world = int(rng.integers(2, 7))
    cuts = np.sort(rng.integers(0, n + 1, world - 1))            # empty shards and runs over several shards included
    got = sharded(keys, vals, cuts)
    for g, op in zip(got, OPS):
        np.testing.assert_array_equal(g, fold_by_runs(keys, vals, op), err_msg=f"op {op} cuts {cuts}")


def test_one_group_over_every_rank_lands_on_the_first():
    keys, vals = np.zeros(10, np.int64), np.arange(10, dtype=np.int64)
    got = sharded(keys, vals, [3, 3, 7])
    assert [list(g) for g in got] == [[0], [45], [0], [9], [10]]


def test_unordered_shards_ask_for_the_fallback():
    recs = [[1, 2, 5, 9, 0, 0], [1, 2, 7, 12, 0, 0]]             # rank 1 starts below rank 0's last key
    assert merge_tail_boundaries(recs, [FOLD_SUM], 0) is None
    recs = [[1, 2, 5, 9, 0, 0], [0, 2, 9, 12, 0, 0]]             # rank 1's keys were not in order locally
    assert merge_tail_boundaries(recs, [FOLD_SUM], 1) is None


def test_sum_merges_wrap_like_int64():
    big = 2 ** 63 - 1
    recs = [[1, 1, 4, 4, big, big], [1, 1, 4, 4, 2, 2]]
    assert merge_tail_boundaries(recs, [FOLD_SUM], 0) == (False, [-(2 ** 63) + 1])
    assert merge_tail_boundaries(recs, [FOLD_SUM], 1) == (True, None)


@pytest.mark.gpu
@pytest.mark.parametrize("q", ["q03.vdl", "q19.vdl"])
@pytest.mark.parametrize("world,sf", [(2, 0.01), (3, 0.01), (5, 0.01), (4, 0.002)])      # (4, 0.002): the last rank's shard is empty
def test_emit_plans_keep_their_result_sharded(catalog, q, world, sf):
    from mplan2vdl_b200 import synth
    from mplan2vdl_b200.executor import Context
    from util import q19_columns
    text = plan_text(q)
    if q == "q19.vdl":
        text, cols = q19_columns(catalog, sf=sf)
    else:
        rows_all = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
        cols = host_columns(catalog, tpch.plan_columns(text), rows_all, sf=sf)
    nli = len(cols["lineitem.lineitem_l_orderkey_l_linenumber_pkey"])
    want = run_oracle(text, cols)
    ctxs, plans, recs = [], [], []
    for rank in range(world):
        start, n = tpch.shard_range(nli, rank, world)
        ctx = Context(0)
        for k, v in cols.items():
            ctx.upload_column(k, v[start:start + n] if k.startswith("lineitem.") else v)
        plan = ctx.plan(text)
        ops = plan.tail_info()
        assert ops is not None and plan.num_emits >= 1
        plan.tail_enable(True)
        plan.set_row_base(start)
        outs = plan.run()
        rec = plan.tail_boundary()
        runs = rec[1]
        assert all(len(v) == runs for v in outs.values())
        assert rec[4:] == [int(v[0]) if runs else 0 for v in outs.values()] + [int(v[-1]) if runs else 0 for v in outs.values()]
        recs.append(rec)
        ctxs.append(ctx)
        plans.append(plan)
    got = {k: [] for k in want}
    for rank in range(world):
        plans[rank].tail_apply(*merge_tail_boundaries(recs, ops, rank))
        for k, v in plans[rank].outputs().items():
            got[k].append(v)
    assert_same({k: np.concatenate(v) for k, v in got.items()}, want)
    for p, c in zip(plans, ctxs):
        p.close()
        c.close()


@pytest.mark.gpu
def test_plans_of_another_shape_are_not_offered_the_sharded_tail(catalog):
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    cols = host_columns(catalog, ["lineitem." + c for c in ("l_quantity", "l_extendedprice", "l_discount", "l_shipdate")], {"lineitem": 1000})
    for k, v in cols.items():
        ctx.upload_column(k, v)
    plan = ctx.plan(plan_text("q06.vdl"))           # a fused scan: its partial table is what combines
    assert plan.tail_info() is None
    with pytest.raises(Exception):
        plan.tail_enable(True)
    plan.close()
    ctx.close()


def test_emit_plans_over_a_dimension_table_are_refused_not_miscomputed():
    """A probe pass over a replicated table would emit the same survivors on every rank, and a semijoin's dimension side needs
    matches from every rank (Q4, Q20): dist.ShardedPlan refuses them before anything runs."""
    from mplan2vdl_b200.dist import check_shardable

    class P:
        def __init__(self, tables):
            self.num_emits, self._t = len(tables), tables

        def emit_tables(self):
            return self._t
    check_shardable(P([]), 8, "lineitem")                      # scans / probe folds: nothing emitted
    check_shardable(P(["lineitem"]), 8, "lineitem")            # Q3, Q19
    check_shardable(P(["orders", "lineitem"]), 1, "lineitem")  # one GPU: anything goes
    for tables in (["orders"], ["orders", "lineitem"], ["lineitem", "lineitem"]):
        with pytest.raises(NotImplementedError, match="one GPU"):
            check_shardable(P(tables), 2, "lineitem")


@pytest.mark.gpu
def test_emit_group_tables_are_reported(catalog):
    from mplan2vdl_b200 import synth
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    for q, want in (("q03.vdl", ["lineitem"]), ("q06.vdl", [])):
        text = plan_text(q)
        rows = {t: synth.table_rows(catalog, t, 0.002) for t in catalog.tables}
        for k, v in host_columns(catalog, tpch.plan_columns(text), rows, sf=0.002).items():
            try:
                ctx.lookup(k)
            except Exception:
                ctx.upload_column(k, v)
        plan = ctx.plan(text)
        assert plan.emit_tables() == want
        plan.close()
    ctx.close()
