"""TPC-H Q6 / Q1 plans on the GPU (fused single launch and op-at-a-time) against the CPU oracle, bit-exact."""
import numpy as np
import pytest

from mplan2vdl_b200 import synth, tpch
from oracle import sqlref
from oracle.oracle import gen_column
from util import Q1_COLS, Q6_COLS, assert_same, host_columns, plan_text, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows", [1, 17, 1023, 1024, 1025, 300_000, 2_000_003])
def test_q6_parity(catalog, rows):
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": rows})
    want = run_oracle(plan_text("q06.vdl"), cols)
    got, stats = run_gpu(plan_text("q06.vdl"), cols, fuse=True)
    assert stats["fused_scans"] == 1 and stats["nodes"] < stats["statements"]
    assert_same(got, want)
    if rows <= 300_000:
        got_u, stats_u = run_gpu(plan_text("q06.vdl"), cols, fuse=False)
        assert stats_u["fused_scans"] == 0
        assert_same(got_u, want)


@pytest.mark.parametrize("query,colnames,shape", [("q06.vdl", Q6_COLS, "sel3_sum2"), ("q01.vdl", Q1_COLS, "sel1_key2_sum5")])
def test_static_shape_and_generic_kernel_agree(catalog, query, colnames, shape, monkeypatch):
    """The static-shape instantiation is what runs on TPC-H-shaped data; forcing the generic kernel gives the same bits."""
    cols = host_columns(catalog, ["lineitem." + c for c in colnames], {"lineitem": 700_001})
    want = run_oracle(plan_text(query), cols)
    got, stats = run_gpu(plan_text(query), cols)
    assert stats["shape"] == shape
    assert_same(got, want)
    monkeypatch.setenv("VDL_GENERIC_ONLY", "1")
    got_g, stats_g = run_gpu(plan_text(query), cols)
    assert stats_g["shape"] == "generic"
    assert_same(got_g, want)


def test_wide_values_fall_back_to_the_generic_kernel(catalog):
    """Values that do not fit int32 break the static shape's narrowing assumption: the generic kernel must be chosen."""
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": 100_000})
    cols["lineitem.l_extendedprice"] = cols["lineitem.l_extendedprice"] * 1_000_003   # ~1e13: wraps in the sum, exactly as the oracle
    want = run_oracle(plan_text("q06.vdl"), cols)
    got, stats = run_gpu(plan_text("q06.vdl"), cols)
    assert stats["shape"] == "generic"
    assert_same(got, want)


def test_q6_empty_selection(catalog):
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": 5000})
    cols["lineitem.l_quantity"][:] = 5000
    for fuse in (True, False):
        got, _ = run_gpu(plan_text("q06.vdl"), cols, fuse=fuse)
        assert got["revenue"].shape == (0,)


@pytest.mark.parametrize("rows", [1, 33, 1024, 100_003, 1_500_001])
def test_q1_parity(catalog, rows):
    cols = host_columns(catalog, ["lineitem." + c for c in Q1_COLS], {"lineitem": rows})
    want = run_oracle(plan_text("q01.vdl"), cols)
    got, stats = run_gpu(plan_text("q01.vdl"), cols, fuse=True)
    assert stats["fused_scans"] == 1        # all eight Folds share one scan
    assert_same(got, want)
    if rows <= 100_003:
        got_u, _ = run_gpu(plan_text("q01.vdl"), cols, fuse=False)
        assert_same(got_u, want)


def test_q1_many_groups_overflow_the_slot_map(catalog):
    """More distinct keys than lane-private slots: the extra keys take the global-atomic path, same answer."""
    cols = host_columns(catalog, ["lineitem." + c for c in Q1_COLS], {"lineitem": 200_000})
    rng = np.random.default_rng(0)
    cols["lineitem.l_returnflag"] = (16 + 8 * rng.integers(0, 7, 200_000)).astype(np.int64)
    cols["lineitem.l_linestatus"] = (16 + 8 * rng.integers(0, 4, 200_000)).astype(np.int64)
    want = run_oracle(plan_text("q01.vdl"), cols)
    assert len(want["count_order"]) == 28
    got, _ = run_gpu(plan_text("q01.vdl"), cols, fuse=True)
    assert_same(got, want)


def test_device_generator_matches_host_generator(catalog):
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    rows, off, seed = 100_003, 12_345, synth.seed_for(1)
    for name in ["lineitem." + c for c in Q1_COLS] + ["lineitem.lineitem_orders", "lineitem.l_orderkey", "orders.o_orderkey"]:
        spec = synth.column_spec(catalog, name, 1)
        v = ctx.fill_synthetic(name, spec, rows, seed, off)
        np.testing.assert_array_equal(ctx.download(v), gen_column(spec, rows, off, seed).astype(np.int64), err_msg=name)
    ctx.close()


def test_q6_sf1_in_place_and_sql(catalog):
    """BASELINE config 1 (SF1), columns generated in HBM; checked against the SQL-level numpy evaluation."""
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    text = plan_text("q06.vdl")
    names = tpch.plan_columns(text)
    info = tpch.load_synthetic(ctx, catalog, names, 1)
    assert info["rows"]["lineitem"] == 6_001_215
    got = ctx.plan(text).run()
    cols = host_columns(catalog, names, {"lineitem": 6_001_215}, sf=1)
    assert_same(got, sqlref.q6(cols))
    ctx.close()


def test_q6_full_size_linearity(catalog):
    """SF100-sized property: the revenue of the whole table equals the wrapped sum over 4 row-range shards."""
    from mplan2vdl_b200.executor import Context
    sf, text = 100, plan_text("q06.vdl")
    names = tpch.plan_columns(text)
    ctx = Context(0)
    tpch.load_synthetic(ctx, catalog, names, sf)
    whole = ctx.plan(text).run()["revenue"]
    for n in names:
        ctx.drop_column(n)
    total = np.int64(0)
    with np.errstate(over="ignore"):
        for rank in range(4):
            tpch.load_synthetic(ctx, catalog, names, sf, rank=rank, world=4)
            total = total + ctx.plan(text).run()["revenue"][0]
            for n in names:
                ctx.drop_column(n)
    assert whole.shape == (1,) and whole[0] == total
    ctx.close()


@pytest.mark.parametrize("q", ["q03.vdl", "q05.vdl", "q12.vdl"])
@pytest.mark.parametrize("sf", [0.002, 0.02])
def test_fk_join_plans_parity(catalog, q, sf):
    """BASELINE config 3: the FK-join plans (Gather through join-index columns, Scatter-built validity / inverse
    index over the dimension, FoldSelect compaction, radix Partition on the composite key) op-at-a-time on the GPU."""
    text = plan_text(q)
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    want = run_oracle(text, cols)
    got, stats = run_gpu(text, cols)
    assert_same(got, want)
    assert stats["probe_folds"] + stats["probe_emits"] >= 1        # the join chain runs on the probe kernel
    got_u, _ = run_gpu(text, cols, fuse=False)
    assert_same(got_u, want)
    assert len(next(iter(want.values()))) > 0 or q == "q12.vdl"   # (Q12 selects ~0.1 % of the rows: may be empty at SF 0.002)


def test_cli_prints_the_server_json_and_the_decoded_csv():
    """`python -m mplan2vdl_b200` stands where the HTTP Voodoo server + resolve.py stood (eval_query.sh:18-26)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    plan = os.path.join(root, "plans", "q05.vdl")
    out = subprocess.run([sys.executable, "-m", "mplan2vdl_b200", plan, "--sf", "0.01"], cwd=root, capture_output=True, text=True, check=True).stdout
    doc = json.loads(out)
    assert list(doc["results"]["tmp0"]) == [".n_name__nation__n_name"] and list(doc["results"]["tmp1"]) == [".revenue"]
    csv = subprocess.run([sys.executable, "-m", "mplan2vdl_b200", "-", "--sf", "0.01", "--csv"], cwd=root, input=open(plan).read(),
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert csv[0] == ".n_name,.revenue" and len(csv) == 1 + len(doc["results"]["tmp1"][".revenue"])


def test_q19_parity(catalog):
    """Q19: the OR of three conjunctions does not normalise to a conjunct, so the plan mixes the probe kernel (the
    join below the OR) with op-at-a-time evaluation above it; fused, unfused and oracle agree."""
    from util import q19_columns
    text, cols = q19_columns(catalog, sf=0.02)
    want = run_oracle(text, cols)
    assert len(want["revenue"]) == 1
    got, stats = run_gpu(text, cols)
    assert_same(got, want)
    got_u, _ = run_gpu(text, cols, fuse=False)
    assert_same(got_u, want)


def test_avg_over_a_join_is_a_post_op_of_the_probe(catalog):
    """AVG = Divide(FoldSum x, FoldSum 1) (Vlite.hs:1038-1041) over a joined space: both Folds run in the one probe pass
    and the Divide inside its finalize kernel -- no emitted vectors, no op-at-a-time tail."""
    from mplan2vdl_b200 import tpch_queries as Q, vlite
    from mplan2vdl_b200.vlite import Bin, Cast, GroupBy, Join, Lit, Project, Ref, Select, Table
    orders = Select(Table("orders", [("orders.o_orderdate", None), ("orders.%TID%", None)]),
                    Q.between(Lit(Q.DATE, Q.day(1994, 1, 1)), Ref("orders.o_orderdate"), Lit(Q.DATE, Q.day(1996, 1, 1))))
    lineitem = Table("lineitem", Q.li("l_quantity", "l_returnflag") + [("lineitem.lineitem_orders", "lineitem.%lineitem_orders")])
    j = Join(lineitem, orders, [Bin("Eq", Ref("lineitem.%lineitem_orders"), Ref("orders.%TID%"))])
    g = GroupBy(j, [("lineitem.l_returnflag", None)], [(("FChoose", Ref("lineitem.l_returnflag")), None),
                                                        (("Avg", Cast(None, Ref("lineitem.l_quantity"))), "L1.L1"), (("Count",), "L2.L2")])
    text = vlite.translate(catalog, Project(g, [(Ref("lineitem.l_returnflag"), None), (Ref("L1"), "L1.avg_qty"), (Ref("L2"), "L2.cnt")]))
    rows = {t: synth.table_rows(catalog, t, 0.01) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=0.01)
    want = run_oracle(text, cols)
    got, stats = run_gpu(text, cols)
    assert_same(got, want)
    assert stats["probe_folds"] == 1 and stats["probe_emits"] == 0 and stats["launches"] <= 4
    assert len(want["avg_qty"]) == 3


@pytest.mark.parametrize("q", ["q01", "q03", "q05", "q06", "q12", "q19"])
@pytest.mark.parametrize("fuse", [True, False])
def test_cuda_path_reproduces_the_committed_answers(catalog, q, fuse):
    """libvdl_cuda (fused paths, and op-at-a-time) against tests/golden/tpch_sf0.01_answers.json: known answers that do
    not depend on the CPU oracle being run next to it."""
    from util import golden_case
    text, cols, want = golden_case(catalog, q)
    got, _ = run_gpu(text, cols, fuse=fuse)
    assert_same(got, want)
