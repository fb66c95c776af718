"""Randomised plans (tests/fuzz_plans.py): the program the translator restatement prints for a random query must give
the same bits through the CPU oracle, the fused GPU paths (fused scan / FK-join probe, whichever the fusion passes
pick) and the op-at-a-time GPU path."""
import pytest

import fuzz_plans
from mplan2vdl_b200 import synth, tpch, vlite
from util import assert_same, host_columns, run_gpu, run_oracle

SF = 0.004


def _run(catalog, rel, need_gpu=True, sf=SF):
    text = vlite.translate(catalog, rel)
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    want = run_oracle(text, cols)
    if not need_gpu:
        return want, None
    got, stats = run_gpu(text, cols, fuse=True)
    assert_same(got, want)
    got_u, stats_u = run_gpu(text, cols, fuse=False)
    assert stats_u["fused_scans"] == 0 and stats_u["probe_folds"] == 0 and stats_u["probe_emits"] == 0
    assert_same(got_u, want)
    stats["loads"] = text.count(",Load,")
    return want, stats


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(40))
def test_random_single_table_plans(catalog, seed):
    _, stats = _run(catalog, fuzz_plans.single_table(seed))
    # every select -> map -> grouped fold over one table runs as ONE fused launch: the TMA-staged scan, or -- IN lists,
    # column-vs-column comparisons, CASE sums -- a probe fold (a bare COUNT(*) reads no column: per-op)
    assert stats["fused_scans"] + stats["probe_folds"] >= 1 or stats["loads"] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(40))
def test_random_fk_join_plans(catalog, seed):
    _, stats = _run(catalog, fuzz_plans.join_query(seed, catalog))
    assert stats["probe_folds"] + stats["probe_emits"] >= 1      # the join chain runs on the probe kernel


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [100, 101, 102, 103, 104, 105])
def test_random_plans_many_tiles(catalog, seed):
    """~300 K lineitem rows: hundreds of probe tiles / scan tiles, look-back across tiles, several CTAs per scan."""
    _run(catalog, fuzz_plans.single_table(seed), sf=0.05)
    _run(catalog, fuzz_plans.join_query(seed, catalog), sf=0.05)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_plans_translate_and_run_in_the_oracle(catalog, seed):
    for rel in (fuzz_plans.single_table(seed), fuzz_plans.join_query(seed, catalog)):
        want, _ = _run(catalog, rel, need_gpu=False)
        assert len(want) >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(0, 40, 2))
@pytest.mark.parametrize("jit", [True, False])
def test_random_plans_through_map_clusters(catalog, seed, jit, monkeypatch):
    """The same random plans with the probe pass switched off (VDL_NO_PROBE): predicates, CASE arithmetic and FK fetches
    stay in the op-at-a-time remainder, which build_clusters (vdl_plan.cu) turns into vdl_op_map register programs --
    interpreted (VDL_NO_JIT) or compiled at run time for the vectors of >= 65536 rows."""
    monkeypatch.setenv("VDL_NO_PROBE", "1")
    if not jit:
        monkeypatch.setenv("VDL_NO_JIT", "1")
    sf = 0.02                       # ~120 K lineitem rows: long enough for the specialised kernels
    nodes = 0
    for rel in (fuzz_plans.single_table(seed), fuzz_plans.join_query(seed, catalog)):
        text = vlite.translate(catalog, rel)
        rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
        cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
        got, stats = run_gpu(text, cols, fuse=True)
        assert stats["probe_folds"] == 0 and stats["probe_emits"] == 0
        assert_same(got, run_oracle(text, cols))
        nodes += stats["map_nodes"]
    assert nodes >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(0, 40, 3))
def test_random_single_table_plans_through_the_runtime_compiled_scan(catalog, seed, monkeypatch):
    """The fused scan specialised at run time (NVRTC) to the shape of each random descriptor -- forced for these small
    tables with VDL_SCAN_JIT_MIN_ROWS=0 -- against the oracle, and against the generic kernel on the same data."""
    monkeypatch.setenv("VDL_SCAN_JIT_MIN_ROWS", "0")
    rel = fuzz_plans.single_table(seed)
    text = vlite.translate(catalog, rel)
    sf = 0.05
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    want = run_oracle(text, cols)
    got, stats = run_gpu(text, cols)
    assert_same(got, want)
    if stats["fused_scans"]:
        assert stats["shape"].startswith("jit:"), stats["shape"]
        monkeypatch.setenv("VDL_GENERIC_ONLY", "1")
        got_g, stats_g = run_gpu(text, cols)
        assert stats_g["shape"] == "generic"
        assert_same(got_g, want)


@pytest.mark.gpu
@pytest.mark.parametrize("query,colnames", [("q06.vdl", ["l_quantity", "l_extendedprice", "l_discount", "l_shipdate"]),
                                            ("q01.vdl", ["l_quantity", "l_extendedprice", "l_discount", "l_shipdate", "l_tax", "l_returnflag", "l_linestatus"])])
@pytest.mark.parametrize("narrow", [False, True])
def test_runtime_compiled_scan_on_tpch_shapes(catalog, query, colnames, narrow, monkeypatch):
    """Q6 / Q1 through the run-time compiled shape, with the reference's column widths and with the narrow storage format
    (every column that fits as int32): the same bits as the oracle either way."""
    import numpy as np
    from util import plan_text
    monkeypatch.setenv("VDL_SCAN_JIT_MIN_ROWS", "0")
    cols = host_columns(catalog, ["lineitem." + c for c in colnames], {"lineitem": 700_001})
    if narrow:
        cols = {k: (v.astype(np.int32) if v.min() >= -2**31 and v.max() < 2**31 else v) for k, v in cols.items()}
    want = run_oracle(plan_text(query), cols)
    got, stats = run_gpu(plan_text(query), cols)
    assert stats["shape"].startswith("jit:")
    assert_same(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(0, 40, 2))
def test_random_fk_join_plans_through_the_runtime_compiled_probe(catalog, seed, monkeypatch):
    """The probe descriptor printed as CUDA C and compiled at run time (forced for these small tables), fold and emit
    modes, against the oracle; the interpreter kernel on the same plans is test_random_fk_join_plans."""
    monkeypatch.setenv("VDL_PROBE_JIT_MIN_ROWS", "0")
    for rel, sf in ((fuzz_plans.join_query(seed, catalog), 0.02), (fuzz_plans.single_table(seed), 0.02)):
        text = vlite.translate(catalog, rel)
        rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
        cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
        got, stats = run_gpu(text, cols, fuse=True)
        assert_same(got, run_oracle(text, cols))


@pytest.mark.gpu
@pytest.mark.parametrize("q", ["q03", "q05", "q12", "q19", "q04", "q09", "q14", "q16"])
def test_tpch_join_plans_through_the_runtime_compiled_probe(catalog, q, monkeypatch):
    from util import plan_text, q19_columns
    monkeypatch.setenv("VDL_PROBE_JIT_MIN_ROWS", "0")
    sf = 0.05
    if q == "q19":
        text, cols = q19_columns(catalog, sf=sf)
    else:
        text = plan_text(q + ".vdl")
        rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
        cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    got, stats = run_gpu(text, cols, fuse=True)
    assert_same(got, run_oracle(text, cols))
