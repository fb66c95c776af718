"""Oracle (plan interpreter) against the independent SQL-level numpy evaluation, Q6 and Q1."""
import numpy as np
import pytest

from oracle import sqlref
from util import Q1_COLS, Q6_COLS, assert_same, host_columns, plan_text, run_oracle


@pytest.mark.parametrize("rows", [1, 17, 4096, 250_001])
def test_q6_oracle_matches_sql(catalog, rows):
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": rows})
    assert_same(run_oracle(plan_text("q06.vdl"), cols), sqlref.q6(cols))


def test_q6_empty_selection_yields_empty_output(catalog):
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": 1000})
    cols["lineitem.l_quantity"][:] = 5000          # no row passes l_quantity < 24.00
    r = run_oracle(plan_text("q06.vdl"), cols)
    assert r["revenue"].shape == (0,)
    assert_same(r, sqlref.q6(cols))


def test_q6_selectivity_is_as_designed(catalog):
    cols = host_columns(catalog, ["lineitem." + c for c in Q6_COLS], {"lineitem": 1_000_000})
    sd, d, q = cols["lineitem.l_shipdate"], cols["lineitem.l_discount"], cols["lineitem.l_quantity"]
    sel = ((sd >= 728294) & (sd < 728659) & (d >= 5) & (d <= 7) & (q < 2400)).mean()
    assert abs(sel - 365 / 2526 * 3 / 11 * 575 / 1226) < 2e-3     # ~1.8 %


@pytest.mark.parametrize("rows", [1, 33, 100_003])
def test_q1_oracle_matches_sql(catalog, rows):
    cols = host_columns(catalog, ["lineitem." + c for c in Q1_COLS], {"lineitem": rows})
    got = run_oracle(plan_text("q01.vdl"), cols)
    assert_same(got, sqlref.q1(cols))
    if rows > 1000:
        assert len(got["count_order"]) == 6            # the recipe populates 6 of the 32 key slots


def test_q1_threads_do_not_change_results(catalog):
    cols = host_columns(catalog, ["lineitem." + c for c in Q1_COLS], {"lineitem": 50_000})
    assert_same(run_oracle(plan_text("q01.vdl"), cols, threads=1), run_oracle(plan_text("q01.vdl"), cols, threads=8))


@pytest.mark.parametrize("q", ["q01", "q03", "q05", "q06", "q12", "q19"])
def test_oracle_reproduces_the_committed_answers(catalog, q):
    """The plan interpreter against tests/golden/tpch_sf0.01_answers.json (the SQL-level evaluation, committed)."""
    from util import golden_case
    text, cols, want = golden_case(catalog, q)
    assert_same(run_oracle(text, cols), want)
