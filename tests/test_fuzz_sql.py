"""Random queries (tests/fuzz_plans.py): translator restatement + CPU oracle against the direct numpy evaluation of the
relational IR (tests/ir_eval.py) -- a second opinion on the oracle that shares neither the Voodoo lowering nor the op
semantics with it.  CPU only; the GPU suite compares the CUDA paths with the oracle on the same plans."""
import numpy as np
import pytest

import fuzz_plans
import ir_eval
from mplan2vdl_b200 import synth, tpch, vlite
from util import host_columns, run_oracle

SF = 0.004


def check(catalog, query, sf=SF):
    text = vlite.translate(catalog, query)
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    got = list(run_oracle(text, cols).values())
    want = ir_eval.evaluate(cols, query)
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        np.testing.assert_array_equal(np.asarray(g, dtype=np.int64), w, err_msg=f"output column {k}")
    return want


@pytest.mark.parametrize("seed", range(40))
def test_single_table_queries_agree_with_the_direct_evaluation(catalog, seed):
    check(catalog, fuzz_plans.single_table(seed))


@pytest.mark.parametrize("seed", range(40))
def test_fk_join_queries_agree_with_the_direct_evaluation(catalog, seed):
    check(catalog, fuzz_plans.join_query(seed, catalog))


def test_the_fuzzer_draws_non_trivial_queries(catalog):
    sizes = [len(check(catalog, fuzz_plans.join_query(s, catalog))[0]) for s in range(12)]
    assert sum(1 for n in sizes if n > 0) >= 6 and max(sizes) > 1
