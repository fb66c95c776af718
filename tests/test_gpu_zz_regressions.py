"""Regressions found by ad-hoc sweeps (tools/gpu/fuzz_sweep.py).  Kept in a file that sorts last among the GPU tests."""
import pytest

import fuzz_plans
from test_gpu_fuzz import _run


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [128, 183])
def test_folds_that_read_no_fact_column_stay_per_op(catalog, seed):
    """Found by tools/gpu/fuzz_sweep.py: a bare COUNT(*) (seed 183) and a COUNT under a selection that folds to "never"
    (seed 128: 0.10 < l_discount <= 0.10) give the probe pass no column of the fact table to walk; they must not be
    claimed by it (the plan used to fail at run time with "probe: no fact column among the leaves")."""
    _, stats = _run(catalog, fuzz_plans.single_table(seed))
    assert stats["probe_folds"] == 0
