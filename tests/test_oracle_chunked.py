"""The chunked oracle (oracle/chunked.py: fact table generated and consumed a row range at a time, Folds merged by key,
the statements above the Folds run once on the merged vectors) equals the whole-table oracle bit for bit -- it is what
checks bench.py's result over all 600,037,902 rows of SF100 in the same run."""
import numpy as np
import pytest

from mplan2vdl_b200 import synth, tpch
from oracle import chunked
from util import assert_same, host_columns, plan_text, run_oracle


@pytest.mark.parametrize("q,sf,chunk", [("q06", 0.01, 7_001), ("q06", 0.01, 10**9), ("q01", 0.01, 9_973), ("q01", 0.002, 1_000),
                                        ("q05", 0.01, 20_011), ("q03", 0.01, 13_337), ("q12", 0.01, 30_000), ("q19", 0.01, 25_000)])
def test_chunked_equals_whole(catalog, q, sf, chunk):
    text = plan_text(q + ".vdl")
    rows = {t: synth.table_rows(catalog, t, sf) for t in catalog.tables}
    cols = host_columns(catalog, tpch.plan_columns(text), rows, sf=sf)
    want = run_oracle(text, cols)
    got, co = chunked.run_chunked(text, catalog, sf, chunk_rows=chunk)
    assert co.rows == rows["lineitem"]
    assert_same(got, want)


def test_split_plan_cuts_at_the_folds():
    chunk_plan, tail_plan, cut, keys = chunked.split_plan(plan_text("q01.vdl"), "lineitem")
    assert len(cut) == 10 and len(keys) == 1                   # Q1: ten Folds (2 FoldChoose, 7 sums incl. the AVG numerators, 1 count) over one sorted key
    assert "Divide" in tail_plan and "Divide" not in chunk_plan  # AVG's Divide runs once, above the merge (Vlite.hs:1038-1041)
    assert "lineitem." not in tail_plan


def test_empty_chunks_and_empty_selection(catalog):
    text = plan_text("q06.vdl")
    co = chunked.ChunkedOracle(text)
    cols = host_columns(catalog, tpch.plan_columns(text), {"lineitem": 5000})
    cols["lineitem.l_quantity"][:] = 5000                      # nothing passes l_quantity < 24
    co.add_chunk(cols)
    co.add_chunk({k: v[:0] for k, v in cols.items()})
    out = co.finish()
    assert out["revenue"].shape == (0,)                        # G14: no run, no row


def test_unsorted_groups_are_refused():
    plan = "\n".join(["1,Load,lineitem.a", "2,Load,lineitem.b", "3,FoldSum,val,Id 1,val,Id 2,val", "4,Project,s,Id 3,val", "5,MaterializeCompact,Id 4"]) + "\n"
    co = chunked.ChunkedOracle(plan)
    with pytest.raises(chunked.NotChunkable):
        co.add_chunk({"lineitem.a": np.array([1, 2, 1], dtype=np.int64), "lineitem.b": np.array([1, 1, 1], dtype=np.int64)})
