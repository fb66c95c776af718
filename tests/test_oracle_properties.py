"""Property tests of the CPU oracle against direct numpy statements of each op (hypothesis): the reference pins no vectors
(tests/Tests.hs:17-18), so besides the hand-computed cases of test_oracle_ops.py the oracle is held to the algebra of its own
definitions on random inputs -- Fold by runs vs reduceat, Partition vs numpy's stable argsort, Scatter/Gather inverses,
FoldSelect vs flatnonzero, the two-level Fold vs the literal "group the level-1 results by the head's group"."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle.oracle import Oracle

I64 = np.int64
small = st.integers(min_value=-50, max_value=50)
vec = st.lists(small, min_size=0, max_size=200)


def run(plan, **cols):
    o = Oracle()
    for k, v in cols.items():
        o.bind("t." + k, np.asarray(v, dtype=I64))
    return o.run(plan)


FOLD = "1,Load,t.g\n2,Load,t.d\n3,{op},val,Id 1,val,Id 2,val\n4,Project,out,Id 3,val\n5,MaterializeCompact,Id 4\n"


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 3), small), min_size=0, max_size=200))
def test_folds_by_runs_are_reduceat_over_the_run_heads(rows):
    g = np.array([r[0] for r in rows], dtype=I64)
    d = np.array([r[1] for r in rows], dtype=I64)
    heads = np.flatnonzero(np.r_[True, g[1:] != g[:-1]]) if len(g) else np.zeros(0, dtype=np.intp)
    want = {"FoldSum": np.add.reduceat(d, heads) if len(g) else [], "FoldMin": np.minimum.reduceat(d, heads) if len(g) else [],
            "FoldMax": np.maximum.reduceat(d, heads) if len(g) else [], "FoldChoose": d[heads] if len(g) else [],
            "FoldCount": np.diff(np.r_[heads, len(g)]) if len(g) else []}
    for op, w in want.items():
        np.testing.assert_array_equal(run(FOLD.format(op=op), g=g, d=d)["out"], np.asarray(w, dtype=I64), err_msg=op)


PART = "1,Load,t.k\n2,RangeC,val,{lo},{cnt},1\n3,Partition,val,Id 1,val,Id 2,val\n4,RangeV,val,0,Id 1,1\n5,Scatter,Id 1,Id 4,val,Id 3,val\n" \
       "6,Project,perm,Id 3,val\n7,MaterializeCompact,Id 6\n8,Project,sorted,Id 5,val\n9,MaterializeCompact,Id 8\n"


@settings(max_examples=150, deadline=None)
@given(vec, st.integers(-20, 20), st.integers(1, 40))
def test_partition_is_the_inverse_of_the_stable_argsort_of_the_bucket(keys, lo, cnt):
    k = np.array(keys, dtype=I64)
    r = run(PART.format(lo=lo, cnt=cnt), k=k)
    bucket = np.clip(k - lo, 0, cnt)                      # number of pivots lo, lo+1, ... strictly below the value (G3)
    order = np.argsort(bucket, kind="stable")
    want = np.empty(len(k), dtype=I64)
    want[order] = np.arange(len(k))
    np.testing.assert_array_equal(r["perm"], want)
    np.testing.assert_array_equal(r["sorted"], k[order])  # Scatter by the permutation sorts the vector


SEL = "1,Load,t.p\n2,Load,t.x\n3,RangeV,val,0,Id 1,1\n4,FoldSelect,val,Id 3,val,Id 1,val\n5,Gather,Id 2,Id 4,val\n" \
      "6,Project,idx,Id 4,val\n7,MaterializeCompact,Id 6\n8,Project,kept,Id 5,val\n9,MaterializeCompact,Id 8\n"


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.integers(-1, 2), small), min_size=0, max_size=200))
def test_fold_select_is_flatnonzero_and_gather_follows_it(rows):
    p = np.array([r[0] for r in rows], dtype=I64)
    x = np.array([r[1] for r in rows], dtype=I64)
    r = run(SEL, p=p, x=x)
    np.testing.assert_array_equal(r["idx"], np.flatnonzero(p != 0))
    np.testing.assert_array_equal(r["kept"], x[p != 0])


TWO = "1,Load,t.g\n2,Load,t.h\n3,Load,t.d\n4,{op},val,Id 2,val,Id 3,val\n5,{op},val,Id 1,val,Id 4,val\n6,Project,out,Id 5,val\n7,MaterializeCompact,Id 6\n"


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 2), st.integers(0, 3), small), min_size=1, max_size=120))
def test_two_level_fold_groups_the_level_one_results_by_the_group_at_each_head(rows):
    """make2LevelFold (Vlite.hs:1181-1192) in the dense model, for ANY level-1 groups h (refining g or not)."""
    g = np.array([r[0] for r in rows], dtype=I64)
    h = np.array([r[1] for r in rows], dtype=I64)
    d = np.array([r[2] for r in rows], dtype=I64)
    heads1 = np.flatnonzero(np.r_[True, h[1:] != h[:-1]])
    level1 = np.add.reduceat(d, heads1)
    eff = g[heads1]
    heads2 = np.flatnonzero(np.r_[True, eff[1:] != eff[:-1]])
    np.testing.assert_array_equal(run(TWO.format(op="FoldSum"), g=g, h=h, d=d)["out"], np.add.reduceat(level1, heads2))
    m1 = np.maximum.reduceat(d, heads1)
    np.testing.assert_array_equal(run(TWO.format(op="FoldMax"), g=g, h=h, d=d)["out"], np.maximum.reduceat(m1, heads2))


CROSS = "1,Load,t.a\n2,Load,t.b\n3,CrossProductOuter,Id 1,Id 2\n4,CrossProductInner,Id 1,Id 2\n5,Gather,Id 1,Id 3,val\n6,Gather,Id 2,Id 4,val\n" \
        "7,Project,l,Id 5,val\n8,MaterializeCompact,Id 7\n9,Project,r,Id 6,val\n10,MaterializeCompact,Id 9\n"


@settings(max_examples=60, deadline=None)
@given(st.lists(small, min_size=0, max_size=12), st.lists(small, min_size=0, max_size=12))
def test_cross_product_pairs_are_left_major(a, b):
    r = run(CROSS, a=a, b=b)
    np.testing.assert_array_equal(r["l"], np.repeat(np.asarray(a, dtype=I64), len(b)))
    np.testing.assert_array_equal(r["r"], np.tile(np.asarray(b, dtype=I64), len(a)))
