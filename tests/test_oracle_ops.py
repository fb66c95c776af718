"""Known-answer tests for every Voodoo op the oracle implements (hand-computed expectations).
The reference holds no per-op vectors (tests/Tests.hs:17-18), so these pin OUR decisions
 and are re-used verbatim as the CUDA per-op parity inputs."""
import numpy as np
import pytest

from oracle.oracle import Oracle, OracleError

I64 = np.int64


def run(plan, **cols):
    o = Oracle()
    for k, v in cols.items():
        o.bind("t." + k, np.asarray(v))
    return o.run(plan)


def binop_plan(op):
    return f"1,Load,t.a\n2,Project,val,Id 1,a\n3,Load,t.b\n4,Project,val,Id 3,b\n5,{op},val,Id 2,val,Id 4,val\n6,Project,out,Id 5,val\n7,MaterializeCompact,Id 6\n"


A = np.array([5, -7, 0, 2**62, -2**63, 9, 3], dtype=I64)
B = np.array([3, 2, 0, 4, -1, -2, 3], dtype=I64)


@pytest.mark.parametrize("op,expect", [
    ("Add", [8, -5, 0, 2**62 + 4, 2**63 - 1, 7, 6]),
    ("Subtract", [2, -9, 0, 2**62 - 4, -2**63 + 1, 11, 0]),
    ("Multiply", [15, -14, 0, 0, -2**63, -18, 9]),          # 2^62*4 wraps to 0; INT64_MIN*-1 wraps to itself
    ("Divide", [1, -3, 0, 2**60, -2**63, -4, 1]),           # truncation; x/0 := 0; INT64_MIN/-1 wraps
    ("Modulo", [2, -1, 0, 0, 0, 1, 0]),
    ("Greater", [1, 0, 0, 1, 0, 1, 0]),
    ("Equals", [0, 0, 1, 0, 0, 0, 1]),
    ("LogicalAnd", [1, 1, 0, 1, 1, 1, 1]),
    ("LogicalOr", [1, 1, 0, 1, 1, 1, 1]),
    ("BitwiseAnd", [1, 0, 0, 0, -2**63, 8, 3]),
    ("BitwiseOr", [7, -5, 0, 2**62 + 4, -1, -1, 3]),
    ("BitShift", [0, -2, 0, 2**58, 0, 36, 0]),              # +k arithmetic right; -k left (Vlite.hs:205-208)
])
def test_binary(op, expect):
    r = run(binop_plan(op), a=A, b=B)["out"]
    np.testing.assert_array_equal(r, np.array(expect, dtype=I64))


def test_int32_column_sign_extends():
    a = np.array([-5, 7], dtype=np.int32)
    b = np.array([2**40, 1], dtype=I64)
    r = run(binop_plan("Add"), a=a, b=b)["out"]
    np.testing.assert_array_equal(r, [2**40 - 5, 8])


def test_range_and_constants():
    plan = ("1,Load,t.a\n2,Project,val,Id 1,a\n3,RangeV,val,10,Id 2,3\n4,RangeC,val,-1,4,2\n"
            "5,Project,rv,Id 3,val\n6,MaterializeCompact,Id 5\n7,Project,rc,Id 4,val\n8,MaterializeCompact,Id 7\n")
    r = run(plan, a=np.zeros(3, I64))
    np.testing.assert_array_equal(r["rv"], [10, 13, 16])
    np.testing.assert_array_equal(r["rc"], [-1, 1, 3, 5])


SELECT = ("1,Load,t.p\n2,Project,val,Id 1,p\n3,RangeV,val,0,Id 2,1\n4,FoldSelect,val,Id 3,val,Id 2,val\n"
          "5,Load,t.x\n6,Project,val,Id 5,x\n7,Gather,Id 6,Id 4,val\n"
          "8,Project,pos,Id 4,val\n9,MaterializeCompact,Id 8\n10,Project,sel,Id 7,val\n11,MaterializeCompact,Id 10\n")


def test_fold_select_and_gather():
    r = run(SELECT, p=np.array([0, 2, 0, -1, 1, 0], I64), x=np.array([10, 11, 12, 13, 14, 15], I64))
    np.testing.assert_array_equal(r["pos"], [1, 3, 4])
    np.testing.assert_array_equal(r["sel"], [11, 13, 14])


def test_fold_select_empty_and_full():
    r = run(SELECT, p=np.zeros(4, I64), x=np.arange(4, dtype=I64))
    assert r["pos"].shape == (0,) and r["sel"].shape == (0,)
    r = run(SELECT, p=np.ones(4, I64), x=np.arange(4, dtype=I64))
    np.testing.assert_array_equal(r["pos"], [0, 1, 2, 3])


def test_gather_out_of_range_is_an_error():
    plan = "1,Load,t.x\n2,Project,val,Id 1,x\n3,Load,t.i\n4,Project,val,Id 3,i\n5,Gather,Id 2,Id 4,val\n6,MaterializeCompact,Id 5\n"
    with pytest.raises(OracleError):
        run(plan, x=np.arange(3, dtype=I64), i=np.array([0, 3], I64))


FOLD = ("1,Load,t.g\n2,Project,val,Id 1,g\n3,Load,t.x\n4,Project,val,Id 3,x\n5,{op},val,Id 2,val,Id 4,val\n"
        "6,Project,out,Id 5,val\n7,MaterializeCompact,Id 6\n")


@pytest.mark.parametrize("op,expect", [
    ("FoldSum", [3, 3, 15, 7]), ("FoldMin", [1, 3, 4, 7]), ("FoldMax", [2, 3, 6, 7]),
    ("FoldChoose", [1, 3, 4, 7]), ("FoldCount", [2, 1, 3, 1]),
])
def test_folds_by_runs(op, expect):
    # runs are maximal stretches of equal consecutive group values: group 5 appears twice -> two runs
    g = np.array([5, 5, 9, 2, 2, 2, 5], I64)
    x = np.array([1, 2, 3, 4, 5, 6, 7], I64)
    np.testing.assert_array_equal(run(FOLD.format(op=op), g=g, x=x)["out"], expect)


def test_fold_empty_input_gives_empty_output():
    r = run(FOLD.format(op="FoldSum"), g=np.zeros(0, I64), x=np.zeros(0, I64))
    assert r["out"].shape == (0,)


def test_fold_sum_wraps():
    r = run(FOLD.format(op="FoldSum"), g=np.zeros(2, I64), x=np.array([2**63 - 1, 1], I64))
    np.testing.assert_array_equal(r["out"], [-2**63])


def test_fold_many_threads_matches_single_thread():
    rng = np.random.default_rng(7)
    g = np.sort(rng.integers(0, 50, 100_000)).astype(I64)
    x = rng.integers(-10**12, 10**12, 100_000).astype(I64)
    o = Oracle(); o.bind("t.g", g); o.bind("t.x", x)
    for op in ("FoldSum", "FoldMin", "FoldMax", "FoldChoose", "FoldCount"):
        a = o.run(FOLD.format(op=op), threads=1)["out"]
        b = o.run(FOLD.format(op=op), threads=8)["out"]
        np.testing.assert_array_equal(a, b)
        assert len(a) == len(np.unique(g))


PART = ("1,Load,t.k\n2,Project,val,Id 1,k\n3,RangeC,val,0,4,1\n4,Partition,val,Id 2,val,Id 3,val\n"
        "5,RangeV,val,0,Id 2,1\n6,Scatter,Id 2,Id 5,val,Id 4,val\n"
        "7,Project,perm,Id 4,val\n8,MaterializeCompact,Id 7\n9,Project,sorted,Id 6,val\n10,MaterializeCompact,Id 9\n")


def test_partition_is_a_stable_sort_permutation():
    k = np.array([3, 1, 2, 1, 0, 3, 1], I64)
    r = run(PART, k=k)
    # destination of each row in the stable sort by key
    np.testing.assert_array_equal(r["perm"], [5, 1, 4, 2, 0, 6, 3])
    np.testing.assert_array_equal(r["sorted"], [0, 1, 1, 1, 2, 3, 3])


@pytest.mark.parametrize("domain", [7, 5000, 1 << 30])
def test_large_partition_in_parallel_is_the_stable_argsort(domain):
    """Above 65536 rows every radix pass runs on all threads (per-thread digit counts of contiguous slices): same permutation
    as numpy's stable sort, also over several 11-bit passes; keys that are already in order take the identity shortcut."""
    rng = np.random.default_rng(domain)
    k = rng.integers(0, domain, 300_001).astype(I64)
    plan = PART.replace("3,RangeC,val,0,4,1", f"3,RangeC,val,0,{domain},1")
    for keys in (k, np.sort(k)):
        r = run(plan, k=keys)
        order = np.argsort(keys, kind="stable")
        want = np.empty_like(order)
        want[order] = np.arange(len(keys))
        np.testing.assert_array_equal(r["perm"], want)
        np.testing.assert_array_equal(r["sorted"], keys[order])


def test_partition_clamps_to_pivot_range():
    # bucket = number of pivots below the value: values under the first pivot share bucket 0, above the last share bucket n
    r = run(PART, k=np.array([9, -5, 2, 100, 0], I64))
    np.testing.assert_array_equal(r["sorted"], [-5, 0, 2, 9, 100])


def test_scatter_length_is_the_index_space_of_the_positions():
    # positions produced by FoldSelect over a 6-row predicate index a 6-row space (G2): unwritten slots are 0
    plan = ("1,Load,t.p\n2,Project,val,Id 1,p\n3,RangeV,val,0,Id 2,1\n4,FoldSelect,val,Id 3,val,Id 2,val\n"
            "5,RangeV,val,1,Id 4,0\n6,RangeV,val,0,Id 5,1\n7,Scatter,Id 5,Id 6,val,Id 4,val\n"
            "8,Project,valid,Id 7,val\n9,MaterializeCompact,Id 8\n"
            "10,Scatter,Id 6,Id 6,val,Id 4,val\n11,Project,inv,Id 10,val\n12,MaterializeCompact,Id 11\n")
    r = run(plan, p=np.array([0, 1, 0, 0, 1, 1], I64))
    np.testing.assert_array_equal(r["valid"], [0, 1, 0, 0, 1, 1])
    np.testing.assert_array_equal(r["inv"], [0, 0, 0, 0, 1, 2])


def test_unsupported_ops_are_rejected_loudly():
    with pytest.raises(OracleError):
        run("1,Load,t.a\n2,Semisort,Id 1\n", a=np.zeros(1, I64))


def test_metadata_suffix_is_ignored():
    plan = binop_plan("Add").replace("\n", " ;; Metadata {databounds = (0,1)}\n")
    np.testing.assert_array_equal(run(plan, a=A, b=B)["out"], (A + B))


LIKE = ("1,Load,t.s\n2,Project,val,Id 1,s\n3,Load,t.s.heap\n4,Project,val,Id 3,s.heap\n"
        "5,Like,val,Id 2,val,Id 4,val,{pat}\n6,Project,out,Id 5,val\n7,MaterializeCompact,Id 6\n")
STRINGS = ["PROMO BRUSHED TIN", "ECONOMY ANODIZED BRASS", "", "forest green", "Customer slyly Complaints", "a,b", "%_"]


def like_columns(strings):
    heap, offs = bytearray(8), []
    for s in strings:
        offs.append(len(heap))
        b = s.encode() + b"\0"
        heap += b + bytes(-len(b) % 8)
    return np.array(offs, dtype=I64), np.frombuffer(bytes(heap), dtype=np.uint8)


@pytest.mark.parametrize("pat,expect", [
    ("PROMO%", [1, 0, 0, 0, 0, 0, 0]), ("%BRASS", [0, 1, 0, 0, 0, 0, 0]), ("%green%", [0, 0, 0, 1, 0, 0, 0]),
    ("%Customer%Complaints%", [0, 0, 0, 0, 1, 0, 0]), ("%", [1, 1, 1, 1, 1, 1, 1]), ("", [0, 0, 1, 0, 0, 0, 0]),
    ("_", [0, 0, 0, 0, 0, 0, 0]), ("__", [0, 0, 0, 0, 0, 0, 1]), ("a,b", [0, 0, 0, 0, 0, 1, 0]), ("%r%r%", [0, 0, 0, 1, 0, 0, 0]), ("%R%S%", [1, 1, 0, 0, 0, 0, 0]),
    ("forest gree_", [0, 0, 0, 1, 0, 0, 0]), ("%TIN%TIN", [0, 0, 0, 0, 0, 0, 0]), ("%N", [1, 0, 0, 0, 0, 0, 0]),
])
def test_like_over_a_string_heap(pat, expect):
    """Vdl.hs:244-247, 444-447: data = byte offsets into the column's heap; SQL LIKE, no escape, case-sensitive.  (In SQL
    `%` and `_` in the DATA are ordinary characters: "%_" matches "__" but not "_".)"""
    offs, heap = like_columns(STRINGS)
    o = Oracle()
    o.bind("t.s", offs)
    o.bind("t.s.heap", heap)
    np.testing.assert_array_equal(o.run(LIKE.format(pat=pat))["out"], expect)


def test_like_offset_outside_the_heap_is_an_error():
    offs, heap = like_columns(STRINGS)
    o = Oracle()
    o.bind("t.s", np.array([0, len(heap)], dtype=I64))
    o.bind("t.s.heap", heap)
    with pytest.raises(OracleError, match="outside the heap"):
        o.run(LIKE.format(pat="%"))


def test_a_heap_is_only_a_like_dictionary():
    offs, heap = like_columns(STRINGS)
    o = Oracle()
    o.bind("t.s", offs)
    o.bind("t.s.heap", heap)
    with pytest.raises(OracleError, match="string heap"):
        o.run("1,Load,t.s.heap\n2,Project,val,Id 1,s.heap\n3,MaterializeCompact,Id 2\n")
