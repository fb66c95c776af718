"""The C++ planner on a host without a GPU (vdl_plan_explain: parse, CSE, both fusion passes, map clusters, the tail check;
binding and launching happen at run time): what it makes of the 15 checked-in programs and of the other graph shapes the
translator prints.  These are the host-logic tests of csrc/vdl_plan.cu / vdl_plan_join.inc; the GPU tests check the results."""
import pytest

from mplan2vdl_b200 import tpch_queries, vlite
from mplan2vdl_b200.executor import explain
from mplan2vdl_b200.lib import VdlError
from util import plan_text


def scans(d):
    return [(g["table"], g["columns"], g["predicates"], g["key_parts"], g["domain"], g["folds"], g["posts"]) for g in d["fused_scans"]]


def test_q6_is_one_fused_scan():
    d = explain(plan_text("q06.vdl"))
    assert d["statements"] == 42 and d["nodes"] < 42                    # README.md:40-52: 42 lines; Project / Materialize are aliases
    # 4 columns; shipdate's two comparisons are ONE range, discount's two another, quantity the third; no key; one sum
    assert scans(d) == [("lineitem", 4, 3, 0, 1, 1, 0)]
    assert not d["probe_folds"] and not d["probe_emits"] and not d["map_clusters"] and d["folds_op_at_a_time"] == 0


def test_q1_is_one_fused_scan_with_a_32_slot_key_and_three_averages_as_post_ops():
    d = explain(plan_text("q01.vdl"))
    assert scans(d) == [("lineitem", 7, 1, 2, 32, 8, 3)]                # 2 key parts (returnflag, linestatus), 5 bits; 8 distinct folds
    assert d["folds_op_at_a_time"] == 0 and not d["mergeable_tail"]


@pytest.mark.parametrize("q,leaves,preds,domain,folds", [("q05", 12, 3, 128, 2), ("q12", 6, 4, 32, 3)])
def test_fk_join_plans_with_a_small_key_are_one_probe_fold_pass(q, leaves, preds, domain, folds):
    d = explain(plan_text(q + ".vdl"))
    (g,) = d["probe_folds"]
    assert (g["table"], g["leaves"], g["predicates"], g["domain"], g["folds"]) == ("lineitem", leaves, preds, domain, folds)
    assert not d["fused_scans"] and not d["probe_emits"] and d["folds_op_at_a_time"] == 0


def test_q3_emits_its_survivors_once_and_its_tail_merges_across_shards():
    d = explain(plan_text("q03.vdl"))
    assert [(g["table"], g["vectors"]) for g in d["probe_emits"]] == [("lineitem", 7)]
    assert d["folds_op_at_a_time"] == 4 and d["mergeable_tail"] and len(d["map_clusters"]) == 1


def test_every_checked_in_plan_is_planned_and_semijoin_plans_walk_a_dimension_table_too():
    tables = {}
    for q in ("q01", "q03", "q04", "q05", "q06", "q09", "q10", "q11", "q12", "q14", "q15", "q16", "q18", "q19", "q20"):
        d = explain(plan_text(q + ".vdl"))
        assert d["nodes"] <= d["statements"]
        tables[q] = [g["table"] for g in d["probe_emits"]]
    # the plans dist.check_shardable refuses for N > 1: an emit pass over a replicated table, or more than one pass
    assert tables["q04"] == ["orders", "lineitem"] and tables["q20"] == ["lineitem", "nation"] and tables["q11"] == ["partsupp"]
    assert tables["q03"] == tables["q19"] == ["lineitem"]


def test_without_the_fusion_flag_everything_stays_op_at_a_time():
    d = explain(plan_text("q01.vdl"), fuse=False)
    assert not d["fused_scans"] and not d["probe_folds"] and not d["probe_emits"] and not d["map_clusters"]
    assert d["folds_op_at_a_time"] == 8                                 # 10 aggregates, 8 distinct Folds after CSE (the AVGs share sums and the count)


@pytest.mark.parametrize("strategy", ["shuffle", ("hierarchical", 0), ("hierarchical", 2)])
def test_other_aggregation_strategies_plan_like_the_serial_one(catalog, strategy):
    """Shuffle is an alias; a level-1 Fold that provably refines the groups collapses at parse time."""
    for q in ("q06", "q01"):
        rel = tpch_queries.QUERIES[q](catalog)
        assert scans(explain(vlite.translate(catalog, rel, strategy))) == scans(explain(vlite.translate(catalog, rel)))


def test_an_unprovable_two_level_fold_is_left_alone(catalog):
    """Q1 with a grain above the 32 slots of the reference's metadata: keys OR-ed without a shift, nothing provable, both
    levels stay Folds (and vdl_op_fold evaluates level 2 literally)."""
    d = explain(vlite.translate(catalog, tpch_queries.QUERIES["q01"](catalog), ("hierarchical", 13)))
    assert not d["fused_scans"] and d["folds_op_at_a_time"] == 2 * 8     # the 8 distinct aggregates (AVG = sum / count), twice
    d6 = explain(vlite.translate(catalog, tpch_queries.QUERIES["q06"](catalog), ("hierarchical", 13)))
    assert scans(d6) == [("lineitem", 4, 3, 0, 1, 1, 0)]                 # constant groups: any level 1 refines them


def test_goffset_keeps_a_single_column_key_fusable(catalog):
    base = explain(vlite.translate(catalog, tpch_queries.QUERIES["q05"](catalog)))
    off = explain(vlite.translate(catalog, tpch_queries.QUERIES["q05"](catalog), goffset=3))
    assert len(off["probe_folds"]) == len(base["probe_folds"]) == 1      # key = (a + b * col) & mask: still an affine part
    assert off["probe_folds"][0]["domain"] >= base["probe_folds"][0]["domain"]


def test_rejected_programs_report_the_planner_message():
    for text, what in (("1,Load,t.a\n2,Semisort,Id 1\n", "Semisort"), ("1,Load,t.a\n3,Project,x,Id 1,val\n", "1,2,3"),
                       ("1,Load,t.a\n2,Frobnicate,val,Id 1,val,Id 1,val\n", "unknown op"), ("1,Load,t.a\n", "MaterializeCompact")):
        with pytest.raises(VdlError, match=what):
            explain(text)


def test_random_queries_are_planned_and_mostly_fused(catalog):
    """The 80 random shapes of tests/fuzz_plans.py (the GPU fuzz tests run them for results): every program is accepted and
    claimed by a fused scan or a probe pass (the probe also takes single-table plans whose predicates are IN sets or compare two
    columns); at most two Folds of a plan stay op-at-a-time."""
    import fuzz_plans
    pure_scans = one_probe_fold = 0
    for seed in range(40):
        d = explain(vlite.translate(catalog, fuzz_plans.single_table(seed)))
        assert d["fused_scans"] or d["probe_folds"] or d["probe_emits"], seed
        assert d["folds_op_at_a_time"] <= 2, seed
        pure_scans += len(d["fused_scans"]) == 1 and not d["probe_folds"] and not d["probe_emits"] and d["folds_op_at_a_time"] == 0
        j = explain(vlite.translate(catalog, fuzz_plans.join_query(seed, catalog)))
        assert j["nodes"] <= j["statements"] and (j["probe_folds"] or j["probe_emits"]), seed
        one_probe_fold += len(j["probe_folds"]) == 1 and not j["probe_emits"] and j["folds_op_at_a_time"] == 0
    assert pure_scans >= 20 and one_probe_fold >= 30
