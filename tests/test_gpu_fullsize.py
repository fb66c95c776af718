"""Parity at the BASELINE configs themselves (VERDICT r1 "What's weak" #2): the CUDA path over columns generated in HBM
at SF10 / SF100 against the CPU oracle over the same rows, the fact table generated and consumed on the host a row range
at a time (oracle/chunked.py).  This is where the 32-bit accumulator proofs of the register-slot kernel (RK_N32 /
RK_MADW: they depend on rows per thread) and the peer / all-gather merges are exercised at real sizes."""
import numpy as np
import pytest

from mplan2vdl_b200 import synth, tpch
from oracle import chunked
from util import assert_same, plan_text

pytestmark = pytest.mark.gpu


def _gpu(catalog, text, sf):
    from mplan2vdl_b200.executor import Context
    ctx = Context(0)
    tpch.load_synthetic(ctx, catalog, tpch.plan_columns(text), sf)
    plan = ctx.plan(text)
    got = plan.run()
    stats = plan.stats()
    stats["shape"] = plan.shape(0) if stats["fused_scans"] else None
    plan.close()
    ctx.close()
    return got, stats


@pytest.mark.parametrize("q", ["q01", "q06", "q03", "q05", "q12", "q19"])
def test_sf10_against_the_oracle(catalog, q):
    """BASELINE configs 2 and 3: 59,986,052 lineitem rows, 15 M orders, 1.5 M customers."""
    text = plan_text(q + ".vdl")
    got, stats = _gpu(catalog, text, 10)
    want, co = chunked.run_chunked(text, catalog, 10, chunk_rows=20_000_000)
    assert co.rows == 59_986_052
    assert_same(got, want)
    assert stats["fused_scans"] + stats["probe_folds"] + stats["probe_emits"] >= 1
    if q == "q01":
        assert stats["shape"].startswith("jit:") and len(want["count_order"]) == 6     # the run-time compiled shape of this descriptor


def test_q6_sf100_against_the_oracle(catalog):
    """BASELINE config 4 / the north-star target itself: all 600,037,902 rows, bit-exact."""
    text = plan_text("q06.vdl")
    got, stats = _gpu(catalog, text, 100)
    want, co = chunked.run_chunked(text, catalog, 100, chunk_rows=40_000_000)
    assert co.rows == 600_037_902 and stats["fused_scans"] == 1
    assert_same(got, want)


def test_q1_sf100_two_shards_against_the_oracle(catalog):
    """Q1 at SF100 as two row-range shards (run one after the other on this GPU, partial tables concatenated the way the
    all-gather lays them out, merged by the finalize kernel): 300 M rows per shard = the rows-per-thread regime the
    RK_N32 / RK_MADW proofs must cover.  sum_charge exceeds 2^63 at this size and wraps -- on both sides (G5)."""
    import torch
    from mplan2vdl_b200.dist import DeviceView
    from mplan2vdl_b200.executor import Context
    text, sf, world = plan_text("q01.vdl"), 100, 2
    names = tpch.plan_columns(text)
    ctx = Context(0)
    tables, plan = [], None
    for rank in range(world):
        info = tpch.load_synthetic(ctx, catalog, names, sf, rank=rank, world=world)
        plan = ctx.plan(text)
        plan.set_row_base(info["row_base"])
        plan.run_local()
        assert plan.shape(0).startswith("jit:")
        ptr, cnt = plan.partials(0)
        ctx.synchronize()
        tables.append(torch.as_tensor(DeviceView(ptr, cnt), device="cuda:0").clone())
        if rank < world - 1:
            plan.close()
            for n in names:
                ctx.drop_column(n)
    gathered = torch.cat(tables)
    torch.cuda.synchronize()
    got = plan.finish([gathered.data_ptr()], world)
    plan.close()
    ctx.close()
    want, co = chunked.run_chunked(text, catalog, sf, chunk_rows=30_000_000)
    assert co.rows == 600_037_902
    assert_same(got, want)
