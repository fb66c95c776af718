"""The N>1 path on CPU: world_size-2 gloo.  Checks the row-range sharding, the all-gather plumbing of
mplan2vdl_b200.dist and the merge rule of the partial tables (restated in numpy here; on GPUs the merge is the
finalize kernel, covered by tests/test_gpu_sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mplan2vdl_b200 import tpch
from mplan2vdl_b200.dist import gather_partial_tables
from mplan2vdl_b200.meta import builtin_catalog
from oracle import sqlref
from util import Q1_COLS, host_columns, plan_text, run_oracle

ROWS = 20_000
DOMAIN = 32


def test_shard_ranges_partition_the_table():
    for rows in (0, 1, 4095, 4096, 4097, 6_001_215, 600_037_902):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                start, n = tpch.shard_range(rows, r, world)
                assert start == min(rows, covered) and n >= 0
                assert start % tpch.SHARD_ALIGN == 0 or n == 0
                covered += n
            assert covered == rows


def q1_partial_table(cols, row_base):
    """[sum_qty, sum_ep, sum_disc_price, sum_charge, sum_disc, count, first_row, choose_rf, choose_ls] x DOMAIN."""
    g = {k.split(".")[1]: v.astype(np.int64) for k, v in cols.items()}
    m = g["l_shipdate"] <= 729999
    key = ((((g["l_returnflag"] >> 3) - 2) << 2) | ((g["l_linestatus"] >> 3) - 2)) & 31
    t = np.zeros((9, DOMAIN), dtype=np.int64)
    t[6, :] = np.iinfo(np.int64).max
    dp = g["l_extendedprice"] * (100 - g["l_discount"])
    vals = [g["l_quantity"], g["l_extendedprice"], dp, dp * (100 + g["l_tax"]), g["l_discount"], np.ones_like(key)]
    for k in np.unique(key[m]):
        s = m & (key == k)
        for j, v in enumerate(vals):
            t[j, k] = v[s].sum(dtype=np.int64)
        first = int(np.argmax(s))
        t[6, k] = row_base + first
        t[7, k], t[8, k] = g["l_returnflag"][first], g["l_linestatus"][first]
    return t


def merge(tables):
    """The finalize rule: sums add, first row is the minimum, chooses come from the rank holding the first row."""
    tables = np.stack(tables)
    cnt = tables[:, 5, :].sum(0)
    keys = np.nonzero(cnt > 0)[0]
    best = tables[:, 6, :].argmin(0)
    out = {"l_returnflag__lineitem__l_returnflag": [tables[best[k], 7, k] for k in keys],
           "l_linestatus__lineitem__l_linestatus": [tables[best[k], 8, k] for k in keys]}
    s = tables[:, :6, :].sum(0)
    for name, j in (("sum_qty", 0), ("sum_base_price", 1), ("sum_disc_price", 2), ("sum_charge", 3)):
        out[name] = s[j, keys]
    for name, j in (("avg_qty", 0), ("avg_price", 1), ("avg_disc", 4)):
        out[name] = [sqlref._tdiv(s[j, k], s[5, k]) for k in keys]
    out["count_order"] = s[5, keys]
    return {k: np.asarray(v, dtype=np.int64) for k, v in out.items()}


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cat = builtin_catalog()
    names = ["lineitem." + c for c in Q1_COLS]
    start, n = tpch.shard_range(ROWS, rank, world)
    cols = host_columns(cat, names, {"lineitem": n}, row_offset=start)
    local = torch.from_numpy(q1_partial_table(cols, start).reshape(-1))
    gathered = gather_partial_tables(local, world).view(world, 9, DOMAIN).numpy()
    got = merge(list(gathered))
    if rank == 0:
        q.put({k: v.tolist() for k, v in got.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_partials_merge_to_the_whole_table_answer():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cat = builtin_catalog()
    # shard boundaries are multiples of SHARD_ALIGN, so 20_000 rows split 12288 + 7712
    assert tpch.shard_range(ROWS, 1, 2) == (12288, 7712)
    whole = run_oracle(plan_text("q01.vdl"), host_columns(cat, ["lineitem." + c for c in Q1_COLS], {"lineitem": ROWS}))
    assert list(got) == list(whole)
    for k in whole:
        np.testing.assert_array_equal(np.asarray(got[k]), whole[k], err_msg=k)


def survivors_worker(rank, world, port, q):
    from mplan2vdl_b200.dist import gather_survivors
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for lens in ([3, 5], [0, 4], [2, 0], [0, 0]):                      # uneven and empty shards
        local = torch.arange(lens[rank], dtype=torch.int64) + 100 * rank
        out.append(gather_survivors(local, world).tolist())
    if rank == 1:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_survivor_exchange_concatenates_in_rank_order():
    """The exchange a sharded emit plan (Q3) needs: vectors of different lengths, concatenated in rank order."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=survivors_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [[0, 1, 2, 100, 101, 102, 103, 104], [100, 101, 102, 103], [0, 1], []]
