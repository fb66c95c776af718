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


class _FakeCtx:
    """Stands for executor.Context in the peer-setup test: buffers are integers, `fail` picks what breaks on which rank."""
    device = 0

    def __init__(self, rank, fail):
        self.rank, self.fail, self.freed, self.closed, self.n = rank, fail, [], [], 0

    def ipc_alloc(self, nbytes):
        self.n += 1
        if self.fail == ("alloc", self.rank, self.n):
            raise RuntimeError("out of memory")
        return 1000 * (self.rank + 1) + self.n

    def ipc_export(self, ptr):
        return b"h%d" % ptr

    def ipc_open(self, handle):
        if self.fail[0] == "open" and self.fail[1] == self.rank and handle.endswith(b"%d" % self.fail[2]):
            raise RuntimeError("peer access denied")
        return int(handle[1:]) + 500000

    def ipc_close(self, ptr):
        self.closed.append(ptr)

    def ipc_free(self, ptr):
        self.freed.append(ptr)


class _FakePlan:
    num_partials, num_emits = 3, 0

    def __init__(self):
        self.peers = {}

    def set_row_base(self, b):
        pass

    def exchange_bytes(self, i, world):
        return 4096

    def set_peers(self, i, rank, world, ptrs):
        self.peers[i] = list(ptrs)


def peer_setup_worker(rank, world, port, q):
    from mplan2vdl_b200.dist import ShardedPlan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    # nothing fails; rank 1 cannot open rank 0's SECOND buffer; rank 0 cannot allocate its second buffer
    for fail in (("none", -1, 0), ("open", 1, 2), ("alloc", 0, 2)):
        ctx, plan = _FakeCtx(rank, fail), _FakePlan()
        sp = ShardedPlan.__new__(ShardedPlan)
        sp.ctx, sp.plan, sp.rank, sp.world, sp.group = ctx, plan, rank, world, None
        sp._mine, sp._opened, sp.peer_fallback = [], [], None
        ok = sp._setup_peers()
        res.append((ok, len(plan.peers), sorted(ctx.freed), sorted(ctx.closed), sp.peer_fallback))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_setup_stays_collective_when_one_rank_fails():
    """an ipc_open / ipc_alloc failure on one rank must not leave the other rank in a mismatched collective
    (hang); both fall back together and release what they had mapped."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=peer_setup_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        ok0, npeers0, freed0, closed0, why0 = got[rank][0]
        assert ok0 and npeers0 == 3 and not freed0 and not closed0 and why0 is None
        for case in (1, 2):
            ok, npeers, freed, closed, why = got[rank][case]
            assert not ok and npeers == 0 and why
            assert len(freed) >= 1                      # its own buffers were freed again
    assert len(got[0][1][3]) >= 1                       # rank 0 had opened rank 1's first buffer: closed again


class _FakeTailPlan:
    """Stands for executor.Plan in the sharded-tail test: `execute` folds this rank's shard of a sorted key vector in numpy,
    the tail_* methods do what vdl_plan_tail_boundary / vdl_plan_tail_apply do in the library."""
    num_partials, num_emits = 0, 1
    OPS = [3, 0, 1, 2]            # FoldChoose, FoldSum, FoldMin, FoldMax

    def __init__(self, keys, vals):
        self.keys, self.vals, self.outs = keys, vals, None

    def tail_info(self):
        return list(self.OPS)

    def tail_enable(self, on):
        self.on = on

    def execute(self):
        k, v = self.keys, self.vals
        if len(k) == 0:
            self.outs = [np.zeros(0, np.int64) for _ in self.OPS]
            return
        heads = np.flatnonzero(np.r_[True, k[1:] != k[:-1]])
        self.outs = [v[heads].copy(), np.add.reduceat(v, heads), np.minimum.reduceat(v, heads), np.maximum.reduceat(v, heads)]

    def tail_boundary(self, into):
        runs = len(self.outs[0])
        rec = [1, runs, int(self.keys[0]) if runs else 0, int(self.keys[-1]) if runs else 0] + \
              [int(o[0]) if runs else 0 for o in self.outs] + [int(o[-1]) if runs else 0 for o in self.outs]
        for i, x in enumerate(rec):
            into[i] = x
        return into

    def tail_apply(self, drop_first, last_row=None):
        if last_row is not None:
            for o, x in zip(self.outs, last_row):
                o[-1] = x
        if drop_first:
            self.outs = [o[1:] for o in self.outs]

    def outputs(self, copy=True):
        return {f"out{i}": (o.copy() if copy else o) for i, o in enumerate(self.outs)}


TAIL_KEYS = np.sort(np.random.default_rng(7).integers(0, 40, 500)).astype(np.int64)
TAIL_VALS = np.random.default_rng(8).integers(-50, 50, 500).astype(np.int64)


def tail_worker(rank, world, port, q, cuts):
    from mplan2vdl_b200.dist import ShardedPlan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    for cut in cuts:
        bounds = [0] + list(cut) + [len(TAIL_KEYS)]
        plan = _FakeTailPlan(TAIL_KEYS[bounds[rank]:bounds[rank + 1]], TAIL_VALS[bounds[rank]:bounds[rank + 1]])
        sp = ShardedPlan.__new__(ShardedPlan)
        sp.ctx, sp.plan, sp.rank, sp.world, sp.group = _FakeCtx(rank, ("none", -1, 0)), plan, rank, world, None
        sp.peer_mode, sp._want_peer, sp.tail_ops, sp.tail_mode, sp._tb, sp._tail_slice = False, False, plan.tail_info(), True, None, None
        mine = sp.step()                                   # this rank's slice
        whole = sp.global_result()                         # every rank: the concatenation
        res.append(({k: v.tolist() for k, v in mine.items()}, {k: v.tolist() for k, v in whole.items()}))
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cuts", [(2, [(250,), (0,), (500,), (251,)]), (3, [(100, 300), (10, 11), (200, 200)])])
def test_gloo_sharded_tail_merges_the_groups_that_straddle_ranks(world, cuts):
    """The N>1 host path of emit plans (dist.ShardedPlan._step_tail): every rank folds its own rows, one all-gather of the
    boundary records, the straddling groups merged into the rank where they start; the concatenated slices are the fold of
    the whole vector.  Cuts include empty shards and cuts inside a run."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=tail_worker, args=(r, world, port, q, cuts)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole_plan = _FakeTailPlan(TAIL_KEYS, TAIL_VALS)
    whole_plan.execute()
    want = {k: v.tolist() for k, v in whole_plan.outputs().items()}
    for c in range(len(cuts)):
        for rank in range(world):
            assert got[rank][c][1] == want, (cuts[c], rank)
        glued = {k: sum((got[r][c][0][k] for r in range(world)), []) for k in want}
        assert glued == want
