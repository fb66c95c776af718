"""Row-range sharded execution: one process per GPU, torch.distributed for the plumbing.

The fact table is split by contiguous row range (tpch.shard_range), dimension tables are replicated, every rank
runs the same plan on its shard (vdl_plan_run_local), the per-rank partial aggregate tables -- [accumulators x key
domain] int64, a few bytes to a few KB -- are all-gathered (NCCL over NVLink on GPUs), and every rank merges them in
the finalize kernel (vdl_plan_finish): SUM/MIN/MAX accumulators combine by their op, FoldChoose takes the value
from the rank that holds the smallest first row, empty groups are dropped after the merge.  There is no other
data-path collective: the scan itself never leaves the GPU that owns the rows.
"""
from __future__ import annotations

import torch


class DeviceView:
    """Library-owned device memory exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, n_int64: int):
        self.__cuda_array_interface__ = {"shape": (n_int64,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def gather_partial_tables(local: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """All-gather one rank-local partial table (1-D int64 tensor, CPU/gloo or CUDA/nccl) into [world * n]."""
    import torch.distributed as dist
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local, group=group)
    else:                                   # gloo has no all_gather_into_tensor for every build: use the list form
        parts = list(out.view(world, -1).unbind(0))
        dist.all_gather(parts, local, group=group)
    return out


class ShardedPlan:
    """A plan executed over this rank's shard; step() returns the global result on every rank."""

    def __init__(self, ctx, plan, rank: int, world: int, row_base: int, group=None):
        self.ctx, self.plan, self.rank, self.world, self.group = ctx, plan, rank, world, group
        plan.set_row_base(row_base)
        self._stream = torch.cuda.ExternalStream(ctx.stream, device=ctx.device) if world > 1 else None
        self._gathered = []

    def step(self) -> dict:
        if self.world == 1:
            return self.plan.run()                      # one launch per fused scan: its last thread block finalizes
        self.plan.run_local()
        ptrs = []
        with torch.cuda.stream(self._stream):           # NCCL is ordered after the scan on the library's stream
            for i in range(self.plan.num_fused):
                ptr, n = self.plan.partials(i)
                local = torch.as_tensor(DeviceView(ptr, n), device=f"cuda:{self.ctx.device}")
                g = gather_partial_tables(local, self.world, self.group)
                if len(self._gathered) <= i:
                    self._gathered.append(None)
                self._gathered[i] = g                   # keep alive until finish() has consumed it
                ptrs.append(g.data_ptr())
        return self.plan.finish(ptrs, self.world)
