"""Data-parallel plumbing: one process per GPU, rays sharded across ranks, bucketed all-reduce of
the flat gradient buffer (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference has no distributed code (framework/pipelines.py:306-320 trains on one device); the
path shards naturally because rays are independent: ranks exchange nothing on the data
path.  The trainer normalises every loss by GLOBAL counts (global ray count, all-reduced masked-mean
denominators), so the per-rank gradients simply sum: one bucketed sum-all-reduce per step, started bucket
by bucket from events the backward pass records (trainer.Trainer._all_reduce_and_step).  The helpers here
are the environment / sharding plumbing and a stand-alone bucketed reducer (used by the gloo tests)."""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        import datetime
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        # a mismatched collective must fail in minutes, not hold the GPUs for the default 10
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("SNB_DIST_TIMEOUT_S", "180"))), **kw)
    return rank, local, world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [lo, hi) of n rays for `rank`: sizes differ by at most one, every ray exactly once."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bucket_ranges(table, n_params: int, n_buckets: int = 3) -> List[Tuple[int, int]]:
    """Split the flat gradient into contiguous buckets on tensor boundaries, roughly equal in size.
    The flat layout is [trunk ... | heads ...] (state_dict order); backward produces the heads first, so
    buckets are returned last-to-first: the order in which their gradients become final."""
    target = n_params / n_buckets
    cuts, acc = [0], 0
    for _, off, shape in table:
        n = 1
        for s in shape:
            n *= s
        acc += n
        if acc >= target * len(cuts) and len(cuts) < n_buckets:
            cuts.append(off + n)
    cuts.append(n_params)
    cuts = sorted(set(cuts))
    ranges = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
    return ranges[::-1]


class GradAllReducer:
    """Bucketed, asynchronous sum all-reduce of a flat fp32 gradient buffer."""

    def __init__(self, ranges: List[Tuple[int, int]]):
        self.ranges = ranges
        self.handles = []

    def launch(self, flat_grad: torch.Tensor, bucket: int | None = None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        idx = range(len(self.ranges)) if bucket is None else [bucket]
        for i in idx:
            lo, hi = self.ranges[i]
            self.handles.append(dist.all_reduce(flat_grad[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles = []


def all_reduce_scalar_sum(t: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
