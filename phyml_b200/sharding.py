"""Site sharding across ranks: every rank owns a contiguous block of site
patterns of every buffer; the only exchange is the sum of the per-rank partial lnL (and d lnL).

Two transports for that sum:
  * the engine's own NCCL communicator (plk_comm_init): the all-reduce is enqueued on the engine
    stream right behind the reduction kernel -- the product path on GPUs;
  * a torch.distributed process group (gloo on CPU, used by the world_size-2 tests of the host
    logic; also usable with nccl) -- `dist_reduce_fn`.
"""
from __future__ import annotations

from typing import Callable, Sequence

from .alignment import Patterns, shard_bounds  # noqa: F401  (re-exported)


def dist_reduce_fn(group=None) -> Callable[[Sequence[float]], Sequence[float]]:
    """reduce_fn for LkTree: element-wise SUM over the ranks of a torch.distributed group."""
    import torch
    import torch.distributed as dist

    def reduce(vals):
        backend = dist.get_backend(group)
        dev = "cuda" if backend == "nccl" else "cpu"
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.tolist()

    return reduce


def init_engine_comm(engine, rank: int, world: int, group=None, mode: str = "p2p") -> None:
    """Wire the engine's cross-GPU sum.  mode "p2p": mailboxes over NVLink peer memory written by the
    reduction kernels themselves (CUDA IPC handles travel over `group`); mode "nccl": an ncclAllReduce
    enqueued behind the reduction kernel (the 128-byte unique id travels over `group`)."""
    import torch.distributed as dist

    from .engine import Engine

    if mode == "p2p":
        mine = engine.comm_p2p_export(world)
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        engine.comm_p2p_init(rank, world, handles)
        dist.barrier(group=group)
        return
    uid = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, group=group)
    engine.comm_init(rank, world, uid[0])
