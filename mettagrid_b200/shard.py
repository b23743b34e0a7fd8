"""Multi-GPU plumbing: one process per GPU, environments sharded by index, no per-step collective.

The step path partitions perfectly -- environments share no state and each owns its RNG
(cpp/bindings/mettagrid_c.cpp:52) -- so ranks never exchange data while stepping.  The one collective
is the optional reduction of episode statistics over ranks, the batched analogue of the per-episode
agent-stat averaging in python/src/mettagrid/envs/stats_tracker.py:40-45; it runs over NCCL (NVLink)
on GPUs and over gloo in the CPU tests.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) env range of `rank`; the first `total % world` ranks take one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total_envs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def env_seeds(base_seed: int, begin: int, end: int) -> np.ndarray:
    """Env e (global index) is seeded base_seed + e on whichever rank owns it."""
    return (np.arange(begin, end, dtype=np.int64) + base_seed).astype(np.uint32)


def reduce_agent_stats(values: torch.Tensor, counts: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over all agents of all ranks of a dense per-agent stat table.

    values: [num_local_agents, S] float32 (this rank's agents), counts: scalar tensor = num_local_agents.
    Sums are reduced in float64 so the result does not depend on the rank count."""
    total = values.to(torch.float64).sum(dim=0)
    n = counts.to(torch.float64).reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return (total / n.clamp(min=1)).to(torch.float32)
