"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes check the env sharding and the
episode-stat reduction (the only collective on the path)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mettagrid_b200.shard import env_seeds, reduce_agent_stats, shard_range


def test_shard_ranges_partition():
    for total in (1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert env_seeds(42, 3, 6).tolist() == [45, 46, 47]
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total_envs, agents, stats, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(total_envs, rank, world)
    # every env's agents carry a stat row that depends only on the GLOBAL env index
    rows = []
    for env in range(b, e):
        g = np.random.RandomState(1000 + env)
        rows.append(g.randint(0, 50, size=(agents, stats)).astype(np.float32))
    local = torch.from_numpy(np.concatenate(rows)) if rows else torch.zeros((0, stats))
    mean = reduce_agent_stats(local, torch.tensor(local.shape[0]))
    out_q.put((rank, b, e, mean.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_envs", [5, 16])
def test_stat_reduction_two_ranks_matches_single_process(total_envs):
    world, agents, stats = 2, 3, 6
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_envs, agents, stats, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == 0 and results[0][2] == results[1][1] and results[1][2] == total_envs
    expect = np.concatenate(
        [np.random.RandomState(1000 + env).randint(0, 50, size=(agents, stats)).astype(np.float32) for env in range(total_envs)]
    ).astype(np.float64).mean(0).astype(np.float32)
    for _, _, _, mean in results:
        assert np.array_equal(mean, expect)
