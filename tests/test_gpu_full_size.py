"""C2 at its full size: 4096 envs x 16 agents on one GPU.

* every env against its own CPU oracle for the first 64 ticks (checked at a stride of ticks and at the end);
* a sampled subset of envs against the oracle over 1000 ticks while the whole batch runs;
* a size-independent property: an env's trajectory does not depend on the batch around it (the same envs
  stepped in a small handle give identical bytes), and two runs of the batch are identical."""

import hashlib

import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

N, A = 4096, 16


def _actions(steps, seed):
    rs = np.random.RandomState(seed)
    prim = rs.randint(0, 5, size=(steps, N, A)).astype(np.int32)
    vibe = np.zeros_like(prim)
    m = rs.rand(steps, N, A) < 0.1
    vibe[m] = rs.randint(5, 157, size=int(m.sum()))
    return prim, vibe


def test_all_envs_first_64_ticks_and_sampled_envs_1000_ticks():
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    sim = BatchedSimulation(cases.benchmark_config(A), N, seeds=42)
    assert sim.step_kernel == 16
    P = sim.program
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(N)]
    prim, vibe = _actions(64, 1)
    for t in range(64):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t in (0, 15, 31, 47, 63):
            torch.cuda.synchronize()
            obs = sim.observations.cpu().numpy()
            succ = sim.action_success()
            for e, o in enumerate(oracles):
                assert np.array_equal(obs[e], o.observations()), f"obs differ: tick {t} env {e}"
                assert np.array_equal(succ[e], o.action_success()), f"action_success differs: tick {t} env {e}"
    sample = list(range(0, N, 128))  # 32 envs keep their oracle for the long run
    for e in sample:
        assert sim.get_episode_stats(e) == oracles[e].get_episode_stats()
    keep = {e: oracles[e] for e in sample}
    del oracles
    rs = np.random.RandomState(2)
    for t in range(64, 1000):
        p = rs.randint(0, 5, size=(N, A)).astype(np.int32)
        v = np.where(rs.rand(N, A) < 0.1, rs.randint(5, 157, size=(N, A)), 0).astype(np.int32)
        sim.step(p, v)
        for e, o in keep.items():
            o.step(p[e], v[e])
        if t % 117 == 0 or t == 999:
            torch.cuda.synchronize()
            obs = sim.observations[sample].cpu().numpy()
            for k, e in enumerate(sample):
                assert np.array_equal(obs[k], keep[e].observations()), f"obs differ: tick {t} env {e}"
    sim.check_errors()
    for e, o in keep.items():
        assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ in env {e}"
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    sim.close()


def test_batch_invariance_and_determinism():
    from mettagrid_b200.sim import BatchedSimulation

    cfg = cases.benchmark_config(A)
    prim, vibe = _actions(120, 3)
    digests = []
    for _ in range(2):
        sim = BatchedSimulation(cfg, N, seeds=42)
        h = hashlib.sha256()
        for t in range(120):
            sim.step(prim[t], vibe[t])
            if t % 40 == 39:
                torch.cuda.synchronize()
                h.update(sim.observations.cpu().numpy().tobytes())
        big_obs = sim.observations[1000:1008].cpu().numpy()
        big_stats = [sim.get_episode_stats(e) for e in range(1000, 1008)]
        sim.close()
        digests.append(h.hexdigest())
    assert digests[0] == digests[1]
    # the same eight envs alone: same maps (map seed 42 + e) and env seeds
    small = BatchedSimulation(cfg, 8, seeds=[42 + e for e in range(1000, 1008)], map_seeds=[42 + e for e in range(1000, 1008)])
    for t in range(120):
        small.step(prim[t, 1000:1008], vibe[t, 1000:1008])
    torch.cuda.synchronize()
    assert np.array_equal(small.observations.cpu().numpy(), big_obs)
    assert [small.get_episode_stats(e) for e in range(8)] == big_stats
    small.close()


# ---------------------------------------------------------------------------------------------------------------
# C3 / C4 / toy at the shapes bench.py times (same configs, same recorded MapGen pools, same action sampling):
# every env against its own oracle for the first ticks, 32 sampled envs over a long run.
# ---------------------------------------------------------------------------------------------------------------
def _full_shape(wl, num_envs, all_ticks, check_at, long_ticks, long_check):
    from mettagrid_b200 import workloads as W
    from mettagrid_b200.sim import BatchedSimulation
    from tests.oracle_batch import OracleBatch

    agents = W.WORKLOADS[wl][2]
    cfg = W.make_cfg(agents, wl)
    if wl in ("c3", "c4"):
        sim = BatchedSimulation(cfg, num_envs, seeds=42, maps=[W.make_map(cfg, agents, wl, e) for e in range(num_envs)])
    else:
        sim = BatchedSimulation(cfg, num_envs, seeds=42)
    P = sim.program
    na = len(P.action_names)
    nprim = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
    orc = OracleBatch(sim)
    prim, vibe = W.gen_actions(na, nprim, all_ticks, num_envs, agents, 11)
    for t in range(all_ticks):
        sim.step(prim[t], vibe[t])
        orc.step(prim[t], vibe[t])
        if t in check_at:
            torch.cuda.synchronize()
            bad = orc.first_mismatch(sim.observations.cpu().numpy(), sim.rewards.cpu().numpy(), sim.action_success())
            assert bad is None, f"{wl} tick {t}: {bad}"
    sim.check_errors()
    sample = list(range(0, num_envs, num_envs // 32))
    for e in sample[:8]:
        assert sim.get_episode_stats(e) == orc.orc[e].get_episode_stats(), f"{wl}: stats differ in env {e}"
    keep = OracleBatch.__new__(OracleBatch)
    keep.envs, keep.pool, keep.orc = sample, orc.pool, [orc.orc[e] for e in sample]
    orc.orc = None
    rs = np.random.RandomState(5)
    for t in range(all_ticks, long_ticks):
        p = rs.randint(0, nprim, size=(num_envs, agents)).astype(np.int32)
        v = np.zeros_like(p)
        if na > nprim:
            v = np.where(rs.rand(num_envs, agents) < 0.1, rs.randint(nprim, na, size=(num_envs, agents)), 0).astype(np.int32)
        sim.step(p, v)
        keep.step(p, v)
        if t % long_check == 0 or t == long_ticks - 1:
            torch.cuda.synchronize()
            bad = keep.first_mismatch(sim.observations[sample].cpu().numpy(), sim.rewards[sample].cpu().numpy())
            assert bad is None, f"{wl} tick {t}: {bad}"
    sim.check_errors()
    for i, e in enumerate(sample):
        assert sim.get_episode_stats(e) == keep.orc[i].get_episode_stats(), f"{wl}: stats differ in env {e}"
        assert np.array_equal(sim.dump_objects(e), keep.orc[i].dump_objects()), f"{wl}: objects differ in env {e}"
    keep.close()
    sim.close()


def test_c3_combat_at_bench_shape():
    """BASELINE config 3: 16 384 envs x 24 agents, 37 x 37 arena maps, 500 tokens."""
    _full_shape("c3", 16384, 16, (0, 7, 15), 500, 97)


def test_c4_world_at_bench_shape():
    """BASELINE config 4: 8192 envs x 24 agents, 66 x 66 MapGen maps, 200 tokens, AOE + territory + events."""
    _full_shape("c4", 8192, 12, (0, 5, 11), 400, 83)


def test_toy_preset_at_bench_shape():
    """The reference's perf_benchmark.py toy preset: 4096 envs x 20 agents, 40 x 40 walled maps, 8-way moves."""
    _full_shape("toy", 4096, 32, (0, 15, 31), 600, 101)


def test_toy_preset_many_waves():
    """12 288 envs of the toy preset: past 8192 warps mg_fast_layout switches the static variant to four-warp CTAs
    (k_step_fast<32, static, 4>), the instantiation the large-batch numbers in DESIGN.md come from."""
    _full_shape("toy", 12288, 8, (0, 7), 120, 37)
