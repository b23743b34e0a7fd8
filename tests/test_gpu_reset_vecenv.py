"""On-device reset (mg_reset), set_inventory and the Puffer-style vec-env (8f-2), checked
against fresh oracle instances."""

import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _oracles(sim, seeds=None):
    from oracle.oracle import OracleEnv

    P = sim.program
    seeds = sim.seeds if seeds is None else seeds
    return [OracleEnv(P, sim._init_cells[e], int(seeds[e]), sim._init_gstats[e]) for e in range(sim.num_envs)]


def test_reset_all_and_masked():
    from mettagrid_b200.sim import BatchedSimulation

    cfg = cases.combat_config(None, 3)
    maps = [cases.combat_map(3, seed=s) for s in range(6)]
    sim = BatchedSimulation(cfg, 6, seeds=30, maps=maps)
    P = sim.program
    prim, vibe = cases.random_actions(np.random.RandomState(1), 120, (6, 6), 9, len(P.action_names), 0.3)
    orc = _oracles(sim)
    for t in range(40):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(orc):
            o.step(prim[t, e], vibe[t, e])
    # masked reset: envs 1 and 4 start over, the others continue
    mask = torch.tensor([0, 1, 0, 0, 1, 0], dtype=torch.bool, device="cuda")
    sim.reset(env_mask=mask)
    fresh = _oracles(sim)
    orc[1], orc[4] = fresh[1], fresh[4]
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(orc):
        assert np.array_equal(obs[e], o.observations()), f"after masked reset: env {e}"
    for t in range(40, 80):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(orc):
            o.step(prim[t, e], vibe[t, e])
    torch.cuda.synchronize()
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(orc):
        assert np.array_equal(obs[e], o.observations()), f"env {e}"
        assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats env {e}"
    # full reset with new seeds
    new_seeds = [77 + e for e in range(6)]
    sim.reset(seeds=new_seeds)
    orc = _oracles(sim, new_seeds)
    for t in range(80, 120):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(orc):
            o.step(prim[t, e], vibe[t, e])
    torch.cuda.synchronize()
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(orc):
        assert np.array_equal(obs[e], o.observations())
        assert sim.get_episode_stats(e) == o.get_episode_stats()
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    sim.check_errors()
    sim.close()


def test_set_inventory_mid_episode():
    from mettagrid_b200.sim import BatchedSimulation

    cfg = cases.combat_config(None, 3)
    maps = [cases.combat_map(3, seed=s) for s in (3, 4)]
    sim = BatchedSimulation(cfg, 2, seeds=5, maps=maps)
    orc = _oracles(sim)
    prim, vibe = cases.random_actions(np.random.RandomState(2), 90, (2, 6), 9, len(sim.program.action_names), 0.3)
    for t in range(90):
        if t in (10, 40, 41):
            inv = {"loot": 3, "hp": 7, "energy": 5, "pack": 2} if t != 41 else {"weapon": 9}
            torch.cuda.synchronize()
            sim.set_inventory(1, t % 6, inv)
            orc[1].set_inventory(t % 6, inv)
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(orc):
            o.step(prim[t, e], vibe[t, e])
    torch.cuda.synchronize()
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(orc):
        assert np.array_equal(obs[e], o.observations())
        assert sim.get_episode_stats(e) == o.get_episode_stats()
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    sim.close()


def test_vecenv_autoreset_and_flat_views():
    from mettagrid_b200.vecenv import MettaGridVecEnv

    cfg = cases.walled_config(3, max_steps=12)
    env = MettaGridVecEnv(cfg, 5, seed=9)
    orc = _oracles(env.sim)
    assert env.observations.shape == (15, env.sim.num_tokens, 3) and env.observations.data_ptr() == env.sim.observations.data_ptr()
    rng = np.random.RandomState(4)
    P = env.num_primary
    V = len(env.vibe_action_names)
    for t in range(30):
        a = rng.randint(0, P + P * V, size=15)
        a[rng.rand(15) < 0.7] %= P  # mostly plain moves
        # the reference env rebuilds a finished sim before stepping it (mettagrid_puffer_env.py:299-302)
        for e in range(5):
            if orc[e].terminals().all() or orc[e].truncations().all():
                orc[e] = _oracles(env.sim)[e]
        core = np.where(a >= P, (a - P) // V, a).reshape(5, 3)
        vib = np.where(a >= P, P + (a - P) % V, 0).reshape(5, 3)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(a).cuda())
        for e, o in enumerate(orc):
            o.step(core[e], vib[e])
        torch.cuda.synchronize()
        ob = obs.cpu().numpy().reshape(5, 3, -1, 3)
        for e, o in enumerate(orc):
            assert np.array_equal(ob[e], o.observations()), f"step {t} env {e}"
            assert np.array_equal(term.cpu().numpy().reshape(5, 3)[e], o.terminals())
    assert env.episodes_finished == 10  # 5 envs x 2 completed 12-step episodes within 30 steps
    env.close()


def test_vecenv_desync_episodes_like_early_reset_handler():
    """EarlyResetHandler (envs/early_reset_handler.py): the first episode of env e ends, truncated, at the step
    numpy.random.default_rng(seed_e).integers(1, max_steps + 1) draws; later episodes run to max_steps."""
    from mettagrid_b200.vecenv import MettaGridVecEnv

    max_steps = 20
    env = MettaGridVecEnv(cases.benchmark_config(4, max_steps=max_steps), 6, seed=30, desync_episodes=True)
    early = [int(np.random.default_rng(30 + e).integers(1, max_steps + 1)) for e in range(6)]
    assert len(set(early)) > 1
    zeros = torch.zeros(24, dtype=torch.int64, device="cuda")
    ends = [[] for _ in range(6)]
    for t in range(1, 61):
        _, _, term, trunc, _ = env.step(zeros)
        done = (term.view(6, 4).all(dim=1) | trunc.view(6, 4).all(dim=1)).cpu().numpy()
        for e in range(6):
            if done[e]:
                ends[e].append(t)
    for e in range(6):
        want = [early[e]]
        while want[-1] + max_steps <= 60:
            want.append(want[-1] + max_steps)
        assert ends[e] == want, f"env {e}: episodes ended at {ends[e]}, expected {want}"
    env.close()


def test_set_map_gives_an_env_a_new_map_at_its_next_reset():
    """Host-side map pool: mg_set_map replaces one env's stored initial state."""
    from mettagrid_b200.mapgen import RandomMapConfig, random_map
    from mettagrid_b200.sim import BatchedSimulation, MettaGridError
    from oracle.oracle import OracleEnv

    for cfg, mk in (
        (cases.benchmark_config(4), lambda s: random_map(RandomMapConfig(agents=4, width=20, height=20, seed=s))),  # fast path
        (cases.walled_config(4, max_steps=0), lambda s: random_map(RandomMapConfig(agents=4, width=16, height=12, seed=s,
                                                                    border_width=1, objects={"wall": 20}))),  # fmt: skip
    ):
        sim = BatchedSimulation(cfg, 4, seeds=50)
        P = sim.program
        A = P.num_agents
        rs = np.random.RandomState(8)
        acts = rs.randint(0, 5, size=(40, 4, A)).astype(np.int32)
        zeros = np.zeros((4, A), np.int32)
        for t in range(10):
            sim.step(acts[t], zeros)
        new_map = mk(777)
        sim.set_map(2, new_map)
        mask = torch.zeros(4, dtype=torch.bool, device="cuda")
        mask[2] = True
        sim.reset(mask)
        cells, gs = P.encode_map(new_map, with_stats=True)
        fresh = OracleEnv(P, cells, int(sim.seeds[2]), gs)
        torch.cuda.synchronize()
        assert np.array_equal(sim.observations[2].cpu().numpy(), fresh.observations())
        for t in range(10, 40):
            sim.step(acts[t], zeros)
            fresh.step(acts[t, 2], zeros[2])
        torch.cuda.synchronize()
        assert np.array_equal(sim.observations[2].cpu().numpy(), fresh.observations())
        assert sim.get_episode_stats(2) == fresh.get_episode_stats()
        assert np.array_equal(sim.dump_objects(2), fresh.dump_objects())
        sim.check_errors()
        sim.close()
    # a map with more objects than a fast handle was sized for is refused by the C ABI
    from mettagrid_b200 import native

    sim = BatchedSimulation(cases.benchmark_config(4), 2, seeds=1)
    assert sim.step_kernel == 8
    cells = sim._init_cells[0].copy()
    free = np.argwhere(cells < 0)[:9]
    cells[free[:, 0], free[:, 1]] = cells.max()  # 13 objects for 8 lanes
    assert sim._L.mg_set_map(sim._h, 0, cells.ctypes.data, None) == native.MG_E_INVALID
    assert b"more objects" in sim._L.mg_last_error(sim._h)
    sim.close()


def test_vecenv_action_validation_on_device():
    from mettagrid_b200.vecenv import MettaGridVecEnv

    env = MettaGridVecEnv(cases.benchmark_config(2), 3, seed=2)
    P, V = env.num_primary, len(env.vibe_action_names)
    ok = torch.zeros(6, dtype=torch.int64, device="cuda")
    env.step(ok)
    bad = ok.clone()
    bad[4] = P + P * V  # one past the combined range
    with pytest.raises(ValueError, match="out of range"):
        env.step(bad)
    neg = ok.clone()
    neg[1] = -1
    with pytest.raises(ValueError, match="non-negative"):
        env.step(neg)
    with pytest.raises(ValueError, match="shape"):
        env.step(torch.zeros((6, 3), dtype=torch.int64, device="cuda"))
    pair = torch.zeros((6, 2), dtype=torch.int32, device="cuda")
    pair[:, 1] = V  # vibe column out of range
    with pytest.raises(ValueError, match="Vibe action indices"):
        env.step(pair)
    env.close()
    lazy = MettaGridVecEnv(cases.benchmark_config(2), 3, seed=2, validate=False)
    lazy.step(bad)  # no synchronisation, no exception: the offending agents do nothing
    assert lazy.poll() & 2
    assert lazy.poll() == 0  # reported once
    lazy.close()


@pytest.mark.parametrize("persistent", [False, True])
def test_vecenv_fast_path_autoreset_matches_oracle(persistent):
    """Auto-reset on a fast handle restores finished envs from the post-reset snapshot of the packed block; the
    trajectories must be those of freshly constructed environments, also after the snapshot was invalidated.
    `persistent`: the caller reuses one action tensor, so the steady-state step runs as one captured CUDA graph
    (mg_vecenv_step); a fresh tensor per step makes the handle rebuild the graph and then fall back to plain launches."""
    from mettagrid_b200.vecenv import MettaGridVecEnv

    cfg = cases.benchmark_config(4, max_steps=9)
    env = MettaGridVecEnv(cfg, 7, seed=3)
    assert env.sim.step_kernel == 8
    orc = _oracles(env.sim)
    rng = np.random.RandomState(12)
    P, V = env.num_primary, len(env.vibe_action_names)
    for t in range(40):
        a = rng.randint(0, P + P * V, size=28)
        a[rng.rand(28) < 0.7] %= P
        for e in range(7):
            if orc[e].terminals().all() or orc[e].truncations().all():
                orc[e] = _oracles(env.sim)[e]
        core = np.where(a >= P, (a - P) // V, a).reshape(7, 4)
        vib = np.where(a >= P, P + (a - P) % V, 0).reshape(7, 4)
        if t == 20:  # an inventory edit drops the snapshot: the generic reset path takes over
            env.sim.set_inventory(5, 2, {"heart": 4})
            orc[5].set_inventory(2, {"heart": 4})
        if persistent:
            if t == 0:
                held = torch.zeros(28, dtype=torch.int64, device="cuda")
            held.copy_(torch.from_numpy(a))
            obs, rew, term, trunc, _ = env.step(held)
        else:
            obs, rew, term, trunc, _ = env.step(torch.from_numpy(a).cuda())
        for e, o in enumerate(orc):
            o.step(core[e], vib[e])
        torch.cuda.synchronize()
        ob = obs.cpu().numpy().reshape(7, 4, -1, 3)
        for e, o in enumerate(orc):
            assert np.array_equal(ob[e], o.observations()), f"step {t} env {e}"
            assert np.array_equal(term.cpu().numpy().reshape(7, 4)[e], o.terminals())
    for e, o in enumerate(orc):
        assert env.sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ in env {e}"
        assert np.array_equal(env.sim.dump_objects(e), o.dump_objects())
    assert env.episodes_finished == 7 * 4
    env.close()


def test_vecenv_fast_path_stats_read_mid_episode_then_autoreset():
    """Stats read mid-episode (unpack) must not leak touched keys into the episode that follows a snapshot restore."""
    from mettagrid_b200.vecenv import MettaGridVecEnv

    cfg = cases.benchmark_config(4, max_steps=9)
    env = MettaGridVecEnv(cfg, 3, seed=13)
    assert env.sim.step_kernel == 8
    orc = _oracles(env.sim)
    P = env.num_primary
    rng = np.random.RandomState(2)
    for t in range(14):
        a = rng.randint(0, P, size=12)
        if t < 8:
            a[:] = 1  # everyone pushes north: some moves fail, 'action.failed' gets touched in episode 1
        else:
            a[:] = 0  # episode 2 only noops: 'action.failed' must not exist
        for e in range(3):
            if orc[e].terminals().all() or orc[e].truncations().all():
                orc[e] = _oracles(env.sim)[e]
        env.step(torch.from_numpy(a).cuda())
        for e, o in enumerate(orc):
            o.step(a.reshape(3, 4)[e], np.zeros(4, np.int32))
        if t == 6:
            for e, o in enumerate(orc):
                assert env.sim.get_episode_stats(e) == o.get_episode_stats()
    for e, o in enumerate(orc):
        assert env.sim.get_episode_stats(e) == o.get_episode_stats(), f"env {e}"
    env.close()


def test_masked_reset_with_new_seeds_then_autoreset_uses_the_new_seeds():
    from mettagrid_b200.vecenv import MettaGridVecEnv

    cfg = cases.benchmark_config(4, max_steps=6)
    env = MettaGridVecEnv(cfg, 4, seed=40)
    sim = env.sim
    new_seeds = [900 + e for e in range(4)]
    mask = torch.tensor([1, 0, 0, 0], dtype=torch.bool, device="cuda")
    sim.reset(env_mask=mask, seeds=new_seeds)  # before any step: every later reset of env e uses new_seeds[e]
    orc = _oracles(sim, [900, 40 + 1, 40 + 2, 40 + 3])
    rng = np.random.RandomState(1)
    P = env.num_primary
    for t in range(20):
        a = rng.randint(0, P, size=16)
        for e in range(4):
            if orc[e].terminals().all() or orc[e].truncations().all():
                orc[e] = _oracles(sim, new_seeds)[e]
        obs, *_ = env.step(torch.from_numpy(a).cuda())
        for e, o in enumerate(orc):
            o.step(a.reshape(4, 4)[e], np.zeros(4, np.int32))
        torch.cuda.synchronize()
        ob = obs.cpu().numpy().reshape(4, 4, -1, 3)
        for e, o in enumerate(orc):
            assert np.array_equal(ob[e], o.observations()), f"step {t} env {e}"
    env.close()


@pytest.mark.parametrize("game", ["benchmark", "combat"])
def test_step_info_keys_match_the_reference_payload(game):
    """mettagrid_puffer_env.py:230-282: game / attributes / agent keys, present only once the stat exists."""
    from mettagrid_b200.vecenv import MettaGridVecEnv

    keys = ["game/tokens_written", "env_game/objects.agent.red", "game/never.written", "attributes/steps", "attributes/seed",
            "attributes/map_w", "attributes/max_steps", "agent/action.move.success", "agent/action.failed", "agent/reward_step",
            "agent/reward_episode", "agent/cell.visited", "agent/not.a.stat"]  # fmt: skip
    if game == "combat":
        cfg, kw, nprim = cases.combat_config(None, 2, max_steps=40), {"maps": [cases.combat_map(2, seed=s) for s in range(3)]}, 9
        keys += ["game/attack.hits", "agent/hp.amount", "agent/loot.gained", "agent/attack.landed"]
    else:
        cfg, kw, nprim = cases.benchmark_config(4, max_steps=40), {}, 5
    env = MettaGridVecEnv(cfg, 3, seed=5, step_info_keys=keys, **kw)
    orc = _oracles(env.sim)
    P = env.sim.program
    A = P.num_agents
    rng = np.random.RandomState(3)
    with pytest.raises(ValueError, match="Unsupported step_info_keys"):
        MettaGridVecEnv(cfg, 3, seed=5, step_info_keys=["bogus/key"], **kw)
    for t in range(30):
        a = rng.randint(0, nprim, size=3 * A)
        _, _, _, _, info = env.step(torch.from_numpy(a).cuda())
        assert info["game"].is_cuda and info["agent"].shape == (3, A, len(info["agent_keys"]))
        for e, o in enumerate(orc):
            o.step(a.reshape(3, A)[e], np.zeros(A, np.int32))
        if t % 7 == 0 or t == 29:
            for e, o in enumerate(orc):
                st = o.get_episode_stats()
                want = {}
                for k in keys:
                    raw = k[len("env_"):] if k.startswith("env_") else k
                    if raw.startswith("game/") and raw[5:] in st["game"]:
                        want[raw] = st["game"][raw[5:]]
                want.update({"attributes/steps": float(t + 1), "attributes/seed": float(env.sim.seeds[e]), "attributes/map_w": float(P.width),
                             "attributes/max_steps": 40.0})  # fmt: skip
                per_agent = {}
                for ag in range(A):
                    row = {"reward_step": float(o.rewards()[ag]), "reward_episode": float(o.episode_rewards()[ag])}
                    for k in keys:
                        if k.startswith("agent/") and k[6:] in st["agent"][ag]:
                            row[k[6:]] = st["agent"][ag][k[6:]]
                    per_agent[ag] = row
                want["_per_agent_infos"] = per_agent
                assert env.step_info_payload(e) == want, f"tick {t} env {e}"
    env.close()
