"""GPU parity: the CUDA step, called through the C ABI, against the CPU oracle restatement on the same
seeded inputs -- bit-exact observations, rewards, flags, stats and object state, per step."""

import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _run_pair(cfg, num_envs, steps, seed0=42, p_vibe=0.1, p_invalid=0.0, check_every=1):
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    sim = BatchedSimulation(cfg, num_envs, seeds=seed0)
    P = sim.program
    oracles = []
    for e in range(num_envs):
        oracles.append(OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]))
    A = P.num_agents
    num_primary = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
    prim, vibe = cases.random_actions(np.random.RandomState(seed0), steps, (num_envs, A), num_primary, len(P.action_names),
                                      p_vibe, p_invalid)  # fmt: skip
    torch.cuda.synchronize()
    obs0 = sim.observations.cpu().numpy()
    for e, o in enumerate(oracles):
        assert np.array_equal(obs0[e], o.observations()), f"initial observation differs in env {e}"
    for t in range(steps):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t % check_every and t != steps - 1:
            continue
        torch.cuda.synchronize()
        obs = sim.observations.cpu().numpy()
        rew = sim.rewards.cpu().numpy()
        term = sim.terminals.cpu().numpy()
        trunc = sim.truncations.cpu().numpy()
        succ = sim.action_success()
        for e, o in enumerate(oracles):
            assert np.array_equal(obs[e], o.observations()), f"obs differ: step {t} env {e}"
            assert np.array_equal(rew[e].view(np.uint32), o.rewards().view(np.uint32)), f"rewards differ: step {t} env {e}"
            assert np.array_equal(term[e], o.terminals()) and np.array_equal(trunc[e], o.truncations())
            assert np.array_equal(succ[e], o.action_success()), f"action_success differs: step {t} env {e}"
    sim.check_errors()
    er = sim.episode_rewards()
    for e, o in enumerate(oracles):
        assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ in env {e}"
        assert np.array_equal(sim.dump_objects(e), o.dump_objects()), f"object state differs in env {e}"
        assert np.array_equal(er[e].view(np.uint32), o.episode_rewards().view(np.uint32))
    sim.close()


@pytest.mark.parametrize("num_agents", [1, 2, 4, 16])
def test_benchmark_config_matches_oracle(num_agents):
    _run_pair(cases.benchmark_config(num_agents), num_envs=8, steps=200)


def test_invalid_actions_and_vibes():
    _run_pair(cases.benchmark_config(5), num_envs=6, steps=150, p_vibe=0.5, p_invalid=0.1)


def test_walls_eight_way_truncation():
    cfg = cases.walled_config(6, directions=["north", "south", "west", "east", "northwest", "northeast", "southwest", "southeast"],
                              max_steps=40)  # fmt: skip
    _run_pair(cfg, num_envs=5, steps=60)


def test_local_position_and_last_action_move():
    import mettagrid_b200.config as C

    cfg = cases.walled_config(3, max_steps=0, global_obs=C.GlobalObsConfig(local_position=True, last_action_move=True),
                              width=9, height=7, walls=6, num_tokens=200)  # fmt: skip
    cfg.game.obs.width, cfg.game.obs.height = 7, 5
    _run_pair(cfg, num_envs=4, steps=120)


def test_many_envs_c2_shape():
    # the C2 shape at reduced env count; every env checked at a stride of steps
    _run_pair(cases.benchmark_config(16), num_envs=64, steps=64, check_every=8)


@pytest.mark.parametrize("base", [256, 10])
def test_combat_handlers_inventory_rewards(base):
    """C3-style game: handler chains, inventory limits with modifiers, on_use/on_tick, rewards (logf)."""
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    cfg = cases.combat_config(None, 4, token_value_base=base, max_steps=150)
    maps = [cases.combat_map(4, seed=s) for s in range(12)]
    sim = BatchedSimulation(cfg, 12, seeds=100, maps=maps)
    P = sim.program
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(12)]
    prim, vibe = cases.random_actions(np.random.RandomState(3), 200, (12, 8), 9, len(P.action_names), 0.3, 0.01)
    for t in range(200):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t % 5 == 0 or t == 199:
            torch.cuda.synchronize()
            obs, rew = sim.observations.cpu().numpy(), sim.rewards.cpu().numpy()
            for e, o in enumerate(oracles):
                assert np.array_equal(obs[e], o.observations()), f"obs differ: step {t} env {e}"
                assert np.array_equal(rew[e].view(np.uint32), o.rewards().view(np.uint32)), f"rewards differ: step {t} env {e}"
    sim.check_errors()
    for e, o in enumerate(oracles):
        assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ in env {e}"
        assert np.array_equal(sim.dump_objects(e), o.dump_objects()), f"object state differs in env {e}"
    sim.close()


@pytest.mark.parametrize("spawn", [True, False])
def test_world_aoe_territory_events_queries(spawn):
    """C4-style game: fixed + mobile AOE, territory handlers and aoe_mask tokens, events with max_targets
    (RNG shuffles), tag queries, tag add/remove with handlers, raycast spawn, remove-when-empty."""
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    cfg = cases.world_config(None, 3, spawn=spawn, max_steps=0)
    maps = [cases.world_map(3, seed=40 + s) for s in range(10)]
    sim = BatchedSimulation(cfg, 10, seeds=500, maps=maps)
    P = sim.program
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(10)]
    prim, vibe = cases.random_actions(np.random.RandomState(8), 300, (10, 6), 9, len(P.action_names), 0.2)
    torch.cuda.synchronize()
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(oracles):
        assert np.array_equal(obs[e], o.observations()), f"initial obs differ: env {e}"
    for t in range(300):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t % 3 == 0 or t == 299:
            torch.cuda.synchronize()
            obs, rew = sim.observations.cpu().numpy(), sim.rewards.cpu().numpy()
            for e, o in enumerate(oracles):
                assert np.array_equal(obs[e], o.observations()), f"obs differ: step {t} env {e}"
                assert np.array_equal(rew[e].view(np.uint32), o.rewards().view(np.uint32)), f"rewards differ: step {t} env {e}"
    sim.check_errors()
    for e, o in enumerate(oracles):
        assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ in env {e}"
        assert np.array_equal(sim.dump_objects(e), o.dump_objects()), f"object state differs in env {e}"
    sim.close()


def test_network_queries_push_clear():
    """Materialized / closure / raycast queries, query-inventory transfers with stats, push, clear-inventory,
    game-value filter, ratio / min / max rewards, game on_tick, on_after_use."""
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    cfg = cases.network_config(None, 4)
    maps = [cases.network_map(4, seed=70 + s) for s in range(8)]
    sim = BatchedSimulation(cfg, 8, seeds=900, maps=maps)
    P = sim.program
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(8)]
    prim, vibe = cases.random_actions(np.random.RandomState(9), 300, (8, 4), 5, len(P.action_names), 0.1)
    for t in range(300):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t % 4 == 0 or t == 299:
            torch.cuda.synchronize()
            obs, rew = sim.observations.cpu().numpy(), sim.rewards.cpu().numpy()
            for e, o in enumerate(oracles):
                assert np.array_equal(obs[e], o.observations()), f"obs differ: step {t} env {e}"
                assert np.array_equal(rew[e].view(np.uint32), o.rewards().view(np.uint32)), f"rewards differ: step {t} env {e}"
    sim.check_errors()
    for e, o in enumerate(oracles):
        assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ in env {e}"
        assert np.array_equal(sim.dump_objects(e), o.dump_objects()), f"object state differs in env {e}"
    sim.close()


@pytest.mark.parametrize("walls", [24, 26, 27, 40])
def test_object_count_around_sparse_threshold(walls):
    """<= 32 objects take the one-lane-per-object observation path, more take the per-cell path: both sides of
    the switch, 8-way moves so agents shuffle around the walls."""
    import mettagrid_b200.config as C
    from mettagrid_b200.mapgen import RandomMapConfig

    cfg = C.MettaGridConfig(
        game=C.GameConfig(
            num_agents=6,
            obs=C.ObsConfig(width=9, height=7, num_tokens=150, global_obs=C.GlobalObsConfig(local_position=True)),
            max_steps=0,
            actions=C.ActionsConfig(noop=C.NoopActionConfig(), move=C.MoveActionConfig(allowed_directions=list(cases.EIGHT_WAY))),
            objects={"wall": C.WallConfig()},
            map_builder=RandomMapConfig(agents=6, width=11, height=9, seed=5, objects={"wall": walls}),
        )
    )
    _run_pair(cfg, num_envs=6, steps=120, p_vibe=0.2)
