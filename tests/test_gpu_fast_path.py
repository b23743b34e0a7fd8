"""GPU parity for k_step_fast (mettagrid_b200/csrc/mg_fast.cu): handler-free programs in sparse environments.

Every case runs the same seeded inputs through the fast kernel, the generic plain kernel
(METTAGRID_B200_NO_FAST=1 at mg_create) and the CPU oracle, and requires bit-exact agreement per step on
observations, rewards, flags, action_success, and at the end on stats (including touched keys), object state
and episode rewards."""

import os

import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _sparse_config(num_agents, walls, width, height, obs_w, obs_h, num_tokens=120, max_steps=0, directions=None, **gobs):
    import mettagrid_b200.config as C
    from mettagrid_b200.mapgen import RandomMapConfig

    mv = C.MoveActionConfig(allowed_directions=list(directions)) if directions else C.MoveActionConfig()
    obs = C.ObsConfig(num_tokens=num_tokens, width=obs_w, height=obs_h)
    if gobs:
        obs.global_obs = C.GlobalObsConfig(**gobs)
    return C.MettaGridConfig(
        game=C.GameConfig(
            num_agents=num_agents,
            obs=obs,
            max_steps=max_steps,
            actions=C.ActionsConfig(noop=C.NoopActionConfig(), move=mv, change_vibe=C.ChangeVibeActionConfig()),
            objects={"wall": C.WallConfig()},
            map_builder=RandomMapConfig(agents=num_agents, width=width, height=height, seed=11, border_width=0,
                                        objects={"wall": walls}),  # fmt: skip
        )
    )


def _make(cfg, num_envs, seed0, fast, maps=None):
    from mettagrid_b200.sim import BatchedSimulation

    old = os.environ.pop("METTAGRID_B200_NO_FAST", None)
    try:
        if not fast:
            os.environ["METTAGRID_B200_NO_FAST"] = "1"
        return BatchedSimulation(cfg, num_envs, seeds=seed0, maps=maps)
    finally:
        os.environ.pop("METTAGRID_B200_NO_FAST", None)
        if old is not None:
            os.environ["METTAGRID_B200_NO_FAST"] = old


def _triple(cfg, num_envs, steps, expect_lanes, seed0=7, p_vibe=0.2, p_invalid=0.02, check_every=1, maps=None):
    from oracle.oracle import OracleEnv

    fast = _make(cfg, num_envs, seed0, True, maps)
    slow = _make(cfg, num_envs, seed0, False, maps)
    assert fast.step_kernel == expect_lanes, f"fast path not selected: kernel {fast.step_kernel}"
    assert slow.step_kernel == 1
    P = fast.program
    oracles = [OracleEnv(P, fast._init_cells[e], int(fast.seeds[e]), fast._init_gstats[e]) for e in range(num_envs)]
    A = P.num_agents
    num_primary = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
    prim, vibe = cases.random_actions(np.random.RandomState(seed0), steps, (num_envs, A), num_primary, len(P.action_names),
                                      p_vibe, p_invalid)  # fmt: skip
    for t in range(steps):
        fast.step(prim[t], vibe[t])
        slow.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t % check_every and t != steps - 1:
            continue
        torch.cuda.synchronize()
        for name, sim in (("fast", fast), ("plain", slow)):
            obs = sim.observations.cpu().numpy()
            rew = sim.rewards.cpu().numpy()
            term, trunc = sim.terminals.cpu().numpy(), sim.truncations.cpu().numpy()
            succ = sim.action_success()
            for e, o in enumerate(oracles):
                assert np.array_equal(obs[e], o.observations()), f"{name}: obs differ at step {t} env {e}"
                assert np.array_equal(rew[e].view(np.uint32), o.rewards().view(np.uint32)), f"{name}: rewards step {t} env {e}"
                assert np.array_equal(term[e], o.terminals()) and np.array_equal(trunc[e], o.truncations()), f"{name}: flags"
                assert np.array_equal(succ[e], o.action_success()), f"{name}: action_success step {t} env {e}"
    for name, sim in (("fast", fast), ("plain", slow)):
        sim.check_errors()
        er = sim.episode_rewards()
        steps_now = sim.current_steps
        for e, o in enumerate(oracles):
            assert sim.get_episode_stats(e) == o.get_episode_stats(), f"{name}: stats differ in env {e}"
            assert np.array_equal(sim.dump_objects(e), o.dump_objects()), f"{name}: object state differs in env {e}"
            assert np.array_equal(er[e].view(np.uint32), o.episode_rewards().view(np.uint32))
            assert int(steps_now[e]) == o.current_step
        sim.close()


def test_c2_shape_16_lanes():
    # the benchmark game: 16 agents, 13 x 13 window, 157 actions; 2 envs per warp, ragged last CTA
    cfg = cases.benchmark_config(16)
    _triple(cfg, num_envs=37, steps=96, expect_lanes=16, check_every=4, p_vibe=0.3)


@pytest.mark.parametrize("agents,lanes", [(1, 8), (3, 8), (8, 8), (9, 16), (17, 32), (32, 32)])
def test_agent_counts_and_group_sizes(agents, lanes):
    cfg = cases.benchmark_config(agents)
    _triple(cfg, num_envs=9, steps=60, expect_lanes=lanes, check_every=3)


def test_walls_eight_way_small_window_truncation():
    eight = ["north", "south", "west", "east", "northwest", "northeast", "southwest", "southeast"]
    cfg = _sparse_config(6, walls=20, width=12, height=9, obs_w=7, obs_h=5, max_steps=45, directions=eight,
                         local_position=True, last_action_move=True)  # fmt: skip
    _triple(cfg, num_envs=11, steps=70, expect_lanes=32)


def test_crowded_map_move_conflicts():
    # 14 agents on 5 x 4 cells: most moves collide, so the shuffled resolution order decides every tick
    cfg = _sparse_config(14, walls=2, width=5, height=4, obs_w=5, obs_h=5, num_tokens=200)
    _triple(cfg, num_envs=16, steps=120, expect_lanes=16, p_vibe=0.5)


def test_token_budget_overflow_is_reported():
    from mettagrid_b200.sim import MettaGridError

    cfg = _sparse_config(8, walls=0, width=4, height=4, obs_w=7, obs_h=7, num_tokens=12)
    sim = _make(cfg, 4, 3, True)
    assert sim.step_kernel == 8
    A = sim.program.num_agents
    with pytest.raises(MettaGridError):
        for _ in range(3):
            sim.step(np.zeros((4, A), np.int32), np.zeros((4, A), np.int32))
            sim.check_errors()
    sim.close()


def test_rng_stream_across_state_wraparound():
    # 624 / 8 draws per tick: the state index wraps every 78 ticks; 400 ticks cross it five times
    cfg = cases.benchmark_config(16)
    _triple(cfg, num_envs=3, steps=400, expect_lanes=16, check_every=50, p_invalid=0.0)


def test_reset_and_set_inventory_on_fast_handle():
    from oracle.oracle import OracleEnv

    cfg = cases.benchmark_config(4)
    sim = _make(cfg, 5, 21, True)
    assert sim.step_kernel == 8
    P = sim.program
    A = P.num_agents
    rs = np.random.RandomState(5)
    acts = rs.randint(0, 5, size=(40, 5, A)).astype(np.int32)
    zeros = np.zeros((5, A), np.int32)
    for t in range(10):
        sim.step(acts[t], zeros)
    sim.set_inventory(2, 1, {"heart": 3, "ore_red": 7})
    sim.reset()
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(5)]
    sim.set_inventory(2, 1, {"heart": 3, "ore_red": 7})
    oracles[2].set_inventory(1, {"heart": 3, "ore_red": 7})
    for t in range(10, 40):
        sim.step(acts[t], zeros)
        for e, o in enumerate(oracles):
            o.step(acts[t, e], zeros[e])
    torch.cuda.synchronize()
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(oracles):
        assert np.array_equal(obs[e], o.observations()), f"obs differ in env {e}"
        assert sim.get_episode_stats(e) == o.get_episode_stats()
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    sim.close()


def test_introspection_and_masked_reset_between_ticks():
    """Getters, set_buffers and masked resets move state between the packed block and the generic arrays;
    none of it may disturb the episode."""
    from oracle.oracle import OracleEnv

    cfg = cases.benchmark_config(16)
    N = 6
    sim = _make(cfg, N, 33, True)
    assert sim.step_kernel == 16
    P = sim.program
    A = P.num_agents
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(N)]
    prim, vibe = cases.random_actions(np.random.RandomState(9), 90, (N, A), 5, len(P.action_names), 0.3, 0.05)
    for t in range(90):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t % 7 == 3:  # read everything mid-episode
            for e, o in enumerate(oracles):
                assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ at step {t} env {e}"
                assert np.array_equal(sim.dump_objects(e), o.dump_objects())
            assert [int(x) for x in sim.current_steps] == [o.current_step for o in oracles]
        if t == 40:  # rebuild two environments, keep the rest running
            mask = torch.zeros(N, dtype=torch.bool, device="cuda")
            mask[1] = mask[4] = True
            sim.reset(mask)
            for e in (1, 4):
                oracles[e] = OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e])
        if t == 60:  # a full reset right after a masked one
            sim.reset()
            oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(N)]
        torch.cuda.synchronize()
        obs = sim.observations.cpu().numpy()
        for e, o in enumerate(oracles):
            assert np.array_equal(obs[e], o.observations()), f"obs differ at step {t} env {e}"
    for e, o in enumerate(oracles):
        assert sim.get_episode_stats(e) == o.get_episode_stats()
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    sim.close()


def test_long_token_lists_with_vibe_edits():
    # every agent carries 9 resources: 1 tag + (vibe) + 9 inventory + group + id = 12-13 cached tokens, more than the
    # eight kept in registers / the packed block, so the tail lives in the object record and vibe edits shift it
    cfg = cases.benchmark_config(7, num_tokens=200)
    cfg.game.agent.inventory.initial = {"ore_red": 3, "ore_blue": 300, "ore_green": 1, "battery_red": 9, "battery_blue": 2,
                                        "heart": 70000 % 65536, "armor": 5, "laser": 1, "blueprint": 8}  # fmt: skip
    _triple(cfg, num_envs=10, steps=80, expect_lanes=8, p_vibe=0.6, check_every=2)


def test_truncating_episodes_and_small_token_base():
    # max_steps with episode_truncates (truncations, not terminals) and base-10 inventory digits in the token builder
    cfg = cases.benchmark_config(5, num_tokens=220, max_steps=30)
    cfg.game.episode_truncates = True
    cfg.game.obs.token_value_base = 10
    cfg.game.agent.inventory.initial = {"ore_red": 7, "heart": 345, "laser": 12}
    _triple(cfg, num_envs=7, steps=45, expect_lanes=8, p_vibe=0.4)


def test_token_stats_beyond_the_exact_float_range():
    """tokens_written / tokens_free_space are float32 sums of per-agent integers.  Below 2^24 every add is exact and the
    kernel adds a tick's total at once; beyond it the reference's per-agent add order decides the rounding and the
    kernel replays it.  36 000 ticks of 16 agents push both stats past 2^24."""
    from oracle.oracle import OracleEnv

    cfg = cases.benchmark_config(16)
    sim = _make(cfg, 3, 77, True)
    assert sim.step_kernel == 16
    P = sim.program
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(3)]
    rs = np.random.RandomState(1)
    pool = rs.randint(0, 5, size=(64, 3, 16)).astype(np.int32)
    zeros = np.zeros((3, 16), np.int32)
    steps = 36000
    for t in range(steps):
        sim.step(pool[t & 63], zeros)
        for e, o in enumerate(oracles):
            o.step(pool[t & 63, e], zeros[e])
    torch.cuda.synchronize()
    for e, o in enumerate(oracles):
        got, want = sim.get_episode_stats(e), o.get_episode_stats()
        assert want["game"]["tokens_free_space"] > 2**24, "the run must leave the exact range to test the replay"
        assert got == want, f"stats differ in env {e}"
        assert np.array_equal(sim.observations[e].cpu().numpy(), o.observations())
    sim.close()


def test_set_buffers_mid_episode_on_a_fast_handle_sees_the_moved_agents():
    """mg_set_buffers re-runs _init_buffers (mettagrid_c.cpp:1165-1184): the observations it writes must show the agents
    where they are now -- the fast kernel does not maintain the occupancy grid, k_fast_unpack rebuilds it."""
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    sim = BatchedSimulation(cases.benchmark_config(8), 6, seeds=21)
    assert sim.step_kernel >= 8
    P = sim.program
    orc = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(6)]
    prim, vibe = cases.random_actions(np.random.RandomState(3), 60, (6, 8), 5, len(P.action_names), 0.2)
    for t in range(60):
        if t in (25, 26, 50):
            N, A, T = 6, 8, P.num_tokens
            bufs = (torch.empty((N, A, T, 3), dtype=torch.uint8, device="cuda"), torch.zeros((N, A), dtype=torch.bool, device="cuda"),
                    torch.zeros((N, A), dtype=torch.bool, device="cuda"), torch.zeros((N, A), dtype=torch.float32, device="cuda"),
                    torch.zeros((N, A), dtype=torch.int32, device="cuda"), torch.zeros((N, A), dtype=torch.int32, device="cuda"))  # fmt: skip
            sim.set_buffers(*bufs)
            torch.cuda.synchronize()
            obs = sim.observations.cpu().numpy()
            for e, o in enumerate(orc):
                o.reinit_buffers()
                assert np.array_equal(obs[e], o.observations()), f"set_buffers at tick {t}: env {e}"
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(orc):
            o.step(prim[t, e], vibe[t, e])
    torch.cuda.synchronize()
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(orc):
        assert np.array_equal(obs[e], o.observations())
        assert sim.get_episode_stats(e) == o.get_episode_stats()
    sim.close()


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("game", ["benchmark", "walled", "combat"])
def test_step_host_leaves_the_callers_buffers_bound(game, pinned, monkeypatch):
    """mg_step_host steps on the handle's own staging set and must not rebind the handle: device steps, host steps and
    resets interleave on one handle, and the caller's tensors keep receiving the device steps' results."""
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    if game == "benchmark":
        cfg, maps, nprim = cases.benchmark_config(4), None, 5
    elif game == "walled":
        cfg, maps, nprim = cases.walled_config(4, max_steps=0), None, 5
    else:
        cfg, maps, nprim = cases.combat_config(None, 2), [cases.combat_map(2, seed=s) for s in range(3)], 9
    sim = BatchedSimulation(cfg, 3, seeds=8, maps=maps)
    P = sim.program
    A, T = P.num_agents, P.num_tokens
    mk = lambda: [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(3)]  # noqa: E731
    orc = mk()
    prim, vibe = cases.random_actions(np.random.RandomState(6), 50, (3, A), nprim, len(P.action_names), 0.2)
    h_obs = np.zeros((3, A, T, 3), np.uint8)
    if pinned:  # opt-in: pinned rows are written by the step kernels directly (no staging copy): same bytes
        monkeypatch.setenv("METTAGRID_B200_DIRECT_OBS", "1")
        keep = torch.zeros((3, A, T, 3), dtype=torch.uint8).pin_memory()
        h_obs = keep.numpy()
    h_rew, h_term, h_trunc = np.zeros((3, A), np.float32), np.zeros((3, A), np.uint8), np.zeros((3, A), np.uint8)
    for t in range(50):
        if t == 30:
            sim.reset()
            orc = mk()
            torch.cuda.synchronize()
            assert np.array_equal(sim.observations.cpu().numpy(), np.stack([o.observations() for o in orc])), "reset obs"
        host = t % 3 == 1
        if host:
            sim.step_host(prim[t], vibe[t], h_obs, h_rew, h_term, h_trunc)
        else:
            sim.step(prim[t], vibe[t])
            torch.cuda.synchronize()
        for e, o in enumerate(orc):
            o.step(prim[t, e], vibe[t, e])
        got = h_obs if host else sim.observations.cpu().numpy()
        rew = h_rew if host else sim.rewards.cpu().numpy()
        for e, o in enumerate(orc):
            assert np.array_equal(got[e], o.observations()), f"tick {t} ({'host' if host else 'device'}): env {e}"
            assert np.array_equal(rew[e].view(np.uint32), o.rewards().view(np.uint32))
    for e, o in enumerate(orc):
        assert sim.get_episode_stats(e) == o.get_episode_stats()
    sim.close()
