"""GPU parity for the static layer of k_step_fast (mettagrid_b200/csrc/mg_fast.cu, k_step_fast<G, true>): handler-free
games on maps with more objects than lanes -- walls kept as a bitmap and a cell list outside the object lanes.

Same bar as tests/test_gpu_fast_path.py: the fast kernel, the generic plain kernels and the CPU oracle step the same
seeded inputs and must agree bit for bit per step on observations, rewards, flags and action_success, and at the end on
stats (touched keys included -- `cell.visited` carries every wall's `visited` stamp), object state and episode rewards."""

import numpy as np
import pytest

from tests import cases
from tests.test_gpu_fast_path import _make, _sparse_config, _triple

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

EIGHT = ["north", "south", "west", "east", "northwest", "northeast", "southwest", "southeast"]


def _maps(cfg, n, seed0=100):
    from mettagrid_b200.mapgen import random_map

    return [random_map(cfg.game.map_builder, seed=seed0 + e) for e in range(n)]


def test_toy_preset_selects_the_static_fast_kernel():
    # the reference's perf_benchmark.py "toy" preset: 40 x 40, border + 4 % walls (220 objects), 20 agents, 8-way, 11 x 11
    from mettagrid_b200 import workloads as W

    cfg = W.toy_config(max_steps=70)
    _triple(cfg, num_envs=6, steps=90, expect_lanes=64 + 32, check_every=3, maps=_maps(cfg, 6))


@pytest.mark.parametrize(
    "agents,lanes,walls,width,height,obs_w,obs_h,tokens",
    [
        (6, 8, 60, 16, 12, 7, 5, 200),      # elliptical window, 8 lanes: 4 envs per warp with different wall counts
        (12, 16, 90, 20, 20, 11, 11, 200),  # 16 lanes
        (3, 8, 40, 9, 9, 15, 15, 200),      # the widest window: it always hangs over the map's edge
        (5, 8, 50, 70, 6, 3, 3, 200),       # a map row spans three bitmap words; the smallest window
        (32, 32, 100, 24, 24, 9, 13, 200),  # every lane an agent
        (2, 8, 330, 20, 20, 5, 5, 200),     # almost everything is wall: most moves are blocked
        (4, 8, 260, 20, 20, 15, 15, 200),   # ~100 tokens per row: past the staged prefix, work list overflows
        (5, 8, 60, 14, 14, 9, 9, 201),      # 3T not a multiple of 8: the byte-wise stream-out
        (3, 8, 45, 12, 10, 7, 7, 36),       # T below the staged prefix; rows are 108 bytes (not a multiple of 8)
        (20, 32, 300, 30, 30, 13, 13, 252),  # crowded and walled: long lists of dynamic objects per observer
        (1, 8, 45, 12, 40, 11, 11, 200),     # a single agent (no shuffle draws) on a tall map
    ],
)
def test_walled_maps_static_layer(agents, lanes, walls, width, height, obs_w, obs_h, tokens):
    cfg = _sparse_config(agents, walls=walls, width=width, height=height, obs_w=obs_w, obs_h=obs_h, num_tokens=tokens,
                         max_steps=50, directions=EIGHT, local_position=True, last_action_move=True)  # fmt: skip
    n = 9
    _triple(cfg, num_envs=n, steps=70, expect_lanes=64 + lanes, check_every=2, p_vibe=0.3, maps=_maps(cfg, n))


def test_walled_8way_fixture_config_selects_the_static_kernel():
    # tests/cases.py: the config behind the walled golden fixtures, on a map large enough to leave the object lanes
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    cfg = cases.walled_config(8, max_steps=40)
    maps = [random_map(RandomMapConfig(agents=8, width=22, height=17, seed=300 + e, border_width=1, objects={"wall": 30}))
            for e in range(5)]  # fmt: skip
    _triple(cfg, num_envs=5, steps=60, expect_lanes=64 + 8, maps=maps)


def test_token_budget_overflow_is_reported_with_walls():
    from mettagrid_b200.sim import MettaGridError

    cfg = _sparse_config(4, walls=60, width=10, height=10, obs_w=9, obs_h=9, num_tokens=14)
    sim = _make(cfg, 4, 3, True, _maps(cfg, 4))
    assert sim.step_kernel == 64 + 8
    A = sim.program.num_agents
    with pytest.raises(MettaGridError):
        for _ in range(3):
            sim.step(np.zeros((4, A), np.int32), np.zeros((4, A), np.int32))
            sim.check_errors()
    sim.close()


def test_introspection_masked_reset_and_set_buffers_on_a_static_handle():
    """Getters, set_buffers and resets move state between the packed block + static layer and the generic arrays
    (k_fast_pack_static / k_fast_unpack): the walls' `visited` stamps and the occupancy grid must survive the trip."""
    from oracle.oracle import OracleEnv

    cfg = _sparse_config(7, walls=70, width=18, height=14, obs_w=9, obs_h=9, num_tokens=200, directions=EIGHT)
    N = 6
    maps = _maps(cfg, N)
    sim = _make(cfg, N, 33, True, maps)
    assert sim.step_kernel == 64 + 8
    P = sim.program
    A = P.num_agents
    T = P.num_tokens

    def fresh(e):
        return OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e])

    orc = [fresh(e) for e in range(N)]
    nprim = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
    prim, vibe = cases.random_actions(np.random.RandomState(9), 90, (N, A), nprim, len(P.action_names), 0.3, 0.05)
    for t in range(90):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(orc):
            o.step(prim[t, e], vibe[t, e])
        if t % 7 == 3:
            for e, o in enumerate(orc):
                assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ at step {t} env {e}"
                assert np.array_equal(sim.dump_objects(e), o.dump_objects()), f"objects differ at step {t} env {e}"
        if t == 30:  # _init_buffers against new buffers: reads the occupancy grid k_fast_unpack rebuilt
            bufs = (torch.empty((N, A, T, 3), dtype=torch.uint8, device="cuda"), torch.zeros((N, A), dtype=torch.bool, device="cuda"),
                    torch.zeros((N, A), dtype=torch.bool, device="cuda"), torch.zeros((N, A), dtype=torch.float32, device="cuda"),
                    torch.zeros((N, A), dtype=torch.int32, device="cuda"), torch.zeros((N, A), dtype=torch.int32, device="cuda"))  # fmt: skip
            sim.set_buffers(*bufs)
            for o in orc:
                o.reinit_buffers()
        if t == 40:
            mask = torch.zeros(N, dtype=torch.bool, device="cuda")
            mask[1] = mask[4] = True
            sim.reset(mask)
            for e in (1, 4):
                orc[e] = fresh(e)
        if t == 60:
            sim.reset()
            orc = [fresh(e) for e in range(N)]
        torch.cuda.synchronize()
        obs = sim.observations.cpu().numpy()
        for e, o in enumerate(orc):
            assert np.array_equal(obs[e], o.observations()), f"obs differ at step {t} env {e}"
    sim.check_errors()
    for e, o in enumerate(orc):
        assert sim.get_episode_stats(e) == o.get_episode_stats()
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    sim.close()


def test_set_map_on_a_static_handle():
    from mettagrid_b200 import native
    from mettagrid_b200.mapgen import RandomMapConfig, random_map
    from oracle.oracle import OracleEnv

    cfg = _sparse_config(4, walls=50, width=16, height=12, obs_w=7, obs_h=7, num_tokens=200)
    sim = _make(cfg, 3, 50, True, _maps(cfg, 3))
    assert sim.step_kernel == 64 + 8
    P = sim.program
    A = P.num_agents
    rs = np.random.RandomState(8)
    acts = rs.randint(0, 5, size=(40, 3, A)).astype(np.int32)
    zeros = np.zeros((3, A), np.int32)
    for t in range(10):
        sim.step(acts[t], zeros)
    new_map = random_map(RandomMapConfig(agents=4, width=16, height=12, seed=777, border_width=0, objects={"wall": 35}))
    sim.set_map(1, new_map)
    mask = torch.zeros(3, dtype=torch.bool, device="cuda")
    mask[1] = True
    sim.reset(mask)
    cells, gs = P.encode_map(new_map, with_stats=True)
    o = OracleEnv(P, cells, int(sim.seeds[1]), gs)
    for t in range(10, 40):
        sim.step(acts[t], zeros)
        o.step(acts[t, 1], zeros[1])
    torch.cuda.synchronize()
    assert np.array_equal(sim.observations[1].cpu().numpy(), o.observations())
    assert sim.get_episode_stats(1) == o.get_episode_stats()
    assert np.array_equal(sim.dump_objects(1), o.dump_objects())
    sim.check_errors()
    # more walls than the handle's static layer was sized for
    big = random_map(RandomMapConfig(agents=4, width=16, height=12, seed=5, border_width=0, objects={"wall": 120}))
    cells, _ = P.encode_map(big, with_stats=True)
    cells = np.ascontiguousarray(cells, dtype=np.int16)
    assert sim._L.mg_set_map(sim._h, 0, cells.ctypes.data, None) == native.MG_E_INVALID
    sim.close()


def test_vecenv_autoreset_on_a_static_handle():
    """Auto-reset restores finished envs from the post-reset snapshot; the walls' stamps start over with them."""
    from mettagrid_b200.vecenv import MettaGridVecEnv
    from oracle.oracle import OracleEnv

    cfg = _sparse_config(4, walls=45, width=14, height=11, obs_w=7, obs_h=7, num_tokens=200, max_steps=9)
    N = 5
    env = MettaGridVecEnv(cfg, N, seed=3, maps=_maps(cfg, N))
    sim = env.sim
    assert sim.step_kernel == 64 + 8

    def fresh(e):
        return OracleEnv(sim.program, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e])

    orc = [fresh(e) for e in range(N)]
    rng = np.random.RandomState(12)
    P, V = env.num_primary, len(env.vibe_action_names)
    for t in range(32):
        a = rng.randint(0, P + P * V, size=N * 4)
        a[rng.rand(N * 4) < 0.7] %= P
        for e in range(N):
            if orc[e].terminals().all() or orc[e].truncations().all():
                orc[e] = fresh(e)
        core = np.where(a >= P, (a - P) // V, a).reshape(N, 4)
        vib = np.where(a >= P, P + (a - P) % V, 0).reshape(N, 4)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(a).cuda())
        for e, o in enumerate(orc):
            o.step(core[e], vib[e])
        torch.cuda.synchronize()
        ob = obs.cpu().numpy().reshape(N, 4, -1, 3)
        for e, o in enumerate(orc):
            assert np.array_equal(ob[e], o.observations()), f"step {t} env {e}"
    for e, o in enumerate(orc):
        assert sim.get_episode_stats(e) == o.get_episode_stats(), f"stats differ in env {e}"
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    env.close()


def test_long_token_lists_with_vibe_edits_on_a_static_handle():
    # every agent carries 9 resources: 12-13 cached tokens, more than the eight kept in the packed block -- the tail
    # lives in the agent's object record, which the static variant finds through the agent's slot, and vibe edits shift it
    cfg = _sparse_config(7, walls=60, width=15, height=13, obs_w=9, obs_h=9, num_tokens=200, directions=EIGHT)
    cfg.game.agent.inventory.initial = {"ore_red": 3, "ore_blue": 300, "ore_green": 1, "battery_red": 9, "battery_blue": 2,
                                        "heart": 70000 % 65536, "armor": 5, "laser": 1, "blueprint": 8}  # fmt: skip
    n = 8
    _triple(cfg, num_envs=n, steps=80, expect_lanes=64 + 8, p_vibe=0.6, check_every=2, maps=_maps(cfg, n))


def test_set_inventory_and_reset_on_a_static_handle():
    from oracle.oracle import OracleEnv

    cfg = _sparse_config(4, walls=55, width=14, height=12, obs_w=7, obs_h=7, num_tokens=200)
    N = 5
    sim = _make(cfg, N, 21, True, _maps(cfg, N))
    assert sim.step_kernel == 64 + 8
    P = sim.program
    A = P.num_agents
    rs = np.random.RandomState(5)
    acts = rs.randint(0, 5, size=(40, N, A)).astype(np.int32)
    zeros = np.zeros((N, A), np.int32)
    for t in range(10):
        sim.step(acts[t], zeros)
    sim.set_inventory(2, 1, {"heart": 3, "ore_red": 7})  # mid-episode: the agent's token cache is rebuilt from its record
    sim.reset()
    orc = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(N)]
    sim.set_inventory(2, 1, {"heart": 3, "ore_red": 7})
    orc[2].set_inventory(1, {"heart": 3, "ore_red": 7})
    for t in range(10, 40):
        sim.step(acts[t], zeros)
        for e, o in enumerate(orc):
            o.step(acts[t, e], zeros[e])
        if t == 25:
            sim.set_inventory(4, 3, {"laser": 2})
            orc[4].set_inventory(3, {"laser": 2})
    torch.cuda.synchronize()
    obs = sim.observations.cpu().numpy()
    for e, o in enumerate(orc):
        assert np.array_equal(obs[e], o.observations()), f"obs differ in env {e}"
        assert sim.get_episode_stats(e) == o.get_episode_stats()
        assert np.array_equal(sim.dump_objects(e), o.dump_objects())
    sim.check_errors()
    sim.close()
