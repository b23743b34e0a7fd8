"""Live pinning (build container only): the oracle restatement AND our compiler against the real
reference -- its own Python lowering and its C++ step -- on fresh seeds, bit-exact per step."""

import numpy as np
import pytest

from mettagrid_b200.compiler import compile_config
from oracle.oracle import OracleEnv
from tests import cases, golden_cases as gc

pytestmark = pytest.mark.reference


def _compare(ref_cfg, mirror_cfg, grid, seed, steps, p_vibe=0.3, p_invalid=0.02):
    from mettagrid.config.mettagrid_c_config import convert_to_cpp_game_config, rename_map_agents
    from mettagrid.mettagrid_c import MettaGrid

    c_cfg, renames = convert_to_cpp_game_config(ref_cfg.game)
    env = MettaGrid(c_cfg, rename_map_agents(grid.tolist(), renames), seed)
    prog = compile_config(ref_cfg, *grid.shape)
    if mirror_cfg is not None:  # our mirror classes lower to the identical program
        assert np.array_equal(prog.blob, compile_config(mirror_cfg, *grid.shape).blob)
    cells, gs = prog.encode_map(grid, with_stats=True)
    orc = OracleEnv(prog, cells, seed, gs)
    assert np.array_equal(env.observations(), orc.observations())
    nprim = sum(1 for n in prog.action_names if not n.startswith("change_vibe_"))
    prim, vibe = cases.random_actions(np.random.RandomState(seed + 100), steps, (prog.num_agents,), nprim,
                                      len(prog.action_names), p_vibe, p_invalid)  # fmt: skip
    for t in range(steps):
        env.actions()[:] = prim[t]
        env.vibe_actions()[:] = vibe[t]
        env.step()
        orc.step(prim[t], vibe[t])
        assert np.array_equal(env.observations(), orc.observations()), f"obs differ at step {t}"
        assert np.array_equal(env.rewards().view(np.uint32), orc.rewards().view(np.uint32)), f"rewards differ at step {t}"
        assert np.array_equal(env.action_success(), orc.action_success())
    assert gc.clean_stats(env.get_episode_stats()) == gc.clean_stats(orc.get_episode_stats())
    assert np.array_equal(env.get_episode_rewards().view(np.uint32), orc.episode_rewards().view(np.uint32))


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_combat_fresh_seeds(reference_pkg, seed):
    from tests import refns

    ns = refns.reference_namespace()
    grid = cases.combat_map(3, seed=seed)
    _compare(cases.combat_config(ns, 3), cases.combat_config(None, 3), grid, seed, 400)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_world_fresh_seeds(reference_pkg, seed):
    """AOE, territory + aoe_mask, events with max_targets, queries, tags, spawn, remove-when-empty."""
    from tests import refns

    ns = refns.reference_namespace()
    grid = cases.world_map(3, seed=seed)
    _compare(cases.world_config(ns, 3), cases.world_config(None, 3), grid, seed, 350, p_vibe=0.2, p_invalid=0.0)


@pytest.mark.parametrize("seed", [31, 32])
def test_network_fresh_seeds(reference_pkg, seed):
    """Materialized / closure / raycast queries, query-inventory transfers, push, clear-inventory, value filters."""
    from tests import refns

    ns = refns.reference_namespace()
    grid = cases.network_map(4, seed=seed)
    _compare(cases.network_config(ns, 4), cases.network_config(None, 4), grid, seed, 350, p_vibe=0.1, p_invalid=0.0)


@pytest.mark.parametrize("agents", [1, 2, 8, 16])
def test_benchmark_fresh_seeds(reference_pkg, agents):
    from mettagrid_b200 import config as C
    from tests import refns

    ns = refns.reference_namespace()
    grid = gc._bench_map(agents, 100 + agents)
    _compare(gc._bench(ns, agents), gc._bench(C, agents), grid, 77, 300, p_vibe=0.1)


def test_ref_driver_agrees_with_reference_lowering(reference_pkg):
    """oracle/ref_driver.py (used on the GPU box, where the reference's Python cannot exist) builds
    the same game as the reference's own convert_to_cpp_game_config."""
    from mettagrid.config.mettagrid_c_config import convert_to_cpp_game_config, rename_map_agents
    from mettagrid.mettagrid_c import MettaGrid

    from oracle.ref_driver import RefEnv
    from tests import refns

    ns = refns.reference_namespace()
    grid = cases.combat_map(3, seed=4)
    c_cfg, renames = convert_to_cpp_game_config(cases.combat_config(ns, 3).game)
    a = MettaGrid(c_cfg, rename_map_agents(grid.tolist(), renames), 4)
    b = RefEnv(cases.combat_config(None, 3), grid, 4)
    prim, vibe = cases.random_actions(np.random.RandomState(4), 300, (6,), 9, 12, 0.3)
    assert np.array_equal(a.observations(), b.observations())
    for t in range(300):
        a.actions()[:] = prim[t]
        a.vibe_actions()[:] = vibe[t]
        a.step()
        b.step(prim[t], vibe[t])
        assert np.array_equal(a.observations(), b.observations()) and np.array_equal(a.rewards(), b.rewards())
    assert a.get_episode_stats() == b.get_episode_stats()


@pytest.mark.parametrize("game", ["world", "network", "beam_chain", "event_targets", "many_tagged"])
def test_ref_driver_agrees_with_reference_lowering_on_world_systems(reference_pkg, game):
    """The driver's pybind configs for AOE / territory / events / queries / tag handlers / spawn (what bench.py's C4
    cpu_baseline runs on the GPU box) give the same trajectories as the reference's own Python lowering."""
    from mettagrid.config.mettagrid_c_config import convert_to_cpp_game_config, rename_map_agents
    from mettagrid.mettagrid_c import MettaGrid

    from mettagrid_b200 import config as C
    from oracle.ref_driver import RefEnv
    from tests import edge_cases as ec, refns

    ns = refns.reference_namespace()
    mk, grid, nprim = {
        "world": (lambda n: cases.world_config(n, 3), cases.world_map(3, seed=8), 9),
        "network": (lambda n: cases.network_config(n, 4), cases.network_map(4, seed=8), 5),
        "beam_chain": (lambda n: ec.beam_chain_config(n if n is not None else C), ec.beam_chain_map(8), 5),
        "event_targets": (lambda n: ec.event_targets_config(n if n is not None else C), ec.event_targets_map(8), 5),
        "many_tagged": (lambda n: ec.many_tagged_config(n if n is not None else C), ec.many_tagged_map(8), 5),
    }[game]
    c_cfg, renames = convert_to_cpp_game_config(mk(ns).game)
    a = MettaGrid(c_cfg, rename_map_agents(grid.tolist(), renames), 8)
    b = RefEnv(mk(None), grid, 8)
    A = a.observations().shape[0]
    na = len(compile_config(mk(None), *grid.shape).action_names)
    prim, vibe = cases.random_actions(np.random.RandomState(8), 250, (A,), nprim, na, 0.2)
    assert np.array_equal(a.observations(), b.observations())
    for t in range(250):
        a.actions()[:] = prim[t]
        a.vibe_actions()[:] = vibe[t]
        a.step()
        b.step(prim[t], vibe[t])
        assert np.array_equal(a.observations(), b.observations()) and np.array_equal(a.rewards(), b.rewards()), f"{game}: step {t}"
    assert a.get_episode_stats() == b.get_episode_stats()


@pytest.mark.parametrize("seed", [41, 42])
def test_toy_preset_fresh_seeds(reference_pkg, seed):
    """The reference's perf_benchmark.py "toy" preset (benchmarks/perf/perf_benchmark.py:35-77): 40 x 40 map with a wall
    border and 4 % walls, 20 agents, noop + 8-way moves, 11 x 11 window, 200 tokens -- the workload the static layer of
    the fast path is measured on."""
    from mettagrid_b200 import workloads as W
    from mettagrid_b200.mapgen import random_map
    from tests import refns

    ns = refns.reference_namespace()
    mirror = W.toy_config(max_steps=150)
    ref = ns.MettaGridConfig(
        game=ns.GameConfig(
            num_agents=20,
            max_steps=150,
            obs=ns.ObsConfig(width=11, height=11, num_tokens=200),
            actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(allowed_directions=list(W.EIGHT_WAY))),
            objects={"wall": ns.WallConfig()},
        )
    )
    grid = random_map(mirror.game.map_builder, seed=seed)
    _compare(ref, mirror, grid, seed, 220, p_vibe=0.0, p_invalid=0.03)
