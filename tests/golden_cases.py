"""Registry of golden cases: name -> (config builder taking a class namespace, map builder, seed, steps,
action sampling).  The fixtures under tests/golden/*.npz are produced from these by
tests/golden/make_golden.py running the REAL reference (its own Python lowering + C++ step)."""

from __future__ import annotations

import hashlib
import json

import numpy as np

from tests import cases, edge_cases as ec


def _bench(ns, agents, **kw):
    return ns.MettaGridConfig(
        game=ns.GameConfig(
            num_agents=agents,
            obs=ns.ObsConfig(num_tokens=kw.get("num_tokens", 100)),
            max_steps=kw.get("max_steps", 0),
            actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig()),
        )
    )


def _bench_trunc(ns, agents):
    """Benchmark game with a step limit that truncates, base-10 inventory digits and agents that start with inventory."""
    cfg = _bench(ns, agents, num_tokens=220, max_steps=60)
    cfg.game.episode_truncates = True
    cfg.game.obs.token_value_base = 10
    cfg.game.agent.inventory.initial = {"ore_red": 7, "heart": 345, "laser": 12}
    return cfg


def _bench_rich(ns, agents):
    """Benchmark game whose agents start with nine resources: 12-13 tokens per agent, initial inventory order."""
    cfg = _bench(ns, agents, num_tokens=200)
    cfg.game.agent.inventory.initial = {"ore_red": 3, "ore_blue": 300, "ore_green": 1, "battery_red": 9, "battery_blue": 2,
                                        "heart": 70000 % 65536, "armor": 5, "laser": 1, "blueprint": 8}  # fmt: skip
    return cfg


def _bench_map(agents, seed):
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    return random_map(RandomMapConfig(agents=agents, width=20, height=20, seed=seed))


def _walled(ns, agents, **kw):
    return ns.MettaGridConfig(
        game=ns.GameConfig(
            num_agents=agents,
            obs=ns.ObsConfig(width=7, height=9, num_tokens=150, global_obs=ns.GlobalObsConfig(local_position=True, last_action_move=True)),
            max_steps=kw.get("max_steps", 60),
            episode_truncates=kw.get("truncates", True),
            actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(allowed_directions=list(cases.EIGHT_WAY))),
            objects={"wall": ns.WallConfig()},
        )
    )


def _toy(ns, agents=20, max_steps=250):
    """The reference's perf_benchmark.py "toy" preset (benchmarks/perf/perf_benchmark.py:35-77), as mettagrid_b200.workloads.toy_config."""
    eight = ["north", "south", "west", "east", "northwest", "northeast", "southwest", "southeast"]
    return ns.MettaGridConfig(
        game=ns.GameConfig(
            num_agents=agents,
            max_steps=max_steps,
            obs=ns.ObsConfig(width=11, height=11, num_tokens=200),
            actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(allowed_directions=eight)),
            objects={"wall": ns.WallConfig()},
        )
    )


def _toy_map(agents, seed):
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    return random_map(RandomMapConfig(width=40, height=40, agents=agents, border_width=1, seed=seed, objects={"wall": 64}))


def _walled_big_map(agents, seed):
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    return random_map(RandomMapConfig(agents=agents, width=22, height=17, seed=seed, border_width=1, objects={"wall": 30}))


def _walled_map(agents, seed):
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    return random_map(RandomMapConfig(agents=agents, width=14, height=10, seed=seed, border_width=1, objects={"wall": 15}))


CASES = {
    # name: (cfg builder(ns), map builder(), env seed, steps, p_vibe, p_invalid)
    "c1_a1": (lambda ns: _bench(ns, 1), lambda: _bench_map(1, 42), 42, 300, 0.1, 0.0),
    "c1_a4": (lambda ns: _bench(ns, 4), lambda: _bench_map(4, 42), 42, 1000, 0.0, 0.0),
    "c1_a16": (lambda ns: _bench(ns, 16), lambda: _bench_map(16, 42), 42, 400, 0.1, 0.0),
    "c1_a5_invalid": (lambda ns: _bench(ns, 5), lambda: _bench_map(5, 7), 9, 300, 0.4, 0.1),
    "c1_a5_trunc_base10": (lambda ns: _bench_trunc(ns, 5), lambda: _bench_map(5, 21), 21, 90, 0.4, 0.02),
    "c1_a7_nine_resources": (lambda ns: _bench_rich(ns, 7), lambda: _bench_map(7, 31), 31, 120, 0.6, 0.0),
    "walled_8way": (lambda ns: _walled(ns, 6), lambda: _walled_map(6, 3), 5, 90, 0.1, 0.02),
    "toy_a20": (lambda ns: _toy(ns), lambda: _toy_map(20, 42), 42, 300, 0.0, 0.02),  # 240 objects: the fast path's static layer
    "walled_8way_big": (lambda ns: _walled(ns, 6, max_steps=100), lambda: _walled_big_map(6, 4), 6, 140, 0.1, 0.02),  # 7 x 9 elliptical window, static layer
    "combat_3v3": (lambda ns: cases.combat_config(ns, 3), lambda: cases.combat_map(3, seed=1), 1, 500, 0.3, 0.0),
    "combat_4v4_base16": (lambda ns: cases.combat_config(ns, 4, token_value_base=16, max_steps=250),
                          lambda: cases.combat_map(4, seed=5), 5, 300, 0.3, 0.01),  # fmt: skip
    "world_3v3": (lambda ns: cases.world_config(ns, 3), lambda: cases.world_map(3, seed=2), 2, 400, 0.2, 0.0),
    "network_4": (lambda ns: cases.network_config(ns, 4), lambda: cases.network_map(4, seed=6), 6, 400, 0.1, 0.0),
    # corner cases the reference's own tests encode (tests/edge_cases.py)
    "c1_a8_wrong_stream": (lambda ns: _bench(ns, 8), lambda: _bench_map(8, 11), 11, 400, 0.0, 0.0),
    "beam_chain": (lambda ns: ec.beam_chain_config(ns), lambda: ec.beam_chain_map(3), 3, 500, 0.35, 0.0),
    "aoe_round": (lambda ns: ec.aoe_round_config(ns), lambda: ec.aoe_round_map(4), 4, 400, 0.0, 0.0),
    "dyn_limits": (lambda ns: ec.dyn_limits_config(ns), lambda: ec.dyn_limits_map(5), 5, 500, 0.3, 0.0),
    "event_targets": (lambda ns: ec.event_targets_config(ns), lambda: ec.event_targets_map(6), 6, 300, 0.0, 0.0),
    "many_tagged": (lambda ns: ec.many_tagged_config(ns), lambda: ec.many_tagged_map(7), 7, 200, 0.0, 0.0),
    "attack_mutation": (lambda ns: ec.attack_mutation_config(ns), lambda: ec.attack_mutation_map(8), 8, 500, 0.3, 0.0),
    "world_2v2_nospawn_trunc": (lambda ns: cases.world_config(ns, 2, spawn=False, max_steps=120, num_tokens=160),
                                lambda: cases.world_map(2, width=15, height=12, seed=9), 9, 150, 0.25, 0.01),  # fmt: skip
}


# cases no Python config of the reference can express: recorded through oracle/ref_driver.py (pybind configs built
# directly, checked against the reference's own lowering in tests/test_oracle_vs_reference.py) from our mirror classes
VIA_DRIVER = {"attack_mutation"}


def case_actions(name, prog):
    _, _, seed, steps, p_vibe, p_inv = CASES[name]
    A = prog.num_agents
    nprim = sum(1 for n in prog.action_names if not n.startswith("change_vibe_"))
    if name == "c1_a8_wrong_stream":
        return ec.wrong_stream_actions(np.random.RandomState(seed), steps, A, nprim, len(prog.action_names))
    if name == "c1_a4":  # the SURVEY 8(c) known-answer run
        return np.random.RandomState(42).randint(0, 5, size=(1000, 4)).astype(np.int32), np.zeros((1000, 4), np.int32)
    return cases.random_actions(np.random.RandomState(seed), steps, (A,), nprim, len(prog.action_names), p_vibe, p_inv)


def step_digest(obs, rewards, success) -> bytes:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(obs).tobytes())
    h.update(np.ascontiguousarray(rewards, dtype=np.float32).tobytes())
    h.update(np.ascontiguousarray(success).astype(np.uint8).tobytes())
    return h.digest()


def clean_stats(stats: dict) -> dict:
    """Drop the per-value 'action.invalid_index.<N>' keys (not tracked by the engines; the total is)."""
    out = {"game": dict(stats["game"]), "agent": []}
    for a in stats["agent"]:
        out["agent"].append({k: v for k, v in a.items() if not k.startswith("action.invalid_index.")})
    return out


def stats_json(stats: dict) -> str:
    return json.dumps(clean_stats(stats), sort_keys=True)
