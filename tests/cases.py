"""Shared test configurations, written with the host-side mirror classes so they run on the GPU box
(where the reference package does not exist).  C1/C2 = the reference benchmark config
(benchmarks/test_mettagrid_env_benchmark.py:21-29)."""

from __future__ import annotations

import numpy as np

from mettagrid_b200 import config as C
from mettagrid_b200.mapgen import RandomMapConfig


def benchmark_config(num_agents: int, map_seed: int = 42, width: int = 20, height: int = 20, num_tokens: int = 100,
                     max_steps: int = 0):  # fmt: skip
    return C.MettaGridConfig(
        game=C.GameConfig(
            num_agents=num_agents,
            obs=C.ObsConfig(num_tokens=num_tokens),
            max_steps=max_steps,
            actions=C.ActionsConfig(noop=C.NoopActionConfig(), move=C.MoveActionConfig()),
            map_builder=RandomMapConfig(agents=num_agents, width=width, height=height, seed=map_seed),
        )
    )


def walled_config(num_agents: int, map_seed: int = 3, width: int = 16, height: int = 12, walls: int = 20,
                  directions=None, max_steps: int = 50, **obs_kw):  # fmt: skip
    mv = C.MoveActionConfig(allowed_directions=list(directions)) if directions else C.MoveActionConfig()
    return C.MettaGridConfig(
        game=C.GameConfig(
            num_agents=num_agents,
            obs=C.ObsConfig(num_tokens=obs_kw.pop("num_tokens", 120), **obs_kw),
            max_steps=max_steps,
            actions=C.ActionsConfig(noop=C.NoopActionConfig(), move=mv),
            objects={"wall": C.WallConfig()},
            map_builder=RandomMapConfig(agents=num_agents, width=width, height=height, seed=map_seed, border_width=1,
                                        objects={"wall": walls}),  # fmt: skip
        )
    )


def random_actions(rng: np.random.RandomState, steps: int, shape, num_primary: int, num_actions: int, p_vibe: float = 0.1,
                   p_invalid: float = 0.0):  # fmt: skip
    """'Effective' action sampling: primary uniform over non-vibe actions, vibe stream with
    probability p_vibe; optionally a few out-of-range indices."""
    prim = rng.randint(0, num_primary, size=(steps,) + tuple(shape)).astype(np.int32)
    vibe = np.zeros_like(prim)
    if num_actions > num_primary:
        m = rng.rand(*prim.shape) < p_vibe
        vibe[m] = rng.randint(num_primary, num_actions, size=int(m.sum()))
    if p_invalid > 0:
        m = rng.rand(*prim.shape) < p_invalid
        prim[m] = rng.choice([-1, num_actions, 9999], size=int(m.sum()))
        m = rng.rand(*prim.shape) < p_invalid
        vibe[m] = rng.choice([-3, num_actions + 5], size=int(m.sum()))
    return prim, vibe
