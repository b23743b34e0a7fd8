"""The PettingZoo-style adapter (mettagrid_b200/pettingzoo_env.py) against the CPU oracle: dict API, primary-action
indexing, agents dropping out at the end of the episode, reset with a new seed."""

import numpy as np
import pytest

from tests import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_parallel_env_matches_oracle():
    from mettagrid_b200.pettingzoo_env import MettaGridParallelEnv
    from oracle.oracle import OracleEnv

    cfg = cases.walled_config(3, max_steps=15)
    env = MettaGridParallelEnv(cfg, seed=4)
    sim = env._sim
    P = sim.program
    for episode, seed in enumerate((4, 9)):
        obs, infos = env.reset(seed=seed)
        orc = OracleEnv(P, sim._init_cells[0], seed, sim._init_gstats[0])
        assert sorted(obs) == [0, 1, 2] and all(np.array_equal(obs[a], orc.observations()[a]) for a in obs)
        rs = np.random.RandomState(episode)
        for t in range(15):
            acts = {a: int(rs.randint(0, env.num_actions)) for a in env.agents}
            full = np.zeros(3, np.int32)
            for a, i in acts.items():
                full[a] = env._action_indices[i]
            o, r, term, trunc, _ = env.step(acts)
            orc.step(full, np.zeros(3, np.int32))
            for a in o:
                assert np.array_equal(o[a], orc.observations()[a]), f"episode {episode} step {t} agent {a}"
                assert r[a] == float(orc.rewards()[a])
                assert term[a] == bool(orc.terminals()[a]) and trunc[a] == bool(orc.truncations()[a])
        assert env.agents == []  # every agent left at max_steps
    with pytest.raises(ValueError, match="out of range"):
        env.reset()
        env.step({0: env.num_actions})
    env.close()
