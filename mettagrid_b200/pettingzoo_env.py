"""PettingZoo-style parallel env over the CUDA step (SURVEY 8b "Callers").

``MettaGridParallelEnv`` mirrors ``MettaGridPettingZooEnv`` (python/src/mettagrid/envs/pettingzoo_env.py:22-160): integer
agent ids, ``reset(seed) -> (observations, infos)``, ``step({agent: action}) -> (observations, rewards, terminations,
truncations, infos)`` as dicts keyed by agent id, actions indexing the PRIMARY (non-vibe) action names, agents leaving
``self.agents`` once terminal or truncated.  One environment, host-side numpy results through ``mg_step_host`` -- the
dict API is a host API; batched rollouts use ``vecenv.MettaGridVecEnv``.  pettingzoo / gymnasium are not imported:
the class follows the ParallelEnv protocol by duck typing.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from .sim import BatchedSimulation


class MettaGridParallelEnv:
    def __init__(self, cfg: Any, seed: int = 0, device: int | None = None, **kw):
        self._sim = BatchedSimulation(cfg, 1, seeds=[seed], device=device, **kw)
        P = self._sim.program
        self._seed = seed
        # PolicyEnvInterface.action_names: the primary actions (policy/policy_env_interface.py:110-119)
        self._action_indices = [i for i, n in enumerate(P.action_names) if not n.startswith("change_vibe_")]
        self.possible_agents: list[int] = list(range(P.num_agents))
        self.agents: list[int] = self.possible_agents.copy()
        A, T = P.num_agents, P.num_tokens
        self.observation_shape = (T, 3)  # Box(low=0, high=255, shape=(T, 3), uint8) in the reference
        self.num_actions = len(self._action_indices)  # Discrete(num_actions)
        self._actions = np.zeros((1, A), dtype=np.int32)
        self._vibe = np.zeros((1, A), dtype=np.int32)
        self._obs = np.zeros((1, A, T, 3), dtype=np.uint8)
        self._rew = np.zeros((1, A), dtype=np.float32)
        self._term = np.zeros((1, A), dtype=np.uint8)
        self._trunc = np.zeros((1, A), dtype=np.uint8)

    def reset(self, seed: int | None = None, options: dict | None = None):
        if seed is not None:
            self._seed = seed
        self._sim.reset(seeds=[self._seed])  # a new simulation with this seed (:89-96)
        self.agents = self.possible_agents.copy()
        obs = self._sim.observations.cpu().numpy()[0]
        return {a: obs[a] for a in self.agents}, {a: {} for a in self.agents}

    def step(self, actions: dict):
        for a in self.agents:
            if a in actions:
                idx = int(np.asarray(actions[a], dtype=np.int32).reshape(()).item())
                if idx < 0 or idx >= len(self._action_indices):
                    raise ValueError(f"Action index {idx} out of range for PettingZoo action space.")
                self._actions[0, a] = self._action_indices[idx]
        self._sim.step_host(self._actions, self._vibe, self._obs, self._rew, self._term, self._trunc)
        self._sim.check_errors()
        obs_d = {a: self._obs[0, a].copy() for a in self.agents}
        rew_d = {a: float(self._rew[0, a]) for a in self.agents}
        term_d = {a: bool(self._term[0, a]) for a in self.agents}
        trunc_d = {a: bool(self._trunc[0, a]) for a in self.agents}
        info_d = {a: {} for a in self.agents}
        self.agents = [a for a in self.agents if not (term_d[a] or trunc_d[a])]
        return obs_d, rew_d, term_d, trunc_d, info_d

    def observation_space(self, agent: int):
        return self.observation_shape

    def action_space(self, agent: int):
        return self.num_actions

    def close(self) -> None:
        self._sim.close()
