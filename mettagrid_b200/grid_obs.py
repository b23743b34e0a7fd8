"""Dense grid observations on the device.

``GridObsVecEnv`` stands where the reference's ``GridObsWrapper`` stands
(python/src/mettagrid/envs/grid_obs_wrapper.py:25-129): it wraps the vectorised env and hands out
``(num_agents, C, H, W)`` float32 observations instead of token rows -- here a CUDA tensor produced by the
``k_obs_to_grid`` kernel (csrc/mg_gridobs.cu) right after the step kernel, so observations never leave HBM.
"""

from __future__ import annotations

import torch

from .vecenv import MettaGridVecEnv


class GridObsVecEnv:
    def __init__(self, env: MettaGridVecEnv):
        self._env = env
        C, H, W = env.sim.grid_obs_shape()
        self.single_observation_shape = (C, H, W)  # Box(low=0, high=inf, shape=(C, H, W), float32) in the reference
        self._grid = torch.empty((env.num_agents, C, H, W), dtype=torch.float32, device=env.sim.device)

    def _convert(self) -> torch.Tensor:
        return self._env.sim.grid_observations(out=self._grid)

    def reset(self, seed: int | None = None):
        _, info = self._env.reset(seed=seed)
        return self._convert(), info

    def step(self, actions):
        _, rewards, terminals, truncations, info = self._env.step(actions)
        return self._convert(), rewards, terminals, truncations, info

    @property
    def num_agents(self) -> int:
        return self._env.num_agents

    @property
    def action_names(self):
        return self._env.action_names

    def close(self) -> None:
        self._env.close()
