"""Many CPU oracle instances stepped from a thread pool (the ctypes calls release the GIL) -- test helper."""

from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from oracle.oracle import OracleEnv


class OracleBatch:
    def __init__(self, sim, envs=None):
        P = sim.program
        self.envs = list(range(sim.num_envs)) if envs is None else list(envs)
        self.pool = ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4))
        self.orc = list(self.pool.map(lambda e: OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]), self.envs))

    def _chunks(self):
        n = len(self.orc)
        k = max(1, n // (4 * self.pool._max_workers))
        return [range(i, min(i + k, n)) for i in range(0, n, k)]

    def step(self, prim: np.ndarray, vibe: np.ndarray):
        """prim / vibe: [N, A] for the whole batch; each oracle takes its env's row."""

        def work(rng):
            for i in rng:
                self.orc[i].step(prim[self.envs[i]], vibe[self.envs[i]])

        list(self.pool.map(work, self._chunks()))

    def first_mismatch(self, obs: np.ndarray, rewards: np.ndarray | None = None, success: np.ndarray | None = None):
        """obs: [N, A, T, 3] of the whole batch (or [len(envs), ...] when `envs` was a subset and obs was gathered)."""
        sub = obs.shape[0] == len(self.envs) and len(self.envs) != 0

        def work(rng):
            for i in rng:
                j = i if sub else self.envs[i]
                o = self.orc[i]
                if not np.array_equal(obs[j], o.observations()):
                    return f"observations differ in env {self.envs[i]}"
                if rewards is not None and not np.array_equal(rewards[j].view(np.uint32), o.rewards().view(np.uint32)):
                    return f"rewards differ in env {self.envs[i]}"
                if success is not None and not np.array_equal(success[j], o.action_success()):
                    return f"action_success differs in env {self.envs[i]}"
            return None

        for r in self.pool.map(work, self._chunks()):
            if r:
                return r
        return None

    def close(self):
        self.pool.shutdown()
