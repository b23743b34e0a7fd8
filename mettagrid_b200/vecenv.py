"""PufferLib-style vectorised env over a BatchedSimulation.

Mirrors ``MettaGridPufferEnv.reset / step`` (python/src/mettagrid/envs/mettagrid_puffer_env.py:285-408)
for a batch: flat ``[N*A, ...]`` views of the same CUDA buffers, the combined-index action decoding of
``:313-394`` done on the device, and auto-reset of finished environments (``:299-302``: an env whose
agents are all terminal or all truncated is rebuilt before its next step) through ``mg_reset`` with a
device mask -- observations never leave HBM.

``desync_episodes`` restates ``EarlyResetHandler`` (python/src/mettagrid/envs/early_reset_handler.py:6-27): the FIRST
episode of every environment is truncated at a step drawn from ``numpy.random.default_rng(env seed)`` in
``[1, max_steps]`` so that the environments of a batch do not all end on the same tick.
"""

from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .sim import BatchedSimulation


def decode_actions(actions: torch.Tensor, num_primary: int, vibe_action_ids: torch.Tensor):
    """Reference decoding (mettagrid_puffer_env.py:313-394) on device.

    1-D ``[n]``: values in ``[0, P)`` are primary actions with no vibe change; values in
    ``[P, P + P*V)`` encode ``offset = v - P``, ``primary = offset // V``, ``vibe = offset % V``.
    2-D ``[n, 2]``: column 0 primary, column 1 an index into the vibe action list.
    Returns (core_actions, vibe_actions) as int32; vibe_actions holds env action ids (0 = none)."""
    P, V = int(num_primary), int(vibe_action_ids.numel())
    a = actions.to(torch.int64)
    if a.dim() >= 2 and a.shape[-1] in (1, 2) and actions.dim() == 2:
        core = a[:, 0]
        if a.shape[1] == 1:
            vibe = torch.zeros_like(core)
        else:
            if V <= 0:
                raise ValueError("Received 2D actions with vibe column, but environment has no configured vibe action space")
            raw = a[:, 1]
            if bool((raw < 0).any()) or bool((raw >= V).any()):
                raise ValueError(f"Vibe action indices out of range [0,{V}), min={int(raw.min())} max={int(raw.max())}")
            vibe = vibe_action_ids.to(a.device)[raw]
    elif a.dim() == 1:
        if bool((a < 0).any()):
            raise ValueError(f"Actions must be non-negative, got min={int(a.min())}")
        enc = a >= P
        core, vibe = a.clone(), torch.zeros_like(a)
        if bool(enc.any()):
            if V <= 0:
                raise ValueError("Received encoded vibe actions, but environment has no configured vibe action space")
            if bool((a >= P + P * V).any()):
                raise ValueError(f"Action indices out of range [0, {P + P * V}), min={int(a.min())} max={int(a.max())}")
            off = a - P
            core = torch.where(enc, off // V, a)
            vibe = torch.where(enc, vibe_action_ids.to(a.device)[(off % V).clamp(min=0)], vibe)
    else:
        raise ValueError(f"Expected step actions shape [num_agents] or [num_agents,2], got {tuple(actions.shape)}")
    if bool((core < 0).any()) or bool((core >= P).any()):
        raise ValueError(f"Core actions out of range [0,{P}), min={int(core.min())} max={int(core.max())}")
    return core.to(torch.int32), vibe.to(torch.int32)


class MettaGridVecEnv:
    """N envs x A agents presented as N*A agents, buffers on the GPU.

    ``step`` is one asynchronous C-ABI call (``mg_vecenv_step``, csrc/mg_vecenv.cu): finished environments are
    found and rebuilt, actions decoded, the tick run and the early-reset truncation applied on the device.
    With ``validate=True`` (default, the reference's behaviour) out-of-range actions raise ``ValueError`` at
    once, which costs one host synchronisation per step; ``validate=False`` defers the check to ``poll()``."""

    def __init__(self, cfg: Any, num_envs: int, seed: int = 0, desync_episodes: bool = False, validate: bool = True, **kw):
        self.sim = BatchedSimulation(cfg, num_envs, seeds=seed, **kw)
        s = self.sim
        game = getattr(cfg, "game", cfg)
        names = s.program.action_names
        self.action_names = [n for n in names if not n.startswith("change_vibe_")]
        self.vibe_action_names = [n for n in names if n.startswith("change_vibe_")]
        self.num_primary = len(self.action_names)
        vibe_ids = [i for i, n in enumerate(names) if n.startswith("change_vibe_")]
        self._vibe_ids = torch.tensor(vibe_ids, dtype=torch.int64, device=s.device)
        self.num_envs, self.agents_per_env = s.num_envs, s.num_agents
        self.num_agents = s.num_envs * s.num_agents
        self.validate = validate
        self._episodes = 0
        self._max_steps = int(game.max_steps)
        self._desync = bool(desync_episodes) and self._max_steps > 0
        self._vibe_ids_host = np.asarray(vibe_ids, dtype=np.int32)
        self._configure()

    def _configure(self):
        early = None
        if self._desync:  # early_reset_handler.py:15-20: one draw per env from a generator seeded with the env's seed
            early = np.asarray([int(np.random.default_rng(int(sd)).integers(1, self._max_steps + 1)) for sd in self.sim.seeds],
                               dtype=np.int64)  # fmt: skip
        v = self._vibe_ids_host
        self.sim._check(self.sim._L.mg_vecenv_configure(self.sim._h, self.num_primary, v.ctypes.data if v.size else None, int(v.size),
                                                        None if early is None else early.ctypes.data))  # fmt: skip

    # flat zero-copy views, the layout PufferLib hands to policies
    @property
    def observations(self) -> torch.Tensor:
        return self.sim.observations.view(self.num_agents, self.sim.num_tokens, 3)

    @property
    def rewards(self) -> torch.Tensor:
        return self.sim.rewards.view(-1)

    @property
    def terminals(self) -> torch.Tensor:
        return self.sim.terminals.view(-1)

    @property
    def truncations(self) -> torch.Tensor:
        return self.sim.truncations.view(-1)

    def poll(self) -> int:
        """Synchronise, raise for any invalid action seen since the last poll, return the error bits (0)."""
        import ctypes

        n, err = ctypes.c_int(0), ctypes.c_int(0)
        self.sim._check(self.sim._L.mg_vecenv_poll(self.sim._h, ctypes.byref(n), ctypes.byref(err)))
        self._episodes = int(n.value)
        return int(err.value)

    @property
    def episodes_finished(self) -> int:
        self.poll()
        return self._episodes

    def reset(self, seed: int | None = None):
        seeds = None if seed is None else [seed + e for e in range(self.num_envs)]
        self.sim.reset(seeds=seeds)
        self._configure()  # step counters and, with desync, a fresh first-episode cut
        return self.observations, {}

    def step(self, actions: torch.Tensor):
        s = self.sim
        a = torch.as_tensor(actions, device=s.device)
        if a.dtype not in (torch.int32, torch.int64):
            a = a.to(torch.int64)
        if a.dim() == 2 and a.shape[1] == 1:
            a = a[:, 0]
        if not ((a.dim() == 1 and a.shape[0] == self.num_agents) or (a.dim() == 2 and a.shape == (self.num_agents, 2))):
            raise ValueError(f"Expected step actions shape [num_agents] or [num_agents,2], got {tuple(a.shape)}")
        a = a.contiguous()
        # an env can only finish through max_steps (plain termination / truncation) or the early-reset cut
        s._check(s._L.mg_vecenv_step(s._h, a.data_ptr(), int(a.dtype == torch.int64), a.dim(), int(self._max_steps > 0), s._stream()))
        if self.validate and self.poll():
            decode_actions(a, self.num_primary, self._vibe_ids)  # raises the reference's ValueError for this tensor
            raise ValueError("invalid actions")
        return self.observations, self.rewards, self.terminals, self.truncations, {}

    def close(self):
        self.sim.close()
