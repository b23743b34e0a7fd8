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

from typing import Any, Sequence

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

    def __init__(self, cfg: Any, num_envs: int, seed: int = 0, desync_episodes: bool = False, validate: bool = True,
                 step_info_keys: Sequence[str] | None = None, **kw):  # fmt: skip
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
        self._configure_step_info_keys(step_info_keys)

    def _configure(self):
        early = None
        if self._desync:  # early_reset_handler.py:15-20: one draw per env from a generator seeded with the env's seed
            early = np.asarray([int(np.random.default_rng(int(sd)).integers(1, self._max_steps + 1)) for sd in self.sim.seeds],
                               dtype=np.int64)  # fmt: skip
        v = self._vibe_ids_host
        self.sim._check(self.sim._L.mg_vecenv_configure(self.sim._h, self.num_primary, v.ctypes.data if v.size else None, int(v.size),
                                                        None if early is None else early.ctypes.data))  # fmt: skip

    # ---- step_info_keys (mettagrid_puffer_env.py:132-184): same key grammar, gathered on the device -----------------
    def _configure_step_info_keys(self, keys: Sequence[str] | None) -> None:
        """'game/<stat>', 'team/<team>/<stat>', 'attributes/{seed,map_w,map_h,steps,max_steps}' (an 'env_' prefix is
        accepted and dropped like the reference does) and 'agent/<stat>' incl. reward_step / reward_episode."""
        self._info_game: list[tuple[str, int]] = []  # (payload key, game stat id | -1 steps)
        self._info_attr: list[tuple[str, str]] = []
        self._info_agent: list[tuple[str, int]] = []
        P = self.sim.program
        for key in dict.fromkeys(str(k) for k in (keys or [])):
            if key.startswith("agent/"):
                name = key[len("agent/"):]
                if not name:
                    raise ValueError("step_info_keys contains invalid entry 'agent/' (missing key suffix)")
                sid = {"reward_step": -1, "reward_episode": -2}.get(name)
                if sid is None:
                    sid = P.agent_stat_names.index(name) if name in P.agent_stat_names else None
                if sid is not None:  # a stat the program never writes is never present: the reference omits it too
                    self._info_agent.append((name, sid))
                continue
            raw = key[len("env_"):] if key.startswith("env_") else key
            if raw.startswith("game/") or raw.startswith("team/"):
                if raw.startswith("game/"):
                    stat = raw[len("game/"):]
                    if not stat:
                        raise ValueError("step_info_keys contains invalid entry 'game/' (missing key suffix)")
                else:
                    rest = raw[len("team/"):]
                    cut = rest.find("/")
                    if cut <= 0:
                        raise ValueError(f"step_info_keys entry {key!r}: expected 'team/{{team}}/{{stat}}'")
                    if not rest[cut + 1:]:
                        raise ValueError(f"step_info_keys entry {key!r}: missing stat key after team name")
                    stat = rest
                if stat in P.game_stat_names:
                    self._info_game.append((raw, P.game_stat_names.index(stat)))
                continue
            if raw.startswith("attributes/"):
                attr = raw[len("attributes/"):]
                if not attr:
                    raise ValueError("step_info_keys contains invalid entry 'attributes/' (missing key suffix)")
                if attr not in ("seed", "map_w", "map_h", "steps", "max_steps"):
                    raise ValueError(f"Unsupported step_info_keys attribute {raw!r}. Supported: seed, map_w, map_h, steps, max_steps.")
                if attr == "steps":
                    self._info_game.append((raw, -1))
                else:
                    self._info_attr.append((raw, attr))
                continue
            raise ValueError(f"Unsupported step_info_keys entry {key!r}; expected 'game/...', 'attributes/...', 'team/...', or 'agent/...'.")
        s = self.sim
        g = np.asarray([i for _, i in self._info_game], dtype=np.int32)
        a = np.asarray([i for _, i in self._info_agent], dtype=np.int32)
        s._check(s._L.mg_info_configure(s._h, int(g.size), g.ctypes.data if g.size else None, int(a.size), a.ctypes.data if a.size else None))
        N, A = self.num_envs, self.agents_per_env
        self.info_game = torch.zeros((N, max(len(self._info_game), 1)), dtype=torch.float32, device=s.device)
        self.info_game_present = torch.zeros((N, max(len(self._info_game), 1)), dtype=torch.uint8, device=s.device)
        self.info_agent = torch.zeros((N, A, max(len(self._info_agent), 1)), dtype=torch.float32, device=s.device)
        self.info_agent_present = torch.zeros((N, A, max(len(self._info_agent), 1)), dtype=torch.uint8, device=s.device)

    @property
    def has_step_info(self) -> bool:
        return bool(self._info_game or self._info_agent or self._info_attr)

    def gather_step_info(self) -> dict:
        """Device tensors of the configured stats after the latest step (asynchronous; no host synchronisation):
        {'game': [N, Kg], 'game_present', 'game_keys', 'agent': [N, A, Ka], 'agent_present', 'agent_keys'}."""
        s = self.sim
        s._check(s._L.mg_info_gather(s._h, self.info_game.data_ptr(), self.info_game_present.data_ptr(), self.info_agent.data_ptr(),
                                     self.info_agent_present.data_ptr(), s._stream()))  # fmt: skip
        return {"game": self.info_game, "game_present": self.info_game_present, "game_keys": [k for k, _ in self._info_game],
                "agent": self.info_agent, "agent_present": self.info_agent_present, "agent_keys": [k for k, _ in self._info_agent]}  # fmt: skip

    def step_info_payload(self, env: int) -> dict:
        """The reference's info payload for ONE env as a host dict (_build_step_info_payload, :230-282) -- for tests
        and debugging; training loops read gather_step_info()'s tensors."""
        t = self.gather_step_info()
        torch.cuda.current_stream(self.sim.device).synchronize()
        P = self.sim.program
        gv, gp = t["game"][env].cpu().numpy(), t["game_present"][env].cpu().numpy()
        out: dict = {k: float(gv[i]) for i, (k, _) in enumerate(self._info_game) if gp[i]}
        for raw, attr in self._info_attr:
            out[raw] = float({"seed": int(self.sim.seeds[env]), "map_w": P.width, "map_h": P.height, "max_steps": self._max_steps}[attr])
        if self._info_agent:
            av, ap = t["agent"][env].cpu().numpy(), t["agent_present"][env].cpu().numpy()
            out["_per_agent_infos"] = {a: {k: float(av[a, i]) for i, (k, _) in enumerate(self._info_agent) if ap[a, i]}
                                       for a in range(self.agents_per_env)}  # fmt: skip
        return out

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
        info = self.gather_step_info() if self.has_step_info else {}
        return self.observations, self.rewards, self.terminals, self.truncations, info

    def close(self):
        self.sim.close()
