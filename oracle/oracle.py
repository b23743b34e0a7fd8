"""ctypes wrapper around the CPU oracle restatement (oracle/mg_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline leg -- never by the product package ``mettagrid_b200``.
"""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "_build" / "libmg_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _DIR / "mg_oracle.cpp"
    hdr = _DIR.parent / "include" / "mg_program.h"
    if (
        force
        or not _LIB_PATH.exists()
        or _LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime)
    ):
        _LIB_PATH.parent.mkdir(exist_ok=True)
        subprocess.check_call(
            ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", str(src), "-o", str(_LIB_PATH)]
        )
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        vp, i32p, u8p, f32p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float)
        L.mgo_create.restype = vp
        L.mgo_create.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint32, vp]
        L.mgo_destroy.argtypes = [vp]
        L.mgo_step.argtypes = [vp, vp, vp]
        L.mgo_reinit_buffers.argtypes = [vp]
        for name, rt in [("mgo_obs", u8p), ("mgo_rewards", f32p), ("mgo_episode_rewards", f32p), ("mgo_terminals", u8p),
                         ("mgo_truncations", u8p), ("mgo_success", u8p)]:  # fmt: skip
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [vp]
        L.mgo_error.argtypes = [vp]
        L.mgo_error_info.argtypes = [vp]
        L.mgo_current_step.argtypes = [vp]
        L.mgo_current_step.restype = ctypes.c_uint32
        L.mgo_agent_stats.argtypes = [vp, vp, vp]
        L.mgo_game_stats.argtypes = [vp, vp, vp]
        L.mgo_dump_objects.argtypes = [vp, vp, ctypes.c_int]
        L.mgo_agent_state.argtypes = [vp, vp]
        L.mgo_set_inventory.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_int]
        L.mgo_test_shuffle.argtypes = [ctypes.c_uint32, ctypes.c_int, vp, vp]
        _lib = L
    return _lib


def shuffle_kat(seed: int, n: int):
    """(permutation of iota(n), next raw MT19937 output) -- SURVEY H1 known answers."""
    out = np.zeros(n, dtype=np.int32)
    nxt = ctypes.c_uint32(0)
    lib().mgo_test_shuffle(seed, n, out.ctypes.data, ctypes.byref(nxt))
    return out.tolist(), nxt.value


class OracleEnv:
    """One environment stepped by the CPU restatement.  Mirrors the slice of
    ``mettagrid_c.MettaGrid`` the parity tests use (bindings/mettagrid_py.cpp:251-312)."""

    def __init__(self, program, init_cells: np.ndarray, seed: int, init_game_stats: np.ndarray | None = None):
        self.program = program
        self._L = lib()
        blob = np.ascontiguousarray(program.blob, dtype=np.int32)
        cells = np.ascontiguousarray(init_cells, dtype=np.int16)
        gs = None if init_game_stats is None else np.ascontiguousarray(init_game_stats, dtype=np.float32)
        self._h = self._L.mgo_create(blob.ctypes.data, blob.size, cells.ctypes.data, seed & 0xFFFFFFFF, None if gs is None else gs.ctypes.data)
        if not self._h:
            raise RuntimeError("oracle: failed to create environment (bad program or map)")
        self.A = program.num_agents
        self.T = program.num_tokens
        self.actions = np.zeros(self.A, dtype=np.int32)
        self.vibe_actions = np.zeros(self.A, dtype=np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.mgo_destroy(self._h)
            self._h = None

    def step(self, actions=None, vibe_actions=None):
        if actions is not None:
            self.actions[:] = actions
        if vibe_actions is not None:
            self.vibe_actions[:] = vibe_actions
        self._L.mgo_step(self._h, self.actions.ctypes.data, self.vibe_actions.ctypes.data)

    def reinit_buffers(self):
        self._L.mgo_reinit_buffers(self._h)

    def _arr(self, fn, n, dtype):
        return np.ctypeslib.as_array(fn(self._h), shape=(n,)).view(dtype).copy()

    def observations(self):
        return self._arr(self._L.mgo_obs, self.A * self.T * 3, np.uint8).reshape(self.A, self.T, 3)

    def rewards(self):
        return self._arr(self._L.mgo_rewards, self.A, np.float32)

    def episode_rewards(self):
        return self._arr(self._L.mgo_episode_rewards, self.A, np.float32)

    def terminals(self):
        return self._arr(self._L.mgo_terminals, self.A, np.uint8).astype(bool)

    def truncations(self):
        return self._arr(self._L.mgo_truncations, self.A, np.uint8).astype(bool)

    def action_success(self):
        return self._arr(self._L.mgo_success, self.A, np.uint8).astype(bool)

    @property
    def error(self):
        return self._L.mgo_error(self._h)

    @property
    def current_step(self):
        return self._L.mgo_current_step(self._h)

    def stats_arrays(self):
        S, G = len(self.program.agent_stat_names), len(self.program.game_stat_names)
        av, at = np.zeros((self.A, S), np.float32), np.zeros((self.A, S), np.uint8)
        gv, gt = np.zeros(G, np.float32), np.zeros(G, np.uint8)
        self._L.mgo_agent_stats(self._h, av.ctypes.data, at.ctypes.data)
        self._L.mgo_game_stats(self._h, gv.ctypes.data, gt.ctypes.data)
        return av, at, gv, gt

    def get_episode_stats(self):
        """Same shape as the reference's get_episode_stats() (bindings/mettagrid_py.cpp:161-179)."""
        av, at, gv, gt = self.stats_arrays()
        game = {n: float(gv[i]) for i, n in enumerate(self.program.game_stat_names) if gt[i]}
        agents = [
            {n: float(av[a, i]) for i, n in enumerate(self.program.agent_stat_names) if at[a, i]} for a in range(self.A)
        ]
        return {"game": game, "agent": agents}

    def dump_objects(self):
        R = len(self.program.resource_names)
        cap = self.program.hdr("MGH_MAX_OBJECTS")
        out = np.zeros((cap, 8 + 2 * R + self.program.hdr("MGH_TAG_WORDS")), dtype=np.int32)
        n = self._L.mgo_dump_objects(self._h, out.ctypes.data, cap)
        return out[:n]

    def agent_state(self):
        out = np.zeros((self.A, 4), dtype=np.int32)
        self._L.mgo_agent_state(self._h, out.ctypes.data)
        return out

    def grid_objects(self, *a, **k):
        """Same dict as MettaGrid.grid_objects() (bindings/mettagrid_py.cpp:28-139)."""
        from mettagrid_b200.replay import grid_objects

        return grid_objects(self.program, self.dump_objects(), self.agent_state(), *a, **k)

    def set_inventory(self, agent: int, inventory: dict):
        """MettaGrid.set_inventory(agent_id, {resource: amount}) (bindings/mettagrid_py.cpp:203-209)."""
        from mettagrid_b200.compiler import pybind_dict_order

        ids = {self.program.resource_names.index(k): int(v) for k, v in inventory.items()}
        order = pybind_dict_order(list(ids.keys()))
        items = np.asarray(order, dtype=np.int32)
        amounts = np.asarray([ids[k] for k in order], dtype=np.int32)
        self._L.mgo_set_inventory(self._h, agent, items.ctypes.data, amounts.ctypes.data, len(order))
