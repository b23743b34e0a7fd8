"""ctypes binding of the C ABI (include/mettagrid_b200.h).

There is no CPU fallback: if the CUDA extension is missing or cannot be loaded, importing this
module's ``lib()`` raises.
"""

from __future__ import annotations

import ctypes
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ.get("METTAGRID_B200_LIB", _PKG / "libmettagrid_b200.so"))
_lib = None

MG_OK, MG_E_INVALID, MG_E_CUDA, MG_E_UNSUPPORTED, MG_E_ENV = 0, -1, -2, -3, -4

# every symbol include/mettagrid_b200.h declares: (restype, argtypes)
_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
SYMBOLS = {
    "mg_create": (_i, [_vp, _sz, _i, _vp, _vp, _vp, _i, ctypes.POINTER(_vp)]),
    "mg_destroy": (None, [_vp]),
    "mg_last_error": (ctypes.c_char_p, [_vp]),
    "mg_set_buffers": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mg_step": (_i, [_vp, _vp]),
    "mg_step_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mg_reset": (_i, [_vp, _vp, _vp, _vp]),
    "mg_set_map": (_i, [_vp, _i, _vp, _vp]),
    "mg_poll_errors": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "mg_get_episode_rewards": (_i, [_vp, _vp]),
    "mg_get_action_success": (_i, [_vp, _vp]),
    "mg_get_current_steps": (_i, [_vp, _vp]),
    "mg_get_agent_stats": (_i, [_vp, _i, _vp, _vp]),
    "mg_get_game_stats": (_i, [_vp, _i, _vp, _vp]),
    "mg_dump_objects": (_i, [_vp, _i, _vp, _i]),
    "mg_get_agent_state": (_i, [_vp, _i, _vp]),
    "mg_set_inventory": (_i, [_vp, _i, _i, _vp, _vp, _i]),
    "mg_vecenv_configure": (_i, [_vp, _i, _vp, _i, _vp]),
    "mg_vecenv_step": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mg_vecenv_poll": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "mg_token_summary": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "mg_info_configure": (_i, [_vp, _i, _vp, _i, _vp]),
    "mg_info_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mg_grid_obs_configure": (_i, [_vp, _i, _vp]),
    "mg_obs_to_grid": (_i, [_vp, _vp, _i, _vp, _vp]),
    "mg_num_envs": (_i, [_vp]),
    "mg_num_agents": (_i, [_vp]),
    "mg_num_tokens": (_i, [_vp]),
    "mg_state_bytes": (_sz, [_vp]),
    "mg_step_kernel": (_i, [_vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -m mettagrid_b200.build). "
                "mettagrid_b200 has no CPU fallback."
            )
        L = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
