"""Record tests/golden/grid_obs.npz by running the REFERENCE's GridObsWrapper._convert
(/root/reference/python/src/mettagrid/envs/grid_obs_wrapper.py:58-96) on token rows taken from the committed
step fixtures.  gymnasium / pufferlib are absent in this image, so the wrapper module is imported with inert
stand-ins for the three names it pulls in at import time; ``_convert`` itself is the reference's code, unmodified.

    python tests/golden/make_golden_grid_obs.py
"""

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
REF = Path("/root/reference/python/src/mettagrid/envs/grid_obs_wrapper.py")


def load_reference_wrapper():
    for name, attrs in {"gymnasium": {}, "gymnasium.spaces": {"Box": object}, "mettagrid": {}, "mettagrid.envs": {},
                        "mettagrid.envs.mettagrid_puffer_env": {"MettaGridPufferEnv": object}, "mettagrid.policy": {},
                        "mettagrid.policy.policy_env_interface": {"PolicyEnvInterface": object}}.items():  # fmt: skip
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("ref_grid_obs_wrapper", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.GridObsWrapper


def main():
    import mettagrid_b200.config as C
    from mettagrid_b200.compiler import compile_config
    from oracle.grid_obs import scale_table
    from tests import golden_cases

    Wrapper = load_reference_wrapper()
    out = {}
    rng = np.random.RandomState(11)
    for case in ("c1_a16", "walled_8way", "combat_3v3", "combat_4v4_base16", "world_3v3"):
        fx = np.load(HERE / f"{case}.npz")
        cfg = golden_cases.CASES[case][0](C)
        grid_map = fx["grid"]
        P = compile_config(cfg, grid_map.shape[0], grid_map.shape[1])
        rows = np.concatenate([fx["obs0"], fx["obs_last"]])
        # adversarial rows: random bytes (out-of-window coordinates, unknown features, padding, many duplicates)
        junk = rng.randint(0, 256, size=(6,) + rows.shape[1:]).astype(np.uint8)
        junk[:3, :, 1] = rng.randint(0, 4, size=junk[:3, :, 1].shape)  # few features: collisions at one cell
        junk[:3, :, 0] = rng.choice([0x00, 0x11, 0x33, 0xFE, 0xFF], size=junk[:3, :, 0].shape)
        rows = np.concatenate([rows, junk])
        w = object.__new__(Wrapper)  # the attributes __init__ derives from the config (grid_obs_wrapper.py:33-48)
        w._obs_height, w._obs_width = P.hdr("MGH_OBS_H"), P.hdr("MGH_OBS_W")
        w._num_features = max(P.feature_ids.values()) + 1
        w._scale = scale_table(P.feature_ids, P.feature_norms)
        w._center_y, w._center_x = w._obs_height // 2, w._obs_width // 2
        grid = w._convert(rows)
        nz = np.nonzero(grid)
        out[f"{case}_rows"] = rows
        out[f"{case}_shape"] = np.asarray(grid.shape)
        out[f"{case}_idx"] = np.stack(nz).astype(np.int32)
        out[f"{case}_val"] = grid[nz]
    np.savez_compressed(HERE / "grid_obs.npz", **out)
    print({k: v.shape for k, v in out.items() if k.endswith("_shape") or k.endswith("_val")})


if __name__ == "__main__":
    main()
