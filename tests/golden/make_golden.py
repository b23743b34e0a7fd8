"""Generate tests/golden/*.npz with the REAL reference (build container only).

    python tests/golden/make_golden.py [case ...]

For every case in tests/golden_cases.py the reference's own Python lowering
(convert_to_cpp_game_config) and C++ step (oracle/_ref) are run; stored per case: the map, the action
sequences, one SHA-256 per step over observations || rewards || action_success, the full observation
tensor at step 0 and at the last step, episode rewards and the final stats dict.
"""

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import reference  # noqa: E402

assert reference.load() is not None, "needs /root/reference and oracle/_ref"

from mettagrid.config.mettagrid_c_config import convert_to_cpp_game_config, rename_map_agents  # noqa: E402
from mettagrid.mettagrid_c import MettaGrid  # noqa: E402

from mettagrid_b200.compiler import compile_config  # noqa: E402
from tests import golden_cases as gc  # noqa: E402
from tests import refns  # noqa: E402

ns = refns.reference_namespace()
out_dir = Path(__file__).resolve().parent
only = set(sys.argv[1:])  # optional: names of the cases to (re)record
for name, (mk_cfg, mk_map, seed, steps, _pv, _pi) in gc.CASES.items():
    if only and name not in only:
        continue
    grid = mk_map()
    if name in gc.VIA_DRIVER:  # e.g. the C++ AttackMutation: unreachable from the reference's Python configs
        from mettagrid_b200 import config as C  # noqa: E402
        from oracle.ref_driver import RefEnv  # noqa: E402

        cfg = mk_cfg(C)
        env = RefEnv(cfg, grid, seed).env
    else:
        cfg = mk_cfg(ns)
        c_cfg, renames = convert_to_cpp_game_config(cfg.game)
        env = MettaGrid(c_cfg, rename_map_agents(grid.tolist(), renames), seed)
    prog = compile_config(cfg, *grid.shape)  # only for action-space sizes
    prim, vibe = gc.case_actions(name, prog)
    steps = len(prim)
    obs0 = env.observations().copy()
    digests = np.zeros((steps, 32), dtype=np.uint8)
    for t in range(steps):
        env.actions()[:] = prim[t]
        env.vibe_actions()[:] = vibe[t]
        env.step()
        digests[t] = np.frombuffer(gc.step_digest(env.observations(), env.rewards(), env.action_success()), dtype=np.uint8)
    np.savez_compressed(
        out_dir / f"{name}.npz",
        grid=grid, seed=seed, prim=prim, vibe=vibe, obs0=obs0, obs_last=env.observations().copy(), digests=digests,
        episode_rewards=env.get_episode_rewards().copy(), terminals=env.terminals().copy(), truncations=env.truncations().copy(),
        stats=np.array(gc.stats_json(env.get_episode_stats())),
    )  # fmt: skip
    print(name, steps, "steps ->", (out_dir / f"{name}.npz").stat().st_size, "bytes")
