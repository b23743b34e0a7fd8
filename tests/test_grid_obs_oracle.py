"""The numpy restatement of GridObsWrapper._convert (oracle/grid_obs.py) against the fixture recorded from the
reference's own code (tests/golden/grid_obs.npz, made by tests/golden/make_golden_grid_obs.py)."""

from pathlib import Path

import numpy as np
import pytest

import mettagrid_b200.config as C
from mettagrid_b200.compiler import compile_config
from oracle.grid_obs import convert, scale_table
from tests import golden_cases

GOLD = Path(__file__).resolve().parent / "golden"
GRID_CASES = ["c1_a16", "walled_8way", "combat_3v3", "combat_4v4_base16", "world_3v3"]


def load_case(case):
    fx = np.load(GOLD / "grid_obs.npz")
    step_fx = np.load(GOLD / f"{case}.npz")
    cfg = golden_cases.CASES[case][0](C)
    P = compile_config(cfg, step_fx["grid"].shape[0], step_fx["grid"].shape[1])
    shape = tuple(int(x) for x in fx[f"{case}_shape"])
    want = np.zeros(shape, dtype=np.float32)
    want[tuple(fx[f"{case}_idx"])] = fx[f"{case}_val"]
    return cfg, P, fx[f"{case}_rows"], want


@pytest.mark.parametrize("case", GRID_CASES)
def test_restatement_matches_reference_convert(case):
    _, P, rows, want = load_case(case)
    Cn = max(P.feature_ids.values()) + 1
    got = convert(rows, Cn, P.hdr("MGH_OBS_H"), P.hdr("MGH_OBS_W"), scale_table(P.feature_ids, P.feature_norms))
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_feature_normalisations_follow_id_map():
    # config/id_map.py:161-235
    P = compile_config(golden_cases.CASES["combat_4v4_base16"][0](C), 11, 13)
    n = P.feature_norms
    assert (n["agent:group"], n["episode_completion_pct"], n["last_action"], n["last_reward"], n["goal"]) == (10, 255, 10, 100, 100)
    assert (n["vibe"], n["tag"], n["lp:east"], n["agent_id"]) == (255, 10, 255, 255)
    assert n["inv:hp"] == 16.0 and n["inv:hp:p1"] == 16.0
    assert sorted(P.feature_ids.values()) == list(range(len(P.feature_ids)))
