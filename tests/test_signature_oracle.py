"""The reference's deterministic episode signature (scripts/deterministic_episode_signature.py, pinned in SURVEY 8c
as 0a8fe5cd...4d24): the CPU oracle, driven through the same config, must produce the same payload -- object export
(grid_objects), stats, episode rewards, action_success -- and therefore the same hash."""

import json
from pathlib import Path

import numpy as np

from mettagrid_b200.compiler import compile_config
from mettagrid_b200.replay import signature_hash, signature_payload
from tests import cases

PIN = "0a8fe5cd26e34f712ba3035e386418fe428a5cdf268a92f7efd8658d5fef4d24"
GOLD = Path(__file__).resolve().parent / "golden" / "deterministic_signature.json"


def normalise(payload):  # tuples -> lists, like a JSON round trip
    return json.loads(json.dumps(payload, sort_keys=True))


def test_fixture_is_the_pinned_signature():
    assert signature_hash(json.loads(GOLD.read_text())) == PIN


def test_oracle_reproduces_the_signature():
    from oracle.oracle import OracleEnv

    cfg = cases.signature_config()
    grid = cfg.game.map_builder.build()
    P = compile_config(cfg, *grid.shape)
    cells, gstats = P.encode_map(grid, with_stats=True)
    env = OracleEnv(P, cells, 42, gstats)
    noop = P.action_names.index("noop")
    for _ in range(2):
        env.step(np.full(1, noop, np.int32), np.zeros(1, np.int32))
    payload = signature_payload(42, env.current_step, env.action_success(), env.episode_rewards(), env.grid_objects(),
                                env.get_episode_stats())  # fmt: skip
    assert normalise(payload) == json.loads(GOLD.read_text())
    assert signature_hash(payload) == PIN


def test_grid_objects_filters():
    from oracle.oracle import OracleEnv

    cfg = cases.signature_config()
    grid = cfg.game.map_builder.build()
    P = compile_config(cfg, *grid.shape)
    cells, gstats = P.encode_map(grid, with_stats=True)
    env = OracleEnv(P, cells, 42, gstats)
    objs = env.grid_objects()
    assert len(objs) == 27 and objs[19]["type_name"] == "agent" and objs[19]["location"] == (3, 4)
    assert objs[19]["inventory"] == {0: 10, 1: 5} and objs[12]["tag_ids"] == [0, 2]
    no_walls = env.grid_objects(ignore_types=["wall"])
    assert sorted(no_walls) == [11, 12, 13, 16, 19]
    box = env.grid_objects(2, 3, 2, 5)  # rows [2, 3), columns [2, 5)
    assert sorted(box) == [11, 12, 13]
