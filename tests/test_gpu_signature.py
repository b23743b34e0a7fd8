"""GPU: the reference's deterministic episode signature through the C ABI (object export, stats, episode rewards,
action_success) -- the hash SURVEY 8c pins."""

import json

import numpy as np
import pytest

from tests import cases
from tests.test_signature_oracle import GOLD, PIN, normalise

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_cuda_step_reproduces_the_signature():
    from mettagrid_b200.replay import signature_hash, signature_payload
    from mettagrid_b200.sim import BatchedSimulation

    cfg = cases.signature_config()
    sim = BatchedSimulation(cfg, 3, seeds=[42, 42, 7])
    noop = sim.program.action_names.index("noop")
    for _ in range(2):
        sim.step(np.full((3, 1), noop, np.int32), np.zeros((3, 1), np.int32))
    torch.cuda.synchronize()
    for env in (0, 1):
        payload = signature_payload(42, sim.current_steps[env], sim.action_success()[env], sim.episode_rewards()[env],
                                    sim.grid_objects(env), sim.get_episode_stats(env))  # fmt: skip
        assert normalise(payload) == json.loads(GOLD.read_text())
        assert signature_hash(payload) == PIN
    sim.close()


def test_grid_objects_match_oracle_mid_episode():
    from mettagrid_b200.sim import BatchedSimulation
    from oracle.oracle import OracleEnv

    cfg = cases.combat_config(None, 3, max_steps=0)
    maps = [cases.combat_map(3, seed=s) for s in range(4)]
    sim = BatchedSimulation(cfg, 4, seeds=11, maps=maps)
    P = sim.program
    oracles = [OracleEnv(P, sim._init_cells[e], int(sim.seeds[e]), sim._init_gstats[e]) for e in range(4)]
    prim, vibe = cases.random_actions(np.random.RandomState(4), 60, (4, 6), 9, len(P.action_names), 0.3, 0.0)
    for t in range(60):
        sim.step(prim[t], vibe[t])
        for e, o in enumerate(oracles):
            o.step(prim[t, e], vibe[t, e])
        if t % 20 == 19:
            for e, o in enumerate(oracles):
                assert sim.grid_objects(e) == o.grid_objects(), f"grid_objects differ at step {t} env {e}"
                assert sim.grid_objects(e, ignore_types=["wall"]) == o.grid_objects(ignore_types=["wall"])
    sim.close()
