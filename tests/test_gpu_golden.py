"""GPU parity against the golden vectors recorded from the REAL reference: the CUDA step (through the
C ABI) must reproduce every per-step digest, the final observations, rewards and stats."""

import json
from pathlib import Path

import numpy as np
import pytest

from mettagrid_b200 import config as C
from tests import golden_cases as gc

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(gc.CASES))
def test_cuda_matches_reference_golden(name):
    from mettagrid_b200.sim import BatchedSimulation

    g = np.load(Path(__file__).parent / "golden" / f"{name}.npz", allow_pickle=False)
    cfg = gc.CASES[name][0](C)
    # two copies of the env in one batch: both must match, and must not interfere
    sim = BatchedSimulation(cfg, 2, seeds=[int(g["seed"])] * 2, maps=[g["grid"], g["grid"]])
    torch.cuda.synchronize()
    assert np.array_equal(sim.observations.cpu().numpy()[0], g["obs0"])
    prim, vibe = g["prim"], g["vibe"]
    for t in range(len(prim)):
        sim.step(np.stack([prim[t]] * 2), np.stack([vibe[t]] * 2))
        torch.cuda.synchronize()
        obs, rew, succ = sim.observations.cpu().numpy(), sim.rewards.cpu().numpy(), sim.action_success()
        for e in range(2):
            assert gc.step_digest(obs[e], rew[e], succ[e]) == g["digests"][t].tobytes(), f"{name}: step {t} env {e}"
    sim.check_errors()
    assert np.array_equal(sim.observations.cpu().numpy()[1], g["obs_last"])
    assert np.array_equal(sim.episode_rewards()[0].view(np.uint32), g["episode_rewards"].view(np.uint32))
    assert np.array_equal(sim.terminals.cpu().numpy()[0], g["terminals"])
    assert np.array_equal(sim.truncations.cpu().numpy()[0], g["truncations"])
    assert gc.clean_stats(sim.get_episode_stats(1)) == json.loads(str(g["stats"]))
    sim.close()
