"""k_token_summary (csrc/mg_gridobs.cu) against the CPU restatement of token_encoder.py:89-113.  Floating point: the
kernel adds tokens in row order, torch / numpy reduce the token axis in their own order -- tolerance 1e-5 relative
(plus 1e-6 absolute for sums that cancel)."""

import numpy as np
import pytest

from tests import cases
from tests.test_token_summary_oracle import random_case

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_token_summary_on_random_rows_and_on_live_observations():
    from mettagrid_b200.sim import BatchedSimulation
    from mettagrid_b200.token_encoder import TokenSummary
    from oracle import token_encoder as te

    sim = BatchedSimulation(cases.combat_config(None, 3, num_tokens=60), 5, seeds=2, maps=[cases.combat_map(3, seed=s) for s in range(5)])
    mod = TokenSummary(sim, hidden_size=192)
    px, py, fe = (m.weight.detach().cpu().numpy() for m in (mod.pos_x_embed, mod.pos_y_embed, mod.feature_embed))
    sc = mod._feature_scale.cpu().numpy()
    # adversarial rows: every coordinate / feature / value byte, invalid tokens in the middle, empty rows
    tok, *_ = random_case(5, rows=64, T=60)
    tok[3, 10:20, 0] = 0xFF
    tok[7] = 0xFF
    got = mod(torch.from_numpy(tok).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, te.summary(tok, px, py, fe, sc), rtol=1e-5, atol=1e-6)
    # the handle's own observation buffer after a few ticks (NULL observations = sim.observations)
    prim, vibe = cases.random_actions(np.random.RandomState(1), 10, (5, 6), 9, len(sim.program.action_names), 0.3)
    for t in range(10):
        sim.step(prim[t], vibe[t])
    got = mod().cpu().numpy()
    obs = sim.observations.cpu().numpy().reshape(-1, 60, 3)
    np.testing.assert_allclose(got, te.summary(obs, px, py, fe, sc), rtol=1e-5, atol=1e-6)
    for hidden in (32, 100, 256):  # column counts that are not a multiple of the warp
        m2 = TokenSummary(sim, hidden_size=hidden)
        w = [m.weight.detach().cpu().numpy() for m in (m2.pos_x_embed, m2.pos_y_embed, m2.feature_embed)]
        np.testing.assert_allclose(m2().cpu().numpy(), te.summary(obs, *w, m2._feature_scale.cpu().numpy()), rtol=1e-5, atol=1e-6)
    sim.close()
