"""Combined-action decoding of the vec-env against a plain restatement of the reference's numpy code
(python/src/mettagrid/envs/mettagrid_puffer_env.py:313-394)."""

import numpy as np
import pytest
import torch

from mettagrid_b200.vecenv import decode_actions


def _ref_decode(actions: np.ndarray, P: int, vibe_ids: np.ndarray):
    V = len(vibe_ids)
    a = actions.astype(np.int64)
    if a.ndim == 2:
        core = a[:, 0]
        vibe = vibe_ids[a[:, 1]] if a.shape[1] == 2 else np.zeros_like(core)
        return core.astype(np.int32), vibe.astype(np.int32)
    enc = a >= P
    core = a.copy()
    vibe = np.zeros_like(a)
    off = a[enc] - P
    core[enc] = off // V
    vibe[enc] = vibe_ids[off % V]
    return core.astype(np.int32), vibe.astype(np.int32)


@pytest.mark.parametrize("P,V", [(5, 152), (9, 3), (1, 1)])
def test_combined_index_decoding(P, V):
    vibe_ids = np.arange(P, P + V)
    rng = np.random.RandomState(0)
    a = rng.randint(0, P + P * V, size=257)
    c1, v1 = _ref_decode(a, P, vibe_ids)
    c2, v2 = decode_actions(torch.from_numpy(a), P, torch.from_numpy(vibe_ids))
    assert np.array_equal(c1, c2.numpy()) and np.array_equal(v1, v2.numpy())
    a2 = np.stack([rng.randint(0, P, size=64), rng.randint(0, V, size=64)], axis=1)
    c1, v1 = _ref_decode(a2, P, vibe_ids)
    c2, v2 = decode_actions(torch.from_numpy(a2), P, torch.from_numpy(vibe_ids))
    assert np.array_equal(c1, c2.numpy()) and np.array_equal(v1, v2.numpy())


def test_decoding_errors_match_reference_messages():
    ids = torch.arange(5, 8)
    with pytest.raises(ValueError, match="non-negative"):
        decode_actions(torch.tensor([-1, 2]), 5, ids)
    with pytest.raises(ValueError, match="out of range"):
        decode_actions(torch.tensor([5 + 5 * 3]), 5, ids)
    with pytest.raises(ValueError, match="no configured vibe action space"):
        decode_actions(torch.tensor([7]), 5, torch.zeros(0, dtype=torch.int64))
    with pytest.raises(ValueError, match="Expected step actions shape"):
        decode_actions(torch.zeros((2, 2, 2), dtype=torch.int64), 5, ids)
