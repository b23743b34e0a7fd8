"""The numpy restatement of the token policy's front end against the reference's own lines evaluated with torch
(python/src/mettagrid/policy/token_encoder.py:89-113; the module itself imports pufferlib, which is absent)."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def reference_lines(tokens, pos_x, pos_y, feat, scale):
    """token_encoder.py:89-113 verbatim in behaviour, on CPU tensors."""
    tokens = torch.from_numpy(tokens)
    coords_byte = tokens[..., 0].to(torch.long)
    y_coords, x_coords = (coords_byte >> 4) & 0x0F, coords_byte & 0x0F
    feature_ids = tokens[..., 1].to(torch.long)
    values = tokens[..., 2].to(torch.float32)
    valid_mask = coords_byte != 0xFF
    fc = torch.clamp(feature_ids, min=0, max=feat.shape[0] - 1)
    emb = torch.from_numpy(pos_x)[x_coords] + torch.from_numpy(pos_y)[y_coords] + torch.from_numpy(feat)[fc]
    sc = torch.from_numpy(scale)[fc]
    emb = emb * (values / (sc + 1e-6)).unsqueeze(-1)
    emb = emb * valid_mask.unsqueeze(-1).to(emb.dtype)
    summary = emb.sum(dim=-2)
    counts = valid_mask.sum(dim=-1, keepdim=True).clamp_min(1).to(torch.float32)
    return (summary / torch.sqrt(counts)).numpy()


def random_case(seed, rows=40, T=60, hidden=192, nfeat=256):
    rs = np.random.RandomState(seed)
    tok = rs.randint(0, 256, size=(rows, T, 3)).astype(np.uint8)
    n_valid = rs.randint(0, T + 1, size=rows)
    for r in range(rows):
        tok[r, n_valid[r]:, :] = 0xFF
        tok[r, : n_valid[r], 0] = np.where(rs.rand(n_valid[r]) < 0.1, 0xFE, rs.randint(0, 0xDD, size=n_valid[r]))
    tables = [rs.randn(256, hidden).astype(np.float32) * 0.1, rs.randn(256, hidden).astype(np.float32) * 0.1,
              rs.randn(nfeat, hidden).astype(np.float32) * 0.1]  # fmt: skip
    scale = np.maximum(rs.randint(1, 300, size=nfeat), 1).astype(np.float32)
    return tok, tables[0], tables[1], tables[2], scale


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_numpy_restatement_matches_the_reference_lines(seed):
    from oracle import token_encoder as te

    tok, px, py, fe, sc = random_case(seed)
    want = reference_lines(tok, px, py, fe, sc)
    got = te.summary(tok, px, py, fe, sc)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
