"""CPU restatement of TokenPolicyNet._encode_tokens up to the pooled summary
(python/src/mettagrid/policy/token_encoder.py:89-113) -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

Plain numpy, float32 like the torch module; the statement order follows the reference line by line.  Pinned in the
build container against the reference's own arithmetic by tests/test_token_summary_oracle.py, which evaluates the same
lines with torch (the module itself needs pufferlib, absent here)."""

from __future__ import annotations

import numpy as np


def summary(tokens: np.ndarray, pos_x: np.ndarray, pos_y: np.ndarray, feat: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """tokens uint8 [rows, T, 3]; pos_x / pos_y float32 [256, H]; feat float32 [F, H]; scale float32 [F] -> float32 [rows, H]"""
    coords = tokens[..., 0].astype(np.int64)  # :92
    x, y = coords & 0x0F, (coords >> 4) & 0x0F  # coordinates(), :24-31
    fid = np.clip(tokens[..., 1].astype(np.int64), 0, feat.shape[0] - 1)  # :94,101
    values = tokens[..., 2].astype(np.float32)  # :95
    valid = coords != 0xFF  # :97
    emb = (pos_x[x] + pos_y[y]) + feat[fid]  # :103
    scaled = (values / (scale[fid] + np.float32(1e-6))).astype(np.float32)[..., None]  # :105-106
    emb = emb * scaled  # :108
    emb = emb * valid[..., None].astype(np.float32)  # :109
    s = emb.sum(axis=-2, dtype=np.float32)  # :111
    counts = np.maximum(valid.sum(axis=-1, keepdims=True), 1).astype(np.float32)  # :112
    return (s / np.sqrt(counts)).astype(np.float32)  # :113
