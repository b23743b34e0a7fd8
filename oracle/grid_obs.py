"""CPU restatement of GridObsWrapper._convert (python/src/mettagrid/envs/grid_obs_wrapper.py:58-96).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Pinned by tests/golden/grid_obs.npz, which
tests/golden/make_golden_grid_obs.py recorded by running the reference's own ``_convert``.
"""

from __future__ import annotations

import numpy as np


def scale_table(feature_ids: dict, feature_norms: dict) -> np.ndarray:
    """grid_obs_wrapper.py:41-46: per-feature normalisation indexed by feature id, at least 1."""
    C = max(feature_ids.values(), default=0) + 1
    scale = np.ones(max(256, C), dtype=np.float32)
    for name, fid in feature_ids.items():
        scale[fid] = max(float(feature_norms[name]), 1.0)
    return scale


def convert(raw_obs: np.ndarray, num_features: int, obs_height: int, obs_width: int, scale: np.ndarray) -> np.ndarray:
    """(rows, T, 3) uint8 tokens -> (rows, C, H, W) float32; duplicates accumulate in token order (np.add.at)."""
    rows = raw_obs.shape[0]
    H, W, C = obs_height, obs_width, num_features
    grid = np.zeros((rows, C, H, W), dtype=np.float32)
    coord = raw_obs[..., 0]
    fid = raw_obs[..., 1].astype(np.int32)
    val = raw_obs[..., 2].astype(np.float32)
    y = (coord >> 4) & 0x0F
    x = coord & 0x0F
    is_global = coord == 0xFE  # :74-77 global tokens sit at the window centre
    y = np.where(is_global, H // 2, y)
    x = np.where(is_global, W // 2, x)
    valid = (coord != 0xFF) & (y < H) & (x < W) & (fid < C) & (fid >= 0)  # :80
    val = (val / scale[np.clip(fid, 0, scale.shape[0] - 1)]) * valid  # :83-84
    r = np.broadcast_to(np.arange(rows)[:, None], coord.shape)
    np.add.at(grid, (r, np.clip(fid, 0, C - 1), np.clip(y, 0, H - 1).astype(np.intp), np.clip(x, 0, W - 1).astype(np.intp)), val)
    return grid
