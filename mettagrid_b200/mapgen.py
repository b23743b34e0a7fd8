"""Host-side map construction for tests and benchmarks.

Map generation is a host pre-step that the step path only consumes (OUT OF
SCOPE).  ``random_map`` follows the reference's ``RandomMapBuilder`` recipe
(python/src/mettagrid/map_builder/random_map.py:27-100) -- border, then a seeded numpy shuffle of the
inner cells -- so the benchmark configs (C1/C2) see the same maps as the reference for the same seed.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class RandomMapConfig:
    """Mirror of RandomMapBuilder.Config (same field names)."""

    width: int = 10
    height: int = 10
    agents: int | dict = 0
    objects: dict = field(default_factory=dict)
    border_width: int = 0
    border_object: str = "wall"
    seed: int | None = None


def random_map(cfg: RandomMapConfig, seed: int | None = None) -> np.ndarray:
    seed = cfg.seed if seed is None else seed
    rng = np.random.default_rng(seed)
    h, w, bw = cfg.height, cfg.width, cfg.border_width
    grid = np.full((h, w), "empty", dtype="<U50")
    if bw > 0:
        grid[:bw, :] = grid[-bw:, :] = cfg.border_object
        grid[:, :bw] = grid[:, -bw:] = cfg.border_object
    ih, iw = max(0, h - 2 * bw), max(0, w - 2 * bw)
    area = ih * iw
    if area <= 0:
        return grid
    if isinstance(cfg.agents, int):
        agents = ["agent.agent"] * cfg.agents
    else:
        agents = ["agent." + name for name, n in cfg.agents.items() for _ in range(n)]
    objects = dict(cfg.objects)
    while sum(objects.values()) + len(agents) > area:
        if all(c <= 1 for c in objects.values()) and len(agents) <= 1:
            break
        objects = {k: max(1, v // 2) for k, v in objects.items()}
    symbols = [name for name, n in objects.items() for _ in range(n)] + agents
    symbols += ["empty"] * (area - len(symbols))
    symbols = np.array(symbols).astype(str)
    rng.shuffle(symbols)
    grid[bw : bw + ih, bw : bw + iw] = symbols.reshape(ih, iw)
    return grid


def ascii_map(lines: list[str], char_to_name: dict[str, str]) -> np.ndarray:
    """Tiny ASCII maps for tests: one character per cell."""
    return np.array([[char_to_name.get(ch, "empty") for ch in row] for row in lines], dtype="<U50")
