"""Shared test configurations: the games live in mettagrid_b200/workloads.py (bench.py and smoke() use them too)."""

from mettagrid_b200.workloads import *  # noqa: F401,F403
from mettagrid_b200.workloads import (  # noqa: F401
    COMBAT_RESOURCES, COMBAT_VIBES, EIGHT_WAY, NETWORK_RESOURCES, WORLD_RESOURCES, benchmark_config, combat_config,
    combat_map, network_config, network_map, random_actions, signature_config, walled_config, world_config, world_map,
)  # fmt: skip
