"""Host-side object export: the dict ``MettaGrid.grid_objects()`` returns
(cpp/bindings/mettagrid_py.cpp:28-139), rebuilt from the engine's object dump, and the payload of the reference's
deterministic episode signature (scripts/deterministic_episode_signature.py:19-50,92-104) on top of it.

Reported per object: id, type_name, location (x = column, y = row), r, c, tag_ids, inventory ({resource id:
amount}, ascending ids like the reference's std::map); per agent additionally agent_id, group_id, vibe,
steps_without_motion, current_stat_reward.  Not reported: the observation-feature echo, inventory_capacities,
last_action_id / last_animation_id and the tag-mutation callables.
"""

from __future__ import annotations

import hashlib
import json
from typing import Sequence

import numpy as np


def grid_objects(program, dump: np.ndarray, agent_state: np.ndarray, min_row: int = -1, max_row: int = -1,
                 min_col: int = -1, max_col: int = -1, ignore_types: Sequence[str] = ()) -> dict:  # fmt: skip
    R = len(program.resource_names)
    TW = program.hdr("MGH_TAG_WORDS")
    use_bounds = min_row >= 0 and max_row >= 0 and min_col >= 0 and max_col >= 0  # :37
    ignore = {program.type_names.index(t) for t in ignore_types if t in program.type_names}  # :40-50
    out: dict = {}
    for row in dump:
        oid, type_id, r, c, vibe, agent = (int(x) for x in row[:6])
        if type_id in ignore:
            continue
        if use_bounds and (r < min_row or r >= max_row or c < min_col or c >= max_col):  # :66-71
            continue
        tags = row[8 + 2 * R : 8 + 2 * R + TW].view(np.uint32)
        d = {
            "id": oid,
            "type_name": program.type_names[type_id],
            "location": (c, r),  # x is the column, y the row (:77-79)
            "r": r,
            "c": c,
            "inventory": {i: int(row[8 + i]) for i in range(R) if row[8 + i] > 0},
            "tag_ids": [32 * w + b for w in range(TW) for b in range(32) if (int(tags[w]) >> b) & 1],
        }
        if agent >= 0:
            st = agent_state[agent]
            d.update(
                agent_id=agent,
                group_id=int(st[1]),
                vibe=vibe,
                steps_without_motion=int(st[2]),
                current_stat_reward=float(st[3:4].view(np.float32)[0]),
            )
        out[oid] = d
    return out


def _round_float(value: float) -> float:
    return round(float(value), 8)


def signature_payload(seed: int, steps: int, action_success, episode_rewards, objects: dict, stats: dict) -> dict:
    """build_signature_payload()'s dict (scripts/deterministic_episode_signature.py:19-50,97-104)."""
    obj_payload = []
    for oid, obj in sorted(objects.items()):
        entry = {
            "id": int(oid),
            "type_name": obj["type_name"],
            "location": [int(obj["r"]), int(obj["c"])],
            "tag_ids": [int(t) for t in obj.get("tag_ids", [])],
            "inventory_items": [(int(k), int(v)) for k, v in obj.get("inventory", {}).items()],
        }
        if "agent_id" in obj:
            entry["agent_id"] = int(obj["agent_id"])
            entry["group_id"] = int(obj["group_id"])
            entry["vibe"] = int(obj["vibe"])
            entry["current_stat_reward"] = _round_float(obj["current_stat_reward"])
        obj_payload.append(entry)
    return {
        "seed": seed,
        "steps": int(steps),
        "action_success": [bool(x) for x in action_success],
        "episode_reward": [_round_float(v) for v in episode_rewards],
        "objects": obj_payload,
        "stats": {  # StatsTracker::to_dict is a std::map: names in ascending order (stats_tracker.hpp:109-115)
            "game": [(n, _round_float(v)) for n, v in sorted(stats["game"].items())],
            "agent": [[(n, _round_float(v)) for n, v in sorted(a.items())] for a in stats["agent"]],
        },
    }


def signature_hash(payload: dict) -> str:
    return hashlib.sha256(json.dumps(payload, sort_keys=True, separators=(",", ":")).encode("utf-8")).hexdigest()
