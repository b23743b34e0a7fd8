"""Config compiler: ``MettaGridConfig`` -> flat int32 game program (``include/mg_program.h``).

This replaces the reference's Python->C++ object lowering
(/root/reference/python/src/mettagrid/config/mettagrid_c_config.py:576-1007,
mettagrid_c_mutations.py:100-269, mettagrid_c_value_config.py:35-99, id_map.py:161-235) with a
lowering to a table the sm_100a step kernel interprets.  It reads config objects by attribute name,
so it accepts both the reference's Pydantic models and ``mettagrid_b200.config`` dataclasses.

The lowering rules are the reference's (file:line cited at each rule); the output format is ours.
"""

from __future__ import annotations

import re
import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np

# ---------------------------------------------------------------------------------------------
# constants: parsed from include/mg_program.h so the C side and this file cannot drift
# ---------------------------------------------------------------------------------------------
_HEADER = Path(__file__).resolve().parent.parent / "include" / "mg_program.h"


def _parse_program_header(path: Path) -> dict[str, int]:
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    consts: dict[str, int] = {}
    for m in re.finditer(r"#define\s+(MG\w+)\s+(0x[0-9A-Fa-f]+|-?\d+)\s*$", text, flags=re.M):
        consts[m.group(1)] = int(m.group(2), 0)
    for m in re.finditer(r"enum\s*\{(.*?)\}\s*;", text, flags=re.S):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, v = [s.strip() for s in item.split("=")]
                nxt = int(v, 0)
            else:
                name = item
            consts[name] = nxt
            nxt += 1
    return consts


K = _parse_program_header(_HEADER)
globals().update(K)  # MGH_*, MGS_*, MGF_*, ... become module constants

_ORIENT = {  # actions/orientation.hpp:6-15
    "north": 0, "south": 1, "west": 2, "east": 3,
    "northwest": 4, "northeast": 5, "southwest": 6, "southeast": 7,
}  # fmt: skip
_DIR_RC = {"north": (-1, 0), "south": (1, 0), "east": (0, 1), "west": (0, -1)}  # mettagrid_c_config.py:165
_TEAM_NAMES = {0: "red", 1: "blue", 2: "green", 3: "yellow", 4: "purple", 5: "orange"}  # :761


def _f32bits(x: float) -> int:
    return struct.unpack("<i", struct.pack("<f", float(x)))[0]


def _ev(x) -> str:
    """Enum or string -> lower-case string."""
    return str(getattr(x, "value", x)).lower()


def _entity(x) -> int:
    s = _ev(x)
    return {"actor": K["MGE_ACTOR"], "target": K["MGE_TARGET"], "source": K["MGE_SOURCE"]}[s]


class CompileError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------
# observation shape (core/observation_shape.cpp:19-66, systems/packed_coordinate.hpp:87-156)
# ---------------------------------------------------------------------------------------------
def _in_shape(dr: int, dc: int, rr: int, cr: int) -> bool:
    if rr == 0 and cr == 0:
        return dr == 0 and dc == 0
    if rr == 0:
        return dr == 0 and abs(dc) <= cr
    if cr == 0:
        return dc == 0 and abs(dr) <= rr
    if rr == cr:
        d2 = dr * dr + dc * dc
        if d2 <= rr * rr:
            return True
        return rr >= 2 and d2 == rr * rr + 1 and (abs(dr) == rr or abs(dc) == cr)
    return dr * dr * cr * cr + dc * dc * rr * rr <= rr * rr * cr * cr


def observation_offsets(obs_h: int, obs_w: int) -> list[tuple[int, int]]:
    """Window offsets in the reference's emission order: Manhattan shells, rows ascending inside a
    shell, -dc before +dc; clipped to the window box and the disc/ellipse shape."""
    rr, cr = obs_h >> 1, obs_w >> 1
    rmin, rmax, cmin, cmax = -(obs_h // 2), obs_h // 2, -(obs_w // 2), obs_w // 2
    out = []
    for d in range(0, rr + cr + 1):
        for dr in range(-d, d + 1):
            rem = d - abs(dr)
            for dc in ([0] if rem == 0 else [-rem, rem]):
                if rmin <= dr <= rmax and cmin <= dc <= cmax and _in_shape(dr, dc, rr, cr):
                    out.append((dr, dc))
    return out


_FAST_BKT = [2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13]
_PRIMES = [17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 103, 109, 113, 127, 137, 139, 149,
           157, 167, 179, 193, 199, 211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541]  # fmt: skip


def pybind_dict_order(keys: list[int]) -> list[int]:
    """Iteration order of the ``std::unordered_map<uint8_t, ...>`` pybind11 builds from a Python dict
    whose keys arrive in ``keys`` order (refined).

    pybind11's map_caster reserves ``len(dict)`` first, so the table has ``next_bkt(n)`` buckets
    (libstdc++ hashtable_c++0x.cc: 2,2,2,3,5,5,7,7,11,11,11,11,13,13 for n < 14, else the next
    prime) and keys collide modulo that count.  libstdc++ links a new node at the global list head
    when its bucket is empty, otherwise at the head of its bucket's chain (hashtable.h
    _M_insert_bucket_begin).  Verified against the reference: initial inventory
    {hp:0, weapon:1, armor:2, mobility:3, energy:5} iterates as 3,2,1,5,0."""
    n = len(keys)
    if n == 0:
        return []
    nb = _FAST_BKT[n] if n < len(_FAST_BKT) else next(p for p in _PRIMES if p >= n)
    order: list[int] = []
    for k in keys:
        b = k % nb
        pos = next((i for i, x in enumerate(order) if x % nb == b), None)
        order.insert(0 if pos is None else pos, k)
    return order


def _digits_needed(max_value: int, base: int) -> int:
    # systems/observation_encoder.hpp:71-85
    n, v = 0, max_value
    while v > 0:
        v //= base
        n += 1
    return max(n, 1)


# ---------------------------------------------------------------------------------------------
@dataclass
class Program:
    """Compiled game program plus the host-side name tables needed to decode results."""

    blob: np.ndarray  # int32
    height: int
    width: int
    num_agents: int
    num_tokens: int
    resource_names: list[str]
    tag_names: list[str]
    vibe_names: list[str]
    action_names: list[str]
    feature_ids: dict[str, int]
    feature_norms: dict[str, float]  # ObservationFeatureSpec.normalization per feature (config/id_map.py:161-235)
    agent_stat_names: list[str]
    game_stat_names: list[str]
    template_names: list[str]  # template index -> canonical cell name
    cell_to_template: dict[str, int]  # every accepted map cell string -> template index
    agent_renames: dict[str, list[str]]  # group cell -> per-agent cells (mettagrid_c_config.py:745-790)
    type_names: list[str]
    features: set = field(default_factory=set)  # engine capabilities this program needs
    object_type_of_template: list[int] = field(default_factory=list)
    objects_stat: dict[str, int] = field(default_factory=dict)  # cell name -> game stat id of objects.<cell>

    def hdr(self, name: str) -> int:
        return int(self.blob[K[name]])

    # -- map encoding ---------------------------------------------------------------------
    def encode_map(self, grid, with_stats: bool = False):
        """String grid [H][W] -> int16 template index per cell (-1 = empty).

        With ``with_stats`` also returns the initial game-stat vector: the reference counts
        ``objects.<cell>`` per created object, keyed by the (renamed) map cell string
        (mettagrid_c.cpp:244).

        Applies the reference's per-agent cell renaming (rename_map_agents,
        mettagrid_c_config.py:549-568): the k-th occurrence, in row-major order, of a group
        cell such as ``agent.red`` becomes ``agent.red.<k>``."""
        grid = np.asarray(grid)
        h, w = grid.shape
        if (h, w) != (self.height, self.width):
            raise CompileError(f"map is {h}x{w} but the program was compiled for {self.height}x{self.width}")
        out = np.full((h, w), -1, dtype=np.int16)
        counters = {k: 0 for k in self.agent_renames}
        gstats = np.zeros(len(self.game_stat_names), dtype=np.float32)
        flat = grid.reshape(-1)
        oflat = out.reshape(-1)
        for i, cell in enumerate(flat.tolist()):
            if cell in ("empty", ".", " "):  # mettagrid_c.cpp:228
                continue
            if cell in counters:
                k = counters[cell]
                names = self.agent_renames[cell]
                if k >= len(names):
                    raise CompileError(f"Map has more '{cell}' cells ({k + 1}) than agents in the group ({len(names)})")
                counters[cell] = k + 1
                cell = names[k]
            t = self.cell_to_template.get(cell)
            if t is None:
                raise CompileError(f"Unknown object type: {cell}")  # mettagrid_c.cpp:232-234
            oflat[i] = t
            if with_stats:
                gstats[self.objects_stat[cell]] += 1.0
        if with_stats:
            return out, gstats
        return out


# ---------------------------------------------------------------------------------------------
class _Builder:
    def __init__(self, game: Any):
        self.g = game
        self.pool: list[int] = []
        self.handlers: list[list[int]] = []
        self.filters: list[list[int]] = []
        self.mutations: list[list[int]] = []
        self.values: list[list[int]] = []
        self.queries: list[list[int]] = []
        self.limits: list[list[int]] = []
        self.aoes: list[list[int]] = []
        self.events: list[list[int]] = []
        self.agent_stats: dict[str, int] = {}
        self.game_stats: dict[str, int] = {}
        self.features: set = set()
        self.dyn_tags: set[int] = set()

    # -- small helpers ---------------------------------------------------------------------
    def plist(self, words) -> int:
        off = len(self.pool)
        self.pool.extend(int(w) for w in words)
        return off

    def astat(self, name: str) -> int:
        return self.agent_stats.setdefault(name, len(self.agent_stats))

    def gstat(self, name: str) -> int:
        return self.game_stats.setdefault(name, len(self.game_stats))

    def tag_mask(self, tag_ids) -> int:
        words = [0] * self.tw
        for t in tag_ids:
            words[t >> 5] |= 1 << (t & 31)
        return self.plist(w - (1 << 32) if w >= (1 << 31) else w for w in words)

    def prefix_tags(self, prefix: str) -> list[int]:
        # mettagrid_c_config.py:74-75 (dict order == sorted tag order)
        return [i for i, n in enumerate(self.tag_names) if n.startswith(prefix)]

    # -- id maps ---------------------------------------------------------------------------
    def build_id_maps(self):
        g = self.g
        self.resource_names = list(g.resource_names)
        if len(self.resource_names) > 13:
            # inventory iteration order is pure most-recent-first only while the
            # libstdc++ unordered_map<uint8_t,...> stays at 13 buckets with no collisions.
            raise CompileError(
                f"{len(self.resource_names)} resources: bit-exact inventory token order is only defined for <= 13"
            )
        self.rid = {n: i for i, n in enumerate(self.resource_names)}
        self.vibes = [v.name for v in g.actions.change_vibe.vibes]
        self.vid = {n: i for i, n in enumerate(self.vibes)}

        # agents list (mettagrid_c_config.py:596-602)
        if g.agents:
            self.agent_cfgs = list(g.agents)
        else:
            self.agent_cfgs = [g.agent] * g.num_agents
            self._default_team = True
        self.explicit_agents = bool(g.agents)

        type_names = {cfg.name for cfg in g.objects.values()} | {a.name for a in self.agent_cfgs}
        self.type_names = sorted(type_names)  # :608-612
        self.type_id = {n: i for i, n in enumerate(self.type_names)}

        tags = set(g.tags) | {mq.tag for mq in g.materialize_queries}  # :631-647
        for cfg in list(g.objects.values()) + self.agent_cfgs:
            tags.update(cfg.tags)
            tags.add(f"type:{cfg.name}")
        self.tag_names = sorted(tags)
        if len(self.tag_names) > 256:
            raise CompileError(f"Too many unique tags ({len(self.tag_names)}). Maximum supported is 256 due to uint8 limit.")
        self.tid = {n: i for i, n in enumerate(self.tag_names)}
        self.tw = max(1, (len(self.tag_names) + 31) // 32)

        # limit name -> resource ids, first definition wins (:620-629)
        self.limit_name_res: dict[str, list[int]] = {}
        for a in self.agent_cfgs:
            for lname, lim in a.inventory.limits.items():
                if lname not in self.limit_name_res:
                    self.limit_name_res[lname] = [self.rid[r] for r in lim.resources if r in self.rid]
        self.default_limit = self.agent_cfgs[0].inventory.default_limit  # :617-618
        self.territory_index = {n: i for i, n in enumerate(g.territories.keys())}  # :656

    def build_feature_ids(self):
        # config/id_map.py:161-235
        g = self.g
        feats: dict[str, int] = {}
        norms: dict[str, float] = {}

        def add(name, norm):
            feats[name] = len(feats)
            norms[name] = float(norm)

        for n, norm in (("agent:group", 10.0), ("episode_completion_pct", 255.0), ("last_action", 10.0),
                        ("last_reward", 100.0), ("goal", 100.0), ("vibe", 255.0), ("tag", 10.0), ("lp:east", 255.0),
                        ("lp:west", 255.0), ("lp:north", 255.0), ("lp:south", 255.0), ("agent_id", 255.0)):  # fmt: skip
            add(n, norm)
        base = g.obs.token_value_base
        # id_map.py:25-38 uses ceil(log_base(65536)); equal to the C++ count for every base >= 2
        self.inv_digits = _digits_needed(65535, base)
        for r in self.resource_names:
            add(f"inv:{r}", base)
            for p in range(1, self.inv_digits):
                add(f"inv:{r}:p{p}", base)
        if g.protocol_details_obs:
            for r in self.resource_names:
                add(f"protocol_input:{r}", 100.0)
            for r in self.resource_names:
                add(f"protocol_output:{r}", 100.0)
        for prefix in g.obs.global_obs.obs:
            add(prefix, base)
            for p in range(1, self.inv_digits):
                add(f"{prefix}:p{p}", base)
        if g.obs.aoe_mask:
            add("aoe_mask", 3.0)
        if g.obs.global_obs.last_action_move:
            add("last_action_move", 1.0)
        if len(feats) > 255:
            raise CompileError(f"{len(feats)} observation features do not fit uint8 feature ids")
        self.feature_ids = feats
        self.feature_norms = norms

    # -- game values (mettagrid_c_value_config.py:35-99) --------------------------------------
    def value(self, gv) -> int:
        if isinstance(gv, (int, float)):
            return self._emit_value([K["MGV_CONST"], 0, _f32bits(gv), 0, 0, 0])
        kind = type(gv).__name__
        scope = K["MGSC_GAME"] if _ev(getattr(gv, "scope", "agent")) == "game" else K["MGSC_AGENT"]
        if kind == "InventoryValue":
            r = self.rid[gv.item]
            # without an actor the reference falls back to the tracker's "<res>.amount" stat
            # (core/game_value.cpp:31-37); that only resolves for GAME scope.
            return self._emit_value([K["MGV_INVENTORY"], scope, r, 0, 0, 0])
        if kind == "StatValue":
            sid = self.gstat(gv.name) if scope == K["MGSC_GAME"] else self.astat(gv.name)
            return self._emit_value([K["MGV_STAT"], scope, sid, int(bool(gv.delta)), 0, 0])
        if kind == "ConstValue":
            return self._emit_value([K["MGV_CONST"], 0, _f32bits(gv.value), 0, 0, 0])
        if kind == "QueryInventoryValue":
            return self._emit_value([K["MGV_QUERY_INVENTORY"], 0, self.rid[gv.item], self.query(gv.query), 0, 0])
        if kind == "QueryCountValue":
            return self._emit_value([K["MGV_QUERY_COUNT"], 0, 0, self.query(gv.query), 0, 0])
        if kind == "SumGameValue":
            kids = [self.value(v) for v in gv.values]
            woff = -1
            if gv.weights is not None and len(gv.weights) > 0:
                if len(gv.weights) != len(kids):
                    raise CompileError("SumGameValue.weights must have same length as values")
                woff = self.plist(_f32bits(w) for w in gv.weights)  # narrowed to float32 like pybind
            return self._emit_value([K["MGV_SUM"], 0, self.plist(kids), len(kids), woff, int(bool(gv.log))])
        if kind == "RatioGameValue":
            return self._emit_value([K["MGV_RATIO"], 0, self.value(gv.numerator), self.value(gv.denominator), 0, 0])
        if kind in ("MaxGameValue", "MinGameValue"):
            kids = [self.value(v) for v in gv.values]
            op = K["MGV_MAX"] if kind == "MaxGameValue" else K["MGV_MIN"]
            return self._emit_value([op, 0, self.plist(kids), len(kids), 0, 0])
        raise CompileError(f"Unknown GameValue type: {type(gv)}")

    def _emit_value(self, words) -> int:
        self.values.append(words)
        return len(self.values) - 1

    # -- queries (mettagrid_c_config.py:83-186) ----------------------------------------------
    def _max_items(self, q) -> int:
        mi = getattr(q, "max_items", None)
        return -1 if mi is None else self.value(float(mi) if isinstance(mi, int) else mi)

    def query(self, q) -> int:
        self.features.add("query")
        if isinstance(q, str):
            from . import config as _c

            q = _c.Query(source=q)
        qt = getattr(q, "query_type", "query")
        if qt == "materialized":
            from . import config as _c

            return self.query(_c.Query(source=q.tag))
        rnd = int(getattr(q, "order_by", None) == "random")
        w = [0] * K["MG_QUERY_WORDS"]
        if qt == "closure":
            src = self.query(q.source)
            cand = self.query(q.candidates)
            e0, en = self.filter_list(q.edge_filters)
            r0, rn = self.filter_list(q.filters)
            w[:9] = [K["MGQ_CLOSURE"], self._max_items(q), rnd, src, cand, e0, en, r0, rn]
        elif qt == "raycast":
            rng_node = self.value(float(q.max_range) if isinstance(q.max_range, int) else q.max_range)
            dirs = [_DIR_RC[d] for d in q.directions]
            b0, bn = self.filter_list(q.blocker, each_must_convert=True)
            src = self.query(q.source)
            w[:10] = [K["MGQ_RAYCAST"], self._max_items(q), 0, src, rng_node,
                      self.plist(x for d in dirs for x in d), len(dirs), b0, bn, int(bool(q.include_blocker))]  # fmt: skip
        else:
            src = q.source
            if not isinstance(src, str):
                inner = self.query(src)
                f0, fn = self.filter_list(q.filters)
                w[:6] = [K["MGQ_FILTERED"], self._max_items(q), rnd, inner, f0, fn]
            else:
                if src not in self.tid:
                    raise CompileError(f"Tag query references unknown tag '{src}'. Add it to GameConfig.tags or object tags.")
                f0, fn = self.filter_list(q.filters)
                w[:6] = [K["MGQ_TAG"], self._max_items(q), rnd, self.tid[src], f0, fn]
        self.queries.append(w)
        return len(self.queries) - 1

    # -- filters (mettagrid_c_config.py:194-386) ----------------------------------------------
    def _lower_filter(self, f) -> list[list[int]] | None:
        """One Python filter -> list of filter nodes (resource filters fan out), or None when the
        reference silently drops it (unknown vibe / tag)."""
        ft = getattr(f, "filter_type", None)
        FW = K["MG_FILTER_WORDS"]

        def node(op, ent=0, a=0, b=0, c=0, d=0):
            return [op, ent, a, b, c, d][:FW]

        if ft == "not":
            kids = self._lower_filter(f.inner) or []
            k0, kn = self._place_filters(kids)
            return [node(K["MGF_NEG"], 0, k0, kn)]
        if ft == "or":
            kids: list[list[int]] = []
            for inner in f.inner:
                low = self._lower_filter(inner)
                if low is None:
                    continue
                if getattr(inner, "filter_type", None) == "resource" and len(low) > 1:
                    # multi-resource inside OR keeps AND semantics via double negation (:376-383)
                    i0, in_ = self._place_filters(low)
                    inner_neg = node(K["MGF_NEG"], 0, i0, in_)
                    j0, jn = self._place_filters([inner_neg])
                    kids.append(node(K["MGF_NEG"], 0, j0, jn))
                else:
                    kids.extend(low)
            k0, kn = self._place_filters(kids)
            return [node(K["MGF_OR"], 0, k0, kn)]
        if ft == "resource":
            out = [node(K["MGF_RESOURCE"], _entity(f.target), self.rid[r], int(m)) for r, m in f.resources.items() if r in self.rid]
            return out or None
        if ft == "vibe":
            if f.vibe not in self.vid:
                return None
            return [node(K["MGF_VIBE"], _entity(f.target), self.vid[f.vibe])]
        if ft == "tag":
            if f.tag not in self.tid:
                return None
            return [node(K["MGF_TAG_PREFIX"], _entity(f.target), self.tag_mask([self.tid[f.tag]]))]
        if ft in ("tag_prefix", "shared_tag_prefix"):
            ids = self.prefix_tags(f.tag_prefix)
            if not ids:
                raise CompileError(f"{type(f).__name__} prefix '{f.tag_prefix}' matched no tags. Available tags: {self.tag_names}")
            if ft == "tag_prefix":
                return [node(K["MGF_TAG_PREFIX"], _entity(f.target), self.tag_mask(ids))]
            return [node(K["MGF_SHARED_TAG_PREFIX"], 0, self.tag_mask(ids))]
        if ft == "max_distance":
            qid = -1 if f.query is None else self.query(f.query)
            return [node(K["MGF_MAX_DISTANCE"], _entity(f.target), qid, int(f.radius))]
        if ft == "game_value":
            mn = f.min
            thr = self.value(float(mn)) if isinstance(mn, int) else self.value(mn)
            self.features.add("game_value")
            return [node(K["MGF_GAME_VALUE"], _entity(f.target), self.value(f.value), thr)]
        if ft == "target_loc_empty":
            return [node(K["MGF_TARGET_LOC_EMPTY"])]
        if ft == "target_is_usable":
            return [node(K["MGF_TARGET_IS_USABLE"])]
        if ft == "periodic":
            start = f.start_on if f.start_on is not None else f.period  # :302-310
            return [node(K["MGF_PERIODIC"], 0, int(f.period), int(start))]
        if ft == "query_resource":
            reqs = [(self.rid[r], int(m)) for r, m in f.requirements.items()]
            return [node(K["MGF_QUERY_RESOURCE"], 0, self.query(f.query), self.plist(x for p in reqs for x in p), len(reqs))]
        return None

    def _place_filters(self, nodes: list[list[int]]) -> tuple[int, int]:
        first = len(self.filters)
        self.filters.extend(nodes)
        return first, len(nodes)

    def filter_list(self, fs, each_must_convert: bool = False) -> tuple[int, int]:
        nodes: list[list[int]] = []
        for f in fs or []:
            low = self._lower_filter(f)
            if low is None:
                if each_must_convert:
                    raise CompileError(f"Failed to convert blocker filter: {f}")
                continue
            nodes.extend(low)
        return self._place_filters(nodes)

    # -- mutations (mettagrid_c_mutations.py:100-269) -----------------------------------------
    def mutation_list(self, ms, context="") -> tuple[int, int]:
        MW = K["MG_MUTATION_WORDS"]
        nodes: list[list[int]] = []

        def node(op, e1=0, e2=0, a=0, b=0, c=0, d=0, e=0):
            nodes.append([op, e1, e2, a, b, c, d, e][:MW])

        for m in ms or []:
            mt = getattr(m, "mutation_type", None)
            if mt == "resource_delta":
                for rname, delta in m.deltas.items():  # one per resource, dict order (:108-119)
                    if rname not in self.rid:
                        raise CompileError(f"ResourceDeltaMutation references unknown resource '{rname}'.")
                    node(K["MGM_RESOURCE_DELTA"], _entity(m.target), 0, self.rid[rname], int(delta))
                self.features.add("inventory")
            elif mt == "resource_transfer":
                for rname, amount in m.resources.items():
                    if rname not in self.rid:
                        raise CompileError(f"ResourceTransferMutation references unknown resource '{rname}'.")
                    node(K["MGM_RESOURCE_TRANSFER"], _entity(m.from_target), _entity(m.to_target), self.rid[rname],
                         int(amount), int(bool(m.remove_source_when_empty)))  # fmt: skip
                    if m.remove_source_when_empty:
                        self.features.add("remove_object")
                self.features.add("inventory")
            elif mt == "clear_inventory":
                if m.limit_name not in self.limit_name_res:
                    raise CompileError(f"ClearInventoryMutation references unknown limit_name '{m.limit_name}'.")
                ids = self.limit_name_res[m.limit_name]
                node(K["MGM_CLEAR_INVENTORY"], _entity(m.target), 0, self.plist(ids), len(ids))
                self.features.add("inventory")
            elif mt == "stats":
                is_agent = _ev(m.target) == "agent"
                sid = self.astat(m.stat) if is_agent else self.gstat(m.stat)
                vnode = self.value(m.source)
                # logStat lowers to "stat := Sum([stat, const])" (mutation/stats_mutation.py:62-86): an increment.  With
                # an integer constant the float adds commute exactly (below 2^24), so an engine may apply them
                # atomically in any order; word 7 marks the pattern, word 2 carries the constant's bits.
                inc, bits = 0, 0
                v = self.values[vnode]
                if v[0] == K["MGV_SUM"] and v[3] == 2 and v[4] < 0 and not v[5]:
                    k0, k1 = self.values[self.pool[v[2]]], self.values[self.pool[v[2] + 1]]
                    want_scope = K["MGSC_AGENT"] if is_agent else K["MGSC_GAME"]
                    if (k0[0] == K["MGV_STAT"] and k0[1] == want_scope and k0[2] == sid and not k0[3] and k1[0] == K["MGV_CONST"]):
                        c = struct.unpack("<f", struct.pack("<I", int(k1[2]) & 0xFFFFFFFF))[0]
                        if float(c).is_integer() and abs(float(c)) <= 1024:
                            inc, bits = 1, int(k1[2])
                node(K["MGM_STATS"], 0, bits, sid, int(is_agent), int(_ev(m.entity) == "actor"), vnode, inc)
                self.features.add("game_value")
            elif mt == "add_tag":
                if m.tag not in self.tid:
                    raise CompileError(f"AddTagMutation references unknown tag '{m.tag}'.")
                node(K["MGM_ADD_TAG"], _entity(m.target), 0, self.tid[m.tag])
                self.dyn_tags.add(self.tid[m.tag])
                self.features.add("tags")
            elif mt == "remove_tag":
                if m.tag not in self.tid:
                    raise CompileError(f"RemoveTagMutation references unknown tag '{m.tag}'.")
                node(K["MGM_REMOVE_TAG"], _entity(m.target), 0, self.tid[m.tag])
                self.features.add("tags")
            elif mt == "remove_tags_with_prefix":
                ids = self.prefix_tags(m.prefix)
                node(K["MGM_REMOVE_TAGS_PREFIX"], _entity(m.target), 0, self.plist(ids), len(ids))
                self.features.add("tags")
            elif mt == "change_vibe":
                node(K["MGM_CHANGE_VIBE"], _entity(m.target), 0, self.vid.get(m.vibe_name, 0))
            elif mt == "set_game_value":
                src = m.source if m.source is not None else float(m.delta)
                node(K["MGM_GAME_VALUE"], _entity(m.target), 0, self.value(m.value), self.value(src))
                self.features.add("game_value")
                self.features.add("inventory")
            elif mt == "recompute_materialized_query":
                ids = self.prefix_tags(m.tag_prefix)
                if not ids:
                    raise CompileError(f"RecomputeMaterializedQueryMutation prefix '{m.tag_prefix}' matched no tags.")
                for t in ids:
                    node(K["MGM_RECOMPUTE_MQ"], 0, 0, t)
                self.features.add("materialized_query")
            elif mt == "query_inventory":
                qid = self.query(m.query)
                for r in m.deltas:
                    if r not in self.rid:
                        raise CompileError(f"QueryInventoryMutation in {context} references unknown resource '{r}'.")
                pairs = [(self.rid[r], int(d)) for r, d in m.deltas.items()]
                has_src = m.source is not None
                stat_off = -1
                if m.transfer_stats:
                    ids = [-1] * len(self.resource_names)
                    for r, sname in m.transfer_stats.items():
                        ids[self.rid[r]] = self.gstat(sname)
                    stat_off = self.plist(ids)
                node(K["MGM_QUERY_INVENTORY"], _entity(m.source) if has_src else 0, 0, qid,
                     self.plist(x for p in pairs for x in p), len(pairs), int(has_src), stat_off)  # fmt: skip
                self.features.add("inventory")
            elif mt == "spawn_object":
                node(K["MGM_SPAWN_OBJECT"], 0, 0, self._deferred_template(m.object_type))
                self.features.add("spawn")
            elif mt == "raycast_spawn":
                rng_node = self.value(float(m.max_range) if isinstance(m.max_range, int) else m.max_range)
                dirs = [_DIR_RC[d] for d in m.directions]
                b0, bn = self.filter_list(m.blocker, each_must_convert=True)
                node(K["MGM_RAYCAST_SPAWN"], 0, bn, self._deferred_template(m.object_type),
                     self.plist(x for d in dirs for x in d), len(dirs), rng_node, b0)  # fmt: skip
                self.features.add("spawn")
            elif mt == "relocate":
                node(K["MGM_RELOCATE"])
            elif mt == "swap":
                node(K["MGM_SWAP"])
            elif mt == "use_target":
                node(K["MGM_USE_TARGET"])
            elif mt == "push_object":
                node(K["MGM_PUSH_OBJECT"])
            elif mt == "attack":
                # The reference's Python lowering has no AttackMutation branch: dropped.
                continue
            elif mt == "cpp_attack":  # the C++ op itself (attack_mutation.hpp:20-38), reachable only through pybind
                for r in (m.weapon, m.armor, m.health):
                    if r not in self.rid:
                        raise CompileError(f"AttackMutation references unknown resource '{r}'.")
                node(K["MGM_ATTACK"], 0, 0, self.rid[m.weapon], self.rid[m.armor], self.rid[m.health], int(m.damage_multiplier_pct))
                self.features.add("inventory")
            else:
                raise CompileError(f"Unknown mutation type: {type(m)}")
        first = len(self.mutations)
        self.mutations.extend(nodes)
        return first, len(nodes)

    def _deferred_template(self, object_type: str) -> int:
        # spawn mutations look the type up in game_config.objects by map key
        # (spawn_object_mutation.cpp:18-22); resolved after templates exist.
        self._spawn_refs.append((len(self.mutations), object_type))
        return -(2 + len(self._spawn_refs) - 1)  # patched in finalize

    # -- handlers (mettagrid_c_config.py:394-425) ---------------------------------------------
    def simple_handler(self, h, context="") -> int:
        f0, fn = self.filter_list(h.filters)
        m0, mn = self.mutation_list(h.mutations, context)
        self.handlers.append([K["MGHK_SIMPLE"], f0, fn, m0, mn])
        return len(self.handlers) - 1

    def any_handler(self, h) -> int:
        if h is None:
            return -1
        ht = getattr(h, "handler_type", "handler")
        if ht == "handler":
            return self.simple_handler(h, f"handler '{getattr(h, 'name', '')}'")
        kids = [k for k in (self.any_handler(c) for c in h.handlers) if k >= 0]
        if not kids:
            return -1
        kind = K["MGHK_FIRST_MATCH"] if ht == "first_match" else K["MGHK_ALL"]
        self.handlers.append([kind, self.plist(kids), len(kids), 0, 0])
        return len(self.handlers) - 1

    def initial_inventory(self, amounts: dict[int, int]) -> list[tuple[int, int]]:
        """(resource, amount) pairs in token EMISSION order.  The C++ side iterates the pybind-built
        config map and inserts each entry at the front of the (collision-free, 13-bucket) inventory
        map (objects/agent.cpp:79-84), so emission order is that iteration order reversed."""
        return [(k, amounts[k]) for k in reversed(pybind_dict_order(list(amounts.keys())))]

    # -- inventory limits -----------------------------------------------------------------------
    def limit_tables(self, limit_defs: list[tuple[list[int], int, int, dict[int, int]]]):
        """limit_defs in definition order -> (limit_of[R], enforce order, _limits iteration order, modifier mask).

        objects/inventory.cpp:14-26: later definitions overwrite the resource->limit mapping;
        a definition no resource maps to any more is unreachable."""
        R = len(self.resource_names)
        base = len(self.limits)
        limit_of = [-1] * R
        first_seen: list[int] = []
        for li, (res, _mn, _mx, _mods) in enumerate(limit_defs):
            for r in res:
                if limit_of[r] < 0:
                    first_seen.append(r)
                limit_of[r] = base + li
        for res, mn, mx, mods in limit_defs:
            members = []  # filled below once the final mapping is known
            self.limits.append([int(mn), int(mx), self.plist(x for kv in mods.items() for x in kv), len(mods), 0, 0, members])
        reachable = sorted({l for l in limit_of if l >= 0})  # canonical enforce order = definition order
        res_order = list(reversed(first_seen))  # unordered_map iteration: most recently inserted first (H2)
        mod_mask = 0
        for l in reachable:
            lim = self.limits[l]
            members = [r for r in res_order if limit_of[r] == l]
            lim[4], lim[5] = self.plist(members), len(members)
            for it in limit_defs[l - base][3]:
                mod_mask |= 1 << it
        for lim in self.limits[base:]:
            lim.pop()  # drop scratch
        return limit_of, reachable, res_order, mod_mask


# ==================================================================================================
# Compile-time effect analysis: which parts of a tick may run one lane per agent instead of serially.
# The reference resolves everything sequentially (cpp/bindings/mettagrid_c.cpp:966-1052); the kernel may only reorder
# work whose reads and writes provably do not meet.  Everything here is conservative: an op this table does not know
# forces the serial path.
# ==================================================================================================
class _Effects:
    def __init__(self, b: "_Builder", templates: list[list[int]], have_aoe: bool, have_tag_index: bool):
        self.b, self.templates = b, templates
        self.have_aoe, self.have_tag_index = have_aoe, have_tag_index
        self._vc: dict[int, int] = {}
        # game stats some value node reads (other than the "stat" operand of an increment itself): an increment of such a
        # stat must stay ordered with its readers
        own = {b.pool[b.values[m[6]][2]] for m in b.mutations if m[0] == K["MGM_STATS"] and m[7]}
        self.read_gstats = {v[2] for i, v in enumerate(b.values)
                            if v[0] == K["MGV_STAT"] and v[1] == K["MGSC_GAME"] and i not in own}
        tag_remove = [0]
        self._tag_remove_class = 0
        for t in templates:  # on_tag_remove handlers may fire from any tag removal
            off, n = t[K["MGT_TAG_REMOVE"]], t[K["MGT_TAG_REMOVE_N"]]
            for i in range(n):
                tag_remove.append(self.handler(b.pool[off + 2 * i + 1], use=None)[0])
        self._tag_remove_class = max(tag_remove)

    # ---- values / filters --------------------------------------------------------------------------------
    def value(self, node: int) -> int:
        if node in self._vc:
            return self._vc[node]
        self._vc[node] = K["MGC_SERIAL"]  # cycle guard
        op, scope, a, b_, c, d = self.b.values[node]
        pool = self.b.pool
        if op in (K["MGV_INVENTORY"], K["MGV_STAT"]):
            cls = K["MGC_SHARED"] if scope == K["MGSC_GAME"] else K["MGC_LOCAL"]  # game stats: shared words + touched bits
        elif op == K["MGV_CONST"]:
            cls = K["MGC_LOCAL"]
        elif op in (K["MGV_SUM"], K["MGV_MAX"], K["MGV_MIN"]):
            cls = max([self.value(pool[a + i]) for i in range(b_)] + [0])
        elif op == K["MGV_RATIO"]:
            cls = max(self.value(a), self.value(b_))
        else:  # query values read arbitrary objects
            cls = K["MGC_SERIAL"]
        self._vc[node] = cls
        return cls

    def filter(self, fi: int) -> int:
        op, ent, a, b_, c, d = self.b.filters[fi]
        if op in (K["MGF_VIBE"], K["MGF_RESOURCE"], K["MGF_SHARED_TAG_PREFIX"], K["MGF_TAG_PREFIX"], K["MGF_TARGET_LOC_EMPTY"],
                  K["MGF_TARGET_IS_USABLE"], K["MGF_PERIODIC"]):  # fmt: skip
            return K["MGC_LOCAL"]
        if op == K["MGF_GAME_VALUE"]:
            return max(self.value(a), self.value(b_))
        if op in (K["MGF_NEG"], K["MGF_OR"]):
            return max([self.filter(a + i) for i in range(b_)] + [0])
        if op == K["MGF_MAX_DISTANCE"] and a < 0:
            return K["MGC_LOCAL"]
        return K["MGC_SERIAL"]

    def filter_reads_actor_inventory(self, fi: int) -> bool:
        op, ent, a, b_, c, d = self.b.filters[fi]
        if op in (K["MGF_NEG"], K["MGF_OR"]):
            return any(self.filter_reads_actor_inventory(a + i) for i in range(b_))
        return op in (K["MGF_RESOURCE"], K["MGF_GAME_VALUE"]) and ent != K["MGE_TARGET"]

    # ---- mutations / handlers ----------------------------------------------------------------------------
    def mutation(self, mi: int, use) -> tuple[int, bool]:
        """(class, displaces).  `use` = (class, displaces) the USE_TARGET op stands for in this context, or None when the
        target cannot be used (empty cell)."""
        op, e1, e2, a, b_, c, d, e = self.b.mutations[mi]
        L, S, X = K["MGC_LOCAL"], K["MGC_SHARED"], K["MGC_SERIAL"]
        if op in (K["MGM_RESOURCE_DELTA"], K["MGM_CLEAR_INVENTORY"], K["MGM_ATTACK"], K["MGM_CHANGE_VIBE"], K["MGM_RELOCATE"]):
            return L, False
        if op == K["MGM_RESOURCE_TRANSFER"]:
            if c:  # remove_source_when_empty: leaves the tag index and drops AOE presence from every agent inside
                return (X if (self.have_aoe or self.have_tag_index) else S), False
            return L, False
        if op == K["MGM_SWAP"]:
            return L, True
        if op == K["MGM_STATS"]:
            if e and (b_ == 1 or a not in self.read_gstats):  # integer increment nobody reads mid-tick: atomic, order-free
                return L, False
            return max(S if b_ == 0 else L, self.value(d)), False
        if op == K["MGM_GAME_VALUE"]:
            vop, vscope = self.b.values[a][0], self.b.values[a][1]
            tgt = S if vscope == K["MGSC_GAME"] else L
            if vop not in (K["MGV_INVENTORY"], K["MGV_STAT"]):
                tgt = L  # the op ignores other target nodes
            return max(tgt, self.value(b_)), False
        if op in (K["MGM_ADD_TAG"], K["MGM_REMOVE_TAG"], K["MGM_REMOVE_TAGS_PREFIX"]):
            return max(S, self._tag_remove_class), False  # tag index order + insertion stamps
        if op == K["MGM_SPAWN_OBJECT"]:
            return S, False  # object slots / ids / tag index in creation order
        if op == K["MGM_USE_TARGET"]:
            return (L, False) if use is None else use
        return X, False  # queries, materialized queries, raycast spawn, push

    def handler(self, h: int, use) -> tuple[int, bool]:
        if h < 0:
            return K["MGC_LOCAL"], False
        kind, a, b_, c, d = self.b.handlers[h]
        if kind == K["MGHK_SIMPLE"]:
            cls = max([self.filter(a + i) for i in range(b_)] + [0])
            disp = False
            for i in range(d):
                mc_, md = self.mutation(c + i, use)
                cls, disp = max(cls, mc_), disp or md
            return cls, disp
        out = [self.handler(self.b.pool[a + i], use) for i in range(b_)]
        return max([o[0] for o in out] + [0]), any(o[1] for o in out)

    # ---- per-agent world phases -----------------------------------------------------------------------------
    def self_local(self, h: int, target_only: bool, actor_is_agent: bool) -> bool:
        """True when the handler tree only reads class-LOCAL state and only writes the TARGET agent's inventory / stats
        (`target_only`: the actor is another object -- an AOE source or a territory proxy)."""
        if h < 0:
            return True
        kind, a, b_, c, d = self.b.handlers[h]
        if kind != K["MGHK_SIMPLE"]:
            return all(self.self_local(self.b.pool[a + i], target_only, actor_is_agent) for i in range(b_))
        return self.rows_self_local(a, b_, c, d, target_only, actor_is_agent)

    def rows_self_local(self, f0, fn, m0, mn, target_only: bool, actor_is_agent: bool) -> bool:
        for i in range(fn):
            if self.filter(f0 + i) != K["MGC_LOCAL"]:
                return False
            if target_only and actor_is_agent and self.filter_reads_actor_inventory(f0 + i):
                return False  # another lane may be editing that agent's inventory
        for i in range(mn):
            op, e1, e2, a, b_, c, d, e = self.b.mutations[m0 + i]
            tgt_ok = (not target_only) or e1 == K["MGE_TARGET"]
            if op in (K["MGM_RESOURCE_DELTA"], K["MGM_CLEAR_INVENTORY"]) and tgt_ok:
                continue
            if op == K["MGM_CHANGE_VIBE"] and not target_only:
                continue
            if op == K["MGM_STATS"] and b_ == 1 and self.value(d) == K["MGC_LOCAL"] and ((not target_only) or c == 0):
                continue
            if op == K["MGM_GAME_VALUE"] and tgt_ok and self.mutation(m0 + i, None)[0] == K["MGC_LOCAL"]:
                continue
            return False
        return True


def analyze_effects(b: "_Builder", templates, move_chain, territories, n_queries_or_mq: int):
    """-> (tmpl_class rows, move_reach, chain_empty_class, par_flags)"""
    TW = b.tw
    have_aoe = len(b.aoes) > 0
    fx = _Effects(b, templates, have_aoe, n_queries_or_mq > 0)
    agent_templates = [t for t in templates if t[K["MGT_KIND"]] == 1]
    after = [fx.handler(t[K["MGT_ON_AFTER_USE"]], use=None) for t in agent_templates]
    after_cls, after_disp = max([x[0] for x in after] + [0]), any(x[1] for x in after)
    # tags that can appear at run time on any object: a static tag filter cannot rule a template out for those
    addable = set(b.dyn_tags)

    def mask_words(off):
        return [b.pool[off + k] & 0xFFFFFFFF for k in range(TW)]

    def may_apply(h: int, t: list[int]) -> bool:
        kind, a, fn, c, d = b.handlers[h]
        if kind != K["MGHK_SIMPLE"]:
            return True
        static = mask_words(t[K["MGT_TAGS"]])
        for i in range(fn):
            op, ent, fa = b.filters[a + i][0], b.filters[a + i][1], b.filters[a + i][2]
            if op == K["MGF_TAG_PREFIX"] and ent == K["MGE_TARGET"]:
                m = mask_words(fa)
                live = any(m[k] & static[k] for k in range(TW)) or any((m[tg >> 5] >> (tg & 31)) & 1 for tg in addable)
                if not live:
                    return False
        return True

    rows = []
    for t in templates:
        use = None
        if t[K["MGT_ON_USE"]] >= 0:
            uc, ud = fx.handler(t[K["MGT_ON_USE"]], use=None)
            use = (max(uc, after_cls), ud or after_disp)
        cls, disp = (use if use is not None else (K["MGC_LOCAL"], False))  # the built-in [usable -> use_target] handler
        for hid, _rng, _acc, builtin in move_chain:
            if builtin != K["MGMB_GENERIC"] or not may_apply(hid, t):
                continue
            hc, hd = fx.handler(hid, use)
            cls, disp = max(cls, hc), disp or hd
        rows.append([cls | (K["MGC_DISPLACES"] if disp else 0)])
    empty_cls = 0
    for hid, _rng, acc, builtin in move_chain:
        if builtin == K["MGMB_GENERIC"] and acc:
            empty_cls = max(empty_cls, fx.handler(hid, None)[0])
    reach = max([r for _h, r, _a, _b in move_chain] + [1])

    par = K["MGP_ACTIONS"]
    if all(fx.self_local(t[K["MGT_ON_TICK"]], False, True) for t in agent_templates):
        par |= K["MGP_ON_TICK"]
    aoe_ok = True
    for t in templates:
        for k in range(t[K["MGT_AOES"]], t[K["MGT_AOES"]] + t[K["MGT_AOES_N"]]):
            row = b.aoes[k]
            aoe_ok = aoe_ok and fx.rows_self_local(row[3], row[4], row[5], row[6], True, t[K["MGT_KIND"]] == 1)
    for terr in territories:
        for off, n in ((terr[2], terr[3]), (terr[4], terr[5]), (terr[6], terr[7])):
            for i in range(n):
                aoe_ok = aoe_ok and fx.self_local(b.pool[off + i], True, False)
    if aoe_ok:
        par |= K["MGP_AOE"]
    return rows, reach, empty_cls, par


def compile_config(cfg: Any, map_height: int | None = None, map_width: int | None = None, spawn_headroom: int | None = None) -> Program:
    """Compile ``MettaGridConfig`` (or its ``.game``) for a given map size.

    The map size fixes the per-env state layout; if omitted it is read from ``game.map_builder``
    (width/height attributes)."""
    game = getattr(cfg, "game", cfg)
    if map_height is None or map_width is None:
        mb = game.map_builder
        map_height, map_width = int(mb.height), int(mb.width)
    b = _Builder(game)
    b._spawn_refs = []
    b.build_id_maps()
    b.build_feature_ids()
    g = game
    R = len(b.resource_names)
    TW = b.tw
    hdr = [0] * K["MGH_HEADER_WORDS"]

    # ---- well-known stats ------------------------------------------------------------------
    for key, name in [
        ("MGH_ST_ACTION_FAILED", "action.failed"), ("MGH_ST_INVALID_INDEX", "action.invalid_index"),
        ("MGH_ST_MAX_SWM", "status.max_steps_without_motion"), ("MGH_ST_CELL_VISITED", "cell.visited"),
        ("MGH_ST_UNIQUE_VISITED", "cell.unique_visited"), ("MGH_ST_MAX_DIST", "cell.max_distance_from_spawn"),
        ("MGH_ST_DEATH", "death"), ("MGH_ST_ACTIONS_SWAP", "actions.swap"),
        ("MGH_ST_NOOP_SUCCESS", "action.noop.success"), ("MGH_ST_NOOP_FAILED", "action.noop.failed"),
        ("MGH_ST_MOVE_SUCCESS", "action.move.success"), ("MGH_ST_MOVE_FAILED", "action.move.failed"),
        ("MGH_ST_VIBE_SUCCESS", "action.change_vibe.success"), ("MGH_ST_VIBE_FAILED", "action.change_vibe.failed"),
    ]:  # fmt: skip
        hdr[K[key]] = b.astat(name)
    res_stats = []
    for r in b.resource_names:  # objects/agent.cpp:106-121, resource_mutation.hpp:81-85
        res_stats += [b.astat(f"{r}.gained"), b.astat(f"{r}.lost"), b.astat(f"{r}.amount"), b.astat(f"{r}.deposited")]
    for key, name in [("MGH_GST_TOKENS_WRITTEN", "tokens_written"), ("MGH_GST_TOKENS_DROPPED", "tokens_dropped"),
                      ("MGH_GST_TOKENS_FREE", "tokens_free_space")]:  # fmt: skip
        hdr[K[key]] = b.gstat(name)
    res_gstats = [b.gstat(f"{r}.amount") for r in b.resource_names]

    # ---- actions (action_handler_factory.cpp:15-78) --------------------------------------------
    actions: list[list[int]] = [[K["MGA_NOOP"], 0, 0, 0]]
    action_names = ["noop"]
    for d in g.actions.move.allowed_directions:
        if d in _ORIENT:
            actions.append([K["MGA_MOVE"], _ORIENT[d], 0, 0])
            action_names.append(f"move_{d}")
    n_vibes = len(b.vibes) if g.actions.change_vibe.enabled else 0  # mettagrid_c_config.py:973
    for i in range(min(n_vibes, 256)):
        actions.append([K["MGA_CHANGE_VIBE"], i, 0, 1])
        action_names.append(f"change_vibe_{b.vibes[i]}")
    max_priority = 1  # the dead Attack handler still contributes priority 1 (attack.hpp:77)

    # move handler chain: custom handlers, then the two defaults (action_handler_factory.cpp:33-45)
    move_chain: list[list[int]] = []
    for i, h in enumerate(g.actions.move.handlers or []):
        hid = b.simple_handler(h, f"handler '{getattr(h, 'name', '') or f'move_handler_{i}'}'")
        max_range, accepts_empty = 1, 0
        for f in h.filters:  # actions/move.hpp:31-39 (top-level filters only)
            ft = getattr(f, "filter_type", None)
            if ft == "max_distance":
                max_range = f.radius if f.radius > 0 else 1
            if ft == "target_loc_empty":
                accepts_empty = 1
        move_chain.append([hid, int(max_range), accepts_empty, K["MGMB_GENERIC"]])
    f0, fn = b._place_filters([[K["MGF_TARGET_LOC_EMPTY"], 0, 0, 0, 0, 0]])
    m0 = len(b.mutations)
    b.mutations.append([K["MGM_RELOCATE"], 0, 0, 0, 0, 0, 0, 0])
    b.handlers.append([K["MGHK_SIMPLE"], f0, fn, m0, 1])
    move_chain.append([len(b.handlers) - 1, 1, 1, K["MGMB_RELOCATE"]])
    f0, fn = b._place_filters([[K["MGF_TARGET_IS_USABLE"], 0, 0, 0, 0, 0]])
    m0 = len(b.mutations)
    b.mutations.append([K["MGM_USE_TARGET"], 0, 0, 0, 0, 0, 0, 0])
    b.handlers.append([K["MGHK_SIMPLE"], f0, fn, m0, 1])
    move_chain.append([len(b.handlers) - 1, 1, 0, K["MGMB_USE_TARGET"]])

    # ---- templates --------------------------------------------------------------------------------
    templates: list[list[int]] = []
    template_names: list[str] = []
    cell_to_template: dict[str, int] = {}
    max_rewards = 0

    def common_template(cfg_obj, kind: int, limit_defs) -> list[int]:
        t = [0] * K["MG_TEMPLATE_WORDS"]
        t[K["MGT_KIND"]] = kind
        t[K["MGT_TYPE_ID"]] = b.type_id[cfg_obj.name]
        t[K["MGT_VIBE"]] = int(cfg_obj.vibe)
        tag_ids = [b.tid[n] for n in list(cfg_obj.tags) + [f"type:{cfg_obj.name}"]]
        t[K["MGT_TAGS"]] = b.tag_mask(tag_ids)
        limit_of, order, res_order, mod_mask = b.limit_tables(limit_defs)
        t[K["MGT_LIMIT_OF"]] = b.plist(limit_of)
        t[K["MGT_LIMIT_ORDER"]], t[K["MGT_LIMIT_ORDER_N"]] = b.plist(order), len(order)
        t[K["MGT_RES_ORDER"]], t[K["MGT_RES_ORDER_N"]] = b.plist(res_order), len(res_order)
        t[K["MGT_MODIFIER_MASK"]] = mod_mask
        t[K["MGT_ON_USE"]] = b.any_handler(cfg_obj.on_use_handler)
        t[K["MGT_ON_TICK"]] = -1
        t[K["MGT_ON_AFTER_USE"]] = -1
        # AOEs (mettagrid_c_config.py:472-492)
        t[K["MGT_AOES"]] = len(b.aoes)
        for aoe in (cfg_obj.aoes or {}).values():
            af0, afn = b.filter_list(aoe.filters)
            am0, amn = b.mutation_list(aoe.mutations, "AOEConfig")
            pd = []
            for rname, delta in aoe.presence_deltas.items():
                if rname not in b.rid:
                    raise CompileError(f"Unknown resource '{rname}' in AOEConfig presence_deltas")
                pd += [b.rid[rname], int(delta)]
            b.aoes.append([int(aoe.radius), int(bool(aoe.is_static)), int(bool(aoe.effect_self)), af0, afn, am0, amn,
                           b.plist(pd), len(pd) // 2, 0])  # fmt: skip
            b.features.add("aoe")
        t[K["MGT_AOES_N"]] = len(b.aoes) - t[K["MGT_AOES"]]
        terr = []
        for tc in cfg_obj.territory_controls or []:
            if tc.territory not in b.territory_index:
                raise CompileError(f"TerritoryControlConfig references unknown territory '{tc.territory}'.")
            terr += [b.territory_index[tc.territory], int(tc.strength), int(tc.decay)]
            b.features.add("territory")
        t[K["MGT_TERR"]], t[K["MGT_TERR_N"]] = b.plist(terr), len(terr) // 3
        # on_tag_remove: prefix -> handler, fanned out per matching tag (:453-469); per tag the
        # handlers keep dict order.
        pairs = []
        for prefix, h in (cfg_obj.on_tag_remove or {}).items():
            ids = b.prefix_tags(prefix)
            if not ids:
                raise CompileError(f"on_tag_remove prefix '{prefix}' matched no tags.")
            hid = b.simple_handler(h, f"on_tag_remove '{prefix}'")
            pairs += [x for tg in ids for x in (tg, hid)]
            b.features.add("tags")
        t[K["MGT_TAG_REMOVE"]], t[K["MGT_TAG_REMOVE_N"]] = b.plist(pairs), len(pairs) // 2
        return t

    def agent_limit_defs(a) -> list:
        defs, configured = [], set()
        for lim in a.inventory.limits.values():  # mettagrid_c_config.py:684-701
            defs.append(([b.rid[n] for n in lim.resources], lim.base, lim.max,
                         {b.rid[n]: bonus for n, bonus in lim.modifiers.items() if n in b.rid}))  # fmt: skip
            configured.update(lim.resources)
        for rname in b.resource_names:
            if rname not in configured:
                defs.append(([b.rid[rname]], b.default_limit, 65535, {}))
        return defs

    # agents: grouped by team in first-appearance order (:661-790)
    team_groups: dict[int, list] = {}
    for a in b.agent_cfgs:
        team = a.team_id if b.explicit_agents else 0
        team_groups.setdefault(team, []).append(a)
    group_id_of = {team: i for i, team in enumerate(team_groups)}
    agent_renames: dict[str, list[str]] = {}
    for team, members in team_groups.items():
        first_tags = set(members[0].tags)
        for a in members[1:]:
            if set(a.tags) != first_tags:
                raise CompileError(f"All agents in team {team} must have identical tags.")
        gid = group_id_of[team]
        gname = _TEAM_NAMES.get(team, f"group_{gid}")
        canonical = f"agent.{gname}"
        per_agent = []
        for idx, a in enumerate(members):
            t = common_template(a, 1, agent_limit_defs(a))
            t[K["MGT_GROUP"]] = gid
            init = b.initial_inventory({b.rid[k]: int(v) for k, v in a.inventory.initial.items()})
            t[K["MGT_INIT_INV"]], t[K["MGT_INIT_INV_N"]] = b.plist(x for p in init for x in p), len(init)
            rew = []
            for ar in a.rewards.values():  # :671-676
                rew += [b.value(ar.reward), int(bool(ar.per_tick))]
            t[K["MGT_REWARDS"]], t[K["MGT_REWARDS_N"]] = b.plist(rew), len(rew) // 2
            if rew:
                b.features.add("game_value")
            max_rewards = max(max_rewards, len(rew) // 2)
            t[K["MGT_ON_TICK"]] = b.any_handler(a.on_tick)
            t[K["MGT_ON_AFTER_USE"]] = b.any_handler(a.on_after_use_handler)
            name = f"{canonical}.{idx}"
            cell_to_template[name] = len(templates)
            per_agent.append(name)
            templates.append(t)
            template_names.append(name)
        cell_to_template[canonical] = cell_to_template[per_agent[0]]
        if len(members) > 1:
            agent_renames[canonical] = per_agent
        aliases = [f"agent.team_{gid}"]
        if team != gid:
            aliases.append(f"agent.team_{team}")
        if gid in _TEAM_NAMES:
            aliases.append(f"agent.{_TEAM_NAMES[gid]}")
        if team in _TEAM_NAMES and team != gid:
            aliases.append(f"agent.{_TEAM_NAMES[team]}")
        if gid == 0:
            aliases += ["agent.default", "agent.agent"]
        for al in aliases:
            cell_to_template[al] = cell_to_template[canonical]
            if canonical in agent_renames:
                agent_renames[al] = agent_renames[canonical]

    # objects (:794-862)
    for key, oc in g.objects.items():
        is_wall = getattr(oc, "pydantic_type", "object") == "wall"
        defs = []
        init: list = []
        if not is_wall:
            inv = oc.inventory
            configured: set = set()
            for lim in inv.limits.values():
                ids = [b.rid[n] for n in lim.resources if n in b.rid]
                configured.update(lim.resources)
                if ids:
                    defs.append((ids, lim.base, lim.max, {b.rid[n]: bn for n, bn in lim.modifiers.items() if n in b.rid}))
            for rname in inv.initial:
                if rname not in configured and rname in b.rid:
                    defs.append(([b.rid[rname]], inv.default_limit, 65535, {}))
            init = b.initial_inventory({b.rid[k]: int(v) for k, v in inv.initial.items() if k in b.rid})
            init = [p for p in init if p[1] > 0]  # grid_object_factory.cpp:83-87
        t = common_template(oc, 0 if is_wall else 2, defs)
        t[K["MGT_INIT_INV"]], t[K["MGT_INIT_INV_N"]] = b.plist(x for p in init for x in p), len(init)
        if init or defs:
            b.features.add("inventory")
        cell_to_template[oc.map_name] = len(templates)
        templates.append(t)
        template_names.append(oc.map_name)

    # ---- territories, events, materialized queries, obs values, game on_tick ---------------------
    territories = []
    for name, tc in g.territories.items():
        pre = b.prefix_tags(tc.tag_prefix)
        lists = []
        for which in ("on_enter", "on_exit", "presence"):
            hs = [b.simple_handler(h, f"territory '{name}'.{which}.{k}") for k, h in getattr(tc, which).items()]
            lists += [b.plist(hs), len(hs)]
        territories.append([b.plist(pre), len(pre)] + lists)
        b.features.add("territory")
    event_index = {name: i for i, name in enumerate(sorted(g.events.keys()))}  # std::map order
    schedule = []
    for name in sorted(g.events.keys()):
        ev = g.events[name]
        qid = b.query(ev.target_query)
        ef0, efn = b.filter_list(ev.filters)
        em0, emn = b.mutation_list(ev.mutations, f"event '{name}'")
        fb = event_index.get(ev.fallback, -1) if ev.fallback else -1
        b.events.append([qid, -1 if ev.max_targets is None else int(ev.max_targets), ef0, efn, em0, emn, fb, 0])
        schedule += [(int(ts), event_index[name]) for ts in ev.timesteps]
        b.features.add("events")
    schedule.sort(key=lambda p: p[0])  # stable: alphabetical within a timestep (event_scheduler.cpp:11-33)
    mqs = []
    for mq in g.materialize_queries:
        mqs.append([b.tid[mq.tag], b.query(mq.query)])
        b.dyn_tags.add(b.tid[mq.tag])
        b.features.add("materialized_query")
    obs_values = [[b.feature_ids[fname], b.value(gv)] for fname, gv in g.obs.global_obs.obs.items()]
    if obs_values:
        b.features.add("game_value")
    game_on_tick = b.any_handler(g.on_tick)

    # inert template for territory proxy cells (core/territory_tracker.cpp:103-107: a bare GridObject)
    proxy = [0] * K["MG_TEMPLATE_WORDS"]
    proxy[K["MGT_KIND"]] = 3
    proxy[K["MGT_TAGS"]] = b.tag_mask([])
    proxy[K["MGT_LIMIT_OF"]] = b.plist([-1] * R)
    for key in ("MGT_ON_USE", "MGT_ON_TICK", "MGT_ON_AFTER_USE"):
        proxy[K[key]] = -1
    proxy_template = len(templates)
    templates.append(proxy)
    template_names.append("<territory_cell>")

    # spawn references: map key lookup; unknown type -> mutation fails at run time (-1)
    for ref_i, (_mi, otype) in enumerate(b._spawn_refs):
        code = -(2 + ref_i)
        tmpl = cell_to_template.get(otype, -1)
        for m in b.mutations:
            if m[0] in (K["MGM_SPAWN_OBJECT"], K["MGM_RAYCAST_SPAWN"]) and m[3] == code:
                m[3] = tmpl

    tmpl_class, move_reach, chain_empty_class, par_flags = analyze_effects(
        b, templates, move_chain, territories, len(b.queries) + len(mqs))

    # ---- header -------------------------------------------------------------------------------------
    gobs = g.obs.global_obs
    flags = (
        (K["MGG_EPISODE_PCT"] if gobs.episode_completion_pct else 0)
        | (K["MGG_LAST_ACTION"] if gobs.last_action else 0)
        | (K["MGG_LAST_ACTION_MOVE"] if gobs.last_action_move else 0)
        | (K["MGG_LAST_REWARD"] if gobs.last_reward else 0)
        | (K["MGG_LOCAL_POSITION"] if gobs.local_position else 0)
    )
    if g.obs.width > 15 or g.obs.height > 15:  # mettagrid_c.cpp:63-68
        raise CompileError(f"Observation window size ({g.obs.width}x{g.obs.height}) exceeds maximum packable size")
    offsets = observation_offsets(g.obs.height, g.obs.width)
    fid = b.feature_ids
    A = len(b.agent_cfgs)
    n_cells = map_height * map_width
    if spawn_headroom is None:
        spawn_headroom = n_cells if "spawn" in b.features else 0
    # token cache: tags an object can carry + vibe + every resource at full width + group/agent id
    max_static_tags = max([bin(int(x) & 0xFFFFFFFF).count("1") for t in templates
                           for x in [sum((b.pool[t[K["MGT_TAGS"]] + k] & 0xFFFFFFFF) << (32 * k) for k in range(TW))]] + [0])
    max_static_tags = max([sum(bin(b.pool[t[K["MGT_TAGS"]] + k] & 0xFFFFFFFF).count("1") for k in range(TW)) for t in templates] + [0])
    tok_cap = max_static_tags + len(b.dyn_tags) + 1 + R * b.inv_digits + 2
    tok_off = (K["MGO_TAGS"] + TW + (R + 1) // 2 + 3) // 4 * 4  # the token cache starts 16-byte aligned (one vector load)
    # at least four token words per record: the fast path and its pack / unpack kernels move the first eight cached
    # tokens as one 16-byte vector, whatever the count
    obj_stride = (tok_off + max((tok_cap + 1) // 2, 4) + 3) // 4 * 4
    agent_stride = K["MGAG_REWARD_PREV"] + max_rewards
    dyn = sorted(b.dyn_tags)
    dyn_slot = [-1] * len(b.tag_names)
    for i, tg in enumerate(dyn):
        dyn_slot[tg] = i

    H = K
    hdr[H["MGH_MAGIC"]] = K["MG_MAGIC"] - (1 << 32) if K["MG_MAGIC"] >= (1 << 31) else K["MG_MAGIC"]
    hdr[H["MGH_VERSION"]] = K["MG_VERSION"]
    for key, v in [
        ("MGH_H", map_height), ("MGH_W", map_width), ("MGH_NUM_AGENTS", A), ("MGH_NUM_TOKENS", g.obs.num_tokens),
        ("MGH_NUM_RESOURCES", R), ("MGH_NUM_TAGS", len(b.tag_names)), ("MGH_TAG_WORDS", TW),
        ("MGH_TOKEN_BASE", g.obs.token_value_base), ("MGH_INV_DIGITS", b.inv_digits),
        ("MGH_MAX_STEPS", g.max_steps), ("MGH_EPISODE_TRUNCATES", int(bool(g.episode_truncates))),
        ("MGH_OBS_H", g.obs.height), ("MGH_OBS_W", g.obs.width), ("MGH_NUM_OFFSETS", len(offsets)),
        ("MGH_GLOBAL_FLAGS", flags),
        ("MGH_FEAT_GROUP", fid["agent:group"]), ("MGH_FEAT_EPISODE_PCT", fid["episode_completion_pct"]),
        ("MGH_FEAT_LAST_ACTION", fid["last_action"]), ("MGH_FEAT_LAST_REWARD", fid["last_reward"]),
        ("MGH_FEAT_VIBE", fid["vibe"]), ("MGH_FEAT_TAG", fid["tag"]),
        ("MGH_FEAT_LP_EAST", fid["lp:east"]), ("MGH_FEAT_LP_WEST", fid["lp:west"]),
        ("MGH_FEAT_LP_NORTH", fid["lp:north"]), ("MGH_FEAT_LP_SOUTH", fid["lp:south"]),
        ("MGH_FEAT_AGENT_ID", fid["agent_id"]), ("MGH_FEAT_AOE_MASK", fid.get("aoe_mask", 0)),
        ("MGH_FEAT_LAST_ACTION_MOVE", fid.get("last_action_move", 0)),
        ("MGH_NUM_ACTIONS", len(actions)), ("MGH_MAX_PRIORITY", max_priority),
        ("MGH_PRIORITY_MASK", sum(1 << p for p in {a[2] for a in actions})),
        ("MGH_NUM_TEMPLATES", len(templates)), ("MGH_OBJ_STRIDE", obj_stride), ("MGH_AGENT_STRIDE", agent_stride),
        ("MGH_COVER_WORDS", (n_cells + 31) // 32), ("MGH_MAX_REWARDS", max_rewards),
        ("MGH_HP_RESOURCE", b.rid.get("hp", -1)),
        ("MGH_NUM_MOVE_HANDLERS", len(move_chain)), ("MGH_NUM_OBS_VALUES", len(obs_values)),
        ("MGH_NUM_EVENTS_SCHED", len(schedule)), ("MGH_NUM_TERRITORIES", len(territories)), ("MGH_NUM_MQ", len(mqs)),
        ("MGH_GAME_ON_TICK", game_on_tick), ("MGH_NUM_DYN_TAGS", len(dyn)), ("MGH_PROXY_TEMPLATE", proxy_template), ("MGH_TOK_CAP", tok_cap),
        ("MGH_MOVE_REACH", move_reach), ("MGH_CHAIN_EMPTY_CLASS", chain_empty_class), ("MGH_PAR_FLAGS", par_flags),
        ("MGH_SPAWN_CLASS", max([tmpl_class[m[3]][0] & 3 for m in b.mutations if m[0] == K["MGM_SPAWN_OBJECT"] and m[3] >= 0] + [0])),
    ]:  # fmt: skip
        hdr[H[key]] = int(v)
    if g.obs.aoe_mask:
        b.features.add("territory")

    # objects.<cell> game stats for every cell name a map may use (mettagrid_c.cpp:244)
    objects_stat = {name: b.gstat(f"objects.{name}") for name in cell_to_template}

    hdr[H["MGH_NUM_AGENT_STATS"]] = len(b.agent_stats)
    hdr[H["MGH_NUM_GAME_STATS"]] = len(b.game_stats)
    # pool capacity: every cell can hold at most one object; spawned objects get fresh slots.
    hdr[H["MGH_MAX_OBJECTS"]] = n_cells + spawn_headroom + 1
    spawnable = {m[3] for m in b.mutations if m[0] in (K["MGM_SPAWN_OBJECT"], K["MGM_RAYCAST_SPAWN"]) and m[3] >= 0}
    hdr[H["MGH_SPAWN_AOES"]] = max([templates[t][K["MGT_AOES_N"]] for t in spawnable] + [0])
    n_aoe_per_t = max([t[K["MGT_AOES_N"]] for t in templates] + [0])
    n_terr_per_t = max([t[K["MGT_TERR_N"]] for t in templates] + [0])
    hdr[H["MGH_MAX_AOE_SOURCES"]] = (n_cells + spawn_headroom) * n_aoe_per_t if n_aoe_per_t else 0
    hdr[H["MGH_MAX_TERR_SOURCES"]] = (n_cells + spawn_headroom) * n_terr_per_t if n_terr_per_t else 0

    # ---- assemble sections --------------------------------------------------------------------------
    body: list[int] = []

    def section(key: str, rows, width: int | None = None):
        while (K["MGH_HEADER_WORDS"] + len(body)) % 4:  # sections start 16-byte aligned (int4 loads)
            body.append(0)
        hdr[H[key]] = K["MGH_HEADER_WORDS"] + len(body)
        for row in rows:
            if width is not None and len(row) != width:
                raise AssertionError(f"{key}: row width {len(row)} != {width}")
            body.extend(int(x) for x in row)

    inv_feats = [[fid[f"inv:{r}"]] + [fid[f"inv:{r}:p{p}"] for p in range(1, b.inv_digits)] for r in b.resource_names]
    section("MGS_OFFSETS", offsets, 2)
    section("MGS_ACTIONS", actions, K["MG_ACTION_WORDS"])
    section("MGS_MOVE_CHAIN", move_chain, K["MG_MOVEH_WORDS"])
    section("MGS_HANDLERS", b.handlers, K["MG_HANDLER_WORDS"])
    section("MGS_FILTERS", b.filters, K["MG_FILTER_WORDS"])
    section("MGS_MUTATIONS", b.mutations, K["MG_MUTATION_WORDS"])
    section("MGS_VALUES", b.values, K["MG_VALUE_WORDS"])
    section("MGS_QUERIES", b.queries, K["MG_QUERY_WORDS"])
    section("MGS_TEMPLATES", templates, K["MG_TEMPLATE_WORDS"])
    section("MGS_LIMITS", b.limits, K["MG_LIMIT_WORDS"])
    section("MGS_AOES", b.aoes, K["MG_AOE_WORDS"])
    section("MGS_EVENTS", b.events, K["MG_EVENT_WORDS"])
    section("MGS_SCHEDULE", schedule, 2)
    section("MGS_TERRITORIES", territories, K["MG_TERR_WORDS"])
    section("MGS_MQ", mqs, 2)
    section("MGS_OBS_VALUES", obs_values, 2)
    section("MGS_RES_STATS", [res_stats[i * 4 : i * 4 + 4] for i in range(R)], 4)
    section("MGS_RES_GSTATS", [[x] for x in res_gstats], 1)
    section("MGS_INV_FEATS", inv_feats, b.inv_digits)
    section("MGS_DYN_TAGS", [[x] for x in dyn_slot], 1)
    section("MGS_TMPL_CLASS", tmpl_class, 1)
    section("MGS_POOL", [b.pool])
    hdr[H["MGH_TOTAL_WORDS"]] = K["MGH_HEADER_WORDS"] + len(body)

    def to_i32(x: int) -> int:
        x &= 0xFFFFFFFF
        return x - (1 << 32) if x >= (1 << 31) else x

    blob = np.array([to_i32(x) for x in hdr + body], dtype=np.int32)

    agent_stat_names = [n for n, _ in sorted(b.agent_stats.items(), key=lambda kv: kv[1])]
    game_stat_names = [n for n, _ in sorted(b.game_stats.items(), key=lambda kv: kv[1])]
    prog = Program(
        blob=blob, height=map_height, width=map_width, num_agents=A, num_tokens=g.obs.num_tokens,
        resource_names=b.resource_names, tag_names=b.tag_names, vibe_names=b.vibes, action_names=action_names,
        feature_ids=dict(fid), feature_norms=dict(b.feature_norms), agent_stat_names=agent_stat_names, game_stat_names=game_stat_names,
        template_names=template_names, cell_to_template=cell_to_template, agent_renames=agent_renames,
        type_names=b.type_names, features=b.features,
        object_type_of_template=[t[K["MGT_TYPE_ID"]] for t in templates], objects_stat=objects_stat,
    )  # fmt: skip
    return prog
