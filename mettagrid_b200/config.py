"""Host-side mirror of the reference's Python configuration objects.

The step kernels sit behind the reference's own ``MettaGridConfig`` API
(/root/reference/python/src/mettagrid/config/mettagrid_config.py:220-368).  The compiler in
``mettagrid_b200.compiler`` reads configuration objects by ATTRIBUTE NAME, so it accepts the
reference's Pydantic models unchanged when that package is importable.  On a machine without the
reference package (the GPU box) these plain dataclasses provide the same names, fields, defaults and
helper constructors, so configs -- and the tests built on them -- read like the reference's own.

Only the fields the step path consumes are mirrored; rendering, talk, policy
and map-generation options are out of scope.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional, Union

# Names of the reference's default vibe set (config/vibes.py).  Only names and their order matter to
# the step path: they fix the action table (change_vibe_<name>) and vibe ids.
_DEFAULT_VIBE_NAMES = """
default junction carbon_a carbon_b oxygen_a oxygen_b germanium_a germanium_b silicon_a silicon_b
heart_a heart_b gear hub chest wall paperclip up down left right up-right down-right down-left
up-left rotate swords shield wrench money factory lightning fire water tree rotate-clockwise
compass pin pushpin diamond coin oil fuel wheat corn carrot rock mountain wood wave dagger bow
hammer alembic test-tube package backpack zero one two three four five six seven eight nine ten
hash asterisk plus minus multiply divide hundred numbers red-heart orange-heart yellow-heart
green-heart blue-heart purple-heart white-heart black-heart brown-heart two-hearts
sparkling-heart growing-heart heart-arrow heart-ribbon revolving-hearts heart-decoration
broken-heart heart-exclamation love-letter grinning grinning-big-eyes grinning-smiling-eyes
beaming smiling halo heart-eyes star-struck kiss tears-of-joy rofl squinting crying sobbing
crying-cat angry pouting swearing fearful anxious monocle confused sleepy yawning drooling
savoring smirking rolling-eyes clown ghost moai skull-crossbones chart-up chart-down rocket
target red-circle orange-circle yellow-circle green-circle blue-circle purple-circle
brown-circle black-circle white-circle orange-square yellow-square purple-square brown-square
white-square red-triangle blue-diamond small-blue-diamond plug sparkle light-shade medium-shade
""".split()


@dataclass(frozen=True)
class Vibe:
    symbol: str
    name: str
    category: str = "misc"


VIBES = [Vibe("", n) for n in _DEFAULT_VIBE_NAMES]


# --------------------------------------------------------------------------------------------
# enums (reference: config/filter/filter.py, config/mutation/mutation.py, stats_mutation.py)
# --------------------------------------------------------------------------------------------
class HandlerTarget:
    ACTOR = "actor"
    TARGET = "target"


class EntityTarget:
    ACTOR = "actor"
    TARGET = "target"


class StatsTarget:
    GAME = "game"
    AGENT = "agent"


class StatsEntity:
    TARGET = "target"
    ACTOR = "actor"


class Scope:
    AGENT = "agent"
    GAME = "game"


def typeTag(name: str) -> str:
    return f"type:{name}"


def tag(name: str) -> str:
    return name


# --------------------------------------------------------------------------------------------
# game values (reference: config/game_value.py)
# --------------------------------------------------------------------------------------------
@dataclass
class InventoryValue:
    item: str
    scope: str = Scope.AGENT


@dataclass
class StatValue:
    name: str
    scope: str = Scope.AGENT
    delta: bool = False


@dataclass
class ConstValue:
    value: float


@dataclass
class QueryInventoryValue:
    query: Any
    item: str


@dataclass
class QueryCountValue:
    query: Any


@dataclass
class SumGameValue:
    values: list
    weights: Optional[list] = None
    log: bool = False


@dataclass
class RatioGameValue:
    numerator: Any
    denominator: Any


@dataclass
class MaxGameValue:
    values: list


@dataclass
class MinGameValue:
    values: list


def val(x) -> ConstValue:
    return ConstValue(value=float(x))


def _parse_scope(s: str, default: str = Scope.AGENT):
    head, dot, rest = s.partition(".")
    if dot and head.lower() in (Scope.AGENT, Scope.GAME):
        return head.lower(), rest
    return default, s


def inv(s: str) -> InventoryValue:
    scope, name = _parse_scope(s)
    return InventoryValue(item=name, scope=scope)


def stat(s: str, delta: bool = False) -> StatValue:
    scope, name = _parse_scope(s)
    return StatValue(name=name, scope=scope, delta=delta)


def weighted_sum(weighted_values, *, log=False, min=None, max=None):
    out: Any = SumGameValue(values=[v for _, v in weighted_values], weights=[w for w, _ in weighted_values], log=log)
    if min is not None:
        out = MaxGameValue(values=[out, val(min)])
    if max is not None:
        out = MinGameValue(values=[out, val(max)])
    return out


@dataclass
class AgentReward:
    reward: Any = field(default_factory=lambda: val(0.0))
    per_tick: bool = False


def reward(value, *, weight=1.0, log=False, min=None, max=None, per_tick=False) -> AgentReward:
    values = value if isinstance(value, list) else [value]
    return AgentReward(reward=weighted_sum([(weight, v) for v in values], log=log, min=min, max=max), per_tick=per_tick)


def inventoryReward(item, *, weight=1.0, max=None, per_tick=False) -> AgentReward:
    return reward(InventoryValue(item=item), weight=weight, max=max, per_tick=per_tick)


# --------------------------------------------------------------------------------------------
# queries (reference: config/query.py, config/raycast_query.py)
# --------------------------------------------------------------------------------------------
@dataclass
class Query:
    source: Any
    filters: list = field(default_factory=list)
    max_items: Any = None
    order_by: Optional[str] = None
    query_type: str = "query"


@dataclass
class MaterializedQuery:
    tag: str
    query: Any
    source: str = ""
    filters: list = field(default_factory=list)
    max_items: Any = None
    order_by: Optional[str] = None
    query_type: str = "materialized"


@dataclass
class ClosureQuery:
    source: Any
    candidates: Any
    edge_filters: list = field(default_factory=list)
    filters: list = field(default_factory=list)
    max_items: Any = None
    order_by: Optional[str] = None
    query_type: str = "closure"


@dataclass
class RaycastQuery:
    source: Any
    max_range: Any = 2
    directions: list = field(default_factory=lambda: ["north", "south", "east", "west"])
    blocker: list = field(default_factory=list)
    include_blocker: bool = True
    max_items: Any = None
    order_by: Optional[str] = None
    query_type: str = "raycast"


def query(source, filters=None) -> Query:
    return Query(source=source, filters=filters if isinstance(filters, list) else [filters] if filters else [])


def num(s: str, filters=None) -> QueryCountValue:
    fl = filters if isinstance(filters, list) else [filters] if filters is not None else []
    return QueryCountValue(query=query(s, fl))


# --------------------------------------------------------------------------------------------
# filters (reference: config/filter/*.py)
# --------------------------------------------------------------------------------------------
@dataclass
class NotFilter:
    inner: Any
    filter_type: str = "not"


@dataclass
class OrFilter:
    inner: list
    filter_type: str = "or"


@dataclass
class ResourceFilter:
    target: str
    resources: dict = field(default_factory=dict)
    filter_type: str = "resource"


@dataclass
class VibeFilter:
    target: str
    vibe: str
    filter_type: str = "vibe"


@dataclass
class TagFilter:
    target: str
    tag: str
    filter_type: str = "tag"


@dataclass
class TagPrefixFilter:
    target: str
    tag_prefix: str
    filter_type: str = "tag_prefix"


@dataclass
class SharedTagPrefixFilter:
    tag_prefix: str
    filter_type: str = "shared_tag_prefix"


@dataclass
class MaxDistanceFilter:
    target: str = HandlerTarget.TARGET
    query: Any = None
    radius: int = 1
    filter_type: str = "max_distance"


@dataclass
class GameValueFilter:
    target: str
    value: Any
    min: Any = 0
    filter_type: str = "game_value"


@dataclass
class PeriodicFilter:
    period: int
    start_on: Optional[int] = None
    filter_type: str = "periodic"


@dataclass
class TargetIsUsableFilter:
    filter_type: str = "target_is_usable"


@dataclass
class TargetLocEmptyFilter:
    filter_type: str = "target_loc_empty"


def isNot(f) -> NotFilter:
    return NotFilter(inner=f)


def anyOf(filters) -> OrFilter:
    return OrFilter(inner=list(filters))


def actorHas(resources) -> ResourceFilter:
    return ResourceFilter(target=HandlerTarget.ACTOR, resources=dict(resources))


def targetHas(resources) -> ResourceFilter:
    return ResourceFilter(target=HandlerTarget.TARGET, resources=dict(resources))


def actorVibe(vibe) -> VibeFilter:
    return VibeFilter(target=HandlerTarget.ACTOR, vibe=vibe)


def targetVibe(vibe) -> VibeFilter:
    return VibeFilter(target=HandlerTarget.TARGET, vibe=vibe)


def hasTag(t) -> TagFilter:
    return TagFilter(target=HandlerTarget.TARGET, tag=t)


def actorHasTag(t) -> TagFilter:
    return TagFilter(target=HandlerTarget.ACTOR, tag=t)


def isA(type_value) -> TagFilter:
    return hasTag(typeTag(type_value))


def hasTagPrefix(prefix, target=HandlerTarget.TARGET) -> TagPrefixFilter:
    return TagPrefixFilter(target=target, tag_prefix=prefix)


def sharedTagPrefix(prefix) -> SharedTagPrefixFilter:
    return SharedTagPrefixFilter(tag_prefix=prefix)


def maxDistance(radius) -> MaxDistanceFilter:
    return MaxDistanceFilter(target=HandlerTarget.TARGET, radius=radius)


def isNear(q, radius=1) -> MaxDistanceFilter:
    if isinstance(q, str):
        q = Query(source=q)
    return MaxDistanceFilter(target=HandlerTarget.TARGET, query=q, radius=radius)


# --------------------------------------------------------------------------------------------
# mutations (reference: config/mutation/*.py)
# --------------------------------------------------------------------------------------------
@dataclass
class ResourceDeltaMutation:
    target: str
    deltas: dict = field(default_factory=dict)
    mutation_type: str = "resource_delta"


@dataclass
class ResourceTransferMutation:
    from_target: str
    to_target: str
    resources: dict = field(default_factory=dict)
    remove_source_when_empty: bool = False
    mutation_type: str = "resource_transfer"


@dataclass
class ClearInventoryMutation:
    target: str
    limit_name: str
    mutation_type: str = "clear_inventory"


@dataclass
class StatsMutation:
    stat: str
    source: Any
    target: str = StatsTarget.GAME
    entity: str = StatsEntity.TARGET
    mutation_type: str = "stats"


@dataclass
class AddTagMutation:
    tag: str
    target: str = EntityTarget.TARGET
    mutation_type: str = "add_tag"


@dataclass
class RemoveTagMutation:
    tag: str
    target: str = EntityTarget.TARGET
    mutation_type: str = "remove_tag"


@dataclass
class RemoveTagsWithPrefixMutation:
    prefix: str
    target: str = EntityTarget.TARGET
    mutation_type: str = "remove_tags_with_prefix"


@dataclass
class SetGameValueMutation:
    value: Any
    delta: float = 0
    target: str = EntityTarget.ACTOR
    source: Any = None
    mutation_type: str = "set_game_value"


@dataclass
class ChangeVibeMutation:
    target: str = EntityTarget.TARGET
    vibe_name: str = "default"
    mutation_type: str = "change_vibe"


@dataclass
class RecomputeMaterializedQueryMutation:
    tag_prefix: str
    mutation_type: str = "recompute_materialized_query"


@dataclass
class QueryInventoryMutation:
    query: Any
    deltas: dict = field(default_factory=dict)
    source: Optional[str] = None
    transfer_stats: dict = field(default_factory=dict)
    mutation_type: str = "query_inventory"


@dataclass
class SpawnObjectMutation:
    object_type: str
    mutation_type: str = "spawn_object"


@dataclass
class RaycastSpawnMutation:
    object_type: str
    directions: list = field(default_factory=lambda: ["north", "south", "east", "west"])
    max_range: Any = 2
    blocker: list = field(default_factory=list)
    mutation_type: str = "raycast_spawn"


@dataclass
class RelocateMutation:
    mutation_type: str = "relocate"


@dataclass
class SwapMutation:
    mutation_type: str = "swap"


@dataclass
class CppAttackMutation:
    """The C++ engine's AttackMutation (cpp/include/mettagrid/handler/mutations/attack_mutation.hpp:20-38): damage =
    max(0, actor[weapon] * pct / 100 - target[armor]) taken from target[health].  The reference's Python lowering never
    emits it (its own `AttackMutation` config class is dropped), so it has no counterpart there: a handler
    holding one can only be built for the reference through the pybind `add_attack_mutation` (oracle/ref_driver.py)."""

    weapon: str
    armor: str
    health: str
    damage_multiplier_pct: int = 100
    mutation_type: str = "cpp_attack"


@dataclass
class UseTargetMutation:
    mutation_type: str = "use_target"


@dataclass
class PushObjectMutation:
    mutation_type: str = "push_object"


def updateTarget(deltas) -> ResourceDeltaMutation:
    return ResourceDeltaMutation(target=EntityTarget.TARGET, deltas=dict(deltas))


def updateActor(deltas) -> ResourceDeltaMutation:
    return ResourceDeltaMutation(target=EntityTarget.ACTOR, deltas=dict(deltas))


def withdraw(resources, *, remove_when_empty=False) -> ResourceTransferMutation:
    return ResourceTransferMutation(
        from_target=EntityTarget.TARGET,
        to_target=EntityTarget.ACTOR,
        resources=dict(resources),
        remove_source_when_empty=remove_when_empty,
    )


def deposit(resources) -> ResourceTransferMutation:
    return ResourceTransferMutation(from_target=EntityTarget.ACTOR, to_target=EntityTarget.TARGET, resources=dict(resources))


def logStat(stat_name, delta=1, target=StatsTarget.GAME, entity=StatsEntity.TARGET, source=None) -> StatsMutation:
    # reference: mutation/stats_mutation.py:62-86 -- the stat is SET to Sum([stat, value])
    prefix = "game." if target == StatsTarget.GAME else ""
    eff = source if source is not None else val(delta)
    return StatsMutation(
        stat=stat_name, target=target, entity=entity, source=SumGameValue(values=[stat(prefix + stat_name), eff])
    )


def addTag(t, target=EntityTarget.TARGET) -> AddTagMutation:
    return AddTagMutation(tag=t, target=target)


def removeTag(t, target=EntityTarget.TARGET) -> RemoveTagMutation:
    return RemoveTagMutation(tag=t, target=target)


def removeTagPrefix(prefix, target=EntityTarget.TARGET) -> RemoveTagsWithPrefixMutation:
    return RemoveTagsWithPrefixMutation(prefix=prefix, target=target)


def recomputeMaterializedQuery(tag_prefix) -> RecomputeMaterializedQueryMutation:
    return RecomputeMaterializedQueryMutation(tag_prefix=tag_prefix)


def queryDelta(q, deltas) -> QueryInventoryMutation:
    return QueryInventoryMutation(query=q, deltas=dict(deltas))


def queryDeposit(q, resources, stat_prefix="") -> QueryInventoryMutation:
    ts = {r: f"{stat_prefix}{r}.deposited" for r in resources} if stat_prefix else {}
    return QueryInventoryMutation(query=q, deltas=dict(resources), source=EntityTarget.ACTOR, transfer_stats=ts)


def queryWithdraw(q, resources, stat_prefix="") -> QueryInventoryMutation:
    ts = {r: f"{stat_prefix}{r}.withdrawn" for r in resources} if stat_prefix else {}
    return QueryInventoryMutation(
        query=q, deltas={k: -v for k, v in resources.items()}, source=EntityTarget.ACTOR, transfer_stats=ts
    )


# --------------------------------------------------------------------------------------------
# handlers (reference: config/handler_config.py, event_config.py, territory_config.py)
# --------------------------------------------------------------------------------------------
@dataclass
class Handler:
    name: str = ""
    filters: list = field(default_factory=list)
    mutations: list = field(default_factory=list)
    handler_type: str = "handler"


@dataclass
class FirstMatch:
    handlers: list = field(default_factory=list)
    handler_type: str = "first_match"


@dataclass
class AllOf:
    handlers: list = field(default_factory=list)
    handler_type: str = "all_of"


def firstMatch(handlers):
    flat = []
    for h in handlers:
        if h is None:
            continue
        flat.extend(h.handlers) if isinstance(h, FirstMatch) else flat.append(h)
    if not flat:
        return None
    return flat[0] if len(flat) == 1 else FirstMatch(handlers=flat)


def allOf(handlers):
    flat = []
    for h in handlers:
        if h is None:
            continue
        flat.extend(h.handlers) if isinstance(h, AllOf) else flat.append(h)
    if not flat:
        return None
    return flat[0] if len(flat) == 1 else AllOf(handlers=flat)


@dataclass
class AOEConfig:
    name: str = ""
    filters: list = field(default_factory=list)
    mutations: list = field(default_factory=list)
    radius: int = 1
    is_static: bool = True
    effect_self: bool = False
    presence_deltas: dict = field(default_factory=dict)
    handler_type: str = "handler"


@dataclass
class EventConfig:
    target_query: Any
    name: str = ""
    timesteps: list = field(default_factory=list)
    filters: list = field(default_factory=list)
    mutations: list = field(default_factory=list)
    max_targets: Optional[int] = None
    fallback: Optional[str] = None
    handler_type: str = "handler"


def periodic(start, period, end=None, end_period=None) -> list:
    # reference: config/event_config.py:27-68
    if period <= 0:
        raise ValueError(f"period must be positive, got {period}")
    if end is None:
        end = 100000
    if end_period is None:
        return list(range(start, end + 1, period))
    if end_period <= 0:
        raise ValueError(f"end_period must be positive, got {end_period}")
    out, t, span = [], start, end - start
    while t <= end:
        out.append(t)
        if span == 0:
            break
        cur = period + ((t - start) / span) * (end_period - period)
        t += max(1, round(cur))
    return out


def once(timestep) -> list:
    return [timestep]


@dataclass
class TerritoryConfig:
    tag_prefix: str
    on_enter: dict = field(default_factory=dict)
    on_exit: dict = field(default_factory=dict)
    presence: dict = field(default_factory=dict)


@dataclass
class TerritoryControlConfig:
    territory: str
    strength: int = 1
    decay: int = 1


# --------------------------------------------------------------------------------------------
# objects / actions / game (reference: config/mettagrid_config.py, action_config.py, obs_config.py)
# --------------------------------------------------------------------------------------------
@dataclass
class ResourceLimitsConfig:
    base: int
    resources: list
    max: int = 65535
    modifiers: dict = field(default_factory=dict)


@dataclass
class InventoryConfig:
    default_limit: int = 65535
    limits: dict = field(default_factory=dict)
    initial: dict = field(default_factory=dict)


@dataclass
class GridObjectConfig:
    name: str = "object"
    map_name: str = ""
    tags: list = field(default_factory=list)
    vibe: int = 0
    aoes: dict = field(default_factory=dict)
    territory_controls: list = field(default_factory=list)
    inventory: InventoryConfig = field(default_factory=InventoryConfig)
    handlers: dict = field(default_factory=dict)
    on_use_handler: Any = None
    on_tag_remove: dict = field(default_factory=dict)
    pydantic_type: str = "object"

    def __post_init__(self):
        if not self.map_name:
            self.map_name = self.name


@dataclass
class WallConfig(GridObjectConfig):
    name: str = "wall"
    pydantic_type: str = "wall"


@dataclass
class AgentConfig(GridObjectConfig):
    name: str = "agent"
    team_id: int = 0
    rewards: dict = field(default_factory=dict)
    on_tick: Any = None
    on_after_use_handler: Any = None


@dataclass
class ActionConfig:
    action_handler: str = ""
    enabled: bool = True
    required_resources: dict = field(default_factory=dict)
    consumed_resources: dict = field(default_factory=dict)


@dataclass
class NoopActionConfig(ActionConfig):
    action_handler: str = "noop"


CardinalDirections = ["north", "south", "west", "east"]
Directions = ["north", "south", "east", "west", "northeast", "northwest", "southeast", "southwest"]


@dataclass
class MoveActionConfig(ActionConfig):
    action_handler: str = "move"
    allowed_directions: list = field(default_factory=lambda: list(CardinalDirections))
    handlers: list = field(default_factory=list)


@dataclass
class ChangeVibeActionConfig(ActionConfig):
    action_handler: str = "change_vibe"
    vibes: list = field(default_factory=lambda: list(VIBES))


@dataclass
class AttackActionConfig(ActionConfig):
    # Legacy: the reference's Attack handler creates no actions and is never reached from _step
    # (actions/attack.hpp:80-82,122-125); only its priority (=1) is observable.
    action_handler: str = "attack"
    enabled: bool = False


@dataclass
class ActionsConfig:
    noop: NoopActionConfig = field(default_factory=NoopActionConfig)
    move: MoveActionConfig = field(default_factory=MoveActionConfig)
    attack: AttackActionConfig = field(default_factory=AttackActionConfig)
    change_vibe: ChangeVibeActionConfig = field(default_factory=ChangeVibeActionConfig)


@dataclass
class GlobalObsConfig:
    episode_completion_pct: bool = True
    last_action: bool = True
    last_action_move: bool = False
    last_reward: bool = True
    goal_obs: bool = False
    local_position: bool = False
    obs: dict = field(default_factory=dict)


@dataclass
class ObsConfig:
    width: int = 13
    height: int = 13
    token_dim: int = 3
    num_tokens: int = 500
    token_value_base: int = 256
    global_obs: GlobalObsConfig = field(default_factory=GlobalObsConfig)
    aoe_mask: bool = False


_DEFAULT_RESOURCES = [
    "ore_red", "ore_blue", "ore_green", "battery_red", "battery_blue", "battery_green",
    "heart", "armor", "laser", "blueprint",
]  # fmt: skip


@dataclass
class GameConfig:
    resource_names: list = field(default_factory=lambda: list(_DEFAULT_RESOURCES))
    vibe_names: list = field(default_factory=list)
    num_agents: int = 24
    max_steps: int = 10000
    episode_truncates: bool = False
    obs: ObsConfig = field(default_factory=ObsConfig)
    agent: AgentConfig = field(default_factory=AgentConfig)
    agents: list = field(default_factory=list)
    actions: ActionsConfig = field(default_factory=ActionsConfig)
    objects: dict = field(default_factory=dict)
    territories: dict = field(default_factory=dict)
    events: dict = field(default_factory=dict)
    map_builder: Any = None
    protocol_details_obs: bool = True
    tags: list = field(default_factory=list)
    materialize_queries: list = field(default_factory=list)
    on_tick: Any = None

    def __post_init__(self):
        self.vibe_names = [v.name for v in self.actions.change_vibe.vibes]


@dataclass
class AsciiMapConfig:
    """Mirror of AsciiMapBuilder.Config (map_builder/ascii.py:12-52): one character per cell."""

    map_data: list
    char_to_map_name: dict

    def build(self):
        import numpy as np

        rows = [list(line) for line in self.map_data]
        if any(len(r) != len(rows[0]) for r in rows):
            raise ValueError("all lines of an ASCII map must have the same length")
        unknown = {ch for r in rows for ch in r} - set(self.char_to_map_name)
        if unknown:
            raise ValueError(f"Unknown character: {sorted(unknown)!r}. Available: {list(self.char_to_map_name)}")
        return np.array([[self.char_to_map_name[ch] for ch in r] for r in rows], dtype="<U50")


@dataclass
class MettaGridConfig:
    label: str = "mettagrid"
    game: GameConfig = field(default_factory=GameConfig)
    desync_episodes: bool = True

    def with_ascii_map(self, map_data, char_to_map_name) -> "MettaGridConfig":
        """config/mettagrid_config.py:341-346"""
        self.game.map_builder = AsciiMapConfig(map_data=map_data, char_to_map_name=char_to_map_name)
        return self

    @staticmethod
    def EmptyRoom(num_agents: int, width: int = 10, height: int = 10, border_width: int = 1,
                  with_walls: bool = False) -> "MettaGridConfig":  # fmt: skip
        """config/mettagrid_config.py:348-368"""
        from .mapgen import RandomMapConfig

        objects = {"wall": WallConfig()} if border_width > 0 or with_walls else {}
        return MettaGridConfig(
            game=GameConfig(
                map_builder=RandomMapConfig(agents=num_agents, width=width, height=height, border_width=border_width),
                actions=ActionsConfig(move=MoveActionConfig(), change_vibe=ChangeVibeActionConfig()),
                num_agents=num_agents,
                objects=objects,
            )
        )


def closureQuery(source, candidates, edge_filters=None, filters=None) -> ClosureQuery:
    as_list = lambda x: x if isinstance(x, list) else [x] if x else []  # noqa: E731
    return ClosureQuery(source=source, candidates=candidates, edge_filters=as_list(edge_filters), filters=as_list(filters))


def materializedQuery(tag, q) -> MaterializedQuery:
    return MaterializedQuery(tag=tag, query=q)


def raycastQuery(source, max_range=2, directions=None, blocker=None, include_blocker=True) -> RaycastQuery:
    return RaycastQuery(source=source, max_range=max_range, directions=directions or ["north", "south", "east", "west"],
                        blocker=list(blocker or []), include_blocker=include_blocker)  # fmt: skip
