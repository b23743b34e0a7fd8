"""Shared test configurations, written with the host-side mirror classes so they run on the GPU box
(where the reference package does not exist).  C1/C2 = the reference benchmark config
(benchmarks/test_mettagrid_env_benchmark.py:21-29)."""

from __future__ import annotations

import numpy as np

from mettagrid_b200 import config as C
from mettagrid_b200.mapgen import RandomMapConfig


def benchmark_config(num_agents: int, map_seed: int = 42, width: int = 20, height: int = 20, num_tokens: int = 100,
                     max_steps: int = 0):  # fmt: skip
    return C.MettaGridConfig(
        game=C.GameConfig(
            num_agents=num_agents,
            obs=C.ObsConfig(num_tokens=num_tokens),
            max_steps=max_steps,
            actions=C.ActionsConfig(noop=C.NoopActionConfig(), move=C.MoveActionConfig()),
            map_builder=RandomMapConfig(agents=num_agents, width=width, height=height, seed=map_seed),
        )
    )


def walled_config(num_agents: int, map_seed: int = 3, width: int = 16, height: int = 12, walls: int = 20,
                  directions=None, max_steps: int = 50, **obs_kw):  # fmt: skip
    mv = C.MoveActionConfig(allowed_directions=list(directions)) if directions else C.MoveActionConfig()
    return C.MettaGridConfig(
        game=C.GameConfig(
            num_agents=num_agents,
            obs=C.ObsConfig(num_tokens=obs_kw.pop("num_tokens", 120), **obs_kw),
            max_steps=max_steps,
            actions=C.ActionsConfig(noop=C.NoopActionConfig(), move=mv),
            objects={"wall": C.WallConfig()},
            map_builder=RandomMapConfig(agents=num_agents, width=width, height=height, seed=map_seed, border_width=1,
                                        objects={"wall": walls}),  # fmt: skip
        )
    )


def random_actions(rng: np.random.RandomState, steps: int, shape, num_primary: int, num_actions: int, p_vibe: float = 0.1,
                   p_invalid: float = 0.0):  # fmt: skip
    """'Effective' action sampling: primary uniform over non-vibe actions, vibe stream with
    probability p_vibe; optionally a few out-of-range indices."""
    prim = rng.randint(0, num_primary, size=(steps,) + tuple(shape)).astype(np.int32)
    vibe = np.zeros_like(prim)
    if num_actions > num_primary:
        m = rng.rand(*prim.shape) < p_vibe
        vibe[m] = rng.randint(num_primary, num_actions, size=int(m.sum()))
    if p_invalid > 0:
        m = rng.rand(*prim.shape) < p_invalid
        prim[m] = rng.choice([-1, num_actions, 9999], size=int(m.sum()))
        m = rng.rand(*prim.shape) < p_invalid
        vibe[m] = rng.choice([-3, num_actions + 5], size=int(m.sum()))
    return prim, vibe


# --------------------------------------------------------------------------------------------------
# C3-style "combat" game built from the primitives the reference actually executes (SURVEY 8a note):
# vibe-gated attack as a move handler, swap with immobile agents, loot transfer, armor branch in a
# FirstMatch, a chest with deposit/withdraw on_use, an altar that converts, on_tick regeneration behind
# a PeriodicFilter ("freeze timer"), inventory limits with a modifier item, inventory/stat rewards.
# `ns` is either mettagrid_b200.config (default) or the reference's classes (tests/refns.py).
# --------------------------------------------------------------------------------------------------
COMBAT_RESOURCES = ["hp", "weapon", "armor", "mobility", "loot", "energy", "pack"]
COMBAT_VIBES = ["default", "swords", "shield"]
EIGHT_WAY = ["north", "south", "west", "east", "northwest", "northeast", "southwest", "southeast"]


def combat_config(ns=None, agents_per_team: int = 3, num_tokens: int = 200, max_steps: int = 0, token_value_base: int = 256):
    if ns is None:
        ns = C
    H, T, A = ns.Handler, ns.HandlerTarget, ns.EntityTarget
    vibes = [ns.Vibe("", n) for n in COMBAT_VIBES]

    attack_armored = H(
        name="attack_armored",
        filters=[ns.actorVibe("swords"), ns.maxDistance(2), ns.actorHas({"weapon": 1}), ns.isA("agent"),
                 ns.isNot(ns.sharedTagPrefix("team:")), ns.targetHas({"armor": 1})],
        mutations=[ns.updateTarget({"armor": -1}), ns.updateActor({"energy": -1}),
                   ns.logStat("attack.blocked", target=ns.StatsTarget.AGENT, entity=ns.StatsEntity.ACTOR)],
    )  # fmt: skip
    attack = H(
        name="attack",
        filters=[ns.actorVibe("swords"), ns.maxDistance(2), ns.actorHas({"weapon": 1}), ns.isA("agent"),
                 ns.isNot(ns.sharedTagPrefix("team:"))],
        mutations=[ns.updateTarget({"hp": -3, "mobility": -2}), ns.withdraw({"loot": -1}),
                   ns.logStat("attack.hits"), ns.logStat("attack.landed", target=ns.StatsTarget.AGENT, entity=ns.StatsEntity.ACTOR)],
    )  # fmt: skip
    swap_immobile = H(
        name="swap_immobile",
        filters=[ns.isA("agent"), ns.isNot(ns.targetHas({"mobility": 1}))],
        mutations=[ns.SwapMutation()],
    )
    need_mobility = H(  # agents without mobility cannot relocate: a failing handler ends the chain early
        name="stuck",
        filters=[ns.TargetLocEmptyFilter(), ns.isNot(ns.actorHas({"mobility": 1}))],
        mutations=[ns.UseTargetMutation()],
    )
    move = ns.MoveActionConfig(allowed_directions=list(EIGHT_WAY), handlers=[attack_armored, attack, swap_immobile, need_mobility])

    def team_agent(team: int, tag: str):
        return ns.AgentConfig(
            team_id=team,
            tags=[tag],
            inventory=ns.InventoryConfig(
                default_limit=50,
                initial={"hp": 10, "weapon": 2, "armor": 1, "mobility": 3, "energy": 5},
                limits={
                    "hp": ns.ResourceLimitsConfig(base=10, resources=["hp"]),
                    "cargo": ns.ResourceLimitsConfig(base=4, max=12, resources=["loot", "energy"], modifiers={"pack": 4}),
                    "mobility": ns.ResourceLimitsConfig(base=3, resources=["mobility"]),
                },
            ),
            rewards={
                "loot": ns.inventoryReward("loot", weight=0.5),
                "hits": ns.reward(ns.stat("attack.landed"), weight=0.25),
                "alive": ns.reward(ns.inv("hp"), weight=0.01, per_tick=True, log=True),
            },
            on_tick=ns.firstMatch([
                H(name="regen", filters=[ns.PeriodicFilter(period=5)], mutations=[ns.updateTarget({"mobility": 1, "energy": 1})]),
                H(name="heal", filters=[ns.PeriodicFilter(period=7, start_on=3), ns.isNot(ns.targetHas({"hp": 10}))],
                  mutations=[ns.updateTarget({"hp": 1})]),
            ]),  # fmt: skip
        )

    chest = ns.GridObjectConfig(
        name="chest",
        inventory=ns.InventoryConfig(initial={"loot": 6, "pack": 1}, limits={"loot": ns.ResourceLimitsConfig(base=20, resources=["loot"])}),
        on_use_handler=ns.firstMatch([
            H(name="deposit", filters=[ns.actorVibe("shield"), ns.actorHas({"loot": 1})], mutations=[ns.deposit({"loot": -1})]),
            H(name="take_pack", filters=[ns.targetHas({"pack": 1})], mutations=[ns.withdraw({"pack": 1})]),
            H(name="withdraw", filters=[ns.targetHas({"loot": 1})], mutations=[ns.withdraw({"loot": 2})]),
        ]),  # fmt: skip
    )
    altar = ns.GridObjectConfig(
        name="altar",
        vibe=2,
        inventory=ns.InventoryConfig(initial={"weapon": 3}),
        on_use_handler=H(
            name="forge",
            filters=[ns.actorHas({"energy": 2}), ns.anyOf([ns.actorVibe("default"), ns.actorVibe("shield")])],
            mutations=[ns.updateActor({"energy": -2, "armor": 1}), ns.updateTarget({"weapon": 1}),
                       ns.ChangeVibeMutation(target=A.ACTOR, vibe_name="swords")],
        ),  # fmt: skip
    )
    agents = [team_agent(0, "team:red") for _ in range(agents_per_team)] + [team_agent(1, "team:blue") for _ in range(agents_per_team)]
    game = ns.GameConfig(
        resource_names=list(COMBAT_RESOURCES),
        num_agents=2 * agents_per_team,
        max_steps=max_steps,
        obs=ns.ObsConfig(width=11, height=11, num_tokens=num_tokens, token_value_base=token_value_base,
                         global_obs=ns.GlobalObsConfig(last_action_move=True)),
        agents=agents,
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=move, change_vibe=ns.ChangeVibeActionConfig(vibes=vibes)),
        objects={"wall": ns.WallConfig(), "chest": chest, "altar": altar},
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def combat_map(agents_per_team: int = 3, width: int = 13, height: int = 11, seed: int = 0) -> np.ndarray:
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    return random_map(RandomMapConfig(width=width, height=height, border_width=1, seed=seed,
                                      agents={"red": agents_per_team, "blue": agents_per_team},
                                      objects={"wall": 8, "chest": 3, "altar": 2}))  # fmt: skip


# --------------------------------------------------------------------------------------------------
# C4-style "world" game: fixed AOEs (tick mutations + presence deltas), one mobile agent aura, a 2-team
# territory with on_enter / presence / on_exit handlers and aoe_mask observation tokens, periodic events
# with max_targets (RNG shuffle), tag queries (isNear, QueryCountValue reward, query-inventory mutation),
# tag add/remove with an on_tag_remove handler, spawn and remove-when-empty, on_tick regeneration.
# --------------------------------------------------------------------------------------------------
WORLD_RESOURCES = ["hp", "energy", "ore", "gem", "shield"]


def world_config(ns=None, agents_per_team: int = 3, num_tokens: int = 200, max_steps: int = 0, spawn: bool = True,
                 aoe_mask: bool = True):  # fmt: skip
    if ns is None:
        ns = C
    H, A = ns.Handler, ns.EntityTarget
    vibes = [ns.Vibe("", n) for n in ["default", "swords", "shield"]]

    def team_agent(team: int, tag: str, aura: bool):
        aoes = {}
        if aura:  # mobile aura: allies nearby gain a shield point while inside, lose it when they leave
            aoes["aura"] = ns.AOEConfig(radius=2, is_static=False, filters=[ns.sharedTagPrefix("team:")],
                                        presence_deltas={"shield": 1})  # fmt: skip
        return ns.AgentConfig(
            team_id=team,
            tags=[tag],
            aoes=aoes,
            inventory=ns.InventoryConfig(
                default_limit=30,
                initial={"hp": 10, "energy": 4},
                limits={"hp": ns.ResourceLimitsConfig(base=12, resources=["hp"])},
            ),
            rewards={
                "gems": ns.inventoryReward("gem", weight=1.0),
                "marked": ns.reward(ns.num("state:marked"), weight=0.125, per_tick=True),
            },
            on_tick=H(name="regen", filters=[ns.PeriodicFilter(period=4)], mutations=[ns.updateTarget({"energy": 1})]),
        )

    healer = ns.GridObjectConfig(  # fixed AOE: heals agents in radius 2 each tick, +1 energy while present
        name="healer",
        aoes={"heal": ns.AOEConfig(radius=2, filters=[ns.isA("agent")], mutations=[ns.updateTarget({"hp": 1})],
                                   presence_deltas={"energy": 1})},  # fmt: skip
    )
    spikes = ns.GridObjectConfig(  # fixed AOE radius 3 that hurts only the other team; two deltas net out per tick
        name="spikes",
        tags=["team:red"],
        aoes={"hurt": ns.AOEConfig(radius=3, filters=[ns.isA("agent"), ns.isNot(ns.sharedTagPrefix("team:"))],
                                   mutations=[ns.updateTarget({"hp": -2}), ns.updateTarget({"hp": 1, "energy": -1})])},  # fmt: skip
    )
    beacon_red = ns.GridObjectConfig(name="beacon_red", tags=["team:red"],
                                     territory_controls=[ns.TerritoryControlConfig(territory="land", strength=8, decay=2)])  # fmt: skip
    beacon_blue = ns.GridObjectConfig(name="beacon_blue", tags=["team:blue"],
                                      territory_controls=[ns.TerritoryControlConfig(territory="land", strength=6, decay=1)])  # fmt: skip
    mine = ns.GridObjectConfig(  # crates of ore that vanish when emptied; using one marks the agent
        name="mine",
        inventory=ns.InventoryConfig(initial={"ore": 3}),
        on_use_handler=H(name="dig", filters=[ns.targetHas({"ore": 1})],
                         mutations=[ns.withdraw({"ore": 1}, remove_when_empty=True), ns.addTag("state:marked", target=A.ACTOR)]),  # fmt: skip
    )
    vault = ns.GridObjectConfig(
        name="vault",
        inventory=ns.InventoryConfig(initial={"gem": 2}),
        on_use_handler=ns.firstMatch([
            H(name="trade", filters=[ns.actorHas({"ore": 2}), ns.isNear(ns.typeTag("healer"), radius=6)],
              mutations=[ns.updateActor({"ore": -2, "gem": 1}), ns.removeTag("state:marked", target=A.ACTOR)]),
            H(name="peek", filters=[ns.actorHasTag("state:marked")], mutations=[ns.logStat("vault.peeks")]),
        ]),  # fmt: skip
    )
    objects = {"wall": ns.WallConfig(), "healer": healer, "spikes": spikes, "beacon_red": beacon_red,
               "beacon_blue": beacon_blue, "mine": mine, "vault": vault}  # fmt: skip
    events = {
        "storm": ns.EventConfig(  # hits 2 random agents every 7 ticks
            name="storm", target_query=ns.query(ns.typeTag("agent")), timesteps=ns.periodic(3, 7, 400), max_targets=2,
            mutations=[ns.updateTarget({"hp": -1}), ns.logStat("storm.hits")],
        ),
        "amnesty": ns.EventConfig(  # clears the marked tag from everyone marked, else (fallback) refills the vaults
            name="amnesty", target_query=ns.query("state:marked"), timesteps=ns.periodic(10, 15, 400), fallback="refill",
            mutations=[ns.removeTag("state:marked")],
        ),
        "refill": ns.EventConfig(
            name="refill", target_query=ns.query(ns.typeTag("vault"), [ns.isNot(ns.targetHas({"gem": 3}))]), timesteps=[],
            mutations=[ns.updateTarget({"gem": 1})],
        ),
    }  # fmt: skip
    if spawn:
        objects["sprout"] = ns.GridObjectConfig(name="sprout", inventory=ns.InventoryConfig(initial={"ore": 1}),
                                                on_use_handler=H(name="pick", mutations=[ns.withdraw({"ore": 1}, remove_when_empty=True)]))  # fmt: skip
        events["growth"] = ns.EventConfig(  # a sprout grows on a healer's north side now and then
            name="growth", target_query=ns.query(ns.typeTag("healer")), timesteps=ns.periodic(5, 11, 400), max_targets=1,
            mutations=[ns.RaycastSpawnMutation(object_type="sprout", directions=["north", "east"], max_range=2, blocker=[ns.isA("wall")])],
        )  # fmt: skip
    agents = [team_agent(0, "team:red", i == 0) for i in range(agents_per_team)] + \
             [team_agent(1, "team:blue", False) for _ in range(agents_per_team)]  # fmt: skip
    on_marked_removed = H(name="unmark", mutations=[ns.updateTarget({"energy": 2})])
    for a in agents:
        a.on_tag_remove = {"state:": on_marked_removed}
    game = ns.GameConfig(
        resource_names=list(WORLD_RESOURCES),
        num_agents=2 * agents_per_team,
        max_steps=max_steps,
        obs=ns.ObsConfig(width=11, height=11, num_tokens=num_tokens, aoe_mask=aoe_mask,
                         global_obs=ns.GlobalObsConfig(obs={"marked": ns.num("state:marked")})),
        agents=agents,
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(allowed_directions=list(EIGHT_WAY)),
                                 change_vibe=ns.ChangeVibeActionConfig(vibes=vibes)),
        objects=objects,
        tags=["state:marked"],
        territories={"land": ns.TerritoryConfig(
            tag_prefix="team:",
            on_enter={"enter": H(filters=[ns.sharedTagPrefix("team:")], mutations=[ns.updateTarget({"shield": 2})])},
            on_exit={"exit": H(filters=[ns.sharedTagPrefix("team:")], mutations=[ns.updateTarget({"shield": -2})])},
            presence={"tax": H(filters=[ns.isNot(ns.sharedTagPrefix("team:")), ns.PeriodicFilter(period=3)],
                               mutations=[ns.updateTarget({"energy": -1})])},
        )},
        events=events,
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def world_map(agents_per_team: int = 3, width: int = 18, height: int = 14, seed: int = 0) -> np.ndarray:
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    return random_map(RandomMapConfig(width=width, height=height, border_width=1, seed=seed,
                                      agents={"red": agents_per_team, "blue": agents_per_team},
                                      objects={"wall": 10, "healer": 2, "spikes": 2, "beacon_red": 2, "beacon_blue": 2,
                                               "mine": 3, "vault": 2}))  # fmt: skip


# --------------------------------------------------------------------------------------------------
# "network" game: materialized + closure + raycast queries, query-inventory transfers with stats,
# push, clear-inventory, game-value filters, ratio / min / max values, game-level on_tick, on_after_use.
# --------------------------------------------------------------------------------------------------
NETWORK_RESOURCES = ["power", "scrap", "gear", "heart"]


def network_config(ns=None, num_agents: int = 4, num_tokens: int = 220, max_steps: int = 0):
    if ns is None:
        ns = C
    H, A = ns.Handler, ns.EntityTarget
    vibes = [ns.Vibe("", n) for n in ["default", "swords"]]
    powered = ns.closureQuery(ns.typeTag("hub"), ns.query(ns.typeTag("relay")), [ns.maxDistance(4)])
    hub = ns.GridObjectConfig(
        name="hub",
        inventory=ns.InventoryConfig(initial={"power": 40}),
        on_use_handler=H(name="charge", filters=[ns.GameValueFilter(target=ns.HandlerTarget.ACTOR, value=ns.inv("power"), min=0),
                                                 ns.isNot(ns.actorHas({"power": 6}))],
                         mutations=[ns.withdraw({"power": 3}), ns.recomputeMaterializedQuery("net:")]),  # fmt: skip
    )
    relay = ns.GridObjectConfig(
        name="relay",
        inventory=ns.InventoryConfig(initial={"scrap": 2}),
        on_use_handler=ns.firstMatch([
            # powered relays hand out gear paid for from every powered node's scrap
            H(name="craft", filters=[ns.hasTag("net:powered"), ns.actorHas({"power": 1})],
              mutations=[ns.updateActor({"power": -1, "gear": 1}),
                         ns.queryWithdraw(ns.query("net:powered"), {"scrap": 1}, stat_prefix="net.")]),
            H(name="shove", mutations=[ns.PushObjectMutation(), ns.recomputeMaterializedQuery("net:")]),
        ]),  # fmt: skip
    )
    crate = ns.GridObjectConfig(
        name="crate",
        inventory=ns.InventoryConfig(initial={"heart": 1}),
        on_use_handler=H(name="push", mutations=[ns.PushObjectMutation()]),
    )
    agent = ns.AgentConfig(
        inventory=ns.InventoryConfig(
            default_limit=20,
            initial={"power": 2},
            limits={"tools": ns.ResourceLimitsConfig(base=3, resources=["gear", "scrap"])},
        ),
        rewards={
            "eff": ns.AgentReward(reward=ns.RatioGameValue(numerator=ns.inv("gear"), denominator=ns.inv("power"))),
            "net": ns.reward(ns.QueryInventoryValue(query=ns.query("net:powered"), item="scrap"), weight=0.05, max=1.0, min=0.1),
            "seen": ns.reward(ns.QueryCountValue(query=ns.raycastQuery(ns.query(ns.typeTag("hub")), max_range=3,
                                                                        blocker=[ns.isA("wall")], include_blocker=False)), weight=0.01, per_tick=True),
        },
        on_after_use_handler=H(name="tired", filters=[ns.actorHas({"gear": 3})],
                               mutations=[ns.ClearInventoryMutation(target=A.ACTOR, limit_name="tools"), ns.logStat("resets")]),
    )  # fmt: skip
    game = ns.GameConfig(
        resource_names=list(NETWORK_RESOURCES),
        num_agents=num_agents,
        max_steps=max_steps,
        obs=ns.ObsConfig(width=9, height=9, num_tokens=num_tokens),
        agent=agent,
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(), change_vibe=ns.ChangeVibeActionConfig(vibes=vibes)),
        objects={"wall": ns.WallConfig(), "hub": hub, "relay": relay, "crate": crate},
        tags=["net:powered"],
        materialize_queries=[ns.materializedQuery("net:powered", powered)],
        on_tick=H(name="leak", filters=[ns.PeriodicFilter(period=6)],
                  mutations=[ns.queryDelta(ns.query("net:powered", [ns.targetHas({"scrap": 1})]), {"scrap": -1}),
                             ns.logStat("leaks")]),
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def network_map(num_agents: int = 4, width: int = 16, height: int = 12, seed: int = 0) -> np.ndarray:
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    return random_map(RandomMapConfig(width=width, height=height, border_width=1, seed=seed, agents=num_agents,
                                      objects={"wall": 6, "hub": 2, "relay": 7, "crate": 4}))  # fmt: skip


# --------------------------------------------------------------------------------------------------
# The game of the reference's deterministic episode signature (scripts/deterministic_episode_signature.py:53-92):
# one agent in a walled 7 x 6 room with a hub and three wires, a materialized closure query, a per-tick log reward.
# --------------------------------------------------------------------------------------------------
def signature_config(ns=None):
    if ns is None:
        ns = C
    cfg = ns.MettaGridConfig.EmptyRoom(num_agents=1, with_walls=True).with_ascii_map(
        [
            ["#", "#", "#", "#", "#", "#", "#"],
            ["#", ".", ".", ".", ".", ".", "#"],
            ["#", ".", "W", "H", "W", ".", "#"],
            ["#", ".", ".", "W", ".", ".", "#"],
            ["#", ".", ".", "@", ".", ".", "#"],
            ["#", "#", "#", "#", "#", "#", "#"],
        ],
        char_to_map_name={"#": "wall", "@": "agent.agent", ".": "empty", "H": "hub", "W": "wire"},
    )
    cfg.game.actions.noop.enabled = True
    cfg.game.resource_names = ["gold", "silver"]
    cfg.game.agent.inventory.initial = {"gold": 10, "silver": 5}
    cfg.game.agent.rewards = {
        "stability": ns.reward(
            ns.weighted_sum([(1.0, ns.InventoryValue(item="gold")), (0.5, ns.InventoryValue(item="silver"))], log=True),
            per_tick=True,
        )
    }
    cfg.game.objects["hub"] = ns.GridObjectConfig(name="hub", map_name="hub", tags=[ns.typeTag("hub")])
    cfg.game.objects["wire"] = ns.GridObjectConfig(name="wire", map_name="wire", tags=[ns.typeTag("wire")])
    cfg.game.materialize_queries = [
        ns.MaterializedQuery(
            tag="connected_one",
            query=ns.ClosureQuery(source=ns.typeTag("hub"), candidates=ns.query(ns.typeTag("wire")),
                                  edge_filters=[ns.maxDistance(1)], max_items=1),  # fmt: skip
        )
    ]
    return cfg


# --------------------------------------------------------------------------------------------------
# The reference's own perf-benchmark "toy" preset (benchmarks/perf/perf_benchmark.py:35-77): 40 x 40
# RandomMapBuilder map, 4 % walls + a wall border, 20 agents, noop + 8-way move, 11 x 11 window, 200 tokens.
# --------------------------------------------------------------------------------------------------
def toy_config(num_agents: int = 20, map_size: int = 40, density: float = 0.04, seed: int = 42, max_steps: int = 0):
    return C.MettaGridConfig(
        game=C.GameConfig(
            num_agents=num_agents,
            max_steps=max_steps,
            obs=C.ObsConfig(width=11, height=11, num_tokens=200),
            actions=C.ActionsConfig(noop=C.NoopActionConfig(), move=C.MoveActionConfig(allowed_directions=list(EIGHT_WAY))),
            objects={"wall": C.WallConfig()},
            map_builder=RandomMapConfig(width=map_size, height=map_size, agents=num_agents, border_width=1, seed=seed,
                                        objects={"wall": int(map_size * map_size * density)}),  # fmt: skip
        )
    )


# --------------------------------------------------------------------------------------------------
# BASELINE.json workloads: what bench.py times and what tests/test_gpu_full_size.py checks.
# --------------------------------------------------------------------------------------------------
C3_OBJECTS = {"wall": 10, "chest": 6, "altar": 4}
C4_OBJECTS = {"wall": 120, "healer": 6, "spikes": 6, "beacon_red": 5, "beacon_blue": 5, "mine": 12, "vault": 6}

WORKLOADS = {
    # name: (description, default envs per GPU, agents)
    "c2": ("C2: reference benchmark game 20x20, {A} agents/env, 13x13 obs, 100 tokens", 4096, 16),
    "c3": ("C3: combat-heavy game on the arena layout 37x37 (MapGen 25x25 + border 6), {A} agents in 2 teams, handler "
           "chains, 11x11 obs, 500 tokens", 16384, 24),
    "c4": ("C4: world game on MapGen 66x66 (64x64 + border), {A} agents, fixed+mobile AOE, territory + aoe_mask, events, "
           "queries, 11x11 obs, 200 tokens", 8192, 24),
    "toy": ("toy: the reference perf_benchmark preset, 40x40 + wall border, 4 % walls, {A} agents, 8-way move, 11x11 obs, "
            "200 tokens", 4096, 20),
}  # fmt: skip


def make_cfg(agents: int, workload: str = "c2"):
    if workload == "c3":
        return combat_config(None, agents // 2, num_tokens=500)
    if workload == "c4":
        return world_config(None, agents // 2, num_tokens=200)
    if workload == "toy":
        return toy_config(agents)
    return benchmark_config(agents)


_POOLS: dict = {}


def map_pool(workload: str) -> list[np.ndarray]:
    """The 64 recorded map instances of C3 / C4 (built by the reference's MapGen, tools/make_workload_maps.py)."""
    if workload not in _POOLS:
        from pathlib import Path

        z = np.load(Path(__file__).resolve().parent / "workload_maps" / f"{workload}.npz")
        names = z["names"]
        _POOLS[workload] = [names[g] for g in z["grids"]]
    return _POOLS[workload]


def make_map(cfg, agents: int, workload: str, index: int) -> np.ndarray:
    """Map of env `index`: C2 / toy = RandomMapBuilder with seed 42 + index; C3 / C4 = the recorded MapGen pool, cycled
    (24 agents only: other agent counts fall back to the repo's RandomMapBuilder restatement)."""
    from mettagrid_b200.mapgen import random_map

    if workload in ("c3", "c4"):
        if agents == 24:
            pool = map_pool(workload)
            return pool[index % len(pool)]
        if workload == "c3":
            return random_map(RandomMapConfig(width=37, height=37, border_width=6, seed=42 + index,
                                              agents={"red": agents // 2, "blue": agents // 2}, objects=dict(C3_OBJECTS)))  # fmt: skip
        return random_map(RandomMapConfig(width=66, height=66, border_width=1, seed=42 + index,
                                          agents={"red": agents // 2, "blue": agents // 2}, objects=dict(C4_OBJECTS)))  # fmt: skip
    return random_map(cfg.game.map_builder, seed=42 + index)


def algo_bytes(P, workload: str) -> float:
    """B_io + B_state per agent-step (550 for C1/C2, as the survey computes it)."""
    if workload == "c2":
        return 550.0
    T, A, R = P.num_tokens, P.num_agents, len(P.resource_names)
    b_io = 3 * T + 4 + 4 + 4 + 1 + 1 + 8
    b_state = (2 * P.height * P.width + 16) / A + 2 * (4 + 4 + 4 + 1 + 2 * R + 4 * R) + 4 * 8
    return float(b_io + b_state)


def gen_actions(num_actions: int, num_primary: int, steps: int, envs: int, agents: int, seed: int, mode: str = "effective"):
    """SURVEY 8(d), both from np.random.RandomState(...).randint like the reference's benchmark
    (benchmarks/test_mettagrid_env_benchmark.py:44-49):
    'effective': primary uniform over the non-vibe actions, vibe change with p = 0.1;
    'verbatim': uniform over ALL action ids into the primary buffer, vibe buffer left 0 -- about 97 % of the ids of the
    benchmark game are change_vibe ids, which the primary stream ignores."""
    rng = np.random.RandomState(seed)
    if mode == "verbatim":
        prim = rng.randint(0, num_actions, size=(steps, envs, agents)).astype(np.int32)
        return prim, np.zeros_like(prim)
    prim = rng.randint(0, num_primary, size=(steps, envs, agents)).astype(np.int32)
    vibe = np.zeros_like(prim)
    if num_actions > num_primary:
        m = rng.rand(steps, envs, agents) < 0.1
        vibe[m] = rng.randint(num_primary, num_actions, size=int(m.sum()))
    return prim, vibe
