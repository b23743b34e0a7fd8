"""Small games that pin the corner cases the reference's own tests encode, written against a class
namespace like mettagrid_b200/workloads.py so that tests/golden/make_golden.py can record them from the real
reference:

* beam_chain      -- tests/test_move_handlers.py:60-172: explicit handler chain, a range-5 beam through empty cells that
                     stops at walls, swap with immobile agents, a handler whose second mutation fails after its first
                     was applied (no rollback), spawn-on-move.
* aoe_round       -- tests/test_aoe_round_radius.py: Euclidean AOE radii (dr^2 + dc^2 <= r^2), agents walking across
                     the rim of radius-2 / radius-3 / radius-4 sources, presence deltas entering and leaving.
* dyn_limits      -- tests/test_dynamic_inventory_limits.py: limits with modifiers, min floor / max cap, one modifier
                     item feeding two limits so that losing it violates both at once.
* event_targets   -- tests/test_event_max_targets.py: max_targets 1 / 3 / None, filters on targets, an event that never applies -> fallback.
* many_tagged     -- a tag with more than 64 members queried by an event (max_targets -> RNG order) and by isNear:
                     the engine's ordered tag lists overflow and fall back to scanning (MG_TAG_LIST_CAP).
"""

from __future__ import annotations

import numpy as np

from mettagrid_b200.mapgen import RandomMapConfig, random_map
from mettagrid_b200.workloads import EIGHT_WAY


def _vibes(ns, names):
    return [ns.Vibe("", n) for n in names]


# ---------------------------------------------------------------------------------------------------------------
def beam_chain_config(ns, agents_per_team: int = 2):
    H, T = ns.Handler, ns.HandlerTarget
    handlers = [
        H(name="zap_beam", filters=[ns.VibeFilter(target=T.ACTOR, vibe="swords"), ns.MaxDistanceFilter(target=T.TARGET, radius=5),
                                    ns.isA("agent")],
          mutations=[ns.ResourceDeltaMutation(target="target", deltas={"mobility": -1})]),
        # first mutation lands, the second fails on a wall (no on_use): the handler reports failure, nothing is rolled back
        H(name="probe", filters=[ns.isA("wall"), ns.actorHas({"energy": 1})],
          mutations=[ns.updateActor({"energy": -1}), ns.UseTargetMutation()]),
        H(name="build", filters=[ns.TargetLocEmptyFilter(), ns.VibeFilter(target=T.ACTOR, vibe="shield"), ns.actorHas({"energy": 3})],
          mutations=[ns.updateActor({"energy": -3}), ns.SpawnObjectMutation(object_type="block")]),
        H(name="move", filters=[ns.TargetLocEmptyFilter()], mutations=[ns.RelocateMutation()]),
        H(name="swap_immobile", filters=[ns.isA("agent"), ns.isNot(ns.ResourceFilter(target=T.TARGET, resources={"mobility": 1}))],
          mutations=[ns.SwapMutation()]),
        H(name="on_use", filters=[ns.TargetIsUsableFilter()], mutations=[ns.UseTargetMutation()]),
    ]  # fmt: skip

    def agent(team, tag):
        return ns.AgentConfig(
            team_id=team, tags=[tag],
            inventory=ns.InventoryConfig(
                limits={"mobility": ns.ResourceLimitsConfig(base=2, max=2, resources=["mobility"]),
                        "energy": ns.ResourceLimitsConfig(base=9, resources=["energy"])},
                initial={"mobility": 2, "energy": 6}),
            on_tick=ns.Handler(name="recover", filters=[ns.PeriodicFilter(period=9)], mutations=[ns.updateTarget({"mobility": 1, "energy": 1})]),
        )  # fmt: skip

    block = ns.GridObjectConfig(name="block", inventory=ns.InventoryConfig(initial={"energy": 2}),
                                on_use_handler=H(name="salvage", mutations=[ns.withdraw({"energy": 1}, remove_when_empty=True)]))  # fmt: skip
    game = ns.GameConfig(
        resource_names=["mobility", "energy"],
        num_agents=2 * agents_per_team,
        max_steps=0,
        obs=ns.ObsConfig(width=9, height=9, num_tokens=140),
        agents=[agent(0, "team:red") for _ in range(agents_per_team)] + [agent(1, "team:blue") for _ in range(agents_per_team)],
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(handlers=handlers),
                                 change_vibe=ns.ChangeVibeActionConfig(vibes=_vibes(ns, ["default", "swords", "shield"]))),
        objects={"wall": ns.WallConfig(), "block": block},
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def beam_chain_map(seed: int = 0, agents_per_team: int = 2):
    return random_map(RandomMapConfig(width=12, height=9, border_width=1, seed=seed,
                                      agents={"red": agents_per_team, "blue": agents_per_team}, objects={"wall": 9}))  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------
def aoe_round_config(ns, num_agents: int = 5):
    def source(name, radius, res, presence):
        return ns.GridObjectConfig(
            name=name, aoes={"aoe": ns.AOEConfig(radius=radius, mutations=[ns.updateTarget({res: 10})],
                                                 presence_deltas={presence: 1})})  # fmt: skip

    game = ns.GameConfig(
        resource_names=["energy", "heat", "light", "near2", "near3", "near4"],
        num_agents=num_agents,
        max_steps=0,
        obs=ns.ObsConfig(width=7, height=7, num_tokens=120),
        agent=ns.AgentConfig(inventory=ns.InventoryConfig(
            default_limit=5,
            limits={"energy": ns.ResourceLimitsConfig(base=1000, resources=["energy"]),
                    "heat": ns.ResourceLimitsConfig(base=255, resources=["heat"]),
                    "light": ns.ResourceLimitsConfig(base=40000, resources=["light"])})),
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(allowed_directions=list(EIGHT_WAY))),
        objects={"wall": ns.WallConfig(), "src2": source("src2", 2, "energy", "near2"), "src3": source("src3", 3, "heat", "near3"),
                 "src4": source("src4", 4, "light", "near4")},
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def aoe_round_map(seed: int = 0, num_agents: int = 5):
    return random_map(RandomMapConfig(width=15, height=13, border_width=1, seed=seed, agents=num_agents,
                                      objects={"src2": 2, "src3": 1, "src4": 1, "wall": 3}))  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------
def dyn_limits_config(ns, num_agents: int = 3):
    H = ns.Handler
    agent = ns.AgentConfig(
        inventory=ns.InventoryConfig(
            default_limit=7,
            limits={
                "gear": ns.ResourceLimitsConfig(base=4, resources=["gear"]),
                "battery": ns.ResourceLimitsConfig(base=2, max=9, resources=["battery"], modifiers={"gear": 3}),
                # the same modifier item also widens the shared ammo / flare limit: dropping one gear can break both
                "ordnance": ns.ResourceLimitsConfig(base=1, max=8, resources=["ammo", "flare"], modifiers={"gear": 2, "battery": 1}),
                "gold": ns.ResourceLimitsConfig(base=100, max=50, resources=["gold"]),  # max below min: max wins
            },
            initial={"gear": 2, "battery": 5, "ammo": 2, "flare": 1, "gold": 3},
        ),
    )
    depot = ns.GridObjectConfig(
        name="depot",
        inventory=ns.InventoryConfig(initial={"gear": 20, "battery": 60, "ammo": 60, "flare": 60, "gold": 200},
                                     limits={"all": ns.ResourceLimitsConfig(base=60000, resources=["gear", "battery", "ammo", "flare", "gold"])}),
        on_use_handler=ns.firstMatch([
            H(name="stock_gear", filters=[ns.actorVibe("swords")], mutations=[ns.withdraw({"gear": 1})]),
            H(name="stock_all", mutations=[ns.withdraw({"battery": 4, "ammo": 3, "flare": 3, "gold": 30})]),
        ]),
    )  # fmt: skip
    shredder = ns.GridObjectConfig(
        name="shredder",
        on_use_handler=H(name="shred", filters=[ns.actorHas({"gear": 1})], mutations=[ns.updateActor({"gear": -1})]),
    )
    game = ns.GameConfig(
        resource_names=["gear", "battery", "ammo", "flare", "gold"],
        num_agents=num_agents,
        max_steps=0,
        obs=ns.ObsConfig(width=7, height=7, num_tokens=160),
        agent=agent,
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(),
                                 change_vibe=ns.ChangeVibeActionConfig(vibes=_vibes(ns, ["default", "swords"]))),
        objects={"wall": ns.WallConfig(), "depot": depot, "shredder": shredder},
    )
    return ns.MettaGridConfig(game=game)


def dyn_limits_map(seed: int = 0, num_agents: int = 3):
    return random_map(RandomMapConfig(width=9, height=8, border_width=1, seed=seed, agents=num_agents,
                                      objects={"depot": 5, "shredder": 5, "wall": 2}))  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------
def event_targets_config(ns, num_agents: int = 4):
    def ev(name, q, start, period, max_targets, filters, mutations, fallback=None):
        return ns.EventConfig(name=name, target_query=q, timesteps=ns.periodic(start, period, 300), max_targets=max_targets,
                              filters=filters, mutations=mutations, fallback=fallback)  # fmt: skip

    crate = ns.GridObjectConfig(name="crate", tags=["loot:crate"], inventory=ns.InventoryConfig(initial={"coin": 1}))
    events = {
        "one": ev("one", ns.query(ns.typeTag("crate")), 2, 5, 1, [], [ns.updateTarget({"coin": 1}), ns.logStat("ev.one")]),
        "three": ev("three", ns.query("loot:crate"), 3, 7, 3, [ns.isNot(ns.targetHas({"coin": 4}))], [ns.updateTarget({"coin": 2}), ns.logStat("ev.three")]),
        "all": ev("all", ns.query(ns.typeTag("agent")), 4, 6, None, [ns.targetHas({"coin": 1})], [ns.updateTarget({"coin": -1}), ns.logStat("ev.all")]),
        # no target ever passes the filter: nothing is applied and the fallback event runs instead
        "none": ev("none", ns.query(ns.typeTag("agent")), 5, 11, 1, [ns.targetHas({"coin": 1000})], [ns.updateTarget({"coin": 50})], fallback="rescue"),
        "rescue": ns.EventConfig(name="rescue", target_query=ns.query(ns.typeTag("agent"), [ns.isNot(ns.targetHas({"coin": 2}))]), timesteps=[],
                                 max_targets=2, mutations=[ns.updateTarget({"coin": 3}), ns.logStat("ev.rescue")]),
    }  # fmt: skip
    game = ns.GameConfig(
        resource_names=["coin"],
        num_agents=num_agents,
        max_steps=0,
        obs=ns.ObsConfig(width=7, height=7, num_tokens=120),
        agent=ns.AgentConfig(inventory=ns.InventoryConfig(default_limit=30, initial={"coin": 2}),
                             rewards={"coin": ns.inventoryReward("coin", weight=0.25)}),
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig()),
        objects={"wall": ns.WallConfig(), "crate": crate},
        tags=["loot:crate"],
        events=events,
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def event_targets_map(seed: int = 0, num_agents: int = 4):
    return random_map(RandomMapConfig(width=11, height=9, border_width=1, seed=seed, agents=num_agents, objects={"crate": 7, "wall": 3}))


# ---------------------------------------------------------------------------------------------------------------
def many_tagged_config(ns, num_agents: int = 4):
    H = ns.Handler
    pillar = ns.GridObjectConfig(name="pillar", tags=["deco:pillar"], inventory=ns.InventoryConfig(initial={"moss": 1}))
    shrine = ns.GridObjectConfig(
        name="shrine",
        on_use_handler=ns.firstMatch([
            H(name="bless", filters=[ns.isNear("deco:pillar", radius=2)], mutations=[ns.updateActor({"moss": 1}), ns.logStat("blessed")]),
            H(name="curse", mutations=[ns.updateActor({"moss": -1})]),
        ]),
    )  # fmt: skip
    events = {
        "grow": ns.EventConfig(name="grow", target_query=ns.query("deco:pillar"), timesteps=ns.periodic(2, 4, 200), max_targets=5,
                               mutations=[ns.updateTarget({"moss": 1})]),
        "prune": ns.EventConfig(name="prune", target_query=ns.query(ns.typeTag("pillar"), [ns.targetHas({"moss": 3})]),
                                timesteps=ns.periodic(9, 9, 200), max_targets=None, mutations=[ns.updateTarget({"moss": -2}), ns.logStat("pruned")]),
    }  # fmt: skip
    game = ns.GameConfig(
        resource_names=["moss"],
        num_agents=num_agents,
        max_steps=0,
        obs=ns.ObsConfig(width=9, height=9, num_tokens=200),
        agent=ns.AgentConfig(inventory=ns.InventoryConfig(default_limit=20, initial={"moss": 2}),
                             rewards={"pillars": ns.reward(ns.num("deco:pillar"), weight=0.01, per_tick=True)}),
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig()),
        objects={"wall": ns.WallConfig(), "pillar": pillar, "shrine": shrine},
        tags=["deco:pillar"],
        events=events,
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def many_tagged_map(seed: int = 0, num_agents: int = 4):
    return random_map(RandomMapConfig(width=20, height=16, border_width=1, seed=seed, agents=num_agents,
                                      objects={"pillar": 90, "shrine": 6, "wall": 4}))  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------
def attack_mutation_config(ns, agents_per_team: int = 3):
    """The C++ AttackMutation (attack_mutation.hpp:20-38: damage = max(0, weapon * pct / 100 - armor) off the target's
    health).  No Python config of the reference lowers to it, so `ns` must be our mirror classes; the
    reference side is built through pybind's add_attack_mutation (oracle/ref_driver.py), which is how its golden
    fixture is recorded."""
    H = ns.Handler
    strike = H(name="strike", filters=[ns.isA("agent"), ns.actorHas({"weapon": 1}), ns.isNot(ns.sharedTagPrefix("team:"))],
               mutations=[ns.CppAttackMutation(weapon="weapon", armor="armor", health="hp", damage_multiplier_pct=150),
                          ns.updateActor({"energy": -1}), ns.logStat("strikes")])  # fmt: skip
    heavy = H(name="heavy", filters=[ns.isA("agent"), ns.actorVibe("swords"), ns.actorHas({"weapon": 3})],
              mutations=[ns.CppAttackMutation(weapon="weapon", armor="armor", health="hp", damage_multiplier_pct=333)])  # fmt: skip

    def agent(team, tag, weapon, armor):
        return ns.AgentConfig(team_id=team, tags=[tag], inventory=ns.InventoryConfig(
            default_limit=40, initial={"hp": 30, "weapon": weapon, "armor": armor, "energy": 9}),
            on_tick=H(name="mend", filters=[ns.PeriodicFilter(period=6)], mutations=[ns.updateTarget({"hp": 2, "energy": 1})]))  # fmt: skip

    armory = ns.GridObjectConfig(name="armory", inventory=ns.InventoryConfig(initial={"weapon": 9, "armor": 9}),
                                 on_use_handler=ns.firstMatch([
                                     H(name="arm", filters=[ns.actorVibe("swords")], mutations=[ns.withdraw({"weapon": 1})]),
                                     H(name="plate", mutations=[ns.withdraw({"armor": 1})])]))  # fmt: skip
    agents = [agent(0, "team:red", 1 + i, i % 3) for i in range(agents_per_team)] + \
             [agent(1, "team:blue", 3 - i % 3, 2 - i % 2) for i in range(agents_per_team)]  # fmt: skip
    game = ns.GameConfig(
        resource_names=["hp", "weapon", "armor", "energy"],
        num_agents=2 * agents_per_team,
        max_steps=0,
        obs=ns.ObsConfig(width=9, height=9, num_tokens=160),
        agents=agents,
        actions=ns.ActionsConfig(noop=ns.NoopActionConfig(), move=ns.MoveActionConfig(handlers=[heavy, strike]),
                                 change_vibe=ns.ChangeVibeActionConfig(vibes=_vibes(ns, ["default", "swords"]))),
        objects={"wall": ns.WallConfig(), "armory": armory},
    )  # fmt: skip
    return ns.MettaGridConfig(game=game)


def attack_mutation_map(seed: int = 0, agents_per_team: int = 3):
    return random_map(RandomMapConfig(width=9, height=8, border_width=1, seed=seed, agents={"red": agents_per_team, "blue": agents_per_team},
                                      objects={"armory": 3, "wall": 2}))  # fmt: skip


def wrong_stream_actions(rng: np.random.RandomState, steps: int, agents: int, num_primary: int, num_actions: int):
    """The reference benchmark's verbatim sampling (uniform over ALL ids into the primary buffer) plus valid
    ids on the wrong stream in the vibe buffer (cpp/bindings/mettagrid_c.cpp:975-978: silently ignored, not failures)."""
    prim = rng.randint(0, num_actions, size=(steps, agents)).astype(np.int32)
    vibe = np.zeros_like(prim)
    u = rng.rand(steps, agents)
    m = (u >= 0.5) & (u < 0.75)
    vibe[m] = rng.randint(num_primary, num_actions, size=int(m.sum()))  # real vibe changes
    m = u >= 0.75
    vibe[m] = rng.randint(0, num_primary, size=int(m.sum()))  # moves / noop in the vibe buffer: ignored
    return prim, vibe
