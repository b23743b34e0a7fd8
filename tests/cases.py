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
