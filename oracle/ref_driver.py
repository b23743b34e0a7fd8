"""Drive the UNMODIFIED reference step (oracle/_ref: the reference's own C++ compiled by
oracle/Makefile.ref) from our config objects -- TEST INFRASTRUCTURE ONLY.

The reference's Python lowering (config/mettagrid_c_config.py) cannot travel to the GPU box, so this
module feeds the reference's pybind config classes directly, using the id maps and the lowering
decisions our compiler already made (``mettagrid_b200.compiler._Builder``): both sides therefore see
the same resource / tag / vibe / type / feature ids.  In the build container
tests/test_oracle_vs_reference.py checks this driver against the reference's own
``convert_to_cpp_game_config`` on the configs it supports.

Used by bench.py (``--impl reference`` and the ``cpu_baseline`` leg) and by tests.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from mettagrid_b200 import compiler as mc
from oracle import reference


class RefUnsupported(NotImplementedError):
    pass


def _mod():
    m = None
    if reference.available():
        pkg = reference.load()
        if pkg is not None:
            import mettagrid.mettagrid_c as m  # noqa: PLC0415
    if m is None:
        m = reference.load_module_only()
    if m is None:
        raise RuntimeError("oracle/_ref is not built: run `make -f oracle/Makefile.ref -j8` in the build container")
    return m


class _Lower:
    """Our config -> reference pybind config objects (mirrors mettagrid_c_config.py:576-1007)."""

    def __init__(self, game: Any):
        self.m = _mod()
        self.g = game
        b = mc._Builder(game)
        b._spawn_refs = []
        b.build_id_maps()
        b.build_feature_ids()
        self.b = b

    def ent(self, x):
        return getattr(self.m.EntityRef, mc._ev(x))

    # -- values ------------------------------------------------------------------------------------
    def value(self, gv):
        m, b = self.m, self.b
        if isinstance(gv, (int, float)):
            c = m.ConstValueConfig()
            c.value = float(gv)
            return c
        kind = type(gv).__name__
        scope = m.GameValueScope.GAME if mc._ev(getattr(gv, "scope", "agent")) == "game" else m.GameValueScope.AGENT
        if kind == "InventoryValue":
            c = m.InventoryValueConfig()
            c.scope, c.id = scope, b.rid[gv.item]
            return c
        if kind == "StatValue":
            c = m.StatValueConfig()
            c.scope, c.stat_name, c.delta = scope, gv.name, bool(gv.delta)
            return c
        if kind == "ConstValue":
            c = m.ConstValueConfig()
            c.value = float(gv.value)
            return c
        if kind == "SumGameValue":
            c = m.SumValueConfig()
            for v in gv.values:
                c.add_value(self.value(v))
            if gv.weights is not None:
                c.weights = list(gv.weights)
            c.log = bool(gv.log)
            return c
        if kind == "RatioGameValue":
            c = m.RatioValueConfig()
            c.numerator, c.denominator = self.value(gv.numerator), self.value(gv.denominator)
            return c
        if kind in ("MaxGameValue", "MinGameValue"):
            c = m.MaxValueConfig() if kind == "MaxGameValue" else m.MinValueConfig()
            for v in gv.values:
                c.add_value(self.value(v))
            return c
        if kind == "QueryInventoryValue":
            c = m.QueryInventoryValueConfig()
            c.id = b.rid[gv.item]
            c.set_query(self.query(gv.query))
            return c
        if kind == "QueryCountValue":
            c = m.QueryCountValueConfig()
            c.set_query(self.query(gv.query))
            return c
        raise RefUnsupported(f"game value {kind}")

    # -- queries (mettagrid_c_config.py:83-186) ------------------------------------------------------
    def _max_items(self, q, cq):
        mi = getattr(q, "max_items", None)
        if mi is not None:
            cq.max_items = self.value(float(mi) if isinstance(mi, int) else mi)
        if getattr(q, "order_by", None) == "random":
            cq.order_by = self.m.QueryOrderBy.random

    def query(self, q):
        m, b = self.m, self.b
        if isinstance(q, str):
            from mettagrid_b200.config import Query  # noqa: PLC0415

            return self.query(Query(source=q))
        qt = getattr(q, "query_type", "query")
        if qt == "materialized":
            from mettagrid_b200.config import Query  # noqa: PLC0415

            return self.query(Query(source=q.tag))
        if qt == "closure":
            cq = m.ClosureQueryConfig()
            cq.set_source(self.query(q.source))
            cq.set_candidates(self.query(q.candidates))
            self._max_items(q, cq)
            self.add_filters(cq, q.edge_filters, prefix="edge_")
            self.add_filters(cq, q.filters, prefix="result_")
            return m.make_query_config(cq)
        if qt == "raycast":
            dirs = {"north": (-1, 0), "south": (1, 0), "east": (0, 1), "west": (0, -1)}
            cq = m.RaycastQueryConfig()
            cq.max_range = self.value(float(q.max_range) if isinstance(q.max_range, int) else q.max_range)
            cq.include_blocker = bool(q.include_blocker)
            cq.directions = [dirs[d] for d in q.directions]
            self._max_items(q, cq)
            cq.set_source(self.query(q.source))
            self.add_filters(cq, q.blocker, prefix="blocker_")
            return m.make_query_config(cq)
        if not isinstance(q.source, str):
            cq = m.FilteredQueryConfig()
            cq.set_source(self.query(q.source))
            self._max_items(q, cq)
            self.add_filters(cq, q.filters)
            return m.make_query_config(cq)
        cq = m.TagQueryConfig()
        cq.tag_id = b.tid[q.source]
        self._max_items(q, cq)
        self.add_filters(cq, q.filters)
        return m.make_query_config(cq)

    # -- filters -----------------------------------------------------------------------------------
    def add_filters(self, target, filters, prefix=""):
        for f in filters or []:
            for key, obj in self.lower_filter(f):
                getattr(target, f"add_{prefix}{key}_filter")(obj)

    def lower_filter(self, f):
        m, b = self.m, self.b
        ft = getattr(f, "filter_type", None)
        if ft == "not":
            neg = m.NegFilterConfig()
            for key, obj in self.lower_filter(f.inner):
                getattr(neg, f"add_{key}_filter")(obj)
            return [("neg", neg)]
        if ft == "or":
            orf = m.OrFilterConfig()
            for inner in f.inner:
                low = self.lower_filter(inner)
                if getattr(inner, "filter_type", None) == "resource" and len(low) > 1:
                    ineg = m.NegFilterConfig()
                    for _, r in low:
                        ineg.add_resource_filter(r)
                    oneg = m.NegFilterConfig()
                    oneg.add_neg_filter(ineg)
                    orf.add_neg_filter(oneg)
                else:
                    for key, obj in low:
                        getattr(orf, f"add_{key}_filter")(obj)
            return [("or", orf)]
        if ft == "resource":
            return [("resource", m.ResourceFilterConfig(entity=self.ent(f.target), resource_id=b.rid[r], min_amount=int(v)))
                    for r, v in f.resources.items() if r in b.rid]  # fmt: skip
        if ft == "vibe":
            return [("vibe", m.VibeFilterConfig(entity=self.ent(f.target), vibe_id=b.vid[f.vibe]))] if f.vibe in b.vid else []
        if ft == "tag":
            if f.tag not in b.tid:
                return []
            return [("tag_prefix", m.TagPrefixFilterConfig(entity=self.ent(f.target), tag_ids=[b.tid[f.tag]]))]
        if ft == "tag_prefix":
            return [("tag_prefix", m.TagPrefixFilterConfig(entity=self.ent(f.target), tag_ids=b.prefix_tags(f.tag_prefix)))]
        if ft == "shared_tag_prefix":
            return [("shared_tag_prefix", m.SharedTagPrefixFilterConfig(tag_ids=b.prefix_tags(f.tag_prefix)))]
        if ft == "max_distance":
            c = m.MaxDistanceFilterConfig()
            c.entity, c.radius = self.ent(f.target), int(f.radius)
            if f.query is not None:
                c.set_source(self.query(f.query))
            return [("max_distance", c)]
        if ft == "query_resource":
            c = m.QueryResourceFilterConfig()
            c.set_query(self.query(f.query))
            c.requirements = [(b.rid[r], int(v)) for r, v in f.requirements.items()]
            return [("query_resource", c)]
        if ft == "game_value":
            mn = f.min
            return [("game_value", m.GameValueFilterConfig(value=self.value(f.value),
                                                           threshold=self.value(float(mn) if isinstance(mn, int) else mn),
                                                           entity=self.ent(f.target)))]  # fmt: skip
        if ft == "target_loc_empty":
            return [("target_loc_empty", m.TargetLocEmptyFilterConfig())]
        if ft == "target_is_usable":
            return [("target_is_usable", m.TargetIsUsableFilterConfig())]
        if ft == "periodic":
            start = f.start_on if f.start_on is not None else f.period
            return [("periodic", m.PeriodicFilterConfig(period=int(f.period), start_on=int(start)))]
        raise RefUnsupported(f"filter {ft}")

    # -- mutations ---------------------------------------------------------------------------------
    def add_mutations(self, target, mutations):
        m, b = self.m, self.b
        for mu in mutations or []:
            mt = getattr(mu, "mutation_type", None)
            if mt == "resource_delta":
                for r, d in mu.deltas.items():
                    target.add_resource_delta_mutation(
                        m.ResourceDeltaMutationConfig(entity=self.ent(mu.target), resource_id=b.rid[r], delta=int(d))
                    )
            elif mt == "resource_transfer":
                for r, amt in mu.resources.items():
                    target.add_resource_transfer_mutation(m.ResourceTransferMutationConfig(
                        source=self.ent(mu.from_target), destination=self.ent(mu.to_target), resource_id=b.rid[r],
                        amount=int(amt), remove_source_when_empty=bool(mu.remove_source_when_empty)))  # fmt: skip
            elif mt == "clear_inventory":
                target.add_clear_inventory_mutation(m.ClearInventoryMutationConfig(
                    entity=self.ent(mu.target), resource_ids=b.limit_name_res[mu.limit_name]))  # fmt: skip
            elif mt == "stats":
                c = m.StatsMutationConfig(stat_name=mu.stat, target=getattr(m.StatsTarget, mc._ev(mu.target)),
                                          entity=getattr(m.StatsEntity, mc._ev(mu.entity)))  # fmt: skip
                c.source = self.value(mu.source)
                target.add_stats_mutation(c)
            elif mt == "change_vibe":
                target.add_change_vibe_mutation(
                    m.ChangeVibeMutationConfig(entity=self.ent(mu.target), vibe_id=b.vid.get(mu.vibe_name, 0))
                )
            elif mt == "set_game_value":
                src = mu.source if mu.source is not None else float(mu.delta)
                target.add_game_value_mutation(m.GameValueMutationConfig(
                    value=self.value(mu.value), target=self.ent(mu.target), source=self.value(src)))  # fmt: skip
            elif mt == "relocate":
                target.add_relocate_mutation(m.RelocateMutationConfig())
            elif mt == "swap":
                target.add_swap_mutation(m.SwapMutationConfig())
            elif mt == "use_target":
                target.add_use_target_mutation(m.UseTargetMutationConfig())
            elif mt == "push_object":
                target.add_push_object_mutation(m.PushObjectMutationConfig())
            elif mt == "attack":
                continue  # dropped by the reference's Python lowering
            elif mt == "cpp_attack":  # handler_bindings.hpp:335-355,543-546: the only way to reach the C++ AttackMutation
                target.add_attack_mutation(m.AttackMutationConfig(
                    weapon_resource=b.rid[mu.weapon], armor_resource=b.rid[mu.armor], health_resource=b.rid[mu.health],
                    damage_multiplier_pct=int(mu.damage_multiplier_pct)))  # fmt: skip
            elif mt == "add_tag":
                target.add_add_tag_mutation(m.AddTagMutationConfig(entity=self.ent(mu.target), tag_id=b.tid[mu.tag]))
            elif mt == "remove_tag":
                target.add_remove_tag_mutation(m.RemoveTagMutationConfig(entity=self.ent(mu.target), tag_id=b.tid[mu.tag]))
            elif mt == "remove_tags_with_prefix":
                target.add_remove_tags_with_prefix_mutation(
                    m.RemoveTagsWithPrefixMutationConfig(entity=self.ent(mu.target), tag_ids=b.prefix_tags(mu.prefix)))
            elif mt == "recompute_materialized_query":
                for tg in b.prefix_tags(mu.tag_prefix):
                    c = m.RecomputeMaterializedQueryMutationConfig()
                    c.tag_id = tg
                    target.add_recompute_materialized_query_mutation(c)
            elif mt == "query_inventory":
                c = m.QueryInventoryMutationConfig()
                c.set_query(self.query(mu.query))
                c.deltas = [(b.rid[r], int(d)) for r, d in mu.deltas.items()]
                if mu.source is not None:
                    c.source, c.has_source = self.ent(mu.source), True
                if mu.transfer_stats:
                    c.transfer_stat_names = [(b.rid[r], sname) for r, sname in mu.transfer_stats.items()]
                target.add_query_inventory_mutation(c)
            elif mt == "spawn_object":
                c = m.SpawnObjectMutationConfig()
                c.object_type = mu.object_type
                target.add_spawn_object_mutation(c)
            elif mt == "raycast_spawn":
                dirs = {"north": (-1, 0), "south": (1, 0), "east": (0, 1), "west": (0, -1)}
                c = m.RaycastSpawnMutationConfig()
                c.object_type = mu.object_type
                c.max_range = self.value(float(mu.max_range) if isinstance(mu.max_range, int) else mu.max_range)
                c.directions = [dirs[d] for d in mu.directions]
                self.add_filters(c, mu.blocker, prefix="blocker_")
                target.add_raycast_spawn_mutation(c)
            else:
                raise RefUnsupported(f"mutation {mt}")

    # -- AOE / territory / tag handlers (mettagrid_c_config.py:452-545) ----------------------------------
    def aoe_configs(self, aoes):
        out = []
        for a in (aoes or {}).values():
            c = self.m.AOEConfig()
            c.radius, c.is_static, c.effect_self = int(a.radius), bool(a.is_static), bool(a.effect_self)
            self.add_filters(c, a.filters)
            self.add_mutations(c, a.mutations)
            c.presence_deltas = [self.m.ResourceDelta(self.b.rid[r], int(d)) for r, d in a.presence_deltas.items()]
            out.append(c)
        return out

    def territory_controls(self, controls, index):
        out = []
        for tc in controls or []:
            c = self.m.TerritoryControlConfig()
            c.strength, c.decay, c.territory_index = int(tc.strength), int(tc.decay), index[tc.territory]
            out.append(c)
        return out

    def tag_remove_handlers(self, cfg, handlers):
        for prefix, h in (handlers or {}).items():
            hc = self.handler_config(h, prefix)
            for tg in self.b.prefix_tags(prefix):
                cfg.add_on_tag_remove_handler(tg, hc)

    def world_extras(self, cfg_obj, src, terr_index):
        if src.aoes:
            cfg_obj.aoe_configs = self.aoe_configs(src.aoes)
        if src.territory_controls:
            cfg_obj.territory_controls = self.territory_controls(src.territory_controls, terr_index)
        if src.on_tag_remove:
            self.tag_remove_handlers(cfg_obj, src.on_tag_remove)

    def handler_config(self, h, name):
        hc = self.m.HandlerConfig(name)
        self.add_filters(hc, h.filters)
        self.add_mutations(hc, h.mutations)
        return hc

    def any_handler(self, h):
        if h is None:
            return None
        ht = getattr(h, "handler_type", "handler")
        if ht == "handler":
            return self.m.Handler(self.handler_config(h, getattr(h, "name", "") or ""))
        kids = [k for k in (self.any_handler(c) for c in h.handlers) if k is not None]
        if not kids:
            return None
        mode = self.m.HandlerMode.FirstMatch if ht == "first_match" else self.m.HandlerMode.All
        return self.m.MultiHandler(kids, mode)

    def limit_defs(self, defs):
        return [self.m.LimitDef(list(res), int(mn), int(mx), dict(mods)) for res, mn, mx, mods in defs]

    # -- the whole game ----------------------------------------------------------------------------
    def game_config(self):
        m, b, g = self.m, self.b, self.g
        terr_index = {name: i for i, name in enumerate(g.territories.keys())}
        objects = {}
        team_groups: dict[int, list] = {}
        for a in b.agent_cfgs:
            team_groups.setdefault(a.team_id if b.explicit_agents else 0, []).append(a)
        renames: dict[str, list[str]] = {}
        for gid, (team, members) in enumerate(team_groups.items()):
            gname = mc._TEAM_NAMES.get(team, f"group_{gid}")
            canonical = f"agent.{gname}"
            per_agent = []
            for idx, a in enumerate(members):
                defs, configured = [], set()
                for lim in a.inventory.limits.values():
                    defs.append(([b.rid[n] for n in lim.resources], lim.base, lim.max,
                                 {b.rid[n]: bn for n, bn in lim.modifiers.items() if n in b.rid}))  # fmt: skip
                    configured.update(lim.resources)
                for rname in b.resource_names:
                    if rname not in configured:
                        defs.append(([b.rid[rname]], b.default_limit, 65535, {}))
                inv = m.InventoryConfig()
                inv.limit_defs = self.limit_defs(defs)
                rc = m.RewardConfig()
                entries = []
                for ar in a.rewards.values():
                    e = m.RewardEntry()
                    e.reward, e.accumulate = self.value(ar.reward), bool(ar.per_tick)
                    entries.append(e)
                rc.entries = entries
                ac = m.AgentConfig(type_id=b.type_id[a.name], type_name=a.name, group_id=gid, group_name=gname,
                                   initial_vibe=int(a.vibe), inventory_config=inv, reward_config=rc,
                                   initial_inventory={b.rid[k]: int(v) for k, v in a.inventory.initial.items()})  # fmt: skip
                ac.tag_ids = [b.tid[n] for n in list(a.tags) + [f"type:{a.name}"]]
                ac.on_tick = self.any_handler(a.on_tick)
                ac.on_use_handler = self.any_handler(a.on_use_handler)
                ac.on_after_use_handler = self.any_handler(a.on_after_use_handler)
                self.world_extras(ac, a, terr_index)
                cell = f"{canonical}.{idx}"
                objects[cell] = ac
                per_agent.append(cell)
            objects[canonical] = objects[per_agent[0]]
            if len(members) > 1:
                renames[canonical] = per_agent
            aliases = [f"agent.team_{gid}"]
            if team != gid:
                aliases.append(f"agent.team_{team}")
            if gid in mc._TEAM_NAMES:
                aliases.append(f"agent.{mc._TEAM_NAMES[gid]}")
            if team in mc._TEAM_NAMES and team != gid:
                aliases.append(f"agent.{mc._TEAM_NAMES[team]}")
            if gid == 0:
                aliases += ["agent.default", "agent.agent"]
            for al in aliases:
                objects[al] = objects[canonical]
                if canonical in renames:
                    renames[al] = renames[canonical]
        for _key, oc in g.objects.items():
            tid = b.type_id[oc.name]
            if getattr(oc, "pydantic_type", "object") == "wall":
                cc = m.WallConfig(type_id=tid, type_name=oc.name, initial_vibe=int(oc.vibe))
            else:
                cc = m.GridObjectConfig(type_id=tid, type_name=oc.name, initial_vibe=int(oc.vibe))
                inv = oc.inventory
                if inv.initial:
                    cc.initial_inventory = {b.rid[k]: int(v) for k, v in inv.initial.items() if k in b.rid}
                defs, configured = [], set()
                for lim in inv.limits.values():
                    ids = [b.rid[n] for n in lim.resources if n in b.rid]
                    configured.update(lim.resources)
                    if ids:
                        defs.append((ids, lim.base, lim.max, {b.rid[n]: bn for n, bn in lim.modifiers.items() if n in b.rid}))
                for rname in inv.initial:
                    if rname not in configured and rname in b.rid:
                        defs.append(([b.rid[rname]], inv.default_limit, 65535, {}))
                if defs:
                    ic = m.InventoryConfig()
                    ic.limit_defs = self.limit_defs(defs)
                    cc.inventory_config = ic
            cc.tag_ids = [b.tid[n] for n in list(oc.tags) + [f"type:{oc.name}"]]
            cc.on_use_handler = self.any_handler(oc.on_use_handler)
            self.world_extras(cc, oc, terr_index)
            objects[oc.map_name] = cc

        gobs = g.obs.global_obs
        global_obs = m.GlobalObsConfig(
            episode_completion_pct=gobs.episode_completion_pct, last_action=gobs.last_action,
            last_action_move=gobs.last_action_move, last_reward=gobs.last_reward, goal_obs=gobs.goal_obs,
            local_position=gobs.local_position, obs=self.obs_values(gobs.obs))  # fmt: skip
        mv = g.actions.move
        move_kw = dict(allowed_directions=list(mv.allowed_directions), required_resources={}, consumed_resources={})
        if mv.handlers:
            move_kw["handlers"] = [self.handler_config(h, getattr(h, "name", "") or f"move_handler_{i}") for i, h in enumerate(mv.handlers)]
        n_vibes = len(b.vibes) if g.actions.change_vibe.enabled else 0
        actions = {
            "noop": m.ActionConfig(required_resources={}, consumed_resources={}),
            "move": m.MoveActionConfig(**move_kw),
            "attack": m.AttackActionConfig(enabled=False),
            "change_vibe": m.ChangeVibeActionConfig(required_resources={}, consumed_resources={}, number_of_vibes=n_vibes),
        }
        cfg = m.GameConfig(
            num_agents=len(b.agent_cfgs), max_steps=int(g.max_steps), episode_truncates=bool(g.episode_truncates),
            obs_width=int(g.obs.width), obs_height=int(g.obs.height), resource_names=list(b.resource_names),
            vibe_names=list(b.vibes), num_observation_tokens=int(g.obs.num_tokens), global_obs=global_obs,
            feature_ids=dict(b.feature_ids), actions=actions, objects=objects,
            tag_id_map={i: n for i, n in enumerate(b.tag_names)}, protocol_details_obs=bool(g.protocol_details_obs),
            token_value_base=int(g.obs.token_value_base), on_tick=self.any_handler(g.on_tick), **self.game_extras(terr_index))  # fmt: skip
        return cfg, renames

    def obs_values(self, obs):
        out = []
        for fname, gv in (obs or {}).items():
            c = self.m.ObsValueConfig()
            c.value, c.feature_id = self.value(gv), self.b.feature_ids[fname]
            out.append(c)
        return out

    def game_extras(self, terr_index):
        """territories / events / materialized queries (mettagrid_c_config.py:428-445,503-525,983-1002)"""
        m, b, g = self.m, self.b, self.g
        kw = {}
        if g.territories:
            terrs = [m.TerritoryConfig() for _ in terr_index]
            for name, tc in g.territories.items():
                c = terrs[terr_index[name]]
                c.tag_prefix_ids = b.prefix_tags(tc.tag_prefix)
                ctx = f"territory '{name}'"
                c.on_enter = [self.handler_config(h, f"{ctx}.on_enter.{k}") for k, h in tc.on_enter.items()]
                c.on_exit = [self.handler_config(h, f"{ctx}.on_exit.{k}") for k, h in tc.on_exit.items()]
                c.presence = [self.handler_config(h, f"{ctx}.presence.{k}") for k, h in tc.presence.items()]
            kw["territories"] = terrs
        if g.events:
            evs = {}
            for name, ev in g.events.items():
                c = m.EventConfig(ev.name)
                c.timesteps = [int(t) for t in ev.timesteps]
                c.max_targets = -1 if ev.max_targets is None else int(ev.max_targets)
                c.fallback = ev.fallback or ""
                c.set_target_query(self.query(ev.target_query))
                self.add_filters(c, ev.filters)
                self.add_mutations(c, ev.mutations)
                evs[name] = c
            kw["events"] = evs
        mqs = []
        for mq in g.materialize_queries:
            c = m.MaterializedQueryTag()
            c.tag_id = b.tid[mq.tag]
            c.set_query(self.query(mq.query))
            mqs.append(c)
        if mqs:
            kw["materialized_queries"] = mqs
        return kw


def rename_agents(grid: np.ndarray, renames: dict[str, list[str]]) -> list[list[str]]:
    counters = {k: 0 for k in renames}
    out = []
    for row in np.asarray(grid).tolist():
        new = []
        for cell in row:
            if cell in counters:
                new.append(renames[cell][counters[cell]])
                counters[cell] += 1
            else:
                new.append(cell)
        out.append(new)
    return out


class RefEnv:
    """One reference ``MettaGrid`` instance built from our config + a string map."""

    def __init__(self, cfg: Any, grid: np.ndarray, seed: int):
        game = getattr(cfg, "game", cfg)
        low = _Lower(game)
        self._cfg, renames = low.game_config()
        self.env = low.m.MettaGrid(self._cfg, rename_agents(grid, renames), int(seed) & 0xFFFFFFFF)
        self.actions = self.env.actions()
        self.vibe_actions = self.env.vibe_actions()

    def step(self, actions=None, vibe_actions=None):
        if actions is not None:
            self.actions[:] = actions
        if vibe_actions is not None:
            self.vibe_actions[:] = vibe_actions
        self.env.step()

    def __getattr__(self, name):
        return getattr(self.env, name)
