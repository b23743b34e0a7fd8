"""Namespace of the REFERENCE's config classes under the names tests/cases.py uses, so one case
builder produces either our mirror dataclasses or the reference's Pydantic models (build container
only)."""

from types import SimpleNamespace


def reference_namespace():
    import mettagrid.config.filter as F
    import mettagrid.config.game_value as GV
    import mettagrid.config.handler_config as HC
    import mettagrid.config.mettagrid_config as MC
    import mettagrid.config.mutation as MU
    import mettagrid.config.reward_config as RC
    from mettagrid.config.event_config import EventConfig, once, periodic
    from mettagrid.config.mutation.change_vibe_mutation import ChangeVibeMutation
    from mettagrid.config.mutation.game_value_mutation import SetGameValueMutation
    from mettagrid.config.mutation.push_object_mutation import PushObjectMutation
    from mettagrid.config.territory_config import TerritoryConfig, TerritoryControlConfig
    from mettagrid.config.vibes import VIBES, Vibe

    ns = SimpleNamespace()
    for mod in (MC, HC, F, MU, GV, RC):
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(ns, k, v)
    ns.EventConfig, ns.once, ns.periodic = EventConfig, once, periodic
    ns.TerritoryConfig, ns.TerritoryControlConfig = TerritoryConfig, TerritoryControlConfig
    ns.VIBES, ns.Vibe = VIBES, Vibe
    ns.ChangeVibeMutation, ns.SetGameValueMutation, ns.PushObjectMutation = ChangeVibeMutation, SetGameValueMutation, PushObjectMutation
    ns.is_reference = True
    return ns
