"""Host logic without a GPU: the compiled-program format, the compiler's lowering rules, and that the
C-ABI shared library loads and exports every symbol include/mettagrid_b200.h declares."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from mettagrid_b200 import config as C
from mettagrid_b200 import compiler as mc
from mettagrid_b200.compiler import K, compile_config, observation_offsets, pybind_dict_order
from tests import cases

ROOT = Path(__file__).resolve().parent.parent


def test_header_constants_parse():
    assert K["MG_MAGIC"] == 0x4D47B200
    assert K["MGH_HEADER_WORDS"] > K["MGS_POOL"] > K["MGS_OFFSETS"] > K["MGH_MAX_TERR_SOURCES"]
    assert K["MGF_PERIODIC"] == 11 and K["MGM_PUSH_OBJECT"] == 17 and K["MGV_MIN"] == 8


def test_observation_offsets_golden_order():
    # golden order from the reference's own C++ test (tests/test_observations.cpp:84-104) for a 3x3 window,
    # and the shape sizes the reference documents: 13x13 -> 121, 11x11 -> 89, 7x7 -> 37, 3x3 -> 5
    assert observation_offsets(3, 3) == [(0, 0), (-1, 0), (0, -1), (0, 1), (1, 0)]
    assert [len(observation_offsets(n, n)) for n in (13, 11, 7, 3)] == [121, 89, 37, 5]
    offs = observation_offsets(13, 13)
    d = [abs(r) + abs(c) for r, c in offs]
    assert d == sorted(d) and offs[0] == (0, 0) and offs[1:5] == [(-1, 0), (0, -1), (0, 1), (1, 0)]


def test_benchmark_program_shape():
    p = compile_config(cases.benchmark_config(16))
    assert p.action_names[:5] == ["noop", "move_north", "move_south", "move_west", "move_east"]
    assert len(p.action_names) == 157 and p.action_names[5] == "change_vibe_default"
    assert p.feature_ids["tag"] == 6 and p.feature_ids["agent:group"] == 0 and p.feature_ids["agent_id"] == 11
    assert p.hdr("MGH_NUM_OFFSETS") == 121 and p.hdr("MGH_MAX_PRIORITY") == 1
    assert p.blob[K["MGH_TOTAL_WORDS"]] == p.blob.size
    for key in [k for k in K if k.startswith("MGS_")]:
        assert p.hdr(key) % 4 == 0, f"{key} must be 16-byte aligned"
    # the default move chain: relocate, then use
    chain = p.blob[p.hdr("MGS_MOVE_CHAIN") :][: 2 * K["MG_MOVEH_WORDS"]].reshape(2, -1)
    assert chain[0, 3] == K["MGMB_RELOCATE"] and chain[1, 3] == K["MGMB_USE_TARGET"]


def test_agent_renames_and_map_encoding():
    cfg = cases.combat_config()
    grid = cases.combat_map(3, seed=1)
    p = compile_config(cfg, *grid.shape)
    assert p.agent_renames["agent.red"] == ["agent.red.0", "agent.red.1", "agent.red.2"]
    cells, gs = p.encode_map(grid, with_stats=True)
    names = [p.template_names[t] for t in cells.reshape(-1) if t >= 0 and p.template_names[t].startswith("agent.")]
    assert sorted(names) == sorted(["agent.red.0", "agent.red.1", "agent.red.2", "agent.blue.0", "agent.blue.1", "agent.blue.2"])
    assert gs[p.game_stat_names.index("objects.wall")] == float((grid == "wall").sum())
    with pytest.raises(mc.CompileError, match="Unknown object type"):
        g2 = grid.copy()
        g2[1, 1] = "dragon"
        p.encode_map(g2)


def test_lowering_rules():
    # multi-resource filter fans out; unknown vibe filters are dropped; periodic start defaults to period
    b = mc._Builder(cases.combat_config().game)
    b._spawn_refs = []
    b.build_id_maps()
    b.build_feature_ids()
    f0, n = b.filter_list([C.actorHas({"hp": 1, "loot": 2}), C.actorVibe("no-such-vibe"), C.PeriodicFilter(period=9)])
    assert n == 3
    assert [f[0] for f in b.filters[f0 : f0 + n]] == [K["MGF_RESOURCE"], K["MGF_RESOURCE"], K["MGF_PERIODIC"]]
    assert b.filters[f0 + 2][2:4] == [9, 9]
    # OR of a multi-resource filter keeps AND semantics through a double negation
    f0, n = b.filter_list([C.anyOf([C.actorHas({"hp": 1, "loot": 2}), C.actorVibe("swords")])])
    orf = b.filters[f0]
    kids = b.filters[orf[2] : orf[2] + orf[3]]
    assert [k[0] for k in kids] == [K["MGF_NEG"], K["MGF_VIBE"]]
    with pytest.raises(mc.CompileError):
        mc._Builder(C.GameConfig(resource_names=[f"r{i}" for i in range(14)], num_agents=1)).build_id_maps()


def test_pybind_dict_order():
    # verified against the reference: {hp:0, weapon:1, armor:2, mobility:3, energy:5} iterates 3,2,1,5,0
    assert pybind_dict_order([0, 1, 2, 3, 5]) == [3, 2, 1, 5, 0]
    assert pybind_dict_order([0, 1, 2]) == [2, 1, 0]
    assert pybind_dict_order([]) == []


def _declared_symbols():
    text = (ROOT / "include" / "mettagrid_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z_]+)\s*\(", text)))


def test_capi_exports_every_declared_symbol():
    from mettagrid_b200 import native
    from mettagrid_b200.build import build_native

    build_native()
    lib = ctypes.CDLL(str(native.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/mettagrid_b200.h but not exported"
    assert sorted(native.SYMBOLS) == declared
    # no compute without a GPU: a bad program is rejected before any CUDA call
    out = ctypes.c_void_p()
    L = native.lib()
    bad = np.zeros(8, dtype=np.int32)
    assert L.mg_create(bad.ctypes.data, bad.size, 1, None, None, None, 0, ctypes.byref(out)) == native.MG_E_INVALID
    assert b"compiled game program" in L.mg_last_error(None)


def test_bad_template_index_is_rejected_before_any_cuda_call():
    from mettagrid_b200 import native
    from mettagrid_b200.build import build_native

    build_native()
    p = compile_config(cases.benchmark_config(2))
    L = native.lib()
    out = ctypes.c_void_p()
    seeds = np.zeros(1, dtype=np.uint32)
    # a NULL map is refused up front
    rc = L.mg_create(p.blob.ctypes.data, p.blob.size, 1, None, None, seeds.ctypes.data, 0, ctypes.byref(out))
    assert rc == native.MG_E_INVALID and b"init_cells" in L.mg_last_error(None)
