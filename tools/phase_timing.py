"""Per-phase cycle split of k_world, the first kernel of the generic tick, (needs the -DMG_PHASE_TIMING variant):
    python -m mettagrid_b200.build --variant prof -DMG_PHASE_TIMING
    METTAGRID_B200_LIB=$PWD/variants/lib_prof.so python tools/phase_timing.py c3 16384
"""
import ctypes
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from mettagrid_b200 import native
from mettagrid_b200 import workloads as W
from mettagrid_b200.sim import BatchedSimulation

wl, N = sys.argv[1], int(sys.argv[2])
A = W.WORKLOADS[wl][2]
cfg = W.make_cfg(A, wl)
maps = [W.make_map(cfg, A, wl, e) for e in range(N)] if wl in ("c3", "c4") else None
sim = BatchedSimulation(cfg, N, seeds=42, maps=maps)
P = sim.program
nprim = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
prim, vibe = W.gen_actions(len(P.action_names), nprim, 16, N, A, 7)
prim, vibe = torch.from_numpy(prim).cuda(), torch.from_numpy(vibe).cuda()
L = native.lib()
out = (ctypes.c_ulonglong * 16)()
for i in range(4):
    sim.actions.copy_(prim[i]); sim.vibe_actions.copy_(vibe[i]); sim.step()
torch.cuda.synchronize()
L.mg_debug_phase_cycles(out, 1)
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    sim.actions.copy_(prim[i % 16]); sim.vibe_actions.copy_(vibe[i % 16]); sim.step()
e1.record(); torch.cuda.synchronize()
L.mg_debug_phase_cycles(out, 0)
names = sys.argv[3].split(",") if len(sys.argv) > 3 else ["prologue", "shuffle+actions", "bookkeeping", "events+on_tick", "aoe+territory+game_on_tick", "coverage (+ hand-over follows)"]
tot = sum(out)
print(f"{wl} N={N}: {e0.elapsed_time(e1) / K * 1000:.0f} us/tick (kernel {sim.step_kernel})")
for n, c in zip(names, out):
    print(f"  {n:30s} {100.0 * c / max(tot, 1):5.1f} %   {c / K / N:9.0f} cycles per env-tick")
