"""Scratch timing helper for the generic kernel: python tools/quick_bench_generic.py c3|c4 ENVS"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench
wl, N = sys.argv[1], int(sys.argv[2])
A = bench.WORKLOADS[wl][2]
cfg = bench.make_cfg(A, wl)
from mettagrid_b200.sim import BatchedSimulation
pool = [bench.make_map(cfg, A, wl, 42 + i) for i in range(16)]
sim = BatchedSimulation(cfg, N, seeds=42, maps=[pool[e % 16] for e in range(N)])
P = sim.program
nprim = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
prim, vibe = bench.gen_actions(len(P.action_names), nprim, 16, N, A, 7)
prim, vibe = torch.from_numpy(prim).cuda(), torch.from_numpy(vibe).cuda()
for i in range(3):
    sim.actions.copy_(prim[i]); sim.vibe_actions.copy_(vibe[i]); sim.step()
torch.cuda.synchronize()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    sim.actions.copy_(prim[i % 16]); sim.vibe_actions.copy_(vibe[i % 16]); sim.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
sim.check_errors()
print(f"{wl} N={N} A={A} kernel={sim.step_kernel}: {ms*1000:.0f} us/step, {N*A/ms*1000:.3e} agent-steps/s")
