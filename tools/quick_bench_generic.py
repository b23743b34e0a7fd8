"""Scratch timing helper for the generic kernels: python tools/quick_bench_generic.py c3|c4|toy ENVS [flush]"""
import sys

sys.path.insert(0, ".")
import torch

from mettagrid_b200 import workloads as W
from mettagrid_b200.sim import BatchedSimulation

wl, N = sys.argv[1], int(sys.argv[2])
do_flush = len(sys.argv) > 3 and sys.argv[3] == "flush"
A = W.WORKLOADS[wl][2]
cfg = W.make_cfg(A, wl)
maps = [W.make_map(cfg, A, wl, e) for e in range(N)] if wl in ("c3", "c4") else None
sim = BatchedSimulation(cfg, N, seeds=42, maps=maps)
P = sim.program
nprim = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
prim, vibe = W.gen_actions(len(P.action_names), nprim, 16, N, A, 7)
prim, vibe = torch.from_numpy(prim).cuda(), torch.from_numpy(vibe).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if do_flush else None
for i in range(5):
    sim.actions.copy_(prim[i]); sim.vibe_actions.copy_(vibe[i]); sim.step()
torch.cuda.synchronize()
K = 20
ms = 0.0
for i in range(K):
    sim.actions.copy_(prim[i % 16]); sim.vibe_actions.copy_(vibe[i % 16])
    if flush is not None:
        flush.fill_(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.step(); e1.record()
    torch.cuda.synchronize()
    ms += e0.elapsed_time(e1)
ms /= K
sim.check_errors()
ab = W.algo_bytes(P, wl)
print(f"{wl} N={N} A={A} kernel={sim.step_kernel}: {ms*1000:.0f} us/tick, {N*A/ms*1000:.3e} agent-steps/s, "
      f"{ab*N*A/ms/1e6:.0f} GB/s algorithmic = {ab*N*A/ms/1e6/6455.6:.4f} of 6455.6")
