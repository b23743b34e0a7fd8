"""Scratch timing helper: python tools/quick_bench.py ENVS AGENTS [flush]  (flush = evict L2 between timed steps)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from mettagrid_b200 import workloads as cases
from mettagrid_b200.sim import BatchedSimulation
N, A = int(sys.argv[1]), int(sys.argv[2])
flush = len(sys.argv) > 3 and sys.argv[3] == "flush"
cfg = cases.benchmark_config(A)
sim = BatchedSimulation(cfg, N, seeds=42)
g = torch.Generator(device='cuda'); g.manual_seed(0)
acts = torch.randint(0, 5, (64, N, A), device='cuda', dtype=torch.int32, generator=g)
for i in range(5):
    sim.actions.copy_(acts[i]); sim.step()
torch.cuda.synchronize()
K = 50
if flush:
    buf = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        sim.actions.copy_(acts[i % 64]); buf.fill_(i & 0xff)
        ev[i][0].record(); sim.step(); ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    ms = sum(ts) / K
    print(f"N={N} A={A} kernel={sim.step_kernel} flushed: mean {ms*1000:.1f} us/step  median {ts[K//2]*1000:.1f}  min {ts[0]*1000:.1f}  {N*A/ms*1000:.3e} agent-steps/s")
else:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        sim.actions.copy_(acts[i % 64]); sim.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/K
    print(f"N={N} A={A} kernel={sim.step_kernel}: {ms*1000:.1f} us/step, {N*A/ms*1000:.3e} agent-steps/s")
sim.check_errors()
