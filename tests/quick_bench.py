import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from tests import cases
from mettagrid_b200.sim import BatchedSimulation
N, A = int(sys.argv[1]), int(sys.argv[2])
cfg = cases.benchmark_config(A)
t0=time.time()
sim = BatchedSimulation(cfg, N, seeds=42)
print("create", time.time()-t0, "state MB", sim.state_bytes/1e6)
g = torch.Generator(device='cuda'); g.manual_seed(0)
acts = torch.randint(0, 5, (64, N, A), device='cuda', dtype=torch.int32, generator=g)
for i in range(5):
    sim.actions.copy_(acts[i]); sim.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K=50
e0.record()
for i in range(K):
    sim.actions.copy_(acts[i % 64]); sim.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/K
print(f"N={N} A={A}: {ms*1000:.1f} us/step, {N*A/ms*1000:.3e} agent-steps/s")
sim.check_errors()
