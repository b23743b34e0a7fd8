"""Scratch timing helper for k_obs_to_grid: python tools/quick_grid_bench.py ENVS AGENTS"""
import sys
sys.path.insert(0, '.')
import torch
from mettagrid_b200 import workloads as cases
from mettagrid_b200.sim import BatchedSimulation
N, A = int(sys.argv[1]), int(sys.argv[2])
sim = BatchedSimulation(cases.benchmark_config(A), N, seeds=42)
acts = torch.randint(0, 5, (N, A), device='cuda', dtype=torch.int32)
for _ in range(3):
    sim.actions.copy_(acts); sim.step()
C, H, W = sim.grid_obs_shape()
out = torch.empty((N * A, C, H, W), dtype=torch.float32, device='cuda')
for _ in range(3):
    sim.grid_observations(out=out)
torch.cuda.synchronize()
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    sim.grid_observations(out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
nbytes = out.numel() * 4 + N * A * sim.num_tokens * 3
print(f"k_obs_to_grid rows={N*A} C={C} H={H} W={W}: {ms*1000:.1f} us, {nbytes/ms/1e6:.0f} GB/s (output {out.numel()*4/1e6:.0f} MB, larger than L2)")
