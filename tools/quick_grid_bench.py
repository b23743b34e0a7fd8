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
# token-policy front end (k_token_summary) next to the torch expression it replaces (token_encoder.py:89-113)
from mettagrid_b200.token_encoder import TokenSummary
mod = TokenSummary(sim, 192)
summ = torch.empty((N * A, 192), dtype=torch.float32, device='cuda')
for _ in range(3):
    mod(out=summ)
torch.cuda.synchronize()
e0.record()
for _ in range(K):
    mod(out=summ)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
tb = N * A * sim.num_tokens * 3 + summ.numel() * 4
print(f"k_token_summary rows={N*A} hidden=192: {ms*1000:.1f} us, {tb/ms/1e6:.0f} GB/s of rows in + summaries out")
def torch_ref(tokens):
    cb = tokens[..., 0].to(torch.long)
    emb = mod.pos_x_embed(cb & 15) + mod.pos_y_embed((cb >> 4) & 15) + mod.feature_embed(tokens[..., 1].to(torch.long))
    emb = emb * (tokens[..., 2].float() / (mod._feature_scale[tokens[..., 1].to(torch.long)] + 1e-6)).unsqueeze(-1)
    valid = cb != 0xFF
    emb = emb * valid.unsqueeze(-1)
    return emb.sum(-2) / valid.sum(-1, keepdim=True).clamp_min(1).float().sqrt()
try:
    rows = sim.observations.view(-1, sim.num_tokens, 3)[: min(N * A, 16384)]
    with torch.no_grad():
        for _ in range(2):
            torch_ref(rows)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            torch_ref(rows)
        e1.record(); torch.cuda.synchronize()
    print(f"torch expression on {rows.shape[0]} rows: {e0.elapsed_time(e1)/5*1000:.1f} us (x{N*A/rows.shape[0]:.0f} for all rows)")
except Exception as ex:
    print("torch reference skipped:", ex)
