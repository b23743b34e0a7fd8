"""Scratch timing helper: vec-env steps per second with auto-reset traffic.  python tools/quick_vecenv_bench.py ENVS MAX_STEPS"""
import sys, time
sys.path.insert(0, '.')
import torch
from mettagrid_b200 import workloads as cases
from mettagrid_b200.vecenv import MettaGridVecEnv
N, MS = int(sys.argv[1]), int(sys.argv[2])
VALIDATE = not (len(sys.argv) > 3 and sys.argv[3] == 'novalidate')
env = MettaGridVecEnv(cases.benchmark_config(16, max_steps=MS), N, seed=1, desync_episodes=MS > 0, validate=VALIDATE)
acts = torch.randint(0, 5, (32, N * 16), device='cuda')
for i in range(20):
    env.step(acts[i % 32])
torch.cuda.synchronize()
K = 200
t0 = time.perf_counter()
for i in range(K):
    env.step(acts[i % 32])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print(f"vec-env N={N} max_steps={MS} validate={VALIDATE}: {dt*1e6:.1f} us/step wall, {N*16/dt:.3e} agent-steps/s, episodes finished {env.episodes_finished}")
