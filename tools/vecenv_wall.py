"""Wall-clock cost of one vectorised-env step in a training-style loop (no synchronisation inside the loop):
python tools/vecenv_wall.py [c2|toy] [ENVS].  Compare with METTAGRID_B200_NO_GRAPH=1 (plain launches)."""
import sys
import time

sys.path.insert(0, ".")
import torch

from mettagrid_b200 import workloads as W
from mettagrid_b200.vecenv import MettaGridVecEnv

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
A = W.WORKLOADS[wl][2]
cfg = W.make_cfg(A, wl)
cfg.game.max_steps = 200
maps = [W.make_map(cfg, A, wl, e) for e in range(N)] if wl == "toy" else None
env = MettaGridVecEnv(cfg, N, seed=1, validate=False, maps=maps)
acts = torch.randint(0, env.num_primary, (N * A,), device="cuda", dtype=torch.int64)
for _ in range(20):
    env.step(acts)
torch.cuda.synchronize()
K = 2000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(K):
    env.step(acts)
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / K * 1e6
print(f"{wl} N={N} kernel={env.sim.step_kernel}: {wall:.1f} us wall per vec-env step, {e0.elapsed_time(e1) / K * 1e3:.1f} us on the device, "
      f"{N * A / wall * 1e6:.3e} agent-steps/s")
