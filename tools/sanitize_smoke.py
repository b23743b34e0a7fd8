"""A small pass over every kernel for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Every step kernel family (fast path, plain, generic with handlers / AOE / territory / events), reset with a mask,
set_buffers mid-episode, the host-buffer step, the vec-env step with auto-reset, step_info_keys and the dense grid
observations run for a few ticks on a handful of environments."""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from mettagrid_b200 import workloads as W
from mettagrid_b200.sim import BatchedSimulation
from mettagrid_b200.vecenv import MettaGridVecEnv

TICKS = int(sys.argv[1]) if len(sys.argv) > 1 else 12
runs = [
    ("fast path", W.benchmark_config(16), None, 5),
    ("plain (walled, 8-way)", W.toy_config(20, 24), None, 9),
    ("combat (handler chains)", W.combat_config(None, 3), [W.combat_map(3, seed=s) for s in range(6)], 9),
    ("world (AOE, territory, events, spawn)", W.world_config(None, 3), [W.world_map(3, seed=s) for s in range(6)], 9),
    ("network (queries, push)", W.network_config(None, 4), [W.network_map(4, seed=s) for s in range(6)], 5),
]
for name, cfg, maps, nprim in runs:
    sim = BatchedSimulation(cfg, 6, seeds=3, maps=maps)
    P = sim.program
    prim, vibe = W.random_actions(np.random.RandomState(0), TICKS, (6, P.num_agents), nprim, len(P.action_names), 0.2, 0.02)
    h_obs = np.zeros((6, P.num_agents, P.num_tokens, 3), np.uint8)
    h_rew, h_t, h_u = np.zeros((6, P.num_agents), np.float32), np.zeros((6, P.num_agents), np.uint8), np.zeros((6, P.num_agents), np.uint8)
    for t in range(TICKS):
        if t % 4 == 3:
            sim.step_host(prim[t], vibe[t], h_obs, h_rew, h_t, h_u)
        else:
            sim.step(prim[t], vibe[t])
        if t == TICKS // 2:
            sim.reset(env_mask=torch.tensor([0, 1, 0, 0, 1, 0], dtype=torch.bool, device="cuda"))
            sim.get_episode_stats(1)
            sim.set_buffers(sim.observations, sim.terminals, sim.truncations, sim.rewards, sim.actions, sim.vibe_actions)
    sim.grid_observations()
    torch.cuda.synchronize()
    sim.check_errors()
    sim.get_episode_stats(0), sim.dump_objects(0)
    sim.close()
    print("ok:", name, "kernel", P.num_agents)
env = MettaGridVecEnv(W.benchmark_config(4, max_steps=5), 5, seed=1, step_info_keys=["game/tokens_written", "agent/action.failed", "agent/reward_episode"])
for t in range(TICKS):
    env.step(torch.randint(0, env.num_primary, (20,), device="cuda"))
env.step_info_payload(0)
env.close()
print("ok: vec-env")
