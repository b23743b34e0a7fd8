"""Debug helper (GPU box): run one case on the CUDA path and the oracle side by side, print the first difference."""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from mettagrid_b200.sim import BatchedSimulation
from oracle.oracle import OracleEnv
from mettagrid_b200 import workloads as cases

name, seed, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = getattr(cases, f"{name}_config")(None)
grid = getattr(cases, f"{name}_map")(seed=seed)
sim = BatchedSimulation(cfg, 1, seeds=[seed], maps=[grid])
P = sim.program
o = OracleEnv(P, sim._init_cells[0], seed, sim._init_gstats[0])
inv = {v: k for k, v in P.feature_ids.items()}
nprim = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
prim, vibe = cases.random_actions(np.random.RandomState(seed), steps, (1, P.num_agents), nprim, len(P.action_names), 0.1)


def check(t):
    torch.cuda.synchronize()
    a, b = sim.observations.cpu().numpy()[0], o.observations()
    bad = False
    if not np.array_equal(a, b):
        bad = True
        for ag in range(P.num_agents):
            if not np.array_equal(a[ag], b[ag]):
                k = int(np.argmax((a[ag] != b[ag]).any(1)))
                print("step", t, "agent", ag, "token", k)
                print(" gpu   ", [(x[0], inv.get(x[1]), x[2]) for x in a[ag][max(0, k - 3) : k + 6].tolist()])
                print(" oracle", [(x[0], inv.get(x[1]), x[2]) for x in b[ag][max(0, k - 3) : k + 6].tolist()])
    r1, r2 = sim.rewards.cpu().numpy()[0], o.rewards()
    if not np.array_equal(r1.view(np.uint32), r2.view(np.uint32)):
        bad = True
        print("step", t, "rewards gpu", r1, "oracle", r2)
    if not np.array_equal(sim.action_success()[0], o.action_success()):
        bad = True
        print("step", t, "success gpu", sim.action_success()[0], "oracle", o.action_success(), "actions", prim[t], vibe[t])
    s1, s2 = sim.get_episode_stats(0), o.get_episode_stats()
    if s1 != s2:
        bad = True
        for ag in range(P.num_agents):
            d = {k: (s1["agent"][ag].get(k), s2["agent"][ag].get(k)) for k in set(s1["agent"][ag]) | set(s2["agent"][ag])
                 if s1["agent"][ag].get(k) != s2["agent"][ag].get(k)}
            if d:
                print("step", t, "agent", ag, "stats (gpu, oracle)", d)
        d = {k: (s1["game"].get(k), s2["game"].get(k)) for k in set(s1["game"]) | set(s2["game"]) if s1["game"].get(k) != s2["game"].get(k)}
        if d:
            print("step", t, "game stats", d)
    d1, d2 = sim.dump_objects(0), o.dump_objects()
    if d1.shape != d2.shape or not np.array_equal(d1, d2):
        bad = True
        print("step", t, "object dumps differ", d1.shape, d2.shape)
        for i in range(min(len(d1), len(d2))):
            if not np.array_equal(d1[i], d2[i]):
                print("  gpu   ", d1[i].tolist())
                print("  oracle", d2[i].tolist())
                break
    try:
        sim.check_errors()
    except Exception as e:
        print("error:", e)
        bad = True
    return bad


if not check(-1):
    for t in range(steps):
        sim.step(prim[t], vibe[t])
        o.step(prim[t, 0], vibe[t, 0])
        if check(t):
            break
    else:
        print("no difference in", steps, "steps")
