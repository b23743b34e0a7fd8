#!/usr/bin/env python
"""Benchmark of the batched MettaGrid step (BASELINE.json metric: agent-steps/s including observations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--agents A]

Workload (config.workload): BASELINE.json configs[1] ("C2") -- the reference benchmark game
(benchmarks/test_mettagrid_env_benchmark.py:21-29: 20x20 RandomMapBuilder map, 16 agents, 13x13
observation window, 100 tokens, noop + 4-way move + 152 vibes, max_steps=0) batched to 4096 envs per
GPU, env e using map seed 42+e and env seed 42+e, "effective" random actions.

A step = one tick of every env.  `value` is measured with all inputs resident in HBM (CUDA events
around mg_step only, L2 flushed between timed steps); `e2e` is the same tick through the C ABI's
host-buffer entry point (mg_step_host: pinned host actions in, observations/rewards/flags out).

`--impl reference` times the UNMODIFIED reference C++ step (oracle/_ref, built from /root/reference
by oracle/Makefile.ref) on the host cores: one worker process per core, each stepping its own share
of a bounded sample of the same workload.  The only places this file touches oracle/ are that arm
and the `cpu_baseline` leg.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "agent_steps_per_sec_incl_obs"
UNIT = "agent-steps/s"
ALGO_BYTES_PER_AGENT_STEP = 550.0  # C1/C2 (T=100, 20x20, R=10, A=16)


WORKLOADS = {
    # name: (description, default envs per GPU, agents)
    "c2": ("C2: reference benchmark game 20x20, {A} agents/env, 13x13 obs, 100 tokens", 4096, 16),
    "c3": ("C3: combat-heavy game 37x37 (25x25 + border 6), {A} agents in 2 teams, handler chains, 11x11 obs, 500 tokens", 16384, 24),
    "c4": ("C4: world game 66x66, {A} agents, fixed+mobile AOE, territory + aoe_mask, events, queries, 11x11 obs, 200 tokens", 8192, 24),
}


def make_cfg(agents: int, workload: str = "c2"):
    from tests import cases

    if workload == "c3":
        return cases.combat_config(None, agents // 2, num_tokens=500)
    if workload == "c4":
        return cases.world_config(None, agents // 2, num_tokens=200)
    return cases.benchmark_config(agents)


def make_map(cfg, agents: int, workload: str, seed: int):
    from mettagrid_b200.mapgen import RandomMapConfig, random_map

    if workload == "c3":
        return random_map(RandomMapConfig(width=37, height=37, border_width=6, seed=seed,
                                          agents={"red": agents // 2, "blue": agents // 2},
                                          objects={"wall": 10, "chest": 6, "altar": 4}))  # fmt: skip
    if workload == "c4":
        return random_map(RandomMapConfig(width=66, height=66, border_width=1, seed=seed,
                                          agents={"red": agents // 2, "blue": agents // 2},
                                          objects={"wall": 120, "healer": 6, "spikes": 6, "beacon_red": 5, "beacon_blue": 5,
                                                   "mine": 12, "vault": 6}))  # fmt: skip
    return random_map(cfg.game.map_builder, seed=seed)


def algo_bytes(P, workload: str) -> float:
    """B_io + B_state per agent-step."""
    if workload == "c2":
        return ALGO_BYTES_PER_AGENT_STEP
    T, A, R = P.num_tokens, P.num_agents, len(P.resource_names)
    b_io = 3 * T + 4 + 4 + 4 + 1 + 1 + 8
    b_state = (2 * P.height * P.width + 16) / A + 2 * (4 + 4 + 4 + 1 + 2 * R + 4 * R) + 4 * 8
    return float(b_io + b_state)


ACTION_MODE = os.environ.get("METTAGRID_BENCH_ACTIONS", "effective")  # --actions (inherited by the CPU workers)


def gen_actions(num_actions: int, num_primary: int, steps: int, envs: int, agents: int, seed: int):
    """SURVEY 8(d), both from np.random.RandomState(...).randint like the reference's benchmark
    (benchmarks/test_mettagrid_env_benchmark.py:44-49):
    'effective' (default): primary uniform over the non-vibe actions, vibe change with p = 0.1;
    'verbatim': uniform over ALL action ids into the primary buffer, vibe buffer left 0 -- about 97 % of the ids
    are change_vibe ids, which the primary stream ignores."""
    rng = np.random.RandomState(seed)
    if ACTION_MODE == "verbatim":
        prim = rng.randint(0, num_actions, size=(steps, envs, agents)).astype(np.int32)
        return prim, np.zeros_like(prim)
    prim = rng.randint(0, num_primary, size=(steps, envs, agents)).astype(np.int32)
    vibe = np.zeros_like(prim)
    m = rng.rand(steps, envs, agents) < 0.1
    vibe[m] = rng.randint(num_primary, num_actions, size=int(m.sum()))
    return prim, vibe


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference C++ env on host cores
# ------------------------------------------------------------------------------------------------
def _ref_worker(rank: int, agents: int, envs: int, env0: int, steps: int, warmup: int, barrier, out_q, workload="c2"):
    try:
        os.sched_setaffinity(0, {rank % os.cpu_count()})
    except Exception:
        pass
    sys.path.insert(0, str(ROOT))
    from mettagrid_b200.compiler import compile_config
    from oracle.ref_driver import RefEnv, RefUnsupported

    cfg = make_cfg(agents, workload)
    sims, kind = [], "reference"
    for e in range(envs):
        grid = make_map(cfg, agents, workload, 42 + env0 + e)
        try:
            sims.append(RefEnv(cfg, grid, 42 + env0 + e))
        except RefUnsupported:  # the GPU-box driver cannot build this game for the reference: time the oracle port
            from oracle.oracle import OracleEnv

            kind = "port"
            prog = compile_config(cfg, *grid.shape)
            cells, gs = prog.encode_map(grid, with_stats=True)
            sims.append(OracleEnv(prog, cells, 42 + env0 + e, gs))
    prog = compile_config(cfg, *make_map(cfg, agents, workload, 0).shape)
    nprim = sum(1 for n in prog.action_names if not n.startswith("change_vibe_"))
    prim, vibe = gen_actions(len(prog.action_names), nprim, 64, envs, agents, 1000 + rank)
    for t in range(warmup):
        for e, s in enumerate(sims):
            s.step(prim[t % 64, e], vibe[t % 64, e])
    barrier.wait()
    t0 = time.perf_counter()
    for t in range(steps):
        for e, s in enumerate(sims):
            s.step(prim[t % 64, e], vibe[t % 64, e])
    dt = time.perf_counter() - t0
    barrier.wait()
    out_q.put((rank, dt, kind))


def run_reference(agents: int, steps: int, warmup: int, envs_per_worker: int, workload: str = "c2"):
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    P = os.cpu_count() or 1
    barrier = ctx.Barrier(P)
    q = ctx.Queue()
    procs = [ctx.Process(target=_ref_worker, args=(r, agents, envs_per_worker, r * envs_per_worker, steps, warmup, barrier, q, workload))
             for r in range(P)]  # fmt: skip
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    times = [r[1] for r in res]
    kind = "port" if any(r[2] == "port" for r in res) else "reference"
    for p in procs:
        p.join()
    dt = max(times)
    total = P * envs_per_worker * agents * steps
    return {"value": total / dt, "ms_per_step": dt / steps * 1e3, "cores": P, "envs": P * envs_per_worker, "seconds": dt, "kind": kind}


def reference_available() -> bool:
    from oracle import reference

    return reference.so_path() is not None


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")  # fmt: skip
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}  # fmt: skip


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from mettagrid_b200.sim import BatchedSimulation

    from mettagrid_b200.shard import env_seeds, reduce_agent_stats, shard_range

    wl = args.workload
    envs = args.envs or WORKLOADS[wl][1]
    A = args.agents or WORKLOADS[wl][2]
    cfg = make_cfg(A, wl)
    env0, env1 = shard_range(world * envs, rank, world)  # weak scaling: `envs` environments per GPU
    if wl == "c2":
        maps = None
        map_seeds = [42 + e for e in range(env0, env1)]
    else:  # map instances are expensive to build on the host: cycle a pool of distinct maps
        pool = [make_map(cfg, A, wl, 42 + i) for i in range(64)]
        maps, map_seeds = [pool[e % 64] for e in range(env0, env1)], None
    sim = BatchedSimulation(cfg, envs, seeds=env_seeds(42, env0, env1), maps=maps, map_seeds=map_seeds, device=local_rank)
    P = sim.program
    num_actions = len(P.action_names)
    num_primary = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
    POOL = 16
    prim_h, vibe_h = gen_actions(num_actions, num_primary, POOL, envs, A, 7 + rank)
    prim = torch.from_numpy(prim_h).cuda()
    vibe = torch.from_numpy(vibe_h).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    steps, warmup = args.steps, max(args.warmup, 3)

    def one_step(i):
        sim.actions.copy_(prim[i % POOL])
        sim.vibe_actions.copy_(vibe[i % POOL])

    for i in range(warmup):
        one_step(i)
        sim.step()
    torch.cuda.synchronize()
    sim.check_errors()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(steps):
        one_step(warmup + i)  # inputs resident in HBM before the timed region of this step
        flush.fill_(i & 0xFF)  # evict L2 between timed iterations
        ev[i][0].record()
        sim.step()
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = steps  # one k_step launch per tick
    clocks = sampler.stop()

    # ---- e2e: host buffers through the C ABI (pinned), copies inside the timed region
    NA = envs * A
    h_prim = torch.from_numpy(prim_h[0].copy()).pin_memory()
    h_vibe = torch.from_numpy(vibe_h[0].copy()).pin_memory()
    h_obs = torch.empty((envs, A, P.num_tokens, 3), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty((envs, A), dtype=torch.float32).pin_memory()
    h_term = torch.empty((envs, A), dtype=torch.uint8).pin_memory()
    h_trunc = torch.empty((envs, A), dtype=torch.uint8).pin_memory()
    hs = lambda t: t.numpy()  # noqa: E731
    for i in range(3):
        sim.step_host(hs(h_prim), hs(h_vibe), hs(h_obs), hs(h_rew), hs(h_term), hs(h_trunc))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        h_prim.copy_(torch.from_numpy(prim_h[i % POOL]))  # next step's host inputs (host memcpy, part of e2e)
        h_vibe.copy_(torch.from_numpy(vibe_h[i % POOL]))
        sim.step_host(hs(h_prim), hs(h_vibe), hs(h_obs), hs(h_rew), hs(h_term), hs(h_trunc))
    e2e_s = time.perf_counter() - t0
    e2e_sum = float(h_rew.sum()) + float(h_obs[0, 0, 0, 0])  # consume the result on the host
    sim.check_errors()

    t = torch.tensor([total_ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the one collective the path has: episode-stat reduction over NVLink (envs/stats_tracker.py:40-45);
    # outside the timed region, a few KB
    av = np.concatenate([sim.stats_arrays(e)[0] for e in range(min(envs, 8))])
    mean_stats = reduce_agent_stats(torch.from_numpy(av).cuda(), torch.tensor(av.shape[0], device="cuda"))
    total_ms, e2e_ms = float(t[0]), float(t[1])
    agent_steps = world * envs * A * steps
    value = agent_steps / (total_ms * 1e-3)
    e2e_value = agent_steps / (e2e_ms * 1e-3)

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured"
        else:
            peak, peak_src = 6650.0, "fallback"
        per_launch_ms = total_ms / steps
        ab = algo_bytes(P, wl)
        achieved = ab * envs * A / (per_launch_ms * 1e-3) / 1e9
        # which kernel mg_step launches for this handle (include/mettagrid_b200.h: mg_step_kernel)
        sk = sim.step_kernel
        kernel = f"k_step_fast<{sk}>" if sk >= 8 else ("k_step<true>" if sk == 1 else "k_step<false>")
        traffic = None
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists() and wl == "c2" and envs == WORKLOADS[wl][1]:  # measured for the default workload only
            traffic = json.loads(tpath.read_text()).get(kernel, {}).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/i32 (f32 stats)", "data": "synthetic",
            "config": {"workload": WORKLOADS[wl][0].format(A=A) + f", {envs} envs/GPU",
                       "envs_per_gpu": envs, "agents_per_env": A, "actions": (f"verbatim: primary uniform over all {num_actions} ids, vibe buffer 0" if ACTION_MODE == "verbatim"
                                   else f"primary uniform over {num_primary}, vibe p=0.1"),
                       "l2": "flushed between timed steps (256 MB fill)", "obs_write_GBps": value * 3 * P.num_tokens / 1e9,
                       "parallelism": f"env-sharded x{world}, no per-step collective",
                       "mean_move_success_per_agent": float(mean_stats[P.agent_stat_names.index("action.move.success")])},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * NA * 4,
                    "d2h_bytes_per_step": NA * (3 * P.num_tokens + 4 + 1 + 1), "check": e2e_sum},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kernel, "peak_source": peak_src,
                         "algorithmic_bytes_per_agent_step": ab},
        }  # fmt: skip
        if world == 1 and not args.no_cpu_baseline and reference_available():
            sim.close()
            cpu_steps = args.cpu_steps if wl == "c2" else max(200, args.cpu_steps // 20)
            cb = run_reference(A, steps=cpu_steps, warmup=50, envs_per_worker=args.cpu_envs_per_worker, workload=wl)
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
                                    "sample": f"{cb['envs']} envs x {cpu_steps} steps of the same game, one process per core "
                                              f"({cb['seconds']:.1f} s)"}  # fmt: skip
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--actions", default="effective", choices=["effective", "verbatim"],
                    help="action sampling (SURVEY 8d): the default drives every agent; 'verbatim' is the reference benchmark's")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--agents", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=40000)
    ap.add_argument("--cpu-envs-per-worker", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global ACTION_MODE
    ACTION_MODE = args.actions
    os.environ["METTAGRID_BENCH_ACTIONS"] = args.actions

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        if not reference_available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (make -f oracle/Makefile.ref)"}))
            return
        # one "step" = one tick of the bounded sample (cores x envs_per_worker envs)
        A = args.agents or WORKLOADS[args.workload][2]
        tps = 30 if args.workload == "c2" else 3
        cb = run_reference(A, steps=max(args.steps, 1) * tps, warmup=max(args.warmup, 3) * tps // 3 + 1,
                           envs_per_worker=args.cpu_envs_per_worker, workload=args.workload)  # fmt: skip
        line = {
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32 (f32 stats)", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][0].format(A=A) + f" -- sample of {cb['envs']} envs on "
                                   f"{cb['cores']} host cores (reference C++ step, one process per core)",
                       "ticks_per_step": tps},
            "cpu_baseline": {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
                             "sample": f"{cb['envs']} envs x {max(args.steps, 1) * tps} ticks"},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }  # fmt: skip
        print(json.dumps(line))
        return
    run_ours(args)


if __name__ == "__main__":
    main()
