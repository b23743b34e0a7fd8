#!/usr/bin/env python
"""Benchmark of the batched MettaGrid step (BASELINE.json metric: agent-steps/s including observations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|toy] [--envs E]

Headline workload (config.workload): BASELINE.json configs[1] ("C2") -- the reference benchmark game
(benchmarks/test_mettagrid_env_benchmark.py:21-29: 20x20 RandomMapBuilder map, 16 agents, 13x13 observation
window, 100 tokens, noop + 4-way move + 152 vibes, max_steps=0) batched to 4096 envs per GPU, env e using map seed
42+e and env seed 42+e, "effective" random actions.

A step = `ticks_per_step` (default 50) consecutive ticks of every env, so that the timed region of K steps lasts long
enough for the clock sampler to see the GPU under load.  `value` is measured with all inputs resident in HBM: one pair
of CUDA events around every mg_step launch, L2 flushed (256 MB fill) before every tick, the flushes outside the event
pairs; ms_per_step is the sum of a step's tick times.  `e2e` is the same ticks through the C ABI's host-buffer entry
point (mg_step_host: pinned host actions in, observations / rewards / flags out, wall clock).

After the headline the default run also times BASELINE configs 3-5 and attaches them as config.extra:
  c3  combat-heavy game, 16384 envs x 24 agents per GPU          c4  world game, 8192 envs x 24 agents per GPU
  c5  the C2 game at 262144 envs TOTAL, split over the N ranks (strong scaling; the headline is weak scaling)
  toy the reference's perf_benchmark.py "toy" preset (40x40 walled map, 20 agents), 4096 envs per GPU
each with its own roofline and, at N=1, its own cpu_baseline.  --no-extras skips them.

`--impl reference` times the UNMODIFIED reference C++ step (oracle/_ref, built from /root/reference by
oracle/Makefile.ref) on the host cores: one worker process per core, each stepping its own share of a bounded sample
of the same workload.  The only places this file touches oracle/ are that arm and the `cpu_baseline` legs.
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

from mettagrid_b200.workloads import WORKLOADS, algo_bytes, gen_actions, make_cfg, make_map  # noqa: E402

METRIC = "agent_steps_per_sec_incl_obs"
UNIT = "agent-steps/s"
DTYPE = "u8/i32 (f32 stats)"
C5_TOTAL_ENVS = 262144

ACTION_MODE = os.environ.get("METTAGRID_BENCH_ACTIONS", "effective")  # --actions (inherited by the CPU workers)


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
        grid = make_map(cfg, agents, workload, env0 + e)
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
    prim, vibe = gen_actions(len(prog.action_names), nprim, 64, envs, agents, 1000 + rank, ACTION_MODE)
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


def cpu_baseline(workload: str, agents: int, cpu_steps: int, envs_per_worker: int):
    steps = cpu_steps if workload in ("c2", "toy") else max(200, cpu_steps // 20)
    cb = run_reference(agents, steps=steps, warmup=50, envs_per_worker=envs_per_worker, workload=workload)
    return {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"],
            "sample": f"{cb['envs']} envs x {steps} ticks of the same game, one process per core ({cb['seconds']:.1f} s)"}  # fmt: skip


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region: NVML (pynvml, ~10 ms period), nvidia-smi as a fallback."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self._stop_evt = index, [], set(), threading.Event()
        self.max_mhz, self.source = None, "nvml"
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()]
            self._h = pynvml.nvmlDeviceGetHandleByIndex(ids[index] if index < len(ids) else index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv, self.source = None, "nvidia-smi"

    def _sample_nvml(self):
        nv = self._nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):  # fmt: skip
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")  # fmt: skip
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for n, v in zip(self.NAMES, out[2:]):
            if v.strip().lower() == "active":
                self.reasons.add(n)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self._nv is not None:  # an NVML call this driver lacks: fall back once
                    self._nv, self.source = None, "nvidia-smi"
            self._stop_evt.wait(0.01 if self._nv is not None else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}  # fmt: skip


def kernel_name(sk: int) -> str:
    """What mg_step launches per tick (include/mettagrid_b200.h: mg_step_kernel)."""
    if sk >= 64:
        return f"k_step_fast<{sk - 64},static>"
    return f"k_step_fast<{sk}>" if sk >= 8 else ("k_world<true>+k_observe+k_finish<true>" if sk == 1 else "k_world<false>+k_observe+k_finish<false>")


def launches_per_tick(sk: int) -> int:
    return 1 if sk >= 8 else 3


def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class Workload:
    """One handle on this rank's GPU plus a pool of device-resident action batches."""

    POOL = 16

    def __init__(self, wl: str, envs: int, agents: int, env0: int, device: int, seed: int, mode: str = None):
        import torch

        from mettagrid_b200.shard import env_seeds
        from mettagrid_b200.sim import BatchedSimulation

        self.wl, self.envs, self.A = wl, envs, agents
        cfg = make_cfg(agents, wl)
        if wl in ("c2", "toy"):
            maps, map_seeds = None, [42 + e for e in range(env0, env0 + envs)]
        else:
            maps, map_seeds = [make_map(cfg, agents, wl, e) for e in range(env0, env0 + envs)], None
        self.sim = BatchedSimulation(cfg, envs, seeds=env_seeds(42, env0, env0 + envs), maps=maps, map_seeds=map_seeds, device=device)
        P = self.P = self.sim.program
        self.num_actions = len(P.action_names)
        self.num_primary = sum(1 for n in P.action_names if not n.startswith("change_vibe_"))
        self.set_actions(seed, mode or ACTION_MODE)
        self.tick = 0
        self.torch = torch

    def set_actions(self, seed: int, mode: str):
        import torch

        self.prim_h, self.vibe_h = gen_actions(self.num_actions, self.num_primary, self.POOL, self.envs, self.A, seed, mode)
        self.prim = torch.from_numpy(self.prim_h).cuda()
        self.vibe = torch.from_numpy(self.vibe_h).cuda()

    def load_actions(self):
        i = self.tick % self.POOL
        self.sim.actions.copy_(self.prim[i])
        self.sim.vibe_actions.copy_(self.vibe[i])
        self.tick += 1

    def run_ticks(self, n: int, flush, timed: bool):
        """n ticks; with `timed`, a CUDA-event pair around each mg_step launch (the L2 flush and the action copy of
        a tick stay outside its pair).  Returns the event pairs."""
        torch = self.torch
        evs = []
        for _ in range(n):
            self.load_actions()  # inputs resident in HBM before the timed region of this tick
            if flush is not None:
                flush.fill_(self.tick & 0xFF)  # evict L2 between timed launches
            if timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                self.sim.step()
                b.record()
                evs.append((a, b))
            else:
                self.sim.step()
        return evs

    def roofline(self, ms_per_tick: float, traffic=None):
        peak, src = peak_hbm()
        ab = algo_bytes(self.P, self.wl)
        achieved = ab * self.envs * self.A / (ms_per_tick * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name(self.sim.step_kernel), "peak_source": src,
                "algorithmic_bytes_per_agent_step": ab, "agent_steps_per_launch": self.envs * self.A}  # fmt: skip

    def close(self):
        self.sim.close()
        self.torch.cuda.empty_cache()


def measured_traffic(kernel: str, wl: str, envs: int):
    """DRAM bytes per launch from the committed ncu capture of this kernel at this shape, else None."""
    tpath = ROOT / "profiles" / "traffic.json"
    if not tpath.exists():
        return None
    rec = json.loads(tpath.read_text()).get(f"{wl}:{kernel}", {})
    return rec.get("dram_bytes_per_launch") if rec.get("envs") == envs else None


def time_extra(wl: str, envs: int, agents: int, env0: int, world: int, rank: int, local_rank: int, ticks: int, flush, args,
               scaling: str):  # fmt: skip
    """One of the BASELINE configs 3-5: a few warm-up ticks, `ticks` timed ticks, max over ranks."""
    import torch
    import torch.distributed as dist

    w = Workload(wl, envs, agents, env0, local_rank, 7 + rank)
    w.run_ticks(3, None, False)
    torch.cuda.synchronize()
    w.sim.check_errors()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = w.run_ticks(ticks, flush, True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    w.sim.check_errors()
    per_tick = ms / ticks
    out = {
        "workload": WORKLOADS[wl][0].format(A=agents) + f", {envs} envs/GPU", "value": world * envs * agents * ticks / (ms * 1e-3),
        "unit": UNIT, "ticks": ticks, "ms_per_tick": per_tick, "envs_per_gpu": envs, "total_envs": world * envs, "scaling": scaling,
        "roofline": w.roofline(per_tick, measured_traffic(kernel_name(w.sim.step_kernel), wl, envs)),
        "gpu_launches": ticks * launches_per_tick(w.sim.step_kernel),
    }  # fmt: skip
    w.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    from mettagrid_b200.shard import bind_to_gpu_numa, reduce_agent_stats, shard_range

    numa = bind_to_gpu_numa(local_rank)  # before any pinned allocation: host staging lands next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = args.workload
    envs = args.envs or WORKLOADS[wl][1]
    A = args.agents or WORKLOADS[wl][2]
    env0, _ = shard_range(world * envs, rank, world)  # weak scaling: `envs` environments per GPU
    w = Workload(wl, envs, A, env0, local_rank, 7 + rank)
    sim, P = w.sim, w.P
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    steps, warmup, tps = args.steps, max(args.warmup, 3), args.ticks_per_step

    w.run_ticks(warmup * tps, None, False)
    torch.cuda.synchronize()
    sim.check_errors()

    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = w.run_ticks(steps * tps, flush, True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = steps * tps * launches_per_tick(sim.step_kernel)  # step-kernel launches in the timed region
    clocks = sampler.stop()

    # ---- e2e: host buffers through the C ABI (pinned), copies inside the timed region
    NA = envs * A
    h_prim = torch.from_numpy(w.prim_h[0].copy()).pin_memory()
    h_vibe = torch.from_numpy(w.vibe_h[0].copy()).pin_memory()
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
    e2e_ticks = steps * min(tps, args.e2e_ticks_per_step)
    t0 = time.perf_counter()
    for i in range(e2e_ticks):
        h_prim.copy_(torch.from_numpy(w.prim_h[i % w.POOL]))  # next tick's host inputs (host memcpy, part of e2e)
        h_vibe.copy_(torch.from_numpy(w.vibe_h[i % w.POOL]))
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
    value = world * envs * A * steps * tps / (total_ms * 1e-3)
    e2e_value = world * envs * A * e2e_ticks / (e2e_ms * 1e-3)
    per_tick_ms = total_ms / (steps * tps)
    roof = w.roofline(per_tick_ms, measured_traffic(kernel_name(sim.step_kernel), wl, envs))

    # ---- the reference benchmark's verbatim action sampling on the same handle (report both)
    other = "verbatim" if ACTION_MODE == "effective" else "effective"
    w.set_actions(107 + rank, other)
    w.run_ticks(3, None, False)
    torch.cuda.synchronize()
    evs = w.run_ticks(40, flush, True)
    torch.cuda.synchronize()
    tv = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    other_line = {"value": world * envs * A * 40 / (float(tv[0]) * 1e-3), "ms_per_tick": float(tv[0]) / 40, "ticks": 40}
    sim.check_errors()
    w.close()

    extra = {}
    if not args.no_extras and wl == "c2" and not args.envs:
        for name in ("c3", "c4", "toy"):
            n, a = WORKLOADS[name][1], WORKLOADS[name][2]
            extra[name] = time_extra(name, n, a, rank * n, world, rank, local_rank, args.extra_ticks, flush, args, "weak")
        n5 = C5_TOTAL_ENVS // world
        extra["c5"] = time_extra("c2", n5, 16, rank * n5, world, rank, local_rank, args.extra_ticks, flush, args, "strong")
        extra["c5"]["workload"] = f"C5: the C2 game at {C5_TOTAL_ENVS} envs in total, {n5} per GPU"

    if rank == 0:
        num_actions, num_primary = w.num_actions, w.num_primary
        desc = {"effective": f"primary uniform over {num_primary}, vibe p=0.1",
                "verbatim": f"verbatim: primary uniform over all {num_actions} ids, vibe buffer 0"}  # fmt: skip
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": WORKLOADS[wl][0].format(A=A) + f", {envs} envs/GPU",
                       "envs_per_gpu": envs, "agents_per_env": A, "ticks_per_step": tps, "ms_per_tick": per_tick_ms,
                       "actions": desc[ACTION_MODE], "other_action_sampling": {"actions": desc[other], **other_line},
                       "l2": "flushed before every timed tick (256 MB fill, outside the event pairs)",
                       "obs_write_GBps": value * 3 * P.num_tokens / 1e9,
                       "parallelism": f"env-sharded x{world}, no per-step collective", "numa_binding": numa,
                       "mean_move_success_per_agent": float(mean_stats[P.agent_stat_names.index("action.move.success")]),
                       "extra": extra},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * NA * 4 * (e2e_ticks // steps),
                    "d2h_bytes_per_step": NA * (3 * P.num_tokens + 4 + 1 + 1) * (e2e_ticks // steps),
                    "ticks_per_step": e2e_ticks // steps, "check": e2e_sum},
            "gpu_launches": launches,
            "roofline": roof,
        }  # fmt: skip
        if world == 1 and not args.no_cpu_baseline and reference_available():
            line["cpu_baseline"] = cpu_baseline(wl, A, args.cpu_steps, args.cpu_envs_per_worker)
            for name, x in extra.items():
                if name == "c5":
                    x["cpu_baseline"] = dict(line["cpu_baseline"], sample=line["cpu_baseline"]["sample"] + " (the C2 game)")
                else:
                    x["cpu_baseline"] = cpu_baseline(name, WORKLOADS[name][2], args.cpu_steps, args.cpu_envs_per_worker)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ticks-per-step", type=int, default=50, help="ticks of every env per timed step")
    ap.add_argument("--e2e-ticks-per-step", type=int, default=10, help="ticks per step of the host-buffer (e2e) leg")
    ap.add_argument("--extra-ticks", type=int, default=30, help="timed ticks of each config.extra workload")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--actions", default="effective", choices=["effective", "verbatim"],
                    help="action sampling (SURVEY 8d): the default drives every agent; 'verbatim' is the reference benchmark's")
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--agents", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=40000)
    ap.add_argument("--cpu-envs-per-worker", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
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
        # one "step" = ticks_per_step ticks of the bounded sample (cores x envs_per_worker envs)
        A = args.agents or WORKLOADS[args.workload][2]
        tps = args.ticks_per_step if args.workload in ("c2", "toy") else max(2, args.ticks_per_step // 10)
        cb = run_reference(A, steps=max(args.steps, 1) * tps, warmup=max(args.warmup, 3) * tps // 3 + 1,
                           envs_per_worker=args.cpu_envs_per_worker, workload=args.workload)  # fmt: skip
        line = {
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"] * tps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
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
