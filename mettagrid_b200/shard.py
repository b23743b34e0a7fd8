"""Multi-GPU plumbing: one process per GPU, environments sharded by index, no per-step collective.

The step path partitions perfectly -- environments share no state and each owns its RNG
(cpp/bindings/mettagrid_c.cpp:52) -- so ranks never exchange data while stepping.  The one collective
is the optional reduction of episode statistics over ranks, the batched analogue of the per-episode
agent-stat averaging in python/src/mettagrid/envs/stats_tracker.py:40-45; it runs over NCCL (NVLink)
on GPUs and over gloo in the CPU tests.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) env range of `rank`; the first `total % world` ranks take one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total_envs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def env_seeds(base_seed: int, begin: int, end: int) -> np.ndarray:
    """Env e (global index) is seeded base_seed + e on whichever rank owns it."""
    return (np.arange(begin, end, dtype=np.int64) + base_seed).astype(np.uint32)


def reduce_agent_stats(values: torch.Tensor, counts: torch.Tensor, group=None) -> torch.Tensor:
    """Mean over all agents of all ranks of a dense per-agent stat table.

    values: [num_local_agents, S] float32 (this rank's agents), counts: scalar tensor = num_local_agents.
    Sums are reduced in float64 so the result does not depend on the rank count."""
    total = values.to(torch.float64).sum(dim=0)
    n = counts.to(torch.float64).reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return (total / n.clamp(min=1)).to(torch.float32)


def gpu_cpu_affinity(device_index: int) -> set[int]:
    """CPUs NVML recommends for a GPU (the cores of its NUMA node); empty when NVML cannot say."""
    try:
        import os

        import pynvml

        pynvml.nvmlInit()
        vis = [int(x) for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip().isdigit()]
        h = pynvml.nvmlDeviceGetHandleByIndex(vis[device_index] if device_index < len(vis) else device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() or 64 + 63) // 64 + 1)
        return {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
    except Exception:
        return set()


def gpu_numa_node(device_index: int) -> int | None:
    """NUMA node of a GPU: from its PCI address in sysfs, else the node that holds the CPUs NVML recommends."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        if node >= 0:
            return node
    except Exception:
        pass
    try:
        import glob

        cpus = gpu_cpu_affinity(device_index)
        if cpus:
            for path in sorted(glob.glob("/sys/devices/system/node/node[0-9]*/cpulist")):
                with open(path) as f:
                    if min(cpus) in _cpulist(f.read()):
                        return int(path.split("/node")[-1].split("/")[0])
    except Exception:
        pass
    return None


def _cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin this process to the cores of its GPU's NUMA node and prefer that node's memory, BEFORE it allocates pinned
    host buffers: the observation rows of a host-resident caller (mg_step_host, 20 MB per step at C2) then cross one
    PCIe root instead of the inter-socket link.  With one rank per GPU and all ranks left on node 0, eight ranks share
    one socket's memory controllers -- the round-1 end-to-end scaling collapse.  Best effort: returns what it did."""
    import ctypes
    import os

    node = gpu_numa_node(device_index)
    info: dict = {"gpu": device_index, "numa_node": node, "cpus": None, "mempolicy": False}
    try:
        cpus = gpu_cpu_affinity(device_index)
        if not cpus and node is not None:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = _cpulist(f.read())
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
    except Exception:
        pass
    if node is None:
        return info
    try:  # set_mempolicy(MPOL_PREFERRED, {node}): x86-64 syscall 238
        mask = ctypes.c_ulong(1 << node)
        rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        info["mempolicy"] = rc == 0
    except Exception:
        pass
    return info
