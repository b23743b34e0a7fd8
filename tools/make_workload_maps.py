"""Record the map pools of the C3 / C4 / toy bench workloads with the REFERENCE's own builders (build container only).

    python tools/make_workload_maps.py

C3: ``MapGen`` 25x25 + border 6 with a ``Random`` scene, the arena layout of builder/envs.py:54-67 (24 agents in two
teams, 10 walls, plus the chests and altars of the combat game).  C4: ``MapGen`` 64x64 + border 1 with a ``Random`` scene
holding the world game's objects.  The pools are stored as small integer grids (mettagrid_b200/workload_maps/*.npz)
because /root/reference does not exist on the GPU box; mettagrid_b200.workloads.map_pool() reads them back.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import reference  # noqa: E402

assert reference.load() is not None, "needs /root/reference and oracle/_ref"
import mettagrid.mapgen.scenes.random as R  # noqa: E402
from mettagrid.mapgen.mapgen import MapGen  # noqa: E402

from mettagrid_b200 import workloads as W  # noqa: E402

POOL = 64


def record(name, make):
    grids = [np.asarray(make(42 + i)) for i in range(POOL)]
    names = sorted({str(x) for g in grids for x in np.unique(g)})
    idx = {n: i for i, n in enumerate(names)}
    arr = np.stack([np.vectorize(idx.get)(g).astype(np.uint8) for g in grids])
    out = ROOT / "mettagrid_b200" / "workload_maps" / f"{name}.npz"
    np.savez_compressed(out, grids=arr, names=np.array(names))
    print(out, arr.shape, names)


def c3(seed):
    cfg = MapGen.Config(num_agents=24, width=25, height=25, border_width=6, instance_border_width=0, seed=seed,
                        instance=R.Random.Config(agents={"red": 12, "blue": 12}, objects=W.C3_OBJECTS))  # fmt: skip
    return cfg.create().build_for_num_agents(24).grid


def c4(seed):
    cfg = MapGen.Config(num_agents=24, width=64, height=64, border_width=1, instance_border_width=0, seed=seed,
                        instance=R.Random.Config(agents={"red": 12, "blue": 12}, objects=W.C4_OBJECTS))  # fmt: skip
    return cfg.create().build_for_num_agents(24).grid


if __name__ == "__main__":
    record("c3", c3)
    record("c4", c4)
