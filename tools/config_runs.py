#!/usr/bin/env python
"""Run the BASELINE.json configs that bench.py does not cover and print one JSON line each.

    python tools/config_runs.py shipped|disc16k|cluster|galaxy [steps]      (galaxy: launch with torchrun on 8 GPUs)
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

nb = G.load_package()
which = sys.argv[1]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
CFG = {
    # name: (n, scenario, extent, field, coverage, default steps)
    "shipped": (16384, nb.SCENARIO_SQUARE, 0.0, 100000, nb.COVERAGE_REFERENCE, 2000),   # configs[0]: nbodyConfig.txt as shipped
    "disc16k": (16384, nb.SCENARIO_DISC, 1e5, 100000, nb.COVERAGE_FULL, 1000),           # configs[1]
    "cluster": (131072, nb.SCENARIO_DISC, 1e5, 200000, nb.COVERAGE_FULL, 200),           # configs[2]
    "galaxy": (4194304, nb.SCENARIO_TWO_GALAXY, 8e5, 3000000, nb.COVERAGE_FULL, 5),      # configs[4]
}
n, kind, extent, field, coverage, steps = CFG[which]
if len(sys.argv) > 2:
    steps = int(sys.argv[2])
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
block0 = nb.generate(kind, n, extent=extent, field_w=field, field_h=field)
sim = nb.Simulation(n, field_w=field, field_h=field, coverage=coverage, device=local, rank=rank, world=world)
if world > 1:
    ids = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sim.comm_init(ids[0])
sim.upload(block0, n)
first = []
for _ in range(min(3, steps)):                      # the first steps separately: they carry most of the merging
    t, _f = sim.step_timed(1, force=False)
    first.append(round(t, 4))
s0 = sim.stats()
rest = steps - len(first)
t0 = time.perf_counter()
ms, _ = sim.step_timed(rest, force=False) if rest > 0 else (0.0, 0.0)
wall = time.perf_counter() - t0
s1 = sim.stats()
if world > 1:
    import torch
    tt = torch.tensor([ms, float(s1["pairs"] - s0["pairs"]), float(s1["candidates"])], dtype=torch.float64)
    mx = tt.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = tt.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ms, pairs, cands = float(mx[0]), float(sm[1]), float(sm[2])
else:
    pairs, cands = float(s1["pairs"] - s0["pairs"]), float(s1["candidates"])
if rank == 0:
    print(json.dumps({"config": which, "n0": n, "n_end": s1["n"], "steps": steps, "n_gpus": world, "first_steps_ms": first,
                      "ms_rest": round(ms, 3), "steps_per_sec": rest / (ms * 1e-3) if ms > 0 else None,
                      "interactions_per_sec": pairs / (ms * 1e-3) if ms > 0 else None, "collision_events": cands,
                      "overflow": s1["overflow"], "wall_s_rest": round(wall, 3)}), flush=True)
sim.close()
if world > 1:
    dist.destroy_process_group()
