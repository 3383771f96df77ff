#!/usr/bin/env python
"""Two-sided force kernel on `world` GPUs (launch with torchrun) against the same run on one GPU: replicas identical,
events / survivors / masses / radii identical to the single-GPU run, velocities within 1e-4 max|v|."""
import json
import os
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import __graft_entry__ as G

nb = G.load_package()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
R = 1e5 * np.sqrt(n / 16384.0)
field = int(R)
block0 = nb.generate(nb.SCENARIO_DISC, n, extent=R, field_w=field, field_h=field)
ids = [nb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, device=local, rank=rank, world=world, event_capacity=1 << 20)
sim.comm_init(ids[0])
sim.upload(block0, n)
two_sided = sim.stats()["pair_halving"]
sim.step(steps)
got, n1 = sim.download()
ev = sim.events()
st = sim.stats()
sim.close()
digest = [(n1, zlib.crc32(got.tobytes()), len(ev), st["overflow"])]
all_d = [None] * world
dist.all_gather_object(all_d, digest[0])
all_ev = [None] * world
dist.all_gather_object(all_ev, ev)
if rank == 0:
    one = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, device=local, event_capacity=1 << 20)
    one.upload(block0, n)
    one.step(steps)
    ref, n_ref = one.download()
    ev_ref = one.events()
    one.close()
    evs = np.concatenate(all_ev)
    evs = evs[np.lexsort((evs["i"], evs["step"]))]      # stable: a row's events stay in visit order
    ok_ev = len(evs) == len(ev_ref) and all(np.array_equal(evs[k], ev_ref[k]) for k in ("step", "i", "j", "kind"))
    res = {"world": world, "n": n, "steps": steps, "two_sided": two_sided, "replicas_identical": len({(d[0], d[1]) for d in all_d}) == 1,
           "n_after": [d[0] for d in all_d], "n_ref": n_ref, "events": len(evs), "events_equal": bool(ok_ev), "overflow": [d[3] for d in all_d]}
    if n1 == n_ref:
        p1, v1, m1, r1 = nb.split(got, n1)
        p2, v2, m2, r2 = nb.split(ref, n_ref)
        res.update({"mass_radius_equal": bool(np.array_equal(m1, m2) and np.array_equal(r1, r2)),
                    "dv_rel_max": float(np.abs(v1 - v2).max() / np.abs(v2).max()), "dp_max": float(np.abs(p1 - p2).max())})
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
