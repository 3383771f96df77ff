#!/usr/bin/env python
"""Force-kernel time of the default path at the given sizes (library picked by NBODY_B200_LIB): kernel experiments."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

nb = G.load_package()
flags = int(os.environ.get("NB_FLAGS", "0"))
for n in [int(a) for a in sys.argv[1:]] or [131072, 1048576]:
    R = 1e5 * np.sqrt(n / 16384.0)
    field = int(R)
    block0 = nb.generate(nb.SCENARIO_DISC, n, extent=R, field_w=field, field_h=field)
    sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, flags=flags)
    sim.upload(block0, n)
    sim.step(3)
    steps = 5
    s0 = sim.stats()
    tot, frc = sim.step_timed(steps, force=True)
    s1 = sim.stats()
    pairs = s1["pairs"] - s0["pairs"]
    got, n1 = sim.download()
    print(json.dumps({"lib": os.environ.get("NBODY_B200_LIB", "default"), "n": n, "pair_halving": s1["pair_halving"], "regs": s1["sym_regs"],
                      "force_ms": frc / steps, "step_ms": tot / steps, "ginter_per_s_step": pairs / (tot * 1e-3) / 1e9,
                      "n_after": n1, "checksum": float(np.abs(got).sum(dtype=np.float64))}), flush=True)
    sim.close()
