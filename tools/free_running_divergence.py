#!/usr/bin/env python
"""How fast does a free-running CUDA trajectory leave the oracle's?  Per step: worst |dv| / max|v|, the 99.9 % quantile
and the number of bodies beyond 1e-4 -- the data behind the K in tests/test_parity_gpu.py's free-running bounds.
    python tools/free_running_divergence.py [steps]
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

nb = G.load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
CASES = [("shipped/reference", nb.SCENARIO_SQUARE, 16384, 100000, 0.0, nb.COVERAGE_REFERENCE),
         ("shipped/full", nb.SCENARIO_SQUARE, 16384, 100000, 0.0, nb.COVERAGE_FULL),
         ("disc16k/full", nb.SCENARIO_DISC, 16384, 100000, 1e5, nb.COVERAGE_FULL)]
for name, kind, n0, field, extent, cov in CASES:
    block0 = nb.generate(kind, n0, extent=extent, field_w=field, field_h=field)
    sim = nb.Simulation(n0, field_w=field, field_h=field, coverage=cov)
    sim.upload(block0, n0)
    cpu, n_cpu = block0.copy(), n0
    par = O.params(field_w=field, field_h=field, coverage=cov)
    rows = []
    for s in range(steps):
        sim.step(1)
        n_cpu, _, _ = O.step(cpu, n_cpu, par)
        got, n_gpu = sim.download()
        if n_gpu != n_cpu:
            rows.append({"step": s, "n_gpu": n_gpu, "n_cpu": n_cpu, "diverged": True})
            break
        _, vg, mg, rg = nb.split(got, n_gpu)
        _, vc, mc, rc = O.split(cpu, n_cpu)
        vmax = float(np.abs(vc).max())
        dv = np.abs(vg - vc).max(axis=1) / vmax
        rows.append({"step": s, "n": n_gpu, "worst": float(dv.max()), "q999": float(np.quantile(dv, 0.999)),
                     "beyond_1e-4": int((dv > 1e-4).sum()), "mr_bits": bool(np.array_equal(mg, mc) and np.array_equal(rg, rc))})
    sim.close()
    first = {b: next((r["step"] for r in rows if r.get("worst", 1) > b), None) for b in (1e-4, 1e-3, 1e-2)}
    print(json.dumps({"case": name, "first_step_beyond": {str(k): v for k, v in first.items()},
                      "per_step": [[r["step"], r.get("worst"), r.get("q999"), r.get("beyond_1e-4")] for r in rows]}), flush=True)
