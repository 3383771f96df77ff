#!/usr/bin/env python
"""Time every force-kernel variant on one GPU and check each against the oracle (3 steps, shipped scenario).

    gpurun -- 'python tools/variant_sweep.py [n ...] > gpurun_out/variants.jsonl'
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
sizes = [int(a) for a in sys.argv[1:]] or [16384, 131072]
NVAR = 6


def check(variant):
    n0 = 16384
    block0 = nb.generate(nb.SCENARIO_SQUARE, n0)
    sim = nb.Simulation(n0, coverage=nb.COVERAGE_REFERENCE, event_capacity=1 << 16, flags=nb.flag_variant(variant))
    sim.upload(block0, n0)
    cpu = block0.copy()
    n_cpu = n0
    par = O.params(coverage=O.COVERAGE_REFERENCE)
    ok = True
    for s in range(3):
        sim.step(1)
        n_cpu, _, ev_cpu = O.step(cpu, n_cpu, par, want_events=True)
        got, n_gpu = sim.download()
        ev = sim.events()
        ok &= n_gpu == n_cpu and len(ev) == len(ev_cpu) and np.array_equal(ev["j"], ev_cpu["j"])
        if ok:
            _, _, mg, rg = nb.split(got, n_gpu)
            _, _, mc, rc = O.split(cpu, n_cpu)
            ok &= np.array_equal(mg, mc) and np.array_equal(rg, rc)
    sim.close()
    return bool(ok)


for variant in range(NVAR):
    ok = check(variant)
    for n in sizes:
        R = 1e5 * np.sqrt(n / 16384.0)
        field = int(R)
        block0 = nb.generate(nb.SCENARIO_DISC, n, extent=R, field_w=field, field_h=field)
        sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, flags=nb.flag_variant(variant))
        sim.upload(block0, n)
        sim.step(3)
        sim.sync()
        steps = 20 if n <= 32768 else 5
        s0 = sim.stats()
        tot, frc = sim.step_timed(steps, force=True)
        s1 = sim.stats()
        sim.upload(block0, n)
        sim.step(3)
        sim.sync()
        tot_g, _ = sim.step_timed(steps, force=False)      # CUDA-graph path, whole step
        pairs = s1["pairs"] - s0["pairs"]
        print(json.dumps({"variant": variant, "parity_ok": ok, "n": n, "regs": s1["force_regs"], "threads": s1["force_threads"],
                          "grid": s1["force_grid"], "force_ms": frc / steps, "step_ms": tot / steps, "graph_step_ms": tot_g / steps,
                          "ginter_per_s_force": pairs / (frc * 1e-3) / 1e9,
                          "frac_roofline_force": pairs * 20 / (frc * 1e-3) / 74.45e12}), flush=True)
        sim.close()
