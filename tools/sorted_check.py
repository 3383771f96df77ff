#!/usr/bin/env python
"""Sorted-j-stream path: parity against the oracle at small n (forced on) and speed against the plain path."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

nb = G.load_package()


def parity(n, field, steps=4):
    block0 = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
    sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, event_capacity=64 * n,
                        sort_min_n=1)
    sim.upload(block0, n)
    cpu, n_cpu = block0.copy(), n
    par = O.params(field_w=field, field_h=field, coverage=O.COVERAGE_FULL)
    ok = True
    for s in range(steps):
        sim.step(1)
        n_cpu, _, ev_cpu = O.step(cpu, n_cpu, par, want_events=True)
        got, n_gpu = sim.download()
        ev = sim.events()
        ok &= n_gpu == n_cpu and len(ev) == len(ev_cpu) and np.array_equal(ev["i"], ev_cpu["i"]) and np.array_equal(ev["j"], ev_cpu["j"])
        if ok:
            pg, vg, mg, rg = nb.split(got, n_gpu)
            pc, vc, mc, rc = O.split(cpu, n_cpu)
            ok &= np.array_equal(mg, mc) and np.array_equal(rg, rc) and np.abs(pg - pc).max() <= 1e-5 * field
            ok &= np.abs(vg - vc).max() <= 1e-3 * np.abs(vc).max()
    st = sim.stats()
    sim.close()
    return bool(ok), st["culled_parts"], st["fast_chunks"]


for n, field in [(300, 2000), (1000, 4000), (5000, 20000), (20000, 60000)]:
    print(json.dumps({"parity_n": n, "result": parity(n, field)}), flush=True)

for n in [int(a) for a in sys.argv[1:]] or [131072, 1048576]:
    R = 1e5 * np.sqrt(n / 16384.0)
    field = int(R)
    block0 = nb.generate(nb.SCENARIO_DISC, n, extent=R, field_w=field, field_h=field)
    for flags in (nb.FLAG_NO_SORT, 0):
        sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, flags=flags)
        sim.upload(block0, n)
        sim.step(3)
        sim.sync()
        steps = 5
        s0 = sim.stats()
        tot, frc = sim.step_timed(steps, force=True)
        s1 = sim.stats()
        prof = sim.step_profile(2)
        pairs = s1["pairs"] - s0["pairs"]
        print(json.dumps({"n": n, "sorted": not flags, "force_ms": frc / steps, "step_ms": tot / steps,
                          "frac_roofline_force": pairs * 20 / (frc * 1e-3) / 74.45e12,
                          "frac_roofline_step": pairs * 20 / (tot * 1e-3) / 74.45e12,
                          "culled_parts": s1["culled_parts"] - s0["culled_parts"], "exact": s1["exact_chunks"] - s0["exact_chunks"],
                          "compact_ms": prof["compact"] / 2, "sort_ms": prof["sort"] / 2}), flush=True)
        sim.close()
