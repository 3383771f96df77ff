#!/usr/bin/env python
"""Two-sided (pair-halving) force kernel: parity against the oracle at small n (sorted order forced on), agreement
with the one-sided kernel and run-to-run determinism at large n, and speed against the one-sided kernel."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402
from oracle import oracle as O  # noqa: E402

nb = G.load_package()
EXTRA = int(__import__("os").environ.get("NB_EXTRA_FLAGS", "0"))      # e.g. 128 = NB_FLAG_SYM_ROWS8 (experimental variant)


def parity(n, field, steps=4, softening=0.0):
    block0 = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
    sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, event_capacity=64 * n,
                        sort_min_n=1, flags=nb.FLAG_PAIR_HALVING | EXTRA, softening=softening)
    sim.upload(block0, n)
    cpu, n_cpu = block0.copy(), n
    par = O.params(field_w=field, field_h=field, coverage=O.COVERAGE_FULL, softening=softening)
    ok, why = True, ""
    used = 0
    for s in range(steps):
        used += sim.stats()["pair_halving"]
        sim.step(1)
        n_cpu, _, ev_cpu = O.step(cpu, n_cpu, par, want_events=True)
        got, n_gpu = sim.download()
        ev = sim.events()
        if not (n_gpu == n_cpu and len(ev) == len(ev_cpu) and np.array_equal(ev["i"], ev_cpu["i"]) and np.array_equal(ev["j"], ev_cpu["j"])):
            ok, why = False, f"step {s}: n {n_gpu} vs {n_cpu}, events {len(ev)} vs {len(ev_cpu)}"
            break
        pg, vg, mg, rg = nb.split(got, n_gpu)
        pc, vc, mc, rc = O.split(cpu, n_cpu)
        dv = np.abs(vg - vc).max() / np.abs(vc).max()
        if not (np.array_equal(mg, mc) and np.array_equal(rg, rc) and np.abs(pg - pc).max() <= 1e-5 * field and dv <= 1e-3):
            ok, why = False, f"step {s}: m/r exact {np.array_equal(mg, mc)} {np.array_equal(rg, rc)}, dp {np.abs(pg - pc).max():.3g}, dv {dv:.3g}"
            break
        why = f"dv {dv:.2e}"
    st = sim.stats()
    sim.close()
    return {"ok": bool(ok), "why": why, "steps_two_sided": used, "exact": st["exact_chunks"], "culled": st["culled_parts"],
            "events": st["candidates"], "n_end": st["n"]}


for n, field, soft in [(1100, 4000, 0.0), (3000, 9000, 0.0), (5000, 20000, 0.0), (20000, 60000, 0.0), (20000, 30000, 0.0), (5000, 20000, 500.0)]:
    print(json.dumps({"parity_n": n, "field": field, "softening": soft, "result": parity(n, field, softening=soft)}), flush=True)

for n in [int(a) for a in sys.argv[1:]] or [131072, 1048576]:
    R = 1e5 * np.sqrt(n / 16384.0)
    field = int(R)
    block0 = nb.generate(nb.SCENARIO_DISC, n, extent=R, field_w=field, field_h=field)
    outs = {}
    for name, flags in (("one_sided", nb.FLAG_ONE_SIDED), ("two_sided", nb.FLAG_PAIR_HALVING | EXTRA), ("two_sided_again", nb.FLAG_PAIR_HALVING | EXTRA)):
        sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, flags=flags, event_capacity=1 << 20)
        sim.upload(block0, n)
        sim.step(3)
        outs[name] = (sim.download(), sim.events())
        if name != "two_sided_again":
            steps = 5
            s0 = sim.stats()
            tot, frc = sim.step_timed(steps, force=True)
            s1 = sim.stats()
            prof = sim.step_profile(2)
            pairs = s1["pairs"] - s0["pairs"]
            print(json.dumps({"n": n, "kernel": name, "pair_halving": s1["pair_halving"], "regs": s1["sym_regs"], "force_ms": frc / steps,
                              "step_ms": tot / steps, "ginter_per_s_step": pairs / (tot * 1e-3) / 1e9,
                              "frac_roofline_force": pairs * 20 / (frc * 1e-3) / 74.45e12,
                              "culled": s1["culled_parts"] - s0["culled_parts"], "exact": s1["exact_chunks"] - s0["exact_chunks"],
                              "finish_ms": prof["finish"] / 2, "compact_ms": prof["compact"] / 2, "sort_ms": prof["sort"] / 2}), flush=True)
        sim.close()
    (b1, n1), e1 = outs["one_sided"]
    (b2, n2), e2 = outs["two_sided"]
    (b3, n3), e3 = outs["two_sided_again"]
    same_events = len(e1) == len(e2) and np.array_equal(e1["i"], e2["i"]) and np.array_equal(e1["j"], e2["j"]) and np.array_equal(e1["kind"], e2["kind"])
    res = {"n": n, "survivors": [n1, n2, n3], "events_equal": bool(same_events), "deterministic": bool(n2 == n3 and np.array_equal(b2, b3))}
    if n1 == n2:
        p1, v1, m1, r1 = nb.split(b1, n1)
        p2, v2, m2, r2 = nb.split(b2, n2)
        res.update({"mass_radius_equal": bool(np.array_equal(m1, m2) and np.array_equal(r1, r2)),
                    "dv_rel_max": float(np.abs(v1 - v2).max() / np.abs(v1).max()), "dp_max": float(np.abs(p1 - p2).max())})
    print(json.dumps(res), flush=True)
