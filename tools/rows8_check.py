import json, sys
sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as G
from oracle import oracle as O
nb = G.load_package()
F = nb.FLAG_PAIR_HALVING | nb.FLAG_SYM_ROWS8
def parity(n, field, steps=3):
    block0 = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
    sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, event_capacity=64 * n, sort_min_n=1, flags=F)
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
            pg, vg, mg, rg = nb.split(got, n_gpu); pc, vc, mc, rc = O.split(cpu, n_cpu)
            ok &= np.array_equal(mg, mc) and np.array_equal(rg, rc) and np.abs(vg - vc).max() <= 1e-3 * np.abs(vc).max()
    st = sim.stats(); sim.close()
    return bool(ok), st["sym_regs"], st["exact_chunks"]
print(json.dumps({"rows8_parity_5000": parity(5000, 20000)}), flush=True)
print(json.dumps({"rows8_parity_20000": parity(20000, 30000)}), flush=True)
n = 262144; R = 1e5 * np.sqrt(n / 16384.0); field = int(R)
block0 = nb.generate(nb.SCENARIO_DISC, n, extent=R, field_w=field, field_h=field)
for name, fl in (("rows4", 0), ("rows8", F)):
    sim = nb.Simulation(n, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, flags=fl)
    sim.upload(block0, n); sim.step(2)
    s0 = sim.stats(); tot, frc = sim.step_timed(4, force=True); s1 = sim.stats()
    got, n1 = sim.download()
    print(json.dumps({"kernel": name, "n": n, "regs": s1["sym_regs"], "force_ms": frc / 4, "ginter_per_s": (s1["pairs"] - s0["pairs"]) / (frc * 1e-3) / 1e9, "n_after": n1}), flush=True)
    sim.close()
