#!/usr/bin/env python
"""A few steps of one BASELINE configuration and nothing else: the command ncu wraps (profiles/README.md).
    python tools/prof_step.py CONFIG [steps] [n_override]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import __graft_entry__ as G  # noqa: E402

nb = G.load_package()
cfg = dict(bench.CONFIGS[sys.argv[1]])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if len(sys.argv) > 3:
    cfg["n"] = int(sys.argv[3])
block0 = bench.make_block(nb, cfg, product=True)
cov = nb.COVERAGE_FULL if cfg["coverage"] == "full" else nb.COVERAGE_REFERENCE
sim = nb.Simulation(cfg["n"], field_w=cfg["field"], field_h=cfg["field"], coverage=cov, flags=nb.FLAG_NO_GRAPH)
sim.upload(block0, cfg["n"])
for _ in range(steps):
    sim.step(1)
st = sim.stats()
print(f"{sys.argv[1]}: n {cfg['n']} -> {st['n']} after {steps} steps, two-sided {st['pair_halving']}, launches {st['kernel_launches']}")
sim.close()
