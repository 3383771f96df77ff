#!/usr/bin/env python
"""Small driver meant for compute-sanitizer (closed on this pool in round 1, so it was only run plain): a few steps at sizes that exercise every code path
(tiny n / exact-only, window edges in reference coverage, part splitting, multi-segment runs, render)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402

nb = G.load_package()
for n, field in [(1, 1000), (130, 2000), (300, 2000), (1000, 4000), (4096, 20000), (20000, 60000)]:
    for cov in (nb.COVERAGE_REFERENCE, nb.COVERAGE_FULL):
        block = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
        sim = nb.Simulation(n, field_w=field, field_h=field, coverage=cov, event_capacity=64 * n + 1024)
        sim.upload(block, n)
        sim.step(3)
        got, n1 = sim.download()
        ev = sim.events()
        img = sim.render(64, 48)
        print(n, cov, n1, len(ev), int((img == 0).sum()), flush=True)
        sim.close()
print("done")
