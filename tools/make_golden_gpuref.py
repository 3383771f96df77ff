#!/usr/bin/env python
"""Generate tests/golden/gpuref_golden.json from the UNMODIFIED reference kernels.

Runs on a B200 box (needs a GPU and the prebuilt oracle/_ref/libnbody_gpuref.so,
which oracle/Makefile compiles from /root/reference/src/nbody.cu in the build
container):

    gpurun -- 'python tools/make_golden_gpuref.py gpurun_out/gpuref_golden.json'

then copy the JSON to tests/golden/.  For every scenario it steps the reference
kernels, records (n, FNV-1a-64 of the whole BodiesData block) after every step,
and -- in the same run -- steps the CPU oracle from the same initial block and
reports whether the two agree BIT FOR BIT (state hash), with diagnostics when
they do not.  The committed JSON is what pins oracle/nbody_oracle.c to the
reference (tests/test_oracle_golden.py replays it on the CPU).
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

SCENARIOS = [
    # name, n, field, steps, dt, growth
    ("shipped", 16384, 100000, 60, 0.2, 0.1),
    ("dense4096", 4096, 20000, 20, 0.2, 0.1),
    ("dense3000", 3000, 12000, 12, 0.5, 0.25),
] + [(f"small{n}", n, 2000, 6, 0.2, 0.1) for n in
     (1, 2, 5, 100, 127, 128, 129, 130, 200, 255, 256, 257, 258, 300, 383, 384, 385, 1000, 1500)]


def main():
    out_path = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "gpuref_golden.json")
    out_path.parent.mkdir(parents=True, exist_ok=True)
    golden = {"generated_by": "tools/make_golden_gpuref.py (unmodified reference kernels on a B200)",
              "scenarios": {}}
    all_ok = True
    for name, n0, field, steps, dt, growth in SCENARIOS:
        block0 = O.init_square(n0, seed=1024, field_w=field, field_h=field)
        par = O.params(dt=dt, growth=growth, field_w=field, field_h=field, coverage=O.COVERAGE_REFERENCE)
        ref = O.GpuRef(block0, n0)
        cpu = block0.copy()
        n_cpu = n0
        trace = []
        ok = True
        t_ref = 0.0
        for s in range(steps):
            n_ref, ms = ref.step(par)
            t_ref += ms
            blk_ref, _ = ref.read()
            h_ref = O.fnv(blk_ref[:6 * n_ref]) if n_ref > 0 else 0
            if n_cpu > 0:
                n_cpu, stats, _ = O.step(cpu, n_cpu, par)
            h_cpu = O.fnv(cpu[:6 * n_cpu]) if n_cpu > 0 else 0
            trace.append({"n": int(n_ref), "fnv": f"{h_ref:016x}"})
            if ok and (n_ref != n_cpu or h_ref != h_cpu):
                ok = False
                all_ok = False
                msg = f"[{name}] step {s}: ref n={n_ref} hash={h_ref:016x}  oracle n={n_cpu} hash={h_cpu:016x}"
                if n_ref == n_cpu and n_ref > 0:
                    pr, vr, mr, rr = O.split(blk_ref, n_ref)
                    pc, vc, mc, rc = O.split(cpu, n_cpu)
                    msg += (f" | pos mismatches {int((pr != pc).any(axis=1).sum())} max|d|={np.abs(pr - pc).max():.3g}"
                            f" vel mismatches {int((vr != vc).any(axis=1).sum())} max|d|={np.abs(vr - vc).max():.3g}"
                            f" mass mismatches {int((mr != mc).sum())} radius mismatches {int((rr != rc).sum())}")
                print(msg, flush=True)
            if n_ref == 0:
                break
        if ref is not None:
            ref.close()
        golden["scenarios"][name] = {"n0": n0, "field": field, "steps": steps, "dt": dt, "growth": growth,
                                     "seed": 1024, "trace": trace, "oracle_bit_exact": ok}
        print(f"[{name}] n0={n0} steps={len(trace)} final n={trace[-1]['n']} oracle_bit_exact={ok} "
              f"ref kernel time {t_ref:.2f} ms", flush=True)
    # final full state of the shipped scenario, compact: survivors' original ordering hash only
    out_path.write_text(json.dumps(golden, indent=1) + "\n")
    print("wrote", out_path, "ALL BIT-EXACT" if all_ok else "MISMATCHES PRESENT")
    return 0 if all_ok else 1


if __name__ == "__main__":
    t = time.time()
    rc = main()
    print(f"{time.time() - t:.1f}s")
    sys.exit(rc)
