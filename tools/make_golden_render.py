#!/usr/bin/env python
"""Generate tests/golden/render_golden.json from the UNMODIFIED reference generateImage kernel
(src/nbody.cu:294-348), launched as the reference's main loop launches it (oracle/gpu_ref_harness.cu,
gpuref_render).  Needs a GPU and the prebuilt oracle/_ref/libnbody_gpuref.so:

    gpurun -- 'python tools/make_golden_render.py gpurun_out/render_golden.json'

then copy the JSON to tests/golden/.  Every scenario is drawn from bodies the reference's own init sequence
generates (seed 1024); `steps` > 0 first runs that many reference steps (only scenarios in which no body dies, so
that the loop's stale launch grid stays inside the body store).  Recorded: FNV-1a-64 of the image bytes and the
number of body pixels.  In the same run the CPU oracle's orc_render is compared byte for byte.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

SCENARIOS = [
    # name, n, field, min_radius, max_radius, width, height, steps
    ("r130", 130, 2000, 20.0, 100.0, 64, 48, 0),                 # 1 block: bodies 128, 129 are not drawn
    ("r1000", 1000, 20000, 200.0, 900.0, 256, 256, 0),           # 7 blocks: 896 of 1000 drawn
    ("r1024", 1024, 20000, 200.0, 900.0, 300, 200, 0),           # all drawn, non-square image
    ("r4096", 4096, 100000, 500.0, 3000.0, 512, 512, 0),         # big discs, many clipped at the borders
    ("tall", 2048, 50000, 300.0, 2500.0, 64, 512, 0),
    ("shipped", 16384, 100000, 50.0, 200.0, 1024, 1024, 0),      # nbodyConfig.txt's own picture of the initial bodies
    ("sparse_stepped", 512, 200000, 100.0, 400.0, 2048, 2048, 3), # 3 reference steps first; nobody dies
    ("sparse_stepped_tail", 640, 300000, 100.0, 500.0, 2048, 1024, 3),
]


def main():
    out_path = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "render_golden.json")
    out_path.parent.mkdir(parents=True, exist_ok=True)
    golden = {"generated_by": "tools/make_golden_render.py (unmodified reference generateImage on a B200)", "scenarios": {}}
    all_ok = True
    for name, n, field, rmin, rmax, w, h, steps in SCENARIOS:
        block = O.init_square(n, field_w=field, field_h=field, min_radius=rmin, max_radius=rmax)
        par = O.params(field_w=field, field_h=field, coverage=O.COVERAGE_REFERENCE)
        ref = O.GpuRef(block, n)
        cpu, n_cpu = block.copy(), n
        grid_n = n
        for _ in range(steps):
            grid_n = n_cpu
            n_ref, _ = ref.step(par)
            n_cpu, _, _ = O.step(cpu, n_cpu, par)
            assert n_ref == n_cpu == n, f"{name}: a body died; choose a sparser scenario"
        img = ref.render(w, h, field, field, grid_n)
        ref.close()
        drawn = 128 * max(1, grid_n // 128)
        mine = O.render(cpu, n_cpu, w, h, field, field, grid_n=grid_n)
        ok = bool(np.array_equal(img, mine))
        all_ok &= ok
        golden["scenarios"][name] = {"n": n, "field": field, "min_radius": rmin, "max_radius": rmax, "width": w, "height": h,
                                     "steps": steps, "grid_n": grid_n, "drawn": min(n, drawn), "fnv": f"{O.fnv(img):016x}",
                                     "body_pixels": int((img == 0).sum()), "oracle_byte_exact": ok}
        print(f"[{name}] {w}x{h} drawn {min(n, drawn)}/{n} body pixels {int((img == 0).sum())} oracle_byte_exact={ok}", flush=True)
        if not ok:
            print(f"    differing pixels: {int((img != mine).sum())}", flush=True)
    out_path.write_text(json.dumps(golden, indent=1) + "\n")
    print("wrote", out_path, "ALL BYTE-EXACT" if all_ok else "MISMATCHES PRESENT")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
