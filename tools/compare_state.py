#!/usr/bin/env python
"""Compare two state dumps written by `nbody --dump-state` (int32 n, then the 24 n-byte BodiesData block),
or two event CSVs written by `--dump-events`.

    python tools/compare_state.py a.bin b.bin [--field W]        exit 0 when survivors, masses and radii are
                                                                 bit-identical and positions / velocities are
                                                                 within the parity tolerances of tests/
    python tools/compare_state.py a.csv b.csv                    first differing event, if any
"""
import sys

import numpy as np


def load_state(path):
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw[:4], dtype=np.int32)[0])
    blk = np.frombuffer(raw[4:4 + 24 * n], dtype=np.float32)
    return n, blk[:2 * n].reshape(n, 2), blk[2 * n:4 * n].reshape(n, 2), blk[4 * n:5 * n], blk[5 * n:6 * n]


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    a, b = argv[1], argv[2]
    if a.endswith(".csv"):
        la, lb = open(a).read().splitlines(), open(b).read().splitlines()
        for k, (x, y) in enumerate(zip(la, lb)):
            if x != y:
                print(f"events differ at line {k + 1}: {x!r} vs {y!r}")
                return 1
        if len(la) != len(lb):
            print(f"event counts differ: {len(la) - 1} vs {len(lb) - 1}")
            return 1
        print(f"{len(la) - 1} events identical")
        return 0
    field = float(argv[argv.index("--field") + 1]) if "--field" in argv else None
    na, pa, va, ma, ra = load_state(a)
    nb_, pb, vb, mb, rb = load_state(b)
    print(f"n: {na} vs {nb_}")
    if na != nb_:
        return 1
    ok = True
    for name, x, y in (("mass", ma, mb), ("radius", ra, rb)):
        same = np.array_equal(x.view(np.uint32), y.view(np.uint32))
        print(f"{name}: {'bit-identical' if same else f'{int((x != y).sum())} differ'}")
        ok &= same
    dp, dv = np.abs(pa - pb).max() if na else 0.0, np.abs(va - vb).max() if na else 0.0
    vmax = max(float(np.abs(va).max()) if na else 0.0, 1e-30)
    extent = field if field else max(float(np.abs(pa).max()) if na else 1.0, 1.0)
    print(f"max |dp| = {dp:.6g} ({dp / extent:.3g} of the field half-width)   max |dv| = {dv:.6g} ({dv / vmax:.3g} of max |v|)")
    ok &= dp <= 1e-5 * extent + 1e-2 * vmax and dv <= 1e-2 * vmax
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv))
