"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/nbody_b200.h
declares, the host-only driver surface (config parser, RNG, initial conditions, plan, P5 writer) matches the
golden vectors produced from the reference's own headers, and the product has no route to the oracle or to
a CPU fallback.  No compute call needs a GPU here."""
import ctypes as C
import json
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def host_golden(golden_dir):
    return json.loads((golden_dir / "host_golden.json").read_text())


def test_library_exports_every_declared_symbol(nb):
    header = (ROOT / "include" / "nbody_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    lib = nb.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert sorted(nb.SYMBOLS) == declared, "the Python mirror's symbol list is out of date"
    assert lib.nb_version() == 103


def test_no_cpu_fallback(nb):
    """Without a GPU nb_create must fail loudly (this container has none; on a GPU box the test is moot)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(nb.NbodyError) as e:
        nb.Simulation(128)
    assert e.value.code == nb.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = ROOT / "ppa-nbody-collisions_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cuh")) + \
            list(pkg.rglob("Makefile")) + [ROOT / "include" / "nbody_b200.h"]:
        text = path.read_text()
        assert "oracle" not in text.lower().replace("the oracle's", ""), f"{path} mentions the oracle"
    deps = subprocess.run(["ldd", str(pkg / "lib" / "libnbody_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps and "gpuref" not in deps


def test_rng_matches_reference_headers(nb, host_golden):
    g = nb.Rng()
    nb.lib().nb_rng_seed(C.byref(g), 1024)
    assert [str(nb.lib().nb_rng_ival64(C.byref(g))) for _ in range(8)] == host_golden["rng"]["seed1024_ival64"]
    nb.lib().nb_rng_seed(C.byref(g), 1024)
    assert [nb.lib().nb_rng_fval(C.byref(g)) for _ in range(8)] == [float(v) for v in host_golden["rng"]["seed1024_fval"]]


def test_initial_conditions_match_reference(nb, oracle, host_golden):
    g = host_golden["init_shipped"]
    block = nb.generate(nb.SCENARIO_SQUARE, g["n"])
    assert f"{oracle.fnv(block):016x}" == g["fnv1a64_block"]
    # other sizes / fields against the oracle's restatement of src/nbody.cu:401-416
    for n, field in [(1, 2000), (300, 2000), (4096, 20000)]:
        assert np.array_equal(nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field),
                              oracle.init_square(n, field_w=field, field_h=field))


def test_synthetic_scenarios_are_well_formed(nb):
    n = 4096
    disc = nb.generate(nb.SCENARIO_DISC, n, extent=1e5)
    pos, vel, m, r = nb.split(disc, n)
    assert (np.hypot(pos[:, 0], pos[:, 1]) <= 1e5 * (1 + 1e-6)).all() and not vel.any()
    assert m.min() >= 1e4 and m.max() <= 1e17 and r.min() >= 50 and r.max() <= 200
    gal = nb.generate(nb.SCENARIO_TWO_GALAXY, n, extent=8e5, field_w=3000000, field_h=3000000)
    pos, vel, m, r = nb.split(gal, n)
    assert (pos[: n // 2, 0] < 0).all() and (pos[n // 2:, 0] > 0).all()
    assert vel[: n // 2, 0].mean() > 300 and vel[n // 2:, 0].mean() < -300
    assert np.array_equal(disc, nb.generate(nb.SCENARIO_DISC, n, extent=1e5)), "generators are deterministic"


def _parse(nb, tmp_path, text):
    cfg_path = tmp_path / "nbodyConfig.txt"
    cfg_path.write_text(text)
    echo_path = tmp_path / "echo.txt"
    fd = os.open(echo_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    try:
        rc, cfg = nb.parse_config(cfg_path, fd)
    finally:
        os.close(fd)
    return rc, cfg, echo_path.read_text()


def test_config_parser_shipped(nb, tmp_path, host_golden):
    g = host_golden["config_shipped"]
    text = Path("/root/reference/nbodyConfig.txt").read_text() if Path("/root/reference/nbodyConfig.txt").exists() else None
    if text is None:
        text = ("particleCount=16384\ntotalIterations=2000\nsave_Image_Every_Xth_Iteration=10\ntimestep=0.2f\n"
                "radiusGrowthRate=0.1f\nminRandBodyMass=1e4f\nmaxRandBodyMass=1e17f\nminRadius=50.f\nmaxRadius=200.f\n"
                "imgWidth=1024\nimgHeight=1024\nfieldWidth=100000\nfieldHeight=100000\nimagePath=iter_img\n")
    rc, cfg, echo = _parse(nb, tmp_path, text)
    assert rc == nb.OK and echo == g["echo"]
    for key, want in g["values"].items():
        got = getattr(cfg, key)
        assert (f"{got:.9g}" if isinstance(got, float) else str(got)) == want, key
    assert cfg.imagePath.decode() == g["imagePath"]


def test_config_parser_quirks(nb, tmp_path, host_golden):
    g = host_golden["config_quirky"]
    rc, cfg, echo = _parse(nb, tmp_path, g["text"])
    assert rc == nb.OK and echo == g["echo"]
    assert cfg.imagePath.decode() == g["imagePath"]
    for key, want in g["values"].items():
        got = getattr(cfg, key)
        assert (f"{got:.9g}" if isinstance(got, float) else str(got)) == want, key


def test_config_parser_errors(nb, tmp_path):
    rc, _, echo = _parse(nb, tmp_path, "particleCount=abc\n")
    assert rc == nb.ERR_INVALID and echo == "particleCount invalid value: stoi\n"       # nbodyConfig.h:41-44
    rc, _, echo = _parse(nb, tmp_path, "radiusGrowthRate=\n")
    assert rc == nb.ERR_INVALID and echo == "growthRate invalid value: stof\n"          # nbodyConfig.h:213-216
    rc, cfg = nb.parse_config(tmp_path / "missing.txt", -1)
    assert rc == nb.ERR_IO


def test_pgm_writer(nb, tmp_path):
    img = (np.arange(6 * 4, dtype=np.uint8).reshape(4, 6) * 7)
    path = tmp_path / "iteration_0.ppm"
    assert nb.lib().nb_write_pgm(str(path).encode(), img.ctypes.data, 6, 4) == nb.OK
    assert path.read_bytes() == b"P5\n6 4\n255\n" + img.tobytes()                       # src/nbody.cu:359-362
    assert nb.lib().nb_write_pgm(str(tmp_path / "nodir" / "x.ppm").encode(), img.ctypes.data, 6, 4) == nb.ERR_IO


def test_plan_matches_oracle_coverage(nb, oracle):
    for n in (1, 2, 100, 127, 128, 129, 130, 200, 255, 256, 257, 300, 16384, 15709, 131072, 1048576):
        for mode in (0, 1):
            p, c = nb.plan(n, coverage=mode), oracle.coverage(n, mode)
            assert (p["blocks"], p["limit_last"], p["n_active"]) == (c["blocks"], c["limit_last"], c["n_active"]), (n, mode)
            assert p["window_len"] == 128 * (c["blocks"] - 1) + c["limit_last"]


def test_driver_binary_is_built(nb):
    assert nb.DRIVER_PATH.exists(), "ppa-nbody-collisions_b200/bin/nbody (drop-in driver) was not built"
    out = subprocess.run([str(nb.DRIVER_PATH), "--bogus"], capture_output=True, text=True)
    assert out.returncode == 2 and "unknown option" in out.stderr


def test_vec2_header_matches_reference_semantics(tmp_path, host_golden):
    """include/nb_vec2.h (API-compatible Vec2f / Vec2<T>) against operator results produced by the reference's
    own include/vec2f.h (tools/make_golden_host.py, same expressions)."""
    src = tmp_path / "v.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "nb_vec2.h"
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main() {
    Vec2f a(3.f, -4.f), b(0.5f, 7.f);
    Vec2f c = a * 3.f; Vec2f d = a / 3.f; Vec2f e = a * b; Vec2f f = a + b; Vec2f g = a - b; Vec2f h = -a;
    Vec2f s(2.5f); Vec2f k = 2.f * b;
    printf("%08x %08x\n", bits(c.X), bits(c.Y)); printf("%08x %08x\n", bits(d.X), bits(d.Y));
    printf("%08x %08x\n", bits(e.X), bits(e.Y)); printf("%08x %08x\n", bits(f.X), bits(f.Y));
    printf("%08x %08x\n", bits(g.X), bits(g.Y)); printf("%08x %08x\n", bits(h.X), bits(h.Y));
    printf("%08x %08x\n", bits(s.X), bits(s.Y)); printf("%08x %08x\n", bits(k.X), bits(k.Y));
    printf("%08x\n", bits(a.length())); printf("%08x\n", bits(Vec2f(1e-3f, 7.f).length()));
    printf("%zu\n", sizeof(Vec2f));
    Vec2<double> q(1.0, 2.0); q /= 3.0; q += Vec2<double>(1.0); q *= 2.0; a[1] = 9.f;
    return (q.X > 0 && a.Element[1] == 9.f && sizeof(Vec2<double>) == 16) ? 0 : 1;
}
''')
    exe = tmp_path / "v"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert [l for l in out if l] == host_golden["vec2f"]
