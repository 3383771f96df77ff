"""The CPU oracle against the golden vectors.

host_golden.json   : produced here by compiling the reference's own headers
                     (tools/make_golden_host.py).
gpuref_golden.json : produced on a B200 by stepping the UNMODIFIED reference
                     kernels (tools/make_golden_gpuref.py); every scenario was
                     bit-exact against the oracle in that run, and this test
                     replays the hashes on the CPU so the pin cannot rot.
"""
import json

import numpy as np
import pytest


@pytest.fixture(scope="module")
def host_golden(golden_dir):
    return json.loads((golden_dir / "host_golden.json").read_text())


@pytest.fixture(scope="module")
def gpuref_golden(golden_dir):
    return json.loads((golden_dir / "gpuref_golden.json").read_text())


def test_rng_stream(oracle, host_golden):
    g = host_golden["rng"]
    assert [str(v) for v in oracle.rng_stream(1024, 8)] == g["seed1024_ival64"]
    assert [str(v) for v in oracle.rng_stream(7, 4)] == g["seed7_ival64"]
    assert [float(v) for v in g["seed1024_fval"]] == oracle.rng_fvals(1024, 8)


def test_initial_bodies_shipped(oracle, host_golden):
    g = host_golden["init_shipped"]
    n = g["n"]
    block = oracle.init_square(n)
    assert f"{oracle.fnv(block):016x}" == g["fnv1a64_block"]
    pos, vel, m, r = oracle.split(block, n)
    for idx, bits in g["bodies_xymr_bits"].items():
        i = int(idx)
        got = [f"{int(v):08x}" for v in np.array([pos[i, 0], pos[i, 1], m[i], r[i]], dtype=np.float32).view(np.uint32)]
        assert got == bits
    assert not vel.any()
    assert float(g["sum_m"]) == pytest.approx(m.astype(np.float64).sum(), rel=1e-15)


def test_coverage_descriptor(oracle):
    # src/nbody.cu:473 (floor) and :194 (n % 129)
    assert oracle.coverage(16384, oracle.COVERAGE_REFERENCE) == {"n": 16384, "blocks": 128, "limit_last": 1, "n_active": 16384}
    assert oracle.coverage(15709, oracle.COVERAGE_REFERENCE) == {"n": 15709, "blocks": 122, "limit_last": 100, "n_active": 15616}
    assert oracle.coverage(100, oracle.COVERAGE_REFERENCE) == {"n": 100, "blocks": 1, "limit_last": 100, "n_active": 100}
    assert oracle.coverage(129, oracle.COVERAGE_REFERENCE) == {"n": 129, "blocks": 1, "limit_last": 0, "n_active": 128}
    assert oracle.coverage(300, oracle.COVERAGE_FULL) == {"n": 300, "blocks": 3, "limit_last": 44, "n_active": 300}


SMALL = ["small1", "small2", "small5", "small100", "small127", "small128", "small129", "small130", "small200",
         "small255", "small256", "small257", "small258", "small300", "small383", "small384", "small385",
         "small1000", "small1500", "dense3000", "dense4096"]


def _replay(oracle, sc, max_steps=None):
    n = sc["n0"]
    block = oracle.init_square(n, seed=sc["seed"], field_w=sc["field"], field_h=sc["field"])
    par = oracle.params(dt=sc["dt"], growth=sc["growth"], field_w=sc["field"], field_h=sc["field"],
                        coverage=oracle.COVERAGE_REFERENCE)
    trace = sc["trace"] if max_steps is None else sc["trace"][:max_steps]
    for s, want in enumerate(trace):
        n, _, _ = oracle.step(block, n, par)
        assert n == want["n"], f"step {s}: n"
        got = f"{oracle.fnv(block[:6 * n]):016x}" if n > 0 else f"{0:016x}"
        assert got == want["fnv"], f"step {s}: state hash"


@pytest.mark.parametrize("name", SMALL)
def test_reference_kernels_small(oracle, gpuref_golden, name):
    sc = gpuref_golden["scenarios"][name]
    assert sc["oracle_bit_exact"]
    _replay(oracle, sc)


def test_reference_kernels_shipped(oracle, gpuref_golden):
    """Shipped nbodyConfig.txt scenario; 16 of the 60 recorded steps keep the CPU suite short."""
    sc = gpuref_golden["scenarios"]["shipped"]
    assert sc["oracle_bit_exact"] and len(sc["trace"]) == 60
    _replay(oracle, sc, max_steps=16)


def test_survivor_trace_matches_survey(gpuref_golden):
    """SURVEY.md 8c listed a provisional survivor trace; the reference kernels confirm it."""
    ns = [t["n"] for t in gpuref_golden["scenarios"]["shipped"]["trace"]]
    assert ns[:12] == [15709, 15682, 15643, 15598, 15535, 15478, 15404, 15321, 15227, 15142, 15036, 14929]
    assert (ns[19], ns[29], ns[39], ns[49], ns[59]) == (14118, 13058, 12026, 11018, 10147)


def test_events_step0_shipped(oracle):
    n = 16384
    block = oracle.init_square(n)
    n1, stats, ev = oracle.step(block, n, oracle.params(), want_events=True)
    assert n1 == 15709 and stats["pairs"] == 266338304 and stats["asserts"] == 0
    assert int((ev["kind"] == oracle.EV_ABSORB).sum()) == 691
    assert int((ev["kind"] == oracle.EV_KILLED).sum()) == 694
    assert (np.diff(ev["i"]) >= 0).all()


def test_rows_match_step(oracle):
    """orc_rows (row sampling used at large N) agrees with the full step."""
    n = 1000
    block = oracle.init_square(n, field_w=4000, field_h=4000)
    par = oracle.params(field_w=4000, field_h=4000, coverage=oracle.COVERAGE_FULL)
    idx = np.arange(0, n, 7)
    out, hits, visited = oracle.rows(block, n, par, idx)
    assert (visited == n - 1).all()
    full = block.copy()
    # emulate commit without compaction by evaluating every row
    allrows, _, _ = oracle.rows(block, n, par, np.arange(n))
    assert np.array_equal(allrows[idx], out)
    n1, stats, _ = oracle.step(full, n, par)
    alive = allrows[:, 4] != 0
    assert n1 == int(alive.sum())
    pos, vel, m, r = oracle.split(full, n1)
    assert np.array_equal(pos, allrows[alive][:, 2:4]) and np.array_equal(vel, allrows[alive][:, 0:2])
    assert np.array_equal(m, allrows[alive][:, 4]) and np.array_equal(r, allrows[alive][:, 5])


@pytest.mark.parametrize("n", [1, 1000, 16385])
def test_scenario_generators_match_the_product(oracle, nb, n):
    """The oracle-side generators of the synthetic BASELINE scenarios (what bench.py's reference arm feeds the
    reference kernels) produce the very bits of the product's nb_generate."""
    assert np.array_equal(oracle.init_disc(n, 8e5).view(np.uint32),
                          nb.generate(nb.SCENARIO_DISC, n, extent=8e5, field_w=800000, field_h=800000).view(np.uint32))
    assert np.array_equal(oracle.init_two_galaxy(n, 8e5).view(np.uint32),
                          nb.generate(nb.SCENARIO_TWO_GALAXY, n, extent=8e5, field_w=3000000, field_h=3000000).view(np.uint32))
    assert np.array_equal(oracle.init_square(n).view(np.uint32), nb.generate(nb.SCENARIO_SQUARE, n).view(np.uint32))


def _render_golden(golden_dir):
    p = golden_dir / "render_golden.json"
    if not p.exists():
        pytest.skip("tests/golden/render_golden.json not generated yet (tools/make_golden_render.py on a GPU box)")
    return json.loads(p.read_text())


def test_render_restatement_matches_the_reference_images(oracle, golden_dir):
    """orc_render against images of the UNMODIFIED generateImage kernel (src/nbody.cu:294-348), launched with the
    reference loop's own grid: byte-exact (FNV of the image bytes), including the bodies its stale grid leaves out."""
    g = _render_golden(golden_dir)
    for name, sc in g["scenarios"].items():
        assert sc["oracle_byte_exact"], name
        block = oracle.init_square(sc["n"], field_w=sc["field"], field_h=sc["field"], min_radius=sc["min_radius"],
                                   max_radius=sc["max_radius"])
        par = oracle.params(field_w=sc["field"], field_h=sc["field"], coverage=oracle.COVERAGE_REFERENCE)
        n = sc["n"]
        for _ in range(sc["steps"]):
            n, _, _ = oracle.step(block, n, par)
        img = oracle.render(block, n, sc["width"], sc["height"], sc["field"], sc["field"], grid_n=sc["grid_n"])
        assert f"{oracle.fnv(img):016x}" == sc["fnv"] and int((img == 0).sum()) == sc["body_pixels"], name
