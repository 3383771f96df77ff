"""Edge cases of the time step on the GPU against the oracle, and the error behaviour of the C ABI."""
import numpy as np
import pytest

from test_parity_gpu import _run_side_by_side

pytestmark = pytest.mark.gpu


def _block(pos, vel, m, r):
    n = len(m)
    return np.concatenate([np.asarray(pos, np.float32).reshape(-1), np.asarray(vel, np.float32).reshape(-1),
                           np.asarray(m, np.float32), np.asarray(r, np.float32)]).astype(np.float32), n


def _random(n, field, rng, rmin=50.0, rmax=200.0, mmin=1e4, mmax=1e17, vmax=0.0):
    pos = rng.uniform(-field, field, (n, 2))
    vel = rng.uniform(-vmax, vmax, (n, 2)) if vmax > 0 else np.zeros((n, 2))
    return _block(pos, vel, rng.uniform(mmin, mmax, n), rng.uniform(rmin, rmax, n))


@pytest.mark.parametrize("coverage", [0, 1])
def test_coincident_and_equal_mass_bodies(nb, oracle, coverage):
    """Identical positions (d2 = 0 is a hit, never a division), equal masses (both absorb, neither dies,
    src/nbody.cu:215), chains of overlapping bodies (victim absorbed twice, absorber deleted)."""
    rng = np.random.default_rng(5)
    block, n = _random(600, 3000, rng)
    pos = block[:2 * n].reshape(n, 2)
    m = block[4 * n:5 * n]
    pos[10] = pos[11]                       # coincident pair
    pos[300] = pos[301] = pos[302]          # coincident triple
    m[300] = m[301] = m[302] = np.float32(5e16)
    m[10] = m[11]
    pos[400:420] = pos[400] + np.arange(20)[:, None] * np.float32(30.0)   # a chain of overlapping bodies
    _run_side_by_side(nb, oracle, block, n, 5, coverage, 3000)


@pytest.mark.parametrize("coverage", [0, 1])
def test_zero_radius_and_zero_mass(nb, oracle, coverage):
    """r = 0 bodies collide only when coincident; bodies uploaded with m = 0 exert no force and are dropped by
    the first compaction (src/nbody.cu:490) - in reference coverage only if they own a thread, like the reference."""
    rng = np.random.default_rng(6)
    block, n = _random(700, 4000, rng, rmin=0.0, rmax=0.0)
    block[4 * n + 5] = 0.0
    block[4 * n + 699] = 0.0
    block[2 * 17:2 * 17 + 2] = block[2 * 18:2 * 18 + 2]          # one coincident pair so that something happens
    _run_side_by_side(nb, oracle, block, n, 4, coverage, 4000)


def test_wall_reflection_and_moving_bodies(nb, oracle):
    """Fast bodies next to the walls: the reference's odd reflection test (position + dv vs +-field -+ r,
    src/nbody.cu:256-261) flips velocities; large dt moves bodies outside the field."""
    rng = np.random.default_rng(7)
    block, n = _random(1000, 5000, rng, vmax=800.0)
    pos = block[:2 * n].reshape(n, 2)
    pos[:100, 0] = 5000 - rng.uniform(0, 150, 100)     # hugging the +x wall
    pos[100:200, 1] = -5000 + rng.uniform(0, 150, 100)
    # re-synchronised every step: with dt = 2 and |v| ~ 1e3 a free run is chaotic within a few steps
    _run_side_by_side(nb, oracle, block, n, 6, 1, 5000, dt=2.0, resync=True)
    _run_side_by_side(nb, oracle, block, n, 6, 0, 5000, dt=2.0, resync=True)


def test_capacity_larger_than_n_and_reupload(nb, oracle):
    n_max, n, field = 5000, 1300, 6000
    block = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
    sim = nb.Simulation(n_max, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL)
    par = oracle.params(field_w=field, field_h=field, coverage=oracle.COVERAGE_FULL)
    for _ in range(2):                                  # the second round re-uploads into a used context
        sim.upload(block, n)
        cpu, n_cpu = block.copy(), n
        for _s in range(3):
            sim.step(1)
            n_cpu, _, _ = oracle.step(cpu, n_cpu, par)
        got, n_gpu = sim.download()
        assert n_gpu == n_cpu == sim.num_bodies()
        assert np.array_equal(got[4 * n_gpu:].view(np.uint32), cpu[4 * n_cpu:6 * n_cpu].view(np.uint32))
    sim.close()


def test_empty_and_single_body(nb):
    sim = nb.Simulation(64, field_w=1000, field_h=1000)
    sim.upload(np.zeros(0, dtype=np.float32), 0)
    sim.step(3)
    assert sim.num_bodies() == 0 and sim.download()[1] == 0
    one = np.array([10, 20, 3, -4, 5e10, 100], dtype=np.float32)
    sim.upload(one, 1)
    sim.step(2)
    got, n = sim.download()
    assert n == 1 and np.allclose(got[:2], [10 + 2 * 0.2 * 3, 20 - 2 * 0.2 * 4]) and got[4] == np.float32(5e10)
    sim.close()


def test_error_codes(nb):
    n, field = 2048, 3000                               # dense: about 10 hits per body
    block = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
    with pytest.raises(nb.NbodyError) as e:
        nb.Simulation(n, field_w=field, field_h=field).upload(np.zeros(6 * (n + 1), np.float32), n + 1)
    assert e.value.code == nb.ERR_CAPACITY
    sim = nb.Simulation(n, field_w=field, field_h=field, candidate_capacity=64)
    sim.upload(block, n)
    sim.step(1)
    with pytest.raises(nb.NbodyError) as e:
        sim.sync()
    assert e.value.code == nb.ERR_CANDIDATE_OVERFLOW and sim.stats()["overflow"] == 1
    sim.close()
    sim = nb.Simulation(n, field_w=field, field_h=field, event_capacity=16)
    sim.upload(block, n)
    sim.step(1)
    with pytest.raises(nb.NbodyError) as e:
        sim.events()
    assert e.value.code == nb.ERR_EVENT_OVERFLOW
    with pytest.raises(nb.NbodyError) as e:
        sim.download(np.zeros(6 * 10, np.float32))
    assert e.value.code == nb.ERR_CAPACITY
    sim.close()
    with pytest.raises(nb.NbodyError) as e:
        nb.Simulation(n, field_w=field, field_h=field).events()
    assert e.value.code == nb.ERR_INVALID
    with pytest.raises(nb.NbodyError) as e:
        nb.Simulation(n, field_w=field, field_h=field, world=2, rank=0).step(1)
    assert e.value.code == nb.ERR_COMM
    with pytest.raises(nb.NbodyError):
        nb.Simulation(0)
    with pytest.raises(nb.NbodyError):
        nb.Simulation(16, coverage=7)


def test_every_tiling_edge_in_reference_coverage(nb, oracle):
    """A sweep over body counts around every multiple of 128 / 129 / 512 (the reference's tile arithmetic,
    src/nbody.cu:186-207,473, and this library's i-block, j-tile and part sizes), dense enough that collisions
    happen at every size: 3 steps each against the oracle, one context re-used through nb_upload."""
    sizes = sorted(set(list(range(1, 13)) + [63, 64, 65] + list(range(126, 133)) + [191, 192, 193] + list(range(254, 261)) +
                       [383, 384, 385, 386, 387, 511, 512, 513, 514, 640, 767, 768, 769, 1023, 1024, 1025, 1151, 1152,
                        1153, 1535, 1536, 1537, 2047, 2048, 2049]))
    sim = nb.Simulation(max(sizes), field_w=3000, field_h=3000, coverage=nb.COVERAGE_REFERENCE, event_capacity=1 << 18)
    par = oracle.params(field_w=3000, field_h=3000, coverage=oracle.COVERAGE_REFERENCE)
    for n in sizes:
        rng = np.random.default_rng(n)
        block, _ = _random(n, 3000 * min(1.0, np.sqrt(n / 600.0)), rng)
        sim.upload(block, n)
        cpu, n_cpu = block.copy(), n
        for s in range(3):
            if n_cpu == 0:
                break
            sim.step(1)
            n_cpu, _, ev_cpu = oracle.step(cpu, n_cpu, par, want_events=True)
            got, n_gpu = sim.download()
            ev = sim.events()
            assert n_gpu == n_cpu, (n, s)
            assert len(ev) == len(ev_cpu) and np.array_equal(ev["i"], ev_cpu["i"]) and np.array_equal(ev["j"], ev_cpu["j"]) \
                and np.array_equal(ev["kind"], ev_cpu["kind"]), (n, s)
            assert np.array_equal(got[4 * n_gpu:].view(np.uint32), cpu[4 * n_cpu:6 * n_cpu].view(np.uint32)), (n, s)
            if n_cpu:
                scale = max(float(np.abs(cpu[2 * n_cpu:4 * n_cpu]).max()), 1e-30)
                assert np.abs(got[2 * n_gpu:4 * n_gpu] - cpu[2 * n_cpu:4 * n_cpu]).max() <= 1e-3 * scale, (n, s)
    sim.close()


def test_every_tiling_edge_in_full_coverage(nb, oracle):
    sizes = [1, 2, 3, 31, 32, 33, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 1536, 2047, 2048, 2049, 3000]
    sim = nb.Simulation(max(sizes), field_w=4000, field_h=4000, coverage=nb.COVERAGE_FULL, event_capacity=1 << 18, sort_min_n=1024)
    par = oracle.params(field_w=4000, field_h=4000, coverage=oracle.COVERAGE_FULL)
    for n in sizes:
        rng = np.random.default_rng(1000 + n)
        block, _ = _random(n, 4000 * min(1.0, np.sqrt(n / 1000.0)), rng)
        sim.upload(block, n)
        cpu, n_cpu = block.copy(), n
        for s in range(3):
            if n_cpu == 0:
                break
            sim.step(1)
            n_cpu, _, ev_cpu = oracle.step(cpu, n_cpu, par, want_events=True)
            got, n_gpu = sim.download()
            ev = sim.events()
            assert n_gpu == n_cpu, (n, s)
            assert len(ev) == len(ev_cpu) and np.array_equal(ev["i"], ev_cpu["i"]) and np.array_equal(ev["j"], ev_cpu["j"]), (n, s)
            assert np.array_equal(got[4 * n_gpu:].view(np.uint32), cpu[4 * n_cpu:6 * n_cpu].view(np.uint32)), (n, s)
    sim.close()
