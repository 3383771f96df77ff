"""Sharded CUDA path (one process per GPU, NCCL allgather of the post-step rows) against the oracle.

Needs >= 2 B200s on the box; skipped otherwise (the host-side sharding logic is covered on the CPU by
tests/test_sharding_cpu.py with gloo).  Run by hand with:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gpu_count() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank: int, world: int, port: int, n0: int, field: int, coverage: int, steps: int, out_dir: str, sort_min_n: int = 0,
            flags: int = 0):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as G
    nb = G.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    block0 = nb.generate(nb.SCENARIO_SQUARE, n0, field_w=field, field_h=field)
    sim = nb.Simulation(n0, field_w=field, field_h=field, coverage=coverage, device=rank, rank=rank, world=world,
                        event_capacity=64 * n0, sort_min_n=sort_min_n, flags=flags)
    sim.comm_init(ids[0])
    sim.upload(block0, n0)
    states = []
    events = []
    for _ in range(steps):
        sim.step(1)
        got, n = sim.download()
        states.append(got.copy())
        events.append(sim.events())
    np.savez(os.path.join(out_dir, f"rank_{rank}.npz"), *states)
    np.save(os.path.join(out_dir, f"events_{rank}.npy"), np.concatenate(events) if events else np.zeros(0))
    sim.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs")
# flags 32 = NB_FLAG_ONE_SIDED: the sorted order with row shards; without it steps on the sorted order run the two-sided
# kernel (blocks of the pair triangle dealt to the ranks, exchange of partial forces and candidate pairs)
@pytest.mark.parametrize("n0,field,coverage,sort_min_n,flags", [(16384, 100000, 0, 0, 0), (16384, 100000, 1, 0, 0), (3000, 12000, 1, 0, 0),
                                                               (700, 3000, 0, 0, 0), (16384, 100000, 1, 1024, 0), (3000, 12000, 1, 2900, 0),
                                                               (16384, 100000, 1, 1024, 32), (20000, 30000, 1, 1024, 0),
                                                               # flags 16 = NB_FLAG_MERGE_CONSERVING: the opt-in merge, sharded
                                                               (3000, 12000, 1, 0, 16), (20000, 30000, 1, 1024, 16)])
def test_two_gpus_match_oracle(oracle, nb, tmp_path, n0, field, coverage, sort_min_n, flags):
    import torch.multiprocessing as mp
    world, steps = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), n0, field, coverage, steps, str(tmp_path), sort_min_n, flags), nprocs=world, join=True)
    block = nb.generate(nb.SCENARIO_SQUARE, n0, field_w=field, field_h=field)
    par = oracle.params(field_w=field, field_h=field, coverage=coverage, merge=1 if flags & 16 else 0)
    n = n0
    ranks = [np.load(tmp_path / f"rank_{r}.npz") for r in range(world)]
    ev_all = np.concatenate([np.load(tmp_path / f"events_{r}.npy") for r in range(world)])
    for s in range(steps):
        n, _, ev_cpu = oracle.step(block, n, par, want_events=True)
        a, b = ranks[0][f"arr_{s}"], ranks[1][f"arr_{s}"]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"step {s}: the replicas diverged"
        assert len(a) == 6 * n
        pg, vg, mg, rg = nb.split(a, n)
        pc, vc, mc, rc = oracle.split(block, n)
        assert np.array_equal(mg.view(np.uint32), mc.view(np.uint32)) and np.array_equal(rg.view(np.uint32), rc.view(np.uint32))
        assert np.abs(pg - pc).max() <= 1e-5 * field
        assert np.abs(vg - vc).max() <= 1e-3 * max(np.abs(vc).max(), 1e-30)
        ev = ev_all[ev_all["step"] == s]
        ev = ev[np.argsort(ev["i"], kind="stable")]           # each rank's list is sorted; ranks own disjoint row ranges
        assert np.array_equal(ev["i"], ev_cpu["i"]) and np.array_equal(ev["j"], ev_cpu["j"]) and np.array_equal(ev["kind"], ev_cpu["kind"])


# ---------------------------------------------------------------------------------------------------
# The sharded two-sided flow at BASELINE sizes and with several steps per nb_step call (asynchronous path),
# against the same run on ONE GPU: replicas bit-identical, events / survivors / masses / radii identical to the
# single-GPU run, velocities within 1e-4 max|v| (force sums differ in rounding only: other summation order).
# ---------------------------------------------------------------------------------------------------
def _disc(nb, n):
    R = 1e5 * np.sqrt(n / 16384.0)               # the shipped scenario's surface density
    field = int(R)
    return nb.generate(nb.SCENARIO_DISC, n, extent=R, field_w=field, field_h=field), field


def _worker_calls(rank: int, world: int, port: int, n0: int, calls, out_dir: str, sort_min_n: int, flags: int, dense: float):
    sys.path.insert(0, str(ROOT))
    import zlib
    import torch
    import torch.distributed as dist
    import __graft_entry__ as G
    nb = G.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = [nb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    block0, field = _disc(nb, n0)
    if dense != 1.0:                              # crowd the bodies: many die per step (n crosses sort_min_n quickly)
        block0[:2 * n0] *= np.float32(dense)
    sim = nb.Simulation(n0, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, device=rank, rank=rank, world=world,
                        event_capacity=64 * n0, sort_min_n=sort_min_n, flags=flags)
    sim.comm_init(ids[0])
    sim.upload(block0, n0)
    ns = []
    for k in calls:
        sim.step(k)                               # k > 1: no host synchronisation between the steps of one call
        ns.append(sim.num_bodies())
    got, n = sim.download()
    ev = sim.events()
    st = sim.stats()
    assert st["overflow"] == 0 and st["events_dropped"] == 0
    np.save(os.path.join(out_dir, f"state_{rank}.npy"), got)
    np.save(os.path.join(out_dir, f"events_{rank}.npy"), ev)
    np.save(os.path.join(out_dir, f"ns_{rank}.npy"), np.array(ns))
    sim.close()
    dist.destroy_process_group()


def _one_gpu(nb, n0, calls, sort_min_n, flags, dense):
    block0, field = _disc(nb, n0)
    if dense != 1.0:
        block0[:2 * n0] *= np.float32(dense)
    one = nb.Simulation(n0, field_w=field, field_h=field, coverage=nb.COVERAGE_FULL, event_capacity=64 * n0,
                        sort_min_n=sort_min_n, flags=flags)
    one.upload(block0, n0)
    ns = []
    for k in calls:
        one.step(k)
        ns.append(one.num_bodies())
    ref, n_ref = one.download()
    ev_ref = one.events()
    one.close()
    return ref, n_ref, ev_ref, ns


def _check_against_one_gpu(nb, tmp_path, world, n0, calls, sort_min_n, flags, dense):
    ref, n_ref, ev_ref, ns_ref = _one_gpu(nb, n0, calls, sort_min_n, flags, dense)
    states = [np.load(tmp_path / f"state_{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(states[0].view(np.uint32), states[r].view(np.uint32)), f"replica {r} differs from replica 0"
    for r in range(world):
        assert list(np.load(tmp_path / f"ns_{r}.npy")) == ns_ref, f"rank {r}: body counts per call"
    assert len(states[0]) == 6 * n_ref
    evs = np.concatenate([np.load(tmp_path / f"events_{r}.npy") for r in range(world)])
    evs = evs[np.lexsort((evs["i"], evs["step"]))]        # stable: a row's events stay in visit order
    assert len(evs) == len(ev_ref)
    for k in ("step", "i", "j", "kind"):
        assert np.array_equal(evs[k], ev_ref[k]), f"event {k}s differ from the single-GPU run"
    p1, v1, m1, r1 = nb.split(states[0], n_ref)
    p2, v2, m2, r2 = nb.split(ref, n_ref)
    assert np.array_equal(m1.view(np.uint32), m2.view(np.uint32)) and np.array_equal(r1.view(np.uint32), r2.view(np.uint32))
    assert np.abs(v1 - v2).max() <= 1e-4 * np.abs(v2).max()
    return ns_ref


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("n0,calls", [(131072, (3,)), (262144, (2, 2))])
def test_sharded_two_sided_matches_one_gpu(nb, tmp_path, world, n0, calls):
    """BASELINE-sized discs, several steps per nb_step call."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker_calls, args=(world, _free_port(), n0, calls, str(tmp_path), 0, 0, 1.0), nprocs=world, join=True)
    _check_against_one_gpu(nb, tmp_path, world, n0, calls, 0, 0, 1.0)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_crossing_sort_threshold_inside_one_call(nb, tmp_path, world):
    """The live body count falls through sort_min_n in the middle of one nb_step(k) call: every rank must keep issuing
    the same collectives (the per-rank graph choice raced with the device)."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    n0, dense, calls = 49152, 0.5, (6, 2)          # oracle: n = 39844 39250 38791 38208 37505 36732 35879 35050
    sort_min_n = 38000                            # ... so the count crosses after the 5th step of the first call
    mp.spawn(_worker_calls, args=(world, _free_port(), n0, calls, str(tmp_path), sort_min_n, 0, dense), nprocs=world, join=True)
    ns = _check_against_one_gpu(nb, tmp_path, world, n0, calls, sort_min_n, 0, dense)
    assert ns[0] < sort_min_n < n0, f"the scenario must cross the threshold inside the first call (n after it: {ns[0]})"


# ---------------------------------------------------------------------------------------------------
# The C++ host: `nbody --gpus N` (one thread + one context per GPU, NCCL between them) against `nbody --gpus 1`
# ---------------------------------------------------------------------------------------------------
def _driver(nb, cwd, *args):
    import subprocess
    r = subprocess.run([str(nb.DRIVER_PATH), "--no-images", *args], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def _read_state(path):
    raw = path.read_bytes()
    n = int(np.frombuffer(raw[:4], dtype=np.int32)[0])
    return np.frombuffer(raw[4:], dtype=np.float32), n


@pytest.mark.parametrize("world", [2, 4])
def test_cpp_driver_multi_gpu(nb, tmp_path, world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cfg = ("particleCount={n}\ntotalIterations={it}\nsave_Image_Every_Xth_Iteration=100\ntimestep=0.2f\nradiusGrowthRate=0.1f\n"
           "minRandBodyMass=1e4f\nmaxRandBodyMass=1e17f\nminRadius=50.f\nmaxRadius=200.f\nimgWidth=32\nimgHeight=32\n"
           "fieldWidth={f}\nfieldHeight={f}\nimagePath=.\n")
    # all-pairs coverage above the sort threshold: the two-sided kernel's integer force sums make the result independent
    # of the number of GPUs, bit for bit -- state, stdout and the merged event list
    (tmp_path / "nbodyConfig.txt").write_text(cfg.format(n=60000, it=5, f=190000))
    out1 = _driver(nb, tmp_path, "--coverage", "full", "--dump-state", "one.bin", "--dump-events", "one.csv")
    outw = _driver(nb, tmp_path, "--coverage", "full", "--gpus", str(world), "--dump-state", "many.bin", "--dump-events", "many.csv")
    strip = lambda o: [ln for ln in o.splitlines() if not ln.startswith("Time taken")]
    assert strip(out1) == strip(outw)
    assert (tmp_path / "one.bin").read_bytes() == (tmp_path / "many.bin").read_bytes()
    assert (tmp_path / "one.csv").read_text() == (tmp_path / "many.csv").read_text()
    assert len((tmp_path / "one.csv").read_text().splitlines()) > 100
    # the reference's own coverage (one-sided kernel, rows sharded): same events and survivors, masses and radii bit for
    # bit, trajectories within the summation-order tolerance
    (tmp_path / "nbodyConfig.txt").write_text(cfg.format(n=5000, it=6, f=20000))
    _driver(nb, tmp_path, "--dump-state", "one.bin", "--dump-events", "one.csv")
    _driver(nb, tmp_path, "--gpus", str(world), "--dump-state", "many.bin", "--dump-events", "many.csv")
    assert (tmp_path / "one.csv").read_text() == (tmp_path / "many.csv").read_text()
    a, na = _read_state(tmp_path / "one.bin")
    b, nbb = _read_state(tmp_path / "many.bin")
    assert na == nbb and na < 5000
    pa, va, ma, ra = nb.split(a, na)
    pb, vb, mb, rb = nb.split(b, nbb)
    assert np.array_equal(ma.view(np.uint32), mb.view(np.uint32)) and np.array_equal(ra.view(np.uint32), rb.view(np.uint32))
    assert np.abs(va - vb).max() <= 1e-4 * np.abs(va).max() and np.abs(pa - pb).max() <= 1e-5 * 20000
