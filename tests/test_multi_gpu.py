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
                                                               (16384, 100000, 1, 1024, 32), (20000, 30000, 1, 1024, 0)])
def test_two_gpus_match_oracle(oracle, nb, tmp_path, n0, field, coverage, sort_min_n, flags):
    import torch.multiprocessing as mp
    world, steps = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), n0, field, coverage, steps, str(tmp_path), sort_min_n, flags), nprocs=world, join=True)
    block = nb.generate(nb.SCENARIO_SQUARE, n0, field_w=field, field_h=field)
    par = oracle.params(field_w=field, field_h=field, coverage=coverage)
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
