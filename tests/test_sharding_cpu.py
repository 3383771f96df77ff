"""The N > 1 path on the CPU: world_size-2 and -3 `gloo` runs of the host-side sharding logic.

Each rank owns the i-block-aligned row range the library's own plan (nb_plan_host, the code the device
runs at the end of every step) assigns it, evaluates only those rows (oracle.rows stands in for the force +
finish kernels), exchanges the post-step rows with one all_gather (the NCCL allgather of the CUDA path) and
then runs the replicated stable compaction.  The result must be bit-identical to the single-process oracle
step for several steps while n shrinks, in both coverage modes, including the reference mode's frozen tail.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n0: int, field: int, coverage: int, steps: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as G
    from oracle import oracle as O
    nb = G.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    block = nb.generate(nb.SCENARIO_SQUARE, n0, field_w=field, field_h=field)
    n = n0
    par = O.params(field_w=field, field_h=field, coverage=coverage, threads=2)
    trace = []
    for _ in range(steps):
        if n == 0:
            break
        plan = nb.plan(n, coverage=coverage, rank=rank, world=world)
        rpr = plan["rows_per_rank"]
        assert rpr % 512 == 0 and plan["row_lo"] == min(rank * rpr, n) and plan["row_hi"] == min((rank + 1) * rpr, n)
        rows = np.arange(plan["row_lo"], plan["row_hi"], dtype=np.int32)
        mine = np.zeros((rpr, 6), dtype=np.float32)           # the rank's chunk of the allgather payload
        if len(rows):
            out, _, _ = O.rows(block, n, par, rows)           # rows >= n_active come back unchanged (frozen tail)
            mine[:len(rows)] = out
        chunks = [torch.zeros(rpr, 6) for _ in range(world)]
        dist.all_gather(chunks, torch.from_numpy(mine))
        post = torch.cat(chunks).numpy()[:n]                  # vx, vy, px, py, m, r of every pre-step body
        keep = post[:, 4] != 0.0                              # replicated compaction, src/nbody.cu:488-510
        s = post[keep]
        n = int(keep.sum())
        block = np.concatenate([s[:, 2:4].reshape(-1), s[:, 0:2].reshape(-1), s[:, 4], s[:, 5]]).astype(np.float32)
        trace.append((n, O.fnv(block) if n else 0))
    np.save(os.path.join(out_dir, f"trace_{rank}.npy"), np.array(trace, dtype=np.uint64))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n0,field,coverage", [(2, 3000, 12000, 0), (2, 3000, 12000, 1), (3, 1500, 6000, 0),
                                                     (2, 300, 2000, 1), (2, 700, 3000, 0)])
def test_sharded_step_equals_single_process(oracle, nb, tmp_path, world, n0, field, coverage):
    import torch.multiprocessing as mp
    steps = 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n0, field, coverage, steps, str(tmp_path)), nprocs=world, join=True)
    # single-process oracle
    block = nb.generate(nb.SCENARIO_SQUARE, n0, field_w=field, field_h=field)
    par = oracle.params(field_w=field, field_h=field, coverage=coverage)
    n = n0
    want = []
    for _ in range(steps):
        if n == 0:
            break
        n, _, _ = oracle.step(block, n, par)
        want.append((n, oracle.fnv(block[:6 * n]) if n else 0))
    want = np.array(want, dtype=np.uint64)
    for r in range(world):
        got = np.load(tmp_path / f"trace_{r}.npy")
        assert np.array_equal(got, want), f"rank {r}: sharded trace differs from the single-process oracle"


def test_plan_shards_cover_all_rows(nb):
    for n in (0, 1, 127, 128, 129, 511, 512, 513, 1023, 1024, 1025, 16384, 15709, 131072, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            for cov in (nb.COVERAGE_REFERENCE, nb.COVERAGE_FULL):
                lo_prev = 0
                active = 0
                for rank in range(world):
                    p = nb.plan(n, coverage=cov, rank=rank, world=world)
                    assert p["row_lo"] == lo_prev and p["row_lo"] <= p["row_act_hi"] <= p["row_hi"] <= n
                    assert p["row_lo"] % 512 == 0 or p["row_lo"] == n
                    assert p["units"] == (p["n_iblocks"] * p["n_jtiles"]) * (p["units"] // max(p["n_iblocks"] * p["n_jtiles"], 1) or 1) \
                        or p["units"] == 0
                    lo_prev = p["row_hi"]
                    active += p["row_act_hi"] - p["row_lo"]
                assert lo_prev == n
                assert active == nb.plan(n, coverage=cov)["n_active"]


def test_two_sided_plan_covers_every_tile_pair_once(nb):
    """Host logic of the two-sided force kernel: which steps use it, the cut of the pair triangle into blocks, and the
    deal of the blocks to the ranks -- every unordered tile pair belongs to exactly one block of exactly one rank."""
    for n, world in ((1023, 1), (40959, 1), (40960, 1), (49152, 1), (131072, 1), (1048576, 1), (4194304, 1), (1000003, 2),
                     (1048576, 8), (4194304, 8), (70000, 3)):
        plans = [nb.plan(n, coverage=nb.COVERAGE_FULL, rank=r, world=world) for r in range(world)]
        p = plans[0]
        uses = n >= 40960
        assert p["sorted"] == int(uses) and p["two_sided"] == int(uses), (n, world, p)
        if not uses:
            continue
        T, S, Q = p["n_jtiles"], p["sym_S"], p["sym_Q"]
        assert T == (n + 511) // 512 and Q <= (256 if world == 1 else 512) and (Q - 1) * S < T <= Q * S
        assert p["sym_blocks"] == Q * (Q + 1) // 2
        assert all(q["sym_S"] == S and q["sym_Q"] == Q and q["sym_blocks"] == p["sym_blocks"] for q in plans)
        if Q > 64:
            Q_check = range(0, p["sym_blocks"], 97)          # sample: the full enumeration is done for small Q below
        else:
            Q_check = range(p["sym_blocks"])
        for b in Q_check:
            R, C = nb.plan_block(Q, b)
            assert 0 <= R <= C < Q and nb.plan_block_index(Q, R, C) == b and nb.plan_block_index(Q, C, R) == b
    for Q in (1, 2, 3, 7, 33):
        seen = {}
        for b in range(Q * (Q + 1) // 2):
            R, C = nb.plan_block(Q, b)
            assert (R, C) not in seen
            seen[(R, C)] = b
        assert len(seen) == Q * (Q + 1) // 2 and all(R <= C for R, C in seen)
        # the half-size diagonal blocks come last (a short tail of the queue)
        assert sorted(seen[(R, R)] for R in range(Q)) == list(range(Q * (Q - 1) // 2, Q * (Q + 1) // 2))
        for world in (1, 2, 3, 8):
            owner = {rc: b % world for rc, b in seen.items()}
            load = [sum(1 for o in owner.values() if o == r) for r in range(world)]
            assert max(load) - min(load) <= 1


def test_two_sided_plan_flags(nb):
    assert nb.plan(131072, flags=nb.FLAG_ONE_SIDED)["two_sided"] == 0 and nb.plan(131072, flags=nb.FLAG_ONE_SIDED)["sorted"] == 1
    assert nb.plan(131072, flags=nb.FLAG_NO_SORT)["sorted"] == 0 and nb.plan(131072, flags=nb.FLAG_NO_SORT)["two_sided"] == 0
    assert nb.plan(131072, coverage=nb.COVERAGE_REFERENCE)["two_sided"] == 0
    assert nb.plan(5000, sort_min_n=1024)["two_sided"] == 1 and nb.plan(1023, sort_min_n=1)["two_sided"] == 0
    assert nb.plan(30000, n_max=131072)["two_sided"] == 0            # a big context whose body count has dropped
    assert nb.plan(131072, flags=nb.FLAG_MERGE_CONSERVING)["sorted"] == 0
    with pytest.raises(nb.NbodyError):
        nb.plan_block(4, 10)
