"""The N > 1 path on the CPU: world_size-2 and -3 `gloo` runs of the host-side sharding logic (row shards of the
one-sided steps; at the end of the file the block deal and the exchange of the two-sided steps).

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
    for n, world in ((1023, 1), (12287, 1), (12288, 1), (16384, 1), (16384, 2), (40959, 1), (40959, 2), (40960, 1), (49152, 1), (131072, 1),
                     (196607, 1), (196608, 1), (262144, 2), (1048576, 1), (4194304, 1), (1000003, 2), (1048576, 8), (4194304, 8), (70000, 3)):
        plans = [nb.plan(n, coverage=nb.COVERAGE_FULL, rank=r, world=world) for r in range(world)]
        p = plans[0]
        # the cell-sorted order and a two-sided kernel from 12288 bodies on: below 196608 bodies the warp-level kernel, from
        # there on the CTA-level kernel whose cut of the pair triangle is checked here
        uses = n >= 12288
        assert p["sorted"] == int(uses) and p["two_sided"] == int(uses), (n, world, p)
        if not uses or n < 196608:                 # the warp-level kernel's range (any number of GPUs): no tile-pair cut
            continue
        T, S, Q = p["n_jtiles"], p["sym_S"], p["sym_Q"]
        # the cut depends on the tile count alone (never on the number of GPUs): one GPU and several round alike
        assert T == (n + 511) // 512 and S == min(8, max(1, T // 1024)) and (Q - 1) * S < T <= Q * S
        assert p["sym_lgu"] == 0 or (S == 1 and (p["sym_blocks"] << (p["sym_lgu"] - 1)) < 16 * 444)
        assert p["sym_lgu"] == 2 or (p["sym_blocks"] << p["sym_lgu"]) >= 16 * 444 or S > 1
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
    # without the sorted order one GPU still runs the warp-level kernel below 196608 bodies (every round pre-tested)
    assert nb.plan(262144, flags=nb.FLAG_NO_SORT)["sorted"] == 0 and nb.plan(262144, flags=nb.FLAG_NO_SORT)["two_sided"] == 0
    assert nb.plan(30000, flags=nb.FLAG_NO_SORT)["sorted"] == 0 and nb.plan(30000, flags=nb.FLAG_NO_SORT)["two_sided"] == 1
    assert nb.plan(30000, flags=nb.FLAG_NO_SORT, world=2)["two_sided"] == 0
    assert nb.plan(131072, coverage=nb.COVERAGE_REFERENCE)["two_sided"] == 0
    assert nb.plan(5000, sort_min_n=1024)["two_sided"] == 1 and nb.plan(1023, sort_min_n=1)["two_sided"] == 0
    assert nb.plan(5000, sort_min_n=1024)["sorted"] == 1 and nb.plan(5000, sort_min_n=1024, world=2)["two_sided"] == 1
    p = nb.plan(30000, n_max=131072)                                  # a big context whose body count has dropped
    assert p["sorted"] == 1 and p["two_sided"] == 1
    assert nb.plan(30000, n_max=131072, world=2)["two_sided"] == 1 and nb.plan(9000, n_max=131072)["two_sided"] == 0
    assert nb.plan(131072, flags=nb.FLAG_MERGE_CONSERVING)["sorted"] == 1 and nb.plan(131072, flags=nb.FLAG_MERGE_CONSERVING)["two_sided"] == 1
    with pytest.raises(nb.NbodyError):
        nb.plan_block(4, 10)


def _pair_block(x, y, m, r, rows, cols, own):
    """Two-sided evaluation of the tile pair rows x cols the way force_sym_kernel defines it: float32 predicate with the
    reference's roundings (d2 = fma(dx, dx, dy * dy) <= (r_i + r_j)^2), hit pairs excluded from the force on BOTH bodies
    and reported for both rows; own tile: every ordered pair is met on its own, self pair skipped."""
    xi, yi, xj, yj = x[rows, None], y[rows, None], x[None, cols], y[None, cols]
    dx = (xj - xi).astype(np.float32)
    dy = (yj - yi).astype(np.float32)
    dy2 = (dy * dy).astype(np.float32)
    d2 = (dx.astype(np.float64) * dx.astype(np.float64) + dy2.astype(np.float64)).astype(np.float32)     # fmaf
    rs = (r[rows, None] + r[None, cols]).astype(np.float32)
    hit = d2 <= (rs * rs).astype(np.float32)
    valid = np.ones_like(hit)
    if own:
        valid &= rows[:, None] != cols[None, :]
    w = np.where(valid & ~hit, np.maximum(d2.astype(np.float64), 1.0) ** -1.5, 0.0)     # non-hit pairs are >= 100 apart
    fi = np.stack([(w * dx * m[None, cols]).sum(axis=1), (w * dy * m[None, cols]).sum(axis=1)], axis=1)
    fj = np.stack([-(w * dx * m[rows, None]).sum(axis=0), -(w * dy * m[rows, None]).sum(axis=0)], axis=1)
    ii, jj = np.nonzero(valid & hit)
    pairs = [(int(rows[a]), int(cols[b])) for a, b in zip(ii, jj)]
    if not own:
        pairs += [(b, a) for a, b in pairs]
    return fi, (None if own else fj), pairs


def _two_sided_worker(rank: int, world: int, port: int, n: int, field: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as G
    nb = G.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    block = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
    pos, _, m, r = nb.split(block, n)
    x, y = pos[:, 0].copy(), pos[:, 1].copy()
    plan = nb.plan(n, rank=rank, world=world, sort_min_n=1024)          # the library's own cut of the pair triangle
    assert plan["two_sided"] == 1
    S, Q, T = plan["sym_S"], plan["sym_Q"], plan["n_jtiles"]
    F = np.zeros((n, 2))
    pairs = []
    tile = lambda t: np.arange(t * 512, min((t + 1) * 512, n))
    for b in range(rank, plan["sym_blocks"], world):                     # rank r takes blocks r, r + W, ...
        R, C = nb.plan_block(Q, b)
        for I in range(R * S, min((R + 1) * S, T)):
            for J in range(I if R == C else C * S, min((C + 1) * S, T)):
                fi, fj, pp = _pair_block(x, y, m, r, tile(I), tile(J), own=(I == J))
                F[tile(I)] += fi
                if fj is not None:
                    F[tile(J)] += fj
                pairs += pp
    # the exchange, as the library does it: the forces are 64-bit fixed-point sums (scale 2^k from a bound on |F|: n * m_max /
    # (2 r_min)^2 * 2^k < 2^62) that meet in ONE integer all-reduce -- exact, so every rank ends with the same bits whatever
    # the deal -- and the hit pairs travel in one all_gather of {count, pairs} per rank
    bound = n * float(m.max()) / (4.0 * float(r.min()) ** 2)
    k = 61 - int(np.floor(np.log2(bound)))
    fixed = torch.from_numpy(np.rint(np.ldexp(F, k)).astype(np.int64))
    dist.all_reduce(fixed, op=dist.ReduceOp.SUM)
    total = np.ldexp(fixed.numpy().astype(np.float64), -k)
    cap = 4 * n
    mine = np.full((cap, 2), -1, dtype=np.int64)
    mine[:len(pairs)] = np.array(pairs, dtype=np.int64).reshape(-1, 2)
    Ps = [torch.zeros(cap, 2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(Ps, torch.from_numpy(mine))
    allp = np.concatenate([p.numpy() for p in Ps])
    allp = allp[allp[:, 0] >= 0]
    np.savez(os.path.join(out_dir, f"two_sided_{rank}.npz"), F=total, pairs=allp[np.lexsort((allp[:, 1], allp[:, 0]))])
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,field", [(2, 1500, 6000), (3, 2600, 9000), (2, 3000, 6000)])
def test_two_sided_deal_and_exchange_equal_the_one_sided_sum(oracle, nb, tmp_path, world, n, field):
    """world_size-2 / -3 gloo run of the sharded two-sided flow: blocks of the pair triangle dealt round-robin by the
    library's plan, every rank's fixed-point partial forces brought together by one integer all_reduce and its hit pairs
    by one all_gather.  The sum must be the all-pairs force on every body and the union of the pairs the oracle's
    event list, on every rank -- the same bits on every rank."""
    import torch.multiprocessing as mp
    mp.spawn(_two_sided_worker, args=(world, _free_port(), n, field, str(tmp_path)), nprocs=world, join=True)
    block = nb.generate(nb.SCENARIO_SQUARE, n, field_w=field, field_h=field)
    pos, _, m, r = nb.split(block, n)
    x, y = pos[:, 0].copy(), pos[:, 1].copy()
    everyone = np.arange(n)
    want_F, _, want_pairs = _pair_block(x, y, m, r, everyone, everyone, own=True)
    want_pairs = np.array(sorted(want_pairs), dtype=np.int64).reshape(-1, 2)
    par = oracle.params(field_w=field, field_h=field, coverage=oracle.COVERAGE_FULL)
    _, _, ev = oracle.step(block.copy(), n, par, want_events=True)
    ev_pairs = np.array(sorted(zip(ev["i"].tolist(), ev["j"].tolist())), dtype=np.int64).reshape(-1, 2)
    assert len(ev_pairs) > 0 and np.array_equal(want_pairs, ev_pairs), "numpy restatement of the predicate differs from the oracle"
    outs = [np.load(tmp_path / f"two_sided_{k}.npz") for k in range(world)]
    for o in outs:
        assert np.array_equal(o["pairs"], ev_pairs)
        assert np.array_equal(o["F"], outs[0]["F"]), "ranks disagree on the summed forces"
        assert np.abs(o["F"] - want_F).max() <= 1e-9 * np.abs(want_F).max()


def test_warp_level_work_items_cover_the_pair_triangle_once(nb):
    """Host logic of the warp-level two-sided kernel (nbody_symw.cu): its work-item ids -- the 2 G "own" items first, then
    the triangle in mirrored row pairs -- decode to (128-row group, run of 64-body chunks) such that every (group, chunk)
    with chunk >= 2 group is met exactly once, for any run length, and the deal to W ranks (ids r, r + W, ...) is a
    partition of the same set."""
    for n in list(range(1, 700, 13)) + [1000, 1024, 5000, 12288, 16384, 16385, 33000, 196607]:
        for run in (1, 2, 4, 8):
            if n > 40000 and run < 4:
                continue                          # (keeps the test fast: 3072 chunks x 1536 groups)
            G, Cn = (n + 127) // 128, (n + 63) // 64
            ids = nb.plan_warp_items(n, run)
            seen = set()
            owner = {}
            for item in range(ids):
                d = nb.plan_warp_item(n, run, item)
                if d is None:
                    continue
                g, lo, hi = d
                assert 0 <= g < G and 2 * g <= lo < hi <= Cn and hi - lo <= run, (n, run, item, d)
                if item < 2 * G:
                    assert (lo >> 1) == g and hi == lo + 1, "the expensive own items have the lowest ids"
                for c in range(lo, hi):
                    assert (g, c) not in seen, (n, run, g, c)
                    seen.add((g, c))
                    owner[(g, c)] = item % 3
            assert seen == {(g, c) for g in range(G) for c in range(2 * g, Cn)}, (n, run)
            if ids >= 30:
                load = [sum(1 for o in owner.values() if o == r) for r in range(3)]
                assert min(load) > 0
    with pytest.raises(nb.NbodyError):
        nb.plan_warp_item(1000, 1, nb.plan_warp_items(1000, 1))


def test_fixed_point_scale_of_the_force_sums(nb):
    """The scale 2^k of the two-sided kernels' 64-bit force sums: n m_max / (2 r_min)^2 * 2^k stays below 2^62 (no body's
    |F| can overflow), one unit is at least 30 bits below a typical force, and radii too small for that make the plan
    fall back to the one-sided kernel instead."""
    for n, m_max, r_min, field in ((1 << 20, 1e17, 50.0, 800000), (16384, 1e17, 50.0, 100000), (1 << 22, 3e18, 50.0, 3000000),
                                   (131072, 1e17, 50.0, 200000), (2000, 1.0, 1.0, 10)):
        k = nb.plan_force_scale(n, m_max, r_min, field)
        assert k is not None
        bound = n * float(np.float32(m_max)) / (4.0 * float(np.float32(r_min)) ** 2)
        assert bound * 2.0 ** k < 2.0 ** 62 and bound * 2.0 ** (k + 2) >= 2.0 ** 62, (n, k)
        typical = n * 0.5 * m_max / float(field) ** 2           # a body inside a disc of the field's size
        assert typical * 2.0 ** k >= 2.0 ** 29, "one unit must be far below a typical force"
    assert nb.plan_force_scale(1 << 20, 1e17, 0.0, 800000) is None          # r_min = 0: no bound on |F|
    assert nb.plan_force_scale(1 << 20, 1e17, 1.0, 800000) is None          # radii tiny against the field: too coarse
    assert nb.plan_force_scale(0, 1e17, 50.0, 800000) is None
    assert nb.plan_force_scale(1 << 20, 0.0, 50.0, 800000) is None
