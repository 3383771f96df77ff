#!/usr/bin/env python
"""bench.py -- ordered pairwise interactions/s (and steps/s) of the ppa-nbody-collisions time step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n BODIES]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the one the metric is quoted on): N = 1,048,576 bodies, uniform
random disc of radius 8e5 in a +-8e5 field (the shipped scenario's surface density), v = 0,
m ~ U[1e4, 1e17], r ~ U[50, 200], dt = 0.2, growth 0.1, collisions on, true all-pairs coverage.
A "step" is one full time step: force + collision detect/merge + integrate + compaction
(+ one NCCL allgather of the post-step rows when sharded over N GPUs; strong scaling: N is fixed).

One JSON line on stdout (rank 0).  `value` = ordered pairs evaluated by all ranks / device time of
the K timed steps (CUDA events on the library's stream, max over ranks, bodies resident in HBM).
`e2e` = the same metric through the C ABI with HOST buffers: every step uploads the BodiesData
block from pinned host memory (nb_upload), steps once and downloads the survivors (nb_download).
`roofline` is the force kernel alone against the FP32 FMA peak at 20 flop per interaction.
`cpu_baseline` times the CPU oracle port on the host cores on a bounded sample of rows (N = 1 only).

--impl reference times the UNMODIFIED reference kernels (oracle/_ref, built from
/root/reference/src/nbody.cu) driven through the reference's own main-loop body on the same GPU:
the reference has no CPU implementation, so its own CUDA path is the honest
"reference on this box"; if that library is absent it falls back to the CPU oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_BODIES = 1 << 20
DISC_R = 8.0e5
FIELD = 800000
FLOP_PER_INTERACTION = 20.0
SM_MAX_MHZ_FALLBACK = 1965.0
METRIC = "pairwise_interactions_per_sec"
UNIT = "interactions/s"


def workload_name(n):
    return (f"N={n} uniform random disc R={DISC_R:g} field +-{FIELD}, v=0, m~U[1e4,1e17], r~U[50,200], dt=0.2, "
            f"growth=0.1, collisions on, all-pairs (BASELINE configs[3])")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return {}


def ncu_traffic_bytes(n: int):
    """dram__bytes_read + dram__bytes_write of one force-kernel launch from the committed ncu capture of the same
    workload (profiles/r01_force_1m_ncu.json); None for other sizes."""
    try:
        prof = json.loads((ROOT / "profiles" / "r01_force_1m_ncu.json").read_text())["metrics"]
        if n != N_BODIES:
            return None
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = prof["dram__bytes_read.sum"], prof["dram__bytes_write.sum"]
        return float(rd["value"]) * scale[rd["unit"]] + float(wr["value"]) * scale[wr["unit"]]
    except Exception:
        return None


def cpu_baseline(block0: np.ndarray, n: int, budget_s: float = 12.0) -> dict:
    """The CPU oracle port (oracle/nbody_oracle.c, OpenMP over rows) on a bounded sample of rows of the
    same workload: every row costs n-1 pair evaluations, so rows x (n-1) / time is the port's rate."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    budget_s = float(os.environ.get("NBODY_BENCH_CPU_BUDGET_S", budget_s))     # tests shorten the sample
    par = O.params(field_w=FIELD, field_h=FIELD, coverage=O.COVERAGE_FULL, threads=cores)
    rng = np.random.default_rng(1)
    probe = np.sort(rng.choice(n, size=min(n, 64 * cores), replace=False)).astype(np.int32)
    t0 = time.perf_counter()
    O.rows(block0, n, par, probe)
    t_probe = max(time.perf_counter() - t0, 1e-6)
    rows = int(min(n, max(len(probe), len(probe) * budget_s / t_probe)))
    rows = max(cores, rows // cores * cores)
    sample = np.sort(rng.choice(n, size=rows, replace=False)).astype(np.int32)
    t0 = time.perf_counter()
    _, _, visited = O.rows(block0, n, par, sample)
    dt = time.perf_counter() - t0
    return {"value": float(visited.sum()) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{rows} random rows x all {n} bodies of the same workload ({visited.sum():.3e} pair evaluations, "
                      f"{dt:.1f} s, OpenMP {cores} threads, oracle/nbody_oracle.c)"}


def run_reference(args, out_stream) -> int:
    """The reference arm: the unmodified ComputeForces/MoveBodies + the reference's per-step
    malloc/H2D/D2H/host compaction, exactly as its main loop does them (oracle/gpu_ref_harness.cu)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    import __graft_entry__ as G
    nb = G.load_package()
    n = args.n
    block0 = nb.generate(nb.SCENARIO_DISC, n, extent=DISC_R, field_w=FIELD, field_h=FIELD)
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "config": {"workload": workload_name(n)}}
    if O.gpuref_available():
        try:
            O.gpuref()
            have_gpu = O.gpuref().gpuref_device_count() > 0
        except OSError:
            have_gpu = False
    else:
        have_gpu = False
    if have_gpu:
        par = O.params(field_w=FIELD, field_h=FIELD, coverage=O.COVERAGE_REFERENCE)
        ref = O.GpuRef(block0, n)
        cur = n
        pairs = 0
        kernel_ms = 0.0
        t0 = None
        for s in range(args.warmup + args.steps):
            if s == args.warmup:
                t0 = time.perf_counter()
            cov = O.coverage(cur, O.COVERAGE_REFERENCE)
            window = 128 * (cov["blocks"] - 1) + cov["limit_last"]
            cur, ms = ref.step(par)
            if s >= args.warmup:
                pairs += cov["n_active"] * max(window - 1, 0)
                kernel_ms += ms
        wall = time.perf_counter() - t0
        ref.close()
        value = pairs / wall
        base.update({
            "value": value, "ms_per_step": wall / args.steps * 1e3, "steps_per_sec": args.steps / wall,
            "kernel_only_value": pairs / (kernel_ms * 1e-3),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                             "sample": "whole workload: the reference has no CPU path (SURVEY.md C1); this is its own "
                                       "unmodified CUDA code (oracle/_ref) on the same B200, driven as its main loop does "
                                       "(per-step cudaMalloc, H2D, 2 kernels, blocking D2H, host compaction); note its "
                                       "coverage drops the pairs SURVEY.md C2 lists"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            # for scale: the CPU restatement of the same algorithm on this box's host cores (bounded sample)
            "cpu_port": cpu_baseline(block0, n, budget_s=6.0)})
    else:
        cb = cpu_baseline(block0, n, budget_s=20.0)
        cb["kind"] = "port"
        base.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0,
                     "note": "oracle/_ref (reference CUDA kernels) unavailable: CPU oracle port on the host cores"})
    print(json.dumps(base), file=out_stream, flush=True)
    return 0


def _claim_stdout():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the returned
    file object is the real stdout, used for the one JSON line."""
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main() -> int:
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_BODIES, help="bodies (default: the BASELINE workload, 1048576)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, out_stream)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as G
    nb = G.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (one process per GPU)", file=sys.stderr)
        return 2
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; this framework has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    block0 = nb.generate(nb.SCENARIO_DISC, n, extent=DISC_R, field_w=FIELD, field_h=FIELD)
    sim = nb.Simulation(n, field_w=FIELD, field_h=FIELD, coverage=nb.COVERAGE_FULL, device=local, rank=rank, world=world)
    if world > 1:
        ids = [nb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.comm_init(ids[0])
    sim.upload(block0, n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    for _ in range(args.warmup):
        sim.step(1)
    sim.sync()
    s0 = sim.stats()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = 0.0
    ms_force = 0.0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the timed events)
        torch.cuda.synchronize()
        t, f = sim.step_timed(1, force=True)
        ms_total += t
        ms_force += f
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    s1 = sim.stats()
    pairs_local = s1["pairs"] - s0["pairs"]
    t_ms = torch.tensor([ms_total, ms_force], dtype=torch.float64, device="cuda")
    p_all = torch.tensor([float(pairs_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(p_all, op=dist.ReduceOp.SUM)
    ms_total_max, ms_force_max = (float(x) for x in t_ms.tolist())
    pairs_all = float(p_all.item())
    value = pairs_all / (ms_total_max * 1e-3)

    # ---- the O(n) kernels against the HBM roofline (2 extra untimed-for-the-metric steps) --------------
    prof_steps = 2
    n_before = sim.stats()["n"]
    prof = sim.step_profile(prof_steps)
    n_after = sim.stats()["n"]
    n_mid = 0.5 * (n_before + n_after)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------------
    e2e = None
    if not args.no_e2e:
        host_in = torch.from_numpy(block0.copy()).pin_memory()
        host_out = torch.empty(6 * n, dtype=torch.float32).pin_memory()
        e2e_steps = max(1, min(args.steps, 3))
        sim.upload_ptr(host_in.data_ptr(), n)
        sim.step(1)
        sim.download_ptr(host_out.data_ptr(), n)            # warm
        barrier()
        d2h = 0
        pairs_e2e = 0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sim.upload_ptr(host_in.data_ptr(), n)
            sim.step(1)
            n_out = sim.download_ptr(host_out.data_ptr(), n)
            d2h += 24 * n_out
        barrier()
        t_e2e = time.perf_counter() - t0
        pairs_e2e = float(n) * (n - 1) * e2e_steps            # every e2e step starts from the same n bodies
        t_t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_t, op=dist.ReduceOp.MAX)
        e2e = {"value": pairs_e2e / float(t_t.item()), "unit": UNIT, "h2d_bytes_per_step": 24 * n,
               "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps,
               "what": "nb_upload(pinned host block) + nb_step(1) + nb_download(pinned host block) per step, wall clock"}

    if rank == 0:
        peaks = measured_peaks()
        sm_max = float(peaks.get("sm_max_mhz", SM_MAX_MHZ_FALLBACK))
        sms = s1["sm_count"]
        peak_tflops = sms * 128 * 2 * sm_max * 1e6 / 1e12
        # issue-slot bound of the two-sided kernel: 12 packed ops (2 dispatch cycles each) + 2 MUFU + 2.5 SHFL per lane
        # give 4 ordered interactions; one dispatch per cycle and SM sub-partition, 32 lanes
        sym_bound = 32 * 4 / 28.5 * 4 * sms * sm_max * 1e6
        two_sided = bool(s1["pair_halving"])
        sorted_order = s1["culled_parts"] > s0["culled_parts"]
        achieved_tflops = FLOP_PER_INTERACTION * (pairs_local / args.steps) / (ms_force_max / args.steps * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "steps_per_sec": args.steps / (ms_total_max * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n), "bodies": n, "coverage": "full",
                       "parallelism": (f"pair-triangle blocks dealt round-robin to {world} GPUs + NCCL allgather of partial forces, "
                                       f"candidate pairs and post-step rows") if world > 1 else "1 GPU",
                       "l2": "flushed between timed iterations (256 MiB memset); each step timed by its own CUDA-event pair",
                       "bodies_after": s1["n"]},
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops, "traffic": ncu_traffic_bytes(n),
                         "traffic_note": "DRAM bytes per launch (ncu, profiles/r01_force_1m_ncu.json); algorithmic HBM bytes of "
                                         "the launch are 32 n (one pass over the 16 B/body rows and sorted tiles) plus, for the "
                                         "two-sided kernel, one write of the per-super-tile partial sums (8 B x 256 per body): "
                                         "the kernel is FP32-issue bound, not HBM bound",
                         "kernel": ("force_sym_kernel<packed f32x2, 8 warps x 4 rows/lane, two-sided>" if two_sided
                                    else "force_kernel<packed f32x2, 8 warps x 2 rows/lane>"),
                         "ms_per_launch": ms_force_max / args.steps,
                         "flop_per_interaction": FLOP_PER_INTERACTION,
                         "peak_source": f"nameplate FP32 FMA: {sms} SMs x 128 lanes x 2 flop x {sm_max:.0f} MHz "
                                        f"(MEASURED_PEAKS.json has no FP32 entry; FFMA probe measured 73.9 TFLOP/s, "
                                        f"profiles/r01_fp32_probe.jsonl)",
                         "share_of_step": ms_force_max / ms_total_max,
                         "measured_ffma_peak": 73.9, "frac_of_measured_ffma_peak": achieved_tflops / 73.9,
                         "pair_halving": two_sided,
                         "note": ("achieved = 20 flop x ORDERED interactions / time, the reference's accounting (SURVEY 8d). The "
                                  "two-sided kernel evaluates each unordered pair once (12 packed f32x2 operations + 2 MUFU + 2.5 "
                                  "SHFL per lane for 4 ordered interactions, against 2 x 9 + 2 x 2 one-sided), so frac can pass 1; "
                                  "against its own issue-slot bound (28.5 dispatch cycles per 128 ordered interactions and SM "
                                  f"sub-partition = {sym_bound / 1e12:.2f}e12 interactions/s) it reaches frac_of_issue_bound")
                                 if two_sided else "achieved = 20 flop x ordered interactions / time (SURVEY 8d)",
                         "frac_of_issue_bound": ((pairs_local / args.steps) / (ms_force_max / args.steps * 1e-3) / sym_bound
                                                 if two_sided else None)},
            "clocks": clocks,
            # force, finish, scatter; + count when sharded; + 8 kernels that rebuild the cell-sorted order and the second
            # force kernel of a sort-capable step (the one the step does not use returns at once); + partial-force
            # reduction and candidate threading around the exchange of the sharded two-sided kernel
            "gpu_launches": (3 + (1 if world > 1 else 0) + (8 if sorted_order else 0) + (1 if two_sided else 0)
                             + (2 if two_sided and world > 1 else 0)) * args.steps,
            "wall_s_timed_region": wall,
            "force": {"grid": s1["force_grid"], "regs": s1["force_regs"],
                      "fast_chunks": s1["fast_chunks"] - s0["fast_chunks"], "exact_chunks": s1["exact_chunks"] - s0["exact_chunks"],
                      "parts_without_pretest": s1["culled_parts"] - s0["culled_parts"], "cell_sorted_order": sorted_order,
                      "two_sided": two_sided, "two_sided_regs": s1["sym_regs"]},
            "collision_events": s1["candidates"] - s0["candidates"],
        }
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic bytes per body: finish reads pm 16 + vel 8 + one partial-sum slab 8 (one-sided) and writes 24;
        # compaction reads 24 (+ 16 for the count pass when sharded) and writes pm 16 + vel 8 + j-tile 16
        # the two-sided kernel leaves one 8-B partial per super-tile (one GPU) or per rank (after the exchange)
        tiles = (n + 511) // 512
        qmax = 256 if world == 1 else 512
        sym_S = (tiles + qmax - 1) // qmax
        sym_Q = (tiles + sym_S - 1) // sym_S
        fin_b, cmp_b = 56.0 + (8.0 * (sym_Q if world == 1 else world) - 8.0 if two_sided else 0.0), (64.0 if world == 1 else 80.0)
        out["hbm_kernels"] = {
            "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)",
            "finish": {"ms_per_launch": prof["finish"] / prof_steps, "bytes_per_body": fin_b,
                       "achieved_gbs": fin_b * n_mid / world / (prof["finish"] / prof_steps * 1e-3) / 1e9},
            "compact": {"ms_per_launch": prof["compact"] / prof_steps, "bytes_per_body": cmp_b,
                        "achieved_gbs": cmp_b * n_mid / (prof["compact"] / prof_steps * 1e-3) / 1e9},
            "allgather_ms": prof["allgather"] / prof_steps,
            "sort_ms": prof["sort"] / prof_steps,
            "note": "O(n) kernels, < 0.2 % of the step at this n; on several GPUs `finish` also spans the partial-force "
                    "reduction, the exchange (NCCL allgather) and the candidate threading of the two-sided kernel",
        }
        for k in ("finish", "compact"):
            out["hbm_kernels"][k]["frac"] = out["hbm_kernels"][k]["achieved_gbs"] / hbm_peak
        if e2e is not None:
            out["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(block0, n)
        print(json.dumps(out), file=out_stream, flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
