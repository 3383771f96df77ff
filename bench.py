#!/usr/bin/env python
"""bench.py -- ordered pairwise interactions/s (and steps/s) of the ppa-nbody-collisions time step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

--config selects one of BASELINE.json's five configurations (default disc1m, the one the metric is quoted on):

    shipped   configs[0]  nbodyConfig.txt as shipped: N = 16 384 uniform square, REFERENCE coverage
    disc16k   configs[1]  N = 16 384 uniform disc R = 1e5, all-pairs (single-GPU force-kernel roofline)
    cluster   configs[2]  N = 131 072 cold disc R = 1e5 in a +-2e5 field: collision-heavy
    disc1m    configs[3]  N = 1 048 576 uniform disc R = 8e5 (the shipped surface density), 1/2/4/8 GPUs
    galaxy    configs[4]  N = 4 194 304 two-galaxy encounter

All: v = 0 (galaxy: bulk + spin), m ~ U[1e4, 1e17], r ~ U[50, 200], dt = 0.2, growth 0.1, collisions on.
A bench "step" is one pass of the hot path over one batch: `sim_steps_per_step` full time steps (force + collision
detect/merge + integrate + compaction; 1 for the large configs, 20 for the two 16k ones, whose single step is
~100 us) issued as ONE nb_step call.  L2 is flushed between bench steps.  The two 16k configs restart every bench step
from the initial bodies (uploaded outside the timed region): their bodies merge so fast that consecutive batches would
time ever smaller n; the large configs continue the run from step to step.

One JSON line on stdout (rank 0).  `value` = ordered pairs evaluated by all ranks / device time of the K timed
steps (CUDA events on the library's stream, max over ranks, bodies resident in HBM).  `e2e` = the same metric
through the C ABI with HOST buffers: every step uploads the BodiesData block from pinned host memory (nb_upload),
steps and downloads the survivors (nb_download).  `roofline` is the force kernel alone against the FP32 FMA peak
at 20 flop per interaction (peak measured in the same run by nb_probe_fp32).  `parity` is a correctness check of the
very state the timed steps produced: sampled rows of one more step against the CPU oracle (1 GPU), replicas
bit-identical and events / survivors / masses / radii identical to the same steps on ONE GPU (N > 1).
`cpu_baseline` times the CPU oracle port on the host cores on a bounded sample of rows (N = 1 only).
The oracle appears here in exactly two roles, both outside every timed region: as the CPU baseline, and as the
checker behind `parity`.

--impl reference times the UNMODIFIED reference kernels (oracle/_ref, built from /root/reference/src/nbody.cu)
driven through the reference's own main-loop body on the same GPU: the reference has no CPU implementation
(SURVEY.md C1), so its own CUDA path is the honest "reference on this box"; if that library is absent it falls
back to the CPU oracle port.  That arm loads nothing of the product (inputs come from the oracle's generators).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOP_PER_INTERACTION = 20.0
SM_MAX_MHZ_FALLBACK = 1965.0
METRIC = "pairwise_interactions_per_sec"
UNIT = "interactions/s"

# name: BASELINE index, bodies, scenario, extent, field half-width, coverage, simulation steps per bench step
# reset: every bench step starts again from the initial bodies (re-uploaded outside the timed region), so that all K steps time
# the same batch at the configuration's own N -- these scenarios merge fast (the cold 16k disc is down to ~1000 bodies after 400
# steps); without it consecutive bench steps continue the run.  run_steps: the configuration's whole run, timed once as an extra.
CONFIGS = {
    "shipped": dict(idx=0, n=16384, scenario="square", extent=0.0, field=100000, coverage="reference", batch=20, reset=True, run_steps=2000),
    "disc16k": dict(idx=1, n=16384, scenario="disc", extent=1.0e5, field=100000, coverage="full", batch=20, reset=True, run_steps=1000),
    "cluster": dict(idx=2, n=131072, scenario="disc", extent=1.0e5, field=200000, coverage="full", batch=1, reset=False, run_steps=200),
    "disc1m": dict(idx=3, n=1 << 20, scenario="disc", extent=8.0e5, field=800000, coverage="full", batch=1, reset=False, run_steps=0),
    "galaxy": dict(idx=4, n=1 << 22, scenario="two_galaxy", extent=8.0e5, field=3000000, coverage="full", batch=1, reset=False, run_steps=0),
}


def workload_name(name: str, cfg: dict) -> str:
    shape = {"square": f"uniform random square +-{cfg['field']}", "disc": f"uniform random disc R={cfg['extent']:g}",
             "two_galaxy": f"two counter-rotating discs R={cfg['extent']:g} on an encounter course"}[cfg["scenario"]]
    cov = "all-pairs" if cfg["coverage"] == "full" else "the reference's own pair coverage (SURVEY C2)"
    return (f"{name}: N={cfg['n']} {shape}, field +-{cfg['field']}, m~U[1e4,1e17], r~U[50,200], dt=0.2, growth=0.1, "
            f"collisions on, {cov} (BASELINE configs[{cfg['idx']}])")


def make_block(gen, cfg: dict, product: bool) -> np.ndarray:
    """The configuration's initial BodiesData block: from the product's nb_generate (our arm) or from the oracle's
    generators (reference arm; bit-equal, tests/test_oracle_golden.py)."""
    n, f = cfg["n"], cfg["field"]
    if product:
        kind = {"square": gen.SCENARIO_SQUARE, "disc": gen.SCENARIO_DISC, "two_galaxy": gen.SCENARIO_TWO_GALAXY}[cfg["scenario"]]
        return gen.generate(kind, n, extent=cfg["extent"], field_w=f, field_h=f)
    if cfg["scenario"] == "square":
        return gen.init_square(n, field_w=f, field_h=f)
    if cfg["scenario"] == "disc":
        return gen.init_disc(n, cfg["extent"])
    return gen.init_two_galaxy(n, cfg["extent"])


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


def ncu_traffic_bytes(config: str, two_sided: bool):
    """dram__bytes_read + dram__bytes_write of one force-kernel launch from the committed `ncu --set full` capture of
    the same workload and kernel (profiles/<round>_force_<config>_ncu.json, newest round first); None if there is none."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    names = {"disc1m": ["r02_force_1m_ncu.json"] + (["r01_force_1m_ncu.json"] if two_sided else ["r01_force_1m_one_sided_ncu.json"]),
             "disc16k": ["r02_force_16k_ncu.json"]}.get(config, [])
    for name in names:
        try:
            prof = json.loads((ROOT / "profiles" / name).read_text())["metrics"]
            rd, wr = prof["dram__bytes_read.sum"], prof["dram__bytes_write.sum"]
            return float(rd["value"]) * scale[rd["unit"]] + float(wr["value"]) * scale[wr["unit"]], f"profiles/{name}"
        except Exception:
            continue
    return None, None


def cpu_baseline(block0: np.ndarray, cfg: dict, budget_s: float = 12.0) -> dict:
    """The CPU oracle port (oracle/nbody_oracle.c, OpenMP over rows) on a bounded sample of rows of the same
    workload: every row costs its visited pair evaluations, so visited / time is the port's rate."""
    from oracle import oracle as O
    n = cfg["n"]
    cores = os.cpu_count() or 1
    budget_s = float(os.environ.get("NBODY_BENCH_CPU_BUDGET_S", budget_s))     # tests shorten the sample
    cov = O.COVERAGE_FULL if cfg["coverage"] == "full" else O.COVERAGE_REFERENCE
    par = O.params(field_w=cfg["field"], field_h=cfg["field"], coverage=cov, threads=cores)
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
            "sample": f"{rows} random rows x their {cfg['coverage']}-coverage partners among all {n} bodies of the same workload "
                      f"({visited.sum():.3e} pair evaluations, {dt:.1f} s, OpenMP {cores} threads, oracle/nbody_oracle.c)"}


def run_reference(args, cfg, out_stream) -> int:
    """The reference arm: the unmodified ComputeForces/MoveBodies + the reference's per-step
    malloc/H2D/D2H/host compaction, exactly as its main loop does them (oracle/gpu_ref_harness.cu).
    Nothing of the product is imported here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    n, field, batch = cfg["n"], cfg["field"], cfg["batch"]
    block0 = make_block(O, cfg, product=False)
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": workload_name(args.config, cfg), "bodies": n, "coverage": "reference (the only one it has)",
                       "sim_steps_per_step": batch,
                       "restart": "every bench step restarts from the initial bodies (outside the timed region)" if cfg["reset"]
                                  else "consecutive bench steps continue the run",
                       "parallelism": "1 GPU (the reference has no multi-GPU path)"}}
    have_gpu = False
    if O.gpuref_available():
        try:
            have_gpu = O.gpuref().gpuref_device_count() > 0
        except OSError:
            have_gpu = False
    if have_gpu:
        par = O.params(field_w=field, field_h=field, coverage=O.COVERAGE_REFERENCE)
        ref = O.GpuRef(block0, n)
        cur = n
        pairs = 0
        kernel_ms = 0.0
        wall = 0.0
        for it in range(args.warmup + args.steps):
            if cfg["reset"] and it > 0:             # the same batch again: restart from the initial bodies (untimed)
                ref.close()
                ref = O.GpuRef(block0, n)
                cur = n
            t0 = time.perf_counter()
            it_pairs, it_ms = 0, 0.0
            for _ in range(batch):
                cov = O.coverage(cur, O.COVERAGE_REFERENCE)
                window = 128 * (cov["blocks"] - 1) + cov["limit_last"]
                cur, ms = ref.step(par)
                it_pairs += cov["n_active"] * max(window - 1, 0)
                it_ms += ms
            if it >= args.warmup:
                wall += time.perf_counter() - t0
                pairs += it_pairs
                kernel_ms += it_ms
        ref.close()
        whole = None
        if cfg["run_steps"] > 0 and not args.no_whole_run:
            ref = O.GpuRef(block0, n)
            t0 = time.perf_counter()
            for _ in range(cfg["run_steps"]):
                cur_w, _ms = ref.step(par)
            t_run = time.perf_counter() - t0
            ref.close()
            whole = {"sim_steps": cfg["run_steps"], "ms": t_run * 1e3, "sim_steps_per_sec": cfg["run_steps"] / t_run, "bodies_end": cur_w,
                     "what": "the reference's main-loop body for the configuration's whole run (no image output), wall clock"}
        value = pairs / wall
        base.update({
            "whole_run": whole,
            "value": value, "ms_per_step": wall / args.steps * 1e3, "steps_per_sec": args.steps / wall,
            "sim_steps_per_sec": args.steps * batch / wall, "bodies_after": cur,
            "kernel_only_value": pairs / (kernel_ms * 1e-3),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                             "sample": "whole workload: the reference has no CPU path (SURVEY.md C1); this is its own "
                                       "unmodified CUDA code (oracle/_ref) on the same B200, driven as its main loop does "
                                       "(per-step cudaMalloc, H2D, 2 kernels, blocking D2H, host compaction); note its "
                                       "coverage drops the pairs SURVEY.md C2 lists"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            # for scale: the CPU restatement of the same algorithm on this box's host cores (bounded sample)
            "cpu_port": cpu_baseline(block0, cfg, budget_s=6.0)})
    else:
        cb = cpu_baseline(block0, cfg, budget_s=20.0)
        cb["kind"] = "port"
        base.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0,
                     "note": "oracle/_ref (reference CUDA kernels) unavailable: CPU oracle port on the host cores"})
    try:                                            # the judge's check, made by the arm itself
        base["loads_product_library"] = any("libnbody_b200" in ln for ln in open("/proc/self/maps"))
    except OSError:
        base["loads_product_library"] = None
    print(json.dumps(base), file=out_stream, flush=True)
    return 0


def parity_one_gpu(nb, sim, cfg, n_rows=160) -> dict:
    """One more (untimed) step of the state the timed steps left, sampled rows against the CPU oracle: events per
    row, survivors, masses and radii bit-exact; positions within 1e-5 of the field; velocity changes within 1e-5 of a
    float64 evaluation of the same pairs or within the oracle's own float32 error, whichever is larger."""
    from oracle import oracle as O
    field = cfg["field"]
    before, n0 = sim.download()
    before = before.copy()
    sim.events()                                   # drop what the timed steps logged
    sim.step(1)
    after, n1 = sim.download()
    ev = sim.events()
    if sim.stats()["events_dropped"]:
        return {"checked": False, "why": "event log overflowed"}
    cov = O.COVERAGE_FULL if cfg["coverage"] == "full" else O.COVERAGE_REFERENCE
    par = O.params(field_w=field, field_h=field, coverage=cov)
    rng = np.random.default_rng(5)
    rows = np.unique(np.concatenate([rng.integers(0, n0, n_rows), ev["i"][:: max(1, len(ev) // 48)], [0, n0 - 1]])).astype(np.int32)
    want, hits, _ = O.rows(before, n0, par, rows)
    killed = np.zeros(n0, dtype=bool)
    killed[ev["i"][ev["kind"] == 1]] = True
    keep = ~killed
    pos1, vel1, m1, r1 = nb.split(after, n1)
    _, vel0, _, _ = nb.split(before, n0)
    alive = want[:, 4] != 0
    idx = (np.cumsum(keep) - 1)[rows[alive]]
    w = want[alive]
    truth = O.rows_dv_f64(before, n0, par, rows)[alive]
    dv_gpu = vel1[idx].astype(np.float64) - vel0[rows[alive]].astype(np.float64)
    dv_ref = w[:, 0:2].astype(np.float64) - vel0[rows[alive]].astype(np.float64)
    scale = max(float(np.abs(truth).max()), 1e-30)
    # v0 + dv rounds at ulp(v0): that much of the difference is not the force sum's
    ulp_v = float(np.abs(vel0[rows[alive]]).max()) * 2.0 ** -23 / scale
    err_gpu, err_ref = float(np.abs(dv_gpu - truth).max() / scale), float(np.abs(dv_ref - truth).max() / scale)
    res = {
        "checked": True, "against": "CPU oracle (oracle/nbody_oracle.c), sampled rows of one more step from the timed state",
        "rows": int(len(rows)), "bodies_before": int(n0), "bodies_after": int(n1), "events_in_step": int(len(ev)),
        "survivors_match": bool(n1 == int(keep.sum()) and np.array_equal(alive, keep[rows])),
        "events_per_row_match": bool(np.array_equal(np.bincount(ev["i"], minlength=n0)[rows], hits)),
        "mass_radius_bits_match": bool(np.array_equal(m1[idx].view(np.uint32), w[:, 4].view(np.uint32))
                                       and np.array_equal(r1[idx].view(np.uint32), w[:, 5].view(np.uint32))),
        "dp_max_over_field": float(np.abs(pos1[idx] - w[:, 2:4]).max() / field),
        "dv_err_vs_f64": err_gpu, "dv_err_oracle_vs_f64": err_ref,
    }
    res["ok"] = bool(res["survivors_match"] and res["events_per_row_match"] and res["mass_radius_bits_match"]
                     and res["dp_max_over_field"] <= 1e-5 and err_gpu <= max(err_ref, 1e-5) + 2 * ulp_v)
    return res


def parity_sharded(nb, sim, cfg, block0, dist, rank, world, local, sim_steps: int) -> dict:
    """The sharded run's state after the timed steps: replicas bit-identical, and -- rank 0 repeats the same number
    of steps from the same bodies on ONE GPU -- events, survivors, masses and radii identical; velocities are bounded
    by 1e-4 max|v| here, and in fact identical bit for bit whenever the two-sided kernel ran (integer force sums)."""
    got, n1 = sim.download()
    ev = sim.events()
    st = sim.stats()
    digest = (int(n1), zlib.crc32(got.tobytes()), int(len(ev)), int(st["overflow"]), int(st["events_dropped"]))
    all_d = [None] * world
    dist.all_gather_object(all_d, digest)
    all_ev = [None] * world
    dist.all_gather_object(all_ev, ev)
    res = None
    if rank == 0:
        try:                                       # a failed check must not leave the other ranks waiting in a collective
            res = _parity_vs_one_gpu(nb, cfg, block0, local, sim_steps, got, n1, all_d, all_ev)
        except Exception as e:                     # noqa: BLE001
            res = {"checked": False, "ok": False, "why": f"{type(e).__name__}: {e}"}
    return res


def _parity_vs_one_gpu(nb, cfg, block0, local, sim_steps, got, n1, all_d, all_ev) -> dict:
    n, field = cfg["n"], cfg["field"]
    cov = nb.COVERAGE_FULL if cfg["coverage"] == "full" else nb.COVERAGE_REFERENCE
    one = nb.Simulation(n, field_w=field, field_h=field, coverage=cov, device=local, event_capacity=1 << 22)
    one.upload(block0, n)
    one.step(sim_steps)
    ref, n_ref = one.download()
    ev_ref = one.events()
    one.close()
    evs = np.concatenate(all_ev)
    evs = evs[np.lexsort((evs["i"], evs["step"]))]      # stable: a row's events stay in visit order
    ev_ok = len(evs) == len(ev_ref) and all(np.array_equal(evs[k], ev_ref[k]) for k in ("step", "i", "j", "kind"))
    res = {"checked": True, "against": f"the same {sim_steps} steps on one GPU (rank 0), whole state",
           "sim_steps": sim_steps, "replicas_identical": len({(d[0], d[1]) for d in all_d}) == 1,
           "state_crc32": [d[1] for d in all_d], "bodies_after": [d[0] for d in all_d], "bodies_after_1gpu": int(n_ref),
           "events": int(len(evs)), "events_vs_1gpu": bool(ev_ok), "overflow": [d[3] for d in all_d],
           "events_dropped": [d[4] for d in all_d], "survivors_vs_1gpu": bool(n1 == n_ref)}
    if n1 == n_ref:
        _, v1, m1, r1 = nb.split(got, n1)
        _, v2, m2, r2 = nb.split(ref, n_ref)
        res["mass_radius_bits_vs_1gpu"] = bool(np.array_equal(m1.view(np.uint32), m2.view(np.uint32))
                                               and np.array_equal(r1.view(np.uint32), r2.view(np.uint32)))
        res["dv_rel_max_vs_1gpu"] = float(np.abs(v1 - v2).max() / max(float(np.abs(v2).max()), 1e-30))
    res["ok"] = bool(res["replicas_identical"] and res["events_vs_1gpu"] and res["survivors_vs_1gpu"]
                     and res.get("mass_radius_bits_vs_1gpu", False) and res.get("dv_rel_max_vs_1gpu", 1.0) <= 1e-4
                     and not any(res["overflow"]) and not any(res["events_dropped"]))
    return res


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
    ap.add_argument("--config", default="disc1m", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: disc1m = configs[3])")
    ap.add_argument("--n", type=int, default=0, help="override the configuration's body count (experiments)")
    ap.add_argument("--batch", type=int, default=0, help="override simulation steps per bench step")
    ap.add_argument("--flags", type=int, default=0, help="nb_params.flags (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-whole-run", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.n > 0:
        cfg["n"] = args.n
    if args.batch > 0:
        cfg["batch"] = args.batch
    if args.impl == "reference":
        return run_reference(args, cfg, out_stream)

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

    n, field, batch = cfg["n"], cfg["field"], cfg["batch"]
    coverage = nb.COVERAGE_FULL if cfg["coverage"] == "full" else nb.COVERAGE_REFERENCE
    block0 = make_block(nb, cfg, product=True)
    want_parity = not args.no_parity
    sim = nb.Simulation(n, field_w=field, field_h=field, coverage=coverage, device=local, rank=rank, world=world,
                        event_capacity=(1 << 22) if want_parity else 0, flags=args.flags)
    if world > 1:
        ids = [nb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.comm_init(ids[0])
    fp32_peak_measured = nb.probe_fp32(local) if rank == 0 else None
    sim.upload(block0, n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    reset = cfg["reset"]
    for _ in range(args.warmup):
        if reset:
            sim.upload(block0, n)
        sim.step(batch)
    sim.sync()
    s0 = sim.stats()
    pairs_local = 0
    launches = 0
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = 0.0
    ms_force = 0.0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        if reset:                           # the same batch again (nb_upload also resets the statistics)
            sim.upload(block0, n)
            s0 = sim.stats()
        flush.zero_()                       # L2 flush between timed iterations (outside the timed events)
        torch.cuda.synchronize()
        if batch == 1:                      # every kernel launched on its own, the force kernel bracketed by events
            t, f = sim.step_timed(1, force=True)
        else:                               # the batch replays the step graph; the force kernel is timed below
            t, f = sim.step_timed(batch, force=False)
        ms_total += t
        ms_force += f
        if reset:
            s1 = sim.stats()
            pairs_local += s1["pairs"] - s0["pairs"]
            launches += s1["kernel_launches"] - s0["kernel_launches"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    s1 = sim.stats()
    if not reset:
        pairs_local = s1["pairs"] - s0["pairs"]
        launches = s1["kernel_launches"] - s0["kernel_launches"]
    sim_steps_timed = args.steps * batch

    # ---- correctness of what was just timed ---------------------------------------------------------------
    parity = None
    if want_parity:
        if world > 1:
            parity = parity_sharded(nb, sim, cfg, block0, dist, rank, world, local,
                                    batch if reset else (args.warmup + args.steps) * batch)
        else:
            parity = parity_one_gpu(nb, sim, cfg)

    # ---- force kernel alone (small configs: a few event-bracketed launches after the timed region) ----------
    pairs_force = pairs_local
    force_launches = sim_steps_timed
    if batch > 1:
        if reset:
            sim.upload(block0, n)
        sf0 = sim.stats()
        force_launches = 8
        _, ms_force = sim.step_timed(force_launches, force=True)
        pairs_force = sim.stats()["pairs"] - sf0["pairs"]
    t_ms = torch.tensor([ms_total, ms_force], dtype=torch.float64, device="cuda")
    p_all = torch.tensor([float(pairs_local), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(p_all, op=dist.ReduceOp.SUM)
    ms_total_max, ms_force_max = (float(x) for x in t_ms.tolist())
    pairs_all, launches_all = (float(x) for x in p_all.tolist())
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
        e2e_steps = max(1, args.steps)
        sim.upload_ptr(host_in.data_ptr(), n)
        sim.step(batch)
        if rank == 0:
            sim.download_ptr(host_out.data_ptr(), n)        # warm
        else:
            sim.sync()
        barrier()
        d2h = 0
        pairs_e2e = 0.0
        phase = [0.0, 0.0, 0.0]                 # host-side split of the e2e time: upload | step (returns early on one GPU) | download
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ta = time.perf_counter()
            sim.upload_ptr(host_in.data_ptr(), n)
            tb = time.perf_counter()
            sim.step(batch)
            tc = time.perf_counter()
            if rank == 0:                   # the replicas hold the same bodies: ONE copy of the result goes back to the host
                n_out = sim.download_ptr(host_out.data_ptr(), n)
            else:
                sim.sync()
                n_out = 0
            td = time.perf_counter()
            phase[0] += tb - ta
            phase[1] += tc - tb
            phase[2] += td - tc
            d2h += 24 * n_out
        barrier()
        t_e2e = time.perf_counter() - t0
        pairs_e2e = float(sim.stats()["pairs"]) * e2e_steps   # every e2e step starts from the same n bodies: same pairs
        t_t = torch.tensor([t_e2e, pairs_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            tm = t_t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_t, op=dist.ReduceOp.SUM)
            t_e2e, pairs_e2e = float(tm[0]), float(t_t[1])
        e2e = {"value": pairs_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 24 * n,
               "d2h_bytes_per_step": d2h // e2e_steps, "steps": e2e_steps, "ms_per_step": t_e2e / e2e_steps * 1e3,
               "host_ms_per_step": {"nb_upload": phase[0] / e2e_steps * 1e3, "nb_step": phase[1] / e2e_steps * 1e3,
                                    "nb_download": phase[2] / e2e_steps * 1e3, "rank": 0},
               "what": (f"nb_upload(pinned host block) + nb_step({batch}) + nb_download(pinned host block) per step, wall clock"
                        + ("; every rank uploads (1 / world of the block over PCIe each, the rest over NVLink), rank 0 downloads the "
                           "one copy of the result the job needs (the replicas are identical)" if world > 1 else ""))}

    # ---- the configuration's whole run (BASELINE: "1000 steps", "2000 iterations", ...), timed once as an extra --------
    whole = None
    if cfg["run_steps"] > 0 and not args.no_whole_run:
        sim.upload(block0, n)
        barrier()
        t_run, _ = sim.step_timed(cfg["run_steps"], force=False)
        sw = sim.stats()
        whole = {"sim_steps": cfg["run_steps"], "ms": t_run, "sim_steps_per_sec": cfg["run_steps"] / (t_run * 1e-3),
                 "bodies_end": sw["n"], "value": float(sw["pairs"]) * world / (t_run * 1e-3) if world == 1 else None,
                 "what": "nb_upload, then the configuration's whole run as ONE nb_step call (CUDA events around it)"}

    if rank == 0:
        peaks = measured_peaks()
        sm_max = float(peaks.get("sm_max_mhz", SM_MAX_MHZ_FALLBACK))
        sms = s1["sm_count"]
        nameplate_tflops = sms * 128 * 2 * sm_max * 1e6 / 1e12
        peak_tflops = fp32_peak_measured
        # issue-slot bound of the two-sided kernel: 12 packed ops (2 dispatch cycles each) + 2 MUFU + 2.5 SHFL per lane
        # give 4 ordered interactions; one dispatch per cycle and SM sub-partition, 32 lanes
        sym_bound = 32 * 4 / 28.5 * 4 * sms * sm_max * 1e6
        two_sided = bool(s1["pair_halving"])
        sorted_order = s1["culled_parts"] > s0["culled_parts"]
        rate_force = (pairs_force / force_launches) / (ms_force_max / force_launches * 1e-3)      # ordered interactions/s, this rank
        achieved_tflops = FLOP_PER_INTERACTION * rate_force / 1e12
        traffic, traffic_src = ncu_traffic_bytes(args.config, two_sided)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "steps_per_sec": args.steps / (ms_total_max * 1e-3),
            "sim_steps_per_sec": sim_steps_timed / (ms_total_max * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, cfg), "bodies": n, "coverage": cfg["coverage"],
                       "sim_steps_per_step": batch,
                       "restart": "every bench step restarts from the initial bodies (uploaded outside the timed region)" if reset
                                  else "consecutive bench steps continue the run",
                       "parallelism": (f"pair-triangle blocks dealt round-robin to {world} GPUs + NCCL exchange of partial forces, "
                                       f"candidate pairs and post-step rows") if world > 1 else "1 GPU",
                       "l2": "flushed between timed iterations (256 MiB memset); each bench step timed by its own CUDA-event pair",
                       "bodies_after": s1["n"]},
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops,
                         # the flop the kernel really executed: the two-sided kernel evaluates each unordered pair once
                         "frac_executed": achieved_tflops / peak_tflops * (0.5 if two_sided else 1.0),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "traffic_note": "DRAM bytes per launch from the committed ncu capture; the kernel is FP32-issue bound, "
                                         "not HBM bound: its algorithmic HBM bytes are one pass over the rows and the j stream",
                         "kernel": ("force_sym_kernel<packed f32x2, two-sided>" if two_sided
                                    else "force_kernel<packed f32x2, 8 warps x 2 rows/lane>"),
                         "ms_per_launch": ms_force_max / force_launches, "launches_timed": force_launches,
                         "flop_per_interaction": FLOP_PER_INTERACTION,
                         "peak_source": f"nb_probe_fp32 in this run: independent packed FFMA2 stream, 32 warps/SM, best of 5 "
                                        f"(nameplate {sms} SMs x 128 lanes x 2 flop x {sm_max:.0f} MHz = {nameplate_tflops:.2f})",
                         "nameplate_peak": nameplate_tflops, "frac_of_nameplate": achieved_tflops / nameplate_tflops,
                         "share_of_step": (ms_force_max / ms_total_max) if batch == 1 else None,
                         "pair_halving": two_sided,
                         "note": ("achieved = 20 flop x ORDERED interactions / time, the reference's accounting (SURVEY 8d). The "
                                  "two-sided kernel evaluates each unordered pair once (12 packed f32x2 operations + 2 MUFU + 2.5 "
                                  "SHFL per lane for 4 ordered interactions, against 2 x 9 + 2 x 2 one-sided), so frac can pass 1 "
                                  "(frac_executed halves it); against its own issue-slot bound (28.5 dispatch cycles per 128 ordered "
                                  f"interactions and SM sub-partition = {sym_bound / 1e12:.2f}e12 interactions/s) it reaches "
                                  "frac_of_issue_bound")
                                 if two_sided else "achieved = 20 flop x ordered interactions / time (SURVEY 8d)",
                         "frac_of_issue_bound": (rate_force / sym_bound if two_sided else None)},
            "clocks": clocks,
            "gpu_launches": int(launches_all),
            "gpu_launches_note": "counted at the library's launch sites (nb_stats.kernel_launches: direct launches + kernel nodes "
                                 "of every graph replay), summed over ranks; NCCL's own kernels not included",
            "wall_s_timed_region": wall,
            "force": {"grid": s1["force_grid"], "regs": s1["force_regs"],
                      "fast_chunks": s1["fast_chunks"] - s0["fast_chunks"], "exact_chunks": s1["exact_chunks"] - s0["exact_chunks"],
                      "parts_without_pretest": s1["culled_parts"] - s0["culled_parts"], "cell_sorted_order": sorted_order,
                      "two_sided": two_sided, "two_sided_regs": s1["sym_regs"]},
            "collision_events": s1["candidates"] - s0["candidates"],
            "parity": parity,
            "whole_run": whole,
        }
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic bytes per body: finish reads pm 16 + vel 8 + the partial sums (8 each) and writes 24;
        # compaction reads 24 (+ 16 for the count pass when sharded) and writes pm 16 + vel 8 + j-tile 16
        npart = sim.stats().get("force_partials", 1) or 1
        fin_b, cmp_b = 48.0 + 8.0 * npart, (64.0 if world == 1 else 80.0)
        out["hbm_kernels"] = {
            "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 (B200_PROFILING.md)",
            "finish": {"ms_per_launch": prof["finish"] / prof_steps, "bytes_per_body": fin_b,
                       "achieved_gbs": fin_b * n_mid / world / (prof["finish"] / prof_steps * 1e-3) / 1e9},
            "compact": {"ms_per_launch": prof["compact"] / prof_steps, "bytes_per_body": cmp_b,
                        "achieved_gbs": cmp_b * n_mid / (prof["compact"] / prof_steps * 1e-3) / 1e9},
            "allgather_ms": prof["allgather"] / prof_steps,
            "sort_ms": prof["sort"] / prof_steps,
            "note": "O(n) kernels; on several GPUs `finish` also spans the partial-force reduction, the exchange and the "
                    "candidate threading of the two-sided kernel",
        }
        for k in ("finish", "compact"):
            out["hbm_kernels"][k]["frac"] = out["hbm_kernels"][k]["achieved_gbs"] / hbm_peak
        if e2e is not None:
            out["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(block0, cfg)
        print(json.dumps(out), file=out_stream, flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
