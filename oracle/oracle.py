"""ctypes binding of the CPU oracle (oracle/nbody_oracle.c) and of the
reference-kernel harness (oracle/_ref/libnbody_gpuref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libnbody_oracle.so"
GPUREF_SO = HERE / "_ref" / "libnbody_gpuref.so"

COVERAGE_REFERENCE = 0
COVERAGE_FULL = 1
EV_ABSORB = 0
EV_KILLED = 1


class OrcParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("growth", C.c_float), ("field_w", C.c_int),
                ("field_h", C.c_int), ("coverage", C.c_int), ("threads", C.c_int), ("softening", C.c_float), ("merge", C.c_int)]


class OrcCov(C.Structure):
    _fields_ = [("n", C.c_int), ("blocks", C.c_int), ("limit_last", C.c_int), ("n_active", C.c_int)]


class OrcRng(C.Structure):
    _fields_ = [("u", C.c_uint64), ("v", C.c_uint64), ("w", C.c_uint64)]


EVENT_DTYPE = np.dtype([("i", np.int32), ("j", np.int32), ("kind", np.int32)])


def build(force: bool = False) -> None:
    """Compile the C restatement (always) and, when /root/reference is present,
    the reference-kernel harness.  Building the checker is not using it."""
    src = HERE / "nbody_oracle.c"
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)
    ref_root = Path(os.environ.get("NBODY_REFERENCE", "/root/reference"))
    if (ref_root / "src" / "nbody.cu").exists():
        harness = HERE / "gpu_ref_harness.cu"
        if force or not GPUREF_SO.exists() or GPUREF_SO.stat().st_mtime < harness.stat().st_mtime:
            subprocess.run(["make", "-C", str(HERE), "ref", f"REFERENCE={ref_root}"],
                           check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(ORACLE_SO))
        fp = C.POINTER(C.c_float)
        L.orc_rng_seed.argtypes = [C.POINTER(OrcRng), C.c_uint64]
        L.orc_rng_ival64.argtypes = [C.POINTER(OrcRng)]
        L.orc_rng_ival64.restype = C.c_uint64
        L.orc_rng_fval.argtypes = [C.POINTER(OrcRng)]
        L.orc_rng_fval.restype = C.c_double
        L.orc_init_square.argtypes = [fp, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                      C.c_float, C.c_float, C.c_float, C.c_float]
        L.orc_init_disc_part.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_uint64] + [C.c_double] * 6 + [C.c_float] * 4
        L.orc_coverage.argtypes = [C.c_int, C.c_int, C.POINTER(OrcCov)]
        L.orc_rows.argtypes = [fp, C.c_int, C.POINTER(OrcParams), C.POINTER(C.c_int), C.c_int,
                               fp, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
        L.orc_rows_dv_f64.argtypes = [fp, C.c_int, C.POINTER(OrcParams), C.POINTER(C.c_int), C.c_int,
                                      C.POINTER(C.c_double)]
        L.orc_render.argtypes = [fp, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_step.argtypes = [fp, C.c_int, C.POINTER(OrcParams), C.c_void_p, C.c_longlong,
                               C.POINTER(C.c_longlong)]
        L.orc_step.restype = C.c_int
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def params(dt=0.2, growth=0.1, field_w=100000, field_h=100000, coverage=COVERAGE_REFERENCE, threads=0, softening=0.0, merge=0):
    return OrcParams(np.float32(dt), np.float32(growth), int(field_w), int(field_h), int(coverage), int(threads),
                     np.float32(softening), int(merge))


def coverage(n: int, mode: int) -> dict:
    c = OrcCov()
    lib().orc_coverage(n, mode, C.byref(c))
    return {"n": c.n, "blocks": c.blocks, "limit_last": c.limit_last, "n_active": c.n_active}


def rng_stream(seed: int, count: int):
    g = OrcRng()
    lib().orc_rng_seed(C.byref(g), seed)
    return [lib().orc_rng_ival64(C.byref(g)) for _ in range(count)]


def rng_fvals(seed: int, count: int):
    g = OrcRng()
    lib().orc_rng_seed(C.byref(g), seed)
    return [lib().orc_rng_fval(C.byref(g)) for _ in range(count)]


def init_square(n, seed=1024, field_w=100000, field_h=100000, min_mass=1e4, max_mass=1e17,
                min_radius=50.0, max_radius=200.0) -> np.ndarray:
    """BodiesData block (6*n float32): pos[n][2], vel[n][2], mass[n], radius[n]."""
    block = np.zeros(6 * n, dtype=np.float32)
    lib().orc_init_square(_fptr(block), n, seed, field_w, field_h,
                          np.float32(min_mass), np.float32(max_mass),
                          np.float32(min_radius), np.float32(max_radius))
    return block


def init_disc(n, extent, seed=1024, min_mass=1e4, max_mass=1e17, min_radius=50.0, max_radius=200.0) -> np.ndarray:
    """Uniform disc of radius `extent`, v = 0 (BASELINE configs[1..3]); bit-equal to the product's nb_generate."""
    block = np.zeros(6 * n, dtype=np.float32)
    lib().orc_init_disc_part(_fptr(block), n, 0, n, seed, 0.0, 0.0, float(extent), 0.0, 0.0, 0.0,
                             np.float32(min_mass), np.float32(max_mass), np.float32(min_radius), np.float32(max_radius))
    return block


def init_two_galaxy(n, extent, seed=1024, min_mass=1e4, max_mass=1e17, min_radius=50.0, max_radius=200.0) -> np.ndarray:
    """Two counter-rotating discs on an encounter course (BASELINE configs[4]); bit-equal to nb_generate."""
    block = np.zeros(6 * n, dtype=np.float32)
    R, n0 = float(extent), n // 2
    mr = (np.float32(min_mass), np.float32(max_mass), np.float32(min_radius), np.float32(max_radius))
    lib().orc_init_disc_part(_fptr(block), n, 0, n0, seed, -1.5 * R, -0.25 * R, R, 400.0, 0.0, 2.0e-4, *mr)
    lib().orc_init_disc_part(_fptr(block), n, n0, n - n0, seed + 1, 1.5 * R, 0.25 * R, R, -400.0, 0.0, -2.0e-4, *mr)
    return block


def split(block: np.ndarray, n: int):
    """Views (pos[n,2], vel[n,2], mass[n], radius[n]) of a BodiesData block."""
    return (block[:2 * n].reshape(n, 2), block[2 * n:4 * n].reshape(n, 2),
            block[4 * n:5 * n], block[5 * n:6 * n])


def step(block: np.ndarray, n: int, par: OrcParams, want_events: bool = False):
    """One step in place.  Returns (new_n, stats dict, events or None)."""
    stats = (C.c_longlong * 3)()
    ev = None
    if want_events:
        cap = max(16, 8 * n)
        while True:
            work = block[:6 * n].copy()
            ev = np.zeros(cap, dtype=EVENT_DTYPE)
            new_n = lib().orc_step(_fptr(work), n, C.byref(par), ev.ctypes.data, cap, stats)
            if stats[1] <= cap:
                block[:6 * n] = work
                ev = ev[:stats[1]]
                break
            cap = int(stats[1])
    else:
        new_n = lib().orc_step(_fptr(block), n, C.byref(par), None, 0, stats)
    return new_n, {"pairs": int(stats[0]), "events": int(stats[1]), "asserts": int(stats[2])}, ev


def rows(block: np.ndarray, n: int, par: OrcParams, row_idx):
    """Post-step (vx, vy, px, py, m, r) of selected rows; returns (out[nrows,6], hits, visited)."""
    idx = np.ascontiguousarray(row_idx, dtype=np.int32)
    out = np.zeros((len(idx), 6), dtype=np.float32)
    hits = np.zeros(len(idx), dtype=np.int32)
    visited = np.zeros(len(idx), dtype=np.int64)
    lib().orc_rows(_fptr(block), n, C.byref(par), idx.ctypes.data_as(C.POINTER(C.c_int)), len(idx),
                   _fptr(out.reshape(-1)), hits.ctypes.data_as(C.POINTER(C.c_int)),
                   visited.ctypes.data_as(C.POINTER(C.c_longlong)))
    return out, hits, visited


def rows_dv_f64(block: np.ndarray, n: int, par: OrcParams, row_idx) -> np.ndarray:
    """Float64 yardstick: dv = dt * G * force of selected rows with the sum evaluated in double
    (same visited pairs, same float32 collision predicate).  Returns [nrows, 2] float64."""
    idx = np.ascontiguousarray(row_idx, dtype=np.int32)
    out = np.zeros((len(idx), 2), dtype=np.float64)
    lib().orc_rows_dv_f64(_fptr(block), n, C.byref(par), idx.ctypes.data_as(C.POINTER(C.c_int)), len(idx),
                          out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def render(block: np.ndarray, n: int, width: int, height: int, field_w: int, field_h: int, grid_n=None) -> np.ndarray:
    """generateImage (src/nbody.cu:294-348) for the n live bodies: uint8[height, width], background 254, bodies 0.
    grid_n: the body count before the step just done -- the reference's loop draws with that step's grid, i.e. only the
    first 128 * max(1, grid_n // 128) bodies (src/nbody.cu:473,535); None draws all."""
    img = np.zeros((height, width), dtype=np.uint8)
    drawn = n if grid_n is None else 128 * max(1, grid_n // 128)
    lib().orc_render(_fptr(np.ascontiguousarray(block[:6 * n])), n, drawn, img.ctypes.data, width, height, field_w, field_h)
    return img


def fnv(a: np.ndarray, h: int = 0) -> int:
    a = np.ascontiguousarray(a)
    return int(lib().orc_fnv1a64(a.ctypes.data, a.nbytes, h))


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---------------------------------------------------------------------------
# Reference kernels (GPU only): oracle/_ref/libnbody_gpuref.so
# ---------------------------------------------------------------------------
_gref = None


def gpuref_available() -> bool:
    return GPUREF_SO.exists()


def gpuref() -> C.CDLL:
    global _gref
    if _gref is None:
        G = C.CDLL(str(GPUREF_SO))
        fp = C.POINTER(C.c_float)
        G.gpuref_open.argtypes = [fp, C.c_int]
        G.gpuref_read.argtypes = [fp]
        G.gpuref_step.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, fp]
        G.gpuref_time_kernels.argtypes = [fp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                          C.c_int, C.c_int, fp]
        G.gpuref_main_in.argtypes = [C.c_char_p]
        if hasattr(G, "gpuref_render"):
            G.gpuref_render.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _gref = G
    return _gref


class GpuRef:
    """The unmodified reference kernels driven step by step (needs a GPU)."""

    def __init__(self, block: np.ndarray, n: int):
        rc = gpuref().gpuref_open(_fptr(np.ascontiguousarray(block[:6 * n])), n)
        if rc != 0:
            raise RuntimeError(f"gpuref_open failed: {rc}")

    def step(self, par: OrcParams):
        ms = C.c_float(0)
        n = gpuref().gpuref_step(par.dt, par.growth, par.field_w, par.field_h, C.byref(ms))
        if n < 0:
            raise RuntimeError(f"gpuref_step failed: {n}")
        return n, float(ms.value)

    def render(self, width: int, height: int, field_w: int, field_h: int, grid_n: int) -> np.ndarray:
        """The reference's generateImage of the current bodies, launched with the grid of a step that started with
        grid_n bodies (src/nbody.cu:529-539).  Raises when that launch would read outside the body store."""
        img = np.zeros((height, width), dtype=np.uint8)
        rc = gpuref().gpuref_render(width, height, field_w, field_h, grid_n, img.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"gpuref_render failed: {rc}")
        return img

    def read(self):
        n = gpuref().gpuref_n()
        block = np.zeros(6 * max(n, 1), dtype=np.float32)
        if n > 0:
            gpuref().gpuref_read(_fptr(block))
        return block[:6 * max(n, 0)], n

    def close(self):
        gpuref().gpuref_close()


def gpuref_time_kernels(block, n, par: OrcParams, warmup=1, reps=3) -> float:
    ms = C.c_float(0)
    rc = gpuref().gpuref_time_kernels(_fptr(np.ascontiguousarray(block[:6 * n])), n, par.dt, par.growth,
                                      par.field_w, par.field_h, warmup, reps, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"gpuref_time_kernels failed: {rc}")
    return float(ms.value) / reps
