"""ppa-nbody-collisions_b200 -- host-side mirror of the C ABI in include/nbody_b200.h.

The product is `lib/libnbody_b200.so` (hand-written sm_100a kernels behind a C ABI,
built in tree by `csrc/Makefile`).  This module is only the ctypes binding a Python
host uses to reach it; it holds no compute and has NO fallback: if the library is
missing, or no sm_100 GPU is present, it raises.

The directory name has a hyphen (it is the reference's name), so import it with
`importlib.import_module("ppa-nbody-collisions_b200")` or `__graft_entry__.load_package()`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("NBODY_B200_LIB", HERE / "lib" / "libnbody_b200.so"))   # override: kernel experiments only
DRIVER_PATH = HERE / "bin" / "nbody"

OK = 0
ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_CANDIDATE_OVERFLOW, ERR_COMM, ERR_IO, ERR_EVENT_OVERFLOW = -1, -2, -3, -4, -5, -6, -7
COVERAGE_REFERENCE, COVERAGE_FULL = 0, 1
EV_ABSORB, EV_KILLED = 0, 1
FLAG_NO_GRAPH, FLAG_SCALAR_FORCE, FLAG_NO_SORT, FLAG_MERGE_CONSERVING = 1, 2, 4, 16
FLAG_ONE_SIDED, FLAG_PAIR_HALVING, FLAG_SYM_ROWS8 = 32, 64, 128
SCENARIO_SQUARE, SCENARIO_DISC, SCENARIO_TWO_GALAXY = 0, 1, 2
UNIQUE_ID_BYTES = 128

# every symbol include/nbody_b200.h declares
SYMBOLS = [
    "nb_create", "nb_destroy", "nb_last_error", "nb_version", "nb_upload", "nb_download", "nb_num_bodies",
    "nb_step", "nb_step_timed", "nb_step_profile", "nb_sync", "nb_get_stats", "nb_events", "nb_comm_unique_id", "nb_comm_init",
    "nb_plan_host", "nb_plan_block", "nb_plan_block_index", "nb_plan_warp_items", "nb_plan_warp_item", "nb_plan_force_scale", "nb_render", "nb_render_grid", "nb_write_pgm", "nb_config_parse", "nb_rng_seed", "nb_rng_ival64", "nb_rng_fval",
    "nb_rng_fval_range", "nb_generate", "nb_probe_fp32",
]


def flag_variant(v: int) -> int:
    """nb_params.flags bits selecting force-kernel variant v (NB_FLAG_VARIANT)."""
    return v << 8


class Params(C.Structure):
    _fields_ = [("n_max", C.c_int), ("dt", C.c_float), ("growth", C.c_float), ("field_w", C.c_int),
                ("field_h", C.c_int), ("grav", C.c_float), ("coverage", C.c_int), ("device", C.c_int),
                ("candidate_capacity", C.c_int), ("event_capacity", C.c_int), ("rank", C.c_int),
                ("world", C.c_int), ("flags", C.c_int), ("sort_min_n", C.c_int), ("softening", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("pairs", C.c_int64), ("candidates", C.c_int64), ("exact_chunks", C.c_int64),
                ("fast_chunks", C.c_int64), ("n", C.c_int32), ("overflow", C.c_int32), ("events_dropped", C.c_int32),
                ("sm_count", C.c_int32), ("force_grid", C.c_int32), ("force_regs", C.c_int32),
                ("row_lo", C.c_int32), ("row_hi", C.c_int32), ("force_threads", C.c_int32),
                ("force_variant", C.c_int32), ("culled_parts", C.c_int64),
                ("pair_halving", C.c_int32), ("sym_regs", C.c_int32), ("kernel_launches", C.c_int64),
                ("force_partials", C.c_int32), ("reserved", C.c_int32)]


class Plan(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "blocks", "limit_last", "limit_first", "n_active", "window_len",
                                          "row_lo", "row_hi", "row_act_hi", "rows_per_rank", "n_iblocks", "n_jtiles")]
    _fields_ = _fields_ + [("units", C.c_int64)] + [(k, C.c_int32) for k in ("sorted", "two_sided", "sym_S", "sym_Q",
                                                                              "sym_blocks", "sym_lgu")]


class Config(C.Structure):
    _fields_ = [("particleCount", C.c_int), ("totalIterations", C.c_int), ("save_Image_Every_Xth_Iteration", C.c_int),
                ("timestep", C.c_float), ("minRandBodyMass", C.c_float), ("maxRandBodyMass", C.c_float),
                ("minRadius", C.c_float), ("maxRadius", C.c_float), ("growthRate", C.c_float),
                ("imgWidth", C.c_int), ("imgHeight", C.c_int), ("fieldWidth", C.c_int), ("fieldHeight", C.c_int),
                ("imagePath", C.c_char * 1024)]


class Rng(C.Structure):
    _fields_ = [("u", C.c_uint64), ("v", C.c_uint64), ("w", C.c_uint64)]


class Scenario(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("seed", C.c_uint64), ("field_w", C.c_int), ("field_h", C.c_int),
                ("min_mass", C.c_float), ("max_mass", C.c_float), ("min_radius", C.c_float), ("max_radius", C.c_float),
                ("extent", C.c_double)]


EVENT_DTYPE = np.dtype([("step", np.int32), ("i", np.int32), ("j", np.int32), ("kind", np.int32)])


class NbodyError(RuntimeError):
    def __init__(self, what: str, code: int, msg: str):
        super().__init__(f"{what} failed ({code}): {msg}")
        self.code = code


def build(force: bool = False) -> None:
    """Compile lib/libnbody_b200.so and bin/nbody in tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", str(HERE / "csrc"), "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", str(HERE / "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libnbody_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])


_lib = None


def lib() -> C.CDLL:
    """The C-ABI library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
    L = C.CDLL(str(LIB_PATH))
    vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.nb_create.argtypes = [C.POINTER(vp), C.POINTER(Params)]
    L.nb_destroy.argtypes = [vp]
    L.nb_destroy.restype = None
    L.nb_last_error.argtypes = [vp]
    L.nb_last_error.restype = C.c_char_p
    L.nb_upload.argtypes = [vp, vp, C.c_int]
    L.nb_download.argtypes = [vp, vp, C.c_int, ip]
    L.nb_num_bodies.argtypes = [vp, ip]
    L.nb_step.argtypes = [vp, C.c_int]
    L.nb_step_timed.argtypes = [vp, C.c_int, fp, fp]
    L.nb_step_profile.argtypes = [vp, C.c_int, fp]
    L.nb_sync.argtypes = [vp]
    L.nb_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.nb_events.argtypes = [vp, vp, C.c_int, ip]
    L.nb_comm_unique_id.argtypes = [vp]
    L.nb_comm_init.argtypes = [vp, vp]
    L.nb_plan_host.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.POINTER(Plan)]
    L.nb_plan_block.argtypes = [C.c_int, C.c_int, ip, ip]
    L.nb_plan_block_index.argtypes = [C.c_int, C.c_int, C.c_int]
    L.nb_plan_warp_items.argtypes = [C.c_int, C.c_int, ip]
    L.nb_plan_warp_item.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip]
    L.nb_plan_force_scale.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, ip]
    L.nb_render.argtypes = [vp, vp, C.c_int, C.c_int]
    L.nb_render_grid.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.nb_write_pgm.argtypes = [C.c_char_p, vp, C.c_int, C.c_int]
    L.nb_config_parse.argtypes = [C.c_char_p, C.POINTER(Config), C.c_int]
    L.nb_rng_seed.argtypes = [C.POINTER(Rng), C.c_uint64]
    L.nb_rng_seed.restype = None
    L.nb_rng_ival64.argtypes = [C.POINTER(Rng)]
    L.nb_rng_ival64.restype = C.c_uint64
    L.nb_rng_fval.argtypes = [C.POINTER(Rng)]
    L.nb_rng_fval.restype = C.c_double
    L.nb_rng_fval_range.argtypes = [C.POINTER(Rng), C.c_double, C.c_double]
    L.nb_rng_fval_range.restype = C.c_double
    L.nb_generate.argtypes = [C.POINTER(Scenario), vp]
    L.nb_probe_fp32.argtypes = [C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


# ---------------------------------------------------------------------------------------------
# host-only helpers (no GPU)
# ---------------------------------------------------------------------------------------------
def generate(kind: int, n: int, seed: int = 1024, field_w: int = 100000, field_h: int = 100000,
             min_mass: float = 1e4, max_mass: float = 1e17, min_radius: float = 50.0, max_radius: float = 200.0,
             extent: float = 0.0) -> np.ndarray:
    """A BodiesData block (6 n float32: pos[n][2], vel[n][2], mass[n], radius[n])."""
    sc = Scenario(kind, n, seed, field_w, field_h, np.float32(min_mass), np.float32(max_mass),
                  np.float32(min_radius), np.float32(max_radius), float(extent))
    block = np.zeros(6 * n, dtype=np.float32)
    rc = lib().nb_generate(C.byref(sc), block.ctypes.data)
    if rc != OK:
        raise NbodyError("nb_generate", rc, "invalid scenario")
    return block


def parse_config(path: str, echo_fd: int = -1) -> tuple[int, Config]:
    cfg = Config()
    rc = lib().nb_config_parse(str(path).encode(), C.byref(cfg), echo_fd)
    return rc, cfg


def plan(n: int, coverage: int = COVERAGE_FULL, rank: int = 0, world: int = 1, force_grid: int = 592, flags: int = 0,
         sort_min_n: int = 0, n_max: int = 0) -> dict:
    p = Params(n_max=max(n, n_max, 1), coverage=coverage, rank=rank, world=world, field_w=1, field_h=1, flags=flags,
               sort_min_n=sort_min_n)
    out = Plan()
    rc = lib().nb_plan_host(C.byref(p), n, force_grid, C.byref(out))
    if rc != OK:
        raise NbodyError("nb_plan_host", rc, "invalid arguments")
    return {k: getattr(out, k) for k, _ in Plan._fields_}


def plan_block(Q: int, b: int):
    """(R, C), R <= C: super-tiles of block b in the two-sided kernel's queue order."""
    R, Cc = C.c_int(), C.c_int()
    rc = lib().nb_plan_block(Q, b, C.byref(R), C.byref(Cc))
    if rc != OK:
        raise NbodyError("nb_plan_block", rc, "invalid arguments")
    return R.value, Cc.value


def plan_block_index(Q: int, X: int, Y: int) -> int:
    return lib().nb_plan_block_index(Q, X, Y)


def plan_warp_items(n: int, run: int) -> int:
    """Work-item ids of the warp-level two-sided kernel for n bodies, `run` chunks per item."""
    ids = C.c_int()
    rc = lib().nb_plan_warp_items(n, run, C.byref(ids))
    if rc != OK:
        raise NbodyError("nb_plan_warp_items", rc, "invalid arguments")
    return ids.value


def plan_warp_item(n: int, run: int, item: int):
    """(group, chunk_lo, chunk_hi) of a work item, or None for a void id."""
    g, lo, hi = C.c_int(), C.c_int(), C.c_int()
    rc = lib().nb_plan_warp_item(n, run, item, C.byref(g), C.byref(lo), C.byref(hi))
    if rc != OK:
        raise NbodyError("nb_plan_warp_item", rc, "invalid arguments")
    return None if g.value < 0 else (g.value, lo.value, hi.value)


def plan_force_scale(n: int, m_max: float, r_min: float, field: int):
    """log2 of the fixed-point scale of the two-sided kernels' force sums, or None when the plan would fall back."""
    k = C.c_int()
    rc = lib().nb_plan_force_scale(n, np.float32(m_max), np.float32(r_min), field, C.byref(k))
    return k.value if rc == OK else None


def split(block: np.ndarray, n: int):
    """Views (pos[n,2], vel[n,2], mass[n], radius[n]) of a BodiesData block."""
    return (block[:2 * n].reshape(n, 2), block[2 * n:4 * n].reshape(n, 2), block[4 * n:5 * n], block[5 * n:6 * n])


# ---------------------------------------------------------------------------------------------
# the simulation context (needs a B200)
# ---------------------------------------------------------------------------------------------
class Simulation:
    """One nb_ctx: one GPU's share of the bodies.  Mirrors the C ABI call for call."""

    def __init__(self, n_max: int, dt: float = 0.2, growth: float = 0.1, field_w: int = 100000, field_h: int = 100000,
                 coverage: int = COVERAGE_REFERENCE, device: int = 0, event_capacity: int = 0,
                 candidate_capacity: int = 0, rank: int = 0, world: int = 1, flags: int = 0, grav: float = 0.0,
                 sort_min_n: int = 0, softening: float = 0.0):
        self._h = C.c_void_p()
        self.params = Params(n_max, np.float32(dt), np.float32(growth), field_w, field_h, np.float32(grav), coverage,
                             device, candidate_capacity, event_capacity, rank, world, flags, sort_min_n, np.float32(softening))
        rc = lib().nb_create(C.byref(self._h), C.byref(self.params))
        if rc != OK:
            self._h = C.c_void_p()
            raise NbodyError("nb_create", rc, (lib().nb_last_error(None) or b"").decode())
        self.n_max = n_max

    def _check(self, what: str, rc: int):
        if rc != OK:
            raise NbodyError(what, rc, (lib().nb_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            lib().nb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, block: np.ndarray, n: int):
        block = np.ascontiguousarray(block[:6 * n], dtype=np.float32)
        self._check("nb_upload", lib().nb_upload(self._h, block.ctypes.data, n))

    def upload_ptr(self, ptr: int, n: int):
        """Upload from a raw host pointer (e.g. pinned torch memory)."""
        self._check("nb_upload", lib().nb_upload(self._h, C.c_void_p(ptr), n))

    def download(self, out: np.ndarray | None = None) -> tuple[np.ndarray, int]:
        if out is None:
            out = np.zeros(6 * self.n_max, dtype=np.float32)
        n = C.c_int(0)
        self._check("nb_download", lib().nb_download(self._h, out.ctypes.data, out.size // 6, C.byref(n)))
        return out[:6 * n.value], n.value

    def download_ptr(self, ptr: int, capacity_n: int) -> int:
        n = C.c_int(0)
        self._check("nb_download", lib().nb_download(self._h, C.c_void_p(ptr), capacity_n, C.byref(n)))
        return n.value

    def num_bodies(self) -> int:
        n = C.c_int(0)
        self._check("nb_num_bodies", lib().nb_num_bodies(self._h, C.byref(n)))
        return n.value

    def step(self, n_steps: int = 1):
        self._check("nb_step", lib().nb_step(self._h, n_steps))

    def step_timed(self, n_steps: int = 1, force: bool = True) -> tuple[float, float]:
        """(ms_total, ms_force) measured with CUDA events on the context's stream."""
        tot, frc = C.c_float(0), C.c_float(0)
        self._check("nb_step_timed", lib().nb_step_timed(self._h, n_steps, C.byref(tot), C.byref(frc) if force else None))
        return tot.value, frc.value

    def step_profile(self, n_steps: int = 1) -> dict:
        """Per-kernel device times in ms summed over n_steps (no graph): force, finish, allgather, compact."""
        ms = (C.c_float * 5)()
        self._check("nb_step_profile", lib().nb_step_profile(self._h, n_steps, ms))
        return {"force": ms[0], "finish": ms[1], "allgather": ms[2], "compact": ms[3], "sort": ms[4]}

    def sync(self):
        self._check("nb_sync", lib().nb_sync(self._h))

    def stats(self) -> dict:
        s = Stats()
        self._check("nb_get_stats", lib().nb_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def events(self, capacity: int | None = None) -> np.ndarray:
        """All records logged since the last call (nb_events).  The default buffer holds the whole log (event_capacity)."""
        if capacity is None:
            capacity = max(int(self.params.event_capacity), 1)
        buf = np.zeros(capacity, dtype=EVENT_DTYPE)
        cnt = C.c_int(0)
        self._check("nb_events", lib().nb_events(self._h, buf.ctypes.data, capacity, C.byref(cnt)))
        return buf[:cnt.value]

    def render(self, w: int, h: int, grid_n: int | None = None) -> np.ndarray:
        """All live bodies, or -- with grid_n, the body count before the last step -- what the reference's stale launch
        grid draws (nb_render_grid)."""
        img = np.zeros((h, w), dtype=np.uint8)
        if grid_n is None:
            self._check("nb_render", lib().nb_render(self._h, img.ctypes.data, w, h))
        else:
            self._check("nb_render_grid", lib().nb_render_grid(self._h, img.ctypes.data, w, h, 128 * max(1, grid_n // 128)))
        return img

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)
        self._check("nb_comm_init", lib().nb_comm_init(self._h, buf))


def probe_fp32(device: int = 0) -> float:
    """Measured FP32 peak of the device in TFLOP/s (packed FFMA2 stream), see nb_probe_fp32."""
    t = C.c_double(0)
    rc = lib().nb_probe_fp32(device, C.byref(t))
    if rc != OK:
        raise NbodyError("nb_probe_fp32", rc, "no usable CUDA device")
    return t.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    rc = lib().nb_comm_unique_id(buf)
    if rc != OK:
        raise NbodyError("nb_comm_unique_id", rc, (lib().nb_last_error(None) or b"").decode())
    return buf.raw
