"""ctypes binding of libcosmogrb_b200.so (the C ABI declared in include/cosmogrb_b200.h).

The product path has no CPU fallback: if the shared library is missing, importing the
symbols raises ``EngineUnavailable`` with the build command, and every sampler call fails
loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# COSMOGRB_B200_LIB selects another build of the same library (tuning runs: variants compiled with -D flags)
LIB_PATH = os.environ.get("COSMOGRB_B200_LIB") or os.path.join(_HERE, "_lib", "libcosmogrb_b200.so")

ABI_VERSION = 7
NBINS = 140
NCHAN = 128
NE_LAMBDA = 75
NE_ENVELOPE = 500
NT_FMAX = 50
SEG = 8192
MERGE_TILE = 2048
DT_TILE = 2048
MD_TILE = 2048
BKG_TILE = 2048
TRIG_TILE = 4096
ENV_MAXW = 256

# detector table block layout (doubles) -- must match include/cosmogrb_b200.h
DT_EDGES = 0
DT_EA_X = 144
DT_EA_Y = 288
DT_BKG_CDF = 432
DT_G75_E = 560
DT_G75_A = 640
DT_G500_E = 720
DT_G500_A = 1220
DT_CUM = 1720
DT_GUIDE = 1720 + NBINS * NCHAN           # NBINS*NCHAN bytes of uint8 search guides
BKG_GUIDE_N = 1024
DT_BKG_GUIDE = DT_GUIDE + NBINS * NCHAN // 8    # BKG_GUIDE_N bytes: guide of the background channel draw
DT_SIZE = DT_BKG_GUIDE + BKG_GUIDE_N // 8

ST_SRC_CAPACITY = 1
ST_BKG_CAPACITY = 2
ST_UNIFORMS_EXHAUSTED = 4
ST_TRIAL_CAP = 8
ST_OUT_CAPACITY = 16
ST_ENVELOPE = 32
ST_DEADTIME_CHAIN = 64
ST_MAJORANT = 128
ST_SQUEEZE = 256
ST_INTERNAL = 512

UNIT_NO_SOURCE = 1
UNIT_NO_BACKGROUND = 2


class EngineUnavailable(RuntimeError):
    pass


# numpy mirrors of the C structs (used to build / read device arrays in bulk)
UNIT_DTYPE = np.dtype(
    [
        ("kind", "<i4"), ("det", "<i4"), ("stream_id", "<u4"), ("flags", "<u4"),
        ("peak_flux", "<f8"), ("ep_start", "<f8"), ("ep_tau", "<f8"), ("alpha", "<f8"),
        ("trise", "<f8"), ("tdecay", "<f8"), ("z", "<f8"),
        ("src_tstart", "<f8"), ("src_tstop", "<f8"),
        ("bkg_tstart", "<f8"), ("bkg_tstop", "<f8"), ("bkg_rate", "<f8"),
        ("fmax", "<f8"), ("gamma_a", "<f8"), ("lgamma_a", "<f8"), ("lgamma_1pa", "<f8"),
        ("src_cap", "<i8"), ("bkg_cap", "<i8"), ("src_off", "<i8"), ("bkg_off", "<i8"),
        ("tot_off", "<i8"), ("tab_off", "<i8"), ("tab_n", "<i8"), ("ea_scale", "<f8"),
        ("etab_off", "<i8"), ("etab_nw", "<i8"),
    ],
    align=False,
)
assert UNIT_DTYPE.itemsize == 224

COUNTS_DTYPE = np.dtype(
    [
        ("n_candidates", "<i8"), ("n_src", "<i8"), ("n_bkg", "<i8"), ("n_total", "<i8"),
        ("n_kept", "<i8"), ("n_trials", "<i8"), ("n_exact", "<i8"), ("status", "<u4"), ("pad", "<u4"),
    ]
)
assert COUNTS_DTYPE.itemsize == 64

TRIGGER_DTYPE = np.dtype([("detected", "<i4"), ("combo", "<i4"), ("n_bins", "<i4"), ("razor", "<i4"),
                          ("time", "<f8"), ("time_scale", "<f8")])
assert TRIGGER_DTYPE.itemsize == 32

WORK_DTYPE = np.dtype([("unit", "<i4"), ("seg", "<i4"), ("start", "<i8")])
assert WORK_DTYPE.itemsize == 16


class TriggerPar(C.Structure):
    _fields_ = [("threshold", C.c_double), ("base_timescale", C.c_double), ("background_duration", C.c_double),
                ("pre_window", C.c_double)]


DEADTIME_REFERENCE = 0
DEADTIME_PHYSICAL = 1


class DeadtimeModel(C.Structure):
    """cgrb_deadtime_model_t"""
    _fields_ = [("mode", C.c_int32), ("pad", C.c_int32), ("tau", C.c_double), ("tau_overflow", C.c_double)]


class RngDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("pad", C.c_int32),
        ("seed", C.c_uint64),
        ("d_u", C.c_void_p),
        ("d_off", C.c_void_p),
        ("d_len", C.c_void_p),
    ]


EXPORTS = (
    "cgrb_abi_version",
    "cgrb_last_error",
    "cgrb_device_info",
    "cgrb_profile_enable",
    "cgrb_profile_read",
    "cgrb_launch_count",
    "cgrb_build_work",
    "cgrb_counts_reset",
    "cgrb_energy_integrated_evolution",
    "cgrb_evolution",
    "cgrb_fmax",
    "cgrb_thin_cum_bytes",
    "cgrb_thin_tables",
    "cgrb_sample_events",
    "cgrb_sample_energy",
    "cgrb_energy_workspace_bytes",
    "cgrb_sample_energy_envelope",
    "cgrb_energy_envelope_tables",
    "cgrb_sample_energy_envelope_prepared",
    "cgrb_digitize",
    "cgrb_sample_background_workspace_bytes",
    "cgrb_sample_background",
    "cgrb_background_channels",
    "cgrb_combine",
    "cgrb_gbm_dead_time",
    "cgrb_merge_dead_time_workspace_bytes",
    "cgrb_merge_dead_time",
    "cgrb_gbm_trigger_workspace_bytes",
    "cgrb_gbm_trigger",
)

_lib = None


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.cgrb_abi_version.restype = i32
    lib.cgrb_last_error.restype = C.c_char_p
    lib.cgrb_device_info.argtypes = [i32, C.POINTER(C.c_int)]
    lib.cgrb_profile_enable.argtypes = [i32]
    lib.cgrb_profile_read.argtypes = [vp, vp, i32]
    lib.cgrb_launch_count.restype = C.c_longlong
    lib.cgrb_build_work.restype = i64
    lib.cgrb_build_work.argtypes = [vp, i32, i32, i64, vp, i64]
    lib.cgrb_counts_reset.argtypes = [vp, i32, vp]
    lib.cgrb_energy_integrated_evolution.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp]
    lib.cgrb_evolution.argtypes = [vp, i32, vp, i64, vp, i64, vp, vp]
    lib.cgrb_fmax.argtypes = [vp, vp, i32, i32, vp]
    rp = C.POINTER(RngDesc)
    lib.cgrb_thin_cum_bytes.argtypes = [i64]
    lib.cgrb_thin_cum_bytes.restype = i64
    lib.cgrb_thin_tables.argtypes = [vp, vp, i32, vp, i64, vp, vp, i64, vp]
    lib.cgrb_sample_events.argtypes = [vp, vp, i32, vp, i64, vp, i32, rp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64,
                                       vp, vp]
    lib.cgrb_sample_energy.argtypes = [vp, vp, i32, vp, i64, i32, rp, i64, vp, vp, vp, vp, vp]
    lib.cgrb_energy_workspace_bytes.argtypes = [i32, i64]
    lib.cgrb_energy_workspace_bytes.restype = i64
    lib.cgrb_sample_energy_envelope.argtypes = [vp, vp, i32, vp, i64, i32, rp, i64, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.cgrb_energy_envelope_tables.argtypes = [vp, vp, i32, i32, vp, i64, vp]
    lib.cgrb_sample_energy_envelope_prepared.argtypes = [vp, vp, i32, vp, i64, i32, rp, i64, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.cgrb_digitize.argtypes = [vp, vp, i32, vp, i64, rp, vp, vp, vp, vp, vp]
    lib.cgrb_sample_background_workspace_bytes.argtypes = [i32, i64]
    lib.cgrb_sample_background_workspace_bytes.restype = i64
    lib.cgrb_sample_background.argtypes = [vp, vp, i32, i64, rp, rp, vp, vp, vp, vp, vp]
    lib.cgrb_background_channels.argtypes = [vp, i32, C.c_uint32, rp, i64, vp, vp]
    lib.cgrb_combine.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    dp = C.POINTER(DeadtimeModel)
    lib.cgrb_gbm_dead_time.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, dp, vp]
    lib.cgrb_merge_dead_time_workspace_bytes.argtypes = [i64]
    lib.cgrb_merge_dead_time_workspace_bytes.restype = i64
    lib.cgrb_merge_dead_time.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, dp, vp]
    lib.cgrb_gbm_trigger_workspace_bytes.argtypes = [i32, i64]
    lib.cgrb_gbm_trigger_workspace_bytes.restype = i64
    lib.cgrb_gbm_trigger.argtypes = [vp, i32, vp, i64, vp, C.POINTER(TriggerPar), vp, vp, vp, i64, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("cgrb_last_error", "cgrb_build_work", "cgrb_launch_count", "cgrb_energy_workspace_bytes",
                        "cgrb_merge_dead_time_workspace_bytes", "cgrb_gbm_trigger_workspace_bytes",
                        "cgrb_sample_background_workspace_bytes",
                        "cgrb_thin_cum_bytes"):
            fn.restype = i32
    if lib.cgrb_abi_version() != ABI_VERSION:
        raise EngineUnavailable(
            f"ABI mismatch: library {lib.cgrb_abi_version()} vs binding {ABI_VERSION}; rebuild the library"
        )
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().cgrb_last_error().decode("utf-8", "replace")
        raise EngineError(f"{what} failed with code {rc}: {msg}")


def build_work(units: np.ndarray, which: int, chunk: int) -> np.ndarray:
    """Host helper: chunk work list (cgrb_build_work)."""
    lib = load()
    units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
    n = lib.cgrb_build_work(units.ctypes.data, len(units), which, chunk, None, 0)
    if n < 0:
        check(int(n), "cgrb_build_work")
    work = np.zeros(int(n), dtype=WORK_DTYPE)
    n2 = lib.cgrb_build_work(units.ctypes.data, len(units), which, chunk, work.ctypes.data, int(n))
    if n2 != n:
        check(int(n2) if n2 < 0 else -1, "cgrb_build_work")
    return work


def profile_enable(on=True):
    check(load().cgrb_profile_enable(1 if on else 0), "cgrb_profile_enable")


def profile_read(cap=65536):
    """-> list of (kernel name, milliseconds) for every launch since profile_enable()."""
    names = C.create_string_buffer(cap * 32)
    ms = (C.c_float * cap)()
    n = load().cgrb_profile_read(names, ms, cap)
    if n < 0:
        check(n, "cgrb_profile_read")
    raw = names.raw
    return [(raw[i * 32:(i + 1) * 32].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n)]


def launch_count():
    return int(load().cgrb_launch_count())
