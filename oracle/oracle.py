"""ctypes wrapper of the C oracle (oracle/cgrb_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package never does.

Parity pinning: the C restatement is checked against the real numba kernels of the reference
(imported through oracle/ref_shims.py in the authoring container) by oracle/make_golden.py;
the vectors it mints live under tests/golden/ and are re-checked by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libcgrb_oracle.so")


class CoRng(C.Structure):
    _fields_ = [
        ("mode", C.c_int),
        ("u", C.c_void_p),
        ("n_u", C.c_int64),
        ("pos", C.c_int64),
        ("exhausted", C.c_int),
        ("mt", C.c_uint32 * 624),
        ("mti", C.c_int),
    ]


class CoSource(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("interp_extrap", C.c_int32),
        ("peak_flux", C.c_double),
        ("ep_start", C.c_double),
        ("ep_tau", C.c_double),
        ("alpha", C.c_double),
        ("trise", C.c_double),
        ("tdecay", C.c_double),
        ("z", C.c_double),
        ("emin", C.c_double),
        ("emax", C.c_double),
    ]


def build(force=False):
    """Compile the oracle with gcc (seconds)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("cgrb_oracle.c", "cgrb_oracle.h")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    d, vp, i64, i32 = C.c_double, C.c_void_p, C.c_int64, C.c_int
    rp, sp = C.POINTER(CoRng), C.POINTER(CoSource)
    L.co_rng_init_array.argtypes = [rp, vp, i64]
    L.co_rng_init_mt.argtypes = [rp, C.c_uint32]
    L.co_rng_next.argtypes = [rp]
    L.co_rng_next.restype = d
    L.co_gammaincc.argtypes = [d, d]
    L.co_gammaincc.restype = d
    L.co_norris.argtypes = [d] * 5
    L.co_norris.restype = d
    L.co_cpl.argtypes = [d] * 6
    L.co_cpl.restype = d
    L.co_interp.argtypes = [vp, vp, i32, d, i32]
    L.co_interp.restype = d
    L.co_folded_cpl.argtypes = [sp, vp, vp, i32, d, d]
    L.co_folded_cpl.restype = d
    L.co_cpl_evolution.argtypes = [sp, d, d]
    L.co_cpl_evolution.restype = d
    L.co_energy_integrated_evolution.argtypes = [sp, vp, vp, i32, d]
    L.co_energy_integrated_evolution.restype = d
    L.co_fmax.argtypes = [sp, vp, vp, i32, d, d]
    L.co_fmax.restype = d
    L.co_sample_events.argtypes = [sp, vp, vp, i32, d, d, d, rp, vp, i64, vp, vp, i64, vp]
    L.co_sample_events.restype = i64
    L.co_sample_energy.argtypes = [sp, vp, vp, i32, vp, i64, rp, vp, i64, vp, vp]
    L.co_sample_energy.restype = i64
    L.co_digitize.argtypes = [vp, i64, vp, i32, vp, i32, rp, vp]
    L.co_digitize.restype = None
    L.co_background_times.argtypes = [d, d, d, rp, vp, i64]
    L.co_background_times.restype = i64
    L.co_background_channels.argtypes = [vp, i32, i64, rp, vp]
    L.co_background_channels.restype = None
    L.co_combine.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp]
    L.co_combine.restype = None
    L.co_gbm_dead_time.argtypes = [vp, vp, i64, vp]
    L.co_gbm_dead_time.restype = i64
    L.co_gbm_dead_time_physical.argtypes = [vp, vp, i64, C.c_double, C.c_double, vp]
    L.co_gbm_dead_time_physical.restype = i64
    L.co_process_detector.argtypes = [sp, vp, i32, vp, vp, vp, i32, vp, d, d, d, d, d, d,
                                      C.c_uint32, C.c_uint32, i64, vp]
    L.co_process_detector.restype = i64
    _lib = L
    return L


def _f8(a):
    return np.ascontiguousarray(a, dtype="f8")


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Rng(object):
    """Uniform source: explicit array (``Rng(u=...)``) or MT19937 (``Rng(seed=...)``), the
    generator behind numba's np.random and numpy's legacy RandomState."""

    def __init__(self, u=None, seed=None):
        self.c = CoRng()
        if u is not None:
            self._u = _f8(u)
            lib().co_rng_init_array(C.byref(self.c), self._u.ctypes.data, len(self._u))
        else:
            lib().co_rng_init_mt(C.byref(self.c), int(seed) & 0xFFFFFFFF)

    @property
    def pos(self):
        return int(self.c.pos)

    @property
    def exhausted(self):
        return bool(self.c.exhausted)

    def next(self):
        return lib().co_rng_next(C.byref(self.c))

    def random_sample(self, n):
        return np.array([self.next() for _ in range(n)])


def source(kind=0, interp_extrap="linear", peak_flux=1e-6, ep_start=500.0, ep_tau=2.0, alpha=-0.66,
           trise=1.0, tdecay=1.0, z=1.0, emin=5.0, emax=5e4, ep=None):
    s = CoSource()
    s.kind = int(kind)
    s.interp_extrap = {"linear": 0, "clamp": 1}[interp_extrap] if isinstance(interp_extrap, str) else int(interp_extrap)
    s.peak_flux, s.ep_tau, s.alpha = float(peak_flux), float(ep_tau), float(alpha)
    s.ep_start = float(ep if ep is not None else ep_start)
    s.trise, s.tdecay, s.z = float(trise), float(tdecay), float(z)
    s.emin, s.emax = float(emin), float(emax)
    return s


def gammaincc(a, x):
    L = lib()
    a, x = np.broadcast_arrays(_f8(a), _f8(x))
    out = np.empty(a.shape)
    of, af, xf = out.reshape(-1), a.reshape(-1), x.reshape(-1)
    for i in range(af.size):
        of[i] = L.co_gammaincc(af[i], xf[i])
    return out


def cpl_evolution(s, energy, time):
    L = lib()
    energy, time = np.atleast_1d(_f8(energy)), np.atleast_1d(_f8(time))
    out = np.empty((len(time), len(energy)))
    for n, t in enumerate(time):
        for m, e in enumerate(energy):
            out[n, m] = L.co_cpl_evolution(C.byref(s), e, t)
    return out


def folded_cpl(s, ea, energy, time):
    L = lib()
    ea_x, ea_y = _f8(ea[0]), _f8(ea[1])
    energy, time = np.atleast_1d(_f8(energy)), np.atleast_1d(_f8(time))
    out = np.empty((len(time), len(energy)))
    for n, t in enumerate(time):
        for m, e in enumerate(energy):
            out[n, m] = L.co_folded_cpl(C.byref(s), ea_x.ctypes.data, ea_y.ctypes.data, len(ea_x), e, t)
    return out


def energy_integrated_evolution(s, ea, time):
    L = lib()
    ea_x, ea_y = _f8(ea[0]), _f8(ea[1])
    time = np.atleast_1d(_f8(time))
    return np.array([L.co_energy_integrated_evolution(C.byref(s), ea_x.ctypes.data, ea_y.ctypes.data,
                                                      len(ea_x), t) for t in time])


def fmax(s, ea, tstart, tstop):
    ea_x, ea_y = _f8(ea[0]), _f8(ea[1])
    return lib().co_fmax(C.byref(s), ea_x.ctypes.data, ea_y.ctypes.data, len(ea_x), tstart, tstop)


def sample_events(s, ea, tstart, tstop, fmax_, rng, cap=None, want_candidates=False):
    """-> times (incl. the spurious tstart), n_candidates[, cand_times, cand_accept]"""
    ea_x, ea_y = _f8(ea[0]), _f8(ea[1])
    if cap is None:
        m = max(fmax_ * (tstop - tstart), 0.0)
        cap = int(m + 10 * np.sqrt(m + 1) + 1024)
    out = np.empty(cap)
    ct = np.empty(cap) if want_candidates else None
    ca = np.zeros(cap, np.uint8) if want_candidates else None
    nc = C.c_int64(0)
    n = lib().co_sample_events(C.byref(s), ea_x.ctypes.data, ea_y.ctypes.data, len(ea_x), tstart, tstop,
                               fmax_, C.byref(rng.c), out.ctypes.data, cap, _ptr(ct), _ptr(ca), cap,
                               C.addressof(nc))
    if n < 0:
        raise RuntimeError("oracle sample_events: capacity or uniform stream exhausted")
    if want_candidates:
        return out[:n].copy(), nc.value, ct[:nc.value].copy(), ca[:nc.value].astype(bool)
    return out[:n].copy(), nc.value


def sample_energy(s, ea, times, rng, offsets=None, max_trials=0):
    """-> energies, trials"""
    ea_x, ea_y = _f8(ea[0]), _f8(ea[1])
    times = _f8(times)
    n = len(times)
    out = np.empty(n)
    tr = np.zeros(n, np.int32)
    off = np.ascontiguousarray(offsets, dtype=np.int64) if offsets is not None else None
    lib().co_sample_energy(C.byref(s), ea_x.ctypes.data, ea_y.ctypes.data, len(ea_x), times.ctypes.data, n,
                           C.byref(rng.c), _ptr(off), int(max_trials), out.ctypes.data, tr.ctypes.data)
    return out, tr


def digitize(energies, edges, cum, rng):
    energies, edges, cum = _f8(energies), _f8(edges), _f8(cum)
    out = np.empty(len(energies))
    lib().co_digitize(energies.ctypes.data, len(energies), edges.ctypes.data, len(edges), cum.ctypes.data,
                      cum.shape[1], C.byref(rng.c), out.ctypes.data)
    return out


def background_times(tstart, tstop, rate, rng, cap=None):
    if cap is None:
        m = rate * (tstop - tstart)
        cap = int(m + 10 * np.sqrt(m + 1) + 1024)
    out = np.empty(cap)
    n = lib().co_background_times(tstart, tstop, rate, C.byref(rng.c), out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("oracle background_times: capacity or uniform stream exhausted")
    return out[:n].copy()


def background_channels(weights, n, rng):
    w = _f8(weights)
    out = np.empty(n, np.int64)
    lib().co_background_channels(w.ctypes.data, len(w), n, C.byref(rng.c), out.ctypes.data)
    return out


def combine(t_bkg, pha_bkg, t_src, pha_src):
    t_bkg, pha_bkg, t_src, pha_src = _f8(t_bkg), _f8(pha_bkg), _f8(t_src), _f8(pha_src)
    n = len(t_bkg) + len(t_src)
    ot, op = np.empty(n), np.empty(n)
    lib().co_combine(t_bkg.ctypes.data, pha_bkg.ctypes.data, len(t_bkg), t_src.ctypes.data,
                     pha_src.ctypes.data, len(t_src), ot.ctypes.data, op.ctypes.data)
    return ot, op


def gbm_dead_time(times, pha):
    times, pha = _f8(times), _f8(pha)
    sel = np.zeros(len(times), np.uint8)
    lib().co_gbm_dead_time(times.ctypes.data, pha.ctypes.data, len(times), sel.ctypes.data)
    return sel.astype(bool)


def gbm_dead_time_physical(times, pha, tau=2.6e-6, tau_overflow=10e-6):
    """serial definition of the non-paralyzable model (not the reference's rule)"""
    times, pha = _f8(times), _f8(pha)
    sel = np.zeros(len(times), np.uint8)
    lib().co_gbm_dead_time_physical(times.ctypes.data, pha.ctypes.data, len(times), float(tau), float(tau_overflow),
                                    sel.ctypes.data)
    return sel.astype(bool)


def process_detector(s, edges, ea, cum, bkg_weights, src_tstart, src_tstop, bkg_tstart, bkg_tstop, bkg_rate,
                     seed_numba=1, seed_numpy=2, max_src_events=0, fmax=0.0):
    """One LightCurve.process() (MT19937 streams). -> (n_events_pre_deadtime, counters[6])"""
    edges, cum, w = _f8(edges), _f8(cum), _f8(bkg_weights)
    ea_x, ea_y = _f8(ea[0]), _f8(ea[1])
    counters = np.zeros(6, np.int64)
    n = lib().co_process_detector(C.byref(s), edges.ctypes.data, len(edges), ea_x.ctypes.data, ea_y.ctypes.data,
                                  cum.ctypes.data, cum.shape[1], w.ctypes.data, src_tstart, src_tstop,
                                  bkg_tstart, bkg_tstop, bkg_rate, float(fmax), seed_numba, seed_numpy, max_src_events,
                                  counters.ctypes.data)
    return int(n), counters
