"""The REAL reference (numba kernels + its own classes), timed on a bounded sample of the bright-burst workload.

TEST / BENCH INFRASTRUCTURE ONLY: used by ``bench.py --impl reference`` and by the ``cpu_baseline`` leg of the GPU arm,
never by the product package.  The reference's files are read from ``/root/reference`` (authoring container) or
from the staged copy ``oracle/_ref`` (GPU box; ``oracle/stage_ref.py``), behind the import shims of
``oracle/ref_shims.py``.  One task = the reference's public per-detector call, ``GBMLightCurve.process()``
(lightcurve/lightcurve.py:167-200: sample_times -> sample_photons -> digitize -> background -> combine -> dead time)
on a window of the burst, built from the reference's own ``Source`` / ``CPLSourceFunction`` / ``Background`` /
``Response`` classes; the window is thinned with the FULL burst's fmax (the reference computes fmax over the Source's
own interval, source.py:88-110, so the attribute is set after construction), which makes the cost per event that of
the whole burst.  The numba kernels are compiled once in the parent; the workers are forked afterwards and inherit
the compiled code, so every host core runs the reference's single-threaded loop on its own windows."""
import multiprocessing as mp
import os
import time

import numpy as np

_STATE = {}


def available():
    """(root, reason): the reference tree to use, or (None, why not)"""
    try:
        import numba  # noqa: F401
    except Exception as e:          # noqa: BLE001
        return None, f"numba not importable: {e}"
    from oracle import stage_ref

    if os.path.isdir(os.path.join(os.environ.get("COSMOGRB_REFERENCE", "/root/reference"), "cosmogrb")):
        return os.environ.get("COSMOGRB_REFERENCE", "/root/reference"), ""
    root = stage_ref.staged_root()
    if root:
        return root, ""
    return None, "no reference tree (/root/reference absent and oracle/_ref not staged)"


def setup(params, det="n0", interp_extrap="linear"):
    """import the reference, build the response / template objects, compile every kernel (one tiny window)"""
    root, why = available()
    if root is None:
        raise RuntimeError(why)
    os.environ["COSMOGRB_REFERENCE"] = root
    from oracle import ref_shims

    ref_shims.REFERENCE_ROOT = root
    ref_shims.install(interp_extrap)
    ns = ref_shims.load()
    from cosmogrb_b200 import workloads as h
    from cosmogrb_b200.instruments.gbm.synthetic_response import GBM_DETECTORS, gbm_tables

    rsp = h.response(det)

    class Rsp(ns.response.Response):
        separation_angle = 30.0
        trigger_time = 0.0
        T0 = 0.0

    rrsp = Rsp(matrix=rsp.matrix, geometric_area=rsp.geometric_area, energy_edges=rsp.energy_edges,
               channel_edges=rsp.channel_edges)
    tmpl = ns.background.BackgroundSpectrumTemplate(gbm_tables()["bkg_counts"][GBM_DETECTORS.index(det)])
    p = dict(params)
    sf = ns.cpl_source.CPLSourceFunction(response=rrsp, peak_flux=p["peak_flux"], trise=p["trise"], tdecay=p["tdecay"],
                                         ep_tau=p["ep_tau"], alpha=p["alpha"], ep_start=p["ep_start"])
    full = ns.source.Source(0.0, p["duration"], sf, z=p["z"])
    _STATE.update(ns=ns, rrsp=rrsp, tmpl=tmpl, p=p, fmax=float(full._fmax), root=root)
    window((0, 0.0, 1e-3, -100.0, -99.9, 500.0))         # compiles sample_events / sample_energy / _digitize / ...
    return _STATE["fmax"]


def window(task):
    """(k, a, b, ba, bb, rate) -> (events, source events): GBMLightCurve.process() on source window [a, b] and
    background window [ba, bb]"""
    k, a, b, ba, bb, rate = task
    ns, p = _STATE["ns"], _STATE["p"]
    sf = ns.cpl_source.CPLSourceFunction(response=_STATE["rrsp"], peak_flux=p["peak_flux"], trise=p["trise"],
                                         tdecay=p["tdecay"], ep_tau=p["ep_tau"], alpha=p["alpha"], ep_start=p["ep_start"])
    src = ns.source.Source(a, b, sf, z=p["z"])
    src._fmax = _STATE["fmax"]                            # thin with the whole burst's fmax
    bkg = ns.background.Background(ba, bb, average_rate=rate, background_spectrum_template=_STATE["tmpl"])
    lc = ns.gbm_lightcurve.GBMLightCurve(src, bkg, _STATE["rrsp"], name="n0", grb_name="bench", tstart=ba, tstop=bb)
    st = lc.process()
    return len(st.times_source) + len(st.times_background), len(st.times_source)


def run_sample(n_windows, window_s, procs, duration, bkg=(-100.0, 300.0, 500.0)):
    """`n_windows` windows of `window_s` seconds spread evenly over the burst (background window scaled to the
    background interval), one process per host core.  Returns (events, source events, seconds)."""
    span = bkg[1] - bkg[0]
    tasks = []
    for k in range(n_windows):
        a = (k + 0.5) * duration / n_windows - 0.5 * window_s
        ba = bkg[0] + a * span / duration
        tasks.append((k, a, a + window_s, ba, ba + window_s * span / duration, bkg[2]))
    ctx = mp.get_context("fork")            # the children inherit the compiled kernels
    with ctx.Pool(procs) as pool:
        pool.map(_noop, range(procs))       # workers up before the clock starts
        t0 = time.perf_counter()
        res = pool.map(window, tasks, chunksize=1)
        dt = time.perf_counter() - t0
    return sum(r[0] for r in res), sum(r[1] for r in res), dt


def _noop(x):
    return x
