"""CPU restatement of cosmogrb's GBM trigger.  TEST INFRASTRUCTURE ONLY: imported by
tests/, never by the product package.

Follows, line by line:
  * ``LightCurveStorage._bin_lightcurve`` / ``binned_counts`` / ``_select_channel`` /
    ``get_idx_over_interval``            cosmogrb/lightcurve/light_curve_storage.py:154-294
  * ``GBMLightCurveAnalyzer``           cosmogrb/instruments/gbm/gbm_lightcurve_analyzer.py:13-271
    (``_compute_detection`` :63-143, ``_run_trigger`` :199-243, ``_calculate_dead_time_per_event``
    :246-262, ``dumb_significance`` :265-271)
  * ``TimeInterval.exposure``           cosmogrb/utils/time_interval.py:31
  * ``GBMTrigger.process``              cosmogrb/instruments/gbm/gbm_trigger.py:119-185

Quirks reproduced on purpose:
  * ``binned_counts`` returns ``((counts / dt) * dt).astype(int)`` (a float round trip; exact for dt = 0.016);
  * ``_run_trigger`` is CALLED with (n_bins_background, n_bins_pre, n_bins_src) but DECLARED as
    (n_bins_background, n_bins_source, n_bins_pre): the "source" window is always the 4 s pre-window
    (250 bins) and the gap before it is the trigger time scale;
  * dead time per event is 2 us / 10 us here (the filter uses 2.6 / 10.6 us);
  * ``dead_time_of_interval`` selects ``tmin <= t <= tmax`` (an event on a bin edge counts twice);
  * ``starts[src_idx] > -1.0`` is required for a detection.
Floating point: the reference sums the exposures with numba ``ndarray.sum()`` under fastmath (association
undefined); window sums here use cumulative sums, so significances agree to ~1e-13 relative and decisions
are identical except within that distance of the threshold (``razor`` in the result).
"""
import numpy as np

BASE_TIMESCALE = 0.016
TRIGGER_ENERGY_RANGES = (("a", 50.0, 300.0), ("b", 25.0, 50.0), ("c", 100.0, None), ("d", 300.0, None))
TRIGGER_TIME_SCALES = {
    "a": [BASE_TIMESCALE * (2 ** i) for i in range(10)][::-1],
    "b": [BASE_TIMESCALE * (2 ** i) for i in range(10)][::-1],
    "c": [BASE_TIMESCALE * (2 ** i) for i in range(9)][::-1],
    "d": [BASE_TIMESCALE * (2 ** i) for i in range(4)][::-1],
}


def select_channel(ebounds, emin, emax, pha):
    """light_curve_storage.py:154-182"""
    idx = np.ones_like(pha, dtype=bool)
    if emin is not None:
        idx &= pha > np.searchsorted(ebounds, emin)
    if emax is not None:
        idx &= pha < np.searchsorted(ebounds, emax)
    return idx


def channel_bounds(ebounds):
    """(lo, hi) per trigger energy range: a channel c is selected iff lo < c < hi (hi = 1 << 30: no bound)."""
    out = []
    for _, emin, emax in TRIGGER_ENERGY_RANGES:
        lo = int(np.searchsorted(ebounds, emin)) if emin is not None else -1
        hi = int(np.searchsorted(ebounds, emax)) if emax is not None else (1 << 30)
        out.append((lo, hi))
    return out


def binned_counts(times, pha, ebounds, dt, emin, emax):
    """light_curve_storage.py:212-294 with tmin = tmax = None"""
    tmin, tmax = times.min(), times.max()
    assert tmin < tmax
    bins = np.arange(tmin, tmax, dt)
    idx = select_channel(ebounds, emin, emax, pha)
    counts, edges = np.histogram(times[idx], bins=bins)
    rate = counts / dt
    xbins = np.vstack([edges[:-1], edges[1:]]).T
    return xbins, (rate * dt).astype(int)


def dead_time_per_interval(times, pha, bins):
    """gbm_lightcurve_analyzer.py:76-78,162-196,246-262: sum of the per-event dead times of the events with
    start <= t <= stop, for every bin"""
    per_event = np.where(pha == 127, 10.0e-6, 2.0e-6)
    lo = np.searchsorted(times, bins[:, 0], side="left")
    hi = np.searchsorted(times, bins[:, 1], side="right")
    csum = np.concatenate(([0.0], np.cumsum(per_event)))
    n_ovf = np.concatenate(([0], np.cumsum(pha == 127)))
    n_all = hi - lo
    k_ovf = n_ovf[hi] - n_ovf[lo]
    # exact-in-counts form: (n - k) * 2 us + k * 10 us
    return (n_all - k_ovf) * 2.0e-6 + k_ovf * 10.0e-6, csum


def run_trigger(n_bins_background, n_bins_source, n_bins_pre, starts, stops, counts, exposure, threshold):
    """gbm_lightcurve_analyzer.py:199-243 (parameter names as DECLARED).  Returns (detected, time, razor)."""
    n = len(counts)
    span = n_bins_background + n_bins_pre + n_bins_source
    n_i = max(0, min(n, len(stops) - span))          # i + span < len(stops)
    if n_i <= 0:
        return False, 0.0, False
    cc = np.concatenate(([0], np.cumsum(counts.astype(np.int64))))
    ce = np.concatenate(([0.0], np.cumsum(exposure)))
    i = np.arange(n_i)
    bkg_exposure = ce[i + n_bins_background] - ce[i]
    background_counts = cc[i + n_bins_background] - cc[i]
    src_idx = i + n_bins_background + n_bins_pre
    src_exposure = ce[src_idx + n_bins_source] - ce[src_idx]
    src_counts = cc[src_idx + n_bins_source] - cc[src_idx]
    with np.errstate(all="ignore"):
        alpha = src_exposure / bkg_exposure
        nb = alpha * background_counts
        sig = (src_counts - nb) / np.sqrt(nb)
    ok = (sig >= threshold) & (starts[src_idx] > -1.0)
    razor = bool(np.any(np.abs(sig - threshold) < 1e-9 * threshold))
    hit = np.nonzero(ok)[0]
    if len(hit):
        return True, float(starts[src_idx[hit[0]]]), razor
    return False, 0.0, razor


def analyze(times, pha, ebounds, n_counts_source, threshold=4.5, background_duration=17.0, pre_window=4.0):
    """GBMLightCurveAnalyzer(lightcurve).  Returns dict(is_detected, detection_time, detection_time_scale,
    razor, plus the intermediate products)."""
    times = np.asarray(times, dtype="f8")
    pha = np.asarray(pha).astype(np.int64)
    out = dict(is_detected=False, detection_time=None, detection_time_scale=None, razor=False)
    if n_counts_source <= 0:                                         # lightcurve_analyzer.py:22-34
        return out
    n_bins_background = int(np.floor(background_duration / BASE_TIMESCALE))
    n_bins_pre = int(np.floor(pre_window / BASE_TIMESCALE))
    bins, _ = binned_counts(times, pha, ebounds, BASE_TIMESCALE, None, None)
    starts, stops = bins[:, 0], bins[:, 1]
    dead, _ = dead_time_per_interval(times, pha, bins)
    exposure = (stops - starts) - dead                               # time_interval.py:31
    out.update(starts=starts, stops=stops, exposure=exposure, counts={})
    for key, emin, emax in TRIGGER_ENERGY_RANGES:
        _, counts = binned_counts(times, pha, ebounds, BASE_TIMESCALE, emin, emax)
        out["counts"][key] = counts
        if counts.sum() == 0:
            continue
        for time_scale in TRIGGER_TIME_SCALES[key]:
            n_bins_src = int(np.floor(time_scale / BASE_TIMESCALE))
            # called as _run_trigger(n_bins_background, n_bins_pre, n_bins_src, ...): see the header
            detected, t, razor = run_trigger(n_bins_background, n_bins_pre, n_bins_src, starts, stops, counts,
                                             exposure, threshold)
            out["razor"] |= razor
            if detected:
                out.update(is_detected=True, detection_time=t, detection_time_scale=time_scale)
                return out
    return out


def gbm_trigger(angles, results, simul_trigger_window=0.5, max_n_dets=12):
    """GBMTrigger.process (gbm_trigger.py:52-185).  ``angles``: {detector name: separation angle} of the NaI
    detectors, ``results``: {name: analyze() dict}.  Returns (is_detected, triggered detectors, times, scales)."""
    names = np.array([n for n in angles if n.startswith("n")])
    order = np.argsort(np.array([angles[n] for n in names]))
    names = names[order]
    is_detected, n_triggered, n_tested = False, 0, 0
    t_dets, t_times, t_scales = [], [], []
    while n_tested < len(names) and not is_detected and n_tested < max_n_dets:
        r = results[names[n_tested]]
        if r["is_detected"]:
            if n_triggered != 0:
                t = r["detection_time"]
                if any((t >= o - simul_trigger_window) and (t <= o + simul_trigger_window) for o in t_times):
                    is_detected = True
            t_dets.append(str(names[n_tested]))
            t_times.append(r["detection_time"])
            t_scales.append(r["detection_time_scale"])
            n_triggered += 1
        n_tested += 1
    return is_detected, t_dets, t_times, t_scales
