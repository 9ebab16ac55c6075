"""GBM trigger analysis of a light curve (cosmogrb/instruments/gbm/gbm_lightcurve_analyzer.py:13-271).

The reference bins the total list at 16 ms in four energy ranges, computes the dead time of every bin and
slides a background / source window pair over the bins for each trigger time scale (numba ``_run_trigger``).
Here the whole analysis is two CUDA launches on the event list (``cgrb_gbm_trigger``); the batched entry
points (``trigger_batch``) run it for every detector of every burst of a batch without the events leaving HBM.
"""
import numpy as np

from ...lightcurve.light_curve_storage import LightCurveStorage

_base_timescale = 0.016

_trigger_energy_ranges = {"a": (50.0, 300.0), "b": (25.0, 50.0), "c": (100.0, None), "d": (300.0, None)}

_trigger_time_scales = {
    "a": [_base_timescale * (2 ** i) for i in range(10)][::-1],
    "b": [_base_timescale * (2 ** i) for i in range(10)][::-1],
    "c": [_base_timescale * (2 ** i) for i in range(9)][::-1],
    "d": [_base_timescale * (2 ** i) for i in range(4)][::-1],
}


def channel_bounds(ebounds):
    """int32 [8]: for trigger energy range r a channel c is selected iff b[2r] < c < b[2r+1]
    (``LightCurveStorage._select_channel``, light_curve_storage.py:154-182: strict comparisons against
    ``ebounds.searchsorted(emin / emax)``; a missing bound never rejects)."""
    ebounds = np.asarray(ebounds, dtype="f8")
    out = np.empty(8, dtype=np.int32)
    for r, (emin, emax) in enumerate(_trigger_energy_ranges.values()):
        out[2 * r] = int(ebounds.searchsorted(emin)) if emin is not None else -1
        out[2 * r + 1] = int(ebounds.searchsorted(emax)) if emax is not None else (1 << 30)
    return out


class LightCurveAnalyzer(object):
    """cosmogrb/lightcurve/lightcurve_analyzer.py:9-60"""

    def __init__(self, lightcurve, instrument):
        assert isinstance(lightcurve, LightCurveStorage)
        assert lightcurve.instrument == instrument, \
            f"The lightcurve was not created for {instrument} but for {lightcurve.instrument}"
        self._lightcurve = lightcurve
        self._is_detected = False
        if self._lightcurve.n_counts_source > 0:
            self._process_dead_time()
            self._compute_detection()

    is_detected = property(lambda self: self._is_detected)


class GBMLightCurveAnalyzer(LightCurveAnalyzer):
    def __init__(self, lightcurve, threshold=4.5, background_duration=17, pre_window=4.0):
        self._threshold = threshold
        self._background_duration = background_duration
        self._pre_window = pre_window
        self._base_timescale = _base_timescale
        self._trigger_time_scales = _trigger_time_scales
        self._trigger_energy_ranges = _trigger_energy_ranges
        self._n_bins_background = int(np.floor(self._background_duration / self._base_timescale))
        self._n_bins_pre = int(np.floor(self._pre_window / self._base_timescale))
        self._detection_time = None
        self._detection_time_scale = None
        self._result = None
        super(GBMLightCurveAnalyzer, self).__init__(lightcurve, instrument="GBM")

    def _process_dead_time(self):
        pass        # per-event dead times (2 us / 10 us, :246-262) are applied inside the device kernel

    def _compute_detection(self):
        from ...engine import get_engine

        lc = self._lightcurve
        r = get_engine().analyze_light_curve(
            lc.times, lc._raw["pha"], channel_bounds(lc.ebounds), lc.n_counts_source, threshold=self._threshold,
            background_duration=float(self._background_duration), pre_window=float(self._pre_window),
            base_timescale=self._base_timescale)
        self._result = r
        if r["detected"]:
            self._is_detected = True
            self._detection_time = float(r["time"])
            self._detection_time_scale = float(r["time_scale"])

    detection_time_scale = property(lambda self: self._detection_time_scale)
    detection_time = property(lambda self: self._detection_time)


def simultaneous_trigger(names, angles, detected, times, time_scales, simul_trigger_window=0.5, max_n_dets=12):
    """The detector loop of ``GBMTrigger.process`` (gbm_trigger.py:119-185): NaI detectors in order of their
    angular distance to the burst; the burst is detected when a second detector fires within
    +-``simul_trigger_window`` of an earlier trigger.  Returns (is_detected, names, times, time scales)."""
    idx = [i for i, n in enumerate(names) if str(n).startswith("n")]
    order = np.argsort(np.asarray([angles[i] for i in idx], dtype="f8"))
    is_detected, n_triggered, n_tested = False, 0, 0
    t_dets, t_times, t_scales = [], [], []
    while n_tested < len(order) and not is_detected and n_tested < max_n_dets:
        i = idx[order[n_tested]]
        if detected[i]:
            t = float(times[i])
            if n_triggered != 0 and any((t >= o - simul_trigger_window) and (t <= o + simul_trigger_window)
                                        for o in t_times):
                is_detected = True
            t_dets.append(str(names[i]))
            t_times.append(t)
            t_scales.append(float(time_scales[i]))
            n_triggered += 1
        n_tested += 1
    return is_detected, t_dets, t_times, t_scales


__all__ = ["GBMLightCurveAnalyzer", "LightCurveAnalyzer", "channel_bounds", "simultaneous_trigger"]
