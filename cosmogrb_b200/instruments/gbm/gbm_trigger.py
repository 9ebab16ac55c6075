"""``GBMTrigger`` (cosmogrb/instruments/gbm/gbm_trigger.py:14-185; base ``GRBDetector``
cosmogrb/grb/grb_detector.py): run the GBM trigger on a saved burst."""
import numpy as np

from ...io.grb_save import GRBSave
from .gbm_lightcurve_analyzer import GBMLightCurveAnalyzer, simultaneous_trigger


class GBMTrigger(object):
    def __init__(self, grb_save_file_name, threshold=4.5, simul_trigger_window=0.5, max_n_dets=12):
        self._grb_save = grb_save_file_name if isinstance(grb_save_file_name, GRBSave) \
            else GRBSave.from_file(grb_save_file_name)
        self._instrument = "GBM"
        self._is_detected = False
        self._extra_info = {}
        self._threshold = threshold
        self._simul_trigger_window = simul_trigger_window
        self._max_n_dets = max_n_dets
        self._triggered_times = []
        self._triggered_detectors = []
        self._triggered_time_scales = []
        self._setup_order_by_distance()

    def _setup_order_by_distance(self):
        names = [n for n in self._grb_save.keys() if n.startswith("n")]
        ang = np.array([self._grb_save[n]["lightcurve"].extra_info["angle"] for n in names], dtype="f8")
        self._lc_names = np.array(names)[ang.argsort()]

    is_detected = property(lambda self: self._is_detected)
    triggered_detectors = property(lambda self: self._triggered_detectors)
    triggered_times = property(lambda self: self._triggered_times)
    triggered_time_scales = property(lambda self: self._triggered_time_scales)
    extra_info = property(lambda self: self._extra_info)
    name = property(lambda self: self._grb_save.name)

    def process(self):
        """two simultaneous triggers (gbm_trigger.py:119-185); detectors are analysed lazily, in order,
        exactly as far as the reference would go"""
        n_triggered = n_tested = 0
        while n_tested < len(self._lc_names) and not self._is_detected and n_tested < self._max_n_dets:
            name = self._lc_names[n_tested]
            a = GBMLightCurveAnalyzer(self._grb_save[name]["lightcurve"], threshold=self._threshold)
            if a.is_detected:
                if n_triggered != 0:
                    t = a.detection_time
                    if any((t >= o - self._simul_trigger_window) and (t <= o + self._simul_trigger_window)
                           for o in self._triggered_times):
                        self._is_detected = True
                self._triggered_detectors.append(str(name))
                self._triggered_times.append(a.detection_time)
                self._triggered_time_scales.append(a.detection_time_scale)
                n_triggered += 1
            n_tested += 1
        self._extra_info["triggered_detectors"] = np.array(self._triggered_detectors, dtype="S10")
        self._extra_info["triggered_times"] = np.array(self._triggered_times)
        self._extra_info["triggered_time_scales"] = np.array(self._triggered_time_scales)


__all__ = ["GBMTrigger", "simultaneous_trigger"]
