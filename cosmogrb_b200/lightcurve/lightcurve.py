"""Per-detector pipeline ``LightCurve.process()`` (cosmogrb/lightcurve/lightcurve.py:9-283).

The reference runs source times -> energies -> channels, background times -> channels, combine,
dead time as six numba/NumPy stages on the host.  Here the whole sequence is one batch of CUDA
launches; ``process()`` runs a batch of one unit and ``process_batch`` (used by ``GRB.go`` and
``Universe.go``) runs many light curves in one batch.
"""
import numpy as np

from .. import _capi
from ..engine import Philox, get_engine
from .light_curve_storage import LightCurveStorage


class LightCurve(object):
    #: dead-time model applied by the engine; the base class applies none (lightcurve.py:265-267)
    _deadtime = False

    def __init__(self, name, source, background, response, instrument, T0=0, grb_name="SynthGRB",
                 tstart=0, tstop=100.0):
        self._source = source
        self._background = background
        self._response = response
        self._tstart = tstart
        self._tstop = tstop
        assert tstop > tstart
        self._time_adjustment = 0
        self._name = name
        self._grb_name = grb_name
        self._T0 = T0
        self._instrument = instrument
        self._extra_info = {}
        self._lc_storage = None

    def set_time_adjustment(self, t):
        self._time_adjustment = t

    # -- engine description -----------------------------------------------------------------------
    def unit_record(self, det_index=0, stream_id=0):
        """The ``cgrb_unit_t`` record of this light curve."""
        src, bkg = self._source, self._background
        u = src.source_function._unit(src_tstart=src.tstart, src_tstop=src.tstop,
                                      fmax=src.fmax if src.fmax is not None else 0.0)
        u["det"], u["stream_id"] = det_index, stream_id
        u["bkg_tstart"], u["bkg_tstop"], u["bkg_rate"] = bkg.tstart, bkg.tstop, bkg.background_rate
        return u

    def device_table(self, interp_extrap):
        tmpl = self._background.spectrum_template
        return self._response.device_table(None if tmpl is None else tmpl.weights, interp_extrap)

    def _storage_from(self, arrays, i):
        def sl(name):
            # views of the (pinned) staging buffers: no second host copy
            o = arrays[name + "_offsets"]
            return arrays[name][int(o[i]):int(o[i + 1])].numpy()

        pha, times = sl("pha_out"), sl("times_out")
        if not self._deadtime:
            pha, times = sl("pha_tot"), sl("times_tot")
        return LightCurveStorage(
            name=self._name, tstart=self._tstart, tstop=self._tstop, time_adjustment=self._time_adjustment,
            pha=pha, times=times, pha_source=sl("pha_src"), times_source=sl("times_src"),
            pha_background=sl("pha_bkg"), times_background=sl("times_bkg"),
            channels=self._response.channels, ebounds=self._response.channel_edges, T0=self._T0,
            instrument=self._instrument, extra_info=self._extra_info)

    def process(self):
        """source + background sampling, combine, dead time -> ``LightCurveStorage``."""
        return process_batch([self])[0]

    def set_storage(self, lc_storage):
        self._lc_storage = lc_storage

    def _filter_deadtime(self):
        pass

    time_adjustment = property(lambda self: self._lc_storage.time_adjustment)
    times = property(lambda self: self._lc_storage.times)
    pha = property(lambda self: self._lc_storage.pha)
    pha_source = property(lambda self: self._lc_storage.pha_source)
    times_source = property(lambda self: self._lc_storage.times_source)
    pha_background = property(lambda self: self._lc_storage.pha_background)
    times_background = property(lambda self: self._lc_storage.times_background)
    name = property(lambda self: self._name)
    response = property(lambda self: self._response)
    extra_info = property(lambda self: self._extra_info)
    lightcurve_storage = property(lambda self: self._lc_storage)


def process_batch(lightcurves, seed=None, stream_ids=None, pinned=None):
    """Run ``LightCurve.process()`` for many light curves as ONE batch of launches and return their
    storages (same order).  ``stream_ids`` fixes the Philox stream of every unit (default: position)."""
    eng = get_engine()
    lcs = list(lightcurves)
    if not lcs:
        return []
    tabs, tab_index, units = [], {}, []
    for i, lc in enumerate(lcs):
        tab = lc.device_table(eng.interp_extrap)
        k = id(tab)
        if k not in tab_index:
            tab_index[k] = len(tabs)
            tabs.append(tab)
        units.append(lc.unit_record(det_index=tab_index[k], stream_id=i if stream_ids is None else stream_ids[i]))
    units = np.concatenate(units)
    need_fmax = bool(np.any(~(units["fmax"] > 0)))
    b = eng.new_batch(units, tabs, compute_fmax=need_fmax)
    all_deadtime = all(lc._deadtime for lc in lcs)
    eng.process(b, Philox(eng.next_seed() if seed is None else seed), keep_merged=not all_deadtime)
    extra = () if all_deadtime else (("times_tot", "tot"), ("pha_tot", "tot"))
    arrays = eng.to_host(b, pinned, extra=extra)
    return [lc._storage_from(arrays, i) for i, lc in enumerate(lcs)]
