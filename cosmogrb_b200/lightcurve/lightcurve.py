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
        self._times = None          # the total list after process() (lightcurve.py:189-190)
        self._pha = None

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
        self._times, self._pha = times, pha
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


#: light curves of one call are cut into chunks of about this many event slots when the call is large:
#: the device->host copy of chunk k then overlaps the kernels of chunk k+1
PIPELINE_MIN_SLOTS = 8_000_000
import os as _os
PIPELINE_CHUNKS = int(_os.environ.get("CGRB_PIPELINE_CHUNKS", 7))      # chunks of a large call (device->host copies of chunk k overlap the kernels of k+1)


def _chunks(units, eng):
    """Positions of the units grouped into pipeline chunks (one chunk = everything for small calls)."""
    from ..engine import capacity

    n = len(units)
    slots = (capacity(units["fmax"], units["src_tstop"] - units["src_tstart"])
             + capacity(units["bkg_rate"], units["bkg_tstop"] - units["bkg_tstart"])).astype("f8")
    total = float(slots.sum())
    if n < 2 or not np.isfinite(total) or total < 2 * PIPELINE_MIN_SLOTS:
        return [np.arange(n)]
    target = max(total / PIPELINE_CHUNKS, PIPELINE_MIN_SLOTS)
    chunks, cur, acc = [], [], 0.0
    for i in range(n):
        cur.append(i)
        acc += slots[i]
        if acc >= target:
            chunks.append(np.array(cur))
            cur, acc = [], 0.0
    if cur:
        chunks.append(np.array(cur))
    return chunks


def process_batch(lightcurves, seed=None, stream_ids=None, pinned=None):
    """Run ``LightCurve.process()`` for many light curves as batched launches and return their storages
    (same order).  ``stream_ids`` fixes the Philox stream of every unit (default: position).  Large calls
    are pipelined: results do not depend on the chunking (Philox counters are per unit)."""
    import torch

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
    all_deadtime = all(lc._deadtime for lc in lcs)
    extra = () if all_deadtime else (("times_tot", "tot"), ("pha_tot", "tot"))
    rng = Philox(eng.next_seed() if seed is None else seed)
    chunks = [np.arange(len(lcs))] if (need_fmax or pinned is not None) else _chunks(units, eng)
    if len(chunks) == 1:
        b = eng.new_batch(units, tabs, compute_fmax=need_fmax)
        eng.process(b, rng, keep_merged=not all_deadtime)
        arrays = eng.to_host(b, pinned, extra=extra)
        return [lc._storage_from(arrays, i) for i, lc in enumerate(lcs)]
    # pipelined: kernels of chunk k+1 are enqueued before the host waits for chunk k's counters
    dtab = eng.upload_tables(tabs)
    compute = torch.cuda.current_stream(eng.device)
    copy_stream = eng.copy_stream()
    pending, results = [], [None] * len(lcs)

    def drain(item):
        pos, b, ev = item
        arrays = eng.to_host_async(b, ev, copy_stream, extra=extra)
        return pos, b, arrays

    drained = []
    for pos in chunks:
        b = eng.new_batch(units[pos], dtab, compute_fmax=False)
        eng.process(b, rng, keep_merged=not all_deadtime)
        ev = torch.cuda.Event()
        ev.record(compute)
        pending.append((pos, b, ev))
        if len(pending) > 1:
            drained.append(drain(pending.pop(0)))
    while pending:
        drained.append(drain(pending.pop(0)))
    copy_stream.synchronize()
    for pos, b, arrays in drained:
        for j, i in enumerate(pos):
            results[int(i)] = lcs[int(i)]._storage_from(arrays, j)
    return results
