"""``LightCurveStorage``: the output container of the per-detector hot path.  Constructor contract
and accessors as cosmogrb/lightcurve/light_curve_storage.py:16-152; of the analysis helpers only
the numeric ones (channel/time selection, binned counts :154-294) are kept -- plotting is out of scope.
"""
import numpy as np


class LightCurveStorage(object):
    def __init__(self, name, tstart, tstop, time_adjustment, pha, times, pha_source, times_source,
                 pha_background, times_background, channels, ebounds, T0, instrument, extra_info):
        self._name = name
        self._tstart = tstart
        self._tstop = tstop
        self._time_adjustment = time_adjustment
        # the event arrays may be views of pinned staging buffers in the engine's compact dtypes
        # (u8 channels); the reference's dtypes -- pha int (light_curve_storage.py:58), pha_source
        # float64, pha_background int64 -- are materialised on first access
        self._raw = dict(pha=pha, pha_source=pha_source, pha_background=pha_background)
        self._conv = {}
        self._times = np.asarray(times)
        self._n_counts = self._times.shape[0]
        self._times_source = np.asarray(times_source)
        self._n_counts_source = self._times_source.shape[0]
        self._times_background = np.asarray(times_background)
        self._n_counts_background = self._times_background.shape[0]
        self._channels = channels
        self._ebounds = ebounds
        self._T0 = T0
        self._instrument = instrument
        self._extra_info = extra_info

    name = property(lambda self: self._name)
    instrument = property(lambda self: self._instrument)
    tstart = property(lambda self: self._tstart)
    tstop = property(lambda self: self._tstop)
    n_counts = property(lambda self: self._n_counts)
    n_counts_source = property(lambda self: self._n_counts_source)
    n_counts_background = property(lambda self: self._n_counts_background)
    T0 = property(lambda self: self._T0)
    channels = property(lambda self: self._channels)
    ebounds = property(lambda self: self._ebounds)
    times = property(lambda self: self._times)
    pha = property(lambda self: self._get("pha", int))
    pha_source = property(lambda self: self._get("pha_source", np.float64))
    times_source = property(lambda self: self._times_source)
    pha_background = property(lambda self: self._get("pha_background", np.int64))
    times_background = property(lambda self: self._times_background)
    time_adjustment = property(lambda self: self._time_adjustment)
    extra_info = property(lambda self: self._extra_info)

    _REF_DTYPE = dict(pha=int, pha_source=np.float64, pha_background=np.int64)

    def _get(self, name, dtype):
        if name not in self._conv:
            a = np.asarray(self._raw[name])
            self._conv[name] = a if a.dtype == np.dtype(dtype) else a.astype(dtype)
        return self._conv[name]

    @property
    def _pha(self):
        return self.pha

    @property
    def _pha_source(self):
        return self.pha_source

    @property
    def _pha_background(self):
        return self.pha_background

    def _select_channel(self, emin, emax, pha, original_idx=None):
        if original_idx is None:
            original_idx = np.ones_like(pha, dtype=bool)
        if emin is not None:
            original_idx = np.logical_and(original_idx, pha > self._ebounds.searchsorted(emin))
        if emax is not None:
            original_idx = np.logical_and(original_idx, pha < self._ebounds.searchsorted(emax))
        return original_idx

    def _select_time(self, tmin, tmax, times, original_idx=None):
        idx = (times >= tmin) & (times <= tmax)
        return idx if original_idx is None else np.logical_and(idx, original_idx)

    def get_idx_over_interval(self, tmin, tmax):
        return self._select_time(tmin, tmax, self._times)

    def _bin_lightcurve(self, dt, emin, emax, times, pha, tmin=None, tmax=None):
        tmin = times.min() if tmin is None else tmin
        tmax = times.max() if tmax is None else tmax
        assert tmin < tmax, f"the specified tmin and tmax are out of order ({tmin} > {tmax})"
        bins = np.arange(tmin, tmax, dt)
        idx = self._select_channel(emin, emax, pha, np.ones_like(times, dtype=bool))
        counts, edges = np.histogram(times[idx], bins=bins)
        return np.vstack([edges[:-1], edges[1:]]).T, counts / dt

    def binned_counts(self, dt, emin, emax, tmin=None, tmax=None):
        bins, rate = self._bin_lightcurve(dt, emin, emax, self._times, self._pha, tmin, tmax)
        return bins, (rate * dt).astype(int)

    def count_spectrum(self, tmin=None, tmax=None, component="total"):
        """counts per PHA channel (numeric core of _bin_spectrum, :463)."""
        times, pha = {"total": (self._times, self._pha), "source": (self._times_source, self._pha_source),
                      "background": (self._times_background, self._pha_background)}[component]
        idx = np.ones_like(times, dtype=bool)
        if tmin is not None:
            idx &= times >= tmin
        if tmax is not None:
            idx &= times <= tmax
        return np.bincount(np.asarray(pha)[idx].astype(int), minlength=len(self._channels))

    def _output(self, as_display=False):
        """summary table (light_curve_storage.py:605-640): a pandas Series when pandas is importable, else an
        ordered dict; ``as_display`` prints it (the reference hands it to IPython's display)"""
        import collections

        d = collections.OrderedDict()
        d["name"] = self._name
        d["instrument"] = self._instrument
        d["tstart"] = self._tstart
        d["tstop"] = self._tstop
        d["time adjustment"] = self._time_adjustment
        d["T0"] = self._T0
        d["n_counts"] = self._n_counts
        d["n_counts_source"] = self._n_counts_source
        d["n_counts_background"] = self._n_counts_background
        if self._extra_info:
            for k, v in self._extra_info.items():
                d[k] = v
        try:
            import pandas as pd

            out = pd.Series(data=d, index=list(d.keys()))
        except ImportError:          # pragma: no cover
            out = d
        if as_display:
            print(out.to_string() if hasattr(out, "to_string") else "\n".join(f"{k:<22}{v}" for k, v in out.items()))
        return out

    def __repr__(self):
        out = self._output()
        return out.to_string() if hasattr(out, "to_string") else "\n".join(f"{k:<22}{v}" for k, v in out.items())

    def info(self):
        self._output(as_display=True)
