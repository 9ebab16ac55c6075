"""``Source`` (cosmogrb/sampler/source.py:15-147): owns the time range, the redshift and fmax."""
import numpy as np

from ..engine import get_engine
from .sampler import Sampler


class Source(Sampler):
    def __init__(self, tstart, tstop, source_function, z, use_plaw_sample=False):
        self._source_function = source_function
        self._use_plaw_sample = use_plaw_sample
        self._energy_grid = np.logspace(np.log10(source_function.emin), np.log10(source_function.emax), 25)
        self._z = z
        super(Source, self).__init__(tstart=tstart, tstop=tstop)
        self._source_function.set_source(self)
        self._fmax = None
        if self._source_function.response is not None:
            self._fmax = self._get_energy_integrated_max()

    @property
    def z(self):
        return self._z

    @property
    def source_function(self):
        return self._source_function

    @property
    def fmax(self):
        return self._fmax

    def set_response(self, response):
        self._source_function.set_response(response)
        self._fmax = self._get_energy_integrated_max()

    def _get_energy_integrated_max(self):
        """max of lambda(t) on 50 uniformly spaced times (source.py:88-110) -- on the device (cgrb_fmax)."""
        eng = get_engine()
        sf = self._source_function
        u = sf._unit(src_tstart=self._tstart, src_tstop=self._tstop)
        b = eng.new_batch(u, sf._tables(eng), compute_fmax=True)
        return float(b.units["fmax"][0])

    def sample_times(self):
        return self._source_function.sample_events(self._tstart, self._tstop, self._fmax)

    def sample_photons(self, times):
        return self._source_function.sample_energy(times)

    def sample_channel(self, photons, response):
        return response.digitize(photons)
