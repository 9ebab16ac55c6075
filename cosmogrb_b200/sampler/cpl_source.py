"""Time-evolving cutoff power law (cosmogrb/sampler/cpl_source.py:15-132; numba kernels
cpl_functions.py replaced by the CUDA engine, unit kind 0)."""
from .source_function import SourceFunction


class CPLSourceFunction(SourceFunction):
    def __init__(self, peak_flux=1e-6, ep_start=300.0, ep_tau=1.0, alpha=-1.0, trise=1.0, tdecay=2,
                 emin=10.0, emax=1e4, response=None):
        self._peak_flux = peak_flux
        self._ep_start = ep_start
        self._ep_tau = ep_tau
        self._trise = trise
        self._tdecay = tdecay
        self._alpha = alpha
        assert alpha < 0.0, "the rejection sampler is slow if alpha is positive"     # cpl_source.py:37-39
        super(CPLSourceFunction, self).__init__(emin=emin, emax=emax, index=alpha, response=response)

    def unit_fields(self):
        return dict(kind=0, peak_flux=self._peak_flux, ep_start=self._ep_start, ep_tau=self._ep_tau,
                    alpha=self._alpha, trise=self._trise, tdecay=self._tdecay)
