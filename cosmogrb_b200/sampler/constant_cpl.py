"""Constant cutoff power law (cosmogrb/sampler/constant_cpl.py:15-111; numba kernels
cpl_constant_functions.py replaced by the CUDA engine, unit kind 1)."""
from .source_function import SourceFunction


class ConstantCPL(SourceFunction):
    def __init__(self, peak_flux=1e-6, ep=300.0, alpha=-1.0, emin=10.0, emax=1e4, response=None):
        self._peak_flux = peak_flux
        self._ep = ep
        self._alpha = alpha
        assert alpha < 0.0, "the rejection sampler is slow if alpha is positive"     # constant_cpl.py:33-35
        super(ConstantCPL, self).__init__(emin=emin, emax=emax, index=alpha, response=response)

    def unit_fields(self):
        return dict(kind=1, peak_flux=self._peak_flux, ep_start=self._ep, ep_tau=1.0, alpha=self._alpha,
                    trise=1.0, tdecay=1.0)
