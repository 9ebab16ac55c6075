"""``SourceFunction`` ABC (cosmogrb/sampler/source_function.py:8-173) over the CUDA engine.

Concrete classes describe themselves to the engine with ``unit_fields()`` (the spectral part of a
``cgrb_unit_t`` record); the numerical work -- ``evolution``, ``energy_integrated_evolution``,
``sample_events``, ``sample_energy`` -- runs on the GPU through the C ABI.
"""
import abc

import numpy as np

from .. import _capi
from ..engine import Philox, get_engine, make_units


class SourceFunction(object, metaclass=abc.ABCMeta):
    def __init__(self, emin=10.0, emax=1.0e4, index=None, response=None):
        self._index = index
        self._response = response
        self._source = None
        self._z = None
        # like the reference the energy range is the response's (source_function.py:24-31)
        self._emin = response.emin if response is not None else emin
        self._emax = response.emax if response is not None else emax

    def set_response(self, response):
        self._response = response
        self._emin = response.emin
        self._emax = response.emax

    @property
    def response(self):
        return self._response

    def set_source(self, source):
        self._source = source
        self._z = source.z

    # -- description for the engine ---------------------------------------------------------------
    @abc.abstractmethod
    def unit_fields(self):
        """dict of cgrb_unit_t fields: kind, peak_flux, ep_start, ep_tau, alpha, trise, tdecay."""

    def _unit(self, **extra):
        u = make_units(1)
        for k, v in self.unit_fields().items():
            u[k] = v
        assert self._z is not None, "the source function is not attached to a Source (no redshift)"
        u["z"] = self._z
        for k, v in extra.items():
            u[k] = v
        return u

    def _tables(self, engine):
        assert self._response is not None, "no response set"
        return [self._response.device_table(None, engine.interp_extrap)]

    # -- numerical interface ------------------------------------------------------------------------
    def evolution(self, energy, time):
        """photon flux [n_time, n_energy] (reference: cpl_evolution)"""
        eng = get_engine()
        return eng.evolution(self._unit(), np.atleast_1d(energy), np.atleast_1d(time))

    def energy_integrated_evolution(self, time):
        """counts/s integrated over the response (reference: 75-point trapezoid); scalar in -> scalar out"""
        eng = get_engine()
        out = eng.energy_integrated_evolution(self._unit(), self._tables(eng), np.atleast_1d(time))
        return float(out[0]) if np.ndim(time) == 0 or len(out) == 1 else out

    def time_integrated_spectrum(self, energy, t1, t2):
        # the reference's numba implementation cannot be called (cpl_functions.py:349-361 omits z)
        raise NotImplementedError("time_integrated_spectrum is broken in the reference and not on the hot path")

    def sample_events(self, tstart, tstop, fmax):
        """NHPP thinning (reference: sample_events): ascending times, tstart prepended."""
        eng = get_engine()
        u = self._unit(src_tstart=tstart, src_tstop=tstop, fmax=fmax, flags=_capi.UNIT_NO_BACKGROUND)
        b = eng.new_batch(u, self._tables(eng), compute_fmax=False)
        eng.sample_events(b, Philox(eng.next_seed()))
        eng.fetch_counts(b)
        eng.check_status(b)
        return b.view("times_src", 0, "src")

    def sample_energy(self, times):
        """CPL photon energies at the given arrival times (reference: sample_energy)."""
        eng = get_engine()
        times = np.ascontiguousarray(times, dtype="f8")
        n = len(times)
        if n == 0:
            return np.zeros(0)
        u = self._unit(src_tstart=float(times.min()), src_tstop=float(max(times.max(), times.min() + 1e-9)),
                       src_cap=n, flags=_capi.UNIT_NO_BACKGROUND)
        # the energy kernels only need ec(t) over the photons' time range, not fmax
        b = eng.new_batch(u, self._tables(eng), compute_fmax=False)
        eng.set_counts(b, n_src=n)
        ts = np.zeros(b.n_src_slots)
        ts[:n] = times
        b.t["times_src"] = eng.to_device(ts)
        eng.sample_energy(b, Philox(eng.next_seed()))
        eng.fetch_counts(b)
        eng.check_status(b)
        return b.t["energy_src"][:n].cpu().numpy()

    @property
    def index(self):
        return self._index

    @property
    def emin(self):
        return self._emin

    @property
    def emax(self):
        return self._emax
