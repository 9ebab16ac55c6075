"""Detector response: effective-area curve, row-cumulative matrix and energy -> PHA channel draw.

Host-side mirror of ``cosmogrb/response/response.py`` (class ``Response`` :28-266).  The small
per-detector tables are built once on the host with NumPy -- the same array operations as the
reference (:74-154), hence bit-identical tables -- and packed into the device table block the
CUDA kernels read (layout in include/cosmogrb_b200.h).  ``digitize`` (:156-174, numba
``_digitize`` :10-25) runs on the GPU.
"""
from __future__ import annotations

import numpy as np

from .. import _capi


def interp_table(x, y, v, interp_extrap="linear"):
    """``interpolation.interp(x, y, v)`` as called at cpl_functions.py:119.

    ``linear``: bracket index clamped to the table, weight left free (linear extrapolation,
    EconForge ``interpolation`` behaviour); ``clamp``: ``np.interp``.
    """
    x = np.asarray(x, dtype="f8")
    y = np.asarray(y, dtype="f8")
    v = np.atleast_1d(np.asarray(v, dtype="f8"))
    if interp_extrap == "clamp":
        return np.interp(v, x, y)
    if interp_extrap != "linear":
        raise ValueError("interp_extrap must be 'linear' or 'clamp'")
    i = np.clip(np.searchsorted(x, v, side="left") - 1, 0, len(x) - 2)
    lam = (v - x[i]) / (x[i + 1] - x[i])
    return (1.0 - lam) * y[i] + lam * y[i + 1]


class Interp1D(object):
    """Stand-in for the reference's jitclass ``Interp1D`` (utils/interpolation.py:62-92):
    interpolate, zero outside [xmin, xmax].  Not used by the kernels (they take
    ``effective_area_packed``), kept because ``Response.effective_area`` is public."""

    def __init__(self, x, y, xmin, xmax, interp_extrap="linear"):
        self.x = np.asarray(x, dtype="f8")
        self.y = np.asarray(y, dtype="f8")
        self.xmin = np.float32(xmin)
        self.xmax = np.float32(xmax)
        self._extrap = interp_extrap

    def evaluate(self, v):
        v = np.atleast_1d(np.asarray(v, dtype="f8"))
        out = interp_table(self.x, self.y, v, self._extrap)
        out[(v > self.xmax) | (v < self.xmin)] = 0.0
        return out


class Response(object):
    def __init__(
        self,
        matrix,
        geometric_area,
        energy_edges,
        channel_edges=None,
        channel_starts_at=0,
    ):
        """
        :param matrix: DRM in cm^2, shape (n_photon_bins, n_channels)
        :param geometric_area: geometric area of the detector (normalises the matrix to probabilities)
        :param energy_edges: photon energy edges (n_photon_bins + 1)
        :param channel_edges: channel energy edges (n_channels + 1)
        """
        self._matrix = np.asarray(matrix).astype("f8")
        self._energy_edges = np.asarray(energy_edges).astype("f8")
        self._energy_width = np.diff(self._energy_edges)
        self._energy_mean = (self._energy_edges[:-1] + self._energy_edges[1:]) / 2.0
        self._emin = self._energy_edges[0]
        self._emax = self._energy_edges[-1]

        self._channel_edges = np.asarray(channel_edges).astype("f8")
        self._channel_width = np.diff(self._channel_edges)
        self._channel_mean = (self._channel_edges[:-1] + self._channel_edges[1:]) / 2.0
        self._channels = np.arange(len(self._channel_width), dtype=np.int64)

        if self._matrix.shape != (len(self._energy_mean), len(self._channels)):
            raise ValueError(
                f"matrix shape {self._matrix.shape} does not match "
                f"({len(self._energy_mean)} photon bins, {len(self._channels)} channels)"
            )

        self._geometric_area = geometric_area
        self._build_effective_area_curve()
        self._construct_probabilities()
        self._device_tables = {}

    # -- tables (reference response.py:74-154) ------------------------------------------------
    def _build_effective_area_curve(self):
        ea_curve = self._matrix.sum(axis=1)
        self.effective_area = Interp1D(
            self._energy_mean, ea_curve.copy(), self._energy_mean.min(), self._energy_mean.max()
        )
        ea_curve[~(ea_curve > 0.0)] = 1.0e-99  # "we do not want zeros"
        self.effective_area_packed = np.vstack((self._energy_mean, ea_curve))
        self._max_energy = self._energy_mean[ea_curve[:-10].argmax()]

    def _construct_probabilities(self):
        prob = self._matrix / self._geometric_area
        self._total_probability_per_bin = prob.sum(axis=1)
        nz = np.where(self._total_probability_per_bin > 0)[0]
        prob[nz, :] = prob[nz, :] / self._total_probability_per_bin[nz, np.newaxis]
        self._probability_matrix = prob
        self._normed_probability_matrix = prob
        self._detection_probability = 1 - np.exp(-self._total_probability_per_bin)
        self._cumulative_maxtrix = np.cumsum(prob, axis=1)

    # -- device table block -------------------------------------------------------------------
    def device_table(self, bkg_weights=None, interp_extrap="linear"):
        """Pack the per-detector tables into one float64 block (include/cosmogrb_b200.h, CGRB_DT_*)."""
        nb, nc = self._matrix.shape
        if (nb, nc) != (_capi.NBINS, _capi.NCHAN):
            raise ValueError(f"the engine is compiled for {_capi.NBINS}x{_capi.NCHAN} responses, got {nb}x{nc}")
        key = (interp_extrap, None if bkg_weights is None else np.asarray(bkg_weights).tobytes())
        if key in self._device_tables:
            return self._device_tables[key]
        t = np.zeros(_capi.DT_SIZE, dtype="f8")
        t[_capi.DT_EDGES:_capi.DT_EDGES + nb + 1] = self._energy_edges
        t[_capi.DT_EA_X:_capi.DT_EA_X + nb] = self.effective_area_packed[0]
        t[_capi.DT_EA_Y:_capi.DT_EA_Y + nb] = self.effective_area_packed[1]
        if bkg_weights is not None:
            # numpy legacy choice(): p -> float64, cdf = cumsum(p); cdf /= cdf[-1]  (background.py:93)
            cdf = np.asarray(bkg_weights).astype("f8").cumsum()
            cdf /= cdf[-1]
            t[_capi.DT_BKG_CDF:_capi.DT_BKG_CDF + nc] = cdf
        else:
            t[_capi.DT_BKG_CDF:_capi.DT_BKG_CDF + nc] = np.arange(1, nc + 1) / nc
        t[_capi.DT_BKG_GUIDE:_capi.DT_BKG_GUIDE + _capi.BKG_GUIDE_N // 8] = background_guide(
            t[_capi.DT_BKG_CDF:_capi.DT_BKG_CDF + nc]).view("f8")
        # 10**linspace(log10 emin, log10 emax, n): cpl_functions.py:208,305
        for n, oe, oa in ((_capi.NE_LAMBDA, _capi.DT_G75_E, _capi.DT_G75_A),
                          (_capi.NE_ENVELOPE, _capi.DT_G500_E, _capi.DT_G500_A)):
            grid = np.power(10.0, np.linspace(np.log10(self._emin), np.log10(self._emax), n))
            t[oe:oe + n] = grid
            t[oa:oa + n] = interp_table(self.effective_area_packed[0], self.effective_area_packed[1],
                                        grid, interp_extrap)
        t[_capi.DT_CUM:_capi.DT_CUM + nb * nc] = self._cumulative_maxtrix.reshape(-1)
        t[_capi.DT_GUIDE:_capi.DT_GUIDE + nb * nc // 8] = channel_search_guide(self._cumulative_maxtrix).view("f8")
        self._device_tables[key] = t
        return t

    @property
    def effective_area_max(self):
        return self._max_energy

    def get_photon_bin(self, energy):
        return np.searchsorted(self._energy_edges, energy) - 1

    def digitize(self, photon_energies):
        """Digitize photon energies into PHA channels through the energy dispersion
        (GPU; reference response.py:156-174).  Returns float64 channel numbers like the reference."""
        from ..engine import get_engine

        return get_engine().digitize(self, np.asarray(photon_energies, dtype="f8"))

    @property
    def energy_edges(self):
        return self._energy_edges

    @property
    def emin(self):
        return self._emin

    @property
    def emax(self):
        return self._emax

    @property
    def channel_edges(self):
        return self._channel_edges

    @property
    def channels(self):
        return self._channels

    @property
    def matrix(self):
        return self._matrix

    @property
    def geometric_area(self):
        return self._geometric_area

    @property
    def cumulative_matrix(self):
        return self._cumulative_maxtrix


def _response_set_function(self, integral_function=None):
    """``integral_function(e1, e2, t1, t2)`` -> integral of the model over [e1, e2] x [t1, t2] (response.py:202-211)."""
    self._integral_function = integral_function


def _response_convolve(self, t1, t2):
    """Expected counts per channel over [t1, t2]: the model integrated over every photon bin, folded through the
    matrix (response.py:213-236; host-side, not on the sampling path).  Non-finite bin fluxes count as zero."""
    e = self._energy_edges
    true_fluxes = np.array([self._integral_function(e[i], e[i + 1], t1, t2) for i in range(len(e) - 1)], dtype="f8")
    true_fluxes[~np.isfinite(true_fluxes)] = 0.0
    return np.dot(true_fluxes, self._matrix)


def _response_to_fits(self, filename, telescope_name="telescope", instrument_name="detector", overwrite=False):
    """Write the matrix as an OGIP ``.rsp`` FITS file (response.py:238-263)."""
    from ..utils.response_file import RSP

    RSP(self._energy_edges, self._channel_edges, self.matrix.T, telescope_name, instrument_name).writeto(
        filename, overwrite=overwrite)


def _response_from_fits(cls, filename, geometric_area, extension=None):
    """Build a ``Response`` from an OGIP response file (SPECRESP MATRIX + EBOUNDS, grouped or plain, e.g. a
    real GBM ``.rsp`` / one matrix of an ``.rsp2``): the real-DRM ingestion route of SURVEY.md §8 f4 that
    needs no ``gbm_drm_gen``.  Not in the reference (its responses come from ``DRMGenTTE``)."""
    from ..utils.response_file import read_response

    r = read_response(filename, extension=extension)
    return cls(matrix=r["matrix"], geometric_area=geometric_area, energy_edges=r["energy_edges"],
               channel_edges=r["channel_edges"])


Response.set_function = _response_set_function
Response.convolve = _response_convolve
Response.to_fits = _response_to_fits
Response.from_fits = classmethod(_response_from_fits)


def background_guide(cdf):
    """uint8 [BKG_GUIDE_N] guide of the device background-channel draw (include/cosmogrb_b200.h, CGRB_DT_BKG_GUIDE):
    guide[m] = min(searchsorted(cdf, m / N, 'right'), NCHAN - 1) -- a start at or before the channel of every
    u in [m / N, (m + 1) / N); the device walks forward from it (numpy's legacy ``choice``, background.py:93)"""
    n = _capi.BKG_GUIDE_N
    return np.minimum(np.searchsorted(cdf, np.arange(n) / float(n), "right"), len(cdf) - 1).astype(np.uint8)


def channel_search_guide(cumulative_matrix):
    """uint8 [NBINS*NCHAN] guide of the device channel search (include/cosmogrb_b200.h, CGRB_DT_GUIDE):
    guide[b][m] = np.searchsorted(cum[b], m / NCHAN, 'left'), the first channel whose cumulative
    probability reaches m/NCHAN (NCHAN if none).  The lower bound of any r in [m/NCHAN, (m+1)/NCHAN) is
    at or after it, so the device walks forward from there and finds exactly what a full search finds."""
    cum = np.asarray(cumulative_matrix, dtype="f8")
    nb, nc = cum.shape
    marks = np.arange(nc, dtype="f8") / nc
    g = np.empty((nb, nc), dtype=np.uint8)
    for b in range(nb):
        g[b] = np.searchsorted(cum[b], marks, side="left")
    return np.ascontiguousarray(g.reshape(-1))


__all__ = ["Response", "interp_table", "Interp1D", "channel_search_guide"]
