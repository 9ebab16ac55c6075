"""Synthetic GBM-shaped detector responses (140 photon bins x 128 PHA channels).

The reference builds its DRMs with ``gbm_drm_gen.DRMGenTTE`` + ``gbmgeometry`` + a position
history file (instruments/gbm/gbm_response.py:35-289, response_generator.py:38-138); none of
these can be installed offline, and BASELINE.json config 3 asks for
"synthetic DRMs of GBM shape".  This module is the deterministic generator described in
SURVEY.md §8(d): the engine itself consumes any (140 x 128) matrix.

Shape of the model
  * photon edges: ``logspace(log10 5, log10 5e4, 141)`` keV (NaI), ``logspace(2, log10 6e4, 141)`` (BGO)
  * channel edges: the real 128-channel EBOUNDS of each detector (data/gbm_tables.npz, decoded
    from the reference's ``data/gbm_cspec/<det>.pha`` by scripts/extract_gbm_tables.py)
  * row i = log-normal photopeak (sigma_lnE 0.25 -> 0.08 with energy) + flat Compton tail below it,
    scaled so that ``row.sum()`` = the effective area ``A0 * exp(-((log10 E - c)/0.8)^2 / 2) * |cos theta|``
  * the lowest ``n_zero_rows`` photon bins are all-zero (exercises the zero-row path of digitize)
  * geometric area as gbm_response.py:154-169 (NaI) / :274-289 (BGO)
"""
from __future__ import annotations

import os

import numpy as np

from ...response import Response

GBM_DETECTORS = ("n0", "n1", "n2", "n3", "n4", "n5", "n6", "n7", "n8", "n9", "na", "nb", "b0", "b1")

# detector boresights in spacecraft coordinates (azimuth, zenith) [deg]: the published GBM
# mounting angles (Meegan et al. 2009, table 1)
_BORESIGHT = {
    "n0": (45.89, 20.58), "n1": (45.11, 45.31), "n2": (58.44, 90.21), "n3": (314.87, 45.24),
    "n4": (303.15, 90.27), "n5": (3.35, 89.79), "n6": (224.93, 20.43), "n7": (224.62, 46.18),
    "n8": (236.61, 89.97), "n9": (135.19, 45.55), "na": (123.73, 90.42), "nb": (183.74, 90.32),
    "b0": (0.0, 90.0), "b1": (180.0, 90.0),
}

_TABLES = None


def gbm_tables():
    """detectors[14], bkg_counts[14,128] (f4), channel_edges[14,129] (f8)."""
    global _TABLES
    if _TABLES is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "data", "gbm_tables.npz")
        with np.load(os.path.normpath(path)) as d:
            _TABLES = {k: d[k] for k in d.files}
    return _TABLES


def _unit_vector(az_deg, zen_deg):
    az, zen = np.deg2rad(az_deg), np.deg2rad(zen_deg)
    return np.array([np.sin(zen) * np.cos(az), np.sin(zen) * np.sin(az), np.cos(zen)])


def separation_angle(det, ra, dec):
    """Angle [rad] between the detector boresight and the source.  Without an orbit/attitude
    model the spacecraft frame is taken aligned with the equatorial frame (z = north pole)."""
    src = _unit_vector(ra, 90.0 - dec)
    bore = _unit_vector(*_BORESIGHT[det])
    return float(np.arccos(np.clip(np.dot(src, bore), -1.0, 1.0)))


def angle_factor(det, angle):
    """Viewing-angle factor of the synthetic effective area (vectorised over ``angle`` [rad])."""
    angle = np.asarray(angle, dtype="f8")
    if det[0] == "b":
        return np.maximum(np.abs(np.sin(angle)), 0.3)
    return np.maximum(np.abs(np.cos(angle)), 0.05)


def separation_angles(ra, dec):
    """[n, 14] separation angles [rad] of n sky positions from the 14 boresights."""
    ra, dec = np.deg2rad(np.atleast_1d(ra)), np.deg2rad(np.atleast_1d(dec))
    zen = 0.5 * np.pi - dec
    src = np.stack([np.sin(zen) * np.cos(ra), np.sin(zen) * np.sin(ra), np.cos(zen)], axis=1)
    bore = np.stack([_unit_vector(*_BORESIGHT[d]) for d in GBM_DETECTORS], axis=0)
    return np.arccos(np.clip(src @ bore.T, -1.0, 1.0))


def synthetic_matrix(det, angle, n_zero_rows=2, factor=None):
    """-> (matrix[140,128] cm^2, energy_edges[141], channel_edges[129]).  ``factor`` overrides the
    viewing-angle factor (1.0 = the base, face-on matrix shared by population runs)."""
    tabs = gbm_tables()
    di = GBM_DETECTORS.index(det)
    cedges = np.asarray(tabs["channel_edges"][di], dtype="f8")
    bgo = det[0] == "b"
    if bgo:
        eedges = np.logspace(np.log10(100.0), np.log10(6.0e4), 141)
        centre, a0 = 3.0, 160.0
    else:
        eedges = np.logspace(np.log10(5.0), np.log10(5.0e4), 141)
        centre, a0 = 2.0, 120.0
    emean = 0.5 * (eedges[:-1] + eedges[1:])
    cmid = np.sqrt(cedges[:-1] * cedges[1:])
    cwid = np.diff(cedges)
    lo, hi = np.log10(eedges[0]), np.log10(eedges[-1])
    frac = (np.log10(emean) - lo) / (hi - lo)
    sigma = 0.25 + (0.08 - 0.25) * frac
    # photopeak, integrated over the channel width (density in E times width)
    lnr = np.log(cmid[None, :] / emean[:, None])
    peak = np.exp(-0.5 * (lnr / sigma[:, None]) ** 2) / (cmid[None, :] * sigma[:, None]) * cwid[None, :]
    # flat Compton tail below the photon energy, 20 % of the peak area, growing with energy
    tail = (cmid[None, :] < emean[:, None]) * cwid[None, :]
    tail_sum = tail.sum(axis=1, keepdims=True)
    peak_sum = peak.sum(axis=1, keepdims=True)
    shape = np.where(peak_sum > 0, peak / np.maximum(peak_sum, 1e-300), 0.0) * (1.0 - 0.2 * frac[:, None]) \
        + np.where(tail_sum > 0, tail / np.maximum(tail_sum, 1e-300), 0.0) * (0.2 * frac[:, None])
    # a small deterministic ripple so that detectors are not identical (seed 20240 + d)
    rng = np.random.default_rng(20240 + di)
    shape = shape * (1.0 + 0.02 * rng.standard_normal(shape.shape)).clip(0.9, 1.1)
    rs = shape.sum(axis=1, keepdims=True)
    shape = np.where(rs > 0, shape / np.maximum(rs, 1e-300), 0.0)
    cosfac = float(angle_factor(det, angle)) if factor is None else float(factor)
    ea = a0 * np.exp(-0.5 * ((np.log10(emean) - centre) / 0.8) ** 2) * cosfac
    matrix = shape * ea[:, None]
    matrix[:n_zero_rows, :] = 0.0
    return matrix, eedges, cedges


_base = {}


def base_response(det):
    """Face-on (viewing-angle factor 1) response of a detector: the shared table of population runs,
    where the per-GRB factor travels in ``cgrb_unit_t.ea_scale``."""
    if det not in _base:
        matrix, eedges, cedges = synthetic_matrix(det, 0.0, factor=1.0)
        _base[det] = Response(matrix=matrix, geometric_area=np.pi * 6.35 ** 2, energy_edges=eedges,
                              channel_edges=cedges)
    return _base[det]


class SyntheticGBMResponse(Response):
    """Stand-in for ``NaIResponse`` / ``BGOResponse`` (gbm_response.py:245-289)."""

    def __init__(self, detector_name, ra, dec, radius, height, trigger_time=0.0, name=None):
        self._detector_name = detector_name
        self._ra, self._dec = ra, dec
        self._radius, self._height = radius, height
        self._separation_angle = separation_angle(detector_name, ra, dec)
        self._trigger_time = trigger_time
        self._name = name
        matrix, eedges, cedges = synthetic_matrix(detector_name, self._separation_angle)
        super(SyntheticGBMResponse, self).__init__(
            matrix=matrix,
            geometric_area=self._compute_geometric_area(),
            energy_edges=eedges,
            channel_edges=cedges,
        )

    def _compute_geometric_area(self):
        # gbm_response.py:154-169
        return np.fabs(np.pi * self._radius ** 2 * np.cos(self._separation_angle)) + np.fabs(
            2 * self._radius * self._height * np.sin(self._separation_angle))

    @property
    def detector_name(self):
        return self._detector_name

    @property
    def separation_angle(self):
        return np.rad2deg(self._separation_angle)

    @property
    def trigger_time(self):
        return self._trigger_time

    @property
    def ra(self):
        return self._ra

    @property
    def dec(self):
        return self._dec

    @property
    def T0(self):
        return self._trigger_time


class GBMResponseFile(SyntheticGBMResponse):
    """A GBM detector response read from an OGIP ``.rsp`` file (SPECRESP MATRIX + EBOUNDS; one matrix of an
    ``.rsp2`` via ``extension``) instead of the synthetic generator: the real-DRM route of SURVEY.md §8 f4 without
    ``gbm_drm_gen``.  The matrix must have the engine's 140 photon bins x 128 channels (what GBM responses have).
    Geometry (separation angle, geometric area) follows the detector table exactly as for the synthetic
    responses (gbm_response.py:154-169, 274-289)."""

    def __init__(self, detector_name, ra, dec, file_name, extension=None, trigger_time=0.0, name=None):
        from ...utils.response_file import read_response

        assert detector_name in GBM_DETECTORS, "must use a proper GBM detector n<> or b<>"
        self._detector_name = detector_name
        self._ra, self._dec = ra, dec
        self._radius, self._height = 0.5 * 12.7, (12.7 if detector_name[0] == "b" else 1.27)
        self._separation_angle = separation_angle(detector_name, ra, dec)
        self._trigger_time = trigger_time
        self._name = name
        self._file_name = str(file_name)
        r = read_response(file_name, extension=extension)
        if r["matrix"].shape != (140, 128):
            raise ValueError(f"{file_name}: matrix is {r['matrix'].shape[0]} photon bins x {r['matrix'].shape[1]} "
                             "channels; the engine's tables are built for GBM's 140 x 128")
        Response.__init__(self, matrix=r["matrix"], geometric_area=self._compute_geometric_area(),
                          energy_edges=r["energy_edges"], channel_edges=r["channel_edges"])

    def _compute_geometric_area(self):
        if self._detector_name[0] == "b":
            a = self._separation_angle + 0.5 * np.pi
            return np.fabs(np.pi * self._radius ** 2 * np.cos(a)) + np.fabs(2 * self._radius * self._height * np.sin(a))
        return SyntheticGBMResponse._compute_geometric_area(self)


class NaIResponse(SyntheticGBMResponse):
    def __init__(self, detector_name, ra, dec, rng=None, save=False, name=None):
        super(NaIResponse, self).__init__(detector_name, ra, dec, radius=0.5 * 12.7, height=1.27, name=name)


class BGOResponse(SyntheticGBMResponse):
    def __init__(self, detector_name, ra, dec, rng=None, save=False, name=None):
        super(BGOResponse, self).__init__(detector_name, ra, dec, radius=0.5 * 12.7, height=12.7, name=name)

    def _compute_geometric_area(self):
        # gbm_response.py:274-289 ("this may be wrong" in the reference; kept literally)
        a = self._separation_angle + 0.5 * np.pi
        return np.fabs(np.pi * self._radius ** 2 * np.cos(a)) + np.fabs(2 * self._radius * self._height * np.sin(a))


__all__ = ["GBM_DETECTORS", "gbm_tables", "synthetic_matrix", "NaIResponse", "BGOResponse", "GBMResponseFile",
           "SyntheticGBMResponse", "separation_angle", "separation_angles", "angle_factor", "base_response"]
