"""OGIP response FITS files: ``RSP`` / ``RMF`` (cosmogrb/utils/response_file.py:8-212, "lifted from 3ML").

Layout written (as the reference's astropy path lays it out): EBOUNDS (CHANNEL K 1-based, E_MIN E, E_MAX E
[keV]) and ``SPECRESP MATRIX`` (``MATRIX`` for an RMF) with ENERG_LO E, ENERG_HI E [keV], N_GRP I = 1,
F_CHAN I = 1, N_CHAN I = n_channels and MATRIX as a fixed-length vector of n_channels D per photon-energy
row.  Energies are cast to float32 as the reference does (response_file.py:189-191).

``read_response`` reads such a file back -- and, beyond what the reference offers, any OGIP response with
grouped, variable-length MATRIX / F_CHAN / N_CHAN columns (the form real GBM ``.rsp`` files take) -- into the
``(matrix[Nphoton, Nchannel], photon edges, channel edges)`` that ``Response`` consumes.
"""
import numpy as np

from . import fits_lite as fl

_OGIP = [
    ("HDUCLASS", "OGIP    ", "format conforms to OGIP standard"),
    ("HDUVERS", "1.1.0   ", "Version of format (OGIP memo CAL/GEN/92-002a)"),
    ("HDUDOC", "OGIP memos CAL/GEN/92-002 & 92-002a", "Documents describing the forma"),
    ("HDUVERS1", "1.0.0   ", "Obsolete - included for backwards compatibility"),
    ("HDUVERS2", "1.1.0   ", "Obsolete - included for backwards compatibility"),
]
_CREATOR = ("CREATOR", "COSMOGRB_B200", "cosmogrb_b200 photon-event engine")


def _ebounds_hdu(ebounds):
    n = len(ebounds) - 1
    header = fl.Header([("EXTNAME", "EBOUNDS", "Extension name")] + _OGIP + [
        ("CHANTYPE", "PI", "Channel type"),
        ("CONTENT", "OGIPResponse Matrix", "File content"),
        ("HDUCLAS1", "RESPONSE", "Extension contains response data  "),
        ("HDUCLAS2", "EBOUNDS ", "Extension contains EBOUNDS"),
        ("TLMIN1", 1, "Minimum legal channel number"),
        _CREATOR,
    ])
    return fl.BinTableHDU([fl.Column("CHANNEL", np.arange(1, n + 1, dtype=np.int64)),
                           fl.Column("E_MIN", ebounds[:-1], unit="keV"),
                           fl.Column("E_MAX", ebounds[1:], unit="keV")], header=header)


def _matrix_hdu(mc_energies, ebounds, matrix, extname, hduclas3, hduclas3_comment, telescope_name, instrument_name):
    n_mc, n_ch = len(mc_energies) - 1, len(ebounds) - 1
    matrix = np.asarray(matrix, dtype="f8")
    assert matrix.shape == (n_ch, n_mc), "Matrix has the wrong shape. Should be %i x %i, got %i x %i" % (
        n_ch, n_mc, matrix.shape[0], matrix.shape[1])
    ones = np.ones(n_mc, np.int16)
    header = fl.Header([("EXTNAME", extname, "Extension name")] + _OGIP + [
        ("HDUCLAS1", "RESPONSE", "dataset relates to spectral response"),
        ("HDUCLAS2", "RSP_MATRIX", "dataset is a spectral response matrix"),
        ("HDUCLAS3", hduclas3, hduclas3_comment),
        ("CHANTYPE", "PI ", "Detector Channel Type in use (PHA or PI)"),
        ("DETCHANS", n_ch, "Number of channels"),
        ("FILTER", "", "Filter used"),
        ("TLMIN4", 1, "Minimum legal channel number"),
        _CREATOR,
        ("TELESCOP", telescope_name, None),
        ("INSTRUME", instrument_name, None),
    ])
    return fl.BinTableHDU([
        fl.Column("ENERG_LO", mc_energies[:-1], unit="keV"), fl.Column("ENERG_HI", mc_energies[1:], unit="keV"),
        fl.Column("N_GRP", ones), fl.Column("F_CHAN", ones), fl.Column("N_CHAN", ones * np.int16(n_ch)),
        fl.Column("MATRIX", np.ascontiguousarray(matrix.T))], header=header)


class _ResponseFile(object):
    _extname, _clas3, _clas3_comment = None, None, None

    def __init__(self, mc_energies, ebounds, matrix, telescope_name, instrument_name):
        mc_energies = np.array(mc_energies, np.float32)
        ebounds = np.array(ebounds, np.float32)
        self._hdus = [fl.PrimaryHDU(), _ebounds_hdu(ebounds),
                      _matrix_hdu(mc_energies, ebounds, matrix, self._extname, self._clas3, self._clas3_comment,
                                  telescope_name, instrument_name)]

    @property
    def hdus(self):
        return self._hdus

    def writeto(self, file_name, overwrite=False, clobber=None):
        fl.write_fits(file_name, self._hdus, overwrite=bool(overwrite if clobber is None else clobber))


class RMF(_ResponseFile):
    """energy dispersion only (response_file.py:156-182); ``matrix`` is [n_channels, n_photon_bins]"""
    _extname, _clas3, _clas3_comment = "MATRIX", "REDIST", "dataset represents energy dispersion only"


class RSP(_ResponseFile):
    """dispersion x effective area in one matrix (response_file.py:185-212)"""
    _extname, _clas3, _clas3_comment = "SPECRESP MATRIX", "FULL", "dataset represents energy dispersion only"


def read_response(file_name, extension=None):
    """-> dict(matrix [n_photon_bins, n_channels] f8 (cm^2), energy_edges [n_photon_bins + 1],
    channel_edges [n_channels + 1], telescope, instrument).  Handles the OGIP channel grouping
    (N_GRP / F_CHAN / N_CHAN, fixed or variable length, TLMIN-based first channel).  ``extension``: index
    among the matrix extensions of a multi-matrix (``.rsp2``) file, default the first."""
    hdus = fl.read_fits(file_name)
    mats = [h for h in hdus[1:] if str(h.name).strip().upper() in ("SPECRESP MATRIX", "MATRIX")]
    if not mats:
        raise fl.FITSError(f"{file_name}: no MATRIX / SPECRESP MATRIX extension")
    mh = mats[extension or 0]
    eb = fl.get_hdu(hdus, "EBOUNDS")
    d, h = mh.data, mh.header
    n_ch = int(h.get("DETCHANS", len(eb.data)))
    names = list(d.dtype.names)
    f_col = names.index("F_CHAN") + 1
    first = int(h.get(f"TLMIN{f_col}", 1))          # OGIP default: channels start at 1 unless TLMIN says 0
    n_e = len(d)
    matrix = np.zeros((n_e, n_ch))
    for r in range(n_e):
        n_grp = int(d["N_GRP"][r])
        f_chan = np.atleast_1d(d["F_CHAN"][r]).astype(np.int64)
        n_chan = np.atleast_1d(d["N_CHAN"][r]).astype(np.int64)
        row = np.asarray(d["MATRIX"][r], dtype="f8").ravel()
        k = 0
        for g in range(n_grp):
            c0, n = int(f_chan[g]) - first, int(n_chan[g])
            matrix[r, c0:c0 + n] = row[k:k + n]
            k += n
    energy_edges = np.append(np.asarray(d["ENERG_LO"], dtype="f8"), float(d["ENERG_HI"][-1]))
    channel_edges = np.append(np.asarray(eb.data["E_MIN"], dtype="f8"), float(eb.data["E_MAX"][-1]))
    return dict(matrix=matrix, energy_edges=energy_edges, channel_edges=channel_edges,
                telescope=h.get("TELESCOP"), instrument=h.get("INSTRUME"), extname=str(mh.name).strip())


__all__ = ["RSP", "RMF", "read_response"]
