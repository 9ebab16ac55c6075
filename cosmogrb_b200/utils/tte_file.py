"""GBM time-tagged-event (TTE) FITS product: ``TTEFile`` (cosmogrb/utils/tte_file.py:25-309).

Same four HDUs as the reference writes through astropy -- PRIMARY (no data), EBOUNDS (CHANNEL I, E_MIN D,
E_MAX D [keV]), EVENTS (TIME D [s], PHA I) and GTI (START D, STOP D) -- with the reference's OGIP keyword
sets, laid out by ``utils.fits_lite`` (astropy is optional).  The three extensions share one keyword table
here instead of three literal lists; what differs per extension is listed in ``_EXTENSIONS``.

Known quirk kept: the reference's keyword lists name ``MJDREFI`` twice (tte_file.py:63-64, :143-144, :226-227),
the second carrying the fractional part.  Both cards are written, in that order, as a header built from that
list holds them; the fraction is ALSO written under its proper OGIP name ``MJDREFF`` so that standard readers
find the epoch.
"""
import numpy as np

from . import fits_lite as fl

_det_translate = {f"n{c}": f"NAI_{i:02d}" for i, c in enumerate("0123456789ab")}
_det_translate.update(b0="BGO_00", b1="BGO_01")

_FIXED_DATE = "2009-05-19T18:49:32"         # the reference stamps every file with this date (tte_file.py:54-56)
_CREATOR = "COSMOGRB_B200"


def _ogip_header(extname, hduclas1, hduclas1_comment, extra_class=(), detchans=True, extra_times=()):
    cards = [
        ("EXTNAME", extname, "Extension name"),
        ("CONTENT", "OGIP PHA data", "File content"),
        ("HDUCLASS", "OGIP    ", "format conforms to OGIP standard"),
        ("HDUVERS", "1.1.0   ", "Version of format (OGIP memo CAL/GEN/92-002a)"),
        ("HDUDOC", "OGIP memos CAL/GEN/92-002 & 92-002a", "Documents describing the forma"),
        ("HDUVERS1", "1.0.0   ", "Obsolete - included for backwards compatibility"),
        ("HDUVERS2", "1.1.0   ", "Obsolete - included for backwards compatibility"),
        ("HDUCLAS1", hduclas1, hduclas1_comment),
    ]
    cards += list(extra_class)
    cards.append(("FILTER", "", "Filter used"))
    if extname == "EBOUNDS":
        cards.append(("CHANTYPE", "PHA", "Channel type"))
    if detchans:
        cards.append(("DETCHANS", 128, "Number of channels"))
    cards += [
        ("TELESCOP", "GLAST", "Name of mission/satelite"),
        ("INSTRUME", "GBM", "Specific instrument used for observation"),
        ("DATE", _FIXED_DATE, "file creation date"),
        ("DATE-OBS", _FIXED_DATE, "Date of start of observation"),
        ("DATE-END", _FIXED_DATE, "Date of end of observation"),
        ("OBSERVER", "God", "The creator of the Universe"),
        ("OBJECT", "GRB080916200", "Burst name in standard format"),
        ("ORIGIN", "MPE", "Name of organization creating the file"),
        ("TIMESYS", "TT", "Time system used in keywords"),
        ("TIMEUNIT", "s", "Time since MJDREF, used in TSTART and TSTOP"),
        ("MJDREFI", 51910, "MJD of GLAST reference epoch"),
        ("MJDREFI", 7.428703703703703e-4, "MJD of GLAST reference epoch"),
        ("MJDREFF", 7.428703703703703e-4, "MJD of GLAST reference epoch, fractional part"),
        ("RADECSYS", "FK5     ", "Stellar reference frame"),
        ("EQUINOX", 2000.0, "Equinox of RA and Dec"),
        ("RA_OBJ", None, "Calculated RA of burst"),
        ("DEC_OBJ", None, "Calculated Dec of burst"),
        ("ERR_RAD", 3.000, "Calculated Location of Error Radius"),
        ("EXTVER", 1, ""),
        ("CH2E_VER", "SPLINE 2.0", ""),
        ("GAIN_COR", 1.0, ""),
        ("TSTART", None, "[GLAST MET] Observation start time"),
        ("TSTOP", None, "[GLAST MET] Observation stop time"),
        ("TRIGTIME", None, "Trigger time realtive to MJDREF, double precision"),
    ]
    cards += list(extra_times)
    cards.append(("DETNAM", None, "Individual detector name"))
    return fl.Header(cards)


def _fill(header, det_name, tstart, tstop, trigger_time, ra, dec):
    def last(key, value):
        # .set() on a duplicated keyword must not touch MJDREFI; these keys are unique
        header.set(key, value)

    last("RA_OBJ", float(ra))
    last("DEC_OBJ", float(dec))
    last("TSTART", float(tstart))
    last("TSTOP", float(tstop))
    last("TRIGTIME", float(trigger_time))
    last("DETNAM", _det_translate[det_name])
    header.set("CREATOR", _CREATOR, "cosmogrb_b200 photon-event engine")
    return header


class TTEFile(object):
    """``TTEFile(det_name, tstart, tstop, trigger_time, ra, dec, channel, emin, emax, pha, time)`` -- same
    arguments as the reference (tte_file.py:264-277); ``writeto(path, overwrite=False)``."""

    def __init__(self, det_name, tstart, tstop, trigger_time, ra, dec, channel, emin, emax, pha, time):
        if det_name not in _det_translate:
            raise KeyError(det_name)
        time = np.asarray(time, dtype="f8")
        pha = np.asarray(pha)
        if time.shape != pha.shape or time.ndim != 1:
            raise ValueError("time and pha must be 1-d arrays of the same length")
        primary = fl.PrimaryHDU(fl.Header([
            ("TSTART", float(tstart), None), ("TSTOP", float(tstop), None), ("TRIGTIME", float(trigger_time), None),
            ("DATE-OBS", _FIXED_DATE, None), ("DATE-END", _FIXED_DATE, None), ("DETNAM", det_name, None),
            ("INSTRUME", "GBM", None), ("TELESCOP", "GLAST", None)]))
        args = (det_name, tstart, tstop, trigger_time, ra, dec)
        ebounds = fl.BinTableHDU(
            [fl.Column("CHANNEL", np.asarray(channel, dtype=np.int16)),
             fl.Column("E_MIN", np.asarray(emin, dtype="f8"), unit="keV"),
             fl.Column("E_MAX", np.asarray(emax, dtype="f8"), unit="keV")],
            header=_fill(_ogip_header("EBOUNDS", "RESPONSE", "These are typically found in RMF files  ",
                                      extra_class=[("HDUCLAS2", "EBOUNDS ", "Fram CAL/GEN/92-002")]), *args))
        events_header = _fill(_ogip_header(
            "EVENTS", "EVENTS", "Extension constains events  ",
            extra_times=[("FIFO_END", None, "Time of the last event in FIFO"),
                         ("PRMT_BEG", None, "Time of the first event in prompt")]), *args)
        events_header.set("PRMT_BEG", float(trigger_time))          # tte_file.py:182-183
        events_header.set("FIFO_END", float(tstop))
        events = fl.BinTableHDU([fl.Column("TIME", time, unit="s"), fl.Column("PHA", pha.astype(np.int16))],
                                header=events_header)
        gti = fl.BinTableHDU(
            [fl.Column("START", np.array([tstart], dtype="f8")), fl.Column("STOP", np.array([tstop], dtype="f8"))],
            header=_fill(_ogip_header("GTI", "GTI", "Extension constains events  ", detchans=False), *args))
        self._hdus = [primary, ebounds, events, gti]

    @property
    def hdus(self):
        return self._hdus

    def writeto(self, file_name, overwrite=False, clobber=None):
        fl.write_fits(file_name, self._hdus, overwrite=bool(overwrite if clobber is None else clobber))


def read_tte(file_name):
    """Read a TTE file back: dict with ``time``, ``pha``, ``channel``, ``emin``, ``emax``, ``gti`` and the
    EVENTS header (``tstart``, ``tstop``, ``trigger_time``, ``det_name``, ``ra``, ``dec``)."""
    hdus = fl.read_fits(file_name)
    ev, eb, gti = fl.get_hdu(hdus, "EVENTS"), fl.get_hdu(hdus, "EBOUNDS"), fl.get_hdu(hdus, "GTI")
    h = ev.header
    return dict(time=ev.data["TIME"], pha=ev.data["PHA"], channel=eb.data["CHANNEL"], emin=eb.data["E_MIN"],
                emax=eb.data["E_MAX"], gti=np.column_stack((gti.data["START"], gti.data["STOP"])),
                tstart=h["TSTART"], tstop=h["TSTOP"], trigger_time=h["TRIGTIME"], det_name=h["DETNAM"],
                ra=h["RA_OBJ"], dec=h["DEC_OBJ"])


__all__ = ["TTEFile", "read_tte"]
