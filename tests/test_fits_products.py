"""TTE / RSP FITS products and response ingestion from OGIP files (§8 f4), CPU only.

The reference writes these files through astropy (utils/tte_file.py, utils/response_file.py, io/gbm_fits.py),
which is not installable here, so byte parity with the reference's files is UNPINNED.  What is checked
instead: (1) the files obey the FITS standard, verified by a byte-level validator written in this test and
independent of the package's own reader; (2) HDU names, column names / TFORM / TUNIT and the header keywords
are the ones the reference's classes declare; (3) values round-trip exactly; (4) the reader understands
files it did not write (a hand-packed OGIP response with grouped variable-length columns).
"""
import struct

import numpy as np
import pytest

from cosmogrb_b200.io import GRBSave, grbsave_to_gbm_fits
from cosmogrb_b200.io.grb_save import write_grb_file
from cosmogrb_b200.response import Response
from cosmogrb_b200.utils import fits_lite as fl
from cosmogrb_b200.utils.response_file import RMF, RSP, read_response
from cosmogrb_b200.utils.tte_file import TTEFile, read_tte
from tests import helpers as h
from tests.test_host_api import _storage

_WIDTH = {"L": 1, "B": 1, "I": 2, "J": 4, "K": 8, "E": 4, "D": 8, "A": 1, "P": 8, "Q": 16}


def validate_fits(path):
    """Independent structural check straight from the FITS standard (4.0 §3, §4, §7.3).  Returns
    [(header dict, list of (ttype, tform, tunit), data offset, NAXIS1, NAXIS2)] per HDU."""
    raw = open(path, "rb").read()
    assert len(raw) % 2880 == 0, "file is not a whole number of 2880-byte blocks"
    pos, out, first = 0, [], True
    while pos < len(raw):
        cards, end = [], False
        while not end:
            block = raw[pos:pos + 2880]
            assert len(block) == 2880
            assert all(32 <= b <= 126 for b in block), "header holds non-printable bytes"
            pos += 2880
            for i in range(0, 2880, 80):
                c = block[i:i + 80].decode("ascii")
                if c.startswith("END" + " " * 77):
                    end = True
                    assert block[i + 80:].strip() == b"", "non-blank cards after END"
                    break
                cards.append(c)
        hd = {}
        for c in cards:
            key = c[:8].rstrip()
            assert key == key.upper() and " " not in key
            if c[8:10] == "= ":
                v = c[10:]
                if v.lstrip().startswith("'"):
                    assert v.startswith("'"), "string value must start in column 11"
                    val = v[1:v.index("'", 1)].rstrip()
                else:
                    val = v.split("/")[0].strip()
                    if val:     # fixed format: numbers and logicals are right-justified to column 30
                        assert len(v.split("/")[0].rstrip()) == 20, c
                hd.setdefault(key, val)
        keys = [c[:8].rstrip() for c in cards]
        if first:
            assert keys[:3] == ["SIMPLE", "BITPIX", "NAXIS"] and hd["SIMPLE"] == "T"
            assert int(hd["NAXIS"]) == 0 and hd.get("EXTEND") == "T"
            out.append((hd, [], pos, 0, 0))
            first = False
            continue
        assert keys[:8] == ["XTENSION", "BITPIX", "NAXIS", "NAXIS1", "NAXIS2", "PCOUNT", "GCOUNT", "TFIELDS"]
        assert hd["XTENSION"] == "BINTABLE" and hd["BITPIX"] == "8" and hd["NAXIS"] == "2" and hd["GCOUNT"] == "1"
        n1, n2, pc, nf = int(hd["NAXIS1"]), int(hd["NAXIS2"]), int(hd["PCOUNT"]), int(hd["TFIELDS"])
        cols, width = [], 0
        for i in range(1, nf + 1):
            tform = hd[f"TFORM{i}"]
            rep = int(tform[:-1]) if tform[:-1] else 1
            width += rep * _WIDTH[tform[-1]]
            cols.append((hd[f"TTYPE{i}"], tform, hd.get(f"TUNIT{i}")))
        assert width == n1, "NAXIS1 does not equal the summed column widths"
        n = n1 * n2 + pc
        out.append((hd, cols, pos, n1, n2))
        pad = raw[pos + n:pos + n + (-n % 2880)]
        assert pad == b"\0" * len(pad), "binary-table padding must be zero bytes"
        pos += n + (-n % 2880)
    assert pos == len(raw)
    return out, raw


def _tte(tmp_path, n=50001, det="n3"):
    rs = np.random.RandomState(5)
    t = np.sort(rs.uniform(-100.0, 300.0, n)) + 5.0e8
    pha = rs.randint(0, 128, n)
    ce = h.response(det).channel_edges
    f = TTEFile(det_name=det, tstart=-100 + 5e8, tstop=300 + 5e8, trigger_time=5e8, ra=312.0, dec=-62.0,
                channel=np.arange(0, 128, dtype=np.int16), emin=ce[:-1], emax=ce[1:], pha=pha.astype(np.int16), time=t)
    fn = tmp_path / f"tte_{det}.fits"
    f.writeto(fn, overwrite=True)
    return fn, t, pha, ce


def test_tte_file_structure_and_round_trip(tmp_path):
    fn, t, pha, ce = _tte(tmp_path)
    hdus, raw = validate_fits(fn)
    assert [x[0].get("EXTNAME") for x in hdus] == [None, "EBOUNDS", "EVENTS", "GTI"]      # tte_file.py:297-308
    prim = hdus[0][0]
    assert prim["DETNAM"] == "n3" and prim["INSTRUME"] == "GBM" and prim["TELESCOP"] == "GLAST"
    assert float(prim["TRIGTIME"]) == 5e8
    assert hdus[1][1] == [("CHANNEL", "I", None), ("E_MIN", "D", "keV"), ("E_MAX", "D", "keV")]
    assert hdus[2][1] == [("TIME", "D", "s"), ("PHA", "I", None)]
    assert hdus[3][1] == [("START", "D", None), ("STOP", "D", None)]
    ev = hdus[2][0]
    for key, want in (("HDUCLASS", "OGIP"), ("HDUCLAS1", "EVENTS"), ("DETCHANS", "128"), ("TELESCOP", "GLAST"),
                      ("INSTRUME", "GBM"), ("TIMESYS", "TT"), ("TIMEUNIT", "s"), ("DETNAM", "NAI_03"),
                      ("OBJECT", "GRB080916200"), ("RADECSYS", "FK5"), ("CH2E_VER", "SPLINE 2.0")):
        assert ev[key] == want, key
    assert float(ev["TSTART"]) == -100 + 5e8 and float(ev["TSTOP"]) == 300 + 5e8
    assert float(ev["PRMT_BEG"]) == 5e8 and float(ev["FIFO_END"]) == 300 + 5e8           # tte_file.py:182-183
    assert float(ev["RA_OBJ"]) == 312.0 and float(ev["DEC_OBJ"]) == -62.0
    assert "FIFO_END" not in hdus[3][0] and "DETCHANS" not in hdus[3][0]                  # the GTI keyword set
    assert hdus[1][0]["HDUCLAS2"] == "EBOUNDS" and hdus[1][0]["CHANTYPE"] == "PHA"
    # big-endian records, decoded by hand: TIME (>f8) + PHA (>i2), 10 bytes per event
    _, _, off, n1, n2 = hdus[2]
    assert (n1, n2) == (10, len(t))
    rec = np.frombuffer(raw, dtype=np.dtype([("t", ">f8"), ("p", ">i2")]), count=n2, offset=off)
    assert np.array_equal(rec["t"], t) and np.array_equal(rec["p"], pha)
    # the package's own reader
    r = read_tte(fn)
    assert np.array_equal(r["time"], t) and np.array_equal(r["pha"], pha)
    assert np.array_equal(r["emin"], ce[:-1]) and np.array_equal(r["emax"], ce[1:])
    assert np.array_equal(r["channel"], np.arange(128)) and r["det_name"] == "NAI_03"
    assert r["gti"].tolist() == [[-100 + 5e8, 300 + 5e8]]
    with pytest.raises(OSError):
        TTEFile("n3", 0, 1, 0, 0, 0, np.arange(128), ce[:-1], ce[1:], pha[:3], t[:3]).writeto(fn)   # no overwrite
    with pytest.raises(KeyError):
        TTEFile("n12", 0, 1, 0, 0, 0, np.arange(128), ce[:-1], ce[1:], pha[:3], t[:3])
    # empty event list
    f0 = tmp_path / "empty.fits"
    TTEFile("b1", 0, 1, 0, 0, 0, np.arange(128), ce[:-1], ce[1:], pha[:0], t[:0]).writeto(f0)
    validate_fits(f0)
    assert len(read_tte(f0)["time"]) == 0 and read_tte(f0)["det_name"] == "BGO_01"


def test_header_cards():
    assert fl.format_card("TSTART", 5e8) == "TSTART  =          500000000.0".ljust(80)
    assert fl.format_card("EXTVER", 1, "") == "EXTVER  =                    1".ljust(80)
    assert fl.format_card("SIMPLE", True) == "SIMPLE  =                    T".ljust(80)
    # a short string fills the value field to column 30, the comment starts in column 32 (astropy's Card and
    # cfitsio alike; held to real GBM files card by card in tests/test_fits_real_gbm.py)
    assert fl.format_card("OBJECT", "GRB", "name") == "OBJECT  = 'GRB     '           / name".ljust(80)
    assert fl.format_card("O", "it's")[10:18] == "'it''s  "
    assert fl.format_card("X", 1e-5)[10:30].strip() == "1.0E-05"
    assert len(fl.format_card("HDUDOC", "OGIP memos CAL/GEN/92-002 & 92-002a", "Documents describing the forma")) == 80
    with pytest.raises(fl.FITSError):
        fl.format_card("TOOLONGKEY", 1)
    with pytest.raises(fl.FITSError):
        fl.format_card("X", float("nan"))
    hd = fl.Header.frombytes(fl.Header([("A", 1, "c"), ("B", "it's", None), ("C", 2.5e-7, None), ("D", None, "undef"),
                                        ("E", False, None), ("MJDREFI", 51910, None), ("MJDREFI", 7.4e-4, None)]).tobytes())
    assert (hd["A"], hd["B"], hd["C"], hd["D"], hd["E"]) == (1, "it's", 2.5e-7, None, False)
    assert hd.comment("A") == "c" and hd.keys().count("MJDREFI") == 2 and hd["MJDREFI"] == 51910


def test_rsp_file_and_response_round_trip(tmp_path):
    rsp = h.response("n0")
    fn = tmp_path / "n0.rsp"
    rsp.to_fits(fn, telescope_name="fermi", instrument_name="n0", overwrite=True)          # response.py:238-263
    hdus, raw = validate_fits(fn)
    assert [x[0].get("EXTNAME") for x in hdus] == [None, "EBOUNDS", "SPECRESP MATRIX"]
    assert hdus[1][1] == [("CHANNEL", "K", None), ("E_MIN", "E", "keV"), ("E_MAX", "E", "keV")]
    assert hdus[2][1] == [("ENERG_LO", "E", "keV"), ("ENERG_HI", "E", "keV"), ("N_GRP", "I", None),
                          ("F_CHAN", "I", None), ("N_CHAN", "I", None), ("MATRIX", "128D", None)]
    m = hdus[2][0]
    assert (m["HDUCLAS1"], m["HDUCLAS2"], m["HDUCLAS3"]) == ("RESPONSE", "RSP_MATRIX", "FULL")
    assert (m["TELESCOP"], m["INSTRUME"], m["DETCHANS"], m["TLMIN4"]) == ("fermi", "n0", "128", "1")
    assert hdus[1][0]["TLMIN1"] == "1" and hdus[1][0]["CHANTYPE"] == "PI"
    r = read_response(fn)
    assert np.array_equal(r["matrix"], rsp.matrix)                          # the matrix is stored in fp64
    # edges go through float32, as in the reference (response_file.py:189-191)
    assert np.array_equal(r["energy_edges"], rsp.energy_edges.astype(np.float32).astype("f8"))
    assert np.array_equal(r["channel_edges"], rsp.channel_edges.astype(np.float32).astype("f8"))
    r2 = Response.from_fits(fn, geometric_area=rsp.geometric_area)
    assert np.array_equal(r2.matrix, rsp.matrix) and r2.geometric_area == rsp.geometric_area
    np.testing.assert_allclose(r2.cumulative_matrix, rsp.cumulative_matrix, rtol=0, atol=0)
    np.testing.assert_allclose(r2.effective_area_packed[1], rsp.effective_area_packed[1], rtol=0, atol=0)
    np.testing.assert_allclose(r2.energy_edges, rsp.energy_edges, rtol=1e-7)
    with pytest.raises(OSError):
        rsp.to_fits(fn)
    RMF(rsp.energy_edges, rsp.channel_edges, rsp.matrix.T, "t", "i").writeto(tmp_path / "n0.rmf")
    hd2, _ = validate_fits(tmp_path / "n0.rmf")
    assert hd2[2][0]["EXTNAME"] == "MATRIX" and hd2[2][0]["HDUCLAS3"] == "REDIST"
    with pytest.raises(AssertionError):
        RSP(rsp.energy_edges, rsp.channel_edges, rsp.matrix, "t", "i")      # wants [n_channels, n_photon_bins]


def _card(k, v, string=False):
    body = f"{k:<8s}= " + (f"'{v:<8s}'" if string else f"{v:>20}")
    return body.ljust(80).encode()


def test_read_grouped_variable_length_ogip_response(tmp_path):
    """A response file this package did not write: hand-packed bytes in the form real GBM .rsp files take --
    N_GRP groups per photon-energy row, F_CHAN / N_CHAN / MATRIX as variable-length arrays ('PI', 'PE') in
    the heap, channels numbered from 1 -- must reconstruct the dense matrix."""
    n_e, n_c = 14, 10          # Response needs > 10 photon bins (response.py:110, ea_curve[:-10])
    rs = np.random.RandomState(2)
    dense = np.zeros((n_e, n_c), dtype=np.float32)
    groups = []
    for r in range(n_e):
        if r == 0:
            groups.append([])                                   # a row with no response at all
            continue
        g = [(1 + r % 3, 2)] if r % 2 else [(0, 2), (5, 3)]     # (first channel 0-based, length)
        for c0, n in g:
            dense[r, c0:c0 + n] = rs.uniform(0.5, 20.0, n).astype(np.float32)
        groups.append(g)
    elo = np.logspace(1, 3, n_e + 1).astype(np.float32)
    heap, main = b"", b""
    for r in range(n_e):
        g = groups[r]
        fch = np.array([c0 + 1 for c0, _ in g], dtype=">i2").tobytes()
        nch = np.array([n for _, n in g], dtype=">i2").tobytes()
        mat = b"".join(dense[r, c0:c0 + n].astype(">f4").tobytes() for c0, n in g)
        o_f, o_n, o_m = len(heap), len(heap) + len(fch), len(heap) + len(fch) + len(nch)
        heap += fch + nch + mat
        main += struct.pack(">ffh", elo[r], elo[r + 1], len(g))
        main += struct.pack(">ii", len(g), o_f) + struct.pack(">ii", len(g), o_n)
        main += struct.pack(">ii", sum(n for _, n in g), o_m)
    row = 4 + 4 + 2 + 8 + 8 + 8
    assert len(main) == row * n_e

    def hdu(cards, data):
        hb = b"".join(cards) + b"END".ljust(80)
        hb += b" " * (-len(hb) % 2880)
        return hb + data + b"\0" * (-len(data) % 2880)

    prim = hdu([_card("SIMPLE", "T"), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", "T")], b"")
    ce = np.linspace(4.0, 1000.0, n_c + 1).astype(np.float32)
    eb = b"".join(struct.pack(">hff", i + 1, ce[i], ce[i + 1]) for i in range(n_c))
    ebounds = hdu([_card("XTENSION", "BINTABLE", True), _card("BITPIX", 8), _card("NAXIS", 2), _card("NAXIS1", 10),
                   _card("NAXIS2", n_c), _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", 3),
                   _card("TTYPE1", "CHANNEL", True), _card("TFORM1", "I", True), _card("TTYPE2", "E_MIN", True),
                   _card("TFORM2", "E", True), _card("TTYPE3", "E_MAX", True), _card("TFORM3", "E", True),
                   _card("EXTNAME", "EBOUNDS", True)], eb)
    matrix = hdu([_card("XTENSION", "BINTABLE", True), _card("BITPIX", 8), _card("NAXIS", 2), _card("NAXIS1", row),
                  _card("NAXIS2", n_e), _card("PCOUNT", len(heap)), _card("GCOUNT", 1), _card("TFIELDS", 6),
                  _card("TTYPE1", "ENERG_LO", True), _card("TFORM1", "E", True),
                  _card("TTYPE2", "ENERG_HI", True), _card("TFORM2", "E", True),
                  _card("TTYPE3", "N_GRP", True), _card("TFORM3", "I", True),
                  _card("TTYPE4", "F_CHAN", True), _card("TFORM4", "PI(2)", True),
                  _card("TTYPE5", "N_CHAN", True), _card("TFORM5", "PI(2)", True),
                  _card("TTYPE6", "MATRIX", True), _card("TFORM6", "PE(10)", True),
                  _card("EXTNAME", "SPECRESP MATRIX", True), _card("DETCHANS", n_c), _card("TLMIN4", 1),
                  _card("TELESCOP", "GLAST", True), _card("INSTRUME", "GBM", True)], main + heap)
    fn = tmp_path / "grouped.rsp"
    fn.write_bytes(prim + ebounds + matrix)
    r = read_response(fn)
    assert r["matrix"].shape == (n_e, n_c) and np.array_equal(r["matrix"], dense.astype("f8"))
    assert np.array_equal(r["energy_edges"], elo.astype("f8")) and np.array_equal(r["channel_edges"], ce.astype("f8"))
    assert (r["telescope"], r["instrument"]) == ("GLAST", "GBM")
    rsp = Response.from_fits(fn, geometric_area=100.0)
    assert rsp.matrix.shape == (n_e, n_c) and rsp.cumulative_matrix[1, -1] == pytest.approx(1.0)
    assert np.all(rsp.cumulative_matrix[0] == 0.0)               # the empty row stays a zero-probability row


def test_grbsave_to_gbm_fits(tmp_path):
    pairs = [_storage(d, seed=i) for i, d in enumerate(("n0", "b0"))]
    fn = tmp_path / "SynthGRB_7_store.h5"
    write_grb_file(fn, name="SynthGRB_7", T0=2.5, z=1.0, duration=2.0, ra=312.0, dec=-62.0,
                   source_params=dict(peak_flux=1e-5), extra_info=None, storages=[p[0] for p in pairs],
                   responses=[p[1] for p in pairs], backend="npz")
    files = grbsave_to_gbm_fits(fn, destination=str(tmp_path))
    assert sorted(files) == ["b0", "n0"]
    for (lc, rsp) in pairs:
        f = files[lc.name]
        assert f["tte"].endswith(f"tte_SynthGRB_7_{lc.name}.fits") and f["rsp"].endswith(f"rsp_SynthGRB_7_{lc.name}.rsp")
        validate_fits(f["tte"])
        validate_fits(f["rsp"])
        r = read_tte(f["tte"])
        # gbm_fits.py:36-56: every time is shifted by the light curve's time_adjustment
        assert np.array_equal(r["time"], lc.times + lc.time_adjustment) and np.array_equal(r["pha"], lc.pha)
        assert r["tstart"] == lc.tstart + lc.time_adjustment and r["tstop"] == lc.tstop + lc.time_adjustment
        assert r["trigger_time"] == 2.5 + lc.time_adjustment and (r["ra"], r["dec"]) == (312.0, -62.0)
        assert np.array_equal(read_response(f["rsp"])["matrix"], rsp.matrix)
    only = grbsave_to_gbm_fits(GRBSave.from_file(fn), destination=str(tmp_path), detectors=["b0"])
    assert list(only) == ["b0"]
    with pytest.raises(AssertionError):
        grbsave_to_gbm_fits(fn, destination=str(tmp_path), detectors=["n5"])


def test_gbm_response_from_file(tmp_path):
    """GBMResponseFile: the detector response of a GBM burst read from an OGIP file -- same
    geometry as the synthetic responses, the file's matrix, the device table the engine consumes."""
    from cosmogrb_b200.instruments.gbm.synthetic_response import BGOResponse, GBMResponseFile, NaIResponse

    for det, cls in (("n4", NaIResponse), ("b0", BGOResponse)):
        syn = cls(det, 312.0, -62.0)
        fn = tmp_path / f"{det}.rsp"
        syn.to_fits(fn, telescope_name="fermi", instrument_name=det)
        rsp = GBMResponseFile(det, 312.0, -62.0, fn, trigger_time=3.5)
        assert np.array_equal(rsp.matrix, syn.matrix) and rsp.geometric_area == syn.geometric_area
        assert rsp.separation_angle == syn.separation_angle and rsp.detector_name == det
        assert (rsp.ra, rsp.dec, rsp.T0, rsp.trigger_time) == (312.0, -62.0, 3.5, 3.5)
        assert np.array_equal(rsp.cumulative_matrix, syn.cumulative_matrix)
        # photon edges went through float32 in the file; everything built from the matrix alone is identical
        a, b = rsp.device_table(h.bkg_weights(det), "linear"), syn.device_table(h.bkg_weights(det), "linear")
        assert a.shape == b.shape
        from cosmogrb_b200 import _capi
        assert np.array_equal(a[_capi.DT_CUM:_capi.DT_GUIDE], b[_capi.DT_CUM:_capi.DT_GUIDE])
        np.testing.assert_allclose(a[_capi.DT_EDGES:_capi.DT_EDGES + 141], b[_capi.DT_EDGES:_capi.DT_EDGES + 141], rtol=1e-7)
    small = Response(matrix=np.ones((20, 16)), geometric_area=1.0, energy_edges=np.logspace(1, 3, 21),
                     channel_edges=np.logspace(1, 3, 17))
    small.to_fits(tmp_path / "small.rsp")
    with pytest.raises(ValueError, match="140 x 128"):
        GBMResponseFile("n0", 0.0, 0.0, tmp_path / "small.rsp")
    with pytest.raises(AssertionError):
        GBMResponseFile("x9", 0.0, 0.0, fn)


def test_response_convolve():
    """Response.set_function / convolve (response.py:202-236): bin integrals folded through the matrix."""
    rsp = h.response("n0")
    rsp.set_function(lambda e1, e2, t1, t2: (t2 - t1) * (np.log(e2 / e1) if e1 > 6.0 else np.inf))   # E^-1, one bad bin
    got = rsp.convolve(0.0, 2.0)
    e = rsp.energy_edges
    flux = 2.0 * np.log(e[1:] / e[:-1])
    flux[e[:-1] <= 6.0] = 0.0
    assert got.shape == (128,) and np.allclose(got, flux @ rsp.matrix, rtol=1e-13)
