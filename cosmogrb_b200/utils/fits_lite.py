"""A small FITS reader / writer in NumPy (FITS standard 4.0: primary HDU without data + BINTABLE extensions).

The reference writes its TTE and RSP products through ``astropy.io.fits`` wrapped by
``responsum.utils.fits_file`` (cosmogrb/utils/fits_file.py:1-19, utils/tte_file.py:25-309,
utils/response_file.py:8-212).  astropy is an optional dependency here, and the event tables of a bright
burst hold 10^7-10^8 rows, so the tables are laid out directly as big-endian NumPy records (one
``tofile`` per HDU) instead of going through a column-object layer.

Supported: header cards (logical / integer / float / string values with comments, COMMENT / HISTORY),
binary tables with scalar and fixed-length vector columns of type B I J K E D A; the reader additionally
resolves variable-length array columns (``P`` / ``Q`` descriptors into the heap), which real OGIP response
files use for their MATRIX / F_CHAN / N_CHAN columns.
"""
from __future__ import annotations

import os
import re

import numpy as np

BLOCK = 2880
CARD = 80

# TFORM letter -> big-endian numpy type
_TFORM = {"L": "S1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8", "A": "S"}
# numpy kind/size -> TFORM letter (the reference's mapping, responsum.utils.fits_file: int16 I, int32 J,
# int64 K, float32 E, float64 D; unsigned types are stored in the signed type of the same width)
_CODE = {("i", 2): "I", ("i", 4): "J", ("i", 8): "K", ("u", 1): "B", ("u", 2): "I", ("u", 4): "J",
         ("f", 4): "E", ("f", 8): "D"}


class FITSError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------
# header
# ------------------------------------------------------------------------------------------------
def _format_value(value):
    if isinstance(value, (bool, np.bool_)):
        return f"{'T' if value else 'F':>20s}"
    if isinstance(value, (int, np.integer)):
        return f"{int(value):>20d}"
    if isinstance(value, (float, np.floating)):
        v = float(value)
        if not np.isfinite(v):
            raise FITSError("FITS headers cannot hold NaN / infinity")
        s = repr(v).upper()                     # shortest round-trip form; 'E' exponent as the standard asks
        if "." not in s and "E" not in s:
            s += ".0"
        elif "E" in s and "." not in s.split("E")[0]:
            m, e = s.split("E")
            s = f"{m}.0E{e}"
        if len(s) > 20:
            # keep the fixed-format value field (columns 11-30): drop trailing mantissa digits, as astropy does
            for p in range(16, 0, -1):
                s = f"{v:.{p}E}"
                if len(s) <= 20:
                    break
        return f"{s:>20s}"
    if value is None:
        return " " * 20                         # undefined value (standard 4.2.1: blank value field)
    s = str(value).replace("'", "''").rstrip()
    # closing quote not before column 20, value field filled to column 30 so that a comment starts in column 32:
    # what astropy's Card and cfitsio both write (checked card by card against real GBM files, tests/test_fits_real_gbm.py)
    return f"{chr(39)}{s:<8s}{chr(39)}".ljust(20)


def format_card(key, value=None, comment=None):
    """One 80-character card image."""
    key = str(key).upper()
    if key in ("COMMENT", "HISTORY", ""):
        return f"{key:<8s}{'' if value is None else str(value)}"[:CARD].ljust(CARD)
    if len(key) > 8 or not re.fullmatch(r"[A-Z0-9_-]*", key):
        raise FITSError(f"illegal FITS keyword {key!r}")
    body = f"{key:<8s}= {_format_value(value)}"
    if len(body) > CARD:
        raise FITSError(f"value of {key} does not fit one card")
    if comment:
        body = f"{body} / {comment}"
    return body[:CARD].ljust(CARD)


def _parse_value(field):
    """value field of a card (columns 11-80) -> (value, comment)"""
    s = field.strip()
    if s.startswith("'"):
        # string: up to the closing single quote (doubled quotes are literal)
        i, out = 1, []
        while i < len(s):
            if s[i] == "'":
                if i + 1 < len(s) and s[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(s[i])
            i += 1
        rest = s[i + 1:]
        comment = rest.split("/", 1)[1].strip() if "/" in rest else None
        return "".join(out).rstrip(), comment
    val, _, comment = s.partition("/")
    val, comment = val.strip(), (comment.strip() or None)
    if val == "":
        return None, comment
    if val == "T":
        return True, comment
    if val == "F":
        return False, comment
    try:
        return int(val), comment
    except ValueError:
        pass
    try:
        return float(val.replace("D", "E")), comment
    except ValueError:
        return val, comment


class Header(object):
    """An ordered list of (keyword, value, comment) cards.  Like a header built from a list of tuples in
    astropy, appended duplicates are kept (the reference's TTE keyword lists carry ``MJDREFI`` twice,
    utils/tte_file.py:63-64); ``set`` updates the first card of that keyword or appends."""

    def __init__(self, cards=()):
        self.cards = []
        for c in cards:
            self.append(*c)

    def append(self, key, value=None, comment=None):
        self.cards.append((str(key).upper(), value, comment))

    def set(self, key, value=None, comment=None):
        key = str(key).upper()
        for i, (k, _, c) in enumerate(self.cards):
            if k == key:
                self.cards[i] = (k, value, comment if comment is not None else c)
                return
        self.append(key, value, comment)

    def index(self, key):
        key = str(key).upper()
        for i, (k, _, _) in enumerate(self.cards):
            if k == key:
                return i
        raise KeyError(key)

    def __contains__(self, key):
        return any(k == str(key).upper() for k, _, _ in self.cards)

    def __getitem__(self, key):
        return self.cards[self.index(key)][1]

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def comment(self, key):
        return self.cards[self.index(key)][2]

    def keys(self):
        return [k for k, _, _ in self.cards]

    def __len__(self):
        return len(self.cards)

    def __iter__(self):
        return iter(self.cards)

    def tobytes(self):
        text = "".join(format_card(*c) for c in self.cards) + "END".ljust(CARD)
        text += " " * (-len(text) % BLOCK)
        return text.encode("ascii")

    @classmethod
    def frombytes(cls, raw):
        h = cls()
        for i in range(0, len(raw), CARD):
            card = raw[i:i + CARD].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                break
            if key in ("COMMENT", "HISTORY", ""):
                if card.strip():
                    h.append(key, card[8:].rstrip())
                continue
            if card[8:10] != "= ":
                h.append(key, card[8:].strip() or None)
                continue
            value, comment = _parse_value(card[10:])
            h.append(key, value, comment)
        return h


# ------------------------------------------------------------------------------------------------
# HDUs
# ------------------------------------------------------------------------------------------------
class Column(object):
    """One binary-table column: ``array`` is [n_rows] (scalar cells) or [n_rows, width] (vector cells)."""

    def __init__(self, name, array, unit=None, format=None):
        a = np.asarray(array)
        if a.dtype.kind in "US":
            a = np.char.encode(a, "ascii") if a.dtype.kind == "U" else a
            width = max(a.dtype.itemsize, 1)
            self.code, self.repeat, self.dtype = "A", width, np.dtype(f"S{width}")
            self.array = a.astype(self.dtype)
        else:
            if a.ndim not in (1, 2):
                raise FITSError(f"column {name}: cells must be scalars or fixed-length vectors")
            try:
                code = _CODE[(a.dtype.kind, a.dtype.itemsize)]
            except KeyError:
                raise FITSError(f"column {name}: dtype {a.dtype} has no FITS binary-table type") from None
            if format is not None:
                m = re.fullmatch(r"(\d*)([BIJKED])", format)
                if not m:
                    raise FITSError(f"column {name}: unsupported TFORM {format!r}")
                code = m.group(2)
            self.code = code
            self.repeat = 1 if a.ndim == 1 else a.shape[1]
            base = np.dtype(_TFORM[code])
            self.dtype = base if a.ndim == 1 else np.dtype((base, (self.repeat,)))
            self.array = a
        self.name, self.unit = str(name), unit

    @property
    def tform(self):
        return f"{self.repeat if (self.repeat != 1 or self.code == 'A') else ''}{self.code}"


class PrimaryHDU(object):
    """Primary HDU without a data array (what ``fits.PrimaryHDU()`` writes)."""
    name = "PRIMARY"
    data = None

    def __init__(self, header=None):
        self.header = header if isinstance(header, Header) else Header(header or ())

    def _full_header(self):
        h = Header([("SIMPLE", True, "conforms to FITS standard"), ("BITPIX", 8, "array data type"),
                    ("NAXIS", 0, "number of array dimensions"), ("EXTEND", True, None)])
        for k, v, c in self.header:
            if k not in ("SIMPLE", "BITPIX", "NAXIS", "EXTEND"):
                h.append(k, v, c)
        return h

    def write(self, f):
        f.write(self._full_header().tobytes())


class BinTableHDU(object):
    """BINTABLE extension from a list of ``Column`` (writer) or a structured array (reader)."""

    def __init__(self, columns=None, header=None, name=None, data=None, units=None):
        self.header = header if isinstance(header, Header) else Header(header or ())
        self.columns = list(columns or [])
        self._data = data
        self._units = units or {}
        if name is not None:
            self.header.set("EXTNAME", name, "Extension name")
        n = {len(c.array) for c in self.columns}
        if len(n) > 1:
            raise FITSError(f"columns of different lengths: {sorted(n)}")

    @property
    def name(self):
        return self.header.get("EXTNAME", "")

    @property
    def n_rows(self):
        if self.columns:
            return len(self.columns[0].array)
        return 0 if self._data is None else len(self._data)

    @property
    def data(self):
        """structured array in native byte order, one field per column"""
        if self._data is None:
            self._data = self._records().astype(self._records().dtype.newbyteorder("="))
        return self._data

    def unit(self, column):
        for c in self.columns:
            if c.name == column:
                return c.unit
        return self._units.get(column)

    def _records(self):
        dt = np.dtype({"names": [c.name for c in self.columns], "formats": [c.dtype for c in self.columns]})
        rec = np.empty(self.n_rows, dtype=dt)
        for c in self.columns:
            rec[c.name] = c.array
        return rec

    def _full_header(self, row_bytes):
        h = Header([("XTENSION", "BINTABLE", "binary table extension"), ("BITPIX", 8, "array data type"),
                    ("NAXIS", 2, "number of array dimensions"), ("NAXIS1", row_bytes, "length of dimension 1"),
                    ("NAXIS2", self.n_rows, "length of dimension 2"), ("PCOUNT", 0, "number of group parameters"),
                    ("GCOUNT", 1, "number of groups"), ("TFIELDS", len(self.columns), "number of table fields")])
        for i, c in enumerate(self.columns, 1):
            h.append(f"TTYPE{i}", c.name)
            h.append(f"TFORM{i}", c.tform)
            if c.unit:
                h.append(f"TUNIT{i}", c.unit)
        reserved = re.compile(r"XTENSION|BITPIX|NAXIS\d*|PCOUNT|GCOUNT|TFIELDS|(TTYPE|TFORM|TUNIT)\d+")
        for k, v, c in self.header:
            if not reserved.fullmatch(k):
                h.append(k, v, c)
        return h

    def write(self, f):
        rec = self._records()
        f.write(self._full_header(rec.dtype.itemsize).tobytes())
        rec.tofile(f)
        pad = -(rec.dtype.itemsize * len(rec)) % BLOCK
        if pad:
            f.write(b"\0" * pad)


def write_fits(path, hdus, overwrite=False):
    """Write [PrimaryHDU, BinTableHDU, ...] (a primary HDU is inserted if the list does not start with one)."""
    path = os.fspath(path)
    if os.path.exists(path) and not overwrite:
        raise OSError(f"File {path!r} already exists (use overwrite=True)")
    hdus = list(hdus)
    if not hdus or not isinstance(hdus[0], PrimaryHDU):
        hdus.insert(0, PrimaryHDU())
    tmp = f"{path}.tmp{os.getpid()}"
    try:
        with open(tmp, "wb") as f:
            for h in hdus:
                h.write(f)
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


# ------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------
def _read_header(f):
    raw = b""
    while True:
        block = f.read(BLOCK)
        if len(block) == 0 and not raw:
            return None
        if len(block) != BLOCK:
            raise FITSError("truncated FITS header")
        raw += block
        for i in range(0, BLOCK, CARD):
            if block[i:i + 8] == b"END     ":
                return Header.frombytes(raw)


def _parse_tform(tform):
    m = re.fullmatch(r"\s*(\d*)([LXBIJKAEDCMPQ])(.*)", str(tform))
    if not m:
        raise FITSError(f"cannot parse TFORM {tform!r}")
    return (int(m.group(1)) if m.group(1) else 1), m.group(2), m.group(3).strip()


def _read_bintable(f, header):
    row_bytes, n_rows = int(header["NAXIS1"]), int(header["NAXIS2"])
    pcount, nf = int(header.get("PCOUNT", 0)), int(header["TFIELDS"])
    names, formats, offsets, varlen, units = [], [], [], {}, {}
    off = 0
    for i in range(1, nf + 1):
        name = str(header.get(f"TTYPE{i}", f"col{i}")).strip()
        repeat, code, rest = _parse_tform(header[f"TFORM{i}"])
        if code in "PQ":
            el = rest[0]
            if el not in _TFORM or el == "A":
                raise FITSError(f"unsupported variable-length element type {el!r}")
            desc = ">i4" if code == "P" else ">i8"
            dt = np.dtype((desc, (2,)))
            varlen[name] = np.dtype(_TFORM[el])
            size = dt.itemsize * repeat
        elif code == "A":
            dt = np.dtype(f"S{max(repeat, 1)}")
            size = repeat
        elif code in _TFORM:
            base = np.dtype(_TFORM[code])
            dt = base if repeat == 1 else np.dtype((base, (repeat,)))
            size = base.itemsize * repeat
        else:
            raise FITSError(f"unsupported TFORM {header[f'TFORM{i}']!r}")
        if repeat == 0:
            continue
        names.append(name)
        formats.append(dt)
        offsets.append(off)
        off += size
        if f"TUNIT{i}" in header:
            units[name] = header[f"TUNIT{i}"]
    if off > row_bytes:
        raise FITSError("columns wider than NAXIS1")
    dt = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": row_bytes})
    n_main = row_bytes * n_rows
    raw = f.read(n_main)
    if len(raw) != n_main:
        raise FITSError("truncated FITS table")
    heap_skip = int(header.get("THEAP", n_main)) - n_main
    heap = f.read(pcount) if pcount else b""
    f.seek(-(n_main + pcount) % BLOCK, os.SEEK_CUR)
    rec = np.frombuffer(raw, dtype=dt, count=n_rows)
    out_dt = []
    for nme in names:
        if nme in varlen:
            out_dt.append((nme, object))
        else:
            fdt = dt.fields[nme][0]
            out_dt.append((nme, fdt.newbyteorder("=") if fdt.kind != "S" else fdt))
    out = np.empty(n_rows, dtype=out_dt)
    for nme in names:
        if nme in varlen:
            el = varlen[nme]
            for r in range(n_rows):
                n_el, start = int(rec[nme][r][0]), int(rec[nme][r][1]) + heap_skip
                out[nme][r] = np.frombuffer(heap, dtype=el, count=n_el, offset=start).astype(el.newbyteorder("="))
        else:
            out[nme] = rec[nme]
    return BinTableHDU(header=header, data=out, units=units)


def read_fits(path):
    """-> list of HDUs (``PrimaryHDU`` first, then one ``BinTableHDU`` per BINTABLE extension; other extension
    types are skipped over)."""
    hdus = []
    with open(path, "rb") as f:
        h = _read_header(f)
        if h is None or h.get("SIMPLE") is not True:
            raise FITSError(f"{path}: not a FITS file")
        naxis = int(h.get("NAXIS", 0))
        n = abs(int(h["BITPIX"])) // 8 if naxis else 0
        for i in range(1, naxis + 1):
            n *= int(h[f"NAXIS{i}"])
        f.seek(n + (-n % BLOCK), os.SEEK_CUR)
        hdus.append(PrimaryHDU(h))
        while True:
            h = _read_header(f)
            if h is None:
                break
            if h.get("XTENSION") == "BINTABLE":
                hdus.append(_read_bintable(f, h))
            else:
                naxis = int(h.get("NAXIS", 0))
                n = abs(int(h["BITPIX"])) // 8 if naxis else 0
                for i in range(1, naxis + 1):
                    n *= int(h[f"NAXIS{i}"])
                n = (n + int(h.get("PCOUNT", 0))) * int(h.get("GCOUNT", 1))
                f.seek(n + (-n % BLOCK), os.SEEK_CUR)
    return hdus


def get_hdu(hdus, name):
    for h in hdus:
        if str(h.name).strip().upper() == name.upper():
            return h
    raise KeyError(name)


__all__ = ["Header", "Column", "PrimaryHDU", "BinTableHDU", "write_fits", "read_fits", "get_hdu", "format_card",
           "FITSError"]
