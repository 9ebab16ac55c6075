"""Decode the GBM input tables shipped as DATA with the reference into one small .npz.

Inputs (read-only, reference tree; run in the authoring container only):
  cosmogrb/data/gbm_backgrounds/<det>.h5 : dataset "counts", 128 x float32 (background
      spectral template used by GBMBackground, instruments/gbm/gbm_background.py:33-41)
  cosmogrb/data/gbm_cspec/<det>.pha      : FITS, HDU "EBOUNDS" (CHANNEL i2, E_MIN f4, E_MAX f4)

h5py/astropy are not installed, so both containers are decoded by hand: the HDF5 files hold a
single contiguous big-endian float32 dataset at byte offset 2048, the FITS files are parsed
from their 2880-byte header blocks.

Output: cosmogrb_b200/data/gbm_tables.npz with
  detectors[14] (str), bkg_counts[14,128] (f4), channel_edges[14,129] (f8)
"""
import os
import sys

import numpy as np

REF = os.environ.get("COSMOGRB_REFERENCE", "/root/reference")
DETS = ("n0", "n1", "n2", "n3", "n4", "n5", "n6", "n7", "n8", "n9", "na", "nb", "b0", "b1")
HERE = os.path.dirname(os.path.abspath(__file__))


def read_bkg(det):
    with open(os.path.join(REF, "cosmogrb/data/gbm_backgrounds", f"{det}.h5"), "rb") as f:
        raw = f.read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    for dt in (">f4", "<f4"):
        v = np.frombuffer(raw, dtype=dt, count=128, offset=2048)
        if np.all(np.isfinite(v)) and np.all(v >= 0) and 1e2 < v.sum() < 1e12:
            return v.astype("f4"), dt
    raise RuntimeError(f"cannot decode {det}.h5")


def read_ebounds(det):
    with open(os.path.join(REF, "cosmogrb/data/gbm_cspec", f"{det}.pha"), "rb") as f:
        raw = f.read()
    pos = 0
    while pos < len(raw):
        cards = {}
        while True:  # header blocks
            block = raw[pos:pos + 2880]
            pos += 2880
            end = False
            for i in range(0, 2880, 80):
                card = block[i:i + 80].decode("ascii", "replace")
                key = card[:8].strip()
                if key == "END":
                    end = True
                    break
                if card[8:10] == "= ":
                    cards[key] = card[10:].split("/")[0].strip().strip("'").strip()
            if end:
                break
        naxis = int(cards.get("NAXIS", 0))
        size = 0
        if naxis > 0:
            size = abs(int(cards["BITPIX"])) // 8
            for k in range(1, naxis + 1):
                size *= int(cards[f"NAXIS{k}"])
            size += int(cards.get("PCOUNT", 0))
        if cards.get("EXTNAME", "") == "EBOUNDS":
            n = int(cards["NAXIS2"])
            rec = np.frombuffer(raw, dtype=[("ch", ">i2"), ("lo", ">f4"), ("hi", ">f4")],
                                count=n, offset=pos)
            return rec["ch"].astype(int), rec["lo"].astype("f8"), rec["hi"].astype("f8")
        pos += (size + 2879) // 2880 * 2880
    raise RuntimeError(f"no EBOUNDS in {det}.pha")


def main():
    bkg = np.zeros((14, 128), "f4")
    edges = np.zeros((14, 129), "f8")
    for i, d in enumerate(DETS):
        v, dt = read_bkg(d)
        bkg[i] = v
        ch, lo, hi = read_ebounds(d)
        assert len(ch) == 128 and np.all(np.diff(ch) == 1)
        assert np.allclose(lo[1:], hi[:-1])
        edges[i, :128] = lo
        edges[i, 128] = hi[-1]
        print(d, dt, "bkg sum", float(v.sum()), "overflow frac", float(v[127] / v.sum()),
              "ebounds", lo[0], hi[-1])
    out = os.path.normpath(os.path.join(HERE, "..", "cosmogrb_b200", "data", "gbm_tables.npz"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, detectors=np.array(DETS), bkg_counts=bkg, channel_edges=edges)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    sys.exit(main())
