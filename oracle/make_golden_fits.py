"""Cut a small real-world FITS fixture out of the GBM CSPEC file the reference ships as DATA.

TEST INFRASTRUCTURE, authoring container only (needs /root/reference).  Usage:
    python oracle/make_golden_fits.py        (writes tests/golden/glg_cspec_n0_head.fits, 14400 bytes)

cosmogrb/data/gbm_cspec/n0.pha is a Fermi-GBM PHAII file written by the GBM ground software through cfitsio
(response_generator.py:58-60 hands it to gbm_drm_gen).  Its first 14400 bytes -- primary header (2 blocks), EBOUNDS
header (2 blocks) and EBOUNDS data (128 rows x 10 bytes, 1 block) -- are a complete FITS file of their own and carry the
file's own CHECKSUM / DATASUM keywords, i.e. an answer key that does not come from this repository.
tests/test_fits_real_gbm.py holds cosmogrb_b200/utils/fits_lite.py (reader AND writer) to it.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("COSMOGRB_REFERENCE", "/root/reference")


def main():
    with open(os.path.join(REF, "cosmogrb", "data", "gbm_cspec", "n0.pha"), "rb") as f:
        raw = f.read(14400 + 80)
    assert raw[:6] == b"SIMPLE" and raw[5760:5768] == b"XTENSION" and raw[14400:14408] == b"XTENSION"
    out = os.path.join(ROOT, "tests", "golden", "glg_cspec_n0_head.fits")
    with open(out, "wb") as f:
        f.write(raw[:14400])
    print("wrote", out)


if __name__ == "__main__":
    sys.exit(main())
