"""CPU: the C-ABI library loads and exports every symbol include/cosmogrb_b200.h declares; the host
work-list helper (no device needed) behaves; the product path fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from cosmogrb_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cosmogrb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgrb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_capi.EXPORTS)
    assert _capi.load().cgrb_abi_version() == _capi.ABI_VERSION


def test_struct_layouts_match_header():
    assert _capi.UNIT_DTYPE.itemsize == 224 and _capi.COUNTS_DTYPE.itemsize == 64 and _capi.WORK_DTYPE.itemsize == 16
    src = open(os.path.join(ROOT, "include", "cosmogrb_b200.h")).read()
    for name in ("EDGES", "EA_X", "EA_Y", "BKG_CDF", "G75_E", "G75_A", "G500_E", "G500_A", "CUM"):
        m = re.search(rf"#define CGRB_DT_{name} (\d+)", src)
        assert int(m.group(1)) == getattr(_capi, f"DT_{name}")


def test_small_structs_and_constants_match_header():
    """ctypes mirrors of the small C structs and the tile / status constants the host relies on"""
    src = open(os.path.join(ROOT, "include", "cosmogrb_b200.h")).read()
    assert ctypes.sizeof(_capi.DeadtimeModel) == 24 and ctypes.sizeof(_capi.TriggerPar) == 32
    assert ctypes.sizeof(_capi.RngDesc) == 40
    for name, val in (("CGRB_TRIG_TILE", _capi.TRIG_TILE), ("CGRB_MD_TILE", _capi.MD_TILE), ("CGRB_DT_TILE", _capi.DT_TILE),
                      ("CGRB_MERGE_TILE", _capi.MERGE_TILE), ("CGRB_ABI_VERSION", _capi.ABI_VERSION),
                      ("CGRB_DEADTIME_PHYSICAL", _capi.DEADTIME_PHYSICAL)):
        m = re.search(rf"#define {name} (\d+)", src)
        assert m and int(m.group(1)) == val, name
    for name, val in (("CGRB_ST_DEADTIME_CHAIN", _capi.ST_DEADTIME_CHAIN), ("CGRB_ST_MAJORANT", _capi.ST_MAJORANT),
                      ("CGRB_ST_ENVELOPE", _capi.ST_ENVELOPE), ("CGRB_ST_OUT_CAPACITY", _capi.ST_OUT_CAPACITY)):
        m = re.search(rf"#define {name} (\d+)u", src)
        assert m and int(m.group(1)) == val, name
    # dead-time model validation happens before any launch: testable without a device
    lib = _capi.load()
    bad = _capi.DeadtimeModel(7, 0, 2.6e-6, 10e-6)
    one = np.zeros(8, dtype=np.uint8).ctypes.data
    rc = lib.cgrb_gbm_dead_time(one, 1, one, 1, one, one, one, one, one, one, None, one, ctypes.byref(bad), None)
    assert rc == -1 and b"dead-time mode" in lib.cgrb_last_error()
    neg = _capi.DeadtimeModel(_capi.DEADTIME_PHYSICAL, 0, -1.0, 10e-6)
    rc = lib.cgrb_merge_dead_time(one, 1, one, 1, one, one, one, one, one, one, one, one, one, ctypes.byref(neg), None)
    assert rc == -1 and b"finite and >= 0" in lib.cgrb_last_error()


def test_build_work():
    u = np.zeros(3, dtype=_capi.UNIT_DTYPE)
    u["src_cap"] = [0, 8192, 20000]
    u["bkg_cap"] = [5, 0, 100]
    w = _capi.build_work(u, 0, 8192)
    assert list(w["unit"]) == [0, 1, 2, 2, 2]            # every unit owns at least one entry
    assert list(w["start"]) == [0, 0, 0, 8192, 16384]
    w2 = _capi.build_work(u, 2, 2048)
    assert (w2["unit"] == 2).sum() == (20000 + 100 + 2 + 2047) // 2048


def test_bad_arguments_are_reported():
    lib = _capi.load()
    assert lib.cgrb_counts_reset(None, 1, None) == -1
    assert b"bad argument" in lib.cgrb_last_error()
    with pytest.raises(_capi.EngineError):
        _capi.check(-1, "x")


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cosmogrb_b200.engine import Engine

    with pytest.raises(_capi.EngineUnavailable):
        Engine()


def test_detector_table_guide_and_layout():
    """the search guide of the detector block: guide[b][m] = searchsorted(cum[b], m/128) -- a start at or before
    the lower bound of every r in [m/128, (m+1)/128), which is what the device walk relies on"""
    from tests import helpers as h
    from cosmogrb_b200.response.response import channel_search_guide

    src = open(os.path.join(ROOT, "include", "cosmogrb_b200.h")).read()
    assert _capi.DT_GUIDE == 1720 + 140 * 128 and _capi.DT_BKG_GUIDE == _capi.DT_GUIDE + 140 * 128 // 8
    assert _capi.DT_SIZE == _capi.DT_BKG_GUIDE + 128
    assert "#define CGRB_DT_BKG_GUIDE (1720 + 140 * 128 + 140 * 128 / 8)" in src
    assert "#define CGRB_DT_SIZE (1720 + 140 * 128 + 140 * 128 / 8 + 128)" in src
    rsp = h.response("n0")
    tab = rsp.device_table(h.bkg_weights("n0"), "linear")
    assert tab.shape == (_capi.DT_SIZE,)
    cum = rsp.cumulative_matrix
    g = tab[_capi.DT_GUIDE:_capi.DT_BKG_GUIDE].view(np.uint8).reshape(_capi.NBINS, _capi.NCHAN)
    assert np.array_equal(g.reshape(-1), channel_search_guide(cum))
    rs = np.random.RandomState(0)
    r = rs.random_sample(20000)
    b = rs.randint(0, _capi.NBINS, len(r))
    m = np.minimum((r * _capi.NCHAN).astype(int), _capi.NCHAN - 1)
    lb = np.array([np.searchsorted(cum[bi], ri, side="left") for bi, ri in zip(b, r)])
    assert np.all(g[b, m] <= lb)
    assert np.mean(lb - g[b, m]) < 3.0          # and the walk is short on the GBM-shaped matrix
    assert _capi.TRIGGER_DTYPE.itemsize == 32


def test_trigger_channel_bounds():
    """channel selection of the trigger == LightCurveStorage._select_channel (light_curve_storage.py:154-182)"""
    from tests import helpers as h
    from oracle import trigger_oracle as trg
    from cosmogrb_b200.instruments.gbm.gbm_lightcurve_analyzer import channel_bounds

    for det in ("n0", "b0"):
        eb = np.asarray(h.response(det).channel_edges)
        cb = channel_bounds(eb)
        pha = np.arange(128)
        for r, (_, emin, emax) in enumerate(trg.TRIGGER_ENERGY_RANGES):
            want = trg.select_channel(eb, emin, emax, pha)
            got = (pha > cb[2 * r]) & (pha < cb[2 * r + 1])
            assert np.array_equal(got, want)
