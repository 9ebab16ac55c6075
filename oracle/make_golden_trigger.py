"""Mint GBM-trigger golden vectors from the REAL reference.

TEST INFRASTRUCTURE, authoring container only (needs /root/reference + numba).  Usage:
    python oracle/make_golden_trigger.py            (writes tests/golden/trigger.npz)

Light curves come from the C oracle's pipeline (MT19937 streams); each is wrapped in the reference's own
``LightCurveStorage`` and analysed by the reference's own ``GBMLightCurveAnalyzer`` (unmodified code from
/root/reference/cosmogrb/instruments/gbm/gbm_lightcurve_analyzer.py).  The numpy restatement in
oracle/trigger_oracle.py is checked against every vector as it is minted.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    from oracle import ref_shims, trigger_oracle as trg
    from oracle.trigger_cases import CASES, light_curve

    ref_shims.install("linear")
    import importlib
    import types

    # the analyzer module imports dask (unused) and expects the package __init__ to export the base class
    sys.modules.setdefault("dask", types.ModuleType("dask"))
    lcs_mod = importlib.import_module("cosmogrb.lightcurve.light_curve_storage")
    base_mod = importlib.import_module("cosmogrb.lightcurve.lightcurve_analyzer")
    sys.modules["cosmogrb.lightcurve"].LightCurveAnalyzer = base_mod.LightCurveAnalyzer
    ana_mod = importlib.import_module("cosmogrb.instruments.gbm.gbm_lightcurve_analyzer")

    cases = CASES
    out = {}
    for name, (det, params, seed) in cases.items():
        rsp, ev = light_curve(det, params, seed)
        lc = lcs_mod.LightCurveStorage(
            name=det, tstart=-100, tstop=300, time_adjustment=0.0, pha=ev["pha"], times=ev["times"],
            pha_source=ev["pha_source"], times_source=ev["times_source"], pha_background=ev["pha_background"],
            times_background=ev["times_background"], channels=rsp.channels, ebounds=rsp.channel_edges, T0=0.0,
            instrument="GBM", extra_info={})
        a = ana_mod.GBMLightCurveAnalyzer(lc, threshold=4.5)
        ref = dict(is_detected=bool(a.is_detected), detection_time=a.detection_time,
                   detection_time_scale=a.detection_time_scale)
        mine = trg.analyze(ev["times"], ev["pha"], rsp.channel_edges, len(ev["times_source"]))
        print(name, ref, "razor" if mine["razor"] else "")
        assert mine["is_detected"] == ref["is_detected"], name
        if ref["is_detected"]:
            assert mine["detection_time"] == ref["detection_time"], (name, mine["detection_time"], ref)
            assert mine["detection_time_scale"] == ref["detection_time_scale"], name
        # binned products of the reference classes, for the stage-level checks
        bins, ctot = lc.binned_counts(trg.BASE_TIMESCALE, None, None)
        assert np.array_equal(bins[:, 0], mine["starts"]) and np.array_equal(bins[:, 1], mine["stops"])
        for key, emin, emax in trg.TRIGGER_ENERGY_RANGES:
            _, c = lc.binned_counts(trg.BASE_TIMESCALE, emin, emax)
            _, c_mine = trg.binned_counts(ev["times"], ev["pha"], rsp.channel_edges, trg.BASE_TIMESCALE, emin, emax)
            assert np.array_equal(c, c_mine), (name, key)
            out[f"{name}/counts_{key}"] = c.astype(np.int32)
        dead = np.array([a.dead_time_of_interval(x, y) for x, y in bins[:2000]])
        expo = (bins[:2000, 1] - bins[:2000, 0]) - dead
        np.testing.assert_allclose(expo, mine["exposure"][:2000], rtol=1e-13)
        # the events themselves are regenerated from the seeds by the (pinned) C oracle: keep a fingerprint
        out[f"{name}/fingerprint"] = np.array([len(ev["times"]), ev["times"].sum(), float(ev["pha"].sum())])
        out[f"{name}/n_counts_source"] = np.int64(len(ev["times_source"]))
        out[f"{name}/is_detected"] = np.bool_(ref["is_detected"])
        out[f"{name}/detection_time"] = np.float64(ref["detection_time"] if ref["is_detected"] else np.nan)
        out[f"{name}/detection_time_scale"] = np.float64(ref["detection_time_scale"] if ref["is_detected"] else np.nan)
        out[f"{name}/exposure_head"] = expo
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "trigger.npz"), **out)
    print("wrote", os.path.join(GOLD, "trigger.npz"))


if __name__ == "__main__":
    main()
