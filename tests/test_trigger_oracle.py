"""CPU: the numpy restatement of the GBM trigger (oracle/trigger_oracle.py) against golden vectors minted from the
REAL reference's GBMLightCurveAnalyzer (oracle/make_golden_trigger.py; no access to the reference here)."""
import os

import numpy as np
import pytest

from oracle import trigger_oracle as trg
from oracle.trigger_cases import CASES, light_curve

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trigger.npz")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", sorted(CASES))
def test_trigger_oracle_matches_reference(gold, name):
    det, params, seed = CASES[name]
    rsp, ev = light_curve(det, params, seed)
    fp = gold[f"{name}/fingerprint"]
    assert len(ev["times"]) == int(fp[0]) and ev["times"].sum() == fp[1] and float(ev["pha"].sum()) == fp[2], \
        "the regenerated light curve is not the one the golden vector was minted from"
    r = trg.analyze(ev["times"], ev["pha"], rsp.channel_edges, len(ev["times_source"]))
    assert r["is_detected"] == bool(gold[f"{name}/is_detected"])
    if r["is_detected"]:
        assert r["detection_time"] == float(gold[f"{name}/detection_time"])
        assert r["detection_time_scale"] == float(gold[f"{name}/detection_time_scale"])
    for key, emin, emax in trg.TRIGGER_ENERGY_RANGES:
        _, c = trg.binned_counts(ev["times"], ev["pha"], rsp.channel_edges, trg.BASE_TIMESCALE, emin, emax)
        assert np.array_equal(c, gold[f"{name}/counts_{key}"])
    head = gold[f"{name}/exposure_head"]
    bins, _ = trg.binned_counts(ev["times"], ev["pha"], rsp.channel_edges, trg.BASE_TIMESCALE, None, None)
    dead, _ = trg.dead_time_per_interval(ev["times"], ev["pha"].astype(np.int64), bins)
    np.testing.assert_allclose(((bins[:, 1] - bins[:, 0]) - dead)[:len(head)], head, rtol=1e-13)


def test_count_round_trip_is_lossless_for_16ms():
    """binned_counts returns ((c / dt) * dt).astype(int) (light_curve_storage.py:292); for dt = 0.016 the float
    round trip happens to be exact for every count that can occur, so the kernels' replica of it is a no-op
    (it is replicated anyway, operation for operation)"""
    c = np.arange(0, 2_000_000)
    assert np.array_equal(((c / trg.BASE_TIMESCALE) * trg.BASE_TIMESCALE).astype(int), c)


def test_simultaneous_trigger_logic():
    """GBMTrigger.process (gbm_trigger.py:119-185): ordered by angle, second trigger within +-0.5 s"""
    from cosmogrb_b200.instruments.gbm.gbm_lightcurve_analyzer import simultaneous_trigger

    names = ["n0", "n1", "n2", "b0", "n3"]
    angles = [30.0, 10.0, 20.0, 5.0, 40.0]
    mk = lambda det, t: dict(is_detected=det, detection_time=t, detection_time_scale=1.024)
    res = {"n0": mk(True, 1.3), "n1": mk(True, 1.0), "n2": mk(False, None), "n3": mk(True, 1.1), "b0": mk(True, 1.0)}
    want = trg.gbm_trigger(dict(zip(names, angles)), res)
    got = simultaneous_trigger(names, angles, [res[n]["is_detected"] for n in names],
                               [res[n]["detection_time"] or 0.0 for n in names], [1.024] * 5)
    assert want[0] is True and want[1] == ["n1", "n0"]
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
    # too far apart: no detection, all NaI detectors tested
    res["n0"] = mk(True, 2.0)
    res["n3"] = mk(True, 3.0)
    want = trg.gbm_trigger(dict(zip(names, angles)), res)
    got = simultaneous_trigger(names, angles, [res[n]["is_detected"] for n in names],
                               [res[n]["detection_time"] or 0.0 for n in names], [1.024] * 5)
    assert want[0] is False and got[0] is False and got[1] == want[1] == ["n1", "n0", "n3"]
    # max_n_dets limits the search
    assert simultaneous_trigger(names, angles, [True] * 5, [1.0] * 5, [1.0] * 5, max_n_dets=1)[0] is False
