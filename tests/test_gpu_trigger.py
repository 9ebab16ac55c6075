"""GPU: the on-device GBM trigger (cgrb_gbm_trigger) against golden vectors of the REAL reference's
GBMLightCurveAnalyzer and against the numpy oracle; the reference-facing classes on top of it."""
import os

import numpy as np
import pytest

from oracle import trigger_oracle as trg
from oracle.trigger_cases import CASES, light_curve
from tests import helpers as h

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trigger.npz")


def _storage(det, rsp, ev, angle=0.0):
    from cosmogrb_b200.lightcurve.light_curve_storage import LightCurveStorage

    return LightCurveStorage(name=det, tstart=-100, tstop=300, time_adjustment=0.0, pha=ev["pha"], times=ev["times"],
                             pha_source=ev["pha_source"], times_source=ev["times_source"],
                             pha_background=ev["pha_background"], times_background=ev["times_background"],
                             channels=rsp.channels, ebounds=rsp.channel_edges, T0=0.0, instrument="GBM",
                             extra_info={"angle": angle})


@pytest.mark.parametrize("name", sorted(CASES))
def test_analyzer_matches_reference(engine, name):
    """GBMLightCurveAnalyzer on the device == the reference's (golden): flag, time and time scale exactly
    (window sums differ from the reference's in the last bits: a decision may only differ if the oracle flags
    a significance within 1e-9 of the threshold)"""
    from cosmogrb_b200.instruments.gbm.gbm_lightcurve_analyzer import GBMLightCurveAnalyzer

    with np.load(GOLD) as z:
        gold = {k: z[k] for k in z.files if k.startswith(name + "/")}
    det, params, seed = CASES[name]
    rsp, ev = light_curve(det, params, seed)
    a = GBMLightCurveAnalyzer(_storage(det, rsp, ev))
    want = bool(gold[f"{name}/is_detected"])
    if a._result["razor"]:
        pytest.skip("a window sits within 1e-9 of the threshold: decision not comparable")
    assert a.is_detected == want
    if want:
        assert a.detection_time == float(gold[f"{name}/detection_time"])
        assert a.detection_time_scale == float(gold[f"{name}/detection_time_scale"])
    else:
        assert a.detection_time is None and a.detection_time_scale is None
    assert a._result["n_bins"] == len(gold[f"{name}/counts_a"])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_trigger_against_oracle_random_light_curves(engine, seed):
    """synthetic light curves with a step in the rate, events on bin edges, overflow channels: device ==
    numpy oracle (thresholds scanned so that several time scales / energy ranges fire)"""
    from cosmogrb_b200.instruments.gbm.gbm_lightcurve_analyzer import channel_bounds

    rs = np.random.RandomState(seed)
    rsp = h.response("n0")
    eb = np.asarray(rsp.channel_edges)
    t = np.sort(np.concatenate([rs.uniform(-100.0, 300.0, 150_000), rs.uniform(5.0, 5.0 + 0.3 * seed, 400 * seed)]))
    grid = np.arange(t[0], t[-1], 0.016)
    t[1000:1010] = grid[70:80]                   # events exactly on bin edges
    t.sort()
    pha = rs.randint(0, 128, len(t))
    pha[rs.random_sample(len(t)) < 0.05] = 127
    for thr in (3.0, 4.5, 6.0, 9.0, 30.0):
        want = trg.analyze(t, pha, eb, 10, threshold=thr)
        got = engine.analyze_light_curve(t, pha, channel_bounds(eb), 10, threshold=thr)
        if want["razor"] or got["razor"]:
            continue
        assert bool(got["detected"]) == want["is_detected"], thr
        if want["is_detected"]:
            assert got["time"] == want["detection_time"] and got["time_scale"] == want["detection_time_scale"], thr
    # no source counts: the analyzer does nothing (lightcurve_analyzer.py:22-34)
    assert not engine.analyze_light_curve(t, pha, channel_bounds(eb), 0)["detected"]


def test_gbm_trigger_on_saved_burst_and_universe(engine, tmp_path):
    """GBMTrigger on saved bursts == the on-device trigger of the population run that wrote them"""
    from cosmogrb_b200.instruments.gbm.gbm_trigger import GBMTrigger
    from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe
    from cosmogrb_b200.universe import synthetic_population
    from cosmogrb_b200.universe.population import Population

    pop = synthetic_population(6, seed=7)
    cols = {k: np.array(getattr(pop, k) if k != "fluxes_latent" else pop.fluxes.latent) for k in
            ("ra", "dec", "distances", "duration", "ep", "alpha", "fluxes_latent", "trise", "tdecay", "tau")}
    cols["fluxes_latent"][:3] = 2e-6          # three bursts that must trigger, three left as drawn
    cols["ep"][:3], cols["distances"][:3] = 300.0, 1.0
    uni = GBM_CPL_Universe(Population(**cols), save_path=str(tmp_path))
    uni.go(seed=5, trigger=True)
    assert uni.detected.shape == (6,) and uni.detected[:3].all()
    for i in range(6):
        g = GBMTrigger(tmp_path / f"SynthGRB_{i}_store.h5")
        g.process()
        assert g.is_detected == bool(uni.detected[i]), i
        tr = uni.trigger_results[i]
        for name, t, sc in zip(g.triggered_detectors, g.triggered_times, g.triggered_time_scales):
            d = h.GBM_DETECTORS.index(name)
            assert tr[d]["detected"] and tr[d]["time"] == t and tr[d]["time_scale"] == sc
