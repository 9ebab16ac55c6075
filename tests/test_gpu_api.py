"""GPU: the reference-facing API surface (GBMGRB_CPL(...).go(), Source / Background / Response
sampler classes, GRB.save + GRBSave.from_file, Universe runs) on the CUDA engine."""
import numpy as np
import pytest

from tests import helpers as h
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
stats = pytest.importorskip("scipy.stats")


@pytest.fixture(scope="module")
def grb(engine):
    from cosmogrb_b200 import gbm

    # test/conftest.py:31-53 (the reference's bright fixture)
    g = gbm.GBMGRB_CPL(ra=312.0, dec=-62.0, z=1.0, peak_flux=1e-5, alpha=-0.66, ep_start=500.0, ep_tau=2.0,
                       trise=0.1, tdecay=1.0, duration=2.0, T0=0.1)
    g.go(seed=2024)
    return g


def test_gbmgrb_cpl_go(grb):
    names = list(grb.lightcurves.keys())
    assert names == list(h.GBM_DETECTORS)
    for name, lc in grb.lightcurves.items():
        st = lc.lightcurve_storage
        assert st.name == name and st.instrument == "GBM" and st.tstart == -100 and st.tstop == 300
        assert st.times_source[0] == 0.0 and st.times_background[0] == -100.0       # the spurious tstart events
        assert np.all(np.diff(st.times) >= 0) and np.all(np.diff(st.times_source) >= 0)
        assert st.pha.dtype.kind == "i" and st.pha_source.dtype == np.float64 and st.pha_background.dtype == np.int64
        assert st.pha.min() >= 0 and st.pha.max() <= 127
        n_pre = st.n_counts_source + st.n_counts_background
        assert 0 < n_pre - st.n_counts < 0.01 * n_pre                   # dead time removes a few events
        assert abs(st.n_counts_background - 1 - 500 * 400) < 6 * np.sqrt(2e5) + 40 * 400    # rate ~ N(500, 10)
        # the total list is the dead-time filtered merge of the two component lists
        wt, wp = orc.combine(st.times_background, st.pha_background, st.times_source, st.pha_source)
        sel = orc.gbm_dead_time(wt, wp)
        assert np.array_equal(st.times, wt[sel]) and np.array_equal(st.pha, wp[sel].astype(int))
    # brighter in the detectors with the larger effective area
    n_src = {k: lc.lightcurve_storage.n_counts_source for k, lc in grb.lightcurves.items()}
    ea = {k: lc.response.effective_area_packed[1].max() for k, lc in grb.lightcurves.items() if k[0] == "n"}
    best, worst = max(ea, key=ea.get), min(ea, key=ea.get)
    assert n_src[best] > n_src[worst]


def test_go_pipelined_equals_single_batch(grb, monkeypatch):
    """Large calls overlap the device->host copy of one chunk of detectors with the kernels of the next;
    the chunking changes nothing (Philox counters are per unit)."""
    from cosmogrb_b200.lightcurve import lightcurve as lcmod

    units = np.concatenate([lc.unit_record(det_index=i) for i, lc in enumerate(grb.lightcurves.values())])
    assert len(lcmod._chunks(units, None)) == 1
    want = {}
    for name, lc in grb.lightcurves.items():
        st = lc.lightcurve_storage
        want[name] = {a: np.array(getattr(st, a)) for a in ("times", "pha", "times_source", "pha_source",
                                                             "times_background", "pha_background")}
    monkeypatch.setattr(lcmod, "PIPELINE_MIN_SLOTS", 1000)
    assert len(lcmod._chunks(units, None)) > 1
    grb.go(seed=2024)                       # same burst object (same background rates), now pipelined
    for name, lc in grb.lightcurves.items():
        st = lc.lightcurve_storage
        for attr, w in want[name].items():
            assert np.array_equal(getattr(st, attr), w), (name, attr)


def test_grb_save_and_reload(grb, tmp_path):
    from cosmogrb_b200 import GRBSave

    fn = tmp_path / "test_grb.h5"
    grb.save(fn)
    g = GRBSave.from_file(fn)
    assert g.name == "SynthGRB" and g.z == 1.0 and g.duration == 2.0 and g.T0 == 0.1
    assert g.source_params["peak_flux"] == 1e-5
    for name, lc in grb.lightcurves.items():
        st, got = lc.lightcurve_storage, g[name]["lightcurve"]
        assert np.array_equal(got.times, st.times) and np.array_equal(got.pha, st.pha)
        assert np.array_equal(got.times_source, st.times_source)
        assert np.array_equal(g[name]["response"].matrix, lc.response.matrix)
        assert got.extra_info["angle"] == pytest.approx(lc.response.separation_angle)


def test_grb_to_gbm_fits_products(grb, tmp_path):
    """go() -> GRB.save -> grbsave_to_gbm_fits (io/gbm_fits.py:9-83): one TTE + one RSP file per detector whose
    events are the dead-time-filtered total list; GBMLightCurve.write_tte (gbm_lightcurve.py:58-77) writes the
    same events; a response read back from its RSP file drives the engine to the same channels."""
    from cosmogrb_b200.io import grbsave_to_gbm_fits
    from cosmogrb_b200.response import Response
    from cosmogrb_b200.utils.tte_file import read_tte

    fn = tmp_path / "burst.h5"
    grb.save(fn)
    files = grbsave_to_gbm_fits(str(fn), destination=str(tmp_path), detectors=["n0", "b1"])
    assert sorted(files) == ["b1", "n0"]
    for det, f in files.items():
        st = grb.lightcurves[det].lightcurve_storage
        r = read_tte(f["tte"])
        assert np.array_equal(r["time"], st.times + st.time_adjustment) and np.array_equal(r["pha"], st.pha)
        assert r["det_name"] == {"n0": "NAI_00", "b1": "BGO_01"}[det]
        rsp = grb.lightcurves[det].response
        r2 = Response.from_fits(f["rsp"], geometric_area=rsp.geometric_area)
        assert np.array_equal(r2.cumulative_matrix, rsp.cumulative_matrix)
        # photon edges went through float32 in the file: digitize agrees wherever the energy is not within
        # float32 rounding of an edge
        rs = np.random.RandomState(3)
        e = 10 ** rs.uniform(np.log10(rsp.emin * 1.01), np.log10(rsp.emax * 0.99), 20000)
        near = np.min(np.abs(e[:, None] / rsp.energy_edges[None, :] - 1.0), axis=1) < 1e-6
        from cosmogrb_b200.engine import Philox, get_engine
        eng = get_engine()
        a, b = eng.digitize(rsp, e, rng=Philox(5)), eng.digitize(r2, e, rng=Philox(5))
        assert near.sum() < 100 and np.array_equal(a[~near], b[~near])
    # a burst whose n0 / b1 responses come from those files (real-DRM route): same matrices, so the same light
    # curves in distribution as the fixture's
    from cosmogrb_b200 import gbm

    g2 = gbm.GBMGRB_CPL(ra=312.0, dec=-62.0, z=1.0, peak_flux=1e-5, alpha=-0.66, ep_start=500.0, ep_tau=2.0,
                        trise=0.1, tdecay=1.0, duration=2.0, T0=0.1,
                        response_files={d: f["rsp"] for d, f in files.items()})
    g2.go(seed=77)
    for det in ("n0", "b1", "n5"):
        a = grb.lightcurves[det].lightcurve_storage.n_counts_source
        b = g2.lightcurves[det].lightcurve_storage.n_counts_source
        assert abs(a - b) < 6.0 * np.sqrt(a + b) + 10, (det, a, b)
    assert type(g2.lightcurves["n0"].response).__name__ == "GBMResponseFile"
    assert type(g2.lightcurves["n5"].response).__name__ == "NaIResponse"
    lc = grb.lightcurves["n3"]
    out = lc.write_tte(str(tmp_path / "n3_direct.fits"))
    r = read_tte(out)
    st = lc.lightcurve_storage
    assert np.array_equal(r["time"], st.times + lc.response.T0) and np.array_equal(r["pha"], st.pha)


def test_go_is_reproducible_and_seed_dependent(engine):
    from cosmogrb_b200 import gbm

    np.random.seed(3)
    kw = dict(ra=10.0, dec=20.0, z=1.0, peak_flux=5e-7, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=1.0,
              tdecay=1.0, duration=20.0, T0=0.1)
    a = gbm.GBMGRB_CPL(**kw)
    np.random.seed(3)
    b = gbm.GBMGRB_CPL(**kw)
    a.go(seed=1)
    b.go(seed=1)
    assert np.array_equal(a.lightcurves["n3"].times, b.lightcurves["n3"].times)
    b.go(seed=2)
    assert not np.array_equal(a.lightcurves["n3"].times_source, b.lightcurves["n3"].times_source)


def test_constant_cpl_grb(engine):
    from cosmogrb_b200 import gbm

    g = gbm.GBMGRB_CPL_Constant(ra=312.0, dec=-62.0, z=1.0, peak_flux=5e-9, alpha=-0.66, ep=500.0, duration=2.0, T0=0.1)
    g.go()
    st = g.lightcurves["n0"].lightcurve_storage
    lam = g.lightcurves["n0"]._source.fmax
    assert abs(st.n_counts_source - 1 - 2.0 * lam) < 6 * np.sqrt(2.0 * lam) + 3


def test_sampler_classes_standalone(engine):
    """Source / CPLSourceFunction / Background / Response used one call at a time, like the
    reference's LightCurve._sample_source (lightcurve.py:74-129)."""
    from cosmogrb_b200.instruments.gbm.gbm_background import GBMBackground
    from cosmogrb_b200.sampler import CPLSourceFunction, Source

    p = h.CONFTEST_BRIGHT
    rsp = h.response("n0")
    sf = CPLSourceFunction(response=rsp, peak_flux=p["peak_flux"], trise=p["trise"], tdecay=p["tdecay"],
                           ep_tau=p["ep_tau"], alpha=p["alpha"], ep_start=p["ep_start"])
    src = Source(0.0, p["duration"], sf, z=p["z"])
    s = h.oracle_source(rsp, **p)
    assert src.fmax == pytest.approx(orc.fmax(s, rsp.effective_area_packed, 0.0, p["duration"]), rel=1e-11)
    assert sf.energy_integrated_evolution(0.5) == pytest.approx(
        orc.energy_integrated_evolution(s, rsp.effective_area_packed, [0.5])[0], rel=1e-11)
    evo = sf.evolution(np.array([10.0, 100.0]), np.array([0.2, 0.5, 1.0]))
    np.testing.assert_allclose(evo, orc.cpl_evolution(s, [10.0, 100.0], [0.2, 0.5, 1.0]), rtol=1e-11)
    times = src.sample_times()
    assert times[0] == 0.0 and np.all(np.diff(times) >= 0) and times[-1] <= p["duration"]
    ref, _ = orc.sample_events(s, rsp.effective_area_packed, 0.0, p["duration"], src.fmax, orc.Rng(seed=5))
    assert abs(len(times) - len(ref)) < 6 * np.sqrt(len(ref))
    assert stats.ks_2samp(times[1:], ref[1:]).pvalue > 0.01
    photons = src.sample_photons(times)
    assert photons.shape == times.shape and np.all(photons >= rsp.emin * (1 - 1e-12))
    pha = src.sample_channel(photons, rsp)
    assert pha.dtype == np.float64 and pha.min() >= 0 and pha.max() <= 127
    bkg = GBMBackground(-100, 300, average_rate=500, detector="n0")
    tb = bkg.sample_times()
    assert tb[0] == -100 and abs(len(tb) - 400 * bkg.background_rate) < 6 * np.sqrt(2e5)
    cb = bkg.sample_channel(size=len(tb))
    assert cb.dtype == np.int64 and len(cb) == len(tb)
    w = h.bkg_weights("n0").astype("f8")
    obs = np.bincount(cb, minlength=128)
    exp = w / w.sum() * len(cb)
    ok = exp >= 10
    assert stats.chi2.sf(((obs[ok] - exp[ok]) ** 2 / exp[ok]).sum(), ok.sum() - 1) > 0.01


def test_universe_run(engine, tmp_path):
    from cosmogrb_b200 import GRBSave
    from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe
    from cosmogrb_b200.universe import synthetic_population

    pop = synthetic_population(12, seed=7)
    fn = tmp_path / "pop.npz"
    pop.writeto(fn)
    uni = GBM_CPL_Universe(str(fn), save_path=str(tmp_path))
    uni.go(seed=11)
    assert uni.event_counts.shape == (12, 14, 3)
    assert np.all(uni.event_counts[:, :, 1] > 1.8e5)
    for i in (0, 11):
        g = GRBSave.from_file(tmp_path / f"SynthGRB_{i}_store.h5")
        assert g.name == f"SynthGRB_{i}" and g.z == pytest.approx(pop.distances[i])
        for d, det in enumerate(h.GBM_DETECTORS):
            st = g[det]["lightcurve"]
            assert (st.n_counts_source, st.n_counts_background, st.n_counts) == tuple(uni.event_counts[i, d])
    uni.save(tmp_path / "survey.h5")
    # results do not depend on the batching (Philox counters are functions of the unit)
    uni2 = GBM_CPL_Universe(str(fn), save_path=None)
    uni2.go(seed=11, grbs_per_batch=5)
    assert uni2.n_batches == 3
    assert np.array_equal(uni2.event_counts, uni.event_counts)


def test_runaway_unit_is_refused_and_skipped(engine):
    """The reference's model lets lambda(t) explode for low-Epeak bursts at late times (the 10-10^4 keV
    normalisation band drifts above the cutoff): fmax x duration reaches 1e15 candidates.  The engine
    refuses such a unit loudly instead of sizing buffers for it, and a population run skips the burst."""
    from cosmogrb_b200 import _capi
    from cosmogrb_b200.instruments.gbm.gbm_universe import GBM_CPL_Universe
    from cosmogrb_b200.universe import synthetic_population
    from cosmogrb_b200.universe.population import Population

    bad = dict(z=4.72, peak_flux=1e-9, alpha=-0.853, ep_start=13.28, ep_tau=2.0, trise=1.0, tdecay=4.0,
               duration=28.3)
    rsp = h.response("n0")
    with pytest.raises(_capi.EngineError, match="max_candidates_per_unit"):
        engine.new_batch(h.unit(**bad), [rsp.device_table(h.bkg_weights("n0"), "linear")])
    pop = synthetic_population(3, seed=7)
    cols = {k: np.array(getattr(pop, k) if k != "fluxes_latent" else pop.fluxes.latent) for k in
            ("ra", "dec", "distances", "duration", "ep", "alpha", "fluxes_latent", "trise", "tdecay", "tau")}
    cols["distances"][1], cols["fluxes_latent"][1], cols["alpha"][1] = bad["z"], bad["peak_flux"], bad["alpha"]
    cols["ep"][1], cols["tau"][1], cols["trise"][1] = bad["ep_start"], bad["ep_tau"], bad["trise"]
    cols["tdecay"][1], cols["duration"][1] = bad["tdecay"], bad["duration"]
    uni = GBM_CPL_Universe(Population(**cols), save_path=None)
    uni.go(seed=3, save=False)
    assert uni.skipped == [1]
    assert np.all(uni.event_counts[1] == -1) and np.all(uni.event_counts[[0, 2], :, 1] > 1.8e5)
    # the switch: runaway="raise" refuses the run before anything is simulated
    uni2 = GBM_CPL_Universe(Population(**cols), save_path=None)
    with pytest.raises(_capi.EngineError, match="do not fit the memory budget"):
        uni2.go(seed=3, save=False, runaway="raise")
    with pytest.raises(ValueError):
        uni2.go(seed=3, save=False, runaway="truncate")


def test_c1_readme_demo(engine):
    """BASELINE.json configs[0]: the README burst through GBMGRB_CPL(...).go().  Per detector the number of
    source events is Poisson around the integral of lambda(t) (oracle, trapezoid on a fine grid) and the
    background fills [-100, 300] s at ~500 Hz."""
    from cosmogrb_b200 import gbm

    g = gbm.GBMGRB_CPL(ra=312.0, dec=-62.0, z=1.0, peak_flux=5e-7, alpha=-0.66, ep_start=500.0, ep_tau=2.0,
                       trise=1.0, tdecay=1.0, duration=80.0, T0=0.1, n_cores=6)     # unknown kwargs are dropped
    g.go(seed=31)
    tg = np.linspace(0.0, 80.0, 4001)
    for name in ("n0", "n3", "n9", "b0"):
        lc = g.lightcurves[name]
        rsp = lc.response
        s = h.oracle_source(rsp, **h.README_DEMO)
        lam = orc.energy_integrated_evolution(s, rsp.effective_area_packed, tg)
        mean = np.trapezoid(lam, tg)
        st = lc.lightcurve_storage
        n_src = st.n_counts_source - 1
        assert abs(n_src - mean) < 5.0 * np.sqrt(mean) + 5.0, (name, n_src, mean)
        assert st.times_source[-1] <= 80.0 and st.times_background[0] == -100.0
        assert abs(st.n_counts_background - 2e5) < 6 * np.sqrt(2e5) + 40 * 400
