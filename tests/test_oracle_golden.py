"""CPU: the C oracle against the golden vectors minted from the REAL reference
(oracle/make_golden.py imported the numba kernels and the real Source / Background / Response /
GBMLightCurve classes from the reference tree).  Nothing here reads the reference."""
import os

import numpy as np
import pytest

from oracle import make_golden as mg
from oracle import oracle as orc
from tests import helpers as h

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"readme": (h.README_DEMO, 0), "bright": (h.CONFTEST_BRIGHT, 0), "long": (h.BRIGHT_LONG, 0),
         "weak": (mg.WEAK, 0), "constant": (mg.CONSTANT, 1)}


def load(name):
    with np.load(os.path.join(GOLD, name)) as d:
        return {k: d[k] for k in d.files}


@pytest.mark.parametrize("extrap", ["linear", "clamp"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_lambda_fmax(extrap, case):
    g = load(f"spectral_{extrap}.npz")
    p, kind = CASES[case]
    rsp = h.response("n0")
    s = h.oracle_source(rsp, kind=kind, interp_extrap=extrap, **p)
    got = orc.energy_integrated_evolution(s, rsp.effective_area_packed, g[f"lambda_{case}_t"])
    np.testing.assert_allclose(got, g[f"lambda_{case}"], rtol=1e-12, atol=1e-300)
    fm = float(g[f"fmax_{case}"])
    assert abs(orc.fmax(s, rsp.effective_area_packed, 0.0, p["duration"]) - fm) <= 1e-12 * fm


def test_evolution():
    g = load("spectral_linear.npz")
    rsp = h.response("n0")
    s = h.oracle_source(rsp, **h.README_DEMO)
    np.testing.assert_allclose(orc.cpl_evolution(s, g["evo_e"], g["evo_t"]), g["evo"], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("extrap", ["linear", "clamp"])
@pytest.mark.parametrize("case", ["readme", "bright"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_source_stream_replay(extrap, case, seed):
    """MT19937(seed) replay of sample_events -> sample_energy -> _digitize on one stream."""
    g = load(f"spectral_{extrap}.npz")
    p, kind = CASES[case]
    rsp = h.response("n0")
    ea = rsp.effective_area_packed
    s = h.oracle_source(rsp, interp_extrap=extrap, **p)
    rng = orc.Rng(seed=seed)
    k = f"{case}_s{seed}"
    t, _ = orc.sample_events(s, ea, 0.0, p["duration"], float(g[f"fmax_{case}"]), rng)
    assert len(t) == len(g[f"events_{k}"])          # same accept/reject decisions
    np.testing.assert_allclose(t, g[f"events_{k}"], rtol=1e-12)
    n_e = len(g[f"energy_{k}"])
    e, tr = orc.sample_energy(s, ea, g[f"events_{k}"][:n_e], rng)
    np.testing.assert_allclose(e, g[f"energy_{k}"], rtol=1e-12)
    assert np.array_equal(tr, g[f"trials_{k}"])
    pha = orc.digitize(g[f"energy_{k}"], rsp.energy_edges, rsp.cumulative_matrix, rng)
    assert np.array_equal(pha, g[f"pha_{k}"])


def test_constant_cpl_stream_replay():
    g = load("spectral_linear.npz")
    p = mg.CONSTANT
    rsp = h.response("n0")
    ea = rsp.effective_area_packed
    s = h.oracle_source(rsp, kind=1, **p)
    rng = orc.Rng(seed=4)
    t, _ = orc.sample_events(s, ea, 0.0, p["duration"], float(g["fmax_constant"]), rng)
    assert len(t) == len(g["events_constant_s4"])
    np.testing.assert_allclose(t, g["events_constant_s4"], rtol=1e-12)
    e, tr = orc.sample_energy(s, ea, g["events_constant_s4"][:300], rng)
    np.testing.assert_allclose(e, g["energy_constant_s4"], rtol=1e-12)


@pytest.mark.parametrize("det", ["n0", "b0"])
def test_digitize(det):
    g = load("digitize.npz")
    rsp = h.response(det)
    e, _ = mg.digitize_inputs(rsp, 60000)
    got = orc.digitize(e, rsp.energy_edges, rsp.cumulative_matrix, orc.Rng(seed=21))
    assert np.array_equal(got, g[f"pha_{det}"].astype("f8"))


def test_dead_time():
    g = load("dead_time.npz")
    assert np.array_equal(orc.gbm_dead_time(g["crafted_t"], g["crafted_pha"]), g["crafted_sel"])
    for i in range(4):
        rate, frac, n = g[f"random{i}_args"]
        t, pha = mg.dead_time_inputs(i, rate, frac, int(n))
        want = np.unpackbits(g[f"random{i}_sel"])[:int(n)].astype(bool)
        assert np.array_equal(orc.gbm_dead_time(t, pha), want)


def test_dead_time_physical_definition():
    """The non-paralyzable model (not computed by the reference): hand-worked case and a
    pure-Python replay of the serial definition."""
    t = np.array([0, 1, 3, 4, 10, 15, 20.5, 20.7, 31.2, 32]) * 1e-6
    pha = np.array([5, 6, 7, 8, 127, 9, 10, 11, 127, 12])
    # windows: 0 -> 2.6; 3 -> 5.6 (4 dropped); 10 (ovf) -> 20; 15 dropped; 20.5 -> 23.1 (20.7 dropped);
    # 31.2 (ovf) -> 41.2; 32 dropped
    want = np.array([1, 0, 1, 0, 1, 0, 1, 0, 1, 0], dtype=bool)
    assert np.array_equal(orc.gbm_dead_time_physical(t, pha), want)
    # the reference's literal rule differs on this input (only overflow events move t_end)
    assert not np.array_equal(orc.gbm_dead_time(t, pha), want)
    rs = np.random.RandomState(3)
    n = 20000
    t = np.cumsum(rs.exponential(1.0 / 2e5, n))
    pha = rs.randint(0, 128, n)
    pha[rs.random_sample(n) < 0.1] = 127
    sel = np.zeros(n, bool)
    t_end = -np.inf
    for i in range(n):
        if i == 0 or t[i] > t_end:
            sel[i] = True
            t_end = t[i] + (10e-6 if pha[i] == 127 else 2.6e-6)
    assert np.array_equal(orc.gbm_dead_time_physical(t, pha, 2.6e-6, 10e-6), sel)
    assert len(orc.gbm_dead_time_physical(t[:0], pha[:0])) == 0
    # zero-length windows keep everything that is strictly later than its predecessor
    assert orc.gbm_dead_time_physical(t, pha, 0.0, 0.0).all()


def test_background():
    g = load("background.npz")
    t = orc.background_times(-10.0, 30.0, 497.25, orc.Rng(seed=9))
    assert len(t) == len(g["times"])
    np.testing.assert_allclose(t, g["times"], rtol=1e-12, atol=1e-12)
    c = orc.background_channels(h.bkg_weights("n0"), len(t), orc.Rng(seed=77))
    assert np.array_equal(c, g["channels"].astype(np.int64))


def test_pipeline_real_classes():
    """The real GBMLightCurve.process() (Source, CPLSourceFunction, Background, Response of the
    reference) for one detector, replayed by chaining the oracle stages on the same two streams."""
    g = load("pipeline.npz")
    p = h.CONFTEST_BRIGHT
    rsp = h.response("n0")
    ea = rsp.effective_area_packed
    s = h.oracle_source(rsp, **p)
    fmax = orc.fmax(s, ea, 0.0, p["duration"])
    assert abs(fmax - float(g["fmax"])) <= 1e-12 * fmax
    rng = orc.Rng(seed=int(g["seeds"][0]))
    ts, _ = orc.sample_events(s, ea, 0.0, p["duration"], fmax, rng)
    es, _ = orc.sample_energy(s, ea, ts, rng)
    ps = orc.digitize(es, rsp.energy_edges, rsp.cumulative_matrix, rng)
    tb = orc.background_times(-20.0, 30.0, float(g["bkg_rate"]), rng)
    pb = orc.background_channels(h.bkg_weights("n0"), len(tb), orc.Rng(seed=int(g["seeds"][1])))
    tt, pt = orc.combine(tb, pb, ts, ps)
    sel = orc.gbm_dead_time(tt, pt)
    np.testing.assert_allclose(ts, g["times_source"], rtol=1e-12)
    assert np.array_equal(ps, g["pha_source"].astype("f8"))
    np.testing.assert_allclose(tb, g["times_background"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(pb, g["pha_background"].astype(np.int64))
    np.testing.assert_allclose(tt[sel], g["times"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(pt[sel], g["pha"].astype("f8"))


def test_gammaincc_against_scipy():
    sc = pytest.importorskip("scipy.special")
    rs = np.random.RandomState(0)
    a = rs.uniform(0.05, 2.0, 5000)
    x = 10 ** rs.uniform(-4, 3, 5000)
    want = sc.gammaincc(a, x)
    got = orc.gammaincc(a, x)
    ok = want > 1e-290
    assert np.max(np.abs(got[ok] - want[ok]) / want[ok]) < 1e-12


# ---- second golden set (oracle/make_golden_r2.py): ConstantCPL at ~1300 events, BGO / n5 detectors, a window of
# ---- the benchmarked bright long burst
EXTRA_CASES = [("constant_n0_s4", "n0", 1, 4), ("constant_n0_s5", "n0", 1, 5), ("constant_b0_s4", "b0", 1, 4),
               ("constant_b0_s5", "b0", 1, 5), ("bright_n5_s0", "n5", 0, 0), ("bright_b0_s0", "b0", 0, 0)]


def extra_params(key):
    from oracle import make_golden_r2 as mg2

    return mg2.CONSTANT_BRIGHT if key.startswith("constant") else h.CONFTEST_BRIGHT


@pytest.mark.parametrize("extrap", ["linear", "clamp"])
@pytest.mark.parametrize("key,det,kind,seed", EXTRA_CASES)
def test_extra_source_stream_replay(extrap, key, det, kind, seed):
    g = load(f"spectral_extra_{extrap}.npz")
    p = extra_params(key)
    rsp = h.response(det)
    ea = rsp.effective_area_packed
    s = h.oracle_source(rsp, kind=kind, interp_extrap=extrap, **p)
    fm = float(g[f"fmax_{key}"])
    assert abs(orc.fmax(s, ea, 0.0, p["duration"]) - fm) <= 1e-12 * fm
    rng = orc.Rng(seed=seed)
    t, nc = orc.sample_events(s, ea, 0.0, p["duration"], fm, rng)
    assert nc == int(g[f"ncand_{key}"]) and len(t) == len(g[f"events_{key}"])
    np.testing.assert_allclose(t, g[f"events_{key}"], rtol=1e-12)
    n_e = len(g[f"energy_{key}"])
    e, tr = orc.sample_energy(s, ea, g[f"events_{key}"][:n_e], rng)
    np.testing.assert_allclose(e, g[f"energy_{key}"], rtol=1e-12)
    assert np.array_equal(tr, g[f"trials_{key}"])
    pha = orc.digitize(g[f"energy_{key}"], rsp.energy_edges, rsp.cumulative_matrix, rng)
    assert np.array_equal(pha, g[f"pha_{key}"].astype("f8"))


def test_extra_bright_long_window():
    from oracle import make_golden_r2 as mg2

    g = load("spectral_extra_linear.npz")
    rsp = h.response("n0")
    s = h.oracle_source(rsp, **h.BRIGHT_LONG)
    t, nc = orc.sample_events(s, rsp.effective_area_packed, mg2.C2_WINDOW[0], mg2.C2_WINDOW[1],
                              float(g["fmax_long_window_n0_s7"]), orc.Rng(seed=7))
    assert nc == int(g["ncand_long_window_n0_s7"])
    np.testing.assert_allclose(t, g["events_long_window_n0_s7"], rtol=1e-12)
