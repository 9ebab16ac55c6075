"""GPU parity against the reference-minted golden vectors DIRECTLY (tests/golden/*.npz were produced by running
the real numba kernels / the real GBMLightCurve.process() of the reference: oracle/make_golden.py,
oracle/make_golden_r2.py).  The CUDA path replays the reference's MT19937 uniform stream (injected mode) and must
reproduce the stored outputs: decisions, trial counts, channels and dead-time survivors bit-exact, times and
energies to 1e-10 relative (north-star bar: 1e-6).

The C oracle appears here only where the reference's serial stream positions are not stored in the golden file
(the per-photon trial counts of the whole-unit replay) and as the second checker of the large C2 window.
"""
import os

import numpy as np
import pytest

from tests import helpers as h
from oracle import make_golden as mg
from oracle import make_golden_r2 as mg2
from oracle import oracle as orc
from cosmogrb_b200.engine import Injected

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TIGHT = 1e-10


def load(name):
    with np.load(os.path.join(GOLD, name)) as d:
        return {k: d[k] for k in d.files}


def _unit_batch(engine, p, det, kind, extrap, fmax=None, window=None, bkg=(-100.0, 300.0, 500.0)):
    rsp = h.response(det)
    u = h.unit(kind=kind, bkg=bkg, **p)
    if window is not None:
        u["src_tstart"], u["src_tstop"] = window
    if fmax is not None:
        u["fmax"] = fmax
    b = engine.new_batch(u, [rsp.device_table(h.bkg_weights(det), extrap)], interp_extrap=extrap,
                         compute_fmax=fmax is None)
    return rsp, b


# golden key, file stem, detector, kind, parameters, MT19937 seed
CHAINS = [(f"readme_s{s}", "spectral", "n0", 0, h.README_DEMO, s) for s in (0, 1, 2)] + \
         [(f"bright_s{s}", "spectral", "n0", 0, h.CONFTEST_BRIGHT, s) for s in (0, 1, 2)] + \
         [("constant_s4", "spectral", "n0", 1, mg.CONSTANT, 4)] + \
         [(f"constant_{d}_s{s}", "spectral_extra", d, 1, mg2.CONSTANT_BRIGHT, s) for d in ("n0", "b0") for s in (4, 5)] + \
         [(f"bright_{d}_s0", "spectral_extra", d, 0, h.CONFTEST_BRIGHT, 0) for d in ("n5", "b0")]


@pytest.mark.parametrize("extrap", ["linear", "clamp"])
@pytest.mark.parametrize("key,stem,det,kind,p,seed", CHAINS, ids=[c[0] for c in CHAINS])
def test_source_chain_replays_the_reference(engine, extrap, key, stem, det, kind, p, seed):
    """sample_events -> sample_energy -> _digitize on ONE numba stream, as LightCurve._sample_source runs them
    (lightcurve.py:74-100; cpl_functions.py:134-286, cpl_constant_functions.py:66-177, response.py:10-25)."""
    g = load(f"{stem}_{extrap}.npz")
    fkey = f"fmax_{key}" if f"fmax_{key}" in g else f"fmax_{key.split('_s')[0]}"
    rsp, b = _unit_batch(engine, p, det, kind, extrap)
    fm = float(g[fkey])
    assert b.units["fmax"][0] == pytest.approx(fm, rel=1e-11)
    want_t = g[f"events_{key}"]
    want_e = g[f"energy_{key}"]
    want_tr = g[f"trials_{key}"].astype(np.int64)
    cap = int(b.units["src_cap"][0])
    n_e = len(want_e)
    u = np.random.RandomState(seed).random_sample(2 * cap + 2 + 2 * int(want_tr.sum()) + n_e + 8)
    engine.sample_events(b, Injected(thin=[u]))
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0
    assert c["n_src"][0] == len(want_t)                       # identical accept / reject decisions
    got_t = b.view("times_src", 0, "src")
    np.testing.assert_allclose(got_t, want_t, rtol=TIGHT)
    if f"ncand_{key}" in g:
        assert c["n_candidates"][0] == int(g[f"ncand_{key}"])
    # the thinning loop consumed u1, u2 per candidate and one last u1
    pos = 2 * int(c["n_candidates"][0]) + 1
    off = pos + np.concatenate(([0], np.cumsum(2 * want_tr)[:-1]))
    engine.set_counts(b, n_src=n_e)
    engine.sample_energy(b, Injected(energy=u, energy_off=[off]), want_trials=True)
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0
    assert np.array_equal(b.t["trials"][:n_e].cpu().numpy(), want_tr)
    np.testing.assert_allclose(b.t["energy_src"][:n_e].cpu().numpy(), want_e, rtol=TIGHT)
    if f"pha_{key}" in g:
        pos_d = pos + 2 * int(want_tr.sum())
        engine.digitize_batch(b, Injected(digitize=[u[pos_d:pos_d + n_e]]))
        assert np.array_equal(b.t["pha_src"][:n_e].cpu().numpy().astype("f8"), g[f"pha_{key}"].astype("f8"))


def test_whole_unit_replays_gbm_lightcurve_process(engine):
    """tests/golden/pipeline.npz is the real GBMLightCurve.process() (lightcurve.py:167-200 with the real Source,
    CPLSourceFunction, Background, Response): the device runs the same unit end to end on the same two uniform
    streams -- source times, energies, channels, background times and channels, combine, dead time -- and all six
    event lists of LightCurveStorage must equal the stored ones."""
    g = load("pipeline.npz")
    p = h.CONFTEST_BRIGHT
    rate = float(g["bkg_rate"])
    rsp, b = _unit_batch(engine, p, "n0", 0, "linear", bkg=(-20.0, 30.0, rate))
    assert b.units["fmax"][0] == pytest.approx(float(g["fmax"]), rel=1e-11)
    n_src, n_bkg = len(g["times_source"]), len(g["times_background"])
    # per-photon trial counts are not in the file: the oracle supplies the serial stream positions only
    s = h.oracle_source(rsp, **p)
    ea = rsp.effective_area_packed
    rng = orc.Rng(seed=int(g["seeds"][0]))
    ts, nc = orc.sample_events(s, ea, 0.0, p["duration"], float(g["fmax"]), rng)
    assert rng.pos == 2 * nc + 1
    _, tr = orc.sample_energy(s, ea, ts, rng)
    tr = tr.astype(np.int64)
    pos_e = 2 * nc + 1
    pos_d = pos_e + 2 * int(tr.sum())
    pos_b = pos_d + n_src
    u = np.random.RandomState(int(g["seeds"][0])).random_sample(pos_b + 2 * int(b.units["bkg_cap"][0]) + 8)
    uc = np.random.RandomState(int(g["seeds"][1])).random_sample(int(b.units["bkg_cap"][0]) + 8)

    engine.sample_events(b, Injected(thin=[u]))
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0 and c["n_src"][0] == n_src and c["n_candidates"][0] == nc
    off = pos_e + np.concatenate(([0], np.cumsum(2 * tr)[:-1]))
    engine.sample_energy(b, Injected(energy=u, energy_off=[off]), want_trials=True)
    assert np.array_equal(b.t["trials"][:n_src].cpu().numpy(), tr)
    engine.digitize_batch(b, Injected(digitize=[u[pos_d:pos_d + n_src]]))
    engine.sample_background(b, Injected(bkg_time=[u[pos_b:]], bkg_chan=[uc]))
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0 and c["n_bkg"][0] == n_bkg
    np.testing.assert_allclose(b.view("times_src", 0, "src"), g["times_source"], rtol=TIGHT)
    assert np.array_equal(b.view("pha_src", 0, "src"), g["pha_source"])
    np.testing.assert_allclose(b.view("times_bkg", 0, "bkg"), g["times_background"], rtol=TIGHT, atol=1e-9)
    assert np.array_equal(b.view("pha_bkg", 0, "bkg"), g["pha_background"])
    # production (one-pass) and diagnostics (combine, then dead time) variants of the last stage
    for fused in (True, False):
        if fused:
            engine.merge_dead_time(b)
        else:
            engine.combine(b)
            engine.dead_time(b)
        c = engine.fetch_counts(b)
        assert c["status"][0] == 0
        assert c["n_total"][0] == n_src + n_bkg
        assert c["n_kept"][0] == len(g["times"]), fused
        np.testing.assert_allclose(b.view("times_out", 0, "kept"), g["times"], rtol=TIGHT, atol=1e-9)
        assert np.array_equal(b.view("pha_out", 0, "kept"), g["pha"]), fused


@pytest.mark.parametrize("extrap", ["linear", "clamp"])
def test_bright_long_window_replays_the_reference(engine, extrap):
    """0.4 s of the BENCHMARKED burst (BASELINE.json configs[1]) near its maximum, thinned by the real reference
    with the whole burst's fmax: 29,899 candidates, decisions bit-exact (squeeze table active on the device)."""
    g = load(f"spectral_extra_{extrap}.npz")
    fm = float(g["fmax_long_window_n0_s7"])
    rsp, b = _unit_batch(engine, h.BRIGHT_LONG, "n0", 0, extrap, fmax=fm, window=mg2.C2_WINDOW)
    assert b.units["tab_n"][0] > 0
    u = np.random.RandomState(7).random_sample(2 * int(b.units["src_cap"][0]) + 2)
    engine.sample_events(b, Injected(thin=[u]))
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0
    assert c["n_candidates"][0] == int(g["ncand_long_window_n0_s7"])
    want = g["events_long_window_n0_s7"]
    assert c["n_src"][0] == len(want)
    np.testing.assert_allclose(b.view("times_src", 0, "src"), want, rtol=TIGHT)
    assert c["n_exact"][0] < c["n_candidates"][0]


def test_bright_long_million_candidate_window_against_the_oracle(engine):
    """>= 1e6 candidates of the benchmarked burst (its rise, 2..17 s: p = lambda/fmax from ~0.1 to ~1) under injected
    uniforms: every accept/reject decision equals the oracle's."""
    p = h.BRIGHT_LONG
    rsp = h.response("n0")
    s = h.oracle_source(rsp, **p)
    ea = rsp.effective_area_packed
    fm = orc.fmax(s, ea, 0.0, p["duration"])
    win = (2.0, 17.0)
    rsp, b = _unit_batch(engine, p, "n0", 0, "linear", fmax=fm, window=win)
    cap = int(b.units["src_cap"][0])
    assert cap > 1_000_000
    u = np.random.RandomState(77).random_sample(2 * cap + 2)
    want_t, want_nc, want_ct, want_acc = orc.sample_events(s, ea, win[0], win[1], fm, orc.Rng(u=u), want_candidates=True)
    assert want_nc > 1_000_000
    engine.sample_events(b, Injected(thin=[u]), want_candidates=True)
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0
    assert c["n_candidates"][0] == want_nc
    got_acc = b.t["cand_accept"][:want_nc].cpu().numpy().astype(bool)
    mism = np.nonzero(got_acc != want_acc)[0]
    assert len(mism) == 0, f"{len(mism)} accept/reject mismatches, first at candidate {mism[:5]}"
    assert 0.2 < want_acc.mean() < 0.95
    np.testing.assert_allclose(b.t["cand_times"][:want_nc].cpu().numpy(), want_ct, rtol=TIGHT)
    np.testing.assert_allclose(b.view("times_src", 0, "src"), want_t, rtol=TIGHT)
    # the squeeze table decided nearly all of them; the rest took the exact 75-point lambda(t)
    assert c["n_exact"][0] < 0.05 * want_nc
