"""GPU parity: every CUDA stage against the CPU oracle on the same injected uniform streams.

Bars (BASELINE.json north_star): accept/reject decisions, channel indices and dead-time survivor
sets bit-exact; event times and energies within 1e-6 relative (fp64; measured agreement is ~1e-13).
All calls go through the C ABI (cosmogrb_b200._capi -> libcosmogrb_b200.so).
"""
import numpy as np
import pytest

from tests import helpers as h
from oracle import oracle as orc
from cosmogrb_b200 import _capi
from cosmogrb_b200.engine import Injected, Philox

pytestmark = pytest.mark.gpu

RTOL = 1e-6          # north-star tolerance for fp64 times / energies
TIGHT = 1e-10        # what we actually expect


def _batch(engine, params, det="n0", kind=0, interp_extrap="linear", **kw):
    rsp = h.response(det)
    u = h.unit(kind=kind, **params, **kw)
    b = engine.new_batch(u, [rsp.device_table(h.bkg_weights(det), interp_extrap)], interp_extrap=interp_extrap)
    return rsp, b


@pytest.mark.parametrize("params,kind", [(h.README_DEMO, 0), (h.CONFTEST_BRIGHT, 0), (h.BRIGHT_LONG, 0),
                                         (dict(z=1.0, peak_flux=5e-9, alpha=-0.66, ep=500.0, duration=2.0), 1)])
@pytest.mark.parametrize("extrap", ["linear", "clamp"])
def test_lambda_and_fmax(engine, params, kind, extrap):
    rsp = h.response("n0")
    s = h.oracle_source(rsp, kind=kind, interp_extrap=extrap, **params)
    u = h.unit(kind=kind, **params)
    tab = rsp.device_table(h.bkg_weights("n0"), extrap)
    times = np.linspace(0.0, params["duration"], 257)
    want = orc.energy_integrated_evolution(s, rsp.effective_area_packed, times)
    got = engine.energy_integrated_evolution(u, [tab], times, interp_extrap=extrap)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-300)
    b = engine.new_batch(u, [tab], interp_extrap=extrap)
    want_fmax = orc.fmax(s, rsp.effective_area_packed, 0.0, params["duration"])
    assert b.units["fmax"][0] == pytest.approx(want_fmax, rel=1e-11)


def test_evolution(engine):
    rsp = h.response("n0")
    s = h.oracle_source(rsp, **h.README_DEMO)
    u = h.unit(**h.README_DEMO)
    e = np.logspace(1, 4, 40)
    t = np.linspace(0, 20, 11)
    want = orc.cpl_evolution(s, e, t)
    got = engine.evolution(u, e, t)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-300)


@pytest.mark.parametrize("params", [h.README_DEMO, h.CONFTEST_BRIGHT])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sample_events_injected(engine, params, seed):
    rsp, b = _batch(engine, params)
    s = h.oracle_source(rsp, **params)
    fmax = float(b.units["fmax"][0])
    cap = int(b.units["src_cap"][0])
    u = np.random.RandomState(seed).random_sample(2 * cap + 2)
    want_t, want_nc, want_ct, want_acc = orc.sample_events(s, rsp.effective_area_packed, 0.0, params["duration"],
                                                            fmax, orc.Rng(u=u), want_candidates=True)
    engine.sample_events(b, Injected(thin=[u]), want_candidates=True)
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0
    assert c["n_candidates"][0] == want_nc
    got_acc = b.t["cand_accept"][:want_nc].cpu().numpy().astype(bool)
    got_ct = b.t["cand_times"][:want_nc].cpu().numpy()
    np.testing.assert_allclose(got_ct, want_ct, rtol=TIGHT)
    mism = np.nonzero(got_acc != want_acc)[0]
    assert len(mism) == 0, f"{len(mism)} accept/reject mismatches, first at candidate {mism[:5]}"
    assert c["n_src"][0] == len(want_t)
    got_t = b.view("times_src", 0, "src")
    assert got_t[0] == 0.0                      # the spurious tstart event
    np.testing.assert_allclose(got_t, want_t, rtol=TIGHT)
    assert np.all(np.diff(got_t) >= 0)


@pytest.mark.parametrize("det,extrap,kind", [("n5", "linear", 0), ("n5", "clamp", 0), ("b0", "linear", 0),
                                             ("b0", "clamp", 0), ("n0", "clamp", 0), ("n3", "linear", 1),
                                             ("b1", "clamp", 1)])
def test_sample_events_injected_detectors_and_variants(engine, det, extrap, kind):
    """the same decision-level check on other NaI detectors, the BGO detectors, the `clamp` interpolation variant
    and the constant CPL (cpl_constant_functions.py:66-105)"""
    params = h.CONFTEST_BRIGHT if kind == 0 else dict(z=1.0, peak_flux=4e-6, alpha=-0.9, ep=300.0, duration=3.0)
    rsp, b = _batch(engine, params, det=det, kind=kind, interp_extrap=extrap)
    s = h.oracle_source(rsp, kind=kind, interp_extrap=extrap, **params)
    fmax = float(b.units["fmax"][0])
    assert fmax == pytest.approx(orc.fmax(s, rsp.effective_area_packed, 0.0, params["duration"]), rel=1e-11)
    cap = int(b.units["src_cap"][0])
    u = np.random.RandomState(13).random_sample(2 * cap + 2)
    want_t, want_nc, want_ct, want_acc = orc.sample_events(s, rsp.effective_area_packed, 0.0, params["duration"],
                                                            fmax, orc.Rng(u=u), want_candidates=True)
    engine.sample_events(b, Injected(thin=[u]), want_candidates=True)
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0 and c["n_candidates"][0] == want_nc and want_nc > 100
    assert np.array_equal(b.t["cand_accept"][:want_nc].cpu().numpy().astype(bool), want_acc)
    np.testing.assert_allclose(b.t["cand_times"][:want_nc].cpu().numpy(), want_ct, rtol=TIGHT)
    np.testing.assert_allclose(b.view("times_src", 0, "src"), want_t, rtol=TIGHT)


@pytest.mark.parametrize("det,extrap,kind", [("n0", "linear", 0), ("n0", "clamp", 0), ("n5", "linear", 0),
                                             ("b0", "linear", 0), ("b0", "clamp", 0), ("n0", "linear", 1),
                                             ("b0", "clamp", 1)])
def test_sample_energy_injected_serial_replay(engine, det, extrap, kind):
    """Strict serial replay: per-photon offsets are the positions the reference's serial stream
    reaches (cumulative 2 x trials), so the GPU consumes exactly the reference's uniforms."""
    params = h.CONFTEST_BRIGHT if kind == 0 else dict(z=1.0, peak_flux=4e-6, alpha=-0.9, ep=300.0, duration=3.0)
    rsp, b = _batch(engine, params, det=det, kind=kind, interp_extrap=extrap)
    s = h.oracle_source(rsp, kind=kind, interp_extrap=extrap, **params)
    rs = np.random.RandomState(5)
    times = np.sort(rs.uniform(0.05, 2.0, 600))
    u = rs.random_sample(600 * 2 * 1500)
    want_e, want_tr = orc.sample_energy(s, rsp.effective_area_packed, times, orc.Rng(u=u))
    assert 2 * int(want_tr.sum()) < len(u)
    off = np.concatenate(([0], np.cumsum(2 * want_tr.astype(np.int64))[:-1]))
    n = len(times)
    assert n <= b.units["src_cap"][0]
    engine.set_counts(b, n_src=n)
    import torch
    ts = np.zeros(b.n_src_slots)
    ts[:n] = times
    b.t["times_src"] = engine.to_device(ts)
    engine.sample_energy(b, Injected(energy=u, energy_off=[off]), want_trials=True)
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0
    got_e = b.t["energy_src"][:n].cpu().numpy()
    got_tr = b.t["trials"][:n].cpu().numpy()
    assert np.array_equal(got_tr, want_tr)
    np.testing.assert_allclose(got_e, want_e, rtol=TIGHT)
    assert c["n_trials"][0] == want_tr.sum()


@pytest.mark.parametrize("det", ["n0", "n5", "b0"])
def test_digitize_bit_exact(engine, det):
    rsp = h.response(det)
    rs = np.random.RandomState(11)
    n = 200_000
    e = np.exp(rs.uniform(np.log(rsp.emin), np.log(rsp.emax), n))
    edges = rsp.energy_edges
    e[:141] = edges                     # exact-edge energies, incl. E == emin (wraps to the last row)
    e[141:300] = rs.uniform(edges[0], edges[3], 159)   # zero-probability rows
    u = rs.random_sample(n)
    u[300:310] = 0.0
    want = orc.digitize(e, edges, rsp.cumulative_matrix, orc.Rng(u=u))
    got = engine.digitize(rsp, e, rng=Injected(digitize=[u]))
    assert got.dtype == np.float64
    assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", [0, 3])
def test_background_injected(engine, seed):
    rsp, b = _batch(engine, h.README_DEMO, flags=_capi.UNIT_NO_SOURCE)
    cap = int(b.units["bkg_cap"][0])
    rs = np.random.RandomState(seed)
    u_t = rs.random_sample(2 * cap + 2)
    u_c = rs.random_sample(cap + 2)
    want_t = orc.background_times(-100.0, 300.0, 500.0, orc.Rng(u=u_t))
    want_c = orc.background_channels(h.bkg_weights("n0"), len(want_t), orc.Rng(u=u_c))
    engine.sample_background(b, Injected(bkg_time=[u_t], bkg_chan=[u_c]))
    c = engine.fetch_counts(b)
    assert c["status"][0] == 0
    assert c["n_bkg"][0] == len(want_t)
    got_t = b.view("times_bkg", 0, "bkg")
    got_c = b.view("pha_bkg", 0, "bkg")
    np.testing.assert_allclose(got_t, want_t, rtol=TIGHT, atol=1e-9)
    assert np.all(np.diff(got_t) >= 0)
    assert np.array_equal(got_c.astype(np.int64), want_c)


def test_dead_time_crafted(engine):
    # the 10-event probe case of SURVEY.md §8c (v)
    t = np.array([0, 1, 3, 4, 10, 15, 20.5, 20.7, 31.2, 32]) * 1e-6
    pha = np.array([5, 6, 7, 8, 127, 9, 10, 11, 127, 12])
    want = np.array([1, 0, 1, 1, 1, 0, 0, 1, 1, 0], dtype=bool)
    assert np.array_equal(orc.gbm_dead_time(t, pha), want)
    assert np.array_equal(engine.gbm_dead_time(t, pha), want)
    # first event in the overflow channel
    pha2 = pha.copy()
    pha2[0] = 127
    assert np.array_equal(engine.gbm_dead_time(t, pha2), orc.gbm_dead_time(t, pha2))
    # single event / two events
    assert np.array_equal(engine.gbm_dead_time(t[:1], pha[:1]), [True])
    assert np.array_equal(engine.gbm_dead_time(t[:2], pha[:2]), [True, False])


@pytest.mark.parametrize("rate,frac127", [(500.0, 0.05), (2e5, 0.05), (2e6, 0.3), (5e6, 1.0)])
def test_dead_time_random(engine, rate, frac127):
    rs = np.random.RandomState(7)
    n = 1_000_000
    t = np.cumsum(rs.exponential(1.0 / rate, n))
    pha = rs.randint(0, 127, n)
    pha[rs.random_sample(n) < frac127] = 127
    want = orc.gbm_dead_time(t, pha)
    got = engine.gbm_dead_time(t, pha)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("rate,frac127", [(500.0, 0.05), (2e5, 0.05), (2e6, 0.3), (5e6, 1.0)])
def test_fused_merge_dead_time(engine, rate, frac127):
    """The one-pass combine + dead-time kernel == the oracle's combine then _gbm_dead_time, bit for bit,
    from sparse lists to rates where the dependence chains outrun the tile halo (slow path)."""
    rs = np.random.RandomState(11)
    n = 30_000 if rate > 1e6 else 1_000_000       # the slow path is O(chain length x log n) per event
    ts = np.cumsum(rs.exponential(1.0 / rate, n))
    tb = np.sort(rs.uniform(0.0, ts[-1], n // 3))
    tb[5:8] = ts[100]                         # ties: background first
    tb.sort()
    ps = rs.randint(0, 127, n)
    ps[rs.random_sample(n) < frac127] = 127
    pb = rs.randint(0, 128, len(tb))
    want_t, want_p = orc.combine(tb, pb, ts, ps)
    sel = orc.gbm_dead_time(want_t, want_p)
    for fused in (True, False):
        got_t, got_p = engine.combine_dead_time(tb, pb, ts, ps, fused=fused)
        assert np.array_equal(got_t, want_t[sel]), fused
        assert np.array_equal(got_p.astype("f8"), want_p[sel]), fused
    # degenerate shapes: one empty list, single events
    for a, b_ in ((0, 5), (5, 0), (1, 1)):
        got_t, got_p = engine.combine_dead_time(tb[:a], pb[:a], ts[:b_], ps[:b_])
        wt, wp = orc.combine(tb[:a], pb[:a], ts[:b_], ps[:b_])
        s2 = orc.gbm_dead_time(wt, wp)
        assert np.array_equal(got_t, wt[s2]) and np.array_equal(got_p.astype("f8"), wp[s2])


def test_dead_time_physical_crafted(engine):
    """CGRB_DEADTIME_PHYSICAL (non-paralyzable) against its serial definition."""
    t = np.array([0, 1, 3, 4, 10, 15, 20.5, 20.7, 31.2, 32]) * 1e-6
    pha = np.array([5, 6, 7, 8, 127, 9, 10, 11, 127, 12])
    want = np.array([1, 0, 1, 0, 1, 0, 1, 0, 1, 0], dtype=bool)
    assert np.array_equal(engine.gbm_dead_time(t, pha, mode="physical"), want)
    assert np.array_equal(engine.gbm_dead_time(t[:1], pha[:1], mode="physical"), [True])
    assert np.array_equal(engine.gbm_dead_time(t[:2], pha[:2], mode="physical"), [True, False])
    # the default stays the reference's literal rule
    assert np.array_equal(engine.gbm_dead_time(t, pha), orc.gbm_dead_time(t, pha))
    # other window lengths, incl. tau > tau_overflow and zero windows
    old = (engine.deadtime_mode, engine.deadtime_tau, engine.deadtime_tau_overflow)
    try:
        for tau, tovf in ((5e-6, 1e-6), (0.0, 0.0), (2.6e-6, 10.6e-6)):
            engine.set_deadtime("physical", tau, tovf)
            assert np.array_equal(engine.gbm_dead_time(t, pha), orc.gbm_dead_time_physical(t, pha, tau, tovf))
    finally:
        engine.set_deadtime(*old)
    with pytest.raises(ValueError):
        engine.gbm_dead_time(t, pha, mode="paralyzable")


@pytest.mark.parametrize("rate,frac127", [(500.0, 0.05), (2e4, 0.05), (2e5, 0.1), (5e5, 0.3)])
def test_dead_time_physical_random(engine, rate, frac127):
    """1e6 random events from GBM-like rates up to 5e5 /s (runs of ~150 events between synchronisation
    points), stand-alone and fused with the merge, bit-exact against the serial definition."""
    rs = np.random.RandomState(17)
    n = 1_000_000
    t = np.cumsum(rs.exponential(1.0 / rate, n))
    pha = rs.randint(0, 127, n)
    pha[rs.random_sample(n) < frac127] = 127
    want = orc.gbm_dead_time_physical(t, pha)
    assert np.array_equal(engine.gbm_dead_time(t, pha, mode="physical"), want)
    # fused: split the list into two interleaved "components" and let the kernel merge them back
    is_src = rs.random_sample(n) < 0.7
    ts, ps, tb, pb = t[is_src], pha[is_src], t[~is_src], pha[~is_src]
    wt, wp = orc.combine(tb, pb, ts, ps)
    sel = orc.gbm_dead_time_physical(wt, wp)
    for fused in (True, False):
        got_t, got_p = engine.combine_dead_time(tb, pb, ts, ps, fused=fused, mode="physical")
        assert np.array_equal(got_t, wt[sel]), fused
        assert np.array_equal(got_p.astype("f8"), wp[sel]), fused
    # fraction lost ~ 1 - 1/(1 + rate * mean window) for a non-paralyzable counter
    mean_tau = (1 - frac127) * 2.6e-6 + frac127 * 10e-6
    assert abs(want.mean() - 1.0 / (1.0 + rate * mean_tau)) < 0.01


def test_dead_time_physical_runaway_chains_take_the_serial_fix_up(engine):
    """rate x window >> 1: no synchronisation point inside the look-back range -> the unit is flagged and redone by
    the serial fix-up kernel of the same call (exact, status clear), not walked in O(n^2) and not refused"""
    n = 60000
    rs = np.random.RandomState(5)
    t = np.cumsum(rs.exponential(1e-7, n))              # 1e7 events/s: every gap far below the 10 us window
    pha = rs.randint(0, 128, n)
    want = orc.gbm_dead_time_physical(t, pha)
    assert 0.005 < want.mean() < 0.05
    assert np.array_equal(engine.gbm_dead_time(t, pha, mode="physical"), want)
    is_src = rs.random_sample(n) < 0.5
    wt, wp = orc.combine(t[~is_src], pha[~is_src], t[is_src], pha[is_src])
    sel = orc.gbm_dead_time_physical(wt, wp)
    for fused in (True, False):
        got_t, got_p = engine.combine_dead_time(t[~is_src], pha[~is_src], t[is_src], pha[is_src], fused=fused,
                                                mode="physical")
        assert np.array_equal(got_t, wt[sel]) and np.array_equal(got_p.astype("f8"), wp[sel]), fused
    # a regular arithmetic train of overflow events (every 11th kept), and the literal rule on the same input
    t2 = np.arange(20000) * 1e-6
    p2 = np.full(20000, 127)
    assert np.array_equal(engine.gbm_dead_time(t2, p2, mode="physical"), orc.gbm_dead_time_physical(t2, p2))
    assert np.array_equal(engine.gbm_dead_time(t2, p2), orc.gbm_dead_time(t2, p2))


def test_combine_and_dead_time_pipeline(engine):
    """Full unit under Philox; the oracle re-does combine + dead time from the GPU's component lists."""
    params = h.CONFTEST_BRIGHT
    rsp, b = _batch(engine, params)
    engine.process(b, Philox(1234), diagnostics=True)
    c = engine.fetch_counts(b)
    engine.check_status(b)
    t_src, p_src = b.view("times_src", 0, "src"), b.view("pha_src", 0, "src")
    t_bkg, p_bkg = b.view("times_bkg", 0, "bkg"), b.view("pha_bkg", 0, "bkg")
    assert c["n_total"][0] == len(t_src) + len(t_bkg)
    want_t, want_p = orc.combine(t_bkg, p_bkg, t_src, p_src)
    got_t, got_p = b.view("times_tot", 0, "tot"), b.view("pha_tot", 0, "tot")
    assert np.array_equal(got_t, want_t)
    assert np.array_equal(got_p.astype("f8"), want_p)
    sel = orc.gbm_dead_time(want_t, want_p)
    got_sel = b.view("selection", 0, "tot").astype(bool)
    assert np.array_equal(got_sel, sel)
    assert c["n_kept"][0] == sel.sum()
    assert np.array_equal(b.view("times_out", 0, "kept"), want_t[sel])
    assert np.array_equal(b.view("pha_out", 0, "kept").astype("f8"), want_p[sel])
    # energies inside the response range, channels valid
    e = b.view("energy_src", 0, "src")
    assert np.all((e >= rsp.emin) & (e <= rsp.emax))
    # the fold fused into the energy kernel == the standalone digitize kernel on the same energies
    # (same channel uniforms: they ride in the Philox call of the accepting trial, replayed from the trial counts),
    # which in turn is the oracle's _digitize given the uniforms (test above)
    fused = b.view("pha_src", 0, "src").copy()
    engine.digitize_batch(b, Philox(1234), trials=b.t["trials"])
    assert np.array_equal(fused, b.view("pha_src", 0, "src"))
    # re-derive the source channels with the oracle from the same Philox uniforms is not possible
    # (device counters), so check determinism instead: same seed -> identical output
    rsp2, b2 = _batch(engine, params)
    engine.process(b2, Philox(1234))          # the production path: fused fold, fused combine + dead time
    c2 = engine.fetch_counts(b2)
    assert c2["n_total"][0] == c["n_total"][0] and c2["n_kept"][0] == c["n_kept"][0]
    assert np.array_equal(b2.view("times_out", 0, "kept"), want_t[sel])
    assert np.array_equal(b2.view("pha_out", 0, "kept").astype("f8"), want_p[sel])


def test_multi_unit_batch_independent_of_batching(engine):
    """14 detectors in one batch == each detector alone (Philox counters depend only on the unit)."""
    params = h.CONFTEST_BRIGHT
    dets = h.GBM_DETECTORS
    tabs = [h.response(d).device_table(h.bkg_weights(d), "linear") for d in dets]
    units = np.concatenate([h.unit(det_index=i, stream_id=16 * 7 + i, **params) for i in range(len(dets))])
    b = engine.new_batch(units, tabs)
    engine.process(b, Philox(99))
    c = engine.fetch_counts(b)
    engine.check_status(b)
    for i in (0, 5, 13):
        bi = engine.new_batch(units[i:i + 1].copy(), tabs)
        engine.process(bi, Philox(99))
        ci = engine.fetch_counts(bi)
        for k in ("n_src", "n_bkg", "n_total", "n_kept", "n_candidates", "n_trials"):
            assert ci[k][0] == c[k][i], k
        assert np.array_equal(bi.view("times_out", 0, "kept"), b.view("times_out", i, "kept"))
        assert np.array_equal(bi.view("pha_out", 0, "kept"), b.view("pha_out", i, "kept"))


@pytest.mark.parametrize("params", [h.README_DEMO, h.BRIGHT_LONG])
def test_squeeze_changes_no_decision(params):
    """Full-size property test: the squeeze table only short-cuts the acceptance test; the accepted
    set with and without it is identical (BRIGHT_LONG: ~2e7 candidates per detector)."""
    from cosmogrb_b200.engine import Engine

    out = {}
    for squeeze in (True, False):
        eng = Engine(squeeze=squeeze, thinning="literal")     # the reference's single-fmax proposal
        rsp = h.response("n0")
        b = eng.new_batch(h.unit(**params), [rsp.device_table(h.bkg_weights("n0"), "linear")])
        assert (b.units["tab_n"][0] > 0) == squeeze
        eng.sample_events(b, Philox(31337))
        c = eng.fetch_counts(b)
        assert c["status"][0] == 0
        out[squeeze] = (int(c["n_candidates"][0]), int(c["n_exact"][0]), b.view("times_src", 0, "src"))
    assert out[True][0] == out[False][0]
    assert out[False][1] == out[False][0]                 # without the table every candidate is exact
    assert out[True][1] < 0.05 * out[True][0], f"squeeze left {out[True][1]} of {out[True][0]} exact"
    assert np.array_equal(out[True][2], out[False][2])


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011) on uint32 arrays: csrc/cgrb_device.cuh philox4x32_10"""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m32, p1 & m32, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m32, p0 & m32
        k0, k1 = (k0 + np.uint64(W0)) & m32, (k1 + np.uint64(W1)) & m32
    return c0, c1, c2, c3


def _u53(a, b):
    """cgrb_device.cuh u53: the top 52 of 64 random bits as the mantissa of a double in [1, 2), minus one"""
    bits = (np.uint64(0x3FF0000000000000) | ((a >> np.uint64(12)) << np.uint64(32)) | ((a << np.uint64(20)) & np.uint64(0xFFFFFFFF))
            | (b >> np.uint64(12)))
    return bits.view("f8") - 1.0


def test_background_under_philox_equals_a_host_replay_of_the_counters(engine):
    """The production background (one pass, fast logarithm, guided channel draw) against a NumPy replay of its
    Philox counters: ONE call per candidate pair m, counter (m, 0), carries the pair's two gap uniforms (40 bits each,
    (k + 1/2) 2^-40) and its two channel draws (24 bits each) -- cgrb_device.cuh bkg_draw; times = tstart + running
    sum of -log(u) / rate (background.py:13-44), channels = searchsorted(cdf, u, 'right') (numpy's legacy choice,
    :93).  Times agree to 1e-13 relative (the device's logarithm and summation order differ from NumPy's in the
    last bits), channels and the count exactly."""
    seed, sid, rate, t0, t1 = 0x1234ABCD5678, 37, 500.0, -100.0, 300.0
    rsp = h.response("n0")
    w = h.bkg_weights("n0")
    u = h.unit(stream_id=sid, bkg=(t0, t1, rate), flags=_capi.UNIT_NO_SOURCE, **h.README_DEMO)
    b = engine.new_batch(u, [rsp.device_table(w, "linear")], compute_fmax=False)
    engine.sample_background(b, Philox(seed))
    c = engine.fetch_counts(b)
    engine.check_status(b)
    got_t, got_p = b.view("times_bkg", 0, "bkg"), b.view("pha_bkg", 0, "bkg")
    n_pairs = (int(b.units["bkg_cap"][0]) + 1) // 2
    m = np.arange(n_pairs, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    x, y, z, ww = _philox4x32_10(m, 0, sid, 4, k0, k1)              # ST_BKG = 4
    u64 = np.uint64
    ga = ((x << u64(8)) | (y >> u64(24))).astype("f8")              # 40-bit integers: exact in a double
    gb = ((z << u64(8)) | ((y >> u64(16)) & u64(0xFF))).astype("f8")
    ug = np.stack(((ga + 0.5) * 2.0 ** -40, (gb + 0.5) * 2.0 ** -40), axis=1).reshape(-1)
    ca, cb = ww >> u64(8), ((ww & u64(0xFF)) << u64(16)) | (y & u64(0xFFFF))
    uc = np.stack((ca.astype("f8") * 2.0 ** -24, cb.astype("f8") * 2.0 ** -24), axis=1).reshape(-1)
    assert ug.min() > 0.0 and ug.max() < 1.0
    times = t0 + np.cumsum(-np.log(ug) / rate)
    n = int(np.searchsorted(times, t1, "right"))
    assert abs(int(c["n_bkg"][0]) - (n + 1)) <= 1                  # a time within 1e-13 of tstop may fall either side
    n = int(c["n_bkg"][0]) - 1
    assert got_t[0] == t0 and np.all(np.diff(got_t) >= 0)
    np.testing.assert_allclose(got_t[1:], times[:n], rtol=1e-13, atol=1e-11)
    cdf = np.asarray(w).astype("f8").cumsum()
    cdf /= cdf[-1]
    want_p = np.minimum(np.searchsorted(cdf, uc[:n], "right"), 127)
    assert np.array_equal(got_p[1:].astype(int), want_p)
    x, y, _, _ = _philox4x32_10(np.uint64(0xFFFFFFFF), 1, sid, 4 | (0xFF << 8), k0, k1)      # event 0: index 0xFFFFFFFFFF
    assert int(got_p[0]) == min(int(np.searchsorted(cdf, _u53(x, y), "right")), 127)


def test_side_stream_branches_change_nothing(engine, monkeypatch):
    """Round 2: for small batches the background chain and the energy sampler's per-unit tables run on side streams
    next to the source chain (cgrb_energy_envelope_tables + cgrb_sample_energy_envelope_prepared); the six lists
    are the ones of the single-stream launch sequence, bit for bit."""
    from cosmogrb_b200 import engine as eng_mod

    dets = ("n0", "b0", "n5")
    tabs = [h.response(d).device_table(h.bkg_weights(d), "linear") for d in dets]
    units = np.concatenate([h.unit(det_index=i, stream_id=40 + i, **h.CONFTEST_BRIGHT) for i in range(len(dets))])
    out = {}
    for on in (False, True, True):
        monkeypatch.setattr(eng_mod, "BACKGROUND_SIDE_STREAM", on)
        b = engine.new_batch(units.copy(), tabs)
        engine.process(b, Philox(4242))
        c = engine.fetch_counts(b)
        engine.check_status(b)
        lists = [b.view(name, i, kind).copy() for i in range(len(dets))
                 for name, kind in (("times_out", "kept"), ("pha_out", "kept"), ("times_src", "src"), ("pha_src", "src"),
                                    ("times_bkg", "bkg"), ("pha_bkg", "bkg"))]
        out.setdefault(on, []).append((c.copy(), lists))
    (c0, l0), (c1, l1), (c2, l2) = out[False][0], out[True][0], out[True][1]
    for k in ("n_src", "n_bkg", "n_kept", "n_candidates", "n_trials", "status"):
        assert np.array_equal(c0[k], c1[k]) and np.array_equal(c0[k], c2[k]), k
    assert all(np.array_equal(a, b_) for a, b_ in zip(l0, l1)) and all(np.array_equal(a, b_) for a, b_ in zip(l0, l2))


def test_short_segments_of_small_units_are_the_same_process(engine, monkeypatch):
    """Round 2: a unit with few candidate slots is thinned in 1024-candidate segments instead of 8192 (the kernels take
    a segment's length from the work list).  The candidates and their uniforms are the same (Philox counters are
    per candidate); only the association of the running gap sums changes, i.e. the last bits of the times."""
    from cosmogrb_b200 import engine as eng_mod

    tab = [h.response("n0").device_table(h.bkg_weights("n0"), "linear")]
    got = {}
    for small in ((0, 1024), (1 << 40, 1024), (1 << 40, 2048)):
        monkeypatch.setattr(eng_mod, "SMALL_UNIT_SEGMENTS", small)
        b = engine.new_batch(h.unit(stream_id=3, **dict(h.CONFTEST_BRIGHT, peak_flux=1e-4)), tab)
        engine.sample_events(b, Philox(777))
        c = engine.fetch_counts(b)
        engine.check_status(b)
        got[small] = (int(c["n_candidates"][0]), b.view("times_src", 0, "src").copy())
    ref_n, ref_t = got[(0, 1024)]
    assert ref_n > 8192                # more than one full-length segment in the reference chunking
    for k in ((1 << 40, 1024), (1 << 40, 2048)):
        assert got[k][0] == ref_n and len(got[k][1]) == len(ref_t)
        np.testing.assert_allclose(got[k][1], ref_t, rtol=1e-13, atol=1e-13)
