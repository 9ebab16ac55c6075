"""Mint golden vectors from the REAL reference (numba kernels imported from /root/reference).

TEST INFRASTRUCTURE, authoring container only (needs /root/reference + numba; never runs on the
GPU box).  Usage:   python oracle/make_golden.py            (writes tests/golden/*.npz)

The reference is not seedable through its API; numba's MT19937 is seeded from
inside an @njit helper, which makes its stream identical to np.random.RandomState(s).random_sample,
and the Python-level ``np.random.seed()`` calls of the reference (OS-entropy reseeds of numpy's
global generator) are pinned by monkeypatching ``np.random.seed`` while the real classes run.

Every vector stores its inputs (or the seeds that regenerate them) next to the reference output.
The same script checks the C oracle against each vector as it is minted (fails loudly otherwise).
One process per ``interp_extrap`` variant because numba caches the compiled kernels.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CONSTANT = dict(z=1.0, peak_flux=5e-9, alpha=-0.66, ep=500.0, duration=2.0)     # test/conftest.py:83-92
WEAK = dict(z=1.0, peak_flux=5e-20, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=0.1, tdecay=1.0,
            duration=2.0)                                                      # test/conftest.py:58-70


def ref_lambda(ns, rsp, p, times, kind):
    ea = rsp.effective_area_packed
    if kind == 0:
        f = ns.cpl_functions.energy_integrated_evolution
        return np.array([f(rsp.emin, rsp.emax, np.array([t]), p["peak_flux"], p["ep_start"], p["ep_tau"], p["alpha"],
                           p["trise"], p["tdecay"], ea, p["z"]) for t in times])
    f = ns.cpl_constant_functions.energy_integrated_evolution
    return np.array([f(rsp.emin, rsp.emax, np.array([t]), p["peak_flux"], p["ep"], p["alpha"], ea, p["z"])
                     for t in times])


def main(extrap):
    from oracle import oracle as orc
    from oracle import ref_shims
    from tests import helpers as h

    ref_shims.install(extrap)
    ns = ref_shims.load()
    os.makedirs(GOLD, exist_ok=True)
    rsp = h.response("n0")
    ea = rsp.effective_area_packed
    out = {}

    # (i) lambda(t), evolution and fmax ---------------------------------------------------------
    cases = {"readme": (h.README_DEMO, 0), "bright": (h.CONFTEST_BRIGHT, 0), "long": (h.BRIGHT_LONG, 0),
             "weak": (WEAK, 0), "constant": (CONSTANT, 1)}
    for name, (p, kind) in cases.items():
        times = np.linspace(0.0, p["duration"], 101)
        lam = ref_lambda(ns, rsp, p, times, kind)
        grid = np.linspace(0.0, p["duration"], 50)
        fmax = ref_lambda(ns, rsp, p, grid, kind).max()          # source.py:88-110
        s = h.oracle_source(rsp, kind=kind, interp_extrap=extrap, **p)
        np.testing.assert_allclose(orc.energy_integrated_evolution(s, ea, times), lam, rtol=1e-12, atol=1e-300)
        assert abs(orc.fmax(s, ea, 0.0, p["duration"]) - fmax) <= 1e-12 * fmax
        out[f"lambda_{name}_t"], out[f"lambda_{name}"], out[f"fmax_{name}"] = times, lam, fmax
    p = h.README_DEMO
    e = np.logspace(1, 4, 25)
    t = np.linspace(0, 10, 6)
    evo = ns.cpl_functions.cpl_evolution(e, t, p["peak_flux"], p["ep_start"], p["ep_tau"], p["alpha"], p["trise"],
                                         p["tdecay"], rsp.emin, rsp.emax, p["z"])
    s = h.oracle_source(rsp, interp_extrap=extrap, **p)
    np.testing.assert_allclose(orc.cpl_evolution(s, e, t), evo, rtol=1e-12, atol=1e-300)
    out["evo_e"], out["evo_t"], out["evo"] = e, t, evo

    # (ii) sample_events, (iii) sample_energy -- one numba stream, as in LightCurve._sample_source -----
    for name in ("readme", "bright"):
        p, kind = cases[name]
        fmax = float(out[f"fmax_{name}"])
        for seed in (0, 1, 2):
            ns.numba_seed(seed)
            times = ns.cpl_functions.sample_events(rsp.emin, rsp.emax, 0.0, p["duration"], p["peak_flux"],
                                                   p["ep_start"], p["ep_tau"], p["alpha"], p["trise"], p["tdecay"],
                                                   ea, fmax, p["z"])
            n_e = min(len(times), 400)
            energies = ns.cpl_functions.sample_energy(times[:n_e], p["peak_flux"], p["ep_start"], p["ep_tau"],
                                                      p["alpha"], p["trise"], p["tdecay"], rsp.emin, rsp.emax, ea,
                                                      p["z"])
            pha = ns.response._digitize(energies, rsp.energy_edges, rsp.cumulative_matrix)
            s = h.oracle_source(rsp, interp_extrap=extrap, **p)
            rng = orc.Rng(seed=seed)
            ot, _ = orc.sample_events(s, ea, 0.0, p["duration"], fmax, rng)
            assert len(ot) == len(times), (name, seed, len(ot), len(times))
            np.testing.assert_allclose(ot, times, rtol=1e-12)
            oe, otr = orc.sample_energy(s, ea, times[:n_e], rng)
            np.testing.assert_allclose(oe, energies, rtol=1e-12)
            op = orc.digitize(energies, rsp.energy_edges, rsp.cumulative_matrix, rng)
            assert np.array_equal(op, pha)
            k = f"{name}_s{seed}"
            out[f"events_{k}"], out[f"energy_{k}"], out[f"trials_{k}"], out[f"pha_{k}"] = times, energies, otr, pha
    # constant CPL (cpl_constant_functions.py)
    p = CONSTANT
    fmax = float(out["fmax_constant"])
    ns.numba_seed(4)
    cc = ns.cpl_constant_functions
    times = cc.sample_events(rsp.emin, rsp.emax, 0.0, p["duration"], p["peak_flux"], p["ep"], p["alpha"], ea, fmax, p["z"])
    energies = cc.sample_energy(times[:300], p["peak_flux"], p["ep"], p["alpha"], rsp.emin, rsp.emax, ea, p["z"])
    s = h.oracle_source(rsp, kind=1, interp_extrap=extrap, **p)
    rng = orc.Rng(seed=4)
    ot, _ = orc.sample_events(s, ea, 0.0, p["duration"], fmax, rng)
    assert len(ot) == len(times)
    np.testing.assert_allclose(ot, times, rtol=1e-12)
    oe, otr = orc.sample_energy(s, ea, times[:300], rng)
    np.testing.assert_allclose(oe, energies, rtol=1e-12)
    out["events_constant_s4"], out["energy_constant_s4"], out["trials_constant_s4"] = times, energies, otr

    if extrap != "linear":
        np.savez_compressed(os.path.join(GOLD, f"spectral_{extrap}.npz"), **out)
        print("wrote", f"spectral_{extrap}.npz")
        return
    np.savez_compressed(os.path.join(GOLD, f"spectral_{extrap}.npz"), **out)

    # (iv) _digitize on crafted energies (regenerated from the seed by the tests) ------------------
    out = {}
    for det in ("n0", "b0"):
        r = h.response(det)
        e, _ = digitize_inputs(r, 60000)
        ns.numba_seed(21)
        pha = ns.response._digitize(e, r.energy_edges, r.cumulative_matrix)
        assert np.array_equal(orc.digitize(e, r.energy_edges, r.cumulative_matrix, orc.Rng(seed=21)), pha)
        out[f"pha_{det}"] = pha.astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, "digitize.npz"), **out)

    # (v) _gbm_dead_time ------------------------------------------------------------------------
    out = {}
    t = np.array([0, 1, 3, 4, 10, 15, 20.5, 20.7, 31.2, 32]) * 1e-6
    pha = np.array([5, 6, 7, 8, 127, 9, 10, 11, 127, 12], dtype="f8")
    _, _, sel = ns.gbm_lightcurve._gbm_dead_time(t, pha, len(t))
    assert np.array_equal(sel.astype(bool), np.array([1, 0, 1, 1, 1, 0, 0, 1, 1, 0], dtype=bool))
    out["crafted_t"], out["crafted_pha"], out["crafted_sel"] = t, pha, sel.astype(bool)
    for i, (rate, frac) in enumerate(((500.0, 0.05), (2e5, 0.05), (2e6, 0.3), (5e6, 1.0))):
        t, pha = dead_time_inputs(i, rate, frac, 500_000)
        _, _, sel = ns.gbm_lightcurve._gbm_dead_time(t, pha.astype("f8"), len(t))
        assert np.array_equal(orc.gbm_dead_time(t, pha), sel.astype(bool))
        out[f"random{i}_sel"] = np.packbits(sel.astype(bool))
        out[f"random{i}_args"] = np.array([rate, frac, 500_000])
    np.savez_compressed(os.path.join(GOLD, "dead_time.npz"), **out)

    # (vi) background times + channels -------------------------------------------------------------
    out = {}
    ns.numba_seed(9)
    bt = ns.background.background_poisson_generator(-10.0, 30.0, 497.25)
    assert np.allclose(orc.background_times(-10.0, 30.0, 497.25, orc.Rng(seed=9)), bt, rtol=1e-12, atol=1e-12)
    tmpl = ns.background.BackgroundSpectrumTemplate(h.gbm_tables()["bkg_counts"][0])
    real_seed = np.random.seed
    np.random.seed = lambda *a: real_seed(77)
    try:
        bc = tmpl.sample_channel(size=len(bt))
    finally:
        np.random.seed = real_seed
    oc = orc.background_channels(h.bkg_weights("n0"), len(bt), orc.Rng(seed=77))
    assert np.array_equal(oc, bc)
    out["times"], out["channels"] = bt, np.asarray(bc).astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, "background.npz"), **out)

    # (vii) the real per-detector pipeline: GBMLightCurve.process() with the real classes -------------
    out = {}
    p = h.CONFTEST_BRIGHT
    Source, CPLSourceFunction = ns.source.Source, ns.cpl_source.CPLSourceFunction
    Background, GBMLightCurve = ns.background.Background, ns.gbm_lightcurve.GBMLightCurve
    RefResponse = ns.response.Response

    class Rsp(RefResponse):
        separation_angle = 30.0
        trigger_time = 0.0
        T0 = 0.0

    rrsp = Rsp(matrix=rsp.matrix, geometric_area=rsp.geometric_area, energy_edges=rsp.energy_edges,
               channel_edges=rsp.channel_edges)
    assert np.array_equal(rrsp._cumulative_maxtrix, rsp.cumulative_matrix)       # host tables bit-identical
    assert np.array_equal(rrsp.effective_area_packed, rsp.effective_area_packed)
    np.random.seed = lambda *a: real_seed(123)
    try:
        np.random.seed()
        sf = CPLSourceFunction(response=rrsp, peak_flux=p["peak_flux"], trise=p["trise"], tdecay=p["tdecay"],
                               ep_tau=p["ep_tau"], alpha=p["alpha"], ep_start=p["ep_start"])
        src = Source(0.0, p["duration"], sf, z=p["z"])
        bkg = Background(-20.0, 30.0, average_rate=500, background_spectrum_template=tmpl)
        lc = GBMLightCurve(src, bkg, rrsp, name="n0", grb_name="golden", tstart=-20.0, tstop=30.0)
        ns.numba_seed(2024)
        st = lc.process()
    finally:
        np.random.seed = real_seed
    rate = bkg._background_rate
    s = h.oracle_source(rsp, interp_extrap=extrap, **p)
    fmax = orc.fmax(s, ea, 0.0, p["duration"])
    assert abs(fmax - src._fmax) <= 1e-12 * fmax
    rng = orc.Rng(seed=2024)
    ts, _ = orc.sample_events(s, ea, 0.0, p["duration"], fmax, rng)
    es, _ = orc.sample_energy(s, ea, ts, rng)
    ps = orc.digitize(es, rsp.energy_edges, rsp.cumulative_matrix, rng)
    tb = orc.background_times(-20.0, 30.0, rate, rng)
    pb = orc.background_channels(h.bkg_weights("n0"), len(tb), orc.Rng(seed=123))
    tt, pt = orc.combine(tb, pb, ts, ps)
    sel = orc.gbm_dead_time(tt, pt)
    np.testing.assert_allclose(ts, st.times_source, rtol=1e-12)
    assert np.array_equal(ps, st.pha_source)
    np.testing.assert_allclose(tb, st.times_background, rtol=1e-12, atol=1e-12)
    assert np.array_equal(pb, st.pha_background)
    np.testing.assert_allclose(tt[sel], st.times, rtol=1e-12, atol=1e-12)
    assert np.array_equal(pt[sel], st.pha)
    out.update(times=st.times, pha=np.asarray(st.pha).astype(np.uint8), times_source=st.times_source,
               pha_source=np.asarray(st.pha_source).astype(np.uint8), times_background=st.times_background,
               pha_background=np.asarray(st.pha_background).astype(np.uint8), bkg_rate=rate, fmax=src._fmax,
               seeds=np.array([2024, 123]))
    np.savez_compressed(os.path.join(GOLD, "pipeline.npz"), **out)
    print("pipeline:", len(st.times_source), "source,", len(st.times_background), "background,", len(st.times), "kept")
    print("golden vectors written to", GOLD)


def digitize_inputs(rsp, n, seed=11):
    """Energies incl. exact edges (E == emin wraps to the last row), zero-probability rows."""
    rs = np.random.RandomState(seed)
    e = np.exp(rs.uniform(np.log(rsp.emin), np.log(rsp.emax), n))
    edges = rsp.energy_edges
    e[:141] = edges
    e[141:300] = rs.uniform(edges[0], edges[3], 159)
    return e, rs


def dead_time_inputs(i, rate, frac127, n):
    rs = np.random.RandomState(100 + i)
    t = np.cumsum(rs.exponential(1.0 / rate, n))
    pha = rs.randint(0, 127, n)
    pha[rs.random_sample(n) < frac127] = 127
    return t, pha


if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(sys.argv[1])
    else:
        for variant in ("linear", "clamp"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), variant])
