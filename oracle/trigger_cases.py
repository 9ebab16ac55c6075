"""Light curves of the GBM-trigger golden vectors (TEST INFRASTRUCTURE): regenerated from seeds by the C
oracle's MT19937 pipeline, so tests/golden/trigger.npz only stores the reference's outputs."""
import numpy as np

from tests import helpers as h

CASES = {
    "bright_n0": ("n0", h.CONFTEST_BRIGHT, 11),
    "bright_n5": ("n5", h.CONFTEST_BRIGHT, 12),
    "medium_n0": ("n0", dict(h.CONFTEST_BRIGHT, peak_flux=6e-7), 13),
    "faint_n0": ("n0", dict(h.CONFTEST_BRIGHT, peak_flux=1.5e-7), 14),
    "weak_n0": ("n0", dict(h.CONFTEST_BRIGHT, peak_flux=5e-20), 15),
    "readme_n3": ("n3", h.README_DEMO, 16),
    "bright_b0": ("b0", dict(h.CONFTEST_BRIGHT, peak_flux=3e-5), 17),
}


def light_curve(det, params, seed, bkg=(-100.0, 300.0, 500.0)):
    """one detector's LightCurve.process() on the CPU oracle: the six event lists"""
    from oracle import oracle as orc
    from tests import helpers as h

    rsp = h.response(det)
    ea = rsp.effective_area_packed
    s = h.oracle_source(rsp, **params)
    fmax = orc.fmax(s, ea, 0.0, params["duration"])
    rng = orc.Rng(seed=seed)
    ts, _ = orc.sample_events(s, ea, 0.0, params["duration"], fmax, rng)
    es, _ = orc.sample_energy(s, ea, ts, rng)
    ps = orc.digitize(es, rsp.energy_edges, rsp.cumulative_matrix, rng)
    tb = orc.background_times(bkg[0], bkg[1], bkg[2], rng)
    pb = orc.background_channels(h.bkg_weights(det), len(tb), orc.Rng(seed=seed + 1000))
    t, p = orc.combine(tb, pb, ts, ps)
    sel = orc.gbm_dead_time(t, p)
    return rsp, dict(times=t[sel], pha=p[sel].astype(np.int64), times_source=ts, pha_source=ps,
                     times_background=tb, pha_background=pb)


