"""Second set of golden vectors from the REAL reference (numba kernels imported from /root/reference).

TEST INFRASTRUCTURE, authoring container only (needs /root/reference + numba; never runs on the GPU
box).  Usage:   python oracle/make_golden_r2.py        (writes tests/golden/spectral_extra_*.npz)

Leaves the files of oracle/make_golden.py untouched and adds what they do not cover:
  * ConstantCPL (cpl_constant_functions.py:66-177) bright enough to produce ~1300 events, on a NaI and a
    BGO detector: sample_events -> sample_energy -> _digitize on one numba stream;
  * the time-evolving CPL (cpl_functions.py:134-286) on detectors n5 and b0;
  * a window of the benchmarked bright long burst (BASELINE.json configs[1]) thinned with the whole burst's
    fmax (cpl_functions.py:134-188).
Every vector is checked against the C oracle as it is minted.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CONSTANT_BRIGHT = dict(z=1.0, peak_flux=5e-6, alpha=-0.66, ep=500.0, duration=2.0)
C2_WINDOW = (60.0, 60.4)        # seconds of the bright long burst, near the pulse maximum


def main(extrap):
    from oracle import oracle as orc
    from oracle import ref_shims
    from tests import helpers as h

    ref_shims.install(extrap)
    ns = ref_shims.load()
    out = {}

    def ref_fmax(rsp, p, kind):
        ea = rsp.effective_area_packed
        grid = np.linspace(0.0, p["duration"], 50)                     # source.py:88-110
        if kind == 0:
            f = ns.cpl_functions.energy_integrated_evolution
            v = [f(rsp.emin, rsp.emax, np.array([t]), p["peak_flux"], p["ep_start"], p["ep_tau"], p["alpha"],
                   p["trise"], p["tdecay"], ea, p["z"]) for t in grid]
        else:
            f = ns.cpl_constant_functions.energy_integrated_evolution
            v = [f(rsp.emin, rsp.emax, np.array([t]), p["peak_flux"], p["ep"], p["alpha"], ea, p["z"]) for t in grid]
        return float(np.max(v))

    def source_chain(key, det, p, kind, seed, n_energy, window=None):
        rsp = h.response(det)
        ea = rsp.effective_area_packed
        fmax = ref_fmax(rsp, p, kind)
        t0, t1 = window if window is not None else (0.0, p["duration"])
        ns.numba_seed(seed)
        if kind == 0:
            times = ns.cpl_functions.sample_events(rsp.emin, rsp.emax, t0, t1, p["peak_flux"], p["ep_start"],
                                                   p["ep_tau"], p["alpha"], p["trise"], p["tdecay"], ea, fmax, p["z"])
        else:
            times = ns.cpl_constant_functions.sample_events(rsp.emin, rsp.emax, t0, t1, p["peak_flux"], p["ep"],
                                                            p["alpha"], ea, fmax, p["z"])
        s = h.oracle_source(rsp, kind=kind, interp_extrap=extrap, **p)
        assert abs(orc.fmax(s, ea, 0.0, p["duration"]) - fmax) <= 1e-12 * fmax
        rng = orc.Rng(seed=seed)
        ot, nc = orc.sample_events(s, ea, t0, t1, fmax, rng)
        assert len(ot) == len(times), (key, len(ot), len(times))
        np.testing.assert_allclose(ot, times, rtol=1e-12)
        out[f"fmax_{key}"], out[f"events_{key}"], out[f"ncand_{key}"] = fmax, times, nc
        if n_energy:
            n_e = min(len(times), n_energy)
            if kind == 0:
                energies = ns.cpl_functions.sample_energy(times[:n_e], p["peak_flux"], p["ep_start"], p["ep_tau"],
                                                          p["alpha"], p["trise"], p["tdecay"], rsp.emin, rsp.emax, ea,
                                                          p["z"])
            else:
                energies = ns.cpl_constant_functions.sample_energy(times[:n_e], p["peak_flux"], p["ep"], p["alpha"],
                                                                   rsp.emin, rsp.emax, ea, p["z"])
            pha = ns.response._digitize(energies, rsp.energy_edges, rsp.cumulative_matrix)
            oe, otr = orc.sample_energy(s, ea, times[:n_e], rng)
            np.testing.assert_allclose(oe, energies, rtol=1e-12)
            op = orc.digitize(energies, rsp.energy_edges, rsp.cumulative_matrix, rng)
            assert np.array_equal(op, pha)
            out[f"energy_{key}"], out[f"trials_{key}"], out[f"pha_{key}"] = energies, otr, pha.astype(np.uint8)
        print(key, len(times), "events of", nc, "candidates")

    for det in ("n0", "b0"):
        for seed in (4, 5):
            source_chain(f"constant_{det}_s{seed}", det, CONSTANT_BRIGHT, 1, seed, 300)
    for det in ("n5", "b0"):
        source_chain(f"bright_{det}_s0", det, h.CONFTEST_BRIGHT, 0, 0, 300)
    source_chain("long_window_n0_s7", "n0", h.BRIGHT_LONG, 0, 7, 0, window=C2_WINDOW)

    np.savez_compressed(os.path.join(GOLD, f"spectral_extra_{extrap}.npz"), **out)
    print("wrote", f"spectral_extra_{extrap}.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(sys.argv[1])
    else:
        for variant in ("linear", "clamp"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), variant])
