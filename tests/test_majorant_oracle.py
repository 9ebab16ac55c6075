"""CPU: the piecewise-majorant proposal of the source thinning, restated in NumPy on the oracle's lambda(t).

The CUDA path (k_thin_nodes / k_thin_margins / k_thin_cum / k_source_thin, Philox only) draws thinning candidates
from a piecewise-constant bound P_j built from the squeeze table instead of the reference's single fmax
(cpl_functions.py:134-188).  Two things are checked here without a GPU, on the reference's own lambda(t) as
restated by the C oracle: (1) the bound holds -- min(p, 1) <= P_j on a grid 16 times finer than the table, for the
fixtures and for random population bursts, and the squeeze margins bound |p - linear interpolation| on every cell
over a sweep of trise / tdecay / duration / table size that includes the pulses whose essential singularity at
t = 0 defeated a pure second-difference margin; (2) operational-time thinning with that bound is the same point process as the reference's loop
(counts Poisson around the same integral, two-sample KS on the arrival times against the oracle's MT19937 path).
"""
import numpy as np
import pytest

from tests import helpers as h
from oracle import oracle as orc

stats = pytest.importorskip("scipy.stats")

SQUEEZE_RATIO, MIN_CAND, MAX_NODES = 8, 2048, 1 << 14       # engine.py: squeeze_ratio / min candidates / max nodes


def _margins(p, pm):
    """k_thin_margins: per cell, 2 x the interpolation error measured at the midpoint + 0.5 x the largest second
    difference of the nodes k-1 .. k+2 + rounding slack; infinite (no squeeze) in cells touching a node where p == 0
    without being identically zero.  (The first cell of a pulse is overwritten by ``_first_cell_bound``.)"""
    n = len(p)
    d2 = np.abs(p[:-2] - 2.0 * p[1:-1] + p[2:])                 # second differences at the interior nodes 1 .. n-2
    idx = np.clip(np.arange(n - 1)[:, None] + np.arange(-1, 3)[None, :], 1, n - 2) - 1
    m = 0.5 * d2[idx].max(axis=1) + 2.0 * np.abs(pm - 0.5 * (p[:-1] + p[1:])) + 1e-12 * (np.abs(p[:-1]) + 1.0)
    z0, z1, zm = p[:-1] == 0, p[1:] == 0, pm == 0
    m[(z0 | z1) & ~(z0 & z1 & zm)] = np.inf
    return m


def _first_cell_bound(s, ea, fmax, h):
    """k_thin_nodes, first cell of a time-evolving CPL: B = sup K x sup G / fmax with K the Norris pulse (its sup over
    [0, h] is at min(h, sqrt(trise tdecay))) and G = lambda / K, evaluated as the constant-CPL rate at unit flux and
    ep(t), at the cell's ends and midpoint"""
    tpk = np.sqrt(s.trise * s.tdecay)
    supK = orc.lib().co_norris(min(max(tpk, 0.0), h), s.peak_flux, 0.0, s.trise, s.tdecay)
    g = []
    for t in (0.0, 0.5 * h, h):
        c = orc.source(kind=1, interp_extrap=s.interp_extrap, peak_flux=1.0, ep_start=s.ep_start / (1.0 + t / s.ep_tau),
                       alpha=s.alpha, z=s.z, emin=s.emin, emax=s.emax)
        g.append(orc.energy_integrated_evolution(c, ea, [0.0])[0])
    supG = max(g) + 2.0 * abs(g[0] - 2.0 * g[1] + g[2])
    return supK * supG / fmax * (1.0 + 1e-6) + 1e-300


def _table(s, ea, fmax, duration, n=None):
    """nodes p_j, margins m_j (inf: no squeeze), bounds P_j and cumulative operational time, as the kernels build them"""
    mean = fmax * duration
    cap = int(np.ceil(mean + 8.0 * np.sqrt(mean))) + 64
    if n is None:
        if cap < MIN_CAND:
            return None
        n = int(np.clip(cap // SQUEEZE_RATIO, 64, MAX_NODES))
    t = np.linspace(0.0, duration, n)
    h = duration / (n - 1)
    p = orc.energy_integrated_evolution(s, ea, t) / fmax
    pm = orc.energy_integrated_evolution(s, ea, t[:-1] + 0.5 * h) / fmax
    m = _margins(p, pm)
    with np.errstate(invalid="ignore"):
        P = np.minimum(1.0, np.maximum(np.maximum(p[:-1], p[1:]) + m, 0.0))
    if s.kind == 0:
        m[0] = np.inf                                           # every candidate of the first cell is exact ...
        P[0] = min(1.0, _first_cell_bound(s, ea, fmax, h))      # ... and the proposal uses the analytic bound
    L = np.concatenate(([0.0], np.cumsum(fmax * h * P)))
    return t, p, m, P, L


def _cases():
    out = [("readme", h.README_DEMO, "n0"), ("bright", h.CONFTEST_BRIGHT, "n3"), ("long", h.BRIGHT_LONG, "b0")]
    rs = np.random.RandomState(4)
    for i in range(9):                                          # population-like draws, bright enough for a table
        t90 = 10 ** rs.normal(1.0, 0.25)
        trise = float(np.clip(rs.normal(1.0, 1.0), 0.05, 5.0))
        out.append((f"pop{i}", dict(z=float(rs.uniform(0.3, 4.0)), peak_flux=float(10 ** rs.uniform(-5.8, -4.0)),
                                    alpha=float(rs.uniform(-1.45, -0.2)), ep=float(10 ** rs.normal(2.3, 0.4)),
                                    ep_tau=float(rs.uniform(1.5, 2.5)), trise=trise,
                                    tdecay=(10 * t90 + trise + np.sqrt(trise) * np.sqrt(20 * t90 + trise)) / 50,
                                    duration=1.5 * t90), ["n1", "n7", "b1"][i % 3]))
    return out


@pytest.mark.parametrize("name,params,det", _cases(), ids=[c[0] for c in _cases()])
def test_majorant_dominates_lambda(name, params, det):
    rsp = h.response(det)
    s = h.oracle_source(rsp, **params)
    ea = rsp.effective_area_packed
    fmax = orc.fmax(s, ea, 0.0, params["duration"])
    tab = _table(s, ea, fmax, params["duration"])
    if tab is None:
        pytest.skip("too few candidates for a table: the literal proposal is used")
    t, p, m, P, L = tab
    n = len(t)
    fine = np.linspace(0.0, params["duration"], 16 * (n - 1) + 1)
    pf = np.minimum(orc.energy_integrated_evolution(s, ea, fine) / fmax, 1.0)
    cell = np.minimum((np.arange(len(fine)) // 16), n - 2)
    # a point on a node belongs to both neighbouring cells: the larger bound applies
    bound = np.maximum(P[cell], P[np.maximum(cell - 1, 0)] * (np.arange(len(fine)) % 16 == 0))
    assert np.all(pf <= bound * (1 + 1e-9) + 1e-300), float(np.max(pf - bound))
    # the squeeze margins themselves: |p - linear interpolation| <= m on every cell
    lin = np.interp(fine, t, p)
    err = np.abs(orc.energy_integrated_evolution(s, ea, fine) / fmax - lin)
    assert np.all(err <= m[cell] + 1e-15), float(np.max(err - m[cell]))
    # and the bound is useful: far fewer expected candidates than fmax * T for a pulse
    assert L[-1] <= fmax * params["duration"] * (1 + 1e-12)
    assert L[-1] >= np.trapezoid(pf, fine) * fmax * (1 - 1e-9)


def test_operational_time_thinning_is_the_reference_process():
    params, det = h.README_DEMO, "n0"          # ~8000 candidates at fmax, ~250 events per realisation
    rsp = h.response(det)
    s = h.oracle_source(rsp, **params)
    ea = rsp.effective_area_packed
    T = params["duration"]
    fmax = orc.fmax(s, ea, 0.0, T)
    t, p, m, P, L = _table(s, ea, fmax, T)
    n = len(t)
    rs = np.random.RandomState(123)
    got, n_cand = [], 0
    for rep in range(40):
        # unit-rate Poisson process on [0, Lambda_T] -> real time through the piecewise-linear Lambda -> thin
        k = rs.poisson(L[-1])
        tau = np.sort(rs.uniform(0.0, L[-1], k))
        j = np.clip(np.searchsorted(L, tau, "right") - 1, 0, n - 2)
        tt = t[j] + (tau - L[j]) / (L[j + 1] - L[j]) * (t[j + 1] - t[j])
        pe = np.minimum(orc.energy_integrated_evolution(s, ea, tt) / fmax, 1.0)
        keep = rs.uniform(size=k) * P[j] <= pe
        got.append(tt[keep])
        n_cand += k
    got = np.concatenate(got)
    ref = np.concatenate([orc.sample_events(s, ea, 0.0, T, fmax, orc.Rng(seed=500 + r))[0][1:] for r in range(40)])
    mean = 40 * np.trapezoid(np.minimum(orc.energy_integrated_evolution(s, ea, np.linspace(0, T, 4001)), fmax),
                             np.linspace(0, T, 4001))
    assert abs(len(got) - mean) < 5 * np.sqrt(mean) and abs(len(ref) - mean) < 5 * np.sqrt(mean)
    assert stats.ks_2samp(got, ref).pvalue > 0.01
    assert n_cand < 0.8 * 40 * fmax * T                         # the point of the bound


# (duration, trise, tdecay, nodes): the advisor's counter-examples to the old margins (T=10, trise=1e-3, tdecay=1,
# n=16384 and T=100, trise=0.01, tdecay=1) and the corners of the population's range, small and large tables
SWEEP = [(10.0, 1e-3, 1.0, 16384), (100.0, 0.01, 1.0, 16384), (2.0, 1e-3, 0.1, 1024), (300.0, 5.0, 0.1, 64),
         (300.0, 1e-3, 0.1, 64), (100.0, 5.0, 1.0, 64), (300.0, 5.0, 60.0, 16384), (10.0, 0.1, 10.0, 1024),
         (100.0, 1.0, 60.0, 64), (2.0, 5.0, 1.0, 1024)]


@pytest.mark.parametrize("T,trise,tdecay,n", SWEEP)
def test_squeeze_margins_bound_the_interpolation_error(T, trise, tdecay, n):
    rsp = h.response("n0")
    ea = rsp.effective_area_packed
    s = h.oracle_source(rsp, z=1.0, peak_flux=1e-5, alpha=-0.66, ep_start=500.0, ep_tau=2.0, trise=trise,
                        tdecay=tdecay, duration=T)
    fmax = orc.fmax(s, ea, 0.0, T)
    assert fmax > 0 and np.isfinite(fmax)
    t, p, m, P, L = _table(s, ea, fmax, T, n=n)
    F = 8 if n > 4096 else 32
    fine = np.linspace(0.0, T, F * (n - 1) + 1)
    cell = np.minimum(np.arange(len(fine)) // F, n - 2)
    err = np.abs(orc.energy_integrated_evolution(s, ea, fine) / fmax - np.interp(fine, t, p))
    assert np.all(err <= m[cell] + 1e-15), (float(np.max(err - m[cell])), np.unique(cell[err > m[cell] + 1e-15])[:8])
    # the first cell of a pulse is never interpolated (the old margins failed there); its majorant is the analytic
    # bound, which must dominate p on the cell and stay far below 1 for a pulse that starts slowly
    assert np.isinf(m[0])
    pf = orc.energy_integrated_evolution(s, ea, fine[:F + 1]) / fmax
    assert np.all(np.minimum(pf, 1.0) <= P[0] * (1 + 1e-12))
    if trise / (T / (n - 1)) > 50:
        assert P[0] < 1e-6
    finite = np.isfinite(m[cell])
    assert np.max(err[finite] / m[cell][finite]) < 0.75         # measured worst case 0.45: head-room, not luck


@pytest.mark.parametrize("name,params,det", _cases()[:4])
def test_cell_bound_from_the_cell_width(name, params, det):
    """Round 2: k_source_thin takes a cell's majorant from the cell's own width in operational time,
    P = (L[j+1] - L[j]) / (fmax h) snapped to 1 above 1 - 1e-9, and the position inside the cell from the stored
    reciprocal width.  The cumulative sums round, so P differs from the tabulated bound -- by far less than the 1e-9
    tolerance of the kernel's self-check -- and a cell clamped at the reference's ceiling reads exactly 1."""
    rsp = h.response(det)
    s = h.oracle_source(rsp, **params)
    ea = rsp.effective_area_packed
    T = params["duration"]
    fmax = orc.fmax(s, ea, 0.0, T)
    tab = _table(s, ea, fmax, T)
    if tab is None:
        pytest.skip("no table for this unit")
    t, p, m, P, L = tab
    hh = T / (len(t) - 1)
    wdt = np.diff(L)
    Pw = wdt * (1.0 / (fmax * hh))
    Pk = np.where(Pw > 1.0 - 1e-9, 1.0, Pw)
    assert np.all(np.abs(Pk - P) <= 1e-10 * np.maximum(P, 1e-300) + 1e-13 * L[-1] / (fmax * hh))
    assert np.all(Pk[P == 1.0] == 1.0)
    # the map tau -> t inside a cell by the stored reciprocal: monotone, inside the cell, ends on the nodes
    invw = np.where(wdt > 0.0, 1.0 / np.where(wdt > 0.0, wdt, 1.0), 0.0)
    rs = np.random.RandomState(5)
    j = rs.randint(0, len(wdt), 20000)
    tau = L[j] + np.sort(rs.uniform(0.0, 1.0, 20000)) * wdt[j]
    fr = np.clip((tau - L[j]) * invw[j], 0.0, 1.0)
    tt = np.minimum(t[j] + fr * hh, t[j + 1])
    assert np.all((tt >= t[j]) & (tt <= t[j + 1]))
    exact = t[j] + np.where(wdt[j] > 0, (tau - L[j]) / np.where(wdt[j] > 0, wdt[j], 1.0), 0.0) * hh
    assert np.all(np.abs(tt - exact) <= 4e-16 * (np.abs(exact) + hh))
