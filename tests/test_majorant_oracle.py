"""CPU: the piecewise-majorant proposal of the source thinning, restated in NumPy on the oracle's lambda(t).

The CUDA path (k_thin_nodes / k_thin_margins / k_thin_cum / k_source_thin, Philox only) draws thinning candidates
from a piecewise-constant bound P_j built from the squeeze table instead of the reference's single fmax
(cpl_functions.py:134-188).  Two things are checked here without a GPU, on the reference's own lambda(t) as
restated by the C oracle: (1) the bound holds -- min(p, 1) <= P_j on a grid 16 times finer than the table, for the
fixtures and for random population bursts (the margins are the same second-difference heuristic the squeeze test
relies on); (2) operational-time thinning with that bound is the same point process as the reference's loop
(counts Poisson around the same integral, two-sample KS on the arrival times against the oracle's MT19937 path).
"""
import numpy as np
import pytest

from tests import helpers as h
from oracle import oracle as orc

stats = pytest.importorskip("scipy.stats")

SQUEEZE_RATIO, MIN_CAND, MAX_NODES = 8, 2048, 1 << 14       # engine.py: squeeze_ratio / min candidates / max nodes


def _table(s, ea, fmax, duration):
    """nodes p_j, margins m_j, bounds P_j and cumulative operational time, as the kernels build them"""
    mean = fmax * duration
    cap = int(np.ceil(mean + 8.0 * np.sqrt(mean))) + 64
    if cap < MIN_CAND:
        return None
    n = int(np.clip(cap // SQUEEZE_RATIO, 64, MAX_NODES))
    t = np.linspace(0.0, duration, n)
    p = orc.energy_integrated_evolution(s, ea, t) / fmax
    d2 = np.abs(p[:-2] - 2.0 * p[1:-1] + p[2:])                 # second differences at the interior nodes 1 .. n-2
    m = np.empty(n - 1)
    for k in range(n - 1):                                      # cell [k, k+1]: nodes k-1 .. k+2, clamped
        c = np.clip(np.arange(k - 1, k + 3), 1, n - 2) - 1
        m[k] = 0.5 * d2[c].max() + 1e-12 * (abs(p[k]) + 1.0)
    P = np.minimum(1.0, np.maximum(np.maximum(p[:-1], p[1:]) + m, 0.0))
    L = np.concatenate(([0.0], np.cumsum(fmax * (duration / (n - 1)) * P)))
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
