"""CPU: the fp32 prefilter of the trigger search (k_trig_eval) never dismisses a window that fires.

The kernel stages the exposure and count prefix sums of a chunk as fp32 relative to the chunk start and dismisses
a (window, gap) pair when  d32 <= 0  or  d32^2 < 0.9 thr^2 B ex eb  (d = S eb - B ex); everything else goes to the
exact fp64 test (gbm_lightcurve_analyzer.py:199-271 restated).  Here the same arithmetic is emulated with NumPy
float32 on synthetic light curves -- quiet background, bright and faint bursts, high dead-time fractions, thresholds
from 1 to 30 -- and compared with the fp64 significance: no window with sig >= thr may be dismissed, and the
prefilter must do its job (dismiss nearly all quiet windows)."""
import numpy as np
import pytest

N_BKG, N_SRC, DT = 1062, 250, 0.016         # floor(17 / 0.016), floor(4 / 0.016)
CH, SPAN = 1792, 1832                        # TRIG_CH, TRIG_SPAN


def _curve(rs, nb, rate, burst):
    lam = np.full(nb, rate * DT)
    if burst:
        k0 = rs.randint(1500, nb - 800)
        lam[k0:k0 + rs.randint(4, 600)] += burst * DT
    counts = rs.poisson(lam).astype(np.int64)
    # exposure per bin = width - dead time (2 us per event, 10 us for a 5 % overflow share), as the kernel does
    ovf = rs.binomial(counts, 0.05)
    expo = DT - ((counts - ovf) * 2e-6 + ovf * 10e-6)
    return counts, expo


@pytest.mark.parametrize("rate,burst,thr", [(500.0, 0.0, 4.5), (500.0, 400.0, 4.5), (500.0, 5e4, 4.5), (3000.0, 2e3, 8.0),
                                            (2.5e4, 2e5, 30.0), (60.0, 100.0, 1.0), (500.0, 1500.0, 2.0)])
def test_prefilter_has_no_false_dismissals(rate, burst, thr):
    rs = np.random.RandomState(int(rate + burst) % 9973)
    nb = 9000
    counts, expo = _curve(rs, nb, rate, burst)
    cs = np.concatenate(([0], np.cumsum(counts)))                 # int64 prefix, cs[k] = sum of bins < k
    es = np.concatenate(([0.0], np.cumsum(expo)))
    thr2f = np.float32(thr * thr)
    min_eb, min_ex = np.float32(0.25) * np.float32(N_BKG * DT), np.float32(0.25) * np.float32(N_SRC * DT)
    n_i_max = nb - (N_BKG + 1 + N_SRC)
    fired = dismissed = total = 0
    for c0 in range(0, n_i_max, CH):
        hi = min(nb, c0 + CH + SPAN)
        assert cs[hi] - cs[c0] < (1 << 24)                        # the kernel's condition for using the prefilter
        s_ef = (es[c0:hi + 1] - es[c0]).astype(np.float32)
        s_cf = (cs[c0:hi + 1] - cs[c0]).astype(np.float32)
        li = np.arange(0, min(c0 + CH, n_i_max) - c0)
        i = c0 + li
        Bf, ebf = s_cf[li + N_BKG] - s_cf[li], s_ef[li + N_BKG] - s_ef[li]
        B, eb = (cs[i + N_BKG] - cs[i]).astype("f8"), es[i + N_BKG] - es[i]
        for gx in range(9, -1, -1):
            n_gap = 1 << gx
            ok = i + N_BKG + n_gap + N_SRC < nb
            ls, gs = li + N_BKG + n_gap, i + N_BKG + n_gap
            ls, gs_ = ls[ok], gs[ok]
            exf, Sf = s_ef[ls + N_SRC] - s_ef[ls], s_cf[ls + N_SRC] - s_cf[ls]
            d32 = Sf * ebf[ok] - Bf[ok] * exf                     # float32 throughout, like the kernel
            kf = np.float32(0.9) * thr2f * Bf[ok] * ebf[ok]
            pre_i = (Bf[ok] > 0) & (ebf[ok] > min_eb)
            dismiss = pre_i & (exf > min_ex) & ((d32 <= 0) | (d32 * d32 < kf * exf))
            ex, S = es[gs_ + N_SRC] - es[gs_], (cs[gs_ + N_SRC] - cs[gs_]).astype("f8")
            with np.errstate(divide="ignore", invalid="ignore"):
                nbk = ex / eb[ok] * B[ok]
                sig = (S - nbk) / np.sqrt(nbk)
            fires = sig >= thr
            assert not np.any(dismiss & fires), (gx, float(sig[dismiss & fires].max()))
            # margin: nothing within 2 % of the threshold is dismissed either
            assert not np.any(dismiss & (sig >= 0.98 * thr))
            fired += int(fires.sum())
            dismissed += int(dismiss.sum())
            total += int(ok.sum())
    if burst == 0.0:
        assert fired == 0 or thr < 2.0
        assert dismissed > 0.999 * total                         # quiet light curve: the exact path stays idle
    if burst >= 1500.0:
        assert fired > 0


@pytest.mark.parametrize("rate,burst,thr", [(500.0, 0.0, 4.5), (500.0, 400.0, 4.5), (500.0, 5e4, 4.5), (3000.0, 2e3, 8.0),
                                            (2.5e4, 2e5, 30.0), (60.0, 100.0, 1.0), (500.0, 1500.0, 2.0), (500.0, 3e5, 4.5)])
def test_level0_source_window_filter_has_no_false_dismissals(rate, burst, thr):
    """Round 2: before the fp32 test the search compares the source-window sum S(m) alone with
    S_min(i) = a_min B (1 - 1e-5) + 0.94 thr sqrt(a_min B), a_min = ex_min / eb(i), ex_min = the smallest source-window
    exposure staged in the chunk (k_trig_eval, `s_need`).  Emulated in float32 like the kernel: a window it skips never
    fires (nor comes within 2 % of the threshold), and on a quiet light curve it skips nearly everything."""
    rs = np.random.RandomState(int(rate + burst) % 9967 + 1)
    nb = 9000
    counts, expo = _curve(rs, nb, rate, burst)
    cs = np.concatenate(([0], np.cumsum(counts)))
    es = np.concatenate(([0.0], np.cumsum(expo)))
    f32 = np.float32
    min_eb, min_ex = f32(0.25) * f32(N_BKG * DT), f32(0.25) * f32(N_SRC * DT)
    thr_lo = f32(0.94) * f32(thr)
    n_i_max = nb - (N_BKG + 1 + N_SRC)
    skipped = total = fired = 0
    for c0 in range(0, n_i_max, CH):
        hi = min(nb, c0 + CH + SPAN)
        n_st = hi - c0 + 1
        s_ef = (es[c0:hi + 1] - es[c0]).astype(f32)
        s_cf = (cs[c0:hi + 1] - cs[c0]).astype(f32)
        m = np.arange(n_st - N_SRC)
        sw = s_cf[m + N_SRC] - s_cf[m]                           # exact: integers below 2^24
        exw = s_ef[m + N_SRC] - s_ef[m]
        exm = exw[N_BKG:].min()
        assert exm > min_ex
        li = np.arange(0, min(c0 + CH, n_i_max) - c0)
        i = c0 + li
        Bf, ebf = s_cf[li + N_BKG] - s_cf[li], s_ef[li + N_BKG] - s_ef[li]
        pre = (Bf > 0) & (ebf > min_eb)
        with np.errstate(divide="ignore", invalid="ignore"):
            ab = (Bf / ebf) * exm
            s_need = np.where(pre, ab * f32(0.99999) + thr_lo * np.sqrt(ab), -np.inf).astype(f32)
        B, eb = (cs[i + N_BKG] - cs[i]).astype("f8"), es[i + N_BKG] - es[i]
        for gx in range(9, -1, -1):
            n_gap = 1 << gx
            ok = i + N_BKG + n_gap + N_SRC < nb
            ls, gs_ = (li + N_BKG + n_gap)[ok], (i + N_BKG + n_gap)[ok]
            skip = sw[ls] < s_need[ok]
            ex, S = es[gs_ + N_SRC] - es[gs_], (cs[gs_ + N_SRC] - cs[gs_]).astype("f8")
            with np.errstate(divide="ignore", invalid="ignore"):
                nbk = ex / eb[ok] * B[ok]
                sig = (S - nbk) / np.sqrt(nbk)
            assert not np.any(skip & (sig >= 0.98 * thr)), (gx, float(sig[skip].max()))
            skipped += int(skip.sum())
            fired += int((sig >= thr).sum())
            total += int(ok.sum())
    if burst == 0.0:
        assert skipped > (0.999 if thr >= 4.0 else 0.5) * total
    if burst >= 1500.0:
        assert fired > 0
