"""CPU: line-by-line Python restatements of three device-side algorithms whose equivalence to the serial reference
rule is an ARGUMENT in the kernels' comments, checked here against the oracle without a GPU:

* ``dt_keep`` (csrc/cgrb_deadtime.cuh): the literal GBM dead-time rule resolved per event from a short look-back
  through overflow events, against the serial loop of gbm_lightcurve.py:80-115 (oracle ``co_gbm_dead_time``);
* ``dt_keep_phys``: the physical (non-paralyzable) rule resolved per event from the nearest synchronisation point,
  against its serial definition (``co_gbm_dead_time_physical``);
* the conversion-free bin index of ``k_trig_bin`` (2^52 floor trick + an ulp-scaled safety margin, exact edge test
  otherwise) against ``np.histogram`` on ``np.arange`` edges (light_curve_storage.py:218-269).
"""
import numpy as np
import pytest

from oracle import oracle as orc

TAU, NORMAL = 10.6e-6, 2.6e-6


def dt_pe(t, p, j):
    return t[j] + (NORMAL if (j == 0 and p[0] != 127) else TAU)


def dt_keep(t, p, i):
    """csrc/cgrb_deadtime.cuh: dt_keep"""
    if i == 0:
        return True
    ti = t[i]
    found = -1
    j = i - 1
    while j >= 0:
        if t[j] + TAU < ti:
            break
        if p[j] == 127 or j == 0:
            found = j
            break
        j -= 1
    if found < 0:
        return True
    head = found
    while head > 0:
        th = t[head]
        q = -1
        j = head - 1
        while j >= 0:
            if t[j] + TAU < th:
                break
            if p[j] == 127 or j == 0:
                q = j
                break
            j -= 1
        if q < 0 or dt_pe(t, p, q) < th:
            break
        head = q
    t_end = dt_pe(t, p, head)
    for j in range(head + 1, i):
        if p[j] == 127 and t[j] > t_end:
            t_end = t[j] + TAU
    return ti > t_end


def dt_keep_phys(t, p, i, tau, tau_ovf):
    """csrc/cgrb_deadtime.cuh: dt_keep_phys (has0 = True, no walk limit)"""
    tau_max = max(tau, tau_ovf)
    s, ts, ps = i, t[i], p[i]
    while s > 0:
        tp = t[s - 1]
        if ts > tp + tau_max:
            break
        s -= 1
        ts, ps = tp, p[s]
    if s == i:
        return True
    t_end = ts + (tau_ovf if ps == 127 else tau)
    kept = True
    for j in range(s + 1, i + 1):
        kept = t[j] > t_end
        if kept:
            t_end = t[j] + (tau_ovf if p[j] == 127 else tau)
    return kept


def _events(rs, n, rate, frac127, ties=False):
    t = np.cumsum(rs.exponential(1.0 / rate, n))
    if ties:
        # adversarial: events exactly one window after another, and exact duplicates
        k = rs.randint(1, n - 1, n // 20)
        t[k] = t[k - 1] + rs.choice([TAU, NORMAL, 10e-6, 0.0], len(k))
        t = np.sort(t)
    p = rs.randint(0, 127, n)
    p[rs.random_sample(n) < frac127] = 127
    return t, p


@pytest.mark.parametrize("rate,frac127,ties", [(500.0, 0.05, False), (2e5, 0.05, True), (2e6, 0.3, True), (5e6, 1.0, False),
                                               (1e5, 0.0, True)])
def test_literal_dead_time_look_back_equals_the_serial_rule(rate, frac127, ties):
    rs = np.random.RandomState(int(rate) % 1000 + int(100 * frac127))
    n = 4000 if rate > 1e6 else 12000
    for first127 in (False, True):
        t, p = _events(rs, n, rate, frac127, ties)
        p[0] = 127 if first127 else 5
        want = orc.gbm_dead_time(t, p)
        got = np.array([dt_keep(t, p, i) for i in range(n)])
        assert np.array_equal(got, want)


@pytest.mark.parametrize("rate,frac127,tau,tovf", [(500.0, 0.05, 2.6e-6, 10e-6), (2e5, 0.1, 2.6e-6, 10e-6),
                                                   (5e5, 0.3, 2.6e-6, 10e-6), (1e5, 0.5, 5e-6, 1e-6), (2e5, 0.2, 0.0, 0.0)])
def test_physical_dead_time_sync_points_equal_the_serial_rule(rate, frac127, tau, tovf):
    rs = np.random.RandomState(int(rate) % 977 + 3)
    n = 12000
    t, p = _events(rs, n, rate, frac127, ties=True)
    want = orc.gbm_dead_time_physical(t, p, tau, tovf)
    got = np.array([dt_keep_phys(t, p, i, tau, tovf) for i in range(n)])
    assert np.array_equal(got, want)


def _bin_fast(tt, t0, delta, nb, last_edge):
    """k_trig_bin's fast path; returns (bin, took_fast_path)"""
    inv_delta = 1.0 / delta
    eps = (abs(t0) + abs(last_edge)) * inv_delta * (64.0 * 2.220446049250313e-16) + 1e-9
    fast_ok = eps < 0.25 and nb > 2
    MAGIC = 6755399441055744.0
    q = (tt - t0) * inv_delta
    m = q + MAGIC
    k = int(np.float64(m).view(np.uint64) & 0xFFFFFFFF)          # __double2loint
    if k >= 1 << 31:
        k -= 1 << 32
    fr = q - (m - MAGIC)
    if fr < 0.0:
        k -= 1
        fr += 1.0
    return k, bool(fast_ok and eps < fr < 1.0 - eps and 1.0 < q < float(nb - 1))


@pytest.mark.parametrize("t0,span", [(-100.0, 400.0), (0.0, 1000.0), (5.0e8, 400.0), (-3.3e9, 123.4), (1e-3, 50.0)])
def test_trigger_bin_index_fast_path_is_exact(t0, span):
    rs = np.random.RandomState(11)
    dt = 0.016
    tmax = t0 + span
    edges = np.arange(t0, tmax, dt)                               # the reference's bin edges
    nb = len(edges) - 1
    delta = (t0 + dt) - t0
    # np.arange's fill rule as restated in the kernel: edge[i] = t0 + i * delta
    assert np.array_equal(edges, t0 + np.arange(len(edges)) * delta)
    tt = np.concatenate((rs.uniform(t0, edges[-1], 20000),
                         edges[rs.randint(1, nb, 300)],                                  # exactly on an edge
                         np.nextafter(edges[rs.randint(1, nb, 300)], -np.inf),           # one ulp below / above
                         np.nextafter(edges[rs.randint(1, nb, 300)], np.inf)))
    want = np.clip(np.searchsorted(edges, tt, "right") - 1, 0, nb - 1)                   # edge[k] <= t < edge[k+1]
    n_fast = 0
    for x, w in zip(tt, want):
        k, fast = _bin_fast(float(x), t0, delta, nb, float(edges[-1]))
        if fast:
            n_fast += 1
            assert k == w, (x, k, w)
    # the fast path must carry nearly all random events unless the times are so large that an ulp is a
    # sizeable fraction of a bin (then everything takes the exact edge test)
    if abs(t0) < 1e9:
        assert n_fast > 0.95 * 20000


def merge_split(A, B, d):
    """csrc/cgrb_deadtime.cuh: merge_split -- number of background events among the first d merged events"""
    nA, nB = len(A), len(B)
    lo, hi = max(0, d - nB), min(d, nA)
    while lo < hi:
        mid = (lo + hi) >> 1
        if A[mid] <= B[d - 1 - mid]:        # ties: background first
            lo = mid + 1
        else:
            hi = mid
    return lo


def test_merge_path_tiles_equal_the_reference_combine():
    """k_merge_plan / k_merge_deadtime (md_merge_general): every tile of the merged order (+ its look-back halo) is rebuilt from the
    merge-path splits of its ends and per-thread serial merges of 9 elements; stitched together the tiles are the
    oracle's combine (lightcurve.py:131-151: stable two-way merge, background first on ties)."""
    rs = np.random.RandomState(2)
    TILE, HALO, NT = 2048, 64, 256
    PER = (TILE + HALO + NT - 1) // NT
    for nA, nB, ties in ((7000, 9000, True), (0, 5000, False), (5000, 0, False), (1, 1, True), (30000, 300, True)):
        A = np.sort(rs.uniform(0, 10, nA))
        B = np.sort(rs.uniform(0, 10, nB))
        if ties and nA > 10 and nB > 10:
            B[5:9] = A[3]
            A[100:103] = B[50] if nB > 50 else A[100:103]
            A.sort()
            B.sort()
        pA, pB = rs.randint(0, 128, nA), rs.randint(0, 128, nB)
        want_t, want_p = orc.combine(A, pA, B, pB)
        total = nA + nB
        got_t, got_p = np.empty(total), np.empty(total)
        for d0 in range(0, total, TILE):
            d1 = min(d0 + TILE, total)
            g0 = max(0, d0 - HALO)
            i0, i1 = merge_split(A, B, g0), (nA if d1 == total else merge_split(A, B, d1))
            j0, j1 = g0 - i0, d1 - i1
            a_, b_ = A[i0:i1], B[j0:j1]
            pa_, pb_ = pA[i0:i1], pB[j0:j1]
            na, nb = len(a_), len(b_)
            n = na + nb
            st, sp = np.empty(n), np.empty(n)
            for tid in range(NT):                   # one thread: PER consecutive outputs from its own split
                q0 = min(tid * PER, n)
                ia = merge_split(a_, b_, q0)
                ib = q0 - ia
                for r in range(PER):
                    q = q0 + r
                    if q >= n:
                        break
                    take_a = ia < na and (ib >= nb or a_[ia] <= b_[ib])
                    st[q], sp[q] = (a_[ia], pa_[ia]) if take_a else (b_[ib], pb_[ib])
                    ia += 1 if take_a else 0
                    ib += 0 if take_a else 1
            l0 = d0 - g0
            got_t[d0:d1], got_p[d0:d1] = st[l0:l0 + (d1 - d0)], sp[l0:l0 + (d1 - d0)]
            # the halo really is the merged order just before the tile
            assert np.array_equal(st[:l0], want_t[g0:d0])
        assert np.array_equal(got_t, want_t) and np.array_equal(got_p, want_p)


def test_background_channel_guide_equals_numpy_choice():
    """k_background: channel = searchsorted(cdf, u, 'right') clamped to the last channel (numpy's legacy choice,
    background.py:93), found from the detector table's 1024-entry guide (guide[m] = searchsorted(cdf, m / 1024,
    'right')) and a forward walk"""
    from tests import helpers as h
    from cosmogrb_b200 import _capi
    from cosmogrb_b200.response.response import background_guide

    rs = np.random.RandomState(9)
    N = _capi.BKG_GUIDE_N
    for det in ("n0", "b1"):
        w = np.asarray(h.bkg_weights(det)).astype("f8")
        cdf = w.cumsum()
        cdf /= cdf[-1]
        guide = background_guide(cdf)
        tab = h.response(det).device_table(h.bkg_weights(det), "linear")
        assert np.array_equal(tab[_capi.DT_BKG_GUIDE:_capi.DT_BKG_GUIDE + N // 8].view(np.uint8), guide)
        u = np.concatenate((rs.random_sample(50000), cdf[:-1], np.nextafter(cdf[:-1], 0), np.arange(N) / float(N),
                            [0.0, np.nextafter(1.0, 0)]))
        want = np.minimum(np.searchsorted(cdf, u, "right"), 127)
        got = np.empty(len(u), dtype=int)
        steps = 0
        for k, uc in enumerate(u):
            j = int(guide[min(max(int(uc * float(N)), 0), N - 1)])
            while j < 127 and not (uc < cdf[j]):
                j += 1
                steps += 1
            got[k] = j
        assert np.array_equal(got, want)
        assert steps < 0.5 * len(u)            # the walk rarely moves at all


# ---- k_merge_deadtime (round 2): chain elements mark their followers inside a tile with halo ------------------------------
def dt_keep_local(t, p, li, g0):
    """csrc/cgrb_deadtime.cuh: dt_keep_local on a tile held in shared memory (local index 0 <-> merged rank g0);
    0 / 1, or 2 when the walk leaves the halo"""
    if g0 + li == 0:
        return 1
    has0 = g0 == 0
    ti = t[li]
    found = -1
    j = li - 1
    while j >= 0:
        if t[j] + TAU < ti:
            return 1
        if p[j] == 127 or (has0 and j == 0):
            found = j
            break
        j -= 1
    if found < 0:
        return 2
    head = found
    while True:
        if has0 and head == 0:
            break
        th = t[head]
        q, gap = -1, False
        jj = head - 1
        while jj >= 0:
            if t[jj] + TAU < th:
                gap = True
                break
            if p[jj] == 127 or (has0 and jj == 0):
                q = jj
                break
            jj -= 1
        if q < 0:
            if gap:
                break
            return 2
        pe = t[q] + (NORMAL if (has0 and q == 0 and p[0] != 127) else TAU)
        if pe < th:
            break
        head = q
    t_end = t[head] + (NORMAL if (has0 and head == 0 and p[0] != 127) else TAU)
    for jj in range(head + 1, li):
        if p[jj] == 127 and t[jj] > t_end:
            t_end = t[jj] + TAU
    return 1 if ti > t_end else 0


def tile_keep_flags(t, p, d0, tile=2048, halo=64):
    """k_merge_deadtime<0>, step 2, for the tile of merged ranks [d0, d0 + tile): every chain element of the staged
    events (overflow channel, the unit's event 0, and the first staged event in place of whatever lies before the
    halo) marks the events that follow it by at most the longest window; only marked events take the exact walk
    (dt_keep_local; the global walk = dt_keep when the chain leaves the halo), everything else is kept.
    Returns (keep flags, number of marked events, number of global walks)."""
    n_all = len(t)
    g0 = max(0, d0 - halo)
    d1 = min(d0 + tile, n_all)
    T, P = t[g0:d1], p[g0:d1]
    n = len(T)
    l0 = d0 - g0
    has0 = g0 == 0
    rare = np.zeros(n, dtype=bool)
    for li in range(n):
        if P[li] == 127 or li == 0:            # li == 0: the unit's event 0 (has0) or the stand-in for the past
            te = T[li] + TAU
            k = li + 1
            while k < n and not (te < T[k]):
                if k >= l0:
                    rare[k] = True
                k += 1
    keep = np.ones(d1 - d0, dtype=bool)
    n_slow = 0
    for e in range(d1 - d0):
        li = l0 + e
        if not rare[li]:
            continue
        k = dt_keep_local(T, P, li, g0)
        if k == 2:
            n_slow += 1
            k = 1 if dt_keep(t, p, d0 + e) else 0
        keep[e] = bool(k)
    return keep, int(rare[l0:].sum()), n_slow


def level_major_order(n_tiles):
    """k_merge_order + k_merge_plan: ticket T -> (unit, level).  hist / above / cum / sorted exactly as the kernels
    build them (the rank of a unit among its equals is whatever the atomics gave: any order works)."""
    n_tiles = np.asarray(n_tiles, dtype=np.int64)
    levels = int(n_tiles.max()) if len(n_tiles) else 0
    hist = np.bincount(n_tiles, minlength=levels + 2)
    above = np.array([hist[k + 1:].sum() for k in range(levels + 1)], dtype=np.int64)    # units with n_u > k
    cum = np.concatenate(([0], np.cumsum(above[:levels])))
    eq = np.zeros(len(n_tiles), dtype=np.int64)
    seen = {}
    for u in np.random.RandomState(5).permutation(len(n_tiles)):      # arrival order of the atomics: arbitrary
        eq[u] = seen.get(int(n_tiles[u]), 0)
        seen[int(n_tiles[u])] = eq[u] + 1
    srt = np.full(int(above[0]) if levels else 0, -1, dtype=np.int64)
    for u, n in enumerate(n_tiles):
        if n > 0:
            srt[above[n] + eq[u]] = u
    out = []
    for T in range(int(cum[levels])):
        lo, hi = 0, levels
        while hi - lo > 1:
            mid = (lo + hi) >> 1
            if cum[mid] <= T:
                lo = mid
            else:
                hi = mid
        out.append((int(srt[T - cum[lo]]), lo))
    return out


@pytest.mark.parametrize("n_tiles", [[3, 1, 4, 1, 5, 0, 2, 6], [7], [0, 0, 2], [1] * 40, list(range(25)) + [100, 3, 3]])
def test_level_major_ticket_order_is_a_dependency_respecting_permutation(n_tiles):
    """every live (unit, tile) pair gets exactly one ticket, and tile k of a unit has a larger ticket than its
    tiles < k (the look-back only ever waits for smaller tickets); the tiles in flight spread over the units"""
    order = level_major_order(n_tiles)
    assert sorted(order) == sorted((u, k) for u, n in enumerate(n_tiles) for k in range(n))
    ticket = {uk: T for T, uk in enumerate(order)}
    for (u, k), T in ticket.items():
        if k > 0:
            assert ticket[(u, k - 1)] < T
    levels = [k for _, k in order]
    assert levels == sorted(levels)


@pytest.mark.parametrize("rate,frac127,ties", [(500.0, 0.05, False), (7.5e4, 0.05, True), (2e5, 0.3, True),
                                               (2e6, 0.3, True), (3e7, 0.02, False), (1e5, 0.0, True)])
def test_tile_chain_marking_equals_the_serial_rule(rate, frac127, ties):
    rs = np.random.RandomState(int(rate) % 991 + int(100 * frac127) + 1)
    n = 3 * 2048 + 700
    for first127 in (False, True):
        t, p = _events(rs, n, rate, frac127, ties)
        p[0] = 127 if first127 else 5
        want = orc.gbm_dead_time(t, p)
        got, n_rare, n_slow = [], 0, 0
        for d0 in range(0, n, 2048):
            k, r, s_ = tile_keep_flags(t, p, d0)
            got.append(k)
            n_rare += r
            n_slow += s_
        assert np.array_equal(np.concatenate(got), want)
        if rate <= 500.0:
            assert n_rare < 0.01 * n           # GBM background: the rare list is (nearly) empty
        if rate >= 3e7:
            assert n_slow > 0                  # more than 64 events inside one window: the halo runs out


# ---- k_merge_deadtime: the global walk (dt_keep_slow) on a sequential cursor over the two lists -----------------------
class MergedCursor:
    """csrc/cgrb_deadtime.cuh: MergedCursor -- position g: the merged prefix [0, g) is A[0, ia) + B[0, ib); ties:
    background (A) first, so the last element of a prefix is B[ib-1] when B[ib-1] >= A[ia-1]"""

    def __init__(self, A, pA, B, pB, g):
        self.A, self.pA, self.B, self.pB = A, pA, B, pB
        self.seek(g)

    def seek(self, g):
        self.ia = merge_split(self.A, self.B, g)
        self.ib = g - self.ia

    def pos(self):
        return self.ia + self.ib

    def prev(self):
        a_prev = self.A[self.ia - 1] if self.ia > 0 else -np.inf
        b_prev = self.B[self.ib - 1] if self.ib > 0 else -np.inf
        if self.ib > 0 and (self.ia == 0 or b_prev >= a_prev):
            self.ib -= 1
            return self.B[self.ib], self.pB[self.ib]
        self.ia -= 1
        return self.A[self.ia], self.pA[self.ia]

    def next(self):
        a_next = self.A[self.ia] if self.ia < len(self.A) else np.inf
        b_next = self.B[self.ib] if self.ib < len(self.B) else np.inf
        if self.ia < len(self.A) and (self.ib >= len(self.B) or a_next <= b_next):
            self.ia += 1
            return self.A[self.ia - 1], self.pA[self.ia - 1]
        self.ib += 1
        return self.B[self.ib - 1], self.pB[self.ib - 1]


def dt_keep_slow(A, pA, B, pB, i):
    """csrc/cgrb_deadtime.cuh: dt_keep_slow -- one backward pass to the head of the chain, one forward pass from it"""
    if i == 0:
        return True
    c = MergedCursor(A, pA, B, pB, i + 1)
    ti, _ = c.prev()
    found, tj, pj = -1, 0.0, 0
    while c.pos() > 0:
        tj, pj = c.prev()
        if tj + TAU < ti:
            break
        if pj == 127 or c.pos() == 0:
            found = c.pos()
            break
    if found < 0:
        return True
    head, th, ph = found, tj, pj
    while head > 0:
        q, tq, pq = -1, 0.0, 0
        while c.pos() > 0:
            tq, pq = c.prev()
            if tq + TAU < th:
                break
            if pq == 127 or c.pos() == 0:
                q = c.pos()
                break
        if q < 0 or tq + (NORMAL if (q == 0 and pq != 127) else TAU) < th:
            break
        head, th, ph = q, tq, pq
    t_end = th + (NORMAL if (head == 0 and ph != 127) else TAU)
    c.seek(head + 1)
    for _ in range(head + 1, i):
        tj, pj = c.next()
        if pj == 127 and tj > t_end:
            t_end = tj + TAU
    return ti > t_end


def test_merged_cursor_walks_the_reference_combine_both_ways():
    rs = np.random.RandomState(12)
    for nA, nB in ((300, 500), (0, 40), (40, 0), (1, 1), (700, 9)):
        A, B = np.sort(rs.randint(0, 60, nA) * 0.25), np.sort(rs.randint(0, 60, nB) * 0.25)    # many exact ties
        pA, pB = rs.randint(0, 128, nA), rs.randint(0, 128, nB)
        want_t, want_p = orc.combine(A, pA, B, pB)
        n = nA + nB
        for g in sorted(set([0, 1, n // 3, n // 2, n - 1, n])):
            c = MergedCursor(A, pA, B, pB, g)
            fwd = [c.next() for _ in range(g, n)]
            assert [x[0] for x in fwd] == list(want_t[g:]) and [x[1] for x in fwd] == list(want_p[g:])
            c = MergedCursor(A, pA, B, pB, g)
            back = [c.prev() for _ in range(g)]
            assert [x[0] for x in back] == list(want_t[:g][::-1]) and [x[1] for x in back] == list(want_p[:g][::-1])
            assert c.pos() == 0


@pytest.mark.parametrize("rate,frac127,ties", [(2e5, 0.3, True), (3e6, 0.02, True), (3e7, 0.02, False), (5e6, 1.0, False)])
def test_global_walk_on_the_cursor_equals_the_serial_rule(rate, frac127, ties):
    """the walk an event takes when its chain leaves the tile's halo: exact for every event, on two lists with ties"""
    rs = np.random.RandomState(int(rate) % 997 + 3)
    n = 2500
    for first127 in (False, True):
        t, p = _events(rs, n, rate, frac127, ties)
        p[0] = 127 if first127 else 5
        is_a = rs.random_sample(n) < 0.3
        # equal times: the reference's combine puts the background (A) events first
        order = np.lexsort((~is_a, t))
        t, p, is_a = t[order], p[order], is_a[order]
        A, pA, B, pB = t[is_a], p[is_a], t[~is_a], p[~is_a]
        ct, cp = orc.combine(A, pA, B, pB)
        assert np.array_equal(ct, t) and np.array_equal(cp, p)
        want = orc.gbm_dead_time(t, p)
        got = np.array([dt_keep_slow(A, pA, B, pB, i) for i in range(n)])
        assert np.array_equal(got, want)


def test_dense_tiles_take_the_long_halo_and_stay_exact():
    """k_merge_plan: a tile with more than MD_DENSE = 8 events to the longest window stages MD_HALO_DENSE = 512 events of
    look-back instead of 64.  Any halo gives the same keep flags; the long one keeps dense tiles off the global walk."""
    rs = np.random.RandomState(77)
    n = 4 * 2048
    t, p = _events(rs, n, 3e6, 0.02)            # ~32 events per window: the peak of a very bright burst
    want = orc.gbm_dead_time(t, p)
    slow = {}
    for halo in (64, 512):
        got, slow[halo] = [], 0
        for d0 in range(0, n, 2048):
            d1, g0 = min(d0 + 2048, n), max(0, d0 - 64)
            dense = d0 > 64 and (t[d1 - 1] - t[g0]) * 8 < (d1 - g0) * TAU       # the rule of k_merge_plan
            assert dense == (d0 > 0)
            k, _, s_ = tile_keep_flags(t, p, d0, halo=halo)
            got.append(k)
            slow[halo] += s_
        assert np.array_equal(np.concatenate(got), want)
    assert slow[64] >= 5 and slow[512] == 0
    tb, _ = _events(rs, n, 500.0, 0.05)         # GBM background: never dense
    assert not any((tb[min(d0 + 2048, n) - 1] - tb[d0 - 64]) * 8 < (min(d0 + 2048, n) - d0 + 64) * TAU for d0 in range(2048, n, 2048))


# ---------------------------------------------------------------------------------------------------------------
# k_sample_energy_env: the lane-refill scheme (csrc/cgrb_energy_kernels.cuh).  A CTA owns the photons of one work item,
# warp q the slice [q per, min((q + 1) per, len)) in 32-bit offsets from the item's first photon; a lane that finished a
# photon takes the warp's next one, trial (photon, j) draws Philox counter (photon, j) whatever lane runs it, and the
# trial counter of the unit is the number of active lanes summed over the loop's iterations.
def _refill_warp(wbeg, wend, trials_of, cap):
    """Python restatement of one warp's loop: returns {photon offset: trials spent}, and the warp's trial count."""
    nxt = wbeg
    active = [False] * 32
    di = [0] * 32
    j = [0] * 32
    done, warp_trials = {}, 0
    while True:
        need = [not a for a in active]
        if any(need):
            before = 0
            for lane in range(32):
                if need[lane]:
                    cand = nxt + before
                    before += 1
                    if cand < wend:
                        di[lane], j[lane], active[lane] = cand, 0, True
            nxt += before
        if not any(active):
            if nxt >= wend:
                break
            continue
        warp_trials += sum(active)
        for lane in range(32):
            if active[lane]:
                j[lane] += 1
                acc = j[lane] >= trials_of(di[lane])          # the trial that accepts this photon
                capped = (not acc) and j[lane] >= cap
                if acc or capped:
                    assert di[lane] not in done
                    done[di[lane]] = j[lane]
                    active[lane] = False
    return done, warp_trials


@pytest.mark.parametrize("length", [0, 1, 31, 32, 33, 255, 256, 257, 2049, 16384])
def test_energy_kernel_lane_refill_covers_every_photon_once(length):
    NW = 8
    rng = np.random.RandomState(length)
    need = rng.geometric(0.89, size=max(length, 1))            # trials a photon needs (1.12 on average, like C2)
    need[::97] = 50                                            # a few photons beyond the cap
    cap = 20
    per = (length + NW - 1) // NW
    seen, total = {}, 0
    for wid in range(NW):
        wbeg, wend = wid * per, min(wid * per + per, length)
        done, wt = _refill_warp(wbeg, wend, lambda d: need[d], cap)
        assert all(wbeg <= d < wend for d in done)
        assert not (set(done) & set(seen))
        seen.update(done)
        total += wt
    assert sorted(seen) == list(range(length))                 # every photon of the item, exactly once
    want = np.minimum(need[:length], cap)
    assert [seen[d] for d in range(length)] == list(want)      # trials per photon do not depend on the lane
    assert total == int(want.sum())                            # per-warp popcount of active lanes == sum of trial counts
