// cgrb_deadtime.cuh -- combine (merge) and GBM dead time: the unfused pair and the one-pass kernel with decoupled look-back
// Part of the single translation unit cgrb_kernels.cu (included in order; see that file).
#pragma once

// ==========================================================================================
// combine (lightcurve.py:131-151): merge-path tiles; inside a tile each element finds its rank
// by a binary search in the other list's slice held in shared memory.
// ==========================================================================================
__global__ void __launch_bounds__(NT) k_combine(const cgrb_unit_t *__restrict__ units,
                                                const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                const double *__restrict__ times_bkg,
                                                const uint8_t *__restrict__ pha_bkg,
                                                const double *__restrict__ times_src,
                                                const uint8_t *__restrict__ pha_src,
                                                double *__restrict__ times_tot, uint8_t *__restrict__ pha_tot,
                                                cgrb_counts_t *counts)
{
    __shared__ double s_t[CGRB_MERGE_TILE];
    __shared__ uint8_t s_p[CGRB_MERGE_TILE];
    __shared__ int64_t s_split[2];
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t nA = counts[wk.unit].n_bkg, nB = counts[wk.unit].n_src;
    const int64_t total = nA + nB;
    if (wk.start == 0 && threadIdx.x == 0)
        counts[wk.unit].n_total = total;
    if (wk.start >= total)
        return;
    const double *A = times_bkg + u->bkg_off;
    const double *B = times_src + u->src_off;
    const int64_t d0 = wk.start, d1 = min(d0 + (int64_t)CGRB_MERGE_TILE, total);
    if (threadIdx.x < 2) {
        const int64_t d = threadIdx.x ? d1 : d0;
        int64_t lo = max((int64_t)0, d - nB), hi = min(d, nA);
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (A[mid] <= B[d - 1 - mid])     // ties: background first
                lo = mid + 1;
            else
                hi = mid;
        }
        s_split[threadIdx.x] = lo;
    }
    __syncthreads();
    const int64_t i0 = s_split[0], i1 = s_split[1];
    const int64_t j0 = d0 - i0, j1 = d1 - i1;
    const int na = (int)(i1 - i0), nb = (int)(j1 - j0);
    for (int k = threadIdx.x; k < na; k += NT) {
        s_t[k] = A[i0 + k];
        s_p[k] = pha_bkg[u->bkg_off + i0 + k];
    }
    for (int k = threadIdx.x; k < nb; k += NT) {
        s_t[na + k] = B[j0 + k];
        s_p[na + k] = pha_src[u->src_off + j0 + k];
    }
    __syncthreads();
    double *out_t = times_tot + u->tot_off + d0;
    uint8_t *out_p = pha_tot + u->tot_off + d0;
    for (int k = threadIdx.x; k < na + nb; k += NT) {
        const double t = s_t[k];
        int rank;
        if (k < na) {
            // number of source events strictly earlier
            int lo = 0, hi = nb;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (s_t[na + mid] < t)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            rank = k + lo;
        } else {
            // number of background events earlier or equal
            int lo = 0, hi = na;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (t < s_t[mid])
                    hi = mid;
                else
                    lo = mid + 1;
            }
            rank = (k - na) + lo;
        }
        out_t[rank] = t;
        out_p[rank] = s_p[k];
    }
}

// ==========================================================================================
// GBM dead time, literal reference semantics (gbm_lightcurve.py:80-115).
//   event 0 is kept and opens a window (10.6 us if overflow channel, else 2.6 us);
//   afterwards an event is kept iff t > t_end, and ONLY a kept overflow event moves t_end.
// The chain of dependence runs only through overflow events closer than 10.6 us to each other,
// so each event resolves its keep flag from a short local look-back (exact for any input).
// ==========================================================================================
#define DT_TAU 10.6e-6
#define DT_NORMAL 2.6e-6

__device__ __forceinline__ double dt_pe(const double *t, const uint8_t *p, int64_t j)
{
    return t[j] + ((j == 0 && p[0] != 127) ? DT_NORMAL : DT_TAU);
}

__device__ __forceinline__ bool dt_keep(const double *__restrict__ t, const uint8_t *__restrict__ p, int64_t i)
{
    if (i == 0)
        return true;
    const double ti = t[i];
    // nearest earlier chain element (overflow event or event 0) whose window could cover ti
    int64_t found = -1;
    for (int64_t j = i - 1; j >= 0; j--) {
        if (t[j] + DT_TAU < ti)
            break;
        if (p[j] == 127 || j == 0) {
            found = j;
            break;
        }
    }
    if (found < 0)
        return true;
    // walk back to a chain element that is certainly kept
    int64_t head = found;
    while (head > 0) {
        const double th = t[head];
        int64_t q = -1;
        for (int64_t j = head - 1; j >= 0; j--) {
            if (t[j] + DT_TAU < th)
                break;
            if (p[j] == 127 || j == 0) {
                q = j;
                break;
            }
        }
        if (q < 0 || dt_pe(t, p, q) < th)
            break;          // nothing before `head` can veto it
        head = q;
    }
    // replay the reference loop from the head
    double t_end = dt_pe(t, p, head);
    for (int64_t j = head + 1; j < i; j++) {
        if (p[j] == 127) {
            const double tj = t[j];
            if (tj > t_end)
                t_end = tj + DT_TAU;
        }
    }
    return ti > t_end;
}

// ==========================================================================================
// GBM dead time, physical (non-paralyzable) model -- SURVEY.md §8 f4; NOT what the reference computes.
//   event 0 is kept; every KEPT event opens a window (tau_overflow if pha == 127, else tau) and an
//   event is kept iff t > t_end of the last kept event.
// The chain runs through every kept event, but it re-synchronises at any event that follows its
// predecessor by more than the longest window: fl(t[s-1] + tau_max) >= any t_end set before s (rounding is
// monotonic), so such an event is certainly kept.  Each event walks back to the nearest synchronisation
// point and replays the serial rule from there (exact for any input; runs are short because the mean gap
// is many windows long).  A walk that leaves the tile's halo (fused kernel: 256 events) or exceeds
// DT_PHYS_MAX_WALK events (event rate x longest window >~ 3: far beyond anything GBM records) flags the unit
// (CGRB_ST_DEADTIME_CHAIN) for the serial fix-up kernel at the end of this file instead of running for O(n^2).
// ==========================================================================================
#define DT_PHYS_MAX_WALK 1024
struct DtModel {
    double tau, tau_ovf, tau_max;
};

struct EventArrays {
    const double *t;
    const uint8_t *p;
    __device__ __forceinline__ void get(int64_t j, double *tt, uint8_t *pp) const
    {
        *tt = t[j];
        *pp = p[j];
    }
};

// keep flag of event i on accessor `m`, whose lowest readable index is 0; `has0` tells whether index 0 is
// the unit's event 0.  Returns 0 / 1, or 2 if the walk would leave the readable range or exceed
// DT_PHYS_MAX_WALK events.
template <class Acc>
__device__ __forceinline__ int dt_keep_phys(const Acc &m, int64_t i, bool has0, const DtModel md)
{
    double ts, tp;
    uint8_t ps, pp;
    m.get(i, &ts, &ps);
    int64_t s = i;
    for (;;) {
        if (s == 0) {
            if (!has0)
                return 2;
            break;
        }
        if (i - s >= DT_PHYS_MAX_WALK)
            return 2;
        m.get(s - 1, &tp, &pp);
        if (ts > tp + md.tau_max)
            break;                          // s is a synchronisation point: certainly kept
        s--;
        ts = tp;
        ps = pp;
    }
    if (s == i)
        return 1;
    double t_end = ts + (ps == 127 ? md.tau_ovf : md.tau);
    bool kept = true;
    for (int64_t j = s + 1; j <= i; j++) {
        m.get(j, &tp, &pp);
        kept = tp > t_end;
        if (kept)
            t_end = tp + (pp == 127 ? md.tau_ovf : md.tau);
    }
    return kept ? 1 : 0;
}

// 0 / 1, or 2 = refused (physical mode only)
template <int MODE>
__device__ __forceinline__ int dt_keep_any(const double *__restrict__ t, const uint8_t *__restrict__ p, int64_t i,
                                           const DtModel md)
{
    if (MODE == 0)
        return dt_keep(t, p, i) ? 1 : 0;
    EventArrays ev{t, p};
    return dt_keep_phys(ev, i, true, md);
}

template <int MODE>
__global__ void __launch_bounds__(NT) k_deadtime_count(const cgrb_unit_t *__restrict__ units,
                                                       const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                       const double *__restrict__ times_tot,
                                                       const uint8_t *__restrict__ pha_tot,
                                                       int32_t *__restrict__ tile_count,
                                                       const cgrb_counts_t *counts, const DtModel md)
{
    __shared__ int s_wi[NW];
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t n = counts[wk.unit].n_total;
    int cnt = 0;
    if (wk.start < n) {
        const double *t = times_tot + u->tot_off;
        const uint8_t *p = pha_tot + u->tot_off;
        const int64_t end = min(wk.start + (int64_t)CGRB_DT_TILE, n);
        for (int64_t i = wk.start + threadIdx.x; i < end; i += NT)
            cnt += dt_keep_any<MODE>(t, p, i, md) == 1 ? 1 : 0;
    }
    int total;
    (void)block_incl_scan_i(cnt, s_wi, &total);
    if (threadIdx.x == 0)
        tile_count[w] = total;
}

template <int MODE>
__global__ void __launch_bounds__(NT) k_deadtime_write(const cgrb_unit_t *__restrict__ units,
                                                       const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                       const int64_t *__restrict__ unit_work_begin,
                                                       const double *__restrict__ times_tot,
                                                       const uint8_t *__restrict__ pha_tot,
                                                       const int32_t *__restrict__ tile_count,
                                                       double *__restrict__ times_out, uint8_t *__restrict__ pha_out,
                                                       uint8_t *__restrict__ selection, cgrb_counts_t *counts,
                                                       const DtModel md)
{
    __shared__ long long s_wl[NW];
    __shared__ int s_wi[NW];
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t n = counts[wk.unit].n_total;
    const int64_t b = unit_work_begin[wk.unit], e = unit_work_begin[wk.unit + 1];
    long long part = 0;
    for (int64_t s = b + threadIdx.x; s < w; s += NT)
        part += tile_count[s];
    long long off = block_sum_ll(part, s_wl);
    if (w == e - 1 && threadIdx.x == 0)
        counts[wk.unit].n_kept = off + tile_count[w];
    if (wk.start >= n)
        return;
    const double *t = times_tot + u->tot_off;
    const uint8_t *p = pha_tot + u->tot_off;
    double *ot = times_out + u->tot_off;
    uint8_t *op = pha_out + u->tot_off;
    const int64_t end = min(wk.start + (int64_t)CGRB_DT_TILE, n);
    for (int64_t base = wk.start; base < end; base += NT) {
        const int64_t i = base + threadIdx.x;
        bool keep = false;
        double ti = 0.0;
        uint8_t pi = 0;
        if (i < end) {
            const bool flagged = (MODE != 0) && ((*(volatile uint32_t *)&counts[wk.unit].status) & CGRB_ST_DEADTIME_CHAIN);
            const int k3 = flagged ? 0 : dt_keep_any<MODE>(t, p, i, md);
            if (MODE != 0 && k3 == 2)
                atomicOr(&counts[wk.unit].status, CGRB_ST_DEADTIME_CHAIN);
            keep = k3 == 1;
            ti = t[i];
            pi = p[i];
            if (selection)
                selection[u->tot_off + i] = keep ? 1 : 0;
        }
        int total;
        int incl = block_incl_scan_i(keep ? 1 : 0, s_wi, &total);
        if (keep) {
            ot[off + incl - 1] = ti;
            op[off + incl - 1] = pi;
        }
        off += total;
    }
}

// ==========================================================================================
// combine + dead time in ONE pass (lightcurve.py:131-151 then gbm_lightcurve.py:42-56): the merged
// pre-dead-time list is not a product of LightCurve.process (light_curve_storage.py:16-80 keeps the two
// component lists and the filtered total), so it never has to exist in HBM.  Each CTA merges one tile
// of the merged order (plus a look-back halo) in shared memory, resolves the keep flags there with the
// same literal rule as dt_keep, and appends the survivors to the unit's output at an offset obtained by a
// decoupled look-back over the preceding tiles' survivor counts (integers: deterministic).
// 9 B read + <= 9 B written per event.
// ==========================================================================================
// merge-path split of diagonal d: number of background events among the first d merged events
__device__ __forceinline__ int64_t merge_split(const double *__restrict__ A, int64_t nA, const double *__restrict__ B,
                                               int64_t nB, int64_t d)
{
    int64_t lo = max((int64_t)0, d - nB), hi = min(d, nA);
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (A[mid] <= B[d - 1 - mid])     // ties: background first
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// merged event g read straight from the two global lists (O(log n) per access: the rare slow path)
struct MergedGlobal {
    const double *A, *B;
    const uint8_t *pA, *pB;
    int64_t nA, nB;
    __device__ __forceinline__ void get(int64_t g, double *t, uint8_t *p) const
    {
        const int64_t ia = merge_split(A, nA, B, nB, g), ib = g - ia;
        if (ia < nA && (ib >= nB || A[ia] <= B[ib])) {
            *t = A[ia];
            *p = pA[ia];
        } else {
            *t = B[ib];
            *p = pB[ib];
        }
    }
};

// Sequential cursor over the merged order of the two global lists: O(1) per step (one load: only the side that
// moved is refilled) instead of a merge-path search per access.  Position g: the merged prefix [0, g) consists of
// A[0, ia) and B[0, ib); ties: background (A) first, so the LAST element of a prefix is B[ib-1] when B[ib-1] >= A[ia-1].
struct MergedCursor {
    const double *A, *B;
    const uint8_t *pA, *pB;
    int64_t nA, nB, ia, ib;
    double a_prev, b_prev;      // A[ia-1], B[ib-1] (-inf when there is none)
    double a_next, b_next;      // A[ia], B[ib] (+inf when there is none)
    __device__ __forceinline__ void seek(const MergedGlobal &m, int64_t g)
    {
        A = m.A, B = m.B, pA = m.pA, pB = m.pB, nA = m.nA, nB = m.nB;
        ia = merge_split(A, nA, B, nB, g);
        ib = g - ia;
        a_prev = ia > 0 ? A[ia - 1] : -INFINITY;
        b_prev = ib > 0 ? B[ib - 1] : -INFINITY;
        a_next = ia < nA ? A[ia] : INFINITY;
        b_next = ib < nB ? B[ib] : INFINITY;
    }
    __device__ __forceinline__ int64_t pos() const { return ia + ib; }
    // step back: returns merged event pos() - 1 and moves there (caller guarantees pos() > 0)
    __device__ __forceinline__ void prev(double *t, uint8_t *p)
    {
        if (ib > 0 && (ia == 0 || b_prev >= a_prev)) {
            ib--;
            *t = b_prev;
            *p = pB[ib];
            b_next = b_prev;
            b_prev = ib > 0 ? B[ib - 1] : -INFINITY;
        } else {
            ia--;
            *t = a_prev;
            *p = pA[ia];
            a_next = a_prev;
            a_prev = ia > 0 ? A[ia - 1] : -INFINITY;
        }
    }
    // step forward: returns merged event pos() and moves past it (caller guarantees pos() < nA + nB)
    __device__ __forceinline__ void next(double *t, uint8_t *p)
    {
        if (ia < nA && (ib >= nB || a_next <= b_next)) {
            *t = a_next;
            *p = pA[ia];
            ia++;
            a_prev = a_next;
            a_next = ia < nA ? A[ia] : INFINITY;
        } else {
            *t = b_next;
            *p = pB[ib];
            ib++;
            b_prev = b_next;
            b_next = ib < nB ? B[ib] : INFINITY;
        }
    }
};

// the literal dead-time rule for merged event i (same walk as dt_keep) on the global lists: one backward pass to
// the head of the chain -- the chain elements met on the way are exactly the overflow events between the head and
// i, because every inner walk stops at the first one -- then one forward pass from the head
__device__ __noinline__ bool dt_keep_slow(const MergedGlobal &m, int64_t i)
{
    if (i == 0)
        return true;
    MergedCursor c;
    c.seek(m, i + 1);
    double ti, tj = 0.0;
    uint8_t pi, pj = 0;
    c.prev(&ti, &pi);                       // event i; the cursor now stands at i
    int64_t found = -1;
    while (c.pos() > 0) {
        c.prev(&tj, &pj);
        if (tj + DT_TAU < ti)
            break;
        if (pj == 127 || c.pos() == 0) {
            found = c.pos();
            break;
        }
    }
    if (found < 0)
        return true;
    int64_t head = found;
    double th = tj;
    uint8_t ph = pj;
    while (head > 0) {
        int64_t q = -1;
        double tq = 0.0;
        uint8_t pq = 0;
        while (c.pos() > 0) {
            c.prev(&tq, &pq);
            if (tq + DT_TAU < th)
                break;
            if (pq == 127 || c.pos() == 0) {
                q = c.pos();
                break;
            }
        }
        if (q < 0 || tq + ((q == 0 && pq != 127) ? DT_NORMAL : DT_TAU) < th)
            break;
        head = q;
        th = tq;
        ph = pq;
    }
    double t_end = th + ((head == 0 && ph != 127) ? DT_NORMAL : DT_TAU);
    c.seek(m, head + 1);
    for (int64_t j = head + 1; j < i; j++) {
        c.next(&tj, &pj);
        if (pj == 127 && tj > t_end)
            t_end = tj + DT_TAU;
    }
    return ti > t_end;
}

// the same rule on the tile in shared memory: local index li <-> merged rank g0 + li.  Returns 0 / 1, or 2
// if the walk would leave the tile's halo (the caller then takes the slow path).
__device__ __forceinline__ int dt_keep_local(const double *t, const uint8_t *p, int li, int64_t g0)
{
    if (g0 + li == 0)
        return 1;
    const bool has0 = (g0 == 0);            // local index 0 is the unit's event 0
    const double ti = t[li];
    int found = -1;
    int j = li - 1;
    for (; j >= 0; j--) {
        if (t[j] + DT_TAU < ti)
            return 1;                       // nothing within the longest window: kept
        if (p[j] == 127 || (has0 && j == 0)) {
            found = j;
            break;
        }
    }
    if (found < 0)
        return 2;                           // ran out of halo
    int head = found;
    for (;;) {
        if (has0 && head == 0)
            break;
        const double th = t[head];
        int q = -1;
        int jj = head - 1;
        bool gap = false;
        for (; jj >= 0; jj--) {
            if (t[jj] + DT_TAU < th) {
                gap = true;
                break;
            }
            if (p[jj] == 127 || (has0 && jj == 0)) {
                q = jj;
                break;
            }
        }
        if (q < 0) {
            if (gap)
                break;                      // nothing before `head` can veto it
            return 2;                       // ran out of halo
        }
        const double pe = t[q] + ((has0 && q == 0 && p[0] != 127) ? DT_NORMAL : DT_TAU);
        if (pe < th)
            break;
        head = q;
    }
    double t_end = t[head] + ((has0 && head == 0 && p[0] != 127) ? DT_NORMAL : DT_TAU);
    for (int jj = head + 1; jj < li; jj++)
        if (p[jj] == 127) {
            const double tj = t[jj];
            if (tj > t_end)
                t_end = tj + DT_TAU;
        }
    return ti > t_end ? 1 : 0;
}

#define MD_TILE 2048
#ifndef MD_HALO
#define MD_HALO 64              // look-back events staged in front of a tile (a multiple of 32)
#endif
#ifndef MD_HALO_DENSE
#define MD_HALO_DENSE 512       // the same for a tile whose events come more than MD_DENSE to the longest window
#endif
#define MD_DENSE 8
#define MD_N (MD_TILE + MD_HALO_DENSE)

// Per-tile records, one thread per ticket: the ~23 dependent probes of each merge-path search are hidden by the
// other tiles' searches, and a CTA of the one-pass kernel gets everything about its tile from ONE 64-byte record.
struct MdTile {
    int64_t a_off, b_off;       // first staged event (halo start) in the flat background / source arrays
    int64_t wb;                 // first tile of the unit in the state array (look-back bound)
    int64_t d0;                 // merged rank of the tile's first event
    int64_t tot_off;            // the unit's slice of the output arrays
    int32_t na, nb;             // staged events of each list: halo + tile
    int32_t nt;                 // the tile's own events
    int32_t unit;
    uint32_t flags;             // MDT_*
    int32_t l0;                 // look-back events staged in front of the tile (0, MD_HALO or MD_HALO_DENSE)
};
static_assert(sizeof(MdTile) == 64, "MdTile is one 64-byte record");
#define MDT_LIVE 1u
#define MDT_HAS0 2u             // the staged range starts at the unit's event 0
#define MDT_LAST 4u             // the unit's last live tile: writes n_kept

__global__ void __launch_bounds__(NT) k_merge_plan(const cgrb_unit_t *__restrict__ units, int64_t n_work,
                                                   const int64_t *__restrict__ unit_work_begin,
                                                   const double *__restrict__ times_bkg,
                                                   const double *__restrict__ times_src, const LevelOrder o,
                                                   MdTile *__restrict__ tiles, const cgrb_counts_t *__restrict__ counts)
{
    const int64_t T = (int64_t)blockIdx.x * NT + threadIdx.x;
    if (T >= n_work)
        return;
    MdTile t;
    t.a_off = t.b_off = t.wb = t.d0 = t.tot_off = 0;
    t.na = t.nb = t.nt = t.unit = t.l0 = 0;
    t.flags = 0u;
    int unit, lo;
    if (level_ticket(o, T, &unit, &lo)) {
        const cgrb_unit_t *u = units + unit;
        const int64_t nA = counts[unit].n_bkg, nB = counts[unit].n_src;
        const int64_t total = nA + nB;
        const int64_t d0 = (int64_t)lo * MD_TILE;
        const double *A = times_bkg + u->bkg_off;
        const double *B = times_src + u->src_off;
        const int64_t d1 = min(d0 + (int64_t)MD_TILE, total);
        int64_t g0 = max((int64_t)0, d0 - MD_HALO);             // first merged rank held in shared memory
        int64_t i0 = merge_split(A, nA, B, nB, g0);
        const int64_t i1 = (d1 == total) ? nA : merge_split(A, nA, B, nB, d1);
        int64_t j0 = g0 - i0;
        if (d0 > MD_HALO) {
            // A dense tile -- more than MD_DENSE events to the longest dead-time window, the peak of a very bright
            // burst -- gets a longer halo: its chains of overflow events reach back hundreds of events, and a walk
            // that leaves the halo goes to the global lists (exact, but slow).  Speed only: any halo is correct.
            const int64_t j1 = d1 - i1;
            const double t_first = fmin(i0 < nA ? A[i0] : INFINITY, j0 < nB ? B[j0] : INFINITY);
            const double t_last = fmax(i1 > 0 ? A[i1 - 1] : -INFINITY, j1 > 0 ? B[j1 - 1] : -INFINITY);
            if ((t_last - t_first) * (double)MD_DENSE < (double)(d1 - g0) * DT_TAU) {
                g0 = max((int64_t)0, d0 - MD_HALO_DENSE);
                i0 = merge_split(A, nA, B, nB, g0);
                j0 = g0 - i0;
            }
        }
        t.a_off = u->bkg_off + i0;
        t.b_off = u->src_off + j0;
        t.wb = unit_work_begin[unit];
        t.d0 = d0;
        t.tot_off = u->tot_off;
        t.na = (int32_t)(i1 - i0);
        t.nb = (int32_t)((d1 - i1) - j0);
        t.nt = (int32_t)(d1 - d0);
        t.l0 = (int32_t)(d0 - g0);
        t.unit = unit;
        t.flags = MDT_LIVE | (g0 == 0 ? MDT_HAS0 : 0u) | (d1 == total ? MDT_LAST : 0u);
    }
    tiles[T] = t;
}

#define MD_PER ((MD_N + NT - 1) / NT)   // merged elements per thread in the merges (9)
#define MD_R (MD_TILE / NT)             // tile elements per thread (8)
#define MD_WORDS (MD_N / 32)            // words of a bitmap over the staged events (66)

// ------------------------------------------------------------------------------------------
// bulk-async (TMA) staging of a contiguous slice of a global list into shared memory.  cp.async.bulk moves
// multiples of 16 bytes between 16-byte aligned addresses, a slice starts anywhere: element k of the slice lands at
// dst[k] with dst chosen by the caller so that (shared address of dst[k]) == (global address of src[k]) mod 16;
// the aligned middle part goes through the bulk engine (one instruction, no registers, no LSU traffic), the
// < 16-byte head and tail are plain loads by the first threads.  Nothing outside [src, src + n) is read.
// ------------------------------------------------------------------------------------------
struct BulkPlan {
    int head, body;         // elements before the aligned part, elements in it (the tail is the rest)
};
template <int ELEM>
__device__ __forceinline__ BulkPlan bulk_plan(const void *src, int n)
{
    constexpr int PER16 = 16 / ELEM;
    BulkPlan pl;
    const int mis = (int)(((uintptr_t)src & 15u) / ELEM);
    pl.head = min(n, (PER16 - mis) & (PER16 - 1));
    pl.body = ((n - pl.head) / PER16) * PER16;
    return pl;
}
__device__ __forceinline__ void bulk_issue(void *dst, const void *src, unsigned bytes, uint64_t *s_bar)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(s_bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// bounded wait on phase `parity` of an mbarrier; false if it never completed (reported, not hung)
__device__ __forceinline__ bool mbar_wait(uint64_t *s_bar, unsigned parity)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(s_bar);
    unsigned long long t_begin = 0;
    for (int spin = 0;; spin++) {
        if ((spin & 255) == 255) {
            const unsigned long long now = global_timer_ns();
            if (t_begin == 0)
                t_begin = now;
            else if (now - t_begin > LB_WAIT_NS)
                return false;
        }
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done)
            return true;
    }
}

#ifdef MD_TIMING
#define MDT_DECL(n) long long tacc[n] = {0}; long long tlast = clock64();
#define MDT_MARK(i) do { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } while (0)
#else
#define MDT_DECL(n)
#define MDT_MARK(i)
#endif
#define MD_SPARSE 64             // a tile whose minority list has at most this many events is merged by insertion
#ifndef MD_CTAS
#define MD_CTAS 5                // resident CTAs per SM the kernel is compiled for (register budget; ~20 KB of shared memory each)
#endif

struct MdSmem {
    alignas(16) double t[MD_N + 4];         // slack: the alignment shift of up to two staged slices
    alignas(16) uint8_t p[MD_N + 48];
    uint32_t rare[MD_WORDS + 2];            // bit li: staged event li needs the exact chain walk
    uint16_t ins[MD_SPARSE];                // insertion merge: majority events preceding minority event i
    uint8_t cblk[MD_WORDS + 6];             // insertion merge: minority events inserted before majority index 32 b
    int wtot[NW];                           // survivors per warp
    long long excl;
    alignas(8) uint64_t full;               // the staged slices have landed
};

__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)),
                 "r"(bytes)
                 : "memory");
}
// The two in-place merges of a staged tile (slice A at t[ta ..] / p[qa ..], slice B at t[tb ..] / p[qb ..]) into
// merged order at t[0 ..] / p[0 ..]; out of line so that their register arrays do not weigh on the main path.
// Called by every thread of the CTA (they contain barriers).
__device__ __noinline__ void md_merge_insertion(MdSmem &sm, int na, int nb, int ta, int tb, int qa, int qb)
{
    const int tid = threadIdx.x, n = na + nb;
    // insertion merge: minority M (nm events) into majority J (nj events).  Ties: background first, so a
    // background event precedes the source events >= it, a source event follows the background events <= it.
    const bool minor_a = na <= nb;
    const int nm = minor_a ? na : nb, nj = n - nm;
    const double *M = sm.t + (minor_a ? ta : tb), *J = sm.t + (minor_a ? tb : ta);
    const uint8_t *pM = sm.p + (minor_a ? qa : qb), *pJ = sm.p + (minor_a ? qb : qa);
    double rt[MD_PER], mt = 0.0;
    uint8_t rp[MD_PER], mp = 0;
    int mdest = 0;
#pragma unroll
    for (int r = 0; r < MD_PER; r++) {
        if (r * NT >= nj)                       // (uniform: the last round only exists under the long halo)
            break;
        const int k = r * NT + tid;
        rt[r] = (k < nj) ? J[k] : 0.0;
        rp[r] = (k < nj) ? pJ[k] : 0;
    }
    if (tid < nm) {
        mt = M[tid];
        mp = pM[tid];
        int lo = 0, hi = nj;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double v = J[mid];
            if (minor_a ? (v < mt) : (v <= mt))
                lo = mid + 1;
            else
                hi = mid;
        }
        sm.ins[tid] = (uint16_t)lo;
        mdest = lo + tid;
    }
    __syncthreads();
    if (tid <= (nj + 31) / 32) {
        int c = 0, hi = nm;                         // minority events inserted before majority index 32 tid
        while (c < hi) {
            const int mid = (c + hi) >> 1;
            if ((int)sm.ins[mid] < 32 * tid)
                c = mid + 1;
            else
                hi = mid;
        }
        sm.cblk[tid] = (uint8_t)c;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MD_PER; r++) {
        if (r * NT >= nj)
            break;
        const int k = r * NT + tid;
        if (k < nj) {
            const int c1 = sm.cblk[(k >> 5) + 1];
            int c = sm.cblk[k >> 5];
            for (int i = c; i < c1; i++)
                c += ((int)sm.ins[i] <= k) ? 1 : 0;
            sm.t[k + c] = rt[r];
            sm.p[k + c] = rp[r];
        }
    }
    if (tid < nm) {
        sm.t[mdest] = mt;
        sm.p[mdest] = mp;
    }
    __syncthreads();
}

__device__ __noinline__ void md_merge_general(MdSmem &sm, int na, int nb, int ta, int tb, int qa, int qb)
{
    const int tid = threadIdx.x, n = na + nb;
    // serial merge of MD_PER consecutive outputs per thread from its merge-path split, in place through
    // registers
    const double *a_ = sm.t + ta, *b_ = sm.t + tb;
    const uint8_t *pa_ = sm.p + qa, *pb_ = sm.p + qb;
    const int per = (n + NT - 1) / NT;          // <= MD_PER; 9 under the short halo
    const int q0 = min(tid * per, n);
    int lo = max(0, q0 - nb), hi = min(q0, na);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a_[mid] <= b_[q0 - 1 - mid])     // ties: background first
            lo = mid + 1;
        else
            hi = mid;
    }
    int ia = lo, ib = q0 - lo;
    double rt[MD_PER];
    uint8_t rp[MD_PER];
#pragma unroll
    for (int r = 0; r < MD_PER; r++) {
        rt[r] = 0.0;
        rp[r] = 0;
        if (r >= per)
            break;
        if (q0 + r < n) {
            const bool take_a = ia < na && (ib >= nb || a_[ia] <= b_[ib]);
            rt[r] = take_a ? a_[ia] : b_[ib];
            rp[r] = take_a ? pa_[ia] : pb_[ib];
            ia += take_a ? 1 : 0;
            ib += take_a ? 0 : 1;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MD_PER; r++)
        if (r < per && q0 + r < n) {
            sm.t[q0 + r] = rt[r];
            sm.p[q0 + r] = rp[r];
        }
    __syncthreads();
}

// One CTA per tile of MD_TILE merged events; blockIdx.x is the tile's TICKET in the level-major order above and
// everything about the tile comes from ONE 64-byte record written by k_merge_plan, so the dependence chain in front
// of the data is record -> staging.  Several small CTAs per SM (MD_CTAS, ~20 KB of shared memory each) overlap one
// tile's latencies with the other tiles' work.  Forward progress of the look-back needs every tile with a smaller
// ticket to be running or finished: CTAs of a 1-D grid are dispatched in blockIdx order (the assumption CUB's
// decoupled look-back scan makes too); a wait that outlasts LB_WAIT_NS gives up and sets CGRB_ST_INTERNAL instead
// of hanging.  (A persistent variant with a producer warp and a two-stage ring was built and measured: CTAs that
// hold claimed tiles behind the one they are processing turn the look-back into a convoy)
//  0. staging: the aligned middle of each list slice (+ MD_HALO events of look-back) by the bulk-async engine, the
//     < 16-byte heads and tails by plain loads.
//  1. merge.  A tile fed by one list only -- nearly every tile of a background-dominated unit -- is already in
//     merged order.  A tile with a sparse minority (<= MD_SPARSE events of one list: the background under a bright
//     burst, the burst's photons in the background) is merged by insertion: every minority event finds its place
//     in the majority by bisection, every majority event shifts by the number of insertions before it (a count per
//     32 events + the few insertions of its own block).  Anything else takes the general merge-path: MD_PER
//     consecutive outputs per thread from its split.  In place, through registers.
//  2. every warp owns 256 consecutive events of the tile, lane l the events 32 r + l (r = 0..7), held in registers
//     from here on.  Literal dead time: only an event that follows a CHAIN element (overflow channel, the unit's
//     event 0; the first staged event stands in for whatever lies before the halo) by at most the longest window
//     can be removed, so the chain elements -- a few per cent of the events -- mark their followers in a bitmap,
//     and only marked events take the exact walk (dt_keep_local, or the global accessor when the chain leaves the
//     halo).  Everything else is kept without further work.
//  3. survivors are ranked by ballots (per warp) and a sum over the warps, warp 0 publishes the tile's count and
//     looks back, and every thread stores its survivors at rank order: lane-consecutive, coalesced stores.
template <int MODE>
__global__ void __launch_bounds__(NT, MD_CTAS) k_merge_deadtime(const cgrb_unit_t *__restrict__ units,
                                                          const double *__restrict__ times_bkg,
                                                          const uint8_t *__restrict__ pha_bkg,
                                                          const double *__restrict__ times_src,
                                                          const uint8_t *__restrict__ pha_src,
                                                          unsigned long long *tile_state,
                                                          const MdTile *__restrict__ tiles,
                                                          double *__restrict__ times_out, uint8_t *__restrict__ pha_out,
                                                          cgrb_counts_t *counts, const DtModel md)
{
    __shared__ MdSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    MDT_DECL(8)
    const MdTile tile = tiles[blockIdx.x];
    if (!(tile.flags & MDT_LIVE))
        return;                 // beyond the last live ticket
    if (tid == 0) {
        mbar_init(&sm.full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < MD_WORDS + 2)
        sm.rare[tid] = 0u;
    __syncthreads();
    const int na = tile.na, nb = tile.nb, n = na + nb;      // n = halo + tile <= MD_N
    const int64_t d0 = tile.d0;
    const int l0 = tile.l0, nt = tile.nt;                   // l0 is 0, MD_HALO or MD_HALO_DENSE
    const int unit = tile.unit;
    const bool has0 = (tile.flags & MDT_HAS0) != 0;
    // ---- 0. staging: slice A at t[ta + k] / p[qa + k], slice B at t[tb + k] / p[qb + k]; shared and global
    // addresses agree mod 16
    const double *gA = times_bkg + tile.a_off, *gB = times_src + tile.b_off;
    const uint8_t *hA = pha_bkg + tile.a_off, *hB = pha_src + tile.b_off;
    const int ta = (int)(((uintptr_t)gA >> 3) & 1u);
    const int tb = ((ta + na + 1) & ~1) + (int)(((uintptr_t)gB >> 3) & 1u);
    const int qa = (int)((uintptr_t)hA & 15u);
    const int qb = ((qa + na + 15) & ~15) + (int)((uintptr_t)hB & 15u);
    bool failed = false;
    {
        const BulkPlan ptA = bulk_plan<8>(gA, na), ptB = bulk_plan<8>(gB, nb);
        const BulkPlan ppA = bulk_plan<1>(hA, na), ppB = bulk_plan<1>(hB, nb);
        const unsigned bulk_bytes = 8u * (unsigned)(ptA.body + ptB.body) + (unsigned)(ppA.body + ppB.body);
        if (tid == 0 && bulk_bytes) {
            mbar_arrive_expect_tx(&sm.full, bulk_bytes);
            if (ptA.body)
                bulk_issue(sm.t + ta + ptA.head, gA + ptA.head, 8u * ptA.body, &sm.full);
            if (ptB.body)
                bulk_issue(sm.t + tb + ptB.head, gB + ptB.head, 8u * ptB.body, &sm.full);
            if (ppA.body)
                bulk_issue(sm.p + qa + ppA.head, hA + ppA.head, (unsigned)ppA.body, &sm.full);
            if (ppB.body)
                bulk_issue(sm.p + qb + ppB.head, hB + ppB.head, (unsigned)ppB.body, &sm.full);
        }
        if (wid >= 1 && wid <= 4) {
            // heads and tails (< 16 bytes each) by plain loads, one warp per slice: a time slice has at most one head
            // and one tail element, a pha slice at most 15 + 15
            const int which = wid - 1;                      // 0: t of A, 1: t of B, 2: p of A, 3: p of B
            const BulkPlan bp = which == 0 ? ptA : which == 1 ? ptB : which == 2 ? ppA : ppB;
            const int cnt = (which & 1) ? nb : na;
            const int tail0 = bp.head + bp.body, ntail = cnt - tail0;
            const int e = lane < bp.head ? lane : (lane - bp.head < ntail ? tail0 + (lane - bp.head) : -1);
            if (e >= 0) {
                if (which == 0)
                    sm.t[ta + e] = gA[e];
                else if (which == 1)
                    sm.t[tb + e] = gB[e];
                else if (which == 2)
                    sm.p[qa + e] = hA[e];
                else
                    sm.p[qb + e] = hB[e];
            }
        }
        if (bulk_bytes)
            failed = !mbar_wait(&sm.full, 0u);
    }
    __syncthreads();
    MDT_MARK(0);

    // ---- 1. merged order of the staged events: T[0 .. n), P[0 .. n) ---------------------------------------
    const double *T = sm.t;
    const uint8_t *P = sm.p;
    if (na == 0 || nb == 0) {
        // a tile fed by one list only: the slice IS the merged order
        T = sm.t + (na ? ta : tb);
        P = sm.p + (na ? qa : qb);
    } else if (min(na, nb) <= MD_SPARSE)
        md_merge_insertion(sm, na, nb, ta, tb, qa, qb);
    else
        md_merge_general(sm, na, nb, ta, tb, qa, qb);
    MDT_MARK(1);

    // ---- 2. keep flags of this thread's events (local indices l0 + 256 wid + 32 r + lane) ---------------------
    const int e0 = wid * (32 * MD_R) + lane;    // tile element of round 0
    unsigned keep_bits = 0;                     // bit r: the event of round r survives
    if (MODE == 0) {
        for (int h = tid; h < l0; h += NT) {
            // halo: its overflow events, and its first event in place of whatever lies before the halo
            if (h == 0 || P[h] == 127) {
                const double te = T[h] + DT_TAU;
                for (int k = max(h + 1, l0); k < n && !(te < T[k]); k++)
                    atomicOr(&sm.rare[k >> 5], 1u << (k & 31));
            }
        }
        unsigned chain_bits = 0;                // bit r: the event of round r is a chain element
#pragma unroll
        for (int r = 0; r < MD_R; r++) {
            const int e = e0 + 32 * r, li = l0 + e;
            chain_bits |= ((e < nt && (P[li] == 127 || (has0 && li == 0))) ? 1u : 0u) << r;
        }
        while (chain_bits) {
            // a chain element: its followers within the longest window need the exact walk
            const int li = l0 + e0 + 32 * (__ffs(chain_bits) - 1);
            chain_bits &= chain_bits - 1u;
            const double te = T[li] + DT_TAU;
            for (int k = li + 1; k < n && !(te < T[k]); k++)
                atomicOr(&sm.rare[k >> 5], 1u << (k & 31));
        }
        __syncthreads();
        MDT_MARK(2);
        const int64_t g0 = d0 - l0;
        keep_bits = (1u << min(max((nt - e0 + 31) >> 5, 0), MD_R)) - 1u;       // this thread's events inside the tile
        unsigned rare_bits = 0;
#pragma unroll
        for (int r = 0; r < MD_R; r++)
            rare_bits |= ((sm.rare[(l0 >> 5) + wid * MD_R + r] >> lane) & 1u) << r;     // word li >> 5: one per warp and round
        rare_bits &= keep_bits;
        while (rare_bits) {
            const int r = __ffs(rare_bits) - 1;
            rare_bits &= rare_bits - 1u;
            const int e = e0 + 32 * r;
            int k3 = dt_keep_local(T, P, l0 + e, g0);
            if (k3 == 2) {
                const cgrb_unit_t *u = units + unit;
                const MergedGlobal mg{times_bkg + u->bkg_off, times_src + u->src_off, pha_bkg + u->bkg_off,
                                      pha_src + u->src_off, counts[unit].n_bkg, counts[unit].n_src};
                k3 = dt_keep_slow(mg, d0 + e) ? 1 : 0;
            }
            if (!k3)
                keep_bits &= ~(1u << r);
        }
    } else {
        // physical mode: a unit already flagged for the serial fix-up needs no more work here
        const bool flagged = ((*(volatile uint32_t *)&counts[unit].status) & CGRB_ST_DEADTIME_CHAIN) != 0;
        const EventArrays ev{T, P};
#pragma unroll
        for (int r = 0; r < MD_R; r++) {
            const int e = e0 + 32 * r, li = l0 + e;
            int keep = 0;
            if (e < nt) {
                // no synchronisation point inside the tile's halo (MD_HALO events closer than the longest
                // window: never at GBM rates): the unit goes to the serial fix-up kernel
                keep = flagged ? 0 : dt_keep_phys(ev, (int64_t)li, has0, md);
                if (keep == 2) {
                    atomicOr(&counts[unit].status, CGRB_ST_DEADTIME_CHAIN);
                    keep = 0;
                }
            }
            keep_bits |= (unsigned)keep << r;
        }
    }
    MDT_MARK(3);
    // rank of each survivor inside its warp's 256 events (8 bits per round), survivors of the warp
    unsigned long long ranks = 0ull;
    int wsum = 0;
#pragma unroll
    for (int r = 0; r < MD_R; r++) {
        const unsigned bal = __ballot_sync(0xffffffffu, (keep_bits >> r) & 1u);
        ranks |= (unsigned long long)(wsum + __popc(bal & ((1u << lane) - 1u))) << (8 * r);
        wsum += __popc(bal);
    }
    if (lane == 0)
        sm.wtot[wid] = wsum;
    __syncthreads();
    MDT_MARK(4);

    // ---- 3. output offset: sum over the warps; warp 0 publishes the tile's count and looks back ----------------
    int wbase = 0, cnt_tile = 0;
#pragma unroll
    for (int q = 0; q < NW; q++) {
        const int c = sm.wtot[q];
        wbase += (q < wid) ? c : 0;
        cnt_tile += c;
    }
    if (wid == 0) {
        const int64_t w = tile.wb + d0 / MD_TILE;
        const long long excl = lookback_exclusive(tile_state, w, tile.wb, cnt_tile, lane, &failed);
        if (lane == 0) {
            sm.excl = excl;
            if (tile.flags & MDT_LAST)
                counts[unit].n_kept = excl + cnt_tile;        // the last live tile of the unit
        }
    }
    if (failed)
        atomicOr(&counts[unit].status, CGRB_ST_INTERNAL);
    __syncthreads();
    MDT_MARK(5);
    const int64_t out0 = tile.tot_off + sm.excl + wbase;
    double *ot = times_out + out0;
    uint8_t *op = pha_out + out0;
#pragma unroll
    for (int r = 0; r < MD_R; r++) {
        if ((keep_bits >> r) & 1u) {
            const int rank = (int)((ranks >> (8 * r)) & 0xffull);
            const int li = l0 + e0 + 32 * r;
            ot[rank] = T[li];
            op[rank] = P[li];
        }
    }
    MDT_MARK(6);
#ifdef MD_TIMING
    if (tid == 0 && (blockIdx.x % 20000) == 777)
        printf("MDT tile %lld: stage %lld merge %lld mark %lld keep %lld count %lld lookback %lld write %lld\n",
               (long long)blockIdx.x, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6]);
#endif
}


// ==========================================================================================
// Physical dead time, serial fix-up.  A unit whose events stay closer than the longest window for more than
// DT_PHYS_MAX_WALK events in a row has no synchronisation point for the parallel resolution above (event rate x
// window > ~5: far beyond anything GBM records); the kernels flag it (CGRB_ST_DEADTIME_CHAIN) and this kernel
// -- one warp per unit, lane 0 working, a no-op for every unflagged unit -- redoes that unit with the serial
// definition itself, merging the two component lists on the fly when there is no merged list, and clears the flag.
// ==========================================================================================
__global__ void __launch_bounds__(128) k_deadtime_phys_fixup(const cgrb_unit_t *__restrict__ units, int n_units,
                                                             const double *__restrict__ times_tot,
                                                             const uint8_t *__restrict__ pha_tot,
                                                             const double *__restrict__ times_bkg,
                                                             const uint8_t *__restrict__ pha_bkg,
                                                             const double *__restrict__ times_src,
                                                             const uint8_t *__restrict__ pha_src,
                                                             double *__restrict__ times_out, uint8_t *__restrict__ pha_out,
                                                             uint8_t *__restrict__ selection, cgrb_counts_t *counts,
                                                             const DtModel md)
{
    const int ui = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ui >= n_units || (threadIdx.x & 31) != 0)
        return;
    if (!(counts[ui].status & CGRB_ST_DEADTIME_CHAIN))
        return;
    const cgrb_unit_t *u = units + ui;
    double *ot = times_out + u->tot_off;
    uint8_t *op = pha_out + u->tot_off;
    int64_t kept = 0;
    double t_end = -INFINITY;
    if (times_tot) {
        const double *t = times_tot + u->tot_off;
        const uint8_t *p = pha_tot + u->tot_off;
        const int64_t n = counts[ui].n_total;
        for (int64_t i = 0; i < n; i++) {
            const double ti = t[i];
            const uint8_t pi = p[i];
            const bool keep = (i == 0) || ti > t_end;
            if (keep) {
                t_end = ti + (pi == 127 ? md.tau_ovf : md.tau);
                ot[kept] = ti;
                op[kept] = pi;
                kept++;
            }
            if (selection)
                selection[u->tot_off + i] = keep ? 1 : 0;
        }
    } else {
        const double *A = times_bkg + u->bkg_off, *B = times_src + u->src_off;
        const uint8_t *pA = pha_bkg + u->bkg_off, *pB = pha_src + u->src_off;
        const int64_t nA = counts[ui].n_bkg, nB = counts[ui].n_src;
        int64_t ia = 0, ib = 0;
        for (int64_t i = 0; i < nA + nB; i++) {
            const bool take_a = ia < nA && (ib >= nB || A[ia] <= B[ib]);     // ties: background first
            const double ti = take_a ? A[ia] : B[ib];
            const uint8_t pi = take_a ? pA[ia] : pB[ib];
            ia += take_a ? 1 : 0;
            ib += take_a ? 0 : 1;
            if (i == 0 || ti > t_end) {
                t_end = ti + (pi == 127 ? md.tau_ovf : md.tau);
                ot[kept] = ti;
                op[kept] = pi;
                kept++;
            }
        }
        counts[ui].n_total = nA + nB;
    }
    counts[ui].n_kept = kept;
    counts[ui].status &= ~CGRB_ST_DEADTIME_CHAIN;
}
