// cgrb_thinning.cuh -- NHPP thinning of the source rate and the homogeneous background (gap pass, segment scan, decision pass)
// Part of the single translation unit cgrb_kernels.cu (included in order; see that file).
#pragma once

// ==========================================================================================
// thinning pass 1: exponential gaps of a segment, their running sum INSIDE the segment (stored: the
// candidate's arrival time minus the segment's start time) and the segment total.  Pass 2 turns the
// totals into segment start times, pass 3 adds the start time -- no gap is drawn twice.
// The summation order is fixed (pairs, tiles of THIN_TILE, segments of CGRB_SEG), so times are reproducible.
// ==========================================================================================
// ------------------------------------------------------------------------------------------
// Piecewise-constant majorant of the thinning (Philox mode only; injected uniforms replay the reference's
// single-fmax loop).  The reference proposes candidates at the constant rate fmax over the whole window and
// keeps each with probability min(p, 1), p(t) = lambda(t) / fmax (cpl_functions.py:134-188): for a pulse that
// fills a fraction of the window nearly all candidates are wasted (8 % acceptance over the §8d population).
// The squeeze table already bounds p on every cell of its uniform time grid: |p - linear interpolation| <=
// margin, so P_j = min(1, max(p_j, p_j+1) + margin_j) >= min(p, 1) on cell j.  Candidates are drawn as a
// unit-rate Poisson process in operational time tau = Lambda(t) = fmax * integral of P (the same stored
// exponential gaps, with rate 1), mapped back through the piecewise-linear Lambda, and kept with probability
// min(p, 1) / P_j: the accepted events are the same process, min(lambda, fmax), from Lambda_T candidates
// instead of fmax * T.  `thin_cum`: per node the cumulative Lambda (double), then per cell the reciprocal of its
// width in operational time (double), then per node a search guide (int32: the cell of tau = g * Lambda_T / (n - 1)).
// ------------------------------------------------------------------------------------------
struct ThinCum {
    const double *L;        // [tab_n] cumulative operational time at the nodes
    const double *invw;     // [tab_n - 1] 1 / (L[j+1] - L[j]) of every cell (0 for an empty cell)
    const int32_t *guide;   // [tab_n - 1]
};
__device__ __forceinline__ ThinCum thin_cum_of(const double *thin_cum, int64_t n_nodes_total, const cgrb_unit_t &u)
{
    ThinCum c;
    c.L = thin_cum + u.tab_off;
    c.invw = thin_cum + n_nodes_total + u.tab_off;
    c.guide = reinterpret_cast<const int32_t *>(thin_cum + 2 * n_nodes_total) + u.tab_off;
    return c;
}
__device__ __forceinline__ bool unit_uses_majorant(const double *thin_cum, const cgrb_rng_t &rng, const cgrb_unit_t &u)
{
    return thin_cum != nullptr && rng.mode == 0 && u.tab_n >= 4 && !(u.flags & CGRB_UNIT_NO_SOURCE);
}
__device__ __forceinline__ double majorant_of(const double2 n0, const double2 n1)
{
    if (n0.y < 0.0)
        return fmin(-n0.y, 1.0);                    // first cell of a pulse: the stored bound, no squeeze
    const double P = fmax(n0.x, n1.x) + n0.y;
    return (P <= 1.0) ? fmax(P, 0.0) : 1.0;        // NaN / inf margins -> 1 (the reference's own ceiling)
}
// candidates a unit can need: N with S_N > Lambda_T, N ~ Lambda_T +- sqrt(Lambda_T)
__device__ __forceinline__ double majorant_cap(double LT)
{
    return LT + 10.0 * sqrt(LT) + 64.0;
}

#define THIN_CUM_K 8
// one CTA per unit: cumulative majorant at the nodes (fixed association)
__global__ void __launch_bounds__(NT) k_thin_cum(const cgrb_unit_t *__restrict__ units, int n_units,
                                                 const double *__restrict__ thin_tab, double *__restrict__ thin_cum)
{
    __shared__ double s_w[NW];
    const int ui = blockIdx.x;
    if (ui >= n_units)
        return;
    const cgrb_unit_t *u = units + ui;
    const int n = (int)u->tab_n;
    if (n < 4 || (u->flags & CGRB_UNIT_NO_SOURCE))
        return;
    const double2 *sq = reinterpret_cast<const double2 *>(thin_tab) + u->tab_off;
    double *L = thin_cum + u->tab_off;
    const double scale = u->fmax * ((u->src_tstop - u->src_tstart) / (double)(n - 1));
    double carry = 0.0;
    if (threadIdx.x == 0)
        L[0] = 0.0;
    // THIN_CUM_K consecutive cells per thread (serial prefix), the thread totals through one block scan
    for (int base = 0; base < n - 1; base += THIN_CUM_K * NT) {
        const int j0 = base + THIN_CUM_K * threadIdx.x;
        double v[THIN_CUM_K];
        double2 prev = (j0 < n) ? sq[j0] : make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < THIN_CUM_K; q++) {
            const int j = j0 + q;
            double2 nxt = prev;
            double x = 0.0;
            if (j < n - 1) {
                nxt = sq[j + 1];
                x = scale * majorant_of(prev, nxt);
            }
            v[q] = (q ? v[q - 1] : 0.0) + x;
            prev = nxt;
        }
        double total;
        const double excl = block_excl_scan(v[THIN_CUM_K - 1], s_w, &total);
#pragma unroll
        for (int q = 0; q < THIN_CUM_K; q++)
            if (j0 + q < n - 1)
                L[j0 + q + 1] = carry + (excl + v[q]);
        carry = carry + total;
    }
}

// search guide of the cumulative table, one thread per entry (node work list: which=4, chunk=256)
__global__ void __launch_bounds__(NT) k_thin_guide(const cgrb_unit_t *__restrict__ units,
                                                   const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                   double *__restrict__ thin_cum, int64_t n_nodes_total)
{
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int n = (int)u->tab_n;
    if (n < 4 || (u->flags & CGRB_UNIT_NO_SOURCE))
        return;
    const int g = (int)wk.start + threadIdx.x;
    if (g >= n - 1)
        return;
    const double *L = thin_cum + u->tab_off;
    int32_t *guide = reinterpret_cast<int32_t *>(thin_cum + 2 * n_nodes_total) + u->tab_off;
    {
        const double wdt = L[g + 1] - L[g];         // cell g: the map tau -> t inside it is one multiplication
        thin_cum[n_nodes_total + u->tab_off + g] = (wdt > 0.0) ? 1.0 / wdt : 0.0;
    }
    const double target = (double)g * L[n - 1] / (double)(n - 1);
    int lo = 0, hi = n - 1;                     // largest j in [0, n-2] with L[j] <= target
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (L[mid] <= target)
            lo = mid;
        else
            hi = mid;
    }
    guide[g] = lo;
}

template <int WHICH>  // 0 = source (prefix -> stage at src_off), 1 = background (prefix -> times_bkg at bkg_off + 1)
__global__ void __launch_bounds__(NT) k_gap_prefix(const cgrb_unit_t *__restrict__ units,
                                                   const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                   cgrb_rng_t rng, double *__restrict__ segsum,
                                                   double *__restrict__ prefix_out, cgrb_counts_t *counts,
                                                   const double *__restrict__ thin_cum, int64_t n_nodes_total)
{
    __shared__ double s_w[NW];
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t cap = (WHICH == 0) ? u->src_cap : u->bkg_cap;
    double rate = (WHICH == 0) ? u->fmax : u->bkg_rate;
    if (WHICH == 0 && unit_uses_majorant(thin_cum, rng, *u)) {
        // operational time: unit-rate gaps; segments beyond the candidates the unit can need are dead
        rate = 1.0;
        const double LT = thin_cum[u->tab_off + u->tab_n - 1];
        if ((double)wk.start > majorant_cap(LT)) {
            if (threadIdx.x == 0)
                segsum[w] = INFINITY;
            return;
        }
    }
    const double inv_f = 1.0 / rate;
    const uint32_t stage = (WHICH == 0) ? ST_SRC_TIME : ST_BKG;
    const uint32_t sid = u->stream_id;
    // a segment ends where the unit's next work entry starts (segments are at most CGRB_SEG candidates; small units
    // are cut into shorter ones by the host so that their serial depth stays low)
    int64_t seg_len = CGRB_SEG;
    if (w + 1 < n_work && work[w + 1].unit == wk.unit)
        seg_len = work[w + 1].start - wk.start;
    const int64_t n = min(seg_len, cap - wk.start);
    double *out = prefix_out + ((WHICH == 0) ? u->src_off : u->bkg_off + 1) + wk.start;
    double carry = 0.0;
    bool exhausted = false;
    for (int64_t base = 0; base < n; base += THIN_TILE) {
        const int64_t j0 = base + 2 * threadIdx.x;          // wk.start and base are even
        double g0 = 0.0, g1 = 0.0;
        if (j0 < n) {
            PairDraw d = pair_draw(rng, wk.unit, sid, stage, (wk.start + j0) >> 1, 0);
            // time - (1/fmax) * log(u)  cpl_functions.py:160
            g0 = (rng.mode == 0) ? inv_f * neglog_fast(d.a) : -(inv_f * log(d.a));
            if (j0 + 1 < n)
                g1 = (rng.mode == 0) ? inv_f * neglog_fast(d.b) : -(inv_f * log(d.b));
            exhausted |= d.exhausted;
        }
        double total;
        const double excl = block_excl_scan(g0 + g1, s_w, &total);
        const double p0 = (carry + excl) + g0, p1 = p0 + g1;
        if (j0 < n)
            out[j0] = p0;
        if (j0 + 1 < n)
            out[j0 + 1] = p1;
        carry = carry + total;
    }
    if (threadIdx.x == 0)
        segsum[w] = carry;
    if (exhausted)
        atomicOr(&counts[wk.unit].status, CGRB_ST_UNIFORMS_EXHAUSTED);
}

// pass 2: sequential (association-exact) scan of the segment sums of each unit: one warp per unit
template <int WHICH>
__global__ void __launch_bounds__(128) k_seg_scan(const cgrb_unit_t *__restrict__ units, int n_units,
                                                  const int64_t *__restrict__ unit_work_begin,
                                                  const double *__restrict__ segsum,
                                                  double *__restrict__ segstart, cgrb_counts_t *counts,
                                                  cgrb_rng_t rng, const double *__restrict__ thin_cum)
{
    const int lane = threadIdx.x & 31;
    const int ui = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ui >= n_units)
        return;
    const cgrb_unit_t *u = units + ui;
    double tstart = (WHICH == 0) ? u->src_tstart : u->bkg_tstart;
    double tstop = (WHICH == 0) ? u->src_tstop : u->bkg_tstop;
    if (WHICH == 0 && unit_uses_majorant(thin_cum, rng, *u)) {
        tstart = 0.0;                                       // operational time runs over [0, Lambda_T]
        tstop = thin_cum[u->tab_off + u->tab_n - 1];
    }
    const int64_t b = unit_work_begin[ui], e = unit_work_begin[ui + 1];
    double t = tstart;
    for (int64_t s0 = b; s0 < e; s0 += 32) {
        double v = (s0 + lane < e) ? segsum[s0 + lane] : 0.0;
        double mine = 0.0;
        for (int j = 0; j < 32; j++) {
            double vj = __shfl_sync(0xffffffffu, v, j);
            if (lane == j)
                mine = t;
            t = t + vj;     // same order on every lane
        }
        if (s0 + lane < e)
            segstart[s0 + lane] = mine;
    }
    const uint32_t disabled = (WHICH == 0) ? (u->flags & CGRB_UNIT_NO_SOURCE) : (u->flags & CGRB_UNIT_NO_BACKGROUND);
    if (lane == 0 && !disabled && !(t > tstop))
        atomicOr(&counts[ui].status, (WHICH == 0) ? CGRB_ST_SRC_CAPACITY : CGRB_ST_BKG_CAPACITY);
}

// ==========================================================================================
// source thinning pass 3 (cpl_functions.py:134-188): candidate time = segment start + stored running
// sum, acceptance by the squeeze table or the exact lambda(t), ordered compaction of the accepted
// times -- in place, into the head of the segment's own slots of `stage`.
// ==========================================================================================
// the exact test is rare under the squeeze table: kept out of line so the common path stays lean
__device__ __noinline__ double lambda_exact(const LambdaTab *tab, const cgrb_unit_t &u, double t)
{
    return lambda_at(tab, u, t);
}

// self-check of the squeeze margins wherever the exact lambda(t) is evaluated anyway: the margin must bound the
// distance between p(t) and the table's linear interpolation (finite margins only)
__device__ __forceinline__ bool squeeze_margin_violated(double pe, double pl, double margin)
{
    return margin >= 0.0 && margin < INFINITY && fabs(pe - pl) > margin * (1.0 + 1e-9) + 1e-15;
}

// The exact test with its self-checks, out of line: nothing of the common path stays live across it.  Returns
// bit 0: accept (v <= p(t)), bit 1: p(t) above the cell's majorant P (P < 1), bit 2: squeeze margin violated.
#define THIN_ACC 1u
#define THIN_MAJ_VIOLATED 2u
#define THIN_SQ_VIOLATED 4u
__device__ __noinline__ unsigned thin_exact(const LambdaTab *tab, const cgrb_unit_t &u, const double2 *__restrict__ sq,
                                            int j, double fr, double t, double v, double P)
{
    const double pe = lambda_at(tab, u, t) / u.fmax;
    unsigned r = (v <= pe) ? THIN_ACC : 0u;         // `if test <= p_test`
    if (sq) {
        const double2 n0 = __ldg(sq + j), n1 = __ldg(sq + j + 1);
        const double pl = n0.x + (n1.x - n0.x) * fr;
        if (P < 1.0 && pe > P * (1.0 + 1e-9))
            r |= THIN_MAJ_VIOLATED;
        if (squeeze_margin_violated(pe, pl, n0.y))
            r |= THIN_SQ_VIOLATED;
    }
    return r;
}

// returns THIN_ACC | violation bits
__device__ __forceinline__ unsigned thin_accept(const LambdaTab *tab, const cgrb_unit_t &u, bool squeeze,
                                                const double2 *__restrict__ sq, double sq_inv_h, double t, double u2,
                                                long long *n_exact)
{
    int kq = 0;
    double fr = 0.0;
    if (squeeze) {
        const double q = (t - u.src_tstart) * sq_inv_h;
        kq = min(max((int)q, 0), (int)u.tab_n - 2);       // tab_n < 2^31 (host: <= 65536 nodes)
        fr = q - (double)kq;
        const double2 n0 = __ldg(sq + kq), n1 = __ldg(sq + kq + 1);
        const double pl = n0.x + (n1.x - n0.x) * fr;
        if (n0.y >= 0.0) {                  // negative: a cell without squeeze
            if (u2 <= pl - n0.y)
                return THIN_ACC;
            if (u2 > pl + n0.y)
                return 0u;
        }
    }
    (*n_exact)++;
    return thin_exact(tab, u, squeeze ? sq : nullptr, kq, fr, t, u2, 1.0);
}


// Majorant mode: candidate at operational time tau -> cell j (guide + short walk over the cumulative table),
// real time t inside the cell, acceptance with probability min(p(t), 1) / P_j by the same squeeze test.
// *t_out receives the real time.  Returns THIN_ACC | violation bits (an exact evaluation that finds p(t) above the
// cell's bound, or farther from the interpolation than the margin).
__device__ __forceinline__ unsigned thin_accept_majorant(const LambdaTab *tab, const cgrb_unit_t &u,
                                                         const double2 *__restrict__ sq, const ThinCum tc, double gscale,
                                                         double h, double inv_fh, double tau, double u2, double *t_out,
                                                         long long *n_exact)
{
    const int n = (int)u.tab_n;
    int j = __ldg(tc.guide + min(max((int)(tau * gscale), 0), n - 2));
    double Lj = __ldg(tc.L + j), Lj1 = __ldg(tc.L + j + 1);
    while (j < n - 2 && Lj1 <= tau) {
        j++;
        Lj = Lj1;
        Lj1 = __ldg(tc.L + j + 1);
    }
    const double2 n0 = __ldg(sq + j), n1 = __ldg(sq + j + 1);
    const double wdt = Lj1 - Lj;                    // = fmax h P_j: expected candidates of the cell
    // the cell's majorant as the proposal actually realises it: candidates fall into the cell at the rate wdt / h, so
    // the acceptance probability that leaves lambda(t) is p / (wdt / (fmax h)) -- one multiplication, and exact even
    // where the rounding of the cumulative sums moved wdt by an ulp of Lambda (majorant_of(n0, n1) to ~1e-12)
    // (a cell clamped at the reference's own ceiling P = 1 -- p may exceed 1 there -- must read exactly 1)
    const double Pw = wdt * inv_fh;
    const double P = (Pw > 1.0 - 1e-9) ? 1.0 : Pw;
    // position inside the cell: a monotone map (one multiplication by the cell's stored reciprocal width)
    const double fr = fmin(fmax((tau - Lj) * __ldg(tc.invw + j), 0.0), 1.0);
    const double tj = u.src_tstart + (double)j * h;
    const double tj1 = (j == n - 2) ? u.src_tstop : u.src_tstart + (double)(j + 1) * h;
    const double t = fmin(fma(fr, h, tj), tj1);
    *t_out = t;
    if (!(P > 0.0))
        return 0u;
    const double v = u2 * P;
    const double pl = n0.x + (n1.x - n0.x) * fr;
    if (n0.y >= 0.0) {                      // negative: a cell without squeeze
        if (v <= pl - n0.y)
            return THIN_ACC;
        if (v > pl + n0.y)
            return 0u;
    }
    (*n_exact)++;
    return thin_exact(tab, u, sq, j, fr, t, v, P);
}

#ifndef THIN_CTAS
#define THIN_CTAS 3
#endif
__global__ void __launch_bounds__(NT, THIN_CTAS) k_source_thin(const double *__restrict__ dtab,
                                                    const cgrb_unit_t *__restrict__ units,
                                                    const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                    cgrb_rng_t rng, const double *__restrict__ segstart,
                                                    int32_t *__restrict__ segcount, double *stage,
                                                    double *__restrict__ cand_times,
                                                    uint8_t *__restrict__ cand_accept,
                                                    const double *__restrict__ thin_tab, cgrb_counts_t *counts,
                                                    const double *__restrict__ thin_cum, int64_t n_nodes_total)
{
    __shared__ LambdaTab tab;
    __shared__ cgrb_unit_t u;
    __shared__ int s_wi[NW];
    __shared__ double s_t[NT];
    __shared__ double s_last;
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    if (threadIdx.x == 0)
        u = units[wk.unit];
    __syncthreads();
    const double t0 = segstart[w];
    // majorant mode: segment starts, running sums and the stop are in operational time until the acceptance test
    const bool maj = unit_uses_majorant(thin_cum, rng, u);
    const ThinCum tc = thin_cum_of(thin_cum, n_nodes_total, u);
    const double stop = maj ? __ldg(tc.L + u.tab_n - 1) : u.src_tstop;
    if ((u.flags & CGRB_UNIT_NO_SOURCE) || !(t0 <= stop)) {   // dead segment
        if (threadIdx.x == 0)
            segcount[w] = 0;
        return;
    }
    build_lambda_tab(&tab, dtab + (size_t)u.det * CGRB_DT_SIZE, u);
    if (threadIdx.x == 0)
        s_last = t0;
    __syncthreads();

    int64_t seg_len = CGRB_SEG;             // (as in k_gap_prefix: up to the unit's next work entry)
    if (w + 1 < n_work && work[w + 1].unit == wk.unit)
        seg_len = work[w + 1].start - wk.start;
    const int64_t n = min(seg_len, u.src_cap - wk.start);
    // squeeze table of p(t) = lambda(t)/fmax (cgrb_thin_tables): decides most candidates without
    // evaluating lambda; the rest take the exact test, so decisions are those of the exact test
    const bool squeeze = thin_tab != nullptr && u.tab_n >= 4;
    const double2 *sq = reinterpret_cast<const double2 *>(thin_tab) + (squeeze ? u.tab_off : 0);
    const double sq_inv_h = squeeze ? (double)(u.tab_n - 1) / (u.src_tstop - u.src_tstart) : 0.0;
    const double maj_h = maj ? (u.src_tstop - u.src_tstart) / (double)(u.tab_n - 1) : 0.0;
    const double maj_gscale = (maj && stop > 0.0) ? (double)(u.tab_n - 1) / stop : 0.0;
    const double maj_inv_fh = (maj && maj_h > 0.0 && u.fmax > 0.0) ? 1.0 / (u.fmax * maj_h) : 0.0;
    unsigned vflags = 0u;                  // THIN_*_VIOLATED bits of the exact evaluations
    double *seg = stage + u.src_off + wk.start;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long n_exact = 0;
    int seg_n = 0;        // accepted so far in this segment
    int n_valid_total = 0;
    // the stored running sums of the NEXT tile are requested before this tile is processed, so their DRAM
    // latency hides behind the acceptance tests
    double nxt_a = (2 * threadIdx.x < n) ? seg[2 * threadIdx.x] : INFINITY;
    double nxt_b = (2 * threadIdx.x + 1 < n) ? seg[2 * threadIdx.x + 1] : INFINITY;
    for (int64_t base = 0; base < n; base += THIN_TILE) {
        const int64_t j0 = base + 2 * threadIdx.x;
        const bool in0 = j0 < n, in1 = j0 + 1 < n;
        double ta = in0 ? t0 + nxt_a : INFINITY;
        double tb = in1 ? t0 + nxt_b : INFINITY;
        {
            const int64_t jn = j0 + THIN_TILE;
            nxt_a = (jn < n) ? seg[jn] : INFINITY;
            nxt_b = (jn + 1 < n) ? seg[jn + 1] : INFINITY;
        }
        // enforce non-decreasing times across association orders (1-ulp inversions)
        s_t[threadIdx.x] = in1 ? tb : ta;
        __syncthreads();
        const double left = (threadIdx.x > 0) ? s_t[threadIdx.x - 1] : s_last;
        ta = fmax(ta, left);
        tb = fmax(tb, ta);
        const double tile_last = in1 ? tb : ta;     // (operational time in majorant mode)
        const bool va = in0 && (ta <= stop);   // `if time > tstop: break`
        const bool vb = in1 && (tb <= stop);
        bool aa = false, ab = false;
        if (va) {
            PairDraw d = pair_draw(rng, wk.unit, u.stream_id, ST_SRC_TIME, (wk.start + j0) >> 1, 1);
            if (maj) {
                const double tau_a = ta, tau_b = tb;
                unsigned ra = thin_accept_majorant(&tab, u, sq, tc, maj_gscale, maj_h, maj_inv_fh, tau_a, d.a, &ta, &n_exact), rb = 0u;
                if (vb)
                    rb = thin_accept_majorant(&tab, u, sq, tc, maj_gscale, maj_h, maj_inv_fh, tau_b, d.b, &tb, &n_exact);
                aa = (ra & THIN_ACC) != 0u;
                ab = (rb & THIN_ACC) != 0u;
                vflags |= ra | rb;
            } else {
                unsigned ra = thin_accept(&tab, u, squeeze, sq, sq_inv_h, ta, d.a, &n_exact), rb = 0u;
                if (vb)
                    rb = thin_accept(&tab, u, squeeze, sq, sq_inv_h, tb, d.b, &n_exact);
                aa = (ra & THIN_ACC) != 0u;
                ab = (rb & THIN_ACC) != 0u;
                vflags |= ra | rb;
            }
        }
        // ordered compaction: per-thread counts (0..2 accepted, 0..2 valid) packed in one int
        const int mine = (int)aa + (int)ab + (((int)va + (int)vb) << 16);
        int incl = warp_incl_scan_i(mine, lane);
        if (lane == 31)
            s_wi[wid] = incl;
        __syncthreads();            // also: every read of this tile's slots is done before any write below
        int before = 0, tile_tot = 0;
#pragma unroll
        for (int q = 0; q < NW; q++) {
            const int c = s_wi[q];
            if (q < wid)
                before += c;
            tile_tot += c;
        }
        const int rank = ((before + incl - mine) & 0xffff) + seg_n;
        if (aa)
            seg[rank] = ta;
        if (ab)
            seg[rank + (int)aa] = tb;
        if (cand_times) {
            if (va) {
                cand_times[u.src_off + wk.start + j0] = ta;
                cand_accept[u.src_off + wk.start + j0] = aa ? 1 : 0;
            }
            if (vb) {
                cand_times[u.src_off + wk.start + j0 + 1] = tb;
                cand_accept[u.src_off + wk.start + j0 + 1] = ab ? 1 : 0;
            }
        }
        const int tile_acc = tile_tot & 0xffff, tile_valid = tile_tot >> 16;
        seg_n += tile_acc;
        n_valid_total += tile_valid;
        if (threadIdx.x == NT - 1)
            s_last = tile_last;
        __syncthreads();
        if (tile_valid < (int)min((int64_t)THIN_TILE, n - base))
            break;      // crossed tstop: everything later is beyond it
    }
    if (vflags & (THIN_MAJ_VIOLATED | THIN_SQ_VIOLATED))
        atomicOr(&counts[wk.unit].status, ((vflags & THIN_MAJ_VIOLATED) ? CGRB_ST_MAJORANT : 0u) |
                                              ((vflags & THIN_SQ_VIOLATED) ? CGRB_ST_SQUEEZE : 0u));
    {
        long long *s_ll = reinterpret_cast<long long *>(s_t);      // s_t is free now
        __syncthreads();
        const long long tot_exact = block_sum_ll(n_exact, s_ll);
        if (threadIdx.x == 0) {
            segcount[w] = seg_n;
            if (n_valid_total)
                atomicAdd((unsigned long long *)&counts[wk.unit].n_candidates, (unsigned long long)n_valid_total);
            if (tot_exact)
                atomicAdd((unsigned long long *)&counts[wk.unit].n_exact, (unsigned long long)tot_exact);
        }
    }
}

// pass 5: gather the per-segment accepted times into the unit's contiguous list; entry 0 is the
// spurious tstart event of cpl_functions.py:153-154.
__global__ void __launch_bounds__(NT) k_source_gather(const cgrb_unit_t *__restrict__ units,
                                                      const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                      const int64_t *__restrict__ unit_work_begin,
                                                      const int32_t *__restrict__ segcount,
                                                      const double *__restrict__ stage,
                                                      double *__restrict__ times_src, cgrb_counts_t *counts)
{
    __shared__ long long s_w[NW];
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t b = unit_work_begin[wk.unit], e = unit_work_begin[wk.unit + 1];
    long long part = 0;
    for (int64_t s = b + threadIdx.x; s < w; s += NT)
        part += segcount[s];
    const long long off = block_sum_ll(part, s_w);
    const int cnt = segcount[w];
    const double *src = stage + u->src_off + wk.start;
    double *dst = times_src + u->src_off + 1 + off;
    for (int j = threadIdx.x; j < cnt; j += NT)
        dst[j] = src[j];
    if (threadIdx.x == 0) {
        if (w == b)
            times_src[u->src_off] = u->src_tstart;
        if (w == e - 1)
            counts[wk.unit].n_src = (u->flags & CGRB_UNIT_NO_SOURCE) ? 0 : 1 + off + cnt;
    }
}

// ==========================================================================================
// background (background.py:13-44 + numpy choice :93) in ONE pass: every candidate is an event; event j of the
// unit is candidate j-1, event 0 is the spurious tstart event.  A CTA owns a tile of BKG_TILE consecutive candidates:
//   1. exponential gaps (one Philox call per candidate pair) and their running sum inside the tile, in registers;
//   2. the tile's start time is the END time of the unit's previous tile, read from a per-tile state word (the
//      tiles are launched level-major, k_level_order: in a population batch the previous tile finished a whole
//      level ago; with few units the chain costs one L2 round trip per tile, while the gaps of all the tiles in
//      flight are drawn in parallel); the tile publishes its own end time at once;
//   3. time = start + running sum, channel from the template's cdf (guide + short walk), list cut at tstop.
// 9 B per event reach HBM (the two-pass version wrote and re-read an 8-byte running sum per candidate: 25 B).
// The running sums are built so that the list is sorted BY CONSTRUCTION: every partial sum that starts a lane, a
// round, a warp or a tile is the very value the element before it was given (see the comments below), and fp64
// addition is monotone.  The association is fixed, so the times are reproducible.
// ==========================================================================================
struct BkgTile {
    int64_t sidx;           // the tile's state word; the unit's previous tile is sidx - 1
    int64_t off;            // slot of the tile's candidate 0 in times_bkg / pha_bkg (the unit's event 1 + k BKG_TILE)
    double tstart, tstop, inv_rate;
    int32_t unit;           // -1: beyond the last live ticket
    int32_t k;              // tile of the unit: candidates [k BKG_TILE, (k+1) BKG_TILE)
    int32_t n;              // candidates of the tile (the unit's capacity may end inside it)
    uint32_t sid;           // Philox stream of the unit
    int32_t det;            // detector table
    uint32_t last;          // the unit's last tile
};
static_assert(sizeof(BkgTile) == 64, "BkgTile is one 64-byte record");

// one thread per ticket: everything a tile needs in ONE record (the kernel's dependence chain in front of its
// arithmetic is a single load)
__global__ void __launch_bounds__(NT) k_background_plan(const cgrb_unit_t *__restrict__ units, int64_t n_tickets,
                                                        const LevelOrder o, BkgTile *__restrict__ tiles)
{
    const int64_t T = (int64_t)blockIdx.x * NT + threadIdx.x;
    if (T >= n_tickets)
        return;
    BkgTile t;
    t.sidx = t.off = 0;
    t.tstart = t.tstop = t.inv_rate = 0.0;
    t.unit = -1;
    t.k = t.n = t.det = 0;
    t.sid = t.last = 0u;
    int unit, level;
    if (level_ticket(o, T, &unit, &level)) {
        const cgrb_unit_t *u = units + unit;
        const int64_t start = (int64_t)level * BKG_TILE;
        t.sidx = o.ubeg[unit] + level;
        t.off = u->bkg_off + 1 + start;
        t.tstart = u->bkg_tstart;
        t.tstop = u->bkg_tstop;
        t.inv_rate = 1.0 / u->bkg_rate;
        t.unit = unit;
        t.k = level;
        t.n = (int)max((int64_t)0, min((int64_t)BKG_TILE, u->bkg_cap - start));
        t.sid = u->stream_id;
        t.det = u->det;
        t.last = (start + BKG_TILE >= u->bkg_cap) ? 1u : 0u;
    }
    tiles[T] = t;
}

#define BKG_NOT_READY 0xFFFFFFFFFFFFFFFFull     // state words are set to this (a NaN no tile can publish) before the launch
#ifndef BKG_CTAS
#define BKG_CTAS 4
#endif
#define BKG_R (BKG_TILE / (2 * NT))             // rounds: a thread owns one candidate PAIR per round (4)

// np.searchsorted(cdf, u, 'right') clamped to the last channel (numpy's legacy choice, background.py:93), started
// from the detector's guide (1024 bytes: guide[m] = searchsorted(cdf, m / 1024, 'right'), built with the table):
// a guide cell holds a channel boundary one time in eight, so the walk is one probe, rarely two.  The 14 tables
// (1 KB + 1 KB each) stay in L1: no staging, no barrier.
__device__ __forceinline__ uint8_t bkg_channel_of(const double *__restrict__ cdf, const uint8_t *__restrict__ guide,
                                                  double uc)
{
    int j = (int)__ldg(guide + min(max((int)(uc * (double)CGRB_BKG_GUIDE_N), 0), CGRB_BKG_GUIDE_N - 1));
    while (j < CGRB_NCHAN - 1 && !(uc < __ldg(cdf + j)))
        j++;
    return (uint8_t)j;
}

__global__ void __launch_bounds__(NT, BKG_CTAS) k_background(const double *__restrict__ dtab,
                                                             const BkgTile *__restrict__ tiles, cgrb_rng_t rng,
                                                             cgrb_rng_t rng_chan, unsigned long long *state,
                                                             double *__restrict__ times_bkg,
                                                             uint8_t *__restrict__ pha_bkg, cgrb_counts_t *counts)
{
    __shared__ double s_wsum[NW];
    __shared__ int s_wi[NW];
    __shared__ double s_t0;
    const BkgTile tl = tiles[blockIdx.x];
    if (tl.unit < 0)
        return;
    const double tstop = tl.tstop;
    const int n = tl.n;
    const int64_t start = (int64_t)tl.k * BKG_TILE;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double *cdf = dtab + (size_t)tl.det * CGRB_DT_SIZE + CGRB_DT_BKG_CDF;
    const uint8_t *guide = reinterpret_cast<const uint8_t *>(dtab + (size_t)tl.det * CGRB_DT_SIZE + CGRB_DT_BKG_GUIDE);
    // the previous tile's end time, asked for now and needed after the gaps (level-major order: it is there)
    unsigned long long prev = BKG_NOT_READY;
    if (threadIdx.x == 0 && tl.k > 0)
        prev = *((volatile unsigned long long *)(state + tl.sidx - 1));

    // ---- 1. gaps and running sums.  Warp w owns the candidates [256 w, 256 w + 256) of the tile, round r the 64
    // candidates from 256 w + 64 r, lane l the pair 2 l, 2 l + 1 of them.
    const double inv_f = tl.inv_rate;
    double p0[BKG_R], p1[BKG_R];
    uint32_t ca[BKG_R], cb[BKG_R];  // channel draws of the pairs (Philox: they ride in the call that gave the gaps)
    bool exhausted = false;
    double rc = 0.0;                // running sum of the warp's earlier rounds
#pragma unroll
    for (int r = 0; r < BKG_R; r++) {
        const int j0 = 256 * wid + 64 * r + 2 * lane;           // start and j0 are even
        double g0 = 0.0, g1 = 0.0;
        ca[r] = cb[r] = 0u;
        if (j0 < n) {
            // time - (1/rate) * log(u)  background.py:33
            if (rng.mode == 0) {
                const BkgDraw d = bkg_draw(philox_at(make_key(rng.seed), tl.sid, ST_BKG, (uint64_t)((start + j0) >> 1), 0u));
                g0 = inv_f * neglog_fast(d.ga);
                if (j0 + 1 < n)
                    g1 = inv_f * neglog_fast(d.gb);
                ca[r] = d.ca;
                cb[r] = d.cb;
            } else {
                PairDraw d = pair_draw(rng, tl.unit, tl.sid, ST_BKG, (start + j0) >> 1, 0);
                g0 = -(inv_f * log(d.a));
                if (j0 + 1 < n)
                    g1 = -(inv_f * log(d.b));
                exhausted |= d.exhausted;
            }
        }
        const double incl = warp_incl_scan(g0 + g1, lane);
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0)
            excl = 0.0;
        // the pair's second sum IS the value the next lane starts from (rc + incl), the round's last one IS the
        // next round's rc: no element can fall below its predecessor.  Inside the pair the two roundings differ.
        p1[r] = rc + incl;
        p0[r] = fmin((rc + excl) + g0, p1[r]);
        rc = rc + __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0)
        s_wsum[wid] = rc;
    __syncthreads();
    // sum over the earlier warps in a fixed order: wbase(w + 1) = wbase(w) + W(w) is the last element of warp w
    double wbase = 0.0, total = 0.0;
#pragma unroll
    for (int q = 0; q < NW; q++) {
        if (q == wid)
            wbase = total;
        total = total + s_wsum[q];
    }

    // ---- 2. start time: the end of the unit's previous tile; publish this tile's end
    if (threadIdx.x == 0) {
        double t0 = tl.tstart;
        bool failed = false;
        if (tl.k > 0) {
            unsigned long long t_begin = 0;
            for (int spin = 0; prev == BKG_NOT_READY; spin++) {
                prev = *((volatile unsigned long long *)(state + tl.sidx - 1));
                if ((spin & 1023) == 1023) {
                    const unsigned long long now = global_timer_ns();
                    if (t_begin == 0)
                        t_begin = now;
                    else if (now - t_begin > LB_WAIT_NS) {
                        failed = true;
                        break;
                    }
                }
            }
            t0 = failed ? INFINITY : __longlong_as_double((long long)prev);
        }
        double t_end = t0 + total;      // == the last candidate's time
        if (!(t_end == t_end))
            t_end = INFINITY;           // never publish the NOT_READY pattern
        *((volatile unsigned long long *)(state + tl.sidx)) = (unsigned long long)__double_as_longlong(t_end);
        s_t0 = t0;
        if (failed)
            atomicOr(&counts[tl.unit].status, CGRB_ST_INTERNAL);
    }
    __syncthreads();
    const double t0 = s_t0;

    if (tl.k == 0 && threadIdx.x == 0) {
        // event 0: tstart, with its own channel draw (sample_channel(size=len(times)))
        double uc;
        if (rng.mode == 0) {
            Philox4 p = philox_at(make_key(rng.seed), tl.sid, ST_BKG, 0xFFFFFFFFFFull, 1u);
            uc = u53(p.x, p.y);
        } else {
            int64_t o = rng_chan.d_off ? rng_chan.d_off[tl.unit] : 0;
            uc = rng_chan.d_u[o];
        }
        times_bkg[tl.off - 1] = tl.tstart;
        pha_bkg[tl.off - 1] = bkg_channel_of(cdf, guide, uc);
    }
    if (t0 > tstop)
        return;         // the list ended in an earlier tile (which recorded the count)

    // ---- 3. times, channels, the cut at tstop
    double *seg_t = times_bkg + tl.off;
    uint8_t *seg_p = pha_bkg + tl.off;
    int n_valid = 0;
#pragma unroll
    for (int r = 0; r < BKG_R; r++) {
        const int j0 = 256 * wid + 64 * r + 2 * lane;
        const double ta = t0 + (wbase + p0[r]), tb = t0 + (wbase + p1[r]);
        const bool va = (j0 < n) && (ta <= tstop);
        const bool vb = (j0 + 1 < n) && (tb <= tstop);
        if (va) {
            double ua, ub = 0.5;
            if (rng.mode == 0) {
                ua = u24d(ca[r]);
                ub = u24d(cb[r]);
            } else {
                const int64_t o = rng_chan.d_off ? rng_chan.d_off[tl.unit] : 0;
                const int64_t len = rng_chan.d_len ? rng_chan.d_len[tl.unit] : ((int64_t)1 << 62);
                const int64_t e0 = start + j0 + 1;              // event index of candidate j0
                ua = 0.5;
                if (e0 < len)
                    ua = rng_chan.d_u[o + e0];
                else
                    exhausted = true;
                if (vb) {
                    if (e0 + 1 < len)
                        ub = rng_chan.d_u[o + e0 + 1];
                    else
                        exhausted = true;
                }
            }
            seg_t[j0] = ta;
            seg_p[j0] = bkg_channel_of(cdf, guide, ua);
            if (vb) {
                seg_t[j0 + 1] = tb;
                seg_p[j0 + 1] = bkg_channel_of(cdf, guide, ub);
            }
        }
        n_valid += __popc(__ballot_sync(0xffffffffu, va)) + __popc(__ballot_sync(0xffffffffu, vb));
    }
    if (lane == 0)
        s_wi[wid] = n_valid;
    __syncthreads();
    if (threadIdx.x == 0) {
        int valid = 0;
#pragma unroll
        for (int q = 0; q < NW; q++)
            valid += s_wi[q];
        // the tile in which the list ends records the count: tstart + the candidates up to tstop
        if (valid < n || tl.last)
            counts[tl.unit].n_bkg = 1 + start + valid;
        if (tl.last && valid == n)
            atomicOr(&counts[tl.unit].status, CGRB_ST_BKG_CAPACITY);    // the buffer ended before tstop
    }
    if (exhausted)
        atomicOr(&counts[tl.unit].status, CGRB_ST_UNIFORMS_EXHAUSTED);
}
