// cgrb_trigger.cuh -- GBM trigger on the device (binning, prefix sums, sliding-window search)
// Part of the single translation unit cgrb_kernels.cu (included in order; see that file).
#pragma once

// ==========================================================================================
// GBM trigger on the device: GBMLightCurveAnalyzer._compute_detection
// (instruments/gbm/gbm_lightcurve_analyzer.py:63-143) on the dead-time-filtered total list of every unit --
// 16 ms binning in the four trigger energy ranges (light_curve_storage.py:154-294), dead time per bin
// (2 us / 10 us per event, :246-262), sliding-window significance (_run_trigger :199-243,
// dumb_significance :265-271) over the trigger time scales, first detection in the reference's order.
// The quirks of the reference are reproduced (see oracle/trigger_oracle.py): the count round trip
// ((c / dt) * dt).astype(int), the swapped window arguments, the inclusive dead-time interval.
// ==========================================================================================
#define TRIG_TILE 4096          // events per CTA (two passes of TRIG_EPT per thread)
#define TRIG_LOCAL 1024         // bins a tile may span and still be accumulated in shared memory (4096 events
                                // at the 500 /s background are 512 bins of 16 ms)
#define TRIG_EPT 8              // events per thread and pass
#define TRIG_PASSES (TRIG_TILE / (TRIG_EPT * NT))
#define TRIG_NCOMBO 33
#define TRIG_NARR 6             // per-bin integer arrays: counts of the 4 ranges, events, overflow events
#define TRIG_CH 1792            // windows per staged chunk of the search (a multiple of NT)
#define TRIG_SPAN 1832          // >= background (1062) + longest gap (512) + source (250) bins
#define TRIG_ST (TRIG_CH + TRIG_SPAN + 4)       // staged prefix entries per chunk (+1, padded)
#define TRIG_EVAL_SMEM ((3 * TRIG_ST + TRIG_CH) * 4)    // dynamic shared memory of k_trig_eval

// exact (double)i for 0 <= i < 2^32 without the quarter-rate I2F.F64: 2^52 + i is representable, subtract 2^52
__device__ __forceinline__ double trig_u2d(unsigned i)
{
    return __hiloint2double(0x43300000, (int)i) - 4503599627370496.0;
}

// np.arange(tmin, tmax, dt): length and the fill rule  edge[i] = start + i * (fl(start + dt) - start)
struct TrigGrid {
    double t0, delta;
    int64_t n_edges;
    __device__ __forceinline__ double edge(int64_t i) const { return __dadd_rn(t0, __dmul_rn((double)i, delta)); }
    // the same for 0 <= i < 2^31, with the exact integer-to-double conversion done by one addition
    __device__ __forceinline__ double edge32(int i) const { return __dadd_rn(t0, __dmul_rn(trig_u2d((unsigned)i), delta)); }
};
__device__ __forceinline__ TrigGrid trig_grid(double tmin, double tmax, double dt)
{
    TrigGrid g;
    g.t0 = tmin;
    g.delta = __dadd_rn(__dadd_rn(tmin, dt), -tmin);
    const double len = ceil(__ddiv_rn(__dadd_rn(tmax, -tmin), dt));
    g.n_edges = (len > 0.0) ? (int64_t)len : 0;
    return g;
}

// Shared-memory histogram of a tile, three packed words per bin (a tile holds TRIG_TILE events, so every 16-bit
// field stays below 2 x TRIG_TILE < 65536): word 0 = range a | range b << 16, word 1 = range c | range d << 16,
// word 2 = events | overflow events << 16 (the two dead-time counters).
#ifndef TRIG_BIN_CTAS
#define TRIG_BIN_CTAS 6
#endif
__global__ void __launch_bounds__(NT, TRIG_BIN_CTAS) k_trig_bin(const cgrb_unit_t *__restrict__ units,
                                                 const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                 const int32_t *__restrict__ chan_bounds, double dt,
                                                 const double *__restrict__ times_out,
                                                 const uint8_t *__restrict__ pha_out,
                                                 const cgrb_counts_t *__restrict__ counts, int64_t bins_per_unit,
                                                 int32_t *__restrict__ bins)
{
    __shared__ int s_h[3][TRIG_LOCAL];
    __shared__ uint8_t s_cm[256];           // channel -> membership mask of the four trigger ranges
    __shared__ int s_k0;
    __shared__ int s_local;
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t n = counts[wk.unit].n_kept;
    if (wk.start >= n || n < 2)
        return;
    const double *t = times_out + u->tot_off;
    const uint8_t *p = pha_out + u->tot_off;
    const TrigGrid g = trig_grid(t[0], t[n - 1], dt);
    if (g.n_edges - 1 < 1 || g.n_edges > 0x7fffffff)    // (more bins than the host can have allocated)
        return;
    const int nb = (int)(g.n_edges - 1);                // np.histogram bins
    const int64_t end = min(wk.start + (int64_t)TRIG_TILE, n);
    int32_t *ub = bins + (size_t)wk.unit * TRIG_NARR * bins_per_unit;
    const int32_t *cb = chan_bounds + 8 * u->det;

    // bin of an event: edge[k] <= t < edge[k+1] (last bin closed), -1 beyond the last edge
    const double inv_delta = 1.0 / g.delta;
    const double last_edge = g.edge32(nb);
    auto bin_of = [&](double tt) -> int {
        const double q = (tt - g.t0) * inv_delta;           // a guess; the exact edges decide
        int k = (q < (double)(nb - 1)) ? (int)q : nb - 1;
        k = max(0, k);
        while (k > 0 && tt < g.edge32(k))
            k--;
        while (k < nb - 1 && tt >= g.edge32(k + 1))
            k++;
        if (tt > last_edge)
            return -1;
        return k;
    };
    // Fast path: q = (t - t0) / delta differs from the position among the floating-point edges by a few ulps of
    // the largest time, so when its fractional part is farther than `eps` from 0 and 1 the bin is floor(q) and
    // the event is on no edge -- no integer <-> double conversions (quarter-rate on the SM), the floor comes
    // from the 2^52 trick.  The rare event near an edge (or outside the interior bins) takes bin_of.
    const double eps = fma((fabs(g.t0) + fabs(last_edge)) * inv_delta, 64.0 * 2.220446049250313e-16, 1e-9);
    const bool fast_ok = eps < 0.25 && nb > 2;
    const double MAGIC = 6755399441055744.0;                // 2^52 + 2^51
    // this thread's events of the first pass (loads issued together)
    double tt[TRIG_EPT];
    int cc[TRIG_EPT];
#pragma unroll
    for (int r = 0; r < TRIG_EPT; r++) {
        const int64_t i = wk.start + r * NT + threadIdx.x;
        const bool in = i < end;
        tt[r] = in ? t[i] : 0.0;
        cc[r] = in ? (int)p[i] : -1;
    }
    if (threadIdx.x == 0) {
        const int ka = bin_of(t[wk.start]);
        int kb = bin_of(t[end - 1]);
        if (kb < 0)
            kb = nb - 1;
        s_k0 = (ka < 0) ? nb : max(0, ka - 1);          // one spare bin below: inclusive dead-time edge
        s_local = (ka >= 0 && kb - s_k0 < TRIG_LOCAL) ? 1 : 0;
    }
    {
        const int c = threadIdx.x;                      // NT == 256 channels values of a byte
        int m = 0;
#pragma unroll
        for (int r = 0; r < 4; r++)
            m |= (c > cb[2 * r] && c < cb[2 * r + 1]) ? (1 << r) : 0;
        s_cm[c] = (uint8_t)m;
    }
    for (int j = threadIdx.x; j < 3 * TRIG_LOCAL; j += NT)
        (&s_h[0][0])[j] = 0;
    __syncthreads();
    const int k0 = s_k0;
    const bool local = s_local != 0;
    for (int pass = 0; pass < TRIG_PASSES; pass++) {
    if (pass > 0) {
#pragma unroll
        for (int r = 0; r < TRIG_EPT; r++) {
            const int64_t i = wk.start + (int64_t)(pass * TRIG_EPT + r) * NT + threadIdx.x;
            const bool in = i < end;
            tt[r] = in ? t[i] : 0.0;
            cc[r] = in ? (int)p[i] : -1;
        }
    }
#pragma unroll
    for (int r = 0; r < TRIG_EPT; r++) {
        const int c = cc[r];
        if (c < 0)
            continue;
        int k;
        bool also_prev = false;
        {
            const double q = (tt[r] - g.t0) * inv_delta;
            const double m = q + MAGIC;
            k = __double2loint(m);                          // round-to-nearest(q) for 0 <= q < 2^31
            double fr = q - (m - MAGIC);                    // in [-0.5, 0.5]
            if (fr < 0.0) {
                k--;
                fr += 1.0;
            }
            if (!(fast_ok && fr > eps && fr < 1.0 - eps && q > 1.0 && q < (double)(nb - 1))) {
                k = bin_of(tt[r]);
                if (k < 0)
                    continue;
                also_prev = (k > 0) && (tt[r] == g.edge32(k));
            }
        }
        const int m = s_cm[c];
        const int add_a = (m & 1) | ((m & 2) << 15);
        const int add_b = ((m >> 2) & 1) | ((m & 8) << 13);
        const int add_c = 1 | ((c == 127) ? (1 << 16) : 0);
        // dead time: every bin with start <= t <= stop (an event on an inner edge belongs to both neighbours:
        // `also_prev`, only possible on the exact path)
        if (local) {
            const int l = k - k0;
            if (add_a)
                atomicAdd(&s_h[0][l], add_a);
            if (add_b)
                atomicAdd(&s_h[1][l], add_b);
            atomicAdd(&s_h[2][l], add_c);
            if (also_prev)
                atomicAdd(&s_h[2][l - 1], add_c);
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (m & (1 << q))
                    atomicAdd(ub + q * bins_per_unit + k, 1);
            atomicAdd(ub + 4 * bins_per_unit + k, 1);
            if (c == 127)
                atomicAdd(ub + 5 * bins_per_unit + k, 1);
            if (also_prev) {
                atomicAdd(ub + 4 * bins_per_unit + k - 1, 1);
                if (c == 127)
                    atomicAdd(ub + 5 * bins_per_unit + k - 1, 1);
            }
        }
    }
    }
    __syncthreads();
    if (local) {
        // The list is time-sorted, so the tile shares bins with its neighbours only at its two ends: the bin of its
        // first event (and the one below it: the inclusive dead-time edge) with the tiles before it, every bin from
        // the one below the NEXT tile's first event on with the tiles after it.  Those take atomics on the zeroed
        // arrays; every bin strictly in between belongs to this tile alone and is stored plainly, zeros included
        // (coalesced, one store per array and bin instead of up to six atomics).
        const int ka_own = k0 + 2;                              // first bin no earlier tile can touch
        int kb_own = nb;                                        // first bin a later tile can touch
        if (end < n) {
            const int kn = bin_of(t[end]);
            kb_own = (kn < 0) ? nb : max(kn - 1, 0);
        }
        for (int l = threadIdx.x; l < TRIG_LOCAL; l += NT) {
            const int k = k0 + l;
            if (k >= nb)
                break;
            const int v0 = s_h[0][l], v1 = s_h[1][l], v2 = s_h[2][l];
            if (k >= ka_own && k < kb_own) {
                ub[0 * bins_per_unit + k] = v0 & 0xffff;
                ub[1 * bins_per_unit + k] = (int)((unsigned)v0 >> 16);
                ub[2 * bins_per_unit + k] = v1 & 0xffff;
                ub[3 * bins_per_unit + k] = (int)((unsigned)v1 >> 16);
                ub[4 * bins_per_unit + k] = v2 & 0xffff;
                ub[5 * bins_per_unit + k] = (int)((unsigned)v2 >> 16);
            } else if (v0 | v1 | v2) {
                if (v0 & 0xffff)
                    atomicAdd(ub + 0 * bins_per_unit + k, v0 & 0xffff);
                if ((unsigned)v0 >> 16)
                    atomicAdd(ub + 1 * bins_per_unit + k, (int)((unsigned)v0 >> 16));
                if (v1 & 0xffff)
                    atomicAdd(ub + 2 * bins_per_unit + k, v1 & 0xffff);
                if ((unsigned)v1 >> 16)
                    atomicAdd(ub + 3 * bins_per_unit + k, (int)((unsigned)v1 >> 16));
                if (v2 & 0xffff)
                    atomicAdd(ub + 4 * bins_per_unit + k, v2 & 0xffff);
                if ((unsigned)v2 >> 16)
                    atomicAdd(ub + 5 * bins_per_unit + k, (int)((unsigned)v2 >> 16));
            }
        }
    }
}

// The exact significance test of one (window start i, gap) pair, from the 64-bit prefix sums in global memory
// (rare: only windows the fp32 prefilter could not rule out).  sig = (S - a B) / sqrt(a B), a = ex / eb, decided
// without the divisions when clear.  Returns 1 if sig >= thr; *razor is set when sig is within 1e-9 of thr.
__device__ __noinline__ int trig_exact(const int32_t *__restrict__ cr, const double *__restrict__ es, int64_t i,
                                       int64_t n_bkg, int64_t n_gap, int64_t n_src, double thr, int *razor)
{
    const int64_t gs = i + n_bkg + n_gap;
    const double eb = es[i + n_bkg] - es[i], ex = es[gs + n_src] - es[gs];
    const double B = (double)(cr[i + n_bkg] - cr[i]), S = (double)(cr[gs + n_src] - cr[gs]);
    const double thr2 = thr * thr;
    const double d = S * eb - B * ex, rhs = thr2 * B * ex * eb;
    const double lhs = d * fabs(d);
    if (B > 0.0 && eb > 0.0 && ex > 0.0 && thr > 0.0 && lhs < rhs * (1.0 - 1e-8))
        return 0;
    if (B > 0.0 && eb > 0.0 && ex > 0.0 && thr > 0.0 && lhs > rhs * (1.0 + 1e-8))
        return 1;
    const double alpha = ex / eb;
    const double nbk = alpha * B;
    const double sig = (S - nbk) / sqrt(nbk);
    if (fabs(sig - thr) < 1e-9 * thr)
        *razor = 1;
    return sig >= thr ? 1 : 0;
}

// One (window start, time scale) pair that survived the source-window filter: the fp32 test on the staged prefix
// sums (B from the global 32-bit prefix: the staged one has been turned into window sums by then), then the exact
// one.  Out of line: a few windows per light curve come here.  Returns 1 on a hit.
__device__ __noinline__ int trig_detail(const int32_t *__restrict__ cr, const double *__restrict__ es,
                                        const float *s_ef, int li, int64_t i, int nbk, int nsr, int n_gap, float Sf,
                                        bool use_pre, float min_eb, float min_ex, float thr2f, double thr, int *razor)
{
    const float *pe = s_ef + li + nbk;
    const float ebf = pe[0] - s_ef[li];
    const float Bf = (float)(cr[i + nbk] - cr[i]);
    if (use_pre && Bf > 0.0f && ebf > min_eb) {
        const float exf = pe[nsr + n_gap] - pe[n_gap];
        const float d32 = Sf * ebf - Bf * exf;
        if (exf > min_ex && (d32 <= 0.0f || d32 * d32 < 0.9f * thr2f * Bf * ebf * exf))
            return 0;
    }
    return trig_exact(cr, es, i, nbk, n_gap, nsr, thr, razor);
}

#define TRIG_LOSS_N 4096
#ifndef TRIG_PB
#define TRIG_PB 4               // bins per thread and round in the prefix pass
#endif
#ifndef TRIG_EVAL_CTAS
#define TRIG_EVAL_CTAS 4
#endif

// one CTA per unit: prefix sums over the bins, then the sliding-window search in the reference's order
__global__ void __launch_bounds__(NT, TRIG_EVAL_CTAS) k_trig_eval(const cgrb_unit_t *__restrict__ units, int n_units,
                                                  cgrb_trigger_par_t par, const double *__restrict__ times_out,
                                                  const cgrb_counts_t *__restrict__ counts, int64_t bins_per_unit,
                                                  const int32_t *__restrict__ bins, int64_t cstride,
                                                  int32_t *__restrict__ csum, double *__restrict__ esum,
                                                  cgrb_trigger_t *__restrict__ out)
{
    const int ui = blockIdx.x;
    if (ui >= n_units)
        return;
#ifdef TRIG_TIMING
    long long tacc[8] = {0}, tlast = clock64();
#define TT_MARK(i) do { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } while (0)
#else
#define TT_MARK(i)
#endif
    const cgrb_unit_t *u = units + ui;
    const int64_t n = counts[ui].n_kept;
    cgrb_trigger_t res;
    res.detected = 0;
    res.combo = -1;
    res.n_bins = 0;
    res.razor = 0;
    res.time = 0.0;
    res.time_scale = 0.0;
    // LightCurveAnalyzer.__init__ (lightcurve_analyzer.py:22-34): nothing to do without source counts
    if (counts[ui].n_src <= 0 || n < 2) {
        if (threadIdx.x == 0)
            out[ui] = res;
        return;
    }
    const double *t = times_out + u->tot_off;
    const double dt = par.base_timescale;
    const TrigGrid g = trig_grid(t[0], t[n - 1], dt);
    const int64_t nb = g.n_edges - 1;
    res.n_bins = (int32_t)max((int64_t)0, min(nb, (int64_t)0x7fffffff));
    if (nb < 1 || nb >= 0x7fffffff || n > 0x7fffffff) {      // (the count prefix sums are 32-bit: <= n_kept)
        if (threadIdx.x == 0) {
            if (n > 0x7fffffff)
                res.razor = 2;
            out[ui] = res;
        }
        return;
    }
    const int32_t *ub = bins + (size_t)ui * TRIG_NARR * bins_per_unit;
    // prefix arrays: entry k = sum over the bins < k; entry 0 sits three elements into a 16-byte aligned row of
    // `cstride` elements (a multiple of 4), so that entries k + 1 .. k + 4 of a thread's four bins are one aligned vector
    int32_t *cs = csum + (size_t)ui * 4 * cstride + 3;
    double *es = esum + (size_t)ui * cstride + 3;
    const bool vec = (bins_per_unit & 3) == 0;          // the bin arrays of a unit are 16-byte aligned rows as well
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // prefix sums: counts after the reference's float round trip, exposures = (stop - start) - dead time.
    // Every thread owns TRIG_PB consecutive bins per round (serial prefix inside), the thread totals of all
    // five quantities go through one block scan: warp scans by shuffles, the 8 warp totals exchanged through
    // double-buffered shared memory (one barrier per round).  Fixed association, so the sums are reproducible.
    __shared__ double s_pw[2][NW];
    __shared__ int s_pi[2][4][NW];
    // (c / dt) * dt truncated loses one count for a fixed set of c: one bit per c < 4096, built once per CTA
    // (16 divisions per thread instead of 4 per bin; 64-bit conversions and divisions are quarter-rate)
    __shared__ unsigned s_loss[TRIG_LOSS_N / 32];
    if (threadIdx.x < TRIG_LOSS_N / 32)
        s_loss[threadIdx.x] = 0u;
    __syncthreads();
    for (int c = threadIdx.x; c < TRIG_LOSS_N; c += NT)
        if ((int)(__dmul_rn(__ddiv_rn((double)c, dt), dt)) < c)
            atomicOr(&s_loss[c >> 5], 1u << (c & 31));
    __syncthreads();
    TT_MARK(0);
    int carry_c[4] = {0, 0, 0, 0};
    double carry_e = 0.0;
    if (threadIdx.x == 0) {
        for (int r = 0; r < 4; r++)
            cs[r * cstride] = 0;
        es[0] = 0.0;
    }
    int par_ = 0;
    for (int64_t base = 0; base < nb; base += TRIG_PB * NT, par_ ^= 1) {
        const int64_t kf = base + (int64_t)TRIG_PB * threadIdx.x;      // this thread's first bin
        int c[4][TRIG_PB];
        double e[TRIG_PB];
        int na[TRIG_PB], no[TRIG_PB];
        if (TRIG_PB == 4 && vec && kf + 3 < bins_per_unit) {
            // bins at and beyond nb are zero (never written), so whole vectors can be read
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int4 v = *reinterpret_cast<const int4 *>(ub + r * bins_per_unit + kf);
                c[r][0] = v.x, c[r][1] = v.y, c[r][2] = v.z, c[r][3] = v.w;
            }
            const int4 va = *reinterpret_cast<const int4 *>(ub + 4 * bins_per_unit + kf);
            const int4 vo = *reinterpret_cast<const int4 *>(ub + 5 * bins_per_unit + kf);
            na[0] = va.x, na[1] = va.y, na[2] = va.z, na[3] = va.w;
            no[0] = vo.x, no[1] = vo.y, no[2] = vo.z, no[3] = vo.w;
        } else {
#pragma unroll
            for (int j = 0; j < TRIG_PB; j++) {
                const bool in = kf + j < nb;
#pragma unroll
                for (int r = 0; r < 4; r++)
                    c[r][j] = in ? ub[r * bins_per_unit + kf + j] : 0;
                na[j] = in ? ub[4 * bins_per_unit + kf + j] : 0;
                no[j] = in ? ub[5 * bins_per_unit + kf + j] : 0;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < TRIG_PB; j++) {
                const int raw = c[r][j];
                // (rate * dt).astype(int)
                c[r][j] = ((unsigned)raw < TRIG_LOSS_N) ? raw - (int)((s_loss[raw >> 5] >> (raw & 31)) & 1u)
                                                       : (int)(__dmul_rn(__ddiv_rn((double)raw, dt), dt));
            }
#pragma unroll
        for (int j = 0; j < TRIG_PB; j++) {
            e[j] = 0.0;
            if (kf + j < nb) {
                const double dead = __dadd_rn(__dmul_rn(trig_u2d((unsigned)(na[j] - no[j])), 2.0e-6),
                                              __dmul_rn(trig_u2d((unsigned)no[j]), 10.0e-6));
                const int kk = (int)(kf + j);
                e[j] = __dadd_rn(__dadd_rn(g.edge32(kk + 1), -g.edge32(kk)), -dead);
            }
        }
        // serial inclusive prefix inside the thread
#pragma unroll
        for (int j = 1; j < TRIG_PB; j++) {
            e[j] += e[j - 1];
#pragma unroll
            for (int r = 0; r < 4; r++)
                c[r][j] += c[r][j - 1];
        }
        const double te = e[TRIG_PB - 1];
        const double ie = warp_incl_scan(te, lane);
        int ic[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            ic[r] = warp_incl_scan_i(c[r][TRIG_PB - 1], lane);
        if (lane == 31) {
            s_pw[par_][wid] = ie;
#pragma unroll
            for (int r = 0; r < 4; r++)
                s_pi[par_][r][wid] = ic[r];
        }
        __syncthreads();
        double before_e = 0.0, tot_e = 0.0;
        int before_c[4] = {0, 0, 0, 0}, tot_c[4] = {0, 0, 0, 0};     // one round: 1024 bins, far below 2^31
#pragma unroll
        for (int q = 0; q < NW; q++) {
            const double we = s_pw[par_][q];
            if (q < wid)
                before_e += we;
            tot_e += we;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int wc = s_pi[par_][r][q];
                if (q < wid)
                    before_c[r] += wc;
                tot_c[r] += wc;
            }
        }
        const double off_e = carry_e + (before_e + (ie - te));
        const bool whole = TRIG_PB == 4 && kf + 3 < nb;     // all four bins exist: aligned vector stores
        if (whole) {
            double2 *d = reinterpret_cast<double2 *>(es + kf + 1);
            d[0] = make_double2(off_e + e[0], off_e + e[1]);
            d[1] = make_double2(off_e + e[2], off_e + e[3]);
        } else {
#pragma unroll
            for (int j = 0; j < TRIG_PB; j++)
                if (kf + j < nb)
                    es[kf + j + 1] = off_e + e[j];
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int off_c = carry_c[r] + (before_c[r] + ic[r] - c[r][TRIG_PB - 1]);
            if (whole)
                *reinterpret_cast<int4 *>(cs + r * cstride + kf + 1) =
                    make_int4(off_c + c[r][0], off_c + c[r][1], off_c + c[r][2], off_c + c[r][3]);
            else {
#pragma unroll
                for (int j = 0; j < TRIG_PB; j++)
                    if (kf + j < nb)
                        cs[r * cstride + kf + j + 1] = off_c + c[r][j];
            }
            carry_c[r] += tot_c[r];
        }
        carry_e += tot_e;
    }
    __syncthreads();
    TT_MARK(1);
    // the search: energy ranges a..d, time scales longest first; the first combination (in that order) with
    // a window at or above the threshold wins, at its first window.  Chunk by chunk of TRIG_CH window starts
    // (+ the longest window span): the exposure prefix is staged once per chunk as fp32 relative to the chunk
    // start and shared by the four ranges; the 32-bit count prefix of one range at a time arrives by cp.async
    // while the previous range is being searched.
    extern __shared__ __align__(16) unsigned char trig_dyn[];
    float *s_ef = reinterpret_cast<float *>(trig_dyn);      // exposure prefix of the chunk
    float *s_sw = s_ef + TRIG_ST;                           // source-window count sums S(m) = C(m + n_src) - C(m)
    int *s_raw = reinterpret_cast<int *>(s_sw + TRIG_ST);   // count prefix of the range in flight
    float *s_need = reinterpret_cast<float *>(s_raw + TRIG_ST);    // per window start: the S a hit needs at least
    __shared__ int s_hit[4][10];            // first firing window start per combination (n_bins < 2^31)
    __shared__ float s_min[NW];
    const int n_scales[4] = {10, 10, 9, 4};
    const int64_t n_bkg = (int64_t)floor(par.background_duration / dt);
    const int64_t n_src = (int64_t)floor(par.pre_window / dt);          // the "source" window (swapped arguments)
    const double thr = par.threshold;
    int razor = 0;
    // the staged layout needs n_bkg + longest gap + n_src <= TRIG_SPAN: true for the reference's constants
    const bool staged_ok = (n_bkg + 512 + n_src <= TRIG_SPAN) && n_bkg > 0 && n_src > 0;
    const bool pre_params_ok = (thr >= 1.0) && (thr < 1e15) && (n_src <= n_bkg);
    const float thr2f = (float)(thr * thr);
    const float thr_lo = 0.94f * (float)thr;                            // < sqrt(0.9) thr: the margin of the fp32 test
    const float min_eb = 0.25f * (float)((double)n_bkg * dt), min_ex = 0.25f * (float)((double)n_src * dt);
    const int nbk = (int)n_bkg, nsr = (int)n_src;
    constexpr int TRIG_IPT = TRIG_CH / NT;                              // window starts per thread and chunk
    constexpr int TRIG_XPT = (TRIG_ST + NT - 1) / NT;                   // staged elements per thread
    if (threadIdx.x < 40)
        (&s_hit[0][0])[threadIdx.x] = INT_MAX;
    unsigned range_ok = 0u;                                             // `if sum(counts) == 0: continue`
#pragma unroll
    for (int r = 0; r < 4; r++)
        range_ok |= (cs[r * cstride + nb] != 0) ? (1u << r) : 0u;
    const int64_t n_i_max = nb - (n_bkg + 1 + n_src);                   // shortest gap: 1 bin
    const int n_chunks = (staged_ok && n_i_max > 0) ? (int)((n_i_max + TRIG_CH - 1) / TRIG_CH) : 0;
    // a combination is identified by key = 16 r + (index of the time scale, longest first): the smallest key with a
    // hit wins.  Later chunks only hold later windows, so once a key has fired only smaller keys are still open.
    int best_key = 64;
    auto next_active = [&](int it) -> int {                             // iteration = 4 chunk + range
        for (int nx = it + 1; nx < 4 * n_chunks; nx++) {
            const int r = nx & 3;
            if (((range_ok >> r) & 1u) && 16 * r < best_key)
                return nx;
        }
        return -1;
    };
    auto prefetch = [&](int nx) {
        const int64_t c0 = (int64_t)(nx >> 2) * TRIG_CH;
        const int n_st = (int)(min(nb, c0 + TRIG_CH + TRIG_SPAN) - c0) + 1;
        const int32_t *src = cs + (nx & 3) * cstride + c0;
        for (int k = threadIdx.x; k < n_st; k += NT) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_raw + k);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + k) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int it = next_active(-1);
    if (it >= 0)
        prefetch(it);
    int staged_chunk = -1;
    float exm = INFINITY;
    while (it >= 0) {
        const int r = it & 3;
        const int64_t c0 = (int64_t)(it >> 2) * TRIG_CH;
        const int64_t hi = min(nb, c0 + TRIG_CH + TRIG_SPAN);
        const int n_st = (int)(hi - c0) + 1;
        const int32_t *cr = cs + r * cstride;
        const int ns = n_scales[r];
        if (staged_chunk != (it >> 2)) {
            staged_chunk = it >> 2;
            // exposure prefix of the chunk (nobody reads s_ef any more: the barrier that closed the last search)
            const double ebase = es[c0];
            const double *ep = es + c0;
            for (int k = threadIdx.x; k < n_st; k += 4 * NT) {
                double ev[4];                           // loads batched four deep
#pragma unroll
                for (int q = 0; q < 4; q++)
                    ev[q] = ep[min(k + q * NT, n_st - 1)];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (k + q * NT < n_st)
                        s_ef[k + q * NT] = (float)(ev[q] - ebase);
            }
            __syncthreads();
            // ex_min: the smallest source-window exposure staged in this chunk
            float e_ = INFINITY;
#pragma unroll
            for (int q = 0; q < TRIG_XPT; q++) {
                const int m = threadIdx.x + q * NT;
                if (m >= nbk && m + nsr < n_st)
                    e_ = fminf(e_, s_ef[m + nsr] - s_ef[m]);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
                e_ = fminf(e_, __shfl_xor_sync(0xffffffffu, e_, d));
            if (lane == 0)
                s_min[wid] = e_;
            __syncthreads();
            exm = s_min[0];
#pragma unroll
            for (int q = 1; q < NW; q++)
                exm = fminf(exm, s_min[q]);
            TT_MARK(2);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                                // the count prefix of (chunk, range) has landed
        TT_MARK(3);
        const bool open = 16 * r < best_key;            // (a prefetch issued before the last hit may be moot)
        const int i_end = (int)(min(c0 + (int64_t)TRIG_CH, n_i_max) - c0);
        // Nearly every window is far below the threshold, which is decided in fp32 (full rate, no 64-bit
        // conversions; counts below 2^24 are exact).  Level 0 works on the source-window sums alone: a hit needs
        //     S >= a B + thr sqrt(a B),   a = ex / eb,
        // and the right-hand side grows with a, so with ex_min as above no window starting at i can fire unless
        // S >= S_min(i) = a_min B + 0.94 thr sqrt(a_min B), a_min = ex_min / eb(i) (0.94 < sqrt(0.9): the margin
        // of the fp32 test in trig_detail, which every survivor takes next, and then the exact one).  One load
        // and one compare per (window, time scale).
        const bool use_pre = pre_params_ok && (s_raw[n_st - 1] - s_raw[0]) < (1 << 24);
        const bool lvl0 = use_pre && exm > min_ex && exm < INFINITY;
        if (open) {
#pragma unroll
            for (int q = 0; q < TRIG_XPT; q++) {
                const int m = threadIdx.x + q * NT;
                if (m + nsr < n_st)
                    s_sw[m] = (float)min(s_raw[m + nsr] - s_raw[m], 1 << 24);
            }
#pragma unroll
            for (int q = 0; q < TRIG_IPT; q++) {
                const int li = threadIdx.x + q * NT;
                float need = -INFINITY;
                if (lvl0 && li < i_end) {
                    const float Bf = (float)(s_raw[li + nbk] - s_raw[li]);
                    const float ebf = s_ef[li + nbk] - s_ef[li];
                    if (Bf > 0.0f && ebf > min_eb) {
                        const float ab = Bf / ebf * exm;        // a_min B
                        need = ab * 0.99999f + thr_lo * sqrtf(ab);
                    }
                }
                s_need[li] = need;
            }
        }
        __syncthreads();                                // s_raw is free again
        TT_MARK(4);
        const int nx = next_active(it);
        if (nx >= 0)
            prefetch(nx);
        if (open) {
            // window i with gap n_gap = 2^gx exists iff i + n_bkg + n_gap + n_src < nb, i.e. n_gap < room(i); the time
            // scales still open in this range are those with an index ns - 1 - gx below s_lim
            const int64_t room0 = nb - n_bkg - n_src - c0;
            const int s_lim = best_key - 16 * r;
            const unsigned open_gx = ((1u << ns) - 1u) & ~((1u << max(ns - s_lim, 0)) - 1u);
#pragma unroll 1
            for (int li = threadIdx.x; li < i_end; li += NT) {
                const int room = (int)min(room0 - li, (int64_t)1024);
                // gaps that exist: 2^gx < room  <=>  gx < ceil(log2(room))
                const unsigned fit = (room > 1) ? ((1u << (32 - __clz(room - 1))) - 1u) : 0u;
                const float need = s_need[li];
                const float *ps = s_sw + li + nbk;
                // all ten source-window sums at once (in bounds of the staged array for every gap), one bit per
                // time scale that could fire: no branch until something does
                unsigned cand = 0u;
#pragma unroll
                for (int gx = 0; gx < 10; gx++)
                    cand |= (ps[1 << gx] >= need) ? (1u << gx) : 0u;
                cand &= open_gx & fit;
                while (cand) {
                    const int gx = 31 - __clz(cand);            // longest time scale first
                    cand &= ~(1u << gx);
                    const int n_gap = 1 << gx;
                    const int64_t i = c0 + li;
                    if (i >= (int64_t)((volatile int *)s_hit[r])[ns - 1 - gx])
                        continue;                       // an earlier window of this combination has fired already
                    if (trig_detail(cr, es, s_ef, li, i, nbk, nsr, n_gap, ps[n_gap], use_pre, min_eb, min_ex, thr2f, thr, &razor) &&
                        g.edge(i + n_bkg + n_gap) > -1.0)
                        atomicMin(&s_hit[r][ns - 1 - gx], (int)i);
                }
            }
        }
        __syncthreads();
        TT_MARK(5);
        for (int key = 16 * r; key < min(best_key, 16 * r + ns); key++)
            if (s_hit[r][key - 16 * r] != INT_MAX) {
                best_key = key;
                break;
            }
        if (best_key == 0)
            break;                  // the first combination fired: nothing later can overrule it
        it = nx;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (best_key < 64) {
        const int r = best_key >> 4, sidx = best_key & 15;
        const int64_t n_gap = (int64_t)1 << (n_scales[r] - 1 - sidx);
        res.detected = 1;
        res.combo = best_key;
        res.time = g.edge((int64_t)s_hit[r][sidx] + n_bkg + n_gap);
        res.time_scale = dt * (double)n_gap;
    }
#ifdef TRIG_TIMING
    if (threadIdx.x == 0 && (blockIdx.x % 400) == 7)
        printf("TT unit %d nb %d: setup %lld prefix %lld | ef-stage %lld wait-counts %lld convert %lld search %lld\n", ui, (int)nb,
               tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5]);
#endif
    __shared__ int s_razor;
    if (threadIdx.x == 0)
        s_razor = 0;
    __syncthreads();
    if (razor)
        atomicOr(&s_razor, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        res.razor = s_razor | (staged_ok ? 0 : 2);
        out[ui] = res;
    }
}

// BackgroundSpectrumTemplate.sample_channel(size) on its own (background.py:80-93): n draws from the
// detector's channel cdf; Philox index = event index (stage ST_BKG, sub-stream 1), or injected u[i].
