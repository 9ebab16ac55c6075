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
#define TRIG_TILE 2048
#define TRIG_LOCAL 256          // bins a tile may span and still be accumulated in shared memory
#define TRIG_NCOMBO 33
#define TRIG_NARR 6             // per-bin integer arrays: counts of the 4 ranges, events, overflow events
#define TRIG_CH 2048            // windows per staged chunk of the search
#define TRIG_SPAN 1832          // >= background (1062) + longest gap (512) + source (250) bins

// np.arange(tmin, tmax, dt): length and the fill rule  edge[i] = start + i * (fl(start + dt) - start)
struct TrigGrid {
    double t0, delta;
    int64_t n_edges;
    __device__ __forceinline__ double edge(int64_t i) const { return __dadd_rn(t0, __dmul_rn((double)i, delta)); }
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

__global__ void __launch_bounds__(NT) k_trig_bin(const cgrb_unit_t *__restrict__ units,
                                                 const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                 const int32_t *__restrict__ chan_bounds, double dt,
                                                 const double *__restrict__ times_out,
                                                 const uint8_t *__restrict__ pha_out,
                                                 const cgrb_counts_t *__restrict__ counts, int64_t bins_per_unit,
                                                 int32_t *__restrict__ bins)
{
    __shared__ int s_h[TRIG_NARR][TRIG_LOCAL];
    __shared__ int64_t s_k0;
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
    const int64_t nb = g.n_edges - 1;                   // np.histogram bins
    if (nb < 1)
        return;
    const int64_t end = min(wk.start + (int64_t)TRIG_TILE, n);
    int32_t *ub = bins + (size_t)wk.unit * TRIG_NARR * bins_per_unit;
    const int32_t *cb = chan_bounds + 8 * u->det;

    // bin of an event: edge[k] <= t < edge[k+1] (last bin closed), -1 beyond the last edge
    const double inv_delta = 1.0 / g.delta;
    auto bin_of = [&](double tt) -> int64_t {
        int64_t k = (int64_t)floor((tt - g.t0) * inv_delta);       // a guess; the exact edges decide
        k = max((int64_t)0, min(k, nb - 1));
        while (k > 0 && tt < g.edge(k))
            k--;
        while (k < nb - 1 && tt >= g.edge(k + 1))
            k++;
        if (tt > g.edge(nb))
            return -1;
        return k;
    };
    if (threadIdx.x == 0) {
        const int64_t ka = bin_of(t[wk.start]);
        int64_t kb = bin_of(t[end - 1]);
        if (kb < 0)
            kb = nb - 1;
        s_k0 = (ka < 0) ? nb : max((int64_t)0, ka - 1);         // one spare bin below: inclusive dead-time edge
        s_local = (ka >= 0 && kb - s_k0 < TRIG_LOCAL) ? 1 : 0;
    }
    for (int j = threadIdx.x; j < TRIG_NARR * TRIG_LOCAL; j += NT)
        (&s_h[0][0])[j] = 0;
    __syncthreads();
    const int64_t k0 = s_k0;
    const bool local = s_local != 0;
    for (int64_t i = wk.start + threadIdx.x; i < end; i += NT) {
        const double tt = t[i];
        const int c = p[i];
        const int64_t k = bin_of(tt);
        if (k < 0)
            continue;
        const int ovf = (c == 127) ? 1 : 0;
        // dead time: every bin with start <= t <= stop (an event on an inner edge belongs to both neighbours)
        const bool also_prev = (k > 0) && (tt == g.edge(k));
        if (local) {
            const int l = (int)(k - k0);
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (c > cb[2 * r] && c < cb[2 * r + 1])
                    atomicAdd(&s_h[r][l], 1);
            atomicAdd(&s_h[4][l], 1);
            if (ovf)
                atomicAdd(&s_h[5][l], 1);
            if (also_prev) {
                atomicAdd(&s_h[4][l - 1], 1);
                if (ovf)
                    atomicAdd(&s_h[5][l - 1], 1);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (c > cb[2 * r] && c < cb[2 * r + 1])
                    atomicAdd(ub + r * bins_per_unit + k, 1);
            atomicAdd(ub + 4 * bins_per_unit + k, 1);
            if (ovf)
                atomicAdd(ub + 5 * bins_per_unit + k, 1);
            if (also_prev) {
                atomicAdd(ub + 4 * bins_per_unit + k - 1, 1);
                if (ovf)
                    atomicAdd(ub + 5 * bins_per_unit + k - 1, 1);
            }
        }
    }
    __syncthreads();
    if (local)
        for (int j = threadIdx.x; j < TRIG_NARR * TRIG_LOCAL; j += NT) {
            const int a = j / TRIG_LOCAL, l = j % TRIG_LOCAL;
            const int v = s_h[a][l];
            if (v && k0 + l < nb)
                atomicAdd(ub + a * bins_per_unit + k0 + l, v);
        }
}

// one CTA per unit: prefix sums over the bins, then the sliding-window search in the reference's order
__global__ void __launch_bounds__(NT, 4) k_trig_eval(const cgrb_unit_t *__restrict__ units, int n_units,
                                                  cgrb_trigger_par_t par, const double *__restrict__ times_out,
                                                  const cgrb_counts_t *__restrict__ counts, int64_t bins_per_unit,
                                                  const int32_t *__restrict__ bins, int64_t *__restrict__ csum,
                                                  double *__restrict__ esum, cgrb_trigger_t *__restrict__ out)
{
    const int ui = blockIdx.x;
    if (ui >= n_units)
        return;
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
    res.n_bins = (int32_t)max((int64_t)0, nb);
    if (nb < 1) {
        if (threadIdx.x == 0)
            out[ui] = res;
        return;
    }
    const int32_t *ub = bins + (size_t)ui * TRIG_NARR * bins_per_unit;
    int64_t *cs = csum + (size_t)ui * 4 * (bins_per_unit + 1);
    double *es = esum + (size_t)ui * (bins_per_unit + 1);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // prefix sums: counts after the reference's float round trip, exposures = (stop - start) - dead time.
    // One block scan per 256 bins carries all five quantities: warp scans by shuffles, the 8 warp totals
    // exchanged through double-buffered shared memory (ONE barrier per round), the next round's bins
    // prefetched before the barrier.  Fixed association, so the sums are reproducible.
    __shared__ double s_pw[2][NW];
    __shared__ int s_pi[2][4][NW];
    long long carry_c[4] = {0, 0, 0, 0};
    double carry_e = 0.0;
    if (threadIdx.x == 0) {
        for (int r = 0; r < 4; r++)
            cs[r * (bins_per_unit + 1)] = 0;
        es[0] = 0.0;
    }
    int raw[TRIG_NARR];
#pragma unroll
    for (int a = 0; a < TRIG_NARR; a++)
        raw[a] = (threadIdx.x < nb) ? ub[a * bins_per_unit + threadIdx.x] : 0;
    int par_ = 0;
    for (int64_t base = 0; base < nb; base += NT, par_ ^= 1) {
        const int64_t k = base + threadIdx.x;
        const bool in = k < nb;
        int c[4];
        double e = 0.0;
#pragma unroll
        for (int r = 0; r < 4; r++)
            c[r] = (int)(__dmul_rn(__ddiv_rn((double)raw[r], dt), dt));         // (rate * dt).astype(int)
        if (in) {
            const int na = raw[4], no = raw[5];
            const double dead = __dadd_rn(__dmul_rn((double)(na - no), 2.0e-6), __dmul_rn((double)no, 10.0e-6));
            e = __dadd_rn(__dadd_rn(g.edge(k + 1), -g.edge(k)), -dead);
        }
        // prefetch the next round
        const int64_t kn = k + NT;
#pragma unroll
        for (int a = 0; a < TRIG_NARR; a++)
            raw[a] = (kn < nb) ? ub[a * bins_per_unit + kn] : 0;
        e = warp_incl_scan(e, lane);
#pragma unroll
        for (int r = 0; r < 4; r++)
            c[r] = warp_incl_scan_i(c[r], lane);
        if (lane == 31) {
            s_pw[par_][wid] = e;
#pragma unroll
            for (int r = 0; r < 4; r++)
                s_pi[par_][r][wid] = c[r];
        }
        __syncthreads();
        double before_e = 0.0, tot_e = 0.0;
        int before_c[4] = {0, 0, 0, 0}, tot_c[4] = {0, 0, 0, 0};     // one round: 256 bins, far below 2^31
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
        if (in) {
            es[k + 1] = carry_e + (before_e + e);
#pragma unroll
            for (int r = 0; r < 4; r++)
                cs[r * (bins_per_unit + 1) + k + 1] = carry_c[r] + (long long)(before_c[r] + c[r]);
        }
        carry_e += tot_e;
#pragma unroll
        for (int r = 0; r < 4; r++)
            carry_c[r] += tot_c[r];
    }
    __syncthreads();
    // the search: energy ranges a..d, time scales longest first; the first combination (in that order) with
    // a window at or above the threshold wins, at its first window.  The prefix sums of a chunk of windows
    // (+ the longest window span) are staged in shared memory; all time scales of a range share the
    // background window of a given i.
    __shared__ float s_ef[TRIG_CH + TRIG_SPAN + 1];     // exposure prefix relative to the chunk start, fp32
    __shared__ int s_c[TRIG_CH + TRIG_SPAN + 1];
    __shared__ long long s_hit[10];
    const int n_scales[4] = {10, 10, 9, 4};
    const int64_t n_bkg = (int64_t)floor(par.background_duration / dt);
    const int64_t n_src = (int64_t)floor(par.pre_window / dt);          // the "source" window (swapped arguments)
    const double thr = par.threshold, thr2 = thr * thr;
    int razor = 0;
    bool found = false;
    // the staged layout needs n_bkg + longest gap + n_src <= TRIG_SPAN: true for the reference's constants
    const bool staged_ok = (n_bkg + 512 + n_src <= TRIG_SPAN) && n_bkg > 0 && n_src > 0;
    const bool pre_params_ok = (thr >= 1.0) && (thr < 1e15) && (n_src <= n_bkg);
    const float thr2f = (float)thr2;
    const float min_eb = 0.25f * (float)((double)n_bkg * dt), min_ex = 0.25f * (float)((double)n_src * dt);
    for (int r = 0; r < 4 && !found && staged_ok; r++) {
        const int64_t *cr = cs + r * (bins_per_unit + 1);
        if (cr[nb] == 0)                                                  // `if sum(counts) == 0: continue`
            continue;
        const int ns = n_scales[r];
        if (threadIdx.x < 10)
            s_hit[threadIdx.x] = LLONG_MAX;
        const int64_t n_i_max = nb - (n_bkg + 1 + n_src);                 // shortest gap: 1 bin
        for (int64_t c0 = 0; c0 < n_i_max; c0 += TRIG_CH) {
            __syncthreads();
            const int64_t hi = min(nb, c0 + TRIG_CH + TRIG_SPAN);
            const long long cbase = cr[c0];
            const double ebase = es[c0];
            for (int64_t k = c0 + threadIdx.x; k <= hi; k += NT) {
                s_ef[k - c0] = (float)(es[k] - ebase);
                s_c[k - c0] = (int)(cr[k] - cbase);
            }
            // fp32 prefilter (see below) needs exact int -> float counts and a sane threshold / window pair
            const bool use_pre = pre_params_ok && (cr[hi] - cbase) < (1ll << 24);
            __syncthreads();
            for (int64_t i = c0 + threadIdx.x; i < min(c0 + (int64_t)TRIG_CH, n_i_max); i += NT) {
                const int li = (int)(i - c0);
                const int Bi = s_c[li + n_bkg] - s_c[li];
                const float ebf = s_ef[li + n_bkg] - s_ef[li], Bf = (float)Bi;
                const bool pre_i = use_pre && Bi > 0 && ebf > min_eb;
                const double B = (double)Bi;
                for (int sidx = 0; sidx < ns; sidx++) {
                    const int64_t n_gap = (int64_t)1 << (ns - 1 - sidx);
                    if (i >= nb - (n_bkg + n_gap + n_src))                // i + span < len(stops)
                        continue;
                    const int ls = li + (int)(n_bkg + n_gap);
                    const int Si = s_c[ls + n_src] - s_c[ls];
                    if (pre_i) {
                        // Nearly every window is far below the threshold: decide that in fp32 (full rate, no
                        // 64-bit conversions).  A hit needs d = S eb - B ex >= thr sqrt(B ex eb); the fp32 value of
                        // d is off by at most ~3e-6 (S eb + B ex), i.e. by < 2.5 % of d for any window that could
                        // fire (counts < 2^24, thr >= 1, source window not longer than the background window),
                        // so d^2 < 0.9 thr^2 B ex eb is a certain miss.  Everything else takes the exact path.
                        const float exf = s_ef[ls + n_src] - s_ef[ls];
                        if (exf > min_ex) {
                            const float d32 = (float)Si * ebf - Bf * exf;
                            if (d32 <= 0.0f || d32 * d32 < 0.9f * (thr2f * Bf * exf * ebf))
                                continue;
                        }
                    }
                    const int64_t gs = i + n_bkg + n_gap;
                    const double eb = es[i + n_bkg] - es[i];
                    const double ex = es[gs + n_src] - es[gs];
                    const double S = (double)Si;
                    // sig = (S - a B) / sqrt(a B), a = ex / eb: decided without the divisions when clear
                    bool hit;
                    const double d = S * eb - B * ex, rhs = thr2 * B * ex * eb;
                    const double lhs = d * fabs(d);
                    if (B > 0.0 && eb > 0.0 && ex > 0.0 && thr > 0.0 && lhs < rhs * (1.0 - 1e-8))
                        hit = false;
                    else if (B > 0.0 && eb > 0.0 && ex > 0.0 && thr > 0.0 && lhs > rhs * (1.0 + 1e-8))
                        hit = true;
                    else {
                        const double alpha = ex / eb;
                        const double nbk = alpha * B;
                        const double sig = (S - nbk) / sqrt(nbk);
                        if (fabs(sig - thr) < 1e-9 * thr)
                            razor = 1;
                        hit = sig >= thr;
                    }
                    if (hit && g.edge(i + n_bkg + n_gap) > -1.0)
                        atomicMin(&s_hit[sidx], (long long)i);
                }
            }
            __syncthreads();
            if (s_hit[0] != LLONG_MAX)
                break;              // the first time scale fired: nothing later can overrule it
        }
        __syncthreads();
        for (int sidx = 0; sidx < ns; sidx++) {
            const long long first = s_hit[sidx];
            if (first != LLONG_MAX) {
                const int64_t n_gap = (int64_t)1 << (ns - 1 - sidx);
                found = true;
                res.detected = 1;
                res.combo = r * 16 + sidx;
                res.time = g.edge(first + n_bkg + n_gap);
                res.time_scale = dt * (double)n_gap;
                break;
            }
        }
        __syncthreads();
    }
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
