// cgrb_energy_kernels.cuh -- photon energies: literal rejection sampler and the tabulated-envelope sampler with the fused fold
// Part of the single translation unit cgrb_kernels.cu (included in order; see that file).
#pragma once

// ==========================================================================================
// photon energies (cpl_functions.py:191-286), literal sampler: one warp per photon.  The 500-point
// envelope maximum is found on log f (monotone in f, no exp); the rejection loop runs 32 trials per
// round, one per lane, and the lowest accepted trial wins, which is the serial loop's result.
// ==========================================================================================
struct EnergyTab {
    double gx[CGRB_NE_ENVELOPE];   // x_m = E_m (1+z)
    double gl[CGRB_NE_ENVELOPE];   // log(A_m) + alpha log(x_m)   (-inf where A_m <= 0)
    double eax[CGRB_NBINS];
    double eay[CGRB_NBINS];
};

__device__ __forceinline__ EaGuess make_ea_guess(const double *eax)
{
    EaGuess g;
    g.log_x0 = log(eax[0]);
    g.inv_dlog = (double)(CGRB_NBINS - 1) / (log(eax[CGRB_NBINS - 1]) - g.log_x0);
    return g;
}

__device__ __forceinline__ PropConst make_prop_const(double emin, double emax, double alpha)
{
    // power-law proposal constants (cpl_functions.py:254-258)
    PropConst pc;
    const double ap1 = alpha + 1.0;
    pc.q_lo = pow(emin, ap1);
    pc.p_span = pow(emax, ap1) - pc.q_lo;
    pc.inv_ap1 = 1.0 / ap1;
    return pc;
}

__global__ void __launch_bounds__(NT) k_sample_energy(const double *__restrict__ dtab,
                                                      const cgrb_unit_t *__restrict__ units,
                                                      const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                      int interp_extrap, cgrb_rng_t rng, int64_t max_trials,
                                                      const double *__restrict__ times_src,
                                                      double *__restrict__ energy_src,
                                                      int32_t *__restrict__ trials_out, cgrb_counts_t *counts)
{
    __shared__ EnergyTab tab;
    __shared__ cgrb_unit_t u;
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const int64_t n_src = counts[wk.unit].n_src;
    if (wk.start >= n_src)
        return;
    if (threadIdx.x == 0)
        u = units[wk.unit];
    __syncthreads();
    const double *dt = dtab + (size_t)u.det * CGRB_DT_SIZE;
    const double opz = 1.0 + u.z;
    const double alpha = u.alpha;
    for (int m = threadIdx.x; m < CGRB_NE_ENVELOPE; m += NT) {
        double x = dt[CGRB_DT_G500_E + m] * opz;
        double A = dt[CGRB_DT_G500_A + m];
        tab.gx[m] = x;
        tab.gl[m] = (A > 0.0) ? log(A) + alpha * log(x) : -INFINITY;
    }
    for (int m = threadIdx.x; m < CGRB_NBINS; m += NT) {
        tab.eax[m] = dt[CGRB_DT_EA_X + m];
        tab.eay[m] = dt[CGRB_DT_EA_Y + m];
    }
    __syncthreads();

    const PropConst pc = make_prop_const(dt[CGRB_DT_EDGES], dt[CGRB_DT_EDGES + CGRB_NBINS], alpha);
    const EaGuess eg = make_ea_guess(tab.eax);

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // chunk end: next work entry of the same unit starts where this chunk ends
    int64_t chunk_end = n_src;
    if (w + 1 < n_work && work[w + 1].unit == wk.unit)
        chunk_end = min(n_src, work[w + 1].start);

    long long my_trials = 0;
    bool exhausted = false, capped = false;
    for (int64_t i = wk.start + wid; i < chunk_end; i += NW) {
        const double t = times_src[u.src_off + i];
        const SpecT s = spec_at(u, t);
        // argmax of the folded spectrum on the 500-point grid (first maximum), via log f
        double best = -INFINITY;
        int bidx = CGRB_NE_ENVELOPE;
        for (int m = lane; m < CGRB_NE_ENVELOPE; m += 32) {
            double g = tab.gl[m] - tab.gx[m] * s.inv_ec;
            if (g > best) {
                best = g;
                bidx = m;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, d);
            int oi = __shfl_xor_sync(0xffffffffu, bidx, d);
            if (ob > best || (ob == best && oi < bidx)) {
                best = ob;
                bidx = oi;
            }
        }
        int idx = (bidx < CGRB_NE_ENVELOPE) ? bidx : 0;
        const PhotonConst ph = literal_photon(u, s, dt, tab.gx, &idx);

        int64_t base_pos = 0;
        if (rng.mode != 0)
            base_pos = rng.d_off[u.src_off + i];
        const int64_t u_len = (rng.mode != 0 && rng.d_len) ? rng.d_len[0] : ((int64_t)1 << 62);

        double x_acc = nan("");
        int64_t ntr = 0;
        for (int64_t round = 0;; round++) {
            const int64_t j = round * 32 + lane;
            double u1, u2;
            bool ok = true;
            if (rng.mode == 0) {
                Philox4 p = philox_at(make_key(rng.seed), u.stream_id, ST_SRC_ENERGY, (uint64_t)i, (uint32_t)j);
                u1 = u53(p.x, p.y);
                u2 = u53(p.z, p.w);
            } else {
                int64_t pos = base_pos + 2 * j;
                if (pos + 1 < u_len) {
                    u1 = rng.d_u[pos];
                    u2 = rng.d_u[pos + 1];
                } else {
                    u1 = 0.5;
                    u2 = 0.0;
                    ok = false;
                }
            }
            double x;
            bool accept = literal_trial(pc, ph, tab.eax, tab.eay, eg, interp_extrap, u1, u2, &x) || !ok;
            unsigned bal = __ballot_sync(0xffffffffu, accept);
            if (bal) {
                int win = __ffs(bal) - 1;
                x_acc = __shfl_sync(0xffffffffu, x, win);
                bool ok_w = __shfl_sync(0xffffffffu, (int)ok, win);
                if (!ok_w) {
                    exhausted = true;
                    x_acc = nan("");
                }
                ntr = round * 32 + win + 1;
                break;
            }
            if (max_trials > 0 && (round + 1) * 32 >= max_trials) {
                capped = true;
                ntr = (round + 1) * 32;
                break;
            }
        }
        if (lane == 0) {
            energy_src[u.src_off + i] = x_acc;
            if (trials_out)
                trials_out[u.src_off + i] = (int32_t)ntr;
            my_trials += ntr;
        }
    }
    if (lane == 0 && my_trials)
        atomicAdd((unsigned long long *)&counts[wk.unit].n_trials, (unsigned long long)my_trials);
    if (lane == 0 && (exhausted || capped))
        atomicOr(&counts[wk.unit].status,
                 (exhausted ? CGRB_ST_UNIFORMS_EXHAUSTED : 0u) | (capped ? CGRB_ST_TRIAL_CAP : 0u));
}

// ==========================================================================================
// photon energies, envelope sampler (Philox only): one thread per photon; see cgrb_energy.cuh.
// Exact rejection sampling of the literal sampler's output density
//     h(x) ~ x^alpha min(A(x) e^{-x w}, 5 A(E_idx) e^{-E_idx w}),   w = (1+z)/ec(t)
// with proposal l = log(x w) from the per-unit piecewise envelope of phi(l) = (alpha+1) l - e^l,
// truncated to the photon's [log(emin w), log(emax w)] by inverse CDF, and A(x) <= Amax.
// Degenerate photons (pulse amplitude 0: the spurious tstart event) take the literal first trial.
// ==========================================================================================
// literal sampler for one photon by one thread (same Philox counters as the warp version); only
// degenerate photons take this path, so it is kept out of line to spare registers.
__device__ __noinline__ double literal_photon_serial(const cgrb_unit_t &u, const double *dt, const double *eax,
                                                     const double *eay, EaGuess eg, int interp_extrap,
                                                     uint64_t seed, int64_t i, double t, int64_t max_trials,
                                                     int64_t *ntr_out, bool *capped)
{
    const double opz = 1.0 + u.z, alpha = u.alpha;
    const SpecT sp = spec_at(u, t);
    const PropConst pc = make_prop_const(dt[CGRB_DT_EDGES], dt[CGRB_DT_EDGES + CGRB_NBINS], alpha);
    double best = -INFINITY;
    int idx = 0;
    // The spurious tstart photon of every unit has pulse amplitude 0: the whole 500-point row is zero, argmax
    // is 0 and the search below (1000 logarithms by ONE lane while its CTA waits at the next barrier -- it was
    // two thirds of this kernel's time on a population) cannot change that.
    if (sp.norm > 0.0)
        for (int m = 0; m < CGRB_NE_ENVELOPE; m++) {
            double A = dt[CGRB_DT_G500_A + m], xg = dt[CGRB_DT_G500_E + m] * opz;
            double g = (A > 0.0) ? log(A) + alpha * log(xg) - xg * sp.inv_ec : -INFINITY;
            if (g > best) {
                best = g;
                idx = m;
            }
        }
    double A_idx = dt[CGRB_DT_G500_A + idx], x_idx = dt[CGRB_DT_G500_E + idx] * opz;
    double f_idx = A_idx * cpl_value(sp, alpha, x_idx, log(x_idx));
    if (!(f_idx > 0.0)) {       // np.argmax of an all-zero row -> 0
        idx = 0;
        A_idx = dt[CGRB_DT_G500_A];
        x_idx = dt[CGRB_DT_G500_E] * opz;
        f_idx = A_idx * cpl_value(sp, alpha, x_idx, log(x_idx));
    }
    PhotonConst ph;
    ph.K1 = f_idx * 5.0 * exp(-alpha * log(dt[CGRB_DT_G500_E + idx]));
    ph.K2 = sp.norm * exp(alpha * (log(opz) - sp.log_ec));
    ph.w = opz * sp.inv_ec;
    for (uint32_t j = 0;; j++) {
        Philox4 p = philox_at(make_key(seed), u.stream_id, ST_SRC_ENERGY, (uint64_t)i, j);
        double x;
        *ntr_out = (int64_t)j + 1;
        if (literal_trial(pc, ph, eax, eay, eg, interp_extrap, u53(p.x, p.y), u53(p.z, p.w), &x))
            return x;
        if (max_trials > 0 && *ntr_out >= max_trials) {
            *capped = true;
            return nan("");
        }
    }
}

// rows of the proposal table: row r of the batch belongs to the unit with etab_off <= r < etab_off + n_w;
// row k of a unit holds the (shifted, unnormalised) cumulative masses of the 256 segments at w = wk[k]:
//     mass_j = exp(logM[j] - xlo[j] wk[k] - shift),  shift = max_j of the exponent.
#define ENV_ROW (CGRB_ENV_SEGS + 1)
__global__ void __launch_bounds__(NT) k_energy_rows(const cgrb_unit_t *__restrict__ units, int n_units,
                                                    const EnergyUnitTab *__restrict__ etab,
                                                    double *__restrict__ rows, uint8_t *__restrict__ row_guides,
                                                    int64_t n_rows)
{
    __shared__ double s_w[NW];
    __shared__ double s_red[NW];
    __shared__ double s_row[ENV_ROW];
    const int64_t r = blockIdx.x;
    if (r >= n_rows)
        return;
    int lo = 0, hi = n_units;               // last unit with etab_off <= r
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (units[mid].etab_off <= r)
            lo = mid;
        else
            hi = mid;
    }
    const EnergyUnitTab *T = etab + lo;
    const int k = (int)(r - units[lo].etab_off);
    const int j = threadIdx.x;               // NT == CGRB_ENV_SEGS
    double v = -INFINITY;
    if (k < T->n_w) {
        const double lm = T->logM[j];
        if (lm > -INFINITY)
            v = lm - T->xlo[j] * T->wk[k];
    }
    // block max
    double m = v;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((j & 31) == 0)
        s_red[j >> 5] = m;
    __syncthreads();
    m = s_red[0];
#pragma unroll
    for (int q = 1; q < NW; q++)
        m = fmax(m, s_red[q]);
    const double e = (v > -INFINITY && m > -INFINITY) ? exp(v - m) : 0.0;
    double total;
    const double incl = block_incl_scan(e, s_w, &total);
    double *row = rows + r * ENV_ROW;
    row[j] = incl - e;
    s_row[j] = incl - e;
    if (j == CGRB_ENV_SEGS - 1) {
        row[CGRB_ENV_SEGS] = incl;
        s_row[CGRB_ENV_SEGS] = incl;
    }
    __syncthreads();
    // guide[m] = segment of y = (m/256) total: the search for y' >= y starts there and walks forward
    const double y = ((double)j / (double)CGRB_ENV_SEGS) * s_row[CGRB_ENV_SEGS];
    int seg = 0;
#pragma unroll
    for (int step = CGRB_ENV_SEGS / 2; step >= 1; step >>= 1)
        if (s_row[seg + step] <= y)
            seg += step;
    row_guides[r * CGRB_ENV_SEGS + j] = (uint8_t)seg;
}

// resident CTAs per SM the kernel is compiled for: 4 (64 registers) since the trial loop's carried state was cut to
// what it needs (3.03 ms on C2 against 3.30 at 3 CTAs / 80 registers; before that cut 4 CTAs lost to their spills)
#ifndef ENV_MIN_CTAS
#define ENV_MIN_CTAS 4
#endif
// CGRB_ENV_HULL_GLOBAL=1 (default): the hull (12 KB, touched once per photon) stays in the unit's table in global
// memory and is read through L1 instead of being copied into every CTA's shared memory -- at 4 CTAs per SM that is
// 48 KB more L1 for the rows and the cumulative matrix the trials live on (k_sample_energy_env 3.04 -> 2.90 ms on C2,
// same output bits); 0 restores the shared-memory copy
#ifndef CGRB_ENV_HULL_GLOBAL
#define CGRB_ENV_HULL_GLOBAL 1
#endif
struct EnvSmem {
    double xlo[CGRB_ENV_SEGS];
    double invMx[CGRB_ENV_SEGS];
    double logAmax[CGRB_ENV_SEGS];
    int kn[CGRB_ENV_SEGS];
    double wk[CGRB_ENV_MAXW + 1];          // wk[n_w] = +inf
#if !CGRB_ENV_HULL_GLOBAL
    double hull_w[CGRB_NE_ENVELOPE + 1];   // hull_w[n_hull] = +inf
    double hull_E[CGRB_NE_ENVELOPE];
    double hull_lA5[CGRB_NE_ENVELOPE];
#endif
    double eax[CGRB_NBINS];
    double eay[CGRB_NBINS];
    double esl[CGRB_NBINS];                // slope of the effective-area curve on [eax[m], eax[m+1]]
    double edges[CGRB_NBINS + 1];
};

// One thread per photon, lanes refill from their warp's slice as they accept.  With pha_src != NULL the
// kernel also folds the accepted energy through the response (Response.digitize) -- the photon's bin
// follows from the effective-area bracket already in hand and the channel search starts from the
// detector's guide table staged in shared memory -- so energies never have to travel through HBM
// (energy_src may be NULL).
template <bool FOLD, bool DIAG, int MINB>   // FOLD: also digitize (pha_src); DIAG: write energies / trial counts
__global__ void __launch_bounds__(NT, MINB) k_sample_energy_env(const double *__restrict__ dtab,
                                                          const cgrb_unit_t *__restrict__ units,
                                                          const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                          int interp_extrap, cgrb_rng_t rng, int64_t max_trials,
                                                          const EnergyUnitTab *__restrict__ etab,
                                                          const double *__restrict__ rows,
                                                          const uint8_t *__restrict__ row_guides,
                                                          const double *__restrict__ times_src,
                                                          double *__restrict__ energy_src,
                                                          uint8_t *__restrict__ pha_src,
                                                          int32_t *__restrict__ trials_out, cgrb_counts_t *counts)
{
    __shared__ EnvSmem sm;
    __shared__ __align__(128) uint8_t s_guide[GUIDE_BYTES];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ cgrb_unit_t u;
    __shared__ unsigned s_trials[NW];
    __shared__ double ck[16];
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const int64_t n_src = counts[wk.unit].n_src;
    if (wk.start >= n_src)
        return;
    const EnergyUnitTab *T = etab + wk.unit;
    if (threadIdx.x == 0)
        u = units[wk.unit];
    const int n_w = T->n_w;
    const int n_hull = T->n_hull;
    for (int j = threadIdx.x; j < CGRB_ENV_SEGS; j += NT) {
        sm.xlo[j] = T->xlo[j];
        sm.invMx[j] = T->invMx[j];
        sm.logAmax[j] = T->logAmax[j];
        sm.kn[j] = T->kn[j];
    }
    for (int j = threadIdx.x; j <= n_w; j += NT)
        sm.wk[j] = (j < n_w) ? T->wk[j] : INFINITY;
#if CGRB_ENV_HULL_GLOBAL
#define HULL_W(idx) (((idx) < n_hull) ? __ldg(T->hull_w + (idx)) : (double)INFINITY)
#define HULL_E(idx) __ldg(T->hull_E + (idx))
#define HULL_LA5(idx) __ldg(T->hull_lA5 + (idx))
#else
#define HULL_W(idx) sm.hull_w[idx]
#define HULL_E(idx) sm.hull_E[idx]
#define HULL_LA5(idx) sm.hull_lA5[idx]
    for (int j = threadIdx.x; j <= n_hull; j += NT) {
        sm.hull_w[j] = (j < n_hull) ? T->hull_w[j] : INFINITY;
        if (j < n_hull) {
            sm.hull_E[j] = T->hull_E[j];
            sm.hull_lA5[j] = T->hull_lA5[j];
        }
    }
#endif
    __syncthreads();
    const double *dt = dtab + (size_t)u.det * CGRB_DT_SIZE;
    if (FOLD && threadIdx.x == 0)
        guide_stage(s_guide, &s_bar, dt);
    for (int m = threadIdx.x; m < CGRB_NBINS; m += NT) {
        const double x0 = dt[CGRB_DT_EA_X + m], y0 = dt[CGRB_DT_EA_Y + m];
        sm.eax[m] = x0;
        sm.eay[m] = y0;
        sm.esl[m] = (m < CGRB_NBINS - 1) ? (dt[CGRB_DT_EA_Y + m + 1] - y0) / (dt[CGRB_DT_EA_X + m + 1] - x0) : 0.0;
    }
    for (int m = threadIdx.x; m <= CGRB_NBINS; m += NT)
        sm.edges[m] = dt[CGRB_DT_EDGES + m];
    // loop constants live in shared memory (broadcast loads) to keep the trial loop's registers free
    if (threadIdx.x == 0) {
        const double alpha = u.alpha;
        ck[1] = T->du;
        ck[2] = T->a;
        ck[3] = T->logw0;
        ck[4] = T->inv_dlogw;
        ck[6] = dt[CGRB_DT_EDGES + CGRB_NBINS];
        // w(t) = (1+z)/ec(t) = wa + wb t   (ep = ep_start / (1 + t/tau); constant CPL: wb = 0)
        const double wa = (1.0 + u.z) * ((alpha == -2.0) ? 0.0001 : (2.0 + alpha)) / u.ep_start;
        ck[7] = wa;
        ck[8] = (u.kind == 0) ? wa / u.ep_tau : 0.0;
        // pulse amplitude is a normal number for t in (tlo, thi): trise/t + t/tdecay < 700
        double tlo = INFINITY, thi = -INFINITY;
        if (u.kind != 0) {
            tlo = -INFINITY;
            thi = INFINITY;
        } else if (u.peak_flux > 0.0 && u.tdecay > 0.0 && u.trise >= 0.0) {
            const double disc = 350.0 * 350.0 - u.trise / u.tdecay;
            if (disc >= 0.0) {
                const double r = 350.0 + sqrt(disc);
                tlo = u.trise / r;
                thi = u.tdecay * r;
            }
        }
        ck[9] = tlo;
        ck[10] = thi;
        ck[11] = 1.0 / (1.0 + u.z);
    }
    __syncthreads();
    if (FOLD)
        guide_wait(&s_bar);
#define c_du ck[1]
#define c_a ck[2]
#define c_logw0 ck[3]
#define c_inv_dlogw ck[4]
#define c_emax ck[6]
#define c_wa ck[7]
#define c_wb ck[8]
#define c_tlo ck[9]
#define c_thi ck[10]
#define c_inv_opz ck[11]
    const double *unit_rows = rows + u.etab_off * ENV_ROW;
    const uint8_t *unit_guides = row_guides + u.etab_off * CGRB_ENV_SEGS;
    const double *cum = dt + CGRB_DT_CUM;

    int64_t chunk_end = n_src;
    if (w + 1 < n_work && work[w + 1].unit == wk.unit)
        chunk_end = min(n_src, work[w + 1].start);
    const uint32_t trial_cap = (max_trials > 0 && max_trials < 0xffffffffll) ? (uint32_t)max_trials : 0xffffffffu;

    // Lane refill: every warp owns a contiguous slice of the chunk; a lane that accepted immediately takes
    // the warp's next photon, so no lane idles on a neighbour's geometric tail.
    // The state a lane carries from trial to trial is kept small (the loop wants ~118 registers and runs in the 64
    // of four CTAs per SM: whatever does not fit is spilled and re-read every trial): photons are
    // addressed relative to the work item's first one (32-bit), everything derived from (k, h, wq) -- the row total,
    // the hull line's slope term -- is re-read or recomputed where it is used, trials are counted per warp (one
    // per active lane and iteration), the status bits are raised where they occur.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (chunk_end - wk.start > 0x7fffffffll) {          // a work item is a few 10^4 photons; 2^31 is a caller's error
        if (threadIdx.x == 0)
            atomicOr(&counts[wk.unit].status, CGRB_ST_INTERNAL);
        return;
    }
    const int len = (int)(chunk_end - wk.start);
    const int per = (len + NW - 1) / NW;
    const int wbeg = wid * per, wend = min(wbeg + per, len);    // the warp's slice, relative to wk.start
    int next = wbeg;                     // warp-uniform
    const int64_t item_off = u.src_off + wk.start;              // photon wk.start + di lives at item_off + di
    if (lane == 0)
        s_trials[wid] = 0;               // the warp's trial count (every active lane of an iteration is one trial)
    bool active = false;
    // per-lane photon state; k (table row) and h (hull position) persist from photon to photon: a lane's
    // photons come in time order, so both usually advance by zero or one step
    int di = 0;
    uint32_t j = 0;
    int k = -1, h = -1;
    double wq = 0.0;
#define PHOTON_I (wk.start + (int64_t)di)
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !active);
        if (need) {
            const int cand = next + __popc(need & ((1u << lane) - 1u));
            next += __popc(need);
            if (!active && cand < wend) {
                di = cand;
                j = 0;
                const double t = times_src[item_off + di];
                bool regular = (t > c_tlo) && (t < c_thi) && n_hull > 0;
                wq = c_wa + c_wb * t;                           // (1+z) / ec(t)
                {
                    int kk = k, steps = 0;
                    if (kk >= 0)
                        while (steps < 4 && sm.wk[kk + 1] <= wq) {
                            kk++;
                            steps++;
                        }
                    if (kk < 0 || steps == 4 || sm.wk[kk] > wq) {
                        kk = (int)floor((log(wq) - c_logw0) * c_inv_dlogw);
                        kk = max(0, min(kk, n_w - 1));
                        if (sm.wk[kk] > wq)
                            kk--;
                        else if (sm.wk[kk + 1] <= wq)
                            kk++;
                    }
                    if (kk < 0 || !(sm.wk[max(kk, 0)] <= wq))   // w below the unit's grid (or NaN)
                        regular = false;
                    k = min(max(kk, 0), n_w - 1);
                }
                if (regular) {
                    const double wh = wq * c_inv_opz;           // 1 / ec
                    int hh = h, steps = 0;
                    if (hh >= 0)
                        while (steps < 4 && HULL_W(hh + 1) <= wh) {
                            hh++;
                            steps++;
                        }
                    if (hh < 0 || steps == 4 || HULL_W(hh) > wh) {
                        int lo = 0, hi = n_hull;                // last entry with hull_w <= wh
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (HULL_W(mid) <= wh)
                                lo = mid;
                            else
                                hi = mid;
                        }
                        hh = lo;
                    }
                    h = hh;
                    const double total = __ldg(unit_rows + (size_t)k * ENV_ROW + CGRB_ENV_SEGS);
                    if (!(total > 0.0) || isinf(total))
                        regular = false;
                }
                if (regular)
                    active = true;
                else {
                    // degenerate photon: the literal loop, right here (rare)
                    int64_t ntr = 0;
                    bool cap1 = false;
                    EaGuess eg = make_ea_guess(sm.eax);
                    const double x = literal_photon_serial(u, dt, sm.eax, sm.eay, eg, interp_extrap, rng.seed, PHOTON_I, t,
                                                           max_trials, &ntr, &cap1);
                    if (cap1)
                        atomicOr(&counts[wk.unit].status, CGRB_ST_TRIAL_CAP);
                    if (DIAG && energy_src)
                        energy_src[item_off + di] = x;
                    if (FOLD)
                        pha_src[item_off + di] = (uint8_t)digitize_one(sm.edges, s_guide, cum, x, rng.seed, u.stream_id, PHOTON_I);
                    if (DIAG && trials_out)
                        trials_out[item_off + di] = -(int32_t)ntr;      // negative: the literal loop (its channel comes from ST_SRC_CHAN)
                    if (ntr)                                            // rare path: straight to the unit's counter
                        atomicAdd((unsigned long long *)&counts[wk.unit].n_trials, (unsigned long long)ntr);
                }
            }
        }
        const unsigned trying = __ballot_sync(0xffffffffu, active);
        if (!trying) {
            if (next >= wend)
                break;
            continue;
        }
        if (lane == 0) {
            unsigned v = s_trials[wid] + __popc(trying);
            if (v & 0x80000000u) {          // never in practice: 2^31 trials in one warp's slice
                atomicAdd((unsigned long long *)&counts[wk.unit].n_trials, (unsigned long long)v);
                v = 0;
            }
            s_trials[wid] = v;
        }
        if (active) {
            const double *row = unit_rows + (size_t)k * ENV_ROW;
            Philox4 p = philox_at(make_key(rng.seed), u.stream_id, ST_SRC_ENERGY, (uint64_t)PHOTON_I, j);
            const TrialDraw td = trial_draw(p);     // proposal, height and (if accepted) the channel from ONE call
            const double U = td.U, V = td.V;
            const double y = U * __ldg(row + CGRB_ENV_SEGS);    // the row's total: an L1 hit behind the Philox rounds
            // segment: row[lo] <= y < row[lo+1], walking forward from the row's guide
            int lo = unit_guides[k * CGRB_ENV_SEGS + min((int)(U * (double)CGRB_ENV_SEGS), CGRB_ENV_SEGS - 1)];
            double c0 = __ldg(row + lo), c1 = __ldg(row + lo + 1);
            while (lo < CGRB_ENV_SEGS - 1 && c1 <= y) {
                lo++;
                c0 = c1;
                c1 = __ldg(row + lo + 1);
            }
            const double den = c1 - c0;
            const double fr = fmin(fmax((den > 1e-30) ? (y - c0) * rcp_fast(den) : (y - c0) / den, 0.0), 1.0);
            const double z = fr * c_du;                         // x = xlo e^z, z in [0, du]
            const double xl = sm.xlo[lo];
            const double x = fmin(xl * exp_small(z), c_emax);
            const double s = x * wq;
            // effective area at x: bracket from the segment's first node, then the linear piece
            int ib = min(max(sm.kn[lo] - 1, 0), CGRB_NBINS - 2);
            ib += (ib < CGRB_NBINS - 2 && sm.eax[ib + 1] < x) ? 1 : 0;   // a segment is narrower than a bin: 0 or 1 step
            while (ib < CGRB_NBINS - 2 && sm.eax[ib + 1] < x)           // (any other table: keep walking)
                ib++;
            double A = sm.eay[ib] + sm.esl[ib] * (x - sm.eax[ib]);
            if (interp_extrap == 1) {
                if (x <= sm.eax[0])
                    A = sm.eay[0];
                else if (x >= sm.eax[CGRB_NBINS - 1])
                    A = sm.eay[CGRB_NBINS - 1];
            }
            A = fmax(A, 0.0);
            const double s_idx = __dmul_rn(HULL_E(h), wq), lA5 = HULL_LA5(h);
            if (sm.logAmax[lo] - xl * wq > lA5 - s_idx)
                A = fmin(A, exp(lA5 + s - s_idx));   // the literal envelope clips f here
            // target / envelope = [A x^a / M_lo] exp(-d), decided from the bounds
            // 1 - d + d^2/2 - d^3/6 <= exp(-d) <= 1 - d + d^2/2 whenever they agree.  The envelope of row k is
            // M_j exp(-xlo_j w_k) exp(-emin (w - w_k)) >= M_j exp(-xlo_j w): the common factor in w - w_k costs
            // nothing, so   d = (x - xlo) w + (xlo - emin) (w - w_k) >= 0.
            const double base = A * sm.invMx[lo] * exp_small(c_a * z);
            const double d = fmax((x - xl) * wq + (xl - sm.xlo[0]) * (wq - sm.wk[k]), 0.0);
            const double qhi = fma(d, fma(d, 0.5, -1.0), 1.0);
            const double qlo = qhi - d * d * d * (1.0 / 6.0);
            j++;
            if (base * qlo > 1.000001)              // the envelope must dominate the target (self-check; never taken)
                atomicOr(&counts[wk.unit].status, CGRB_ST_ENVELOPE);
            bool acc;
            if (V <= base * qlo)
                acc = true;
            else if (V > base * qhi)
                acc = false;
            else
                acc = V <= base * exp(-d);
            const bool cap = !acc && j >= trial_cap;
            if (acc || cap) {
                if (DIAG && energy_src)
                    energy_src[item_off + di] = acc ? x : nan("");
                if (FOLD) {
                    int pha;
                    if (acc) {
                        // photon bin = searchsorted(edges, x, 'left') - 1, walked from the bracket
                        int b = ib + ((sm.edges[ib + 1] < x) ? 1 : 0);   // the bin centre bracket straddles one edge
                        while (b < CGRB_NBINS - 1 && sm.edges[b + 1] < x)
                            b++;
                        while (b >= 0 && !(sm.edges[b] < x))
                            b--;
                        if (b < 0)
                            b += CGRB_NBINS;                    // python negative index (x == emin)
                        pha = nearest_cdf_guided(s_guide + b * CGRB_NCHAN, cum + (size_t)b * CGRB_NCHAN, u32d(p.w));
                    } else
                        pha = digitize_one(sm.edges, s_guide, cum, nan(""), rng.seed, u.stream_id, PHOTON_I);
                    pha_src[item_off + di] = (uint8_t)pha;
                }
                if (DIAG && trials_out)
                    trials_out[item_off + di] = (int32_t)j;
                if (cap)
                    atomicOr(&counts[wk.unit].status, CGRB_ST_TRIAL_CAP);
                active = false;
            }
        }
    }
#undef c_du
#undef c_a
#undef c_logw0
#undef c_inv_dlogw
#undef c_emax
#undef c_wa
#undef c_wb
#undef c_tlo
#undef c_thi
#undef c_inv_opz
#undef PHOTON_I
#undef HULL_W
#undef HULL_E
#undef HULL_LA5
    if (lane == 0 && s_trials[wid])
        atomicAdd((unsigned long long *)&counts[wk.unit].n_trials, (unsigned long long)s_trials[wid]);
}

