/*
 * cgrb_oracle.c -- CPU restatement of cosmogrb's per-photon hot path.
 *
 * TEST INFRASTRUCTURE: parity oracle + "port" CPU baseline.  See cgrb_oracle.h.
 * Plain C99, single-threaded; each function restates the literal behaviour of the cited
 * reference lines (including their quirks) with the uniform random stream made explicit.
 */
#include "cgrb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================
 * uniform source
 * ==================================================================================== */

void co_rng_init_array(co_rng_t *r, const double *u, int64_t n_u)
{
    memset(r, 0, sizeof(*r));
    r->mode = 0;
    r->u = u;
    r->n_u = n_u;
}

/* MT19937 (Matsumoto & Nishimura), init_genrand seeding: the generator behind numba's
 * np.random.* and numpy's legacy RandomState(int_seed). */
void co_rng_init_mt(co_rng_t *r, uint32_t seed)
{
    memset(r, 0, sizeof(*r));
    r->mode = 1;
    r->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->mti = 624;
}

static uint32_t mt_next32(co_rng_t *r)
{
    if (r->mti >= 624) {
        uint32_t *mt = r->mt;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->mti = 0;
    }
    uint32_t y = r->mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

double co_rng_next(co_rng_t *r)
{
    if (r->mode == 0) {
        if (r->pos >= r->n_u) {
            r->exhausted = 1;
            r->pos++;
            return 0.5;
        }
        return r->u[r->pos++];
    }
    /* random_sample(): 53-bit double from two 32-bit draws */
    uint32_t a = mt_next32(r) >> 5, b = mt_next32(r) >> 6;
    r->pos++;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

/* ======================================================================================
 * regularised upper incomplete gamma Q(a,x): scipy.special.gammaincc (cephes igam.c lineage).
 * Branch selection and series follow the published algorithm; the asymptotic a~x branch
 * (a > 20) is never reached on this path (a = 2 + alpha in (0, 2)).
 * ==================================================================================== */

#define CO_MACHEP 1.11022302462515654042E-16
#define CO_MAXLOG 7.09782712893383996843E2
#define CO_MAXITER 2000
static const double co_big = 4.503599627370496e15;
static const double co_biginv = 2.22044604925031308085e-16;

/* x^a e^-x / Gamma(a) */
static double igam_fac(double a, double x)
{
    double ax = a * log(x) - x - lgamma(a);
    if (ax < -CO_MAXLOG)
        return 0.0;
    return exp(ax);
}

static double igam_series(double a, double x)
{
    double ax = igam_fac(a, x);
    if (ax == 0.0)
        return 0.0;
    double r = a, c = 1.0, ans = 1.0;
    for (int i = 0; i < CO_MAXITER; i++) {
        r += 1.0;
        c *= x / r;
        ans += c;
        if (c <= CO_MACHEP * ans)
            break;
    }
    return ans * ax / a;
}

static double igamc_series(double a, double x)
{
    double fac = 1.0, sum = 0.0, term;
    for (int n = 1; n < CO_MAXITER; n++) {
        fac *= -x / n;
        term = fac / (a + n);
        sum += term;
        if (fabs(term) <= CO_MACHEP * fabs(sum))
            break;
    }
    double logx = log(x);
    term = -expm1(a * logx - lgamma(1.0 + a));
    return term - exp(a * logx - lgamma(a)) * sum;
}

static double igamc_continued_fraction(double a, double x)
{
    double ax = igam_fac(a, x);
    if (ax == 0.0)
        return 0.0;
    double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
    double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
    double ans = pkm1 / qkm1, t;
    for (int i = 0; i < CO_MAXITER; i++) {
        c += 1.0;
        y += 1.0;
        z += 2.0;
        double yc = y * c;
        double pk = pkm1 * z - pkm2 * yc;
        double qk = qkm1 * z - qkm2 * yc;
        if (qk != 0) {
            double r = pk / qk;
            t = fabs((ans - r) / r);
            ans = r;
        } else
            t = 1.0;
        pkm2 = pkm1;
        pkm1 = pk;
        qkm2 = qkm1;
        qkm1 = qk;
        if (fabs(pk) > co_big) {
            pkm2 *= co_biginv;
            pkm1 *= co_biginv;
            qkm2 *= co_biginv;
            qkm1 *= co_biginv;
        }
        if (t <= CO_MACHEP)
            break;
    }
    return ans * ax;
}

double co_gammaincc(double a, double x)
{
    if (x < 0 || a < 0)
        return NAN;
    if (a == 0)
        return (x > 0) ? 0.0 : NAN;
    if (x == 0)
        return 1.0;
    if (isinf(a))
        return isinf(x) ? NAN : 1.0;
    if (isinf(x))
        return 0.0;
    if (x > 1.1) {
        if (x < a)
            return 1.0 - igam_series(a, x);
        return igamc_continued_fraction(a, x);
    } else if (x <= 0.5) {
        if (-0.4 / log(x) < a)
            return 1.0 - igam_series(a, x);
        return igamc_series(a, x);
    } else {
        if (x * 1.1 < a)
            return 1.0 - igam_series(a, x);
        return igamc_series(a, x);
    }
}

/* ======================================================================================
 * spectral / temporal functions
 * ==================================================================================== */

/* sampler/temporal_functions.py:5-14 */
double co_norris(double x, double K, double t_start, double t_rise, double t_decay)
{
    if (x > t_start)
        return K * exp(2 * sqrt(t_rise / t_decay)) *
               exp(-t_rise / (x - t_start) - (x - t_start) / t_decay);
    return 0.0;
}

/* sampler/cpl_functions.py:11-44 */
double co_cpl(double x, double alpha, double xp, double F, double a, double b)
{
    double ec;
    if (alpha == -2)
        ec = xp / 0.0001;
    else
        ec = xp / (2 + alpha);
    double g = tgamma(2.0 + alpha);
    double i1 = co_gammaincc(2.0 + alpha, a / ec) * g;
    double i2 = co_gammaincc(2.0 + alpha, b / ec) * g;
    double intflux = -ec * ec * (i2 - i1);
    double erg2keV = 6.24151e8;
    double norm = F * erg2keV / intflux;
    double log_xc = log(ec);
    double log_v = alpha * (log(x) - log_xc) - (x / ec);
    double flux = exp(log_v);
    return norm * flux;
}

/* np.searchsorted(xs, v, side='left') */
static int searchsorted_left(const double *xs, int n, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (xs[mid] < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* np.searchsorted(xs, v, side='right') */
static int searchsorted_right(const double *xs, int n, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (v < xs[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}

/* interpolation.interp(x, y, v) as called at cpl_functions.py:119 / cpl_constant_functions.py:61.
 * extrap 0: bracket index clamped, weight not clamped (linear extrapolation);
 * extrap 1: np.interp semantics (value clamped outside the table). */
double co_interp(const double *xs, const double *ys, int n, double v, int extrap)
{
    if (extrap == 0) {
        int i = searchsorted_left(xs, n, v) - 1;
        if (i < 0)
            i = 0;
        if (i > n - 2)
            i = n - 2;
        double lam = (v - xs[i]) / (xs[i + 1] - xs[i]);
        return (1.0 - lam) * ys[i] + lam * ys[i + 1];
    }
    if (v <= xs[0])
        return ys[0];
    if (v >= xs[n - 1])
        return ys[n - 1];
    int j = searchsorted_right(xs, n, v) - 1; /* xs[j] <= v < xs[j+1] */
    double slope = (ys[j + 1] - ys[j]) / (xs[j + 1] - xs[j]);
    return slope * (v - xs[j]) + ys[j];
}

/* cpl_functions.py:47-100 (kind 0) / cpl_constant_functions.py:11-44 (kind 1) */
double co_cpl_evolution(const co_source_t *s, double energy, double time)
{
    double a = 10.0 * (1 + s->z);
    double b = 1.0e4 * (1 + s->z);
    if (s->kind == 0) {
        double K = co_norris(time, s->peak_flux, 0.0, s->trise, s->tdecay);
        double ep = s->ep_start / (1 + time / s->ep_tau);
        return co_cpl(energy * (1 + s->z), s->alpha, ep, K, a, b);
    }
    return co_cpl(energy * (1 + s->z), s->alpha, s->ep_start, s->peak_flux, a, b);
}

/* cpl_functions.py:103-131 / cpl_constant_functions.py:47-63 */
double co_folded_cpl(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
                     double energy, double time)
{
    return co_interp(ea_x, ea_y, n_ea, energy, s->interp_extrap) * co_cpl_evolution(s, energy, time);
}

/* np.power(10., np.linspace(np.log10(lo), np.log10(hi), n)) */
static void log_grid(double lo, double hi, int n, double *out)
{
    double start = log10(lo), stop = log10(hi);
    double step = (stop - start) / (n - 1);
    for (int i = 0; i < n; i++)
        out[i] = pow(10.0, start + i * step);
    out[n - 1] = pow(10.0, stop);
}

/* cpl_functions.py:289-325 / cpl_constant_functions.py:180-196 */
double co_energy_integrated_evolution(const co_source_t *s, const double *ea_x, const double *ea_y,
                                      int n_ea, double time)
{
    enum { NE = 75 };
    double eg[NE], y[NE];
    log_grid(s->emin, s->emax, NE, eg);
    for (int m = 0; m < NE; m++)
        y[m] = co_folded_cpl(s, ea_x, ea_y, n_ea, eg[m], time);
    double acc = 0.0; /* np.trapz */
    for (int m = 1; m < NE; m++)
        acc += (eg[m] - eg[m - 1]) * (y[m] + y[m - 1]) / 2.0;
    return acc;
}

/* sampler/source.py:88-110 */
double co_fmax(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
               double tstart, double tstop)
{
    enum { NT = 50 };
    double step = (tstop - tstart) / (NT - 1);
    double best = -INFINITY;
    for (int i = 0; i < NT; i++) {
        double t = (i == NT - 1) ? tstop : tstart + i * step;
        double f = co_energy_integrated_evolution(s, ea_x, ea_y, n_ea, t);
        if (f > best || isnan(f)) /* np.max propagates NaN */
            best = f;
        if (isnan(best))
            break;
    }
    return best;
}

/* ======================================================================================
 * samplers
 * ==================================================================================== */

/* cpl_functions.py:134-188 / cpl_constant_functions.py:65-105 */
int64_t co_sample_events(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
                         double tstart, double tstop, double fmax, co_rng_t *rng,
                         double *out_times, int64_t cap,
                         double *out_cand_times, uint8_t *out_accept, int64_t cand_cap,
                         int64_t *n_candidates)
{
    double time = tstart;
    int64_t n = 0, nc = 0;
    if (cap < 1)
        return -1;
    out_times[n++] = time; /* spurious first event (:153-154) */
    for (;;) {
        time = time - (1.0 / fmax) * log(co_rng_next(rng));
        if (time > tstop)
            break;
        double test = co_rng_next(rng);
        double p_test = co_energy_integrated_evolution(s, ea_x, ea_y, n_ea, time) / fmax;
        int acc = (test <= p_test);
        if (out_cand_times) {
            if (nc >= cand_cap)
                return -1;
            out_cand_times[nc] = time;
            out_accept[nc] = (uint8_t)acc;
        }
        nc++;
        if (acc) {
            if (n >= cap)
                return -1;
            out_times[n++] = time;
        }
        if (rng->mode == 0 && rng->exhausted)
            return -1;
    }
    if (n_candidates)
        *n_candidates = nc;
    return n;
}

/* cpl_functions.py:191-286 / cpl_constant_functions.py:108-177 */
int64_t co_sample_energy(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
                         const double *times, int64_t n, co_rng_t *rng, const int64_t *offsets,
                         int64_t max_trials, double *out_energy, int32_t *out_trials)
{
    enum { NG = 500 };
    double egrid[NG];
    log_grid(s->emin, s->emax, NG, egrid);
    double alpha = s->alpha;
    int64_t status = 0;
    for (int64_t i = 0; i < n; i++) {
        double t = times[i];
        /* tmps[i, :] and its argmax (first maximum) (:212-237) */
        int idx = 0;
        double best = co_folded_cpl(s, ea_x, ea_y, n_ea, egrid[0], t);
        for (int m = 1; m < NG; m++) {
            double v = co_folded_cpl(s, ea_x, ea_y, n_ea, egrid[m], t);
            if (v > best) {
                best = v;
                idx = m;
            }
        }
        double C = best * 5;
        if (alpha == -1.0)
            alpha = -1 + 1e-20; /* literal no-op of :247-248 */
        if (offsets)
            rng->pos = offsets[i];
        int64_t trials = 0;
        for (;;) {
            double u = co_rng_next(rng);
            double x = pow((pow(s->emax, alpha + 1) - pow(s->emin, alpha + 1)) * u +
                               pow(s->emin, alpha + 1),
                           1.0 / (alpha + 1.0));
            double y = co_rng_next(rng) * C * pow(x / egrid[idx], alpha);
            trials++;
            if (y <= co_folded_cpl(s, ea_x, ea_y, n_ea, x, t)) {
                out_energy[i] = x;
                break;
            }
            if (max_trials > 0 && trials >= max_trials) {
                out_energy[i] = NAN;
                status = -1;
                break;
            }
        }
        if (out_trials)
            out_trials[i] = (int32_t)trials;
    }
    return status;
}

/* response/response.py:10-25.  searchsorted(...)-1 == -1 wraps to the last matrix row
 * (Python negative index); an energy above the last edge would index one past the matrix in
 * the reference (undefined) and is clamped to the last row here. */
void co_digitize(const double *energies, int64_t n, const double *edges, int n_edges,
                 const double *cum, int n_chan, co_rng_t *rng, double *out_pha)
{
    int n_bins = n_edges - 1;
    for (int64_t i = 0; i < n; i++) {
        int idx = searchsorted_left(edges, n_edges, energies[i]) - 1;
        if (idx < 0)
            idx += n_bins;
        if (idx > n_bins - 1)
            idx = n_bins - 1;
        double r = co_rng_next(rng);
        const double *row = cum + (int64_t)idx * n_chan;
        int best = 0;
        double bd = fabs(row[0] - r);
        for (int j = 1; j < n_chan; j++) {
            double d = fabs(row[j] - r);
            if (d < bd) {
                bd = d;
                best = j;
            }
        }
        out_pha[i] = (double)best;
    }
}

/* sampler/background.py:13-44 */
int64_t co_background_times(double tstart, double tstop, double rate, co_rng_t *rng,
                            double *out_times, int64_t cap)
{
    double fmax = rate;
    double time = tstart;
    int64_t n = 0;
    if (cap < 1)
        return -1;
    out_times[n++] = time;
    for (;;) {
        time = time - (1.0 / fmax) * log(co_rng_next(rng));
        if (time > tstop)
            break;
        double test = co_rng_next(rng);
        double p_test = rate / fmax;
        if (test <= p_test) {
            if (n >= cap)
                return -1;
            out_times[n++] = time;
        }
        if (rng->mode == 0 && rng->exhausted)
            return -1;
    }
    return n;
}

/* sampler/background.py:80-93: np.random.choice(channels, size, p=weights) =
 * cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, random_sample(size), side='right') */
void co_background_channels(const double *weights, int n_chan, int64_t n, co_rng_t *rng,
                            int64_t *out_chan)
{
    double *cdf = (double *)malloc(sizeof(double) * (size_t)n_chan);
    double acc = 0.0;
    for (int j = 0; j < n_chan; j++) {
        acc += weights[j];
        cdf[j] = acc;
    }
    for (int j = 0; j < n_chan; j++)
        cdf[j] /= acc;
    for (int64_t i = 0; i < n; i++)
        out_chan[i] = searchsorted_right(cdf, n_chan, co_rng_next(rng));
    free(cdf);
}

/* lightcurve/lightcurve.py:131-151.  Both inputs are ascending, so argsort of the appended
 * array is a two-way merge; ties resolve background-first (stable order of the append). */
void co_combine(const double *t_bkg, const double *pha_bkg, int64_t n_bkg,
                const double *t_src, const double *pha_src, int64_t n_src,
                double *out_t, double *out_pha)
{
    int64_t i = 0, j = 0, k = 0;
    while (i < n_bkg && j < n_src) {
        if (t_src[j] < t_bkg[i]) {
            out_t[k] = t_src[j];
            out_pha[k++] = pha_src[j++];
        } else {
            out_t[k] = t_bkg[i];
            out_pha[k++] = pha_bkg[i++];
        }
    }
    while (i < n_bkg) {
        out_t[k] = t_bkg[i];
        out_pha[k++] = pha_bkg[i++];
    }
    while (j < n_src) {
        out_t[k] = t_src[j];
        out_pha[k++] = pha_src[j++];
    }
}

/* instruments/gbm/gbm_lightcurve.py:80-115 (literal: t_end only advances on kept overflow events) */
int64_t co_gbm_dead_time(const double *time, const double *pha, int64_t n, uint8_t *selection)
{
    const double dead_time = 2.6e-6;
    const double overflow_dead_time = 10.6e-6;
    if (n <= 0)
        return 0;
    int64_t kept = 1;
    memset(selection, 0, (size_t)n);
    selection[0] = 1;
    double t_end;
    if (pha[0] == 127)
        t_end = time[0] + overflow_dead_time;
    else
        t_end = time[0] + dead_time;
    for (int64_t i = 1; i < n; i++) {
        if (time[i] > t_end) {
            selection[i] = 1;
            kept++;
            if (pha[i] == 127)
                t_end = time[i] + overflow_dead_time;
        }
    }
    return kept;
}

/* Physical (non-paralyzable) GBM dead time -- SURVEY.md §8 f4.  NOT computed by the reference (its loop,
 * restated above, only advances t_end on kept overflow events); this is the serial definition the
 * engine's CGRB_DEADTIME_PHYSICAL mode is checked against: every kept event opens a window of `tau`
 * seconds (`tau_overflow` for the overflow channel 127) and an event is kept iff it falls after the
 * window of the last kept event. */
int64_t co_gbm_dead_time_physical(const double *time, const double *pha, int64_t n, double tau,
                                  double tau_overflow, uint8_t *selection)
{
    if (n <= 0)
        return 0;
    int64_t kept = 1;
    memset(selection, 0, (size_t)n);
    selection[0] = 1;
    double t_end = time[0] + (pha[0] == 127 ? tau_overflow : tau);
    for (int64_t i = 1; i < n; i++) {
        if (time[i] > t_end) {
            selection[i] = 1;
            kept++;
            t_end = time[i] + (pha[i] == 127 ? tau_overflow : tau);
        }
    }
    return kept;
}

/* lightcurve/lightcurve.py:167-200 for one detector (GBMLightCurve), MT19937 streams in the
 * reference's consumption order: numba stream = source times, energies, channels, then
 * background times; numpy global stream = background channels. */
int64_t co_process_detector(const co_source_t *s, const double *edges, int n_edges,
                            const double *ea_x, const double *ea_y, const double *cum, int n_chan,
                            const double *bkg_weights, double src_tstart, double src_tstop,
                            double bkg_tstart, double bkg_tstop, double bkg_rate, double fmax_in,
                            uint32_t seed_numba, uint32_t seed_numpy, int64_t max_src_events,
                            int64_t *counters)
{
    int n_ea = n_edges - 1;
    co_rng_t nb, np;
    co_rng_init_mt(&nb, seed_numba);
    co_rng_init_mt(&np, seed_numpy);
    for (int k = 0; k < 6; k++)
        counters[k] = 0;

    /* fmax_in > 0: a window of a longer source interval, thinned with the full interval's fmax
     * (bounded-sample CPU baseline); otherwise Source.__init__'s own 50-point maximum */
    double fmax = (fmax_in > 0) ? fmax_in : co_fmax(s, ea_x, ea_y, n_ea, src_tstart, src_tstop);

    int64_t cap_src = (int64_t)(fmax * (src_tstop - src_tstart) * 1.2) + 1024;
    double *t_src = (double *)malloc(sizeof(double) * (size_t)cap_src);
    int64_t n_cand = 0;
    int64_t n_src = co_sample_events(s, ea_x, ea_y, n_ea, src_tstart, src_tstop, fmax, &nb, t_src,
                                     cap_src, NULL, NULL, 0, &n_cand);
    if (n_src < 0) {
        counters[5] = -1;
        free(t_src);
        return -1;
    }
    /* bounded sample: the caller may cap the number of source photons carried through the
     * energy/channel stages (the reference's sample_energy needs N x 500 doubles) */
    int64_t n_src_used = n_src;
    if (max_src_events > 0 && n_src_used > max_src_events)
        n_src_used = max_src_events;

    double *e_src = (double *)malloc(sizeof(double) * (size_t)(n_src_used + 1));
    double *pha_src = (double *)malloc(sizeof(double) * (size_t)(n_src_used + 1));
    int32_t *trials = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_src_used + 1));
    co_sample_energy(s, ea_x, ea_y, n_ea, t_src, n_src_used, &nb, NULL, 0, e_src, trials);
    int64_t n_trials = 0;
    for (int64_t i = 0; i < n_src_used; i++)
        n_trials += trials[i];
    co_digitize(e_src, n_src_used, edges, n_edges, cum, n_chan, &nb, pha_src);

    double expect = bkg_rate * (bkg_tstop - bkg_tstart);
    int64_t cap_bkg = (int64_t)(expect + 10.0 * sqrt(expect + 1.0)) + 1024;
    double *t_bkg = (double *)malloc(sizeof(double) * (size_t)cap_bkg);
    int64_t n_bkg = co_background_times(bkg_tstart, bkg_tstop, bkg_rate, &nb, t_bkg, cap_bkg);
    if (n_bkg < 0) {
        counters[5] = -2;
        n_bkg = 0;
    }
    int64_t *c_bkg = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_bkg + 1));
    co_background_channels(bkg_weights, n_chan, n_bkg, &np, c_bkg);
    double *pha_bkg = (double *)malloc(sizeof(double) * (size_t)(n_bkg + 1));
    for (int64_t i = 0; i < n_bkg; i++)
        pha_bkg[i] = (double)c_bkg[i];

    int64_t n_tot = n_bkg + n_src_used;
    double *t_all = (double *)malloc(sizeof(double) * (size_t)(n_tot + 1));
    double *p_all = (double *)malloc(sizeof(double) * (size_t)(n_tot + 1));
    co_combine(t_bkg, pha_bkg, n_bkg, t_src, pha_src, n_src_used, t_all, p_all);
    uint8_t *sel = (uint8_t *)malloc((size_t)(n_tot + 1));
    int64_t kept = co_gbm_dead_time(t_all, p_all, n_tot, sel);

    counters[0] = n_src_used;
    counters[1] = n_bkg;
    counters[2] = kept;
    counters[3] = n_cand;
    counters[4] = n_trials;

    free(t_src); free(e_src); free(pha_src); free(trials); free(t_bkg); free(c_bkg);
    free(pha_bkg); free(t_all); free(p_all); free(sel);
    return n_tot;
}
