/*
 * cgrb_oracle.h -- CPU restatement of cosmogrb's per-photon hot path (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle and the "port" CPU baseline.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product package
 * (cosmogrb_b200/) never links, imports or calls anything declared here.
 *
 * Every function cites the reference file:line (relative to the cosmogrb repository root)
 * whose literal behaviour it restates.  Parity pinning: checked against the real numba
 * kernels imported from the reference tree by oracle/make_golden.py (seeded MT19937 stream
 * replay); the resulting vectors are committed under tests/golden/.
 */
#ifndef CGRB_ORACLE_H
#define CGRB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- uniform source: injected array or MT19937 (numba / numpy legacy stream) ---------- */
typedef struct {
    int mode;            /* 0 = injected array, 1 = MT19937 */
    const double *u;     /* injected stream */
    int64_t n_u;         /* length of injected stream */
    int64_t pos;         /* number of uniforms consumed so far */
    int exhausted;       /* set when an injected stream ran dry (draw returns 0.5) */
    uint32_t mt[624];
    int mti;
} co_rng_t;

void co_rng_init_array(co_rng_t *r, const double *u, int64_t n_u);
void co_rng_init_mt(co_rng_t *r, uint32_t seed);
double co_rng_next(co_rng_t *r);

/* ---- source description -------------------------------------------------------------- */
typedef struct {
    int32_t kind;          /* 0 = time-evolving CPL (cpl_functions.py), 1 = constant CPL (cpl_constant_functions.py) */
    int32_t interp_extrap; /* 0 = linear extrapolation (EconForge interpolation.interp), 1 = clamp (np.interp) */
    double peak_flux;
    double ep_start;       /* kind 1: the constant ep */
    double ep_tau;
    double alpha;
    double trise;
    double tdecay;
    double z;
    double emin;           /* first photon-energy edge of the response (source_function.py:24-31) */
    double emax;           /* last photon-energy edge */
} co_source_t;

/* special functions (cephes-lineage igamc as shipped by scipy.special.gammaincc) */
double co_gammaincc(double a, double x);

/* sampler/temporal_functions.py:5-14 */
double co_norris(double x, double K, double t_start, double t_rise, double t_decay);
/* sampler/cpl_functions.py:11-44 */
double co_cpl(double x, double alpha, double xp, double F, double a, double b);
/* interpolation.interp as used at cpl_functions.py:119 */
double co_interp(const double *xs, const double *ys, int n, double v, int extrap);
/* cpl_functions.py:103-131 / cpl_constant_functions.py:47-63, one (energy, time) point */
double co_folded_cpl(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
                     double energy, double time);
/* cpl_functions.py:47-100 / cpl_constant_functions.py:11-44, one (energy,time) point, no response */
double co_cpl_evolution(const co_source_t *s, double energy, double time);
/* cpl_functions.py:289-325 / cpl_constant_functions.py:180-196 */
double co_energy_integrated_evolution(const co_source_t *s, const double *ea_x, const double *ea_y,
                                      int n_ea, double time);
/* sampler/source.py:88-110 */
double co_fmax(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
               double tstart, double tstop);

/* cpl_functions.py:134-188 / cpl_constant_functions.py:65-105.
 * out_times (capacity cap) receives tstart followed by the accepted times.  Optional
 * out_cand_times / out_accept (capacity cand_cap) receive every candidate that passed the
 * tstop test.  Returns the number of entries in out_times, or -1 if a capacity ran out. */
int64_t co_sample_events(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
                         double tstart, double tstop, double fmax, co_rng_t *rng,
                         double *out_times, int64_t cap,
                         double *out_cand_times, uint8_t *out_accept, int64_t cand_cap,
                         int64_t *n_candidates);

/* cpl_functions.py:191-286 / cpl_constant_functions.py:108-177.
 * If offsets != NULL photon i restarts the injected stream at position offsets[i]
 * (per-photon sub-streams); otherwise the stream is consumed serially as in the reference.
 * out_trials (optional) receives the number of rejection trials of each photon.
 * max_trials bounds the loop (0 = unbounded); returns -1 if it was hit. */
int64_t co_sample_energy(const co_source_t *s, const double *ea_x, const double *ea_y, int n_ea,
                         const double *times, int64_t n, co_rng_t *rng, const int64_t *offsets,
                         int64_t max_trials, double *out_energy, int32_t *out_trials);

/* response/response.py:10-25 */
void co_digitize(const double *energies, int64_t n, const double *edges, int n_edges,
                 const double *cum, int n_chan, co_rng_t *rng, double *out_pha);

/* sampler/background.py:13-44 */
int64_t co_background_times(double tstart, double tstop, double rate, co_rng_t *rng,
                            double *out_times, int64_t cap);
/* sampler/background.py:80-93 (numpy legacy Generator.choice with p) */
void co_background_channels(const double *weights, int n_chan, int64_t n, co_rng_t *rng,
                            int64_t *out_chan);

/* lightcurve/lightcurve.py:131-151: append(bkg, src) then argsort (stable restatement) */
void co_combine(const double *t_bkg, const double *pha_bkg, int64_t n_bkg,
                const double *t_src, const double *pha_src, int64_t n_src,
                double *out_t, double *out_pha);

/* instruments/gbm/gbm_lightcurve.py:80-115; returns the number of survivors */
int64_t co_gbm_dead_time(const double *time, const double *pha, int64_t n, uint8_t *selection);
/* non-paralyzable model (not the reference's; serial definition for CGRB_DEADTIME_PHYSICAL) */
int64_t co_gbm_dead_time_physical(const double *time, const double *pha, int64_t n, double tau,
                                  double tau_overflow, uint8_t *selection);

/* lightcurve/lightcurve.py:167-200 for one detector, MT19937-seeded, used as the CPU baseline.
 * fmax_in > 0 overrides Source._fmax (used to time a window of a longer burst).
 * Returns number of (source + background) events before dead time; counters[0..5] =
 * n_src, n_bkg, n_total_after_deadtime, n_candidates, n_trials, status. */
int64_t co_process_detector(const co_source_t *s, const double *edges, int n_edges,
                            const double *ea_x, const double *ea_y, const double *cum, int n_chan,
                            const double *bkg_weights, double src_tstart, double src_tstop,
                            double bkg_tstart, double bkg_tstop, double bkg_rate, double fmax_in,
                            uint32_t seed_numba, uint32_t seed_numpy, int64_t max_src_events,
                            int64_t *counters);

#ifdef __cplusplus
}
#endif
#endif
