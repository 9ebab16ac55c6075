/*
 * cosmogrb_b200.h -- C ABI of the B200-native photon-event engine (libcosmogrb_b200.so).
 *
 * Drop-in boundary for ONE path of grburgess/cosmogrb: the per-photon hot path
 *   NHPP thinning -> CPL energy draw -> DRM fold -> GBM dead time, batched over
 *   (GRB x detector) "units".
 * The reference has no FFI of its own (it is Python + numba); every entry point below names
 * the reference callable it replaces (file:line relative to the cosmogrb repo root).  The
 * ctypes binding a maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every pointer named d_* is DEVICE memory allocated by the caller (the Python host uses
 *     torch tensors purely as device buffers); h_* is host memory.  The library never frees
 *     or keeps caller memory; it allocates nothing on the device.
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream).  Calls return without synchronising unless stated.
 *   - return value: 0 = ok, <0 = error (CGRB_E_*); cgrb_last_error() gives a message.
 *     Data-dependent failures (capacity exceeded, injected stream exhausted, trial cap hit)
 *     are reported asynchronously through the per-unit status word in cgrb_counts_t.
 *   - units are independent; results do not depend on how units are split over launches,
 *     processes or GPUs (Philox counters are functions of (seed, unit stream id, stage, index)).
 */
#ifndef COSMOGRB_B200_H
#define COSMOGRB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGRB_ABI_VERSION 7

/* error codes */
#define CGRB_OK 0
#define CGRB_E_BADARG (-1)
#define CGRB_E_CUDA (-2)
#define CGRB_E_UNSUPPORTED (-3)

/* per-unit asynchronous status bits (cgrb_counts_t.status) */
#define CGRB_ST_SRC_CAPACITY 1u      /* candidate capacity reached before tstop             */
#define CGRB_ST_BKG_CAPACITY 2u      /* background capacity reached before tstop            */
#define CGRB_ST_UNIFORMS_EXHAUSTED 4u /* injected uniform stream too short                   */
#define CGRB_ST_TRIAL_CAP 8u         /* energy rejection hit max_trials                     */
#define CGRB_ST_OUT_CAPACITY 16u     /* an output buffer was too small                      */
#define CGRB_ST_ENVELOPE 32u         /* energy envelope failed to dominate the target (bug)  */
#define CGRB_ST_MAJORANT 128u        /* thinning majorant failed to dominate lambda(t) (bug)            */
#define CGRB_ST_SQUEEZE 256u         /* squeeze margin failed to bound |p - interpolation| at an exact evaluation (bug) */
#define CGRB_ST_INTERNAL 512u        /* a bounded wait inside a kernel gave up (look-back / bulk copy; never expected), or a work item beyond a kernel's index range */
#define CGRB_ST_DEADTIME_CHAIN 64u   /* physical dead time: unit handed to the serial fix-up (transient; cleared by it) */

/* fixed GBM shape (instruments/gbm: 140 photon bins x 128 PHA channels) */
#define CGRB_NBINS 140
#define CGRB_NCHAN 128
#define CGRB_NE_LAMBDA 75   /* sampler/cpl_functions.py:303 */
#define CGRB_NE_ENVELOPE 500 /* sampler/cpl_functions.py:208 */
#define CGRB_NT_FMAX 50     /* sampler/source.py:100 */

/* candidates per segment: the unit of work of the thinning kernels and the granularity of
 * the floating-point summation order of arrival times (fixed so results are reproducible) */
#define CGRB_SEG 8192

/*
 * Detector table block: one per detector, contiguous doubles in device memory, built on the
 * host (numpy, bit-identical to response/response.py:74-154) by
 * cosmogrb_b200.response.Response.  Offsets in doubles:
 */
#define CGRB_DT_EDGES 0        /* [141] photon energy edges (Response.energy_edges)                   */
#define CGRB_DT_EA_X 144       /* [140] effective_area_packed[0] (bin means)                          */
#define CGRB_DT_EA_Y 288       /* [140] effective_area_packed[1] (row sums, zeros -> 1e-99)           */
#define CGRB_DT_BKG_CDF 432    /* [128] background channel cdf (background.py:93 via numpy choice)    */
#define CGRB_DT_G75_E 560      /* [75]  10**linspace(log10 emin, log10 emax, 75)                      */
#define CGRB_DT_G75_A 640      /* [75]  interp(ea_x, ea_y, G75_E)                                     */
#define CGRB_DT_G500_E 720     /* [500] 10**linspace(log10 emin, log10 emax, 500)                     */
#define CGRB_DT_G500_A 1220    /* [500] interp(ea_x, ea_y, G500_E)                                    */
#define CGRB_DT_CUM 1720       /* [140*128] row-normalised cumulative matrix (16-byte aligned: 13760 B) */
#define CGRB_DT_GUIDE (1720 + 140 * 128) /* [140*128] BYTES: guide[b][m] = searchsorted(cum[b], m/128, 'left')
                                           (0..128): the channel search for r starts at guide[b][floor(128 r)]
                                           and walks forward -- same result as a full search, ~2 probes    */
#define CGRB_DT_BKG_GUIDE (1720 + 140 * 128 + 140 * 128 / 8) /* [1024] BYTES: guide of the background channel draw,
                                           guide[m] = min(searchsorted(bkg_cdf, m/1024, 'right'), 127)               */
#define CGRB_BKG_GUIDE_N 1024
#define CGRB_DT_SIZE (1720 + 140 * 128 + 140 * 128 / 8 + 128) /* 22008 doubles = 176064 B per detector  */

/* one (GRB, detector) work unit; 224 bytes, device array uploaded by the caller */
typedef struct {
    int32_t kind;          /* 0 = CPLSourceFunction (cpl_functions.py), 1 = ConstantCPL (cpl_constant_functions.py) */
    int32_t det;           /* detector table index into d_dtab                                     */
    uint32_t stream_id;    /* Philox stream id (grb_id * 16 + detector)                            */
    uint32_t flags;        /* bit0: source disabled, bit1: background disabled                     */
    double peak_flux;
    double ep_start;       /* kind 1: ep                                                           */
    double ep_tau;
    double alpha;
    double trise;
    double tdecay;
    double z;
    double src_tstart;     /* Source(tstart=0, tstop=duration) gbm_grb.py:61-67                    */
    double src_tstop;
    double bkg_tstart;     /* GBMGRB._background_start/_stop gbm_grb.py:20-21                      */
    double bkg_tstop;
    double bkg_rate;       /* Background._background_rate background.py:125                       */
    double fmax;           /* Source._fmax (source.py:88-110); written by cgrb_fmax                */
    double gamma_a;        /* Gamma(2+alpha)      (host, libm)                                     */
    double lgamma_a;       /* log Gamma(2+alpha)                                                   */
    double lgamma_1pa;     /* log Gamma(3+alpha)                                                   */
    int64_t src_cap;       /* candidate capacity of this unit (host: fmax*T + 8 sigma)             */
    int64_t bkg_cap;       /* background candidate capacity                                        */
    int64_t src_off;       /* offset of this unit in the per-source-event arrays                   */
    int64_t bkg_off;       /* offset of this unit in the per-background-event arrays               */
    int64_t tot_off;       /* offset of this unit in the merged / filtered arrays                  */
    int64_t tab_off;       /* offset (in nodes) of this unit in the thinning squeeze table           */
    int64_t tab_n;         /* nodes of the squeeze table on [src_tstart, src_tstop]; 0 = no squeeze  */
    double ea_scale;       /* multiplies the detector table's effective area (per-GRB viewing-angle
                              factor on a shared base DRM); 1.0 = use the table as is              */
    int64_t etab_off;      /* first row of this unit in the energy proposal table (cgrb_sample_energy_envelope) */
    int64_t etab_nw;       /* rows of this unit (grid of (1+z)/ec(t) over the burst), 1 .. 256           */
} cgrb_unit_t;

#define CGRB_UNIT_NO_SOURCE 1u
#define CGRB_UNIT_NO_BACKGROUND 2u

/* per-unit counters written by the kernels (device array, zeroed by cgrb_counts_reset) */
typedef struct {
    int64_t n_candidates;  /* thinning proposals with t <= tstop                                   */
    int64_t n_src;         /* len(times_source) incl. the spurious tstart event                    */
    int64_t n_bkg;         /* len(times_background) incl. the spurious tstart event                */
    int64_t n_total;       /* merged events entering dead time                                     */
    int64_t n_kept;        /* events surviving dead time                                           */
    int64_t n_trials;      /* energy rejection trials                                              */
    int64_t n_exact;       /* thinning candidates that needed the exact lambda(t) (all, without squeeze) */
    uint32_t status;       /* CGRB_ST_* bits                                                       */
    uint32_t pad;
} cgrb_counts_t;

/* work list entry: a chunk of a unit's index space (built on the host by cgrb_build_work) */
typedef struct {
    int32_t unit;
    int32_t seg;           /* chunk number inside the unit                                         */
    int64_t start;         /* first index of the chunk                                             */
} cgrb_work_t;

/* uniform source */
typedef struct {
    int32_t mode;          /* 0 = Philox4x32-10 (seed, unit stream id, stage, index), 1 = injected  */
    int32_t pad;
    uint64_t seed;
    const double *d_u;     /* injected: uniform buffer                                             */
    const int64_t *d_off;  /* injected: per unit (thin/bkg/digitize) or per photon (energy) start  */
    const int64_t *d_len;  /* injected: per unit number of uniforms available after d_off (may be NULL) */
} cgrb_rng_t;

int cgrb_abi_version(void);
const char *cgrb_last_error(void);
/* number of SMs etc. of the current device: out[0]=sm count, out[1]=cc major, out[2]=cc minor, out[3]=max smem optin */
int cgrb_device_info(int device, int *out4);

/* Optional per-launch timing for benchmarks (no reference counterpart).  While enabled every
 * kernel launch is bracketed by CUDA events on the launching stream.  cgrb_profile_enable(on)
 * clears the records; cgrb_profile_read synchronises on the recorded events and returns the
 * number of records copied (names: cap x 32 chars, may be NULL; ms: cap floats).
 * cgrb_launch_count() = kernels launched by this library since load. */
int cgrb_profile_enable(int on);
int cgrb_profile_read(char *names, float *ms, int cap);
long long cgrb_launch_count(void);

/* host helper: split [0, cap_u) of each unit into chunks of `chunk` indices.
 * h_work may be NULL to query the count.  cap_u = h_units[i].src_cap (which=0, thinning
 * candidates), bkg_cap (which=1), src_cap+bkg_cap+2 (which=2, merged events) or src_cap+1
 * (which=3, source events) or tab_n (which=4, squeeze-table nodes).  Every unit owns at least one entry.  Returns the number of entries. */
int64_t cgrb_build_work(const cgrb_unit_t *h_units, int n_units, int which, int64_t chunk,
                        cgrb_work_t *h_work, int64_t work_cap);

int cgrb_counts_reset(cgrb_counts_t *d_counts, int n_units, void *stream);

/* ---- spectral model --------------------------------------------------------------------- */

/* SourceFunction.energy_integrated_evolution(time) (cpl_source.py:81-96, constant_cpl.py:64-78;
 * numba cpl_functions.py:289-325) for unit `unit_index` at n times. */
int cgrb_energy_integrated_evolution(const double *d_dtab, const cgrb_unit_t *d_units, int unit_index,
                                     int interp_extrap, const double *d_times, int64_t n,
                                     double *d_out, void *stream);

/* SourceFunction.evolution(energy, time) -> [n_time, n_energy] row-major
 * (cpl_source.py:45-60; numba cpl_functions.py:47-100 / cpl_constant_functions.py:11-44). */
int cgrb_evolution(const cgrb_unit_t *d_units, int unit_index, const double *d_energy, int64_t n_energy,
                   const double *d_times, int64_t n_time, double *d_out, void *stream);

/* Source._get_energy_integrated_max (source.py:88-110): fills d_units[i].fmax for every unit. */
int cgrb_fmax(const double *d_dtab, cgrb_unit_t *d_units, int n_units, int interp_extrap, void *stream);

/* ---- samplers --------------------------------------------------------------------------- */

/* Squeeze table for the thinning acceptance test (no reference counterpart; it changes no decision).
 * For every unit with tab_n >= 4 nodes: p_k = lambda(t_k) / fmax on the uniform grid
 * t_k = src_tstart + k (src_tstop - src_tstart) / (tab_n - 1), and a margin m_k bounding the error of
 * linear interpolation on [t_k, t_k+1]: twice the error measured at the cell's midpoint plus the second
 * differences of the neighbouring nodes (x4 safety); infinite -- no squeeze, every candidate exact -- in the
 * first and last cell and in cells touching a node where p == 0 (the pulse's essential singularity at its
 * start).  A margin found too small at an exact evaluation sets CGRB_ST_SQUEEZE.  cgrb_sample_events
 * accepts a candidate outright if u2 <= p_lin - m, rejects it if u2 > p_lin + m, and evaluates
 * lambda(t) exactly (the reference's 75-point trapezoid) only in between, so the accept/reject
 * decisions are those of the exact test.  d_thin_tab: 2 doubles per node (p, m), sum(tab_n) nodes.
 * Work list: which=4, chunk=256.
 *
 * d_thin_cum (optional, NULL = not built): the piecewise-constant MAJORANT of the same table for
 * cgrb_sample_events under Philox -- cgrb_thin_cum_bytes(n_nodes_total) bytes: per node the cumulative
 * operational time Lambda_k = fmax * sum_{j<k} P_j h, P_j = min(1, max(p_j, p_j+1) + m_j) >= min(p, 1) on
 * cell j, then per node an int32 search guide.  n_nodes_total = sum(tab_n) (units' tab_off index both). */
int64_t cgrb_thin_cum_bytes(int64_t n_nodes_total);
int cgrb_thin_tables(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                     const cgrb_work_t *d_work, int64_t n_work, double *d_thin_tab, void *d_thin_cum,
                     int64_t n_nodes_total, void *stream);

/* SourceFunction.sample_events (cpl_source.py:98-114; numba cpl_functions.py:134-188,
 * cpl_constant_functions.py:65-105).  Work list: which=0, chunk <= CGRB_SEG (an even multiple of 512): a segment
 * ends where the unit's next entry starts, so a caller may cut small units into shorter segments (fewer serial
 * rounds per CTA); the chunking must be a function of the unit alone if results are to be independent of the batch.
 *   d_segsum/d_segstart [n_work] doubles, d_segcount [n_work] int32: workspace
 *   d_stage   [sum src_cap] doubles: workspace (accepted times per segment)
 *   d_times_src [sum (src_cap+1)]: output, unit u at src_off; entry 0 is tstart
 *   d_cand_times / d_cand_accept: optional (may be NULL) per-candidate outputs at src_off
 *   d_thin_tab: squeeze table from cgrb_thin_tables, or NULL (every candidate evaluated exactly)
 *   d_thin_cum: majorant table from cgrb_thin_tables, or NULL.  With it, and only under Philox (injected
 *       uniforms always replay the reference's loop), units with a table draw their candidates as a unit-rate
 *       Poisson process in operational time, map them back through Lambda and keep each with probability
 *       min(p, 1) / P_j: the same point process min(lambda, fmax) as the reference's single-fmax thinning
 *       from integral(P) fmax candidates instead of fmax T.  CGRB_ST_MAJORANT flags a violated bound.
 * counts: n_candidates, n_src, status. */
int cgrb_sample_events(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                       const cgrb_work_t *d_work, int64_t n_work, const int64_t *d_unit_work_begin,
                       int interp_extrap, const cgrb_rng_t *rng,
                       double *d_segsum, double *d_segstart, int32_t *d_segcount, double *d_stage,
                       double *d_times_src, double *d_cand_times, uint8_t *d_cand_accept,
                       const double *d_thin_tab, const void *d_thin_cum, int64_t n_nodes_total,
                       cgrb_counts_t *d_counts, void *stream);

/* SourceFunction.sample_energy (cpl_source.py:116-132; numba cpl_functions.py:191-286).
 * Work list: which=3 with any chunk size (photon index space, bounded by counts.n_src).
 * d_trials may be NULL.  Injected mode: rng->d_off is per photon (absolute positions in d_u),
 * rng->d_len[0] is the total length of d_u. */
int cgrb_sample_energy(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                       const cgrb_work_t *d_work, int64_t n_work, int interp_extrap,
                       const cgrb_rng_t *rng, int64_t max_trials, const double *d_times_src,
                       double *d_energy_src, int32_t *d_trials, cgrb_counts_t *d_counts, void *stream);

/* Same contract and the same output DISTRIBUTION as cgrb_sample_energy, Philox only: exact
 * rejection sampling of the literal sampler's output density
 *     h(x) ~ x^alpha min(A(x) e^(-x w), 5 A(E_idx) e^(-E_idx w)),   w = (1+z)/ec(t)
 * from a per-unit tabulated envelope (csrc/cgrb_energy.cuh): 256 equal segments in log x, each bounded
 * by max x^(alpha+1) A(x) over the segment (A is piecewise linear: exact) times exp(-x_lo w_k), on a
 * grid of etab_nw values w_k <= w covering the unit's spectral evolution; one cumulative row per w_k.
 * ~1.1-1.3 trials per photon instead of tens to hundreds.  Every unit owns etab_nw >= 1 rows starting
 * at row etab_off (ascending); n_rows = total rows of the batch.  d_workspace:
 * cgrb_energy_workspace_bytes(n_units, n_rows) bytes of device memory.  Injected uniforms replay the
 * reference's stream and therefore must use cgrb_sample_energy (returns CGRB_E_UNSUPPORTED).
 * Fused fold: with d_pha_src != NULL the kernel also performs Response.digitize on every accepted
 * energy (same Philox counters as cgrb_digitize, hence the same channels) and d_energy_src may then
 * be NULL, so photon energies never travel through HBM.
 * One work item (the photons from its start to the unit's next item) must hold fewer than 2^31 photons
 * (the engine uses 16384); a longer one sets CGRB_ST_INTERNAL on the unit and is not sampled.
 * n_trials counts one trial per active lane and loop iteration, the literal fall-back's trials added. */
int64_t cgrb_energy_workspace_bytes(int n_units, int64_t n_rows);
int cgrb_sample_energy_envelope(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                                const cgrb_work_t *d_work, int64_t n_work, int interp_extrap,
                                const cgrb_rng_t *rng, int64_t max_trials, void *d_workspace, int64_t n_rows,
                                const double *d_times_src, double *d_energy_src, uint8_t *d_pha_src,
                                int32_t *d_trials, cgrb_counts_t *d_counts, void *stream);
/* The same in two steps.  The per-unit envelope tables and rows depend on the unit records and the detector tables
 * only, not on the photons: cgrb_energy_envelope_tables builds them into d_workspace (on any stream, e.g. while the
 * source times of SourceFunction.sample_events are still being drawn on another), and
 * cgrb_sample_energy_envelope_prepared then samples with the tables of d_workspace taken as built (the caller orders
 * the two calls: same stream, or an event).  Together they are cgrb_sample_energy_envelope. */
int cgrb_energy_envelope_tables(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, int interp_extrap,
                                void *d_workspace, int64_t n_rows, void *stream);
int cgrb_sample_energy_envelope_prepared(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                                         const cgrb_work_t *d_work, int64_t n_work, int interp_extrap,
                                         const cgrb_rng_t *rng, int64_t max_trials, void *d_workspace, int64_t n_rows,
                                         const double *d_times_src, double *d_energy_src, uint8_t *d_pha_src,
                                         int32_t *d_trials, cgrb_counts_t *d_counts, void *stream);

/* Response.digitize (response.py:156-174; numba _digitize :10-25): nearest-CDF channel.
 * Work list as for cgrb_sample_energy.  d_trials (may be NULL; Philox only): the per-photon trial counts written by
 * cgrb_sample_energy_envelope -- a photon with a positive count takes its channel uniform from the Philox call of
 * the trial that accepted it, which replays the fold fused into that kernel; otherwise photon i draws from its own
 * counter. */
int cgrb_digitize(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                  const cgrb_work_t *d_work, int64_t n_work, const cgrb_rng_t *rng,
                  const double *d_energy_src, const int32_t *d_trials, uint8_t *d_pha_src,
                  const cgrb_counts_t *d_counts, void *stream);

/* Background.sample_times + sample_channel (background.py:136-174; numba :13-44; numpy choice :93), one pass:
 * tiles of CGRB_BKG_TILE candidates, launched level-major over the units (no work list).  n_tiles: an upper bound
 * on sum over the units with background of max(1, ceil(bkg_cap / CGRB_BKG_TILE)) (the grid size); d_workspace:
 * cgrb_sample_background_workspace_bytes(n_units, n_tiles) bytes.  rng_chan is used for the channel draw in
 * injected mode (one uniform per event, incl. event 0); in Philox mode both come from one counter.
 * counts: n_bkg, status. */
#define CGRB_BKG_TILE 2048
int64_t cgrb_sample_background_workspace_bytes(int n_units, int64_t n_tiles);
int cgrb_sample_background(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, int64_t n_tiles,
                           const cgrb_rng_t *rng_time, const cgrb_rng_t *rng_chan, void *d_workspace,
                           double *d_times_bkg, uint8_t *d_pha_bkg, cgrb_counts_t *d_counts,
                           void *stream);

/* BackgroundSpectrumTemplate.sample_channel(size=n) alone (background.py:80-93): n channel draws
 * from detector table `det`'s background cdf.  Injected mode reads rng->d_u[0..n). */
int cgrb_background_channels(const double *d_dtab, int det, uint32_t stream_id, const cgrb_rng_t *rng,
                             int64_t n, uint8_t *d_out, void *stream);

/* ---- combine + dead time ---------------------------------------------------------------- */

/* LightCurve._combine (lightcurve.py:131-151): two-way merge of the (ascending) background and
 * source lists, background first on ties.  Work list: which=2, chunk=CGRB_MERGE_TILE. */
#define CGRB_MERGE_TILE 2048
int cgrb_combine(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                 const double *d_times_bkg, const uint8_t *d_pha_bkg,
                 const double *d_times_src, const uint8_t *d_pha_src,
                 double *d_times_tot, uint8_t *d_pha_tot, cgrb_counts_t *d_counts, void *stream);

/* Dead-time model (SURVEY.md §8b "dead-time with mode in {reference, physical}").
 *   CGRB_DEADTIME_REFERENCE: the literal semantics of numba _gbm_dead_time (gbm_lightcurve.py:80-115):
 *       event 0 kept, window 2.6 us (10.6 us if it is in the overflow channel 127); later events are kept
 *       iff t > t_end, and ONLY a kept overflow event moves t_end (to t + 10.6 us).  tau / tau_overflow
 *       are ignored.  This is what every parity test compares with the reference.
 *   CGRB_DEADTIME_PHYSICAL: the non-paralyzable detector the reference's docstring describes (§8 f4):
 *       EVERY kept event opens a window (tau_overflow for channel 127, else tau); an event is kept iff
 *       t > t_end of the last kept event.  Not computed by the reference; checked against the serial
 *       definition in oracle/cgrb_oracle.c (co_gbm_dead_time_physical).  Resolved in parallel from
 *       synchronisation points (gaps longer than the longest window); a unit without one inside the
 *       look-back range (event rate x window > ~3) is redone serially by a fix-up kernel of the same call.
 * A NULL model pointer means CGRB_DEADTIME_REFERENCE. */
#define CGRB_DEADTIME_REFERENCE 0
#define CGRB_DEADTIME_PHYSICAL 1
typedef struct {
    int32_t mode;
    int32_t pad;
    double tau;            /* physical mode: window after a kept event, seconds (GBM: 2.6e-6)        */
    double tau_overflow;   /* physical mode: window after a kept overflow-channel event (GBM: 10e-6) */
} cgrb_deadtime_model_t;

/* GBMLightCurve._filter_deadtime (gbm_lightcurve.py:42-56; numba _gbm_dead_time :80-115).
 * Work list: which=2, chunk=CGRB_DT_TILE; d_tile_count [n_work] int32 is
 * workspace.  d_selection (optional) receives the per-event keep flag.  Output is compacted at
 * tot_off; counts.n_kept. */
#define CGRB_DT_TILE 2048
int cgrb_gbm_dead_time(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                       const int64_t *d_unit_work_begin,
                       const double *d_times_tot, const uint8_t *d_pha_tot, int32_t *d_tile_count,
                       double *d_times_out, uint8_t *d_pha_out, uint8_t *d_selection,
                       cgrb_counts_t *d_counts, const cgrb_deadtime_model_t *model, void *stream);

/* LightCurve._combine followed by GBMLightCurve._filter_deadtime in one pass (lightcurve.py:131-151,
 * 189-190; gbm_lightcurve.py:42-56): the merged pre-dead-time list is not one of the products of
 * LightCurve.process, so it is merged tile by tile in shared memory and only the survivors are written
 * (compacted at tot_off; counts.n_total, counts.n_kept).  Bit-identical to cgrb_combine +
 * cgrb_gbm_dead_time.  Work list: which=2, chunk=CGRB_MD_TILE; d_workspace:
 * cgrb_merge_dead_time_workspace_bytes(n_work) bytes (tile states of the decoupled look-back, the tile ticket
 * counter and the merge-path splits; cleared by the call). */
#define CGRB_MD_TILE 2048
int64_t cgrb_merge_dead_time_workspace_bytes(int64_t n_work);
int cgrb_merge_dead_time(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                         const int64_t *d_unit_work_begin, const double *d_times_bkg, const uint8_t *d_pha_bkg,
                         const double *d_times_src, const uint8_t *d_pha_src, void *d_workspace,
                         double *d_times_out, uint8_t *d_pha_out, cgrb_counts_t *d_counts,
                         const cgrb_deadtime_model_t *model, void *stream);

/* ---- GBM trigger ----------------------------------------------------------- */

/* GBMLightCurveAnalyzer (instruments/gbm/gbm_lightcurve_analyzer.py:13-271) for every unit of a batch, on
 * the dead-time-filtered total list left in HBM by cgrb_merge_dead_time / cgrb_gbm_dead_time: 16 ms binning
 * in the four trigger energy ranges (LightCurveStorage.binned_counts, light_curve_storage.py:154-294), dead
 * time per bin, sliding-window significance over the trigger time scales; the first (energy range, time
 * scale) in the reference's order that fires gives the detection time.  2.8e6 events per GRB reduce to
 * 32 bytes per detector without leaving the device.
 *   d_chan_bounds [n_det][8] int32: per detector table and energy range r, a channel c is selected iff
 *       bounds[2r] < c < bounds[2r+1]  (= ebounds.searchsorted(emin) / (emax); no bound: -1 / 1<<30)
 *   bins_per_unit: >= number of 16 ms bins any unit can have (host: ceil(span / base_timescale) + 2)
 *   work list: which=2, chunk=CGRB_TRIG_TILE.  d_workspace: cgrb_gbm_trigger_workspace_bytes(). */
#define CGRB_TRIG_TILE 4096
typedef struct {
    double threshold;            /* 4.5  */
    double base_timescale;       /* 0.016 */
    double background_duration;  /* 17    */
    double pre_window;           /* 4     */
} cgrb_trigger_par_t;
typedef struct {
    int32_t detected;      /* GBMLightCurveAnalyzer.is_detected                                         */
    int32_t combo;         /* 16 * energy range (a..d = 0..3) + index of the time scale (longest = 0), -1 */
    int32_t n_bins;        /* 16 ms bins of the light curve                                             */
    int32_t razor;         /* some window's significance lies within 1e-9 (relative) of the threshold   */
    double time;           /* detection_time                                                            */
    double time_scale;     /* detection_time_scale                                                      */
} cgrb_trigger_t;
int64_t cgrb_gbm_trigger_workspace_bytes(int n_units, int64_t bins_per_unit);
int cgrb_gbm_trigger(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                     const int32_t *d_chan_bounds, const cgrb_trigger_par_t *par, const double *d_times_out,
                     const uint8_t *d_pha_out, const cgrb_counts_t *d_counts, int64_t bins_per_unit,
                     void *d_workspace, cgrb_trigger_t *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* COSMOGRB_B200_H */
