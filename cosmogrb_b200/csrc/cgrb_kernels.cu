// cgrb_kernels.cu -- sm_100a kernels + C ABI of the photon-event engine.
//
// One translation unit: the kernels live in the headers included below (in dependency order), this file
// holds the error / profiling plumbing and the extern "C" entry points.  See include/cosmogrb_b200.h for the
// boundary and DESIGN.md for the data layout and the roofline of each kernel.  Reference lines are relative
// to the cosmogrb repository root.
//   cgrb_device.cuh          Philox, incomplete gamma, the CPL model, searches, warp scans
//   cgrb_energy.cuh          per-unit tables of the energy samplers (k_energy_tables), the literal trial
//   cgrb_common.cuh          block scans, thinning uniforms, the lambda(t) table
//   cgrb_spectral.cuh        k_lambda, k_evolution, k_fmax, k_thin_nodes, k_thin_margins
//   cgrb_thinning.cuh        k_gap_prefix, k_seg_scan, k_source_thin, k_source_gather, k_background
//   cgrb_fold.cuh            guided channel search, guide staging (cp.async.bulk), k_digitize
//   cgrb_energy_kernels.cuh  k_sample_energy, k_energy_rows, k_sample_energy_env
//   cgrb_deadtime.cuh        k_combine, k_deadtime_count/write, k_merge_plan, k_merge_deadtime
//   cgrb_trigger.cuh         k_trig_bin, k_trig_eval
#include <cuda_runtime.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "cgrb_device.cuh"
#include "cgrb_energy.cuh"

using namespace cgrb;

// ==========================================================================================
// error plumbing
// ==========================================================================================
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, const char *detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
static int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return CGRB_E_CUDA;
    }
    return CGRB_OK;
}

// ==========================================================================================
// optional per-launch timing (cgrb_profile_enable / cgrb_profile_read): CUDA events recorded on
// the launching stream around every kernel, so bench.py can report each kernel's live duration.
// ==========================================================================================
struct ProfRec {
    const char *name;
    cudaEvent_t a, b;
};
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static long long g_launches = 0;

static inline void prof_begin(const char *name, cudaStream_t s)
{
    g_launches++;
    if (!g_prof)
        return;
    ProfRec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    g_recs.push_back(r);
}
static inline void prof_end(cudaStream_t s)
{
    if (g_prof)
        cudaEventRecord(g_recs.back().b, s);
}
#define LAUNCH(name, stream, ...)   \
    do {                            \
        prof_begin(name, stream);   \
        __VA_ARGS__;                \
        prof_end(stream);           \
    } while (0)

#include "cgrb_common.cuh"
#include "cgrb_spectral.cuh"
#include "cgrb_thinning.cuh"
#include "cgrb_fold.cuh"
#include "cgrb_energy_kernels.cuh"
#include "cgrb_deadtime.cuh"
#include "cgrb_trigger.cuh"

__global__ void __launch_bounds__(NT) k_background_channels(const double *__restrict__ dtab, int det, uint32_t stream_id,
                                                            cgrb_rng_t rng, int64_t n, uint8_t *__restrict__ out)
{
    __shared__ double s_cdf[CGRB_NCHAN];
    const double *cdf_g = dtab + (size_t)det * CGRB_DT_SIZE + CGRB_DT_BKG_CDF;
    if (threadIdx.x < CGRB_NCHAN)
        s_cdf[threadIdx.x] = cdf_g[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
        double uc;
        if (rng.mode == 0) {
            Philox4 p = philox_at(make_key(rng.seed), stream_id, ST_BKG, (uint64_t)i, 1u);
            uc = u53(p.x, p.y);
        } else
            uc = rng.d_u[i];
        out[i] = (uint8_t)min(searchsorted_right(s_cdf, CGRB_NCHAN, uc), CGRB_NCHAN - 1);
    }
}

__global__ void k_counts_reset(cgrb_counts_t *counts, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        cgrb_counts_t z;
        memset(&z, 0, sizeof(z));
        counts[i] = z;
    }
}

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int cgrb_abi_version(void) { return CGRB_ABI_VERSION; }

int cgrb_profile_enable(int on)
{
    for (auto &r : g_recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_recs.clear();
    g_prof = on != 0;
    return CGRB_OK;
}

int cgrb_profile_read(char *names, float *ms, int cap)
{
    int n = 0;
    for (auto &r : g_recs) {
        if (n >= cap)
            break;
        if (cudaEventSynchronize(r.b) != cudaSuccess)
            return fail(CGRB_E_CUDA, "cgrb_profile_read: event sync failed%s");
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[n] = t;
        if (names) {
            strncpy(names + (size_t)n * 32, r.name, 31);
            names[(size_t)n * 32 + 31] = 0;
        }
        n++;
    }
    return n;
}

long long cgrb_launch_count(void) { return g_launches; }

const char *cgrb_last_error(void) { return g_err; }

int cgrb_device_info(int device, int *out4)
{
    if (!out4)
        return fail(CGRB_E_BADARG, "cgrb_device_info: null output%s");
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess)
        return fail(CGRB_E_CUDA, "cgrb_device_info: %s", cudaGetErrorString(e));
    out4[0] = p.multiProcessorCount;
    out4[1] = p.major;
    out4[2] = p.minor;
    out4[3] = (int)p.sharedMemPerBlockOptin;
    return CGRB_OK;
}

int64_t cgrb_build_work(const cgrb_unit_t *h_units, int n_units, int which, int64_t chunk,
                        cgrb_work_t *h_work, int64_t work_cap)
{
    if (!h_units || n_units < 0 || chunk <= 0 || which < 0 || which > 4)
        return fail(CGRB_E_BADARG, "cgrb_build_work: bad argument%s");
    int64_t n = 0;
    for (int i = 0; i < n_units; i++) {
        int64_t cap = which == 0   ? h_units[i].src_cap
                      : which == 1 ? h_units[i].bkg_cap
                      : which == 2 ? h_units[i].src_cap + h_units[i].bkg_cap + 2
                      : which == 3 ? h_units[i].src_cap + 1
                                   : h_units[i].tab_n;
        int64_t nseg = (cap + chunk - 1) / chunk;
        if (nseg < 1)
            nseg = 1;   // every unit owns at least one entry (it carries the per-unit epilogue)
        for (int64_t s = 0; s < nseg; s++) {
            if (h_work) {
                if (n >= work_cap)
                    return fail(CGRB_E_BADARG, "cgrb_build_work: work_cap too small%s");
                h_work[n].unit = i;
                h_work[n].seg = (int32_t)s;
                h_work[n].start = s * chunk;
            }
            n++;
        }
    }
    return n;
}

int cgrb_counts_reset(cgrb_counts_t *d_counts, int n_units, void *stream)
{
    if (!d_counts || n_units <= 0)
        return fail(CGRB_E_BADARG, "cgrb_counts_reset: bad argument%s");
    LAUNCH("k_counts_reset", (cudaStream_t)stream, k_counts_reset<<<(n_units + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_counts, n_units));
    return check_launch("cgrb_counts_reset");
}

int cgrb_energy_integrated_evolution(const double *d_dtab, const cgrb_unit_t *d_units, int unit_index,
                                     int interp_extrap, const double *d_times, int64_t n, double *d_out,
                                     void *stream)
{
    (void)interp_extrap;   // the grid effective areas in d_dtab were interpolated on the host
    if (!d_dtab || !d_units || !d_times || !d_out || n < 0 || unit_index < 0)
        return fail(CGRB_E_BADARG, "cgrb_energy_integrated_evolution: bad argument%s");
    if (n == 0)
        return CGRB_OK;
    int grid = (int)min((int64_t)4096, (n + NT - 1) / NT);
    LAUNCH("k_lambda", (cudaStream_t)stream, k_lambda<<<grid, NT, 0, (cudaStream_t)stream>>>(d_dtab, d_units, unit_index, d_times, n, d_out));
    return check_launch("cgrb_energy_integrated_evolution");
}

int cgrb_evolution(const cgrb_unit_t *d_units, int unit_index, const double *d_energy, int64_t n_energy,
                   const double *d_times, int64_t n_time, double *d_out, void *stream)
{
    if (!d_units || !d_energy || !d_times || !d_out || n_energy < 0 || n_time < 0 || unit_index < 0)
        return fail(CGRB_E_BADARG, "cgrb_evolution: bad argument%s");
    if (n_energy == 0 || n_time == 0)
        return CGRB_OK;
    int grid = (int)min((int64_t)4096, n_time);
    LAUNCH("k_evolution", (cudaStream_t)stream, k_evolution<<<grid, NT, 0, (cudaStream_t)stream>>>(d_units, unit_index, d_energy, n_energy, d_times,
                                                       n_time, d_out));
    return check_launch("cgrb_evolution");
}

int cgrb_fmax(const double *d_dtab, cgrb_unit_t *d_units, int n_units, int interp_extrap, void *stream)
{
    (void)interp_extrap;
    if (!d_dtab || !d_units || n_units <= 0)
        return fail(CGRB_E_BADARG, "cgrb_fmax: bad argument%s");
    LAUNCH("k_fmax", (cudaStream_t)stream, k_fmax<<<n_units, 64, 0, (cudaStream_t)stream>>>(d_dtab, d_units, n_units));
    return check_launch("cgrb_fmax");
}

int64_t cgrb_thin_cum_bytes(int64_t n_nodes_total) { return 20 * (n_nodes_total > 0 ? n_nodes_total : 0) + 8; }

int cgrb_thin_tables(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work,
                     int64_t n_work, double *d_thin_tab, void *d_thin_cum, int64_t n_nodes_total, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_thin_tab ||
        (d_thin_cum && n_nodes_total <= 0))
        return fail(CGRB_E_BADARG, "cgrb_thin_tables: bad argument%s");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH("k_thin_nodes", s, k_thin_nodes<<<(unsigned)n_work, NT, 0, s>>>(d_dtab, d_units, d_work, n_work, d_thin_tab));
    LAUNCH("k_thin_margins", s, k_thin_margins<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, d_thin_tab));
    if (d_thin_cum) {
        LAUNCH("k_thin_cum", s, k_thin_cum<<<n_units, NT, 0, s>>>(d_units, n_units, d_thin_tab, (double *)d_thin_cum));
        LAUNCH("k_thin_guide", s, k_thin_guide<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, (double *)d_thin_cum,
                                                                               n_nodes_total));
    }
    return check_launch("cgrb_thin_tables");
}

static int check_rng(const cgrb_rng_t *rng, const char *who)
{
    if (!rng)
        return fail(CGRB_E_BADARG, "%s: null rng", who);
    if (rng->mode != 0 && rng->mode != 1)
        return fail(CGRB_E_BADARG, "%s: rng mode must be 0 (philox) or 1 (injected)", who);
    if (rng->mode == 1 && !rng->d_u)
        return fail(CGRB_E_BADARG, "%s: injected rng without uniform buffer", who);
    return CGRB_OK;
}

int cgrb_sample_events(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                       const cgrb_work_t *d_work, int64_t n_work, const int64_t *d_unit_work_begin,
                       int interp_extrap, const cgrb_rng_t *rng, double *d_segsum, double *d_segstart,
                       int32_t *d_segcount, double *d_stage, double *d_times_src, double *d_cand_times,
                       uint8_t *d_cand_accept, const double *d_thin_tab, const void *d_thin_cum,
                       int64_t n_nodes_total, cgrb_counts_t *d_counts, void *stream)
{
    (void)interp_extrap;
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_unit_work_begin || !d_segsum ||
        !d_segstart || !d_segcount || !d_stage || !d_times_src || !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_sample_events: bad argument%s");
    if ((d_cand_times == nullptr) != (d_cand_accept == nullptr))
        return fail(CGRB_E_BADARG, "cgrb_sample_events: cand_times and cand_accept go together%s");
    int rc = check_rng(rng, "cgrb_sample_events");
    if (rc)
        return rc;
    if (d_thin_cum && (!d_thin_tab || n_nodes_total <= 0))
        return fail(CGRB_E_BADARG, "cgrb_sample_events: the majorant table needs the squeeze table and its node count%s");
    cudaStream_t s = (cudaStream_t)stream;
    const double *cum = (rng->mode == 0) ? (const double *)d_thin_cum : nullptr;
    LAUNCH("k_gap_prefix<0>", s, k_gap_prefix<0><<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, *rng, d_segsum, d_stage, d_counts,
                                                                                 cum, n_nodes_total));
    LAUNCH("k_seg_scan<0>", s, k_seg_scan<0><<<(n_units * 32 + 127) / 128, 128, 0, s>>>(d_units, n_units, d_unit_work_begin, d_segsum,
                                                           d_segstart, d_counts, *rng, cum));
    LAUNCH("k_source_thin", s, k_source_thin<<<(unsigned)n_work, NT, 0, s>>>(d_dtab, d_units, d_work, n_work, *rng, d_segstart,
                                                  d_segcount, d_stage, d_cand_times, d_cand_accept, d_thin_tab, d_counts,
                                                  cum, n_nodes_total));
    LAUNCH("k_source_gather", s, k_source_gather<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, d_unit_work_begin, d_segcount,
                                                    d_stage, d_times_src, d_counts));
    return check_launch("cgrb_sample_events");
}

int cgrb_sample_energy(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                       const cgrb_work_t *d_work, int64_t n_work, int interp_extrap, const cgrb_rng_t *rng,
                       int64_t max_trials, const double *d_times_src, double *d_energy_src, int32_t *d_trials,
                       cgrb_counts_t *d_counts, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_times_src || !d_energy_src ||
        !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_sample_energy: bad argument%s");
    if (interp_extrap != 0 && interp_extrap != 1)
        return fail(CGRB_E_BADARG, "cgrb_sample_energy: interp_extrap must be 0 (linear) or 1 (clamp)%s");
    int rc = check_rng(rng, "cgrb_sample_energy");
    if (rc)
        return rc;
    if (rng->mode == 1 && !rng->d_off)
        return fail(CGRB_E_BADARG, "cgrb_sample_energy: injected mode needs per-photon offsets%s");
    LAUNCH("k_sample_energy", (cudaStream_t)stream, k_sample_energy<<<(unsigned)n_work, NT, 0, (cudaStream_t)stream>>>(
        d_dtab, d_units, d_work, n_work, interp_extrap, *rng, max_trials, d_times_src, d_energy_src, d_trials,
        d_counts));
    return check_launch("cgrb_sample_energy");
}

int64_t cgrb_energy_workspace_bytes(int n_units, int64_t n_rows)
{
    int64_t tabs = (int64_t)sizeof(EnergyUnitTab) * (n_units > 0 ? n_units : 0);
    tabs = (tabs + 255) / 256 * 256;
    const int64_t nr = n_rows > 0 ? n_rows : 0;
    return tabs + (int64_t)sizeof(double) * ENV_ROW * nr + (int64_t)CGRB_ENV_SEGS * nr;
}

// what: bit 0 = build the per-unit tables + rows, bit 1 = sample
static int energy_envelope_impl(int what, const char *who, const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                                const cgrb_work_t *d_work, int64_t n_work, int interp_extrap,
                                const cgrb_rng_t *rng, int64_t max_trials, void *d_workspace, int64_t n_rows,
                                const double *d_times_src, double *d_energy_src, uint8_t *d_pha_src,
                                int32_t *d_trials, cgrb_counts_t *d_counts, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_workspace || n_rows < n_units)
        return fail(CGRB_E_BADARG, "%s: bad argument", who);
    if (interp_extrap != 0 && interp_extrap != 1)
        return fail(CGRB_E_BADARG, "%s: interp_extrap must be 0 (linear) or 1 (clamp)", who);
    if (what & 2) {
        if (!d_work || n_work <= 0 || !d_times_src || (!d_energy_src && !d_pha_src) || !d_counts)
            return fail(CGRB_E_BADARG, "%s: bad argument", who);
        int rc = check_rng(rng, who);
        if (rc)
            return rc;
        if (rng->mode != 0)
            return fail(CGRB_E_UNSUPPORTED, "%s: injected uniforms replay the literal sampler; use cgrb_sample_energy", who);
    }
    cudaStream_t s = (cudaStream_t)stream;
    EnergyUnitTab *etab = (EnergyUnitTab *)d_workspace;
    int64_t tabs = ((int64_t)sizeof(EnergyUnitTab) * n_units + 255) / 256 * 256;
    double *rows = (double *)((char *)d_workspace + tabs);
    uint8_t *row_guides = (uint8_t *)(rows + (size_t)ENV_ROW * n_rows);
    if (what & 1) {
    LAUNCH("k_energy_tables", s, k_energy_tables<<<n_units, 256, 0, s>>>(d_dtab, d_units, n_units, interp_extrap, etab));
    LAUNCH("k_energy_rows", s, k_energy_rows<<<(unsigned)n_rows, NT, 0, s>>>(d_units, n_units, etab, rows, row_guides, n_rows));
    }
    if (!(what & 2))
        return check_launch(who);
#define ENV_LAUNCH(F, D, M)                                                                                       \
    LAUNCH("k_sample_energy_env", s, (k_sample_energy_env<F, D, M><<<(unsigned)n_work, NT, 0, s>>>(                \
        d_dtab, d_units, d_work, n_work, interp_extrap, *rng, max_trials, etab, rows, row_guides, d_times_src,    \
        d_energy_src, d_pha_src, d_trials, d_counts)))
    const bool diag = d_energy_src != nullptr || d_trials != nullptr;
    // resident CTAs per SM the production variant is compiled for (register budget); CGRB_ENV_MIN_CTAS=2|3|4
    // switches it for tuning runs
    static int minb = 0;
    if (!minb) {
        const char *e = getenv("CGRB_ENV_MIN_CTAS");
        minb = (e && e[0] == '2') ? 2 : (e && e[0] == '3') ? 3 : (e && e[0] == '4') ? 4 : ENV_MIN_CTAS;
    }
    if (d_pha_src && diag)
        ENV_LAUNCH(true, true, ENV_MIN_CTAS);
    else if (d_pha_src) {
        if (minb == 2)
            ENV_LAUNCH(true, false, 2);
        else if (minb == 4)
            ENV_LAUNCH(true, false, 4);
        else
            ENV_LAUNCH(true, false, 3);
    } else
        ENV_LAUNCH(false, true, ENV_MIN_CTAS);
#undef ENV_LAUNCH
    return check_launch(who);
}

int cgrb_sample_energy_envelope(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                                const cgrb_work_t *d_work, int64_t n_work, int interp_extrap,
                                const cgrb_rng_t *rng, int64_t max_trials, void *d_workspace, int64_t n_rows,
                                const double *d_times_src, double *d_energy_src, uint8_t *d_pha_src,
                                int32_t *d_trials, cgrb_counts_t *d_counts, void *stream)
{
    return energy_envelope_impl(3, "cgrb_sample_energy_envelope", d_dtab, d_units, n_units, d_work, n_work, interp_extrap, rng,
                                max_trials, d_workspace, n_rows, d_times_src, d_energy_src, d_pha_src, d_trials, d_counts, stream);
}

int cgrb_energy_envelope_tables(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, int interp_extrap,
                                void *d_workspace, int64_t n_rows, void *stream)
{
    return energy_envelope_impl(1, "cgrb_energy_envelope_tables", d_dtab, d_units, n_units, nullptr, 0, interp_extrap, nullptr, 0,
                                d_workspace, n_rows, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

int cgrb_sample_energy_envelope_prepared(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                                         const cgrb_work_t *d_work, int64_t n_work, int interp_extrap,
                                         const cgrb_rng_t *rng, int64_t max_trials, void *d_workspace, int64_t n_rows,
                                         const double *d_times_src, double *d_energy_src, uint8_t *d_pha_src,
                                         int32_t *d_trials, cgrb_counts_t *d_counts, void *stream)
{
    return energy_envelope_impl(2, "cgrb_sample_energy_envelope_prepared", d_dtab, d_units, n_units, d_work, n_work, interp_extrap,
                                rng, max_trials, d_workspace, n_rows, d_times_src, d_energy_src, d_pha_src, d_trials, d_counts,
                                stream);
}

int cgrb_digitize(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work,
                  int64_t n_work, const cgrb_rng_t *rng, const double *d_energy_src, const int32_t *d_trials,
                  uint8_t *d_pha_src, const cgrb_counts_t *d_counts, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_energy_src || !d_pha_src ||
        !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_digitize: bad argument%s");
    int rc = check_rng(rng, "cgrb_digitize");
    if (rc)
        return rc;
    LAUNCH("k_digitize", (cudaStream_t)stream, k_digitize<<<(unsigned)n_work, NT, 0, (cudaStream_t)stream>>>(d_dtab, d_units, d_work, n_work, *rng,
                                                                     d_energy_src, d_trials, d_pha_src, d_counts));
    return check_launch("cgrb_digitize");
}

// state words [n_tiles] (8 B), header (64 B), tile records [n_tiles] (64 B), the level-major order's tables:
// cum (int64 [n_tiles + 2]), hist / above (int32 [n_tiles + 2] each), ubeg (int64 [n_units + 1]), sorted / eq (int32 [n_units])
int64_t cgrb_sample_background_workspace_bytes(int n_units, int64_t n_tiles)
{
    if (n_units < 0 || n_tiles < 0)
        return 0;
    return 88 * (n_tiles + 2) + 16 * ((int64_t)n_units + 1) + 256;
}

int cgrb_sample_background(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, int64_t n_tiles,
                           const cgrb_rng_t *rng_time, const cgrb_rng_t *rng_chan, void *d_workspace,
                           double *d_times_bkg, uint8_t *d_pha_bkg, cgrb_counts_t *d_counts, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || n_tiles <= 0 || !d_workspace || !d_times_bkg || !d_pha_bkg || !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_sample_background: bad argument%s");
    int rc = check_rng(rng_time, "cgrb_sample_background");
    if (rc)
        return rc;
    cgrb_rng_t chan = *rng_time;
    if (rng_time->mode == 1) {
        rc = check_rng(rng_chan, "cgrb_sample_background(chan)");
        if (rc)
            return rc;
        if (rng_chan->mode != 1)
            return fail(CGRB_E_BADARG, "cgrb_sample_background: channel rng must be injected too%s");
        chan = *rng_chan;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = (char *)d_workspace;
    unsigned long long *state = (unsigned long long *)ws;
    LevelOrder o;
    o.header = (int64_t *)(ws + 8 * (size_t)n_tiles);
    char *q = ws + 8 * (size_t)n_tiles + 64;
    BkgTile *tiles = (BkgTile *)q;
    q += 64 * (size_t)n_tiles;
    o.cum = (int64_t *)q;
    q += 8 * ((size_t)n_tiles + 2);
    o.ubeg = (int64_t *)q;
    q += 8 * ((size_t)n_units + 1);
    o.hist = (int32_t *)q;
    q += 4 * ((size_t)n_tiles + 2);
    o.above = (int32_t *)q;
    q += 4 * ((size_t)n_tiles + 2);
    o.sorted = (int32_t *)q;
    q += 4 * (size_t)n_units;
    o.eq = (int32_t *)q;
    o.cap = n_tiles;
    // state words: "not ready"; header + histogram: zero
    cudaError_t e = cudaMemsetAsync(state, 0xFF, 8 * (size_t)n_tiles, s);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(o.header, 0, 64, s);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(o.hist, 0, 4 * ((size_t)n_tiles + 2), s);
    if (e != cudaSuccess)
        return fail(CGRB_E_CUDA, "cgrb_sample_background: memset: %s", cudaGetErrorString(e));
    LAUNCH("k_background_order", s, k_level_order<1><<<1, ORDER_NT, 0, s>>>(d_units, nullptr, n_units, d_counts, o));
    LAUNCH("k_background_plan", s, k_background_plan<<<(unsigned)((n_tiles + NT - 1) / NT), NT, 0, s>>>(d_units, n_tiles, o, tiles));
    LAUNCH("k_background", s, k_background<<<(unsigned)n_tiles, NT, 0, s>>>(d_dtab, tiles, *rng_time, chan, state,
                                                                         d_times_bkg, d_pha_bkg, d_counts));
    return check_launch("cgrb_sample_background");
}

int cgrb_background_channels(const double *d_dtab, int det, uint32_t stream_id, const cgrb_rng_t *rng, int64_t n,
                             uint8_t *d_out, void *stream)
{
    if (!d_dtab || det < 0 || n < 0 || !d_out)
        return fail(CGRB_E_BADARG, "cgrb_background_channels: bad argument%s");
    int rc = check_rng(rng, "cgrb_background_channels");
    if (rc)
        return rc;
    if (n == 0)
        return CGRB_OK;
    int grid = (int)min((int64_t)(148 * 8), (n + NT - 1) / NT);
    LAUNCH("k_background_channels", (cudaStream_t)stream,
           k_background_channels<<<grid, NT, 0, (cudaStream_t)stream>>>(d_dtab, det, stream_id, *rng, n, d_out));
    return check_launch("cgrb_background_channels");
}

int cgrb_combine(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                 const double *d_times_bkg, const uint8_t *d_pha_bkg, const double *d_times_src,
                 const uint8_t *d_pha_src, double *d_times_tot, uint8_t *d_pha_tot, cgrb_counts_t *d_counts,
                 void *stream)
{
    if (!d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_times_bkg || !d_pha_bkg || !d_times_src ||
        !d_pha_src || !d_times_tot || !d_pha_tot || !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_combine: bad argument%s");
    LAUNCH("k_combine", (cudaStream_t)stream, k_combine<<<(unsigned)n_work, NT, 0, (cudaStream_t)stream>>>(d_units, d_work, n_work, d_times_bkg,
                                                                 d_pha_bkg, d_times_src, d_pha_src, d_times_tot,
                                                                 d_pha_tot, d_counts));
    return check_launch("cgrb_combine");
}

// validates a dead-time model; *physical = 1 for the non-paralyzable model
static int deadtime_model(const cgrb_deadtime_model_t *model, DtModel *md, int *physical, const char *who)
{
    md->tau = md->tau_ovf = md->tau_max = 0.0;
    *physical = 0;
    if (!model || model->mode == CGRB_DEADTIME_REFERENCE)
        return CGRB_OK;
    if (model->mode != CGRB_DEADTIME_PHYSICAL)
        return fail(CGRB_E_BADARG, "%s: unknown dead-time mode", who);
    if (!(model->tau >= 0.0) || !(model->tau_overflow >= 0.0) || !(model->tau < 1e300) ||
        !(model->tau_overflow < 1e300))
        return fail(CGRB_E_BADARG, "%s: dead-time windows must be finite and >= 0", who);
    md->tau = model->tau;
    md->tau_ovf = model->tau_overflow;
    md->tau_max = model->tau > model->tau_overflow ? model->tau : model->tau_overflow;
    *physical = 1;
    return CGRB_OK;
}

int cgrb_gbm_dead_time(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                       const int64_t *d_unit_work_begin, const double *d_times_tot, const uint8_t *d_pha_tot,
                       int32_t *d_tile_count, double *d_times_out, uint8_t *d_pha_out, uint8_t *d_selection,
                       cgrb_counts_t *d_counts, const cgrb_deadtime_model_t *model, void *stream)
{
    if (!d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_unit_work_begin || !d_times_tot ||
        !d_pha_tot || !d_tile_count || !d_times_out || !d_pha_out || !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_gbm_dead_time: bad argument%s");
    DtModel md;
    int physical;
    if (int rc = deadtime_model(model, &md, &physical, "cgrb_gbm_dead_time"))
        return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!physical) {
        LAUNCH("k_deadtime_count", s, k_deadtime_count<0><<<(unsigned)n_work, NT, 0, s>>>(
            d_units, d_work, n_work, d_times_tot, d_pha_tot, d_tile_count, d_counts, md));
        LAUNCH("k_deadtime_write", s, k_deadtime_write<0><<<(unsigned)n_work, NT, 0, s>>>(
            d_units, d_work, n_work, d_unit_work_begin, d_times_tot, d_pha_tot, d_tile_count, d_times_out,
            d_pha_out, d_selection, d_counts, md));
    } else {
        LAUNCH("k_deadtime_count_phys", s, k_deadtime_count<1><<<(unsigned)n_work, NT, 0, s>>>(
            d_units, d_work, n_work, d_times_tot, d_pha_tot, d_tile_count, d_counts, md));
        LAUNCH("k_deadtime_write_phys", s, k_deadtime_write<1><<<(unsigned)n_work, NT, 0, s>>>(
            d_units, d_work, n_work, d_unit_work_begin, d_times_tot, d_pha_tot, d_tile_count, d_times_out,
            d_pha_out, d_selection, d_counts, md));
        LAUNCH("k_deadtime_phys_fixup", s, k_deadtime_phys_fixup<<<(n_units * 32 + 127) / 128, 128, 0, s>>>(
            d_units, n_units, d_times_tot, d_pha_tot, nullptr, nullptr, nullptr, nullptr, d_times_out, d_pha_out,
            d_selection, d_counts, md));
    }
    return check_launch("cgrb_gbm_dead_time");
}

// tile states [n_work] (8 B each), header (64 B), tile records [n_work] (64 B each), the level-major order's tables:
// hist / above (int32 [n_work + 2] each), cum (int64 [n_work + 2]), sorted / eq (int32 [n_units <= n_work] each)
int64_t cgrb_merge_dead_time_workspace_bytes(int64_t n_work) { return 96 * (n_work > 0 ? n_work : 0) + 512; }

int cgrb_merge_dead_time(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                         const int64_t *d_unit_work_begin, const double *d_times_bkg, const uint8_t *d_pha_bkg,
                         const double *d_times_src, const uint8_t *d_pha_src, void *d_workspace,
                         double *d_times_out, uint8_t *d_pha_out, cgrb_counts_t *d_counts,
                         const cgrb_deadtime_model_t *model, void *stream)
{
    if (!d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_unit_work_begin || !d_times_bkg || !d_pha_bkg ||
        !d_times_src || !d_pha_src || !d_workspace || !d_times_out || !d_pha_out || !d_counts || n_units > n_work)
        return fail(CGRB_E_BADARG, "cgrb_merge_dead_time: bad argument%s");
    DtModel md;
    int physical;
    if (int rc = deadtime_model(model, &md, &physical, "cgrb_merge_dead_time"))
        return rc;
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = (char *)d_workspace;
    unsigned long long *state = (unsigned long long *)ws;
    LevelOrder o;
    o.ubeg = nullptr;
    o.header = (int64_t *)(ws + 8 * (size_t)n_work);
    MdTile *tiles = (MdTile *)(ws + 8 * (size_t)n_work + 64);
    char *q = (char *)(tiles + n_work);
    o.cum = (int64_t *)q;
    q += 8 * ((size_t)n_work + 2);
    o.hist = (int32_t *)q;
    q += 4 * ((size_t)n_work + 2);
    o.above = (int32_t *)q;
    q += 4 * ((size_t)n_work + 2);
    o.sorted = (int32_t *)q;
    q += 4 * (size_t)n_work;
    o.eq = (int32_t *)q;
    o.cap = n_work;
    // zero the tile states + header, and the histogram
    cudaError_t e = cudaMemsetAsync(ws, 0, 8 * (size_t)n_work + 64, s);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(o.hist, 0, 4 * ((size_t)n_work + 2), s);
    if (e != cudaSuccess)
        return fail(CGRB_E_CUDA, "cgrb_merge_dead_time: memset: %s", cudaGetErrorString(e));
    LAUNCH("k_merge_order", s, k_level_order<0><<<1, ORDER_NT, 0, s>>>(d_units, d_unit_work_begin, n_units, d_counts, o));
    LAUNCH("k_merge_plan", s, k_merge_plan<<<(unsigned)((n_work + NT - 1) / NT), NT, 0, s>>>(
        d_units, n_work, d_unit_work_begin, d_times_bkg, d_times_src, o, tiles, d_counts));
    const unsigned md_grid = (unsigned)n_work;      // one tile per CTA; blockIdx.x = ticket in the level-major order
    if (!physical) {
        LAUNCH("k_merge_deadtime", s, k_merge_deadtime<0><<<md_grid, NT, 0, s>>>(
            d_units, d_times_bkg, d_pha_bkg, d_times_src, d_pha_src, state, tiles, d_times_out, d_pha_out, d_counts, md));
    } else {
        LAUNCH("k_merge_deadtime_phys", s, k_merge_deadtime<1><<<md_grid, NT, 0, s>>>(
            d_units, d_times_bkg, d_pha_bkg, d_times_src, d_pha_src, state, tiles, d_times_out, d_pha_out, d_counts, md));
        LAUNCH("k_deadtime_phys_fixup", s, k_deadtime_phys_fixup<<<(n_units * 32 + 127) / 128, 128, 0, s>>>(
            d_units, n_units, nullptr, nullptr, d_times_bkg, d_pha_bkg, d_times_src, d_pha_src, d_times_out,
            d_pha_out, nullptr, d_counts, md));
    }
    return check_launch("cgrb_merge_dead_time");
}

int64_t cgrb_gbm_trigger_workspace_bytes(int n_units, int64_t bins_per_unit)
{
    if (n_units <= 0 || bins_per_unit <= 0)
        return 0;
    const int64_t nu = n_units;
    const int64_t cstride = ((bins_per_unit + 4 + 3) / 4) * 4;       // prefix rows: 3 pad + (bins + 1) entries, a multiple of 4
    return nu * TRIG_NARR * bins_per_unit * 4 + nu * 4 * cstride * 4 + nu * cstride * 8 + 768;
}

int cgrb_gbm_trigger(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                     const int32_t *d_chan_bounds, const cgrb_trigger_par_t *par, const double *d_times_out,
                     const uint8_t *d_pha_out, const cgrb_counts_t *d_counts, int64_t bins_per_unit,
                     void *d_workspace, cgrb_trigger_t *d_out, void *stream)
{
    if (!d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_chan_bounds || !par || !d_times_out || !d_pha_out ||
        !d_counts || bins_per_unit <= 0 || !d_workspace || !d_out)
        return fail(CGRB_E_BADARG, "cgrb_gbm_trigger: bad argument%s");
    if (!(par->base_timescale > 0.0) || !(par->background_duration > 0.0) || !(par->pre_window > 0.0))
        return fail(CGRB_E_BADARG, "cgrb_gbm_trigger: bad trigger parameters%s");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nu = n_units;
    int32_t *bins = (int32_t *)d_workspace;
    const size_t bins_bytes = (size_t)nu * TRIG_NARR * bins_per_unit * 4;
    const int64_t cstride = ((bins_per_unit + 4 + 3) / 4) * 4;
    int32_t *csum = (int32_t *)((char *)d_workspace + ((bins_bytes + 255) / 256) * 256);
    double *esum = (double *)((char *)csum + (((size_t)nu * 4 * cstride * 4 + 255) / 256) * 256);
    cudaError_t e = cudaMemsetAsync(bins, 0, bins_bytes, s);
    if (e != cudaSuccess)
        return fail(CGRB_E_CUDA, "cgrb_gbm_trigger: memset: %s", cudaGetErrorString(e));
    LAUNCH("k_trig_bin", s, k_trig_bin<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, d_chan_bounds,
                                                                      par->base_timescale, d_times_out, d_pha_out,
                                                                      d_counts, bins_per_unit, bins));
    e = cudaFuncSetAttribute(k_trig_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, TRIG_EVAL_SMEM);     // > 48 KB: opt-in
    if (e != cudaSuccess)
        return fail(CGRB_E_CUDA, "cgrb_gbm_trigger: shared memory opt-in: %s", cudaGetErrorString(e));
    LAUNCH("k_trig_eval", s, k_trig_eval<<<n_units, NT, TRIG_EVAL_SMEM, s>>>(d_units, n_units, *par, d_times_out, d_counts,
                                                               bins_per_unit, bins, cstride, csum, esum, d_out));
    return check_launch("cgrb_gbm_trigger");
}

}  // extern "C"

