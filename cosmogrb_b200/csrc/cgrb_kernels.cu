// cgrb_kernels.cu -- sm_100a kernels + C ABI of the photon-event engine.
//
// One translation unit; see include/cosmogrb_b200.h for the boundary and DESIGN.md for the
// data layout and the roofline of each kernel.  Reference lines are relative to the cosmogrb
// repository root.
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

#define NT 256            // threads per CTA of the event kernels
#define NW (NT / 32)

// ==========================================================================================
// shared helpers
// ==========================================================================================

// inclusive block scan of doubles over NT threads; s_w holds NW doubles.  Returns the inclusive
// value, *total = sum over the block.  Contains two __syncthreads and one trailing barrier.
__device__ __forceinline__ double block_incl_scan(double v, double *s_w, double *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_incl_scan(v, lane);
    if (lane == 31)
        s_w[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double w = (lane < NW) ? s_w[lane] : 0.0;
        w = warp_incl_scan(w, lane);
        if (lane < NW)
            s_w[lane] = w;
    }
    __syncthreads();
    double prefix = (wid > 0) ? s_w[wid - 1] : 0.0;
    *total = s_w[NW - 1];
    double r = prefix + v;
    __syncthreads();
    return r;
}

__device__ __forceinline__ int block_incl_scan_i(int v, int *s_w, int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_incl_scan_i(v, lane);
    if (lane == 31)
        s_w[wid] = v;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < NW) ? s_w[lane] : 0;
        w = warp_incl_scan_i(w, lane);
        if (lane < NW)
            s_w[lane] = w;
    }
    __syncthreads();
    int prefix = (wid > 0) ? s_w[wid - 1] : 0;
    *total = s_w[NW - 1];
    int r = prefix + v;
    __syncthreads();
    return r;
}

__device__ __forceinline__ long long block_sum_ll(long long v, long long *s_w)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0)
        s_w[wid] = v;
    __syncthreads();
    long long t = 0;
#pragma unroll
    for (int w = 0; w < NW; w++)
        t += s_w[w];
    __syncthreads();
    return t;
}

// exclusive block scan of doubles (sum of the values of the threads before this one, fixed association)
__device__ __forceinline__ double block_excl_scan(double v, double *s_w, double *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double incl = warp_incl_scan(v, lane);
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0)
        excl = 0.0;
    if (lane == 31)
        s_w[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        double w = (lane < NW) ? s_w[lane] : 0.0;
        w = warp_incl_scan(w, lane);
        if (lane < NW)
            s_w[lane] = w;
    }
    __syncthreads();
    const double prefix = (wid > 0) ? s_w[wid - 1] : 0.0;
    *total = s_w[NW - 1];
    const double r = prefix + excl;
    __syncthreads();
    return r;
}

// Thinning uniforms.  Candidate k of a unit consumes u1 (gap) and u2 (acceptance test for the source,
// channel draw for the background).  Under Philox one call serves the candidate PAIR (2m, 2m+1):
// counter (index m, sub 0) -> (u1 of 2m, u1 of 2m+1), (index m, sub 1) -> (u2 of 2m, u2 of 2m+1), so
// the gap pass and the decision pass each cost half a Philox call per candidate.  Injected streams
// keep the reference's order u1, u2, u1, u2, ... (Appendix A of SURVEY.md): candidate k reads
// u[2k], u[2k+1]; the stream may end on an odd count.
struct PairDraw {
    double a, b;        // values for candidates 2m and 2m+1
    bool exhausted;
};
__device__ __forceinline__ double injected_thin_uniform(const cgrb_rng_t &rng, int64_t off, int64_t len, int64_t k,
                                                        int which, bool *exhausted)
{
    const int64_t pos = 2 * k + which;
    if (pos < len)
        return rng.d_u[off + pos];
    if (which == 1)
        return 0.5;                 // never consulted: the candidate's u1 was the last uniform or beyond
    *exhausted = true;
    return 0.0;                     // -log(0) = +inf: terminates the unit
}
__device__ __forceinline__ PairDraw pair_draw(const cgrb_rng_t &rng, int unit, uint32_t stream_id, uint32_t stage,
                                              int64_t m, int which)
{
    PairDraw d;
    d.exhausted = false;
    if (rng.mode == 0) {
        Philox4 p = philox_at(make_key(rng.seed), stream_id, stage, (uint64_t)m, (uint32_t)which);
        d.a = u53(p.x, p.y);
        d.b = u53(p.z, p.w);
    } else {
        const int64_t off = rng.d_off ? rng.d_off[unit] : 0;
        const int64_t len = rng.d_len ? rng.d_len[unit] : ((int64_t)1 << 62);
        d.a = injected_thin_uniform(rng, off, len, 2 * m, which, &d.exhausted);
        d.b = injected_thin_uniform(rng, off, len, 2 * m + 1, which, &d.exhausted);
    }
    return d;
}
#define THIN_TILE (2 * NT)      // candidates per tile: two per thread

// per-unit tables for lambda(t): lambda = norm * ec^-alpha * sum_m c_m exp(-x_m / ec)
struct LambdaTab {
    double c[CGRB_NE_LAMBDA];
    double x[CGRB_NE_LAMBDA];
};

// builds the table cooperatively; caller syncs afterwards
__device__ __forceinline__ void build_lambda_tab(LambdaTab *tab, const double *dt, const cgrb_unit_t &u)
{
    const double *E = dt + CGRB_DT_G75_E;
    const double *A = dt + CGRB_DT_G75_A;
    const double opz = 1.0 + u.z;
    for (int m = threadIdx.x; m < CGRB_NE_LAMBDA; m += blockDim.x) {
        // np.trapz weights of cpl_functions.py:325
        double w;
        if (m == 0)
            w = 0.5 * (E[1] - E[0]);
        else if (m == CGRB_NE_LAMBDA - 1)
            w = 0.5 * (E[m] - E[m - 1]);
        else
            w = 0.5 * (E[m] - E[m - 1]) + 0.5 * (E[m + 1] - E[m]);
        double x = E[m] * opz;                       // cpl_functions.py:95
        tab->x[m] = x;
        tab->c[m] = w * (A[m] * u.ea_scale) * pow(x, u.alpha);
    }
}

__device__ __forceinline__ double lambda_at(const LambdaTab *tab, const cgrb_unit_t &u, double t)
{
    SpecT s = spec_at(u, t);
    double acc = 0.0;
#pragma unroll 5
    for (int m = 0; m < CGRB_NE_LAMBDA; m++)
        acc += tab->c[m] * exp(-tab->x[m] * s.inv_ec);
    return s.norm * exp(-u.alpha * s.log_ec) * acc;
}

// ==========================================================================================
// spectral model kernels
// ==========================================================================================

// SourceFunction.energy_integrated_evolution at arbitrary times
__global__ void __launch_bounds__(NT) k_lambda(const double *__restrict__ dtab, const cgrb_unit_t *units,
                                               int unit_index, const double *__restrict__ times,
                                               int64_t n, double *__restrict__ out)
{
    __shared__ LambdaTab tab;
    __shared__ cgrb_unit_t u;
    if (threadIdx.x == 0)
        u = units[unit_index];
    __syncthreads();
    build_lambda_tab(&tab, dtab + (size_t)u.det * CGRB_DT_SIZE, u);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT)
        out[i] = lambda_at(&tab, u, times[i]);
}

// SourceFunction.evolution(energy, time): cpl_evolution without the response
__global__ void __launch_bounds__(NT) k_evolution(const cgrb_unit_t *units, int unit_index,
                                                  const double *__restrict__ energy, int64_t n_energy,
                                                  const double *__restrict__ times, int64_t n_time,
                                                  double *__restrict__ out)
{
    __shared__ cgrb_unit_t u;
    if (threadIdx.x == 0)
        u = units[unit_index];
    __syncthreads();
    const double opz = 1.0 + u.z;
    for (int64_t it = blockIdx.x; it < n_time; it += gridDim.x) {
        SpecT s = spec_at(u, times[it]);
        for (int64_t m = threadIdx.x; m < n_energy; m += NT) {
            double x = energy[m] * opz;
            out[it * n_energy + m] = cpl_value(s, u.alpha, x, log(x));
        }
    }
}

// Source._get_energy_integrated_max: one CTA per unit, 50-point time grid (source.py:88-110)
__global__ void __launch_bounds__(64) k_fmax(const double *__restrict__ dtab, cgrb_unit_t *units, int n_units)
{
    __shared__ LambdaTab tab;
    __shared__ cgrb_unit_t u;
    __shared__ double s_f[64];
    const int ui = blockIdx.x;
    if (ui >= n_units)
        return;
    if (threadIdx.x == 0)
        u = units[ui];
    __syncthreads();
    build_lambda_tab(&tab, dtab + (size_t)u.det * CGRB_DT_SIZE, u);
    __syncthreads();
    double f = -INFINITY;
    if (threadIdx.x < CGRB_NT_FMAX) {
        // np.linspace(tstart, tstop, 50)
        double step = (u.src_tstop - u.src_tstart) / (double)(CGRB_NT_FMAX - 1);
        double t = (threadIdx.x == CGRB_NT_FMAX - 1) ? u.src_tstop : u.src_tstart + threadIdx.x * step;
        f = lambda_at(&tab, u, t);
    }
    s_f[threadIdx.x] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        double best = s_f[0];
        bool has_nan = isnan(best);
        for (int i = 1; i < CGRB_NT_FMAX; i++) {
            has_nan |= isnan(s_f[i]);
            if (s_f[i] > best)
                best = s_f[i];
        }
        units[ui].fmax = has_nan ? nan("") : best;   // np.max propagates NaN
    }
}

// ==========================================================================================
// squeeze table of the thinning acceptance probability p(t) = lambda(t) / fmax (see the header):
// nodes on a uniform time grid, then per-interval interpolation-error margins from second differences.
// ==========================================================================================
__global__ void __launch_bounds__(NT) k_thin_nodes(const double *__restrict__ dtab,
                                                   const cgrb_unit_t *__restrict__ units,
                                                   const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                   double *__restrict__ thin_tab)
{
    __shared__ LambdaTab tab;
    __shared__ cgrb_unit_t u;
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    if (threadIdx.x == 0)
        u = units[wk.unit];
    __syncthreads();
    if (u.tab_n < 4 || (u.flags & CGRB_UNIT_NO_SOURCE))
        return;
    build_lambda_tab(&tab, dtab + (size_t)u.det * CGRB_DT_SIZE, u);
    __syncthreads();
    const int64_t k = wk.start + threadIdx.x;
    if (k >= u.tab_n)
        return;
    const double h = (u.src_tstop - u.src_tstart) / (double)(u.tab_n - 1);
    const double t = (k == u.tab_n - 1) ? u.src_tstop : u.src_tstart + (double)k * h;
    thin_tab[2 * (u.tab_off + k)] = lambda_at(&tab, u, t) / u.fmax;
}

__global__ void __launch_bounds__(NT) k_thin_margins(const cgrb_unit_t *__restrict__ units,
                                                     const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                     double *__restrict__ thin_tab)
{
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t n = u->tab_n;
    if (n < 4 || (u->flags & CGRB_UNIT_NO_SOURCE))
        return;
    const int64_t k = wk.start + threadIdx.x;      // interval [k, k+1]
    if (k >= n)
        return;
    const double *p = thin_tab + 2 * u->tab_off;
    // |second differences| at the nodes k-1 .. k+2 (clamped to the interior nodes 1 .. n-2)
    double m = 0.0;
    bool bad = false;
#pragma unroll
    for (int d = -1; d <= 2; d++) {
        int64_t c = min(max(k + d, (int64_t)1), n - 2);
        double d2 = fabs(p[2 * (c - 1)] - 2.0 * p[2 * c] + p[2 * (c + 1)]);
        bad |= !(d2 == d2) || isinf(d2);
        m = fmax(m, d2);
    }
    // linear interpolation error <= h^2/8 sup|p''| ~ d2/8; x4 safety, plus rounding slack
    double pk = fabs(p[2 * k]);
    double margin = 0.5 * m + 1e-12 * (pk + 1.0);
    thin_tab[2 * (u->tab_off + k) + 1] = bad ? INFINITY : margin;
}

// ==========================================================================================
// thinning pass 1: exponential gaps of a segment, their running sum INSIDE the segment (stored: the
// candidate's arrival time minus the segment's start time) and the segment total.  Pass 2 turns the
// totals into segment start times, pass 3 adds the start time -- no gap is drawn twice.
// The summation order is fixed (pairs, tiles of THIN_TILE, segments of CGRB_SEG), so times are reproducible.
// ==========================================================================================
template <int WHICH>  // 0 = source (prefix -> stage at src_off), 1 = background (prefix -> times_bkg at bkg_off + 1)
__global__ void __launch_bounds__(NT) k_gap_prefix(const cgrb_unit_t *__restrict__ units,
                                                   const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                   cgrb_rng_t rng, double *__restrict__ segsum,
                                                   double *__restrict__ prefix_out, cgrb_counts_t *counts)
{
    __shared__ double s_w[NW];
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t cap = (WHICH == 0) ? u->src_cap : u->bkg_cap;
    const double rate = (WHICH == 0) ? u->fmax : u->bkg_rate;
    const double inv_f = 1.0 / rate;
    const uint32_t stage = (WHICH == 0) ? ST_SRC_TIME : ST_BKG;
    const uint32_t sid = u->stream_id;
    const int64_t n = min((int64_t)CGRB_SEG, cap - wk.start);
    double *out = prefix_out + ((WHICH == 0) ? u->src_off : u->bkg_off + 1) + wk.start;
    double carry = 0.0;
    bool exhausted = false;
    for (int64_t base = 0; base < n; base += THIN_TILE) {
        const int64_t j0 = base + 2 * threadIdx.x;          // wk.start and base are even
        double g0 = 0.0, g1 = 0.0;
        if (j0 < n) {
            PairDraw d = pair_draw(rng, wk.unit, sid, stage, (wk.start + j0) >> 1, 0);
            g0 = -(inv_f * log(d.a));                        // time - (1/fmax) * log(u)  cpl_functions.py:160
            if (j0 + 1 < n)
                g1 = -(inv_f * log(d.b));
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
                                                  double *__restrict__ segstart, cgrb_counts_t *counts)
{
    const int lane = threadIdx.x & 31;
    const int ui = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ui >= n_units)
        return;
    const cgrb_unit_t *u = units + ui;
    const double tstart = (WHICH == 0) ? u->src_tstart : u->bkg_tstart;
    const double tstop = (WHICH == 0) ? u->src_tstop : u->bkg_tstop;
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

__device__ __forceinline__ bool thin_accept(const LambdaTab *tab, const cgrb_unit_t &u, bool squeeze,
                                            const double2 *__restrict__ sq, double sq_inv_h, double t, double u2,
                                            long long *n_exact)
{
    if (squeeze) {
        const double q = (t - u.src_tstart) * sq_inv_h;
        const int kq = min(max((int)q, 0), (int)u.tab_n - 2);       // tab_n < 2^31 (host: <= 65536 nodes)
        const double fr = q - (double)kq;
        const double2 n0 = __ldg(sq + kq), n1 = __ldg(sq + kq + 1);
        const double pl = n0.x + (n1.x - n0.x) * fr;
        if (u2 <= pl - n0.y)
            return true;
        if (u2 > pl + n0.y)
            return false;
    }
    (*n_exact)++;
    return u2 <= lambda_exact(tab, u, t) / u.fmax;   // `if test <= p_test`
}

__global__ void __launch_bounds__(NT, 3) k_source_thin(const double *__restrict__ dtab,
                                                    const cgrb_unit_t *__restrict__ units,
                                                    const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                    cgrb_rng_t rng, const double *__restrict__ segstart,
                                                    int32_t *__restrict__ segcount, double *stage,
                                                    double *__restrict__ cand_times,
                                                    uint8_t *__restrict__ cand_accept,
                                                    const double *__restrict__ thin_tab, cgrb_counts_t *counts)
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
    if ((u.flags & CGRB_UNIT_NO_SOURCE) || t0 > u.src_tstop) {   // dead segment
        if (threadIdx.x == 0)
            segcount[w] = 0;
        return;
    }
    build_lambda_tab(&tab, dtab + (size_t)u.det * CGRB_DT_SIZE, u);
    if (threadIdx.x == 0)
        s_last = t0;
    __syncthreads();

    const int64_t n = min((int64_t)CGRB_SEG, u.src_cap - wk.start);
    // squeeze table of p(t) = lambda(t)/fmax (cgrb_thin_tables): decides most candidates without
    // evaluating lambda; the rest take the exact test, so decisions are those of the exact test
    const bool squeeze = thin_tab != nullptr && u.tab_n >= 4;
    const double2 *sq = reinterpret_cast<const double2 *>(thin_tab) + (squeeze ? u.tab_off : 0);
    const double sq_inv_h = squeeze ? (double)(u.tab_n - 1) / (u.src_tstop - u.src_tstart) : 0.0;
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
        const bool va = in0 && (ta <= u.src_tstop);   // `if time > tstop: break`
        const bool vb = in1 && (tb <= u.src_tstop);
        bool aa = false, ab = false;
        if (va) {
            PairDraw d = pair_draw(rng, wk.unit, u.stream_id, ST_SRC_TIME, (wk.start + j0) >> 1, 1);
            aa = thin_accept(&tab, u, squeeze, sq, sq_inv_h, ta, d.a, &n_exact);
            if (vb)
                ab = thin_accept(&tab, u, squeeze, sq, sq_inv_h, tb, d.b, &n_exact);
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
            s_last = in1 ? tb : ta;
        __syncthreads();
        if (tile_valid < (int)min((int64_t)THIN_TILE, n - base))
            break;      // crossed tstop: everything later is beyond it
    }
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
// background (background.py:13-44 + numpy choice :93): every candidate is an event; event j of
// the unit is candidate j-1, event 0 is the spurious tstart event.  Pass 1 (k_gap_prefix<1>) left
// each candidate's running gap sum in its time slot; this pass adds the segment start, draws the
// channel from the template's cdf and cuts the list at tstop.
// ==========================================================================================
__global__ void __launch_bounds__(NT) k_background(const double *__restrict__ dtab,
                                                   const cgrb_unit_t *__restrict__ units,
                                                   const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                   cgrb_rng_t rng, cgrb_rng_t rng_chan,
                                                   const double *__restrict__ segstart,
                                                   double *times_bkg, uint8_t *__restrict__ pha_bkg,
                                                   cgrb_counts_t *counts)
{
    __shared__ double s_cdf[CGRB_NCHAN];
    __shared__ unsigned char s_cg[CGRB_NCHAN];      // guide: searchsorted(cdf, m/128, 'right')
    __shared__ int s_wi[NW];
    __shared__ double s_t[NT];
    __shared__ double s_last;
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    if (u->flags & CGRB_UNIT_NO_BACKGROUND)
        return;
    const double t0 = segstart[w];
    const double tstop = u->bkg_tstop;
    const int64_t off = u->bkg_off;
    const uint32_t sid = u->stream_id;
    const double *cdf_g = dtab + (size_t)u->det * CGRB_DT_SIZE + CGRB_DT_BKG_CDF;
    if (threadIdx.x < CGRB_NCHAN)
        s_cdf[threadIdx.x] = cdf_g[threadIdx.x];
    if (threadIdx.x == 0)
        s_last = t0;
    __syncthreads();
    if (threadIdx.x < CGRB_NCHAN)
        s_cg[threadIdx.x] = (unsigned char)min(searchsorted_right(s_cdf, CGRB_NCHAN, (double)threadIdx.x / CGRB_NCHAN),
                                               CGRB_NCHAN - 1);
    __syncthreads();
    // np.searchsorted(cdf, u, 'right') clamped to the last channel, started from the guide: ~2 probes
    auto channel_of = [&](double uc) -> uint8_t {
        int j = s_cg[min(max((int)(uc * (double)CGRB_NCHAN), 0), CGRB_NCHAN - 1)];
        while (j < CGRB_NCHAN - 1 && !(uc < s_cdf[j]))
            j++;
        return (uint8_t)j;
    };

    if (wk.start == 0 && threadIdx.x == 0) {
        // event 0: tstart, with its own channel draw (sample_channel(size=len(times)))
        double uc;
        if (rng.mode == 0) {
            Philox4 p = philox_at(make_key(rng.seed), sid, ST_BKG, 0xFFFFFFFFFFull, 1u);
            uc = u53(p.x, p.y);
        } else {
            int64_t o = rng_chan.d_off ? rng_chan.d_off[wk.unit] : 0;
            uc = rng_chan.d_u[o];
        }
        times_bkg[off] = u->bkg_tstart;
        pha_bkg[off] = channel_of(uc);
        atomicAdd((unsigned long long *)&counts[wk.unit].n_bkg, 1ull);
    }
    if (t0 > tstop)
        return;

    const int64_t n = min((int64_t)CGRB_SEG, u->bkg_cap - wk.start);
    double *seg_t = times_bkg + off + 1 + wk.start;
    uint8_t *seg_p = pha_bkg + off + 1 + wk.start;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int n_valid_total = 0;
    bool exhausted = false;
    for (int64_t base = 0; base < n; base += THIN_TILE) {
        const int64_t j0 = base + 2 * threadIdx.x;
        const bool in0 = j0 < n, in1 = j0 + 1 < n;
        double ta = in0 ? t0 + seg_t[j0] : INFINITY;
        double tb = in1 ? t0 + seg_t[j0 + 1] : INFINITY;
        s_t[threadIdx.x] = in1 ? tb : ta;
        __syncthreads();
        const double left = (threadIdx.x > 0) ? s_t[threadIdx.x - 1] : s_last;
        ta = fmax(ta, left);
        tb = fmax(tb, ta);
        const bool va = in0 && (ta <= tstop);
        const bool vb = in1 && (tb <= tstop);
        if (va) {
            double ua, ub = 0.5;
            if (rng.mode == 0) {
                PairDraw d = pair_draw(rng, wk.unit, sid, ST_BKG, (wk.start + j0) >> 1, 1);
                ua = d.a;
                ub = d.b;
            } else {
                const int64_t o = rng_chan.d_off ? rng_chan.d_off[wk.unit] : 0;
                const int64_t len = rng_chan.d_len ? rng_chan.d_len[wk.unit] : ((int64_t)1 << 62);
                const int64_t e0 = wk.start + j0 + 1;           // event index of candidate j0
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
            seg_p[j0] = channel_of(ua);
            if (vb) {
                seg_t[j0 + 1] = tb;
                seg_p[j0 + 1] = channel_of(ub);
            }
        }
        const unsigned bala = __ballot_sync(0xffffffffu, va), balb = __ballot_sync(0xffffffffu, vb);
        if (lane == 0)
            s_wi[wid] = __popc(bala) + __popc(balb);
        __syncthreads();
        int tile_valid = 0;
#pragma unroll
        for (int q = 0; q < NW; q++)
            tile_valid += s_wi[q];
        n_valid_total += tile_valid;
        if (threadIdx.x == NT - 1)
            s_last = in1 ? tb : ta;
        __syncthreads();
        if (tile_valid < (int)min((int64_t)THIN_TILE, n - base))
            break;
    }
    if (threadIdx.x == 0 && n_valid_total)
        atomicAdd((unsigned long long *)&counts[wk.unit].n_bkg, (unsigned long long)n_valid_total);
    if (exhausted)
        atomicOr(&counts[wk.unit].status, CGRB_ST_UNIFORMS_EXHAUSTED);
}

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

// lower bound (first index with xs[j] >= v) over n <= 255 ascending entries in a fixed 8 probes:
// no data-dependent trip count, so a warp never diverges on it
__device__ __forceinline__ int lower_bound_255(const double *xs, int n, double v)
{
    int pos = 0;
#pragma unroll
    for (int step = 128; step >= 1; step >>= 1) {
        const int q = pos + step;
        if (q <= n && xs[q - 1] < v)
            pos = q;
    }
    return pos;
}

// first j with row[j] >= r (0..128) for one row of the cumulative matrix.  `guide` (shared memory,
// CGRB_NCHAN bytes of this row) gives a start at or before it -- guide[m] = lower bound of m/128 -- so
// the forward walk is ~1-2 probes of the fp64 row; pathological rows (many channels inside one guide
// cell) finish by bisection.  *v_at = row[result] when result < 128.
__device__ __forceinline__ int guided_lower_bound(const uint8_t *guide, const double *__restrict__ row, double r,
                                                  double *v_at)
{
    const int m = min(max((int)(r * (double)CGRB_NCHAN), 0), CGRB_NCHAN - 1);
    int lo = guide[m];
    double v = 0.0;
    int steps = 0;
    while (lo < CGRB_NCHAN) {
        v = __ldg(row + lo);
        if (!(v < r))
            break;
        lo++;
        if (++steps == 6) {              // rare: finish by bisection on [lo, 128)
            int hi = CGRB_NCHAN;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(row + mid) < r)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (lo < CGRB_NCHAN)
                v = __ldg(row + lo);
            break;
        }
    }
    *v_at = v;
    return lo;
}

// argmin_j |row[j] - r| with the first index on ties (response.py:18-24)
__device__ __forceinline__ int nearest_cdf_guided(const uint8_t *guide, const double *__restrict__ row, double r)
{
    double vhi;
    const int lo = guided_lower_bound(guide, row, r, &vhi);
    if (lo == 0)
        return 0;
    const double vlo = __ldg(row + lo - 1);
    if (lo < CGRB_NCHAN && fabs(vhi - r) < fabs(vlo - r))
        return lo;                       // strictly nearer above; row[lo-1] < r <= row[lo]: no plateau below it
    // tie or nearer below -> lo - 1; np.argmin returns the FIRST index holding that value (plateaus of
    // zero-probability channels): the lower bound of the value itself
    if (lo >= 2 && !(__ldg(row + lo - 2) < vlo)) {
        double dummy;
        return guided_lower_bound(guide, row, vlo, &dummy);
    }
    return lo - 1;
}

// Response.digitize of one photon (response.py:10-25): photon bin by a full search of the edges
__device__ __forceinline__ int digitize_one(const double *edges, const uint8_t *guide, const double *__restrict__ cum,
                                            double x, uint64_t seed, uint32_t stream_id, int64_t i)
{
    int b = lower_bound_255(edges, CGRB_NBINS + 1, x) - 1;      // searchsorted 'left', response.py:14
    if (b < 0)
        b += CGRB_NBINS;                                          // python negative index
    b = min(b, CGRB_NBINS - 1);
    Philox4 p = philox_at(make_key(seed), stream_id, ST_SRC_CHAN, (uint64_t)i, 0u);
    return nearest_cdf_guided(guide + b * CGRB_NCHAN, cum + (size_t)b * CGRB_NCHAN, u53(p.x, p.y));
}

#define GUIDE_BYTES (CGRB_NBINS * CGRB_NCHAN)

// one mbarrier-tracked bulk-async (TMA) copy of a detector's 17,920-byte search guide into shared memory;
// called by one thread, every thread then waits with guide_wait()
__device__ __forceinline__ void guide_stage(uint8_t *s_guide, uint64_t *s_bar, const double *dt)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(s_bar);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(s_guide);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)GUIDE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<const char *>(dt + CGRB_DT_GUIDE)), "r"((unsigned)GUIDE_BYTES), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void guide_wait(uint64_t *s_bar)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(s_bar);
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(0u)
            : "memory");
    }
}

// exp(z) for |z| <= 0.1 to ~3e-17 relative: Taylor polynomial of degree 9
__device__ __forceinline__ double exp_small(double z)
{
    double p = 1.0 / 362880.0;
    p = fma(p, z, 1.0 / 40320.0);
    p = fma(p, z, 1.0 / 5040.0);
    p = fma(p, z, 1.0 / 720.0);
    p = fma(p, z, 1.0 / 120.0);
    p = fma(p, z, 1.0 / 24.0);
    p = fma(p, z, 1.0 / 6.0);
    p = fma(p, z, 0.5);
    p = fma(p, z, 1.0);
    return fma(p, z, 1.0);
}

// 1/d for d in the float range, from the single-precision reciprocal and two Newton steps (~1e-16)
__device__ __forceinline__ double rcp_fast(double d)
{
    double r = (double)__frcp_rn((float)d);
    r = r * (2.0 - d * r);
    return r * (2.0 - d * r);
}

#ifndef ENV_MIN_CTAS
#define ENV_MIN_CTAS 3
#endif
struct EnvSmem {
    double xlo[CGRB_ENV_SEGS];
    double invMx[CGRB_ENV_SEGS];
    double logAmax[CGRB_ENV_SEGS];
    int kn[CGRB_ENV_SEGS];
    double wk[CGRB_ENV_MAXW + 1];          // wk[n_w] = +inf
    double hull_w[CGRB_NE_ENVELOPE + 1];   // hull_w[n_hull] = +inf
    double hull_E[CGRB_NE_ENVELOPE];
    double hull_lA5[CGRB_NE_ENVELOPE];
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
    __shared__ long long s_wl[NW];
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
    for (int j = threadIdx.x; j <= n_hull; j += NT) {
        sm.hull_w[j] = (j < n_hull) ? T->hull_w[j] : INFINITY;
        if (j < n_hull) {
            sm.hull_E[j] = T->hull_E[j];
            sm.hull_lA5[j] = T->hull_lA5[j];
        }
    }
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
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t per = (chunk_end - wk.start + NW - 1) / NW;
    const int64_t wbeg = wk.start + (int64_t)wid * per, wend = min(wbeg + per, chunk_end);
    int64_t next = wbeg;                 // warp-uniform
    long long my_trials = 0;
    bool capped = false, active = false, violated = false;
    // per-lane photon state; k (table row) and h (hull position) persist from photon to photon: a lane's
    // photons come in time order, so both usually advance by zero or one step
    int64_t i = 0;
    uint32_t j = 0;
    int k = -1, h = -1;
    double wq = 0.0, s_idx = 0.0, t_clip = 0.0, total = 0.0;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !active);
        if (need) {
            const int64_t cand = next + __popc(need & ((1u << lane) - 1u));
            next += __popc(need);
            if (!active && cand < wend) {
                i = cand;
                j = 0;
                const double t = times_src[u.src_off + i];
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
                        while (steps < 4 && sm.hull_w[hh + 1] <= wh) {
                            hh++;
                            steps++;
                        }
                    if (hh < 0 || steps == 4 || sm.hull_w[hh] > wh) {
                        int lo = 0, hi = n_hull;                // last entry with hull_w <= wh
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (sm.hull_w[mid] <= wh)
                                lo = mid;
                            else
                                hi = mid;
                        }
                        hh = lo;
                    }
                    h = hh;
                    s_idx = sm.hull_E[h] * wq;
                    t_clip = sm.hull_lA5[h] - s_idx;
                    total = __ldg(unit_rows + (size_t)k * ENV_ROW + CGRB_ENV_SEGS);
                    if (!(total > 0.0) || isinf(total))
                        regular = false;
                }
                if (regular)
                    active = true;
                else {
                    // degenerate photon: the literal loop, right here (rare)
                    int64_t ntr = 0;
                    EaGuess eg = make_ea_guess(sm.eax);
                    const double x = literal_photon_serial(u, dt, sm.eax, sm.eay, eg, interp_extrap, rng.seed, i, t,
                                                           max_trials, &ntr, &capped);
                    if (DIAG && energy_src)
                        energy_src[u.src_off + i] = x;
                    if (FOLD)
                        pha_src[u.src_off + i] = (uint8_t)digitize_one(sm.edges, s_guide, cum, x, rng.seed, u.stream_id, i);
                    if (DIAG && trials_out)
                        trials_out[u.src_off + i] = (int32_t)ntr;
                    my_trials += ntr;
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (next >= wend)
                break;
            continue;
        }
        if (active) {
            const double *row = unit_rows + (size_t)k * ENV_ROW;
            Philox4 p = philox_at(make_key(rng.seed), u.stream_id, ST_SRC_ENERGY, (uint64_t)i, j);
            const double U = u53(p.x, p.y), V = u53(p.z, p.w);
            const double y = U * total;
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
            if (sm.logAmax[lo] - xl * wq > t_clip)
                A = fmin(A, exp(sm.hull_lA5[h] + s - s_idx));   // the literal envelope clips f here
            // target / envelope = [A x^a / M_lo] exp(-d), decided from the bounds
            // 1 - d + d^2/2 - d^3/6 <= exp(-d) <= 1 - d + d^2/2 whenever they agree.  The envelope of row k is
            // M_j exp(-xlo_j w_k) exp(-emin (w - w_k)) >= M_j exp(-xlo_j w): the common factor in w - w_k costs
            // nothing, so   d = (x - xlo) w + (xlo - emin) (w - w_k) >= 0.
            const double base = A * sm.invMx[lo] * exp_small(c_a * z);
            const double d = fmax((x - xl) * wq + (xl - sm.xlo[0]) * (wq - sm.wk[k]), 0.0);
            const double qhi = fma(d, fma(d, 0.5, -1.0), 1.0);
            const double qlo = qhi - d * d * d * (1.0 / 6.0);
            j++;
            violated |= base * qlo > 1.000001;      // the envelope must dominate the target (self-check)
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
                    energy_src[u.src_off + i] = acc ? x : nan("");
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
                        Philox4 pc = philox_at(make_key(rng.seed), u.stream_id, ST_SRC_CHAN, (uint64_t)i, 0u);
                        pha = nearest_cdf_guided(s_guide + b * CGRB_NCHAN, cum + (size_t)b * CGRB_NCHAN, u53(pc.x, pc.y));
                    } else
                        pha = digitize_one(sm.edges, s_guide, cum, nan(""), rng.seed, u.stream_id, i);
                    pha_src[u.src_off + i] = (uint8_t)pha;
                }
                if (DIAG && trials_out)
                    trials_out[u.src_off + i] = (int32_t)j;
                my_trials += j;
                capped |= cap;
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
    const long long tot = block_sum_ll(my_trials, s_wl);
    if (threadIdx.x == 0 && tot)
        atomicAdd((unsigned long long *)&counts[wk.unit].n_trials, (unsigned long long)tot);
    if (capped)
        atomicOr(&counts[wk.unit].status, CGRB_ST_TRIAL_CAP);
    if (violated)
        atomicOr(&counts[wk.unit].status, CGRB_ST_ENVELOPE);
}

// ==========================================================================================
// digitize (response.py:10-25): photon bin by searchsorted on the edges, then the NEAREST value
// of the row-cumulative matrix (argmin |cum - r|, first index on ties), found by binary search.
// The detector's 140x128 cumulative matrix (143,360 B) is staged into shared memory with one
// mbarrier-tracked bulk-async (TMA) copy per CTA when the chunk is large enough to amortise it.
// ==========================================================================================
__global__ void __launch_bounds__(NT) k_digitize(const double *__restrict__ dtab,
                                                 const cgrb_unit_t *__restrict__ units,
                                                 const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                 cgrb_rng_t rng, const double *__restrict__ energy_src,
                                                 uint8_t *__restrict__ pha_src, const cgrb_counts_t *counts)
{
    __shared__ __align__(128) uint8_t s_guide[GUIDE_BYTES];
    __shared__ double s_edges[CGRB_NBINS + 1];
    __shared__ __align__(8) uint64_t s_bar;

    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const int64_t n_src = counts[wk.unit].n_src;
    if (wk.start >= n_src)
        return;
    int64_t chunk_end = n_src;
    if (w + 1 < n_work && work[w + 1].unit == wk.unit)
        chunk_end = min(n_src, work[w + 1].start);
    const cgrb_unit_t *u = units + wk.unit;
    const double *dt = dtab + (size_t)u->det * CGRB_DT_SIZE;
    if (threadIdx.x == 0)
        guide_stage(s_guide, &s_bar, dt);
    for (int m = threadIdx.x; m <= CGRB_NBINS; m += NT)
        s_edges[m] = dt[CGRB_DT_EDGES + m];
    __syncthreads();
    guide_wait(&s_bar);
    const double *cum = dt + CGRB_DT_CUM;
    const uint32_t sid = u->stream_id;
    const int64_t off = u->src_off;

    for (int64_t i = wk.start + threadIdx.x; i < chunk_end; i += NT) {
        const double e = energy_src[off + i];
        if (rng.mode == 0)
            pha_src[off + i] = (uint8_t)digitize_one(s_edges, s_guide, cum, e, rng.seed, sid, i);
        else {
            int idx = lower_bound_255(s_edges, CGRB_NBINS + 1, e) - 1;        // searchsorted 'left', response.py:14
            if (idx < 0)
                idx += CGRB_NBINS;                                            // python negative index
            idx = min(idx, CGRB_NBINS - 1);
            const int64_t o = rng.d_off ? rng.d_off[wk.unit] : 0;
            pha_src[off + i] = (uint8_t)nearest_cdf_guided(s_guide + idx * CGRB_NCHAN, cum + (size_t)idx * CGRB_NCHAN,
                                                           rng.d_u[o + i]);
        }
    }
}

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

__global__ void __launch_bounds__(NT) k_deadtime_count(const cgrb_unit_t *__restrict__ units,
                                                       const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                       const double *__restrict__ times_tot,
                                                       const uint8_t *__restrict__ pha_tot,
                                                       int32_t *__restrict__ tile_count,
                                                       const cgrb_counts_t *counts)
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
            cnt += dt_keep(t, p, i) ? 1 : 0;
    }
    int total;
    (void)block_incl_scan_i(cnt, s_wi, &total);
    if (threadIdx.x == 0)
        tile_count[w] = total;
}

__global__ void __launch_bounds__(NT) k_deadtime_write(const cgrb_unit_t *__restrict__ units,
                                                       const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                       const int64_t *__restrict__ unit_work_begin,
                                                       const double *__restrict__ times_tot,
                                                       const uint8_t *__restrict__ pha_tot,
                                                       const int32_t *__restrict__ tile_count,
                                                       double *__restrict__ times_out, uint8_t *__restrict__ pha_out,
                                                       uint8_t *__restrict__ selection, cgrb_counts_t *counts)
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
            keep = dt_keep(t, p, i);
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
#define MD_TILE 2048
#define MD_HALO 256
#define MD_N (MD_TILE + MD_HALO)
#define MD_FLAG_AGG (1ull << 62)
#define MD_FLAG_INC (2ull << 62)
#define MD_VALUE_MASK ((1ull << 62) - 1ull)

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

// the literal dead-time rule for merged event g (same walk as dt_keep) on the slow accessor
__device__ __noinline__ bool dt_keep_slow(const MergedGlobal &m, int64_t i)
{
    if (i == 0)
        return true;
    double ti, tj;
    uint8_t pi, pj;
    m.get(i, &ti, &pi);
    int64_t found = -1;
    for (int64_t j = i - 1; j >= 0; j--) {
        m.get(j, &tj, &pj);
        if (tj + DT_TAU < ti)
            break;
        if (pj == 127 || j == 0) {
            found = j;
            break;
        }
    }
    if (found < 0)
        return true;
    int64_t head = found;
    while (head > 0) {
        double th;
        uint8_t ph;
        m.get(head, &th, &ph);
        int64_t q = -1;
        double tq = 0.0;
        uint8_t pq = 0;
        for (int64_t j = head - 1; j >= 0; j--) {
            m.get(j, &tq, &pq);
            if (tq + DT_TAU < th)
                break;
            if (pq == 127 || j == 0) {
                q = j;
                break;
            }
        }
        if (q < 0 || tq + ((q == 0 && pq != 127) ? DT_NORMAL : DT_TAU) < th)
            break;
        head = q;
    }
    double th;
    uint8_t ph;
    m.get(head, &th, &ph);
    double t_end = th + ((head == 0 && ph != 127) ? DT_NORMAL : DT_TAU);
    for (int64_t j = head + 1; j < i; j++) {
        m.get(j, &tj, &pj);
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

// merge-path splits of every tile, one thread per tile: the ~23 dependent probes of each search are
// hidden by the other tiles' searches instead of stalling a whole CTA at its first barrier.
// splits[2w] = split at the tile's halo start, splits[2w+1] = split at the tile start.
__global__ void __launch_bounds__(NT) k_merge_splits(const cgrb_unit_t *__restrict__ units,
                                                     const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                     const double *__restrict__ times_bkg,
                                                     const double *__restrict__ times_src,
                                                     int64_t *__restrict__ splits, const cgrb_counts_t *counts)
{
    const int64_t w = (int64_t)blockIdx.x * NT + threadIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t nA = counts[wk.unit].n_bkg, nB = counts[wk.unit].n_src;
    if (wk.start >= nA + nB)
        return;
    const double *A = times_bkg + u->bkg_off;
    const double *B = times_src + u->src_off;
    splits[2 * w] = merge_split(A, nA, B, nB, max((int64_t)0, wk.start - MD_HALO));
    splits[2 * w + 1] = merge_split(A, nA, B, nB, wk.start);
}

#define MD_PER (MD_N / NT)       // merged elements per thread in the serial merge (9)
#define MD_R (MD_TILE / NT)     // tile elements per thread in the dead-time pass (8)

__global__ void __launch_bounds__(NT) k_merge_deadtime(const cgrb_unit_t *__restrict__ units,
                                                       const cgrb_work_t *__restrict__ work, int64_t n_work,
                                                       const int64_t *__restrict__ unit_work_begin,
                                                       const double *__restrict__ times_bkg,
                                                       const uint8_t *__restrict__ pha_bkg,
                                                       const double *__restrict__ times_src,
                                                       const uint8_t *__restrict__ pha_src,
                                                       unsigned long long *tile_state,
                                                       const int64_t *__restrict__ splits,
                                                       double *__restrict__ times_out, uint8_t *__restrict__ pha_out,
                                                       cgrb_counts_t *counts)
{
    __shared__ double s_in_t[MD_N];         // the two input slices, background first
    __shared__ double s_t[MD_N];            // merged
    __shared__ uint8_t s_in_p[MD_N];
    __shared__ uint8_t s_p[MD_N];
    __shared__ int s_cnt[MD_R * NW];        // survivors per (round, warp), then their exclusive prefix
    __shared__ long long s_excl;
    const int64_t w = blockIdx.x;
    if (w >= n_work)
        return;
    const cgrb_work_t wk = work[w];
    const cgrb_unit_t *u = units + wk.unit;
    const int64_t nA = counts[wk.unit].n_bkg, nB = counts[wk.unit].n_src;
    const int64_t total = nA + nB;
    const int64_t wb = unit_work_begin[wk.unit], we = unit_work_begin[wk.unit + 1];
    const double *A = times_bkg + u->bkg_off;
    const double *B = times_src + u->src_off;
    const uint8_t *pA = pha_bkg + u->bkg_off;
    const uint8_t *pB = pha_src + u->src_off;
    const int64_t d0 = wk.start, d1 = min(d0 + (int64_t)MD_TILE, total);
    const int64_t g0 = max((int64_t)0, d0 - MD_HALO);       // first merged rank held in shared memory
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool live = d0 < total;
    if (!live) {
        // a tile beyond the end of the merged list: nothing to merge, and no live tile follows it, so nobody
        // will look back at it (the last LIVE tile writes n_kept)
        if (w == wb && threadIdx.x == 0) {
            counts[wk.unit].n_total = total;
            counts[wk.unit].n_kept = 0;
        }
        return;
    }
    const int l0 = (int)(d0 - g0), nt = (int)(d1 - d0);

    unsigned keep_bits = 0;                 // bit r: tile element r * NT + tid survives
    if (live) {
        // d1 is either the next tile's start (same unit) or the end of the merged list
        const int64_t i0 = splits[2 * w], i1 = (d1 == total) ? nA : splits[2 * (w + 1) + 1];
        const int64_t j0 = g0 - i0, j1 = d1 - i1;
        const int na = (int)(i1 - i0), nb = (int)(j1 - j0);
        for (int k = threadIdx.x; k < na; k += NT) {
            s_in_t[k] = A[i0 + k];
            s_in_p[k] = pA[i0 + k];
        }
        for (int k = threadIdx.x; k < nb; k += NT) {
            s_in_t[na + k] = B[j0 + k];
            s_in_p[na + k] = pB[j0 + k];
        }
        __syncthreads();
        // serial merge of MD_PER consecutive outputs per thread from its merge-path split
        {
            const double *a_ = s_in_t, *b_ = s_in_t + na;
            const int n = na + nb;
            const int q0 = min(threadIdx.x * MD_PER, n);
            int lo = max(0, q0 - nb), hi = min(q0, na);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (a_[mid] <= b_[q0 - 1 - mid])     // ties: background first
                    lo = mid + 1;
                else
                    hi = mid;
            }
            int ia = lo, ib = q0 - lo;
#pragma unroll
            for (int r = 0; r < MD_PER; r++) {
                const int q = q0 + r;
                if (q < n) {
                    const bool take_a = ia < na && (ib >= nb || a_[ia] <= b_[ib]);
                    const int src = take_a ? ia : na + ib;
                    s_t[q] = s_in_t[src];
                    s_p[q] = s_in_p[src];
                    ia += take_a ? 1 : 0;
                    ib += take_a ? 0 : 1;
                }
            }
        }
        __syncthreads();
        // keep flags of the tile's own events (local indices [l0, l0 + nt))
        MergedGlobal mg{A, B, pA, pB, nA, nB};
#pragma unroll
        for (int r = 0; r < MD_R; r++) {
            const int e = r * NT + threadIdx.x;
            int keep = 0;
            if (e < nt) {
                keep = dt_keep_local(s_t, s_p, l0 + e, g0);
                if (keep == 2)
                    keep = dt_keep_slow(mg, d0 + e) ? 1 : 0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0)
                s_cnt[r * NW + wid] = __popc(bal);
            // rank inside the (round, warp) group, kept in the upper bits
            keep_bits |= (unsigned)keep << r;
        }
    } else if (threadIdx.x < MD_R * NW)
        s_cnt[threadIdx.x] = 0;
    __syncthreads();

    // warp 0: exclusive prefix of the 64 group counts, then the decoupled look-back over the unit's
    // earlier tiles for this tile's offset in the unit's output
    if (wid == 0) {
        int c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
        int incl = warp_incl_scan_i(c0 + c1, lane);
        const int cnt_tile = __shfl_sync(0xffffffffu, incl, 31);
        s_cnt[2 * lane] = incl - c0 - c1;
        s_cnt[2 * lane + 1] = incl - c1;
        if (lane == 0) {
            const unsigned long long mine = (unsigned long long)cnt_tile | (w == wb ? MD_FLAG_INC : MD_FLAG_AGG);
            atomicExch(tile_state + w, mine);
        }
        long long excl = 0;
        if (w > wb) {
            // 256 predecessors per round (8 loads in flight per lane): hundreds of tiles run concurrently, so
            // the nearest tile with an inclusive prefix is typically that far back
            int64_t idx = w - 1;
            bool done = false;
            while (!done) {
                unsigned long long st[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int64_t my = idx - 32 * q - lane;
                    st[q] = (my >= wb) ? *((volatile unsigned long long *)(tile_state + my)) : MD_FLAG_INC;
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (done)
                        break;
                    const int64_t my = idx - 32 * q - lane;
                    while ((st[q] >> 62) == 0ull)
                        st[q] = *((volatile unsigned long long *)(tile_state + my));
                    const unsigned inc = __ballot_sync(0xffffffffu, (st[q] >> 62) == 2ull);
                    const int first_inc = inc ? (__ffs(inc) - 1) : 32;     // nearest tile with an inclusive prefix
                    long long v = (lane <= first_inc && my >= wb) ? (long long)(st[q] & MD_VALUE_MASK) : 0;
#pragma unroll
                    for (int dlt = 16; dlt > 0; dlt >>= 1)
                        v += __shfl_xor_sync(0xffffffffu, v, dlt);
                    excl += v;
                    if (inc)
                        done = true;
                }
                idx -= 256;
            }
            if (lane == 0)
                atomicExch(tile_state + w, (unsigned long long)(excl + cnt_tile) | MD_FLAG_INC);
        }
        if (lane == 0) {
            s_excl = excl;
            if (w == wb)
                counts[wk.unit].n_total = total;
            if (d1 == total)
                counts[wk.unit].n_kept = excl + cnt_tile;     // the last live tile of the unit
        }
    }
    __syncthreads();
    if (!live)
        return;
    double *ot = times_out + u->tot_off + s_excl;
    uint8_t *op = pha_out + u->tot_off + s_excl;
#pragma unroll
    for (int r = 0; r < MD_R; r++) {
        // survivors of (round r, this warp) before this lane
        unsigned bal = __ballot_sync(0xffffffffu, (keep_bits >> r) & 1u);
        if ((keep_bits >> r) & 1u) {
            const int rank = s_cnt[r * NW + wid] + __popc(bal & ((1u << lane) - 1u));
            const int li = l0 + r * NT + threadIdx.x;
            ot[rank] = s_t[li];
            op[rank] = s_p[li];
        }
    }
}

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
    __shared__ double s_w[NW];
    __shared__ long long s_l[NW];
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
    // prefix sums: counts after the reference's float round trip, exposures = (stop - start) - dead time
    long long carry_c[4] = {0, 0, 0, 0};
    double carry_e = 0.0;
    if (threadIdx.x == 0) {
        for (int r = 0; r < 4; r++)
            cs[r * (bins_per_unit + 1)] = 0;
        es[0] = 0.0;
    }
    for (int64_t base = 0; base < nb; base += NT) {
        const int64_t k = base + threadIdx.x;
        const bool in = k < nb;
        long long c[4];
        double e = 0.0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int raw = in ? ub[r * bins_per_unit + k] : 0;
            c[r] = (long long)(__dmul_rn(__ddiv_rn((double)raw, dt), dt));      // (rate * dt).astype(int)
        }
        if (in) {
            const int na = ub[4 * bins_per_unit + k], no = ub[5 * bins_per_unit + k];
            const double dead = __dadd_rn(__dmul_rn((double)(na - no), 2.0e-6), __dmul_rn((double)no, 10.0e-6));
            e = __dadd_rn(__dadd_rn(g.edge(k + 1), -g.edge(k)), -dead);
        }
        // block inclusive scans (fixed association)
        double tot_e;
        const double incl_e = block_incl_scan(e, s_w, &tot_e);
        if (in)
            es[k + 1] = carry_e + incl_e;
        carry_e += tot_e;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            long long v = c[r];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d)
                    v += o;
            }
            if (lane == 31)
                s_l[wid] = v;
            __syncthreads();
            long long before = 0, tot = 0;
#pragma unroll
            for (int q = 0; q < NW; q++) {
                if (q < wid)
                    before += s_l[q];
                tot += s_l[q];
            }
            if (in)
                cs[r * (bins_per_unit + 1) + k + 1] = carry_c[r] + before + v;
            carry_c[r] += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    // the search: energy ranges a..d, time scales longest first; the first combination (in that order) with
    // a window at or above the threshold wins, at its first window.  The prefix sums of a chunk of windows
    // (+ the longest window span) are staged in shared memory; all time scales of a range share the
    // background window of a given i.
    __shared__ double s_e[TRIG_CH + TRIG_SPAN + 1];
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
            for (int64_t k = c0 + threadIdx.x; k <= hi; k += NT) {
                s_e[k - c0] = es[k];
                s_c[k - c0] = (int)(cr[k] - cbase);
            }
            __syncthreads();
            for (int64_t i = c0 + threadIdx.x; i < min(c0 + (int64_t)TRIG_CH, n_i_max); i += NT) {
                const int li = (int)(i - c0);
                const double eb = s_e[li + n_bkg] - s_e[li];
                const double B = (double)(s_c[li + n_bkg] - s_c[li]);
                for (int sidx = 0; sidx < ns; sidx++) {
                    const int64_t n_gap = (int64_t)1 << (ns - 1 - sidx);
                    if (i >= nb - (n_bkg + n_gap + n_src))                // i + span < len(stops)
                        continue;
                    const int ls = li + (int)(n_bkg + n_gap);
                    const double ex = s_e[ls + n_src] - s_e[ls];
                    const double S = (double)(s_c[ls + n_src] - s_c[ls]);
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

int cgrb_thin_tables(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work,
                     int64_t n_work, double *d_thin_tab, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_thin_tab)
        return fail(CGRB_E_BADARG, "cgrb_thin_tables: bad argument%s");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH("k_thin_nodes", s, k_thin_nodes<<<(unsigned)n_work, NT, 0, s>>>(d_dtab, d_units, d_work, n_work, d_thin_tab));
    LAUNCH("k_thin_margins", s, k_thin_margins<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, d_thin_tab));
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
                       uint8_t *d_cand_accept, const double *d_thin_tab, cgrb_counts_t *d_counts, void *stream)
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
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH("k_gap_prefix<0>", s, k_gap_prefix<0><<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, *rng, d_segsum, d_stage, d_counts));
    LAUNCH("k_seg_scan<0>", s, k_seg_scan<0><<<(n_units * 32 + 127) / 128, 128, 0, s>>>(d_units, n_units, d_unit_work_begin, d_segsum,
                                                           d_segstart, d_counts));
    LAUNCH("k_source_thin", s, k_source_thin<<<(unsigned)n_work, NT, 0, s>>>(d_dtab, d_units, d_work, n_work, *rng, d_segstart,
                                                  d_segcount, d_stage, d_cand_times, d_cand_accept, d_thin_tab, d_counts));
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

int cgrb_sample_energy_envelope(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                                const cgrb_work_t *d_work, int64_t n_work, int interp_extrap,
                                const cgrb_rng_t *rng, int64_t max_trials, void *d_workspace, int64_t n_rows,
                                const double *d_times_src, double *d_energy_src, uint8_t *d_pha_src,
                                int32_t *d_trials, cgrb_counts_t *d_counts, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_times_src ||
        (!d_energy_src && !d_pha_src) || !d_counts || !d_workspace || n_rows < n_units)
        return fail(CGRB_E_BADARG, "cgrb_sample_energy_envelope: bad argument%s");
    if (interp_extrap != 0 && interp_extrap != 1)
        return fail(CGRB_E_BADARG, "cgrb_sample_energy_envelope: interp_extrap must be 0 (linear) or 1 (clamp)%s");
    int rc = check_rng(rng, "cgrb_sample_energy_envelope");
    if (rc)
        return rc;
    if (rng->mode != 0)
        return fail(CGRB_E_UNSUPPORTED,
                    "cgrb_sample_energy_envelope: injected uniforms replay the literal sampler; use cgrb_sample_energy%s");
    cudaStream_t s = (cudaStream_t)stream;
    EnergyUnitTab *etab = (EnergyUnitTab *)d_workspace;
    int64_t tabs = ((int64_t)sizeof(EnergyUnitTab) * n_units + 255) / 256 * 256;
    double *rows = (double *)((char *)d_workspace + tabs);
    uint8_t *row_guides = (uint8_t *)(rows + (size_t)ENV_ROW * n_rows);
    LAUNCH("k_energy_tables", s, k_energy_tables<<<n_units, 256, 0, s>>>(d_dtab, d_units, n_units, interp_extrap, etab));
    LAUNCH("k_energy_rows", s, k_energy_rows<<<(unsigned)n_rows, NT, 0, s>>>(d_units, n_units, etab, rows, row_guides, n_rows));
#define ENV_LAUNCH(F, D, M)                                                                                       \
    LAUNCH("k_sample_energy_env", s, (k_sample_energy_env<F, D, M><<<(unsigned)n_work, NT, 0, s>>>(                \
        d_dtab, d_units, d_work, n_work, interp_extrap, *rng, max_trials, etab, rows, row_guides, d_times_src,    \
        d_energy_src, d_pha_src, d_trials, d_counts)))
    const bool diag = d_energy_src != nullptr || d_trials != nullptr;
    // resident CTAs per SM the production variant is compiled for (register budget); CGRB_ENV_MIN_CTAS=2|3
    // switches it for tuning runs
    static int minb = 0;
    if (!minb) {
        const char *e = getenv("CGRB_ENV_MIN_CTAS");
        minb = (e && e[0] == '2') ? 2 : (e && e[0] == '4') ? 4 : ENV_MIN_CTAS;
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
    return check_launch("cgrb_sample_energy_envelope");
}

int cgrb_digitize(const double *d_dtab, const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work,
                  int64_t n_work, const cgrb_rng_t *rng, const double *d_energy_src, uint8_t *d_pha_src,
                  const cgrb_counts_t *d_counts, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_energy_src || !d_pha_src ||
        !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_digitize: bad argument%s");
    int rc = check_rng(rng, "cgrb_digitize");
    if (rc)
        return rc;
    LAUNCH("k_digitize", (cudaStream_t)stream, k_digitize<<<(unsigned)n_work, NT, 0, (cudaStream_t)stream>>>(d_dtab, d_units, d_work, n_work, *rng,
                                                                     d_energy_src, d_pha_src, d_counts));
    return check_launch("cgrb_digitize");
}

int cgrb_sample_background(const double *d_dtab, const cgrb_unit_t *d_units, int n_units,
                           const cgrb_work_t *d_work, int64_t n_work, const int64_t *d_unit_work_begin,
                           const cgrb_rng_t *rng_time, const cgrb_rng_t *rng_chan, double *d_segsum,
                           double *d_segstart, double *d_times_bkg, uint8_t *d_pha_bkg,
                           cgrb_counts_t *d_counts, void *stream)
{
    if (!d_dtab || !d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_unit_work_begin || !d_segsum ||
        !d_segstart || !d_times_bkg || !d_pha_bkg || !d_counts)
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
    LAUNCH("k_gap_prefix<1>", s, k_gap_prefix<1><<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, *rng_time, d_segsum, d_times_bkg, d_counts));
    LAUNCH("k_seg_scan<1>", s, k_seg_scan<1><<<(n_units * 32 + 127) / 128, 128, 0, s>>>(d_units, n_units, d_unit_work_begin, d_segsum,
                                                           d_segstart, d_counts));
    LAUNCH("k_background", s, k_background<<<(unsigned)n_work, NT, 0, s>>>(d_dtab, d_units, d_work, n_work, *rng_time, chan, d_segstart,
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

int cgrb_gbm_dead_time(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                       const int64_t *d_unit_work_begin, const double *d_times_tot, const uint8_t *d_pha_tot,
                       int32_t *d_tile_count, double *d_times_out, uint8_t *d_pha_out, uint8_t *d_selection,
                       cgrb_counts_t *d_counts, void *stream)
{
    if (!d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_unit_work_begin || !d_times_tot ||
        !d_pha_tot || !d_tile_count || !d_times_out || !d_pha_out || !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_gbm_dead_time: bad argument%s");
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH("k_deadtime_count", s, k_deadtime_count<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, d_times_tot, d_pha_tot,
                                                     d_tile_count, d_counts));
    LAUNCH("k_deadtime_write", s, k_deadtime_write<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, d_unit_work_begin, d_times_tot,
                                                     d_pha_tot, d_tile_count, d_times_out, d_pha_out,
                                                     d_selection, d_counts));
    return check_launch("cgrb_gbm_dead_time");
}

int64_t cgrb_merge_dead_time_workspace_bytes(int64_t n_work) { return 24 * (n_work > 0 ? n_work : 0); }

int cgrb_merge_dead_time(const cgrb_unit_t *d_units, int n_units, const cgrb_work_t *d_work, int64_t n_work,
                         const int64_t *d_unit_work_begin, const double *d_times_bkg, const uint8_t *d_pha_bkg,
                         const double *d_times_src, const uint8_t *d_pha_src, void *d_workspace,
                         double *d_times_out, uint8_t *d_pha_out, cgrb_counts_t *d_counts, void *stream)
{
    if (!d_units || n_units <= 0 || !d_work || n_work <= 0 || !d_unit_work_begin || !d_times_bkg || !d_pha_bkg ||
        !d_times_src || !d_pha_src || !d_workspace || !d_times_out || !d_pha_out || !d_counts)
        return fail(CGRB_E_BADARG, "cgrb_merge_dead_time: bad argument%s");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_workspace, 0, 8 * (size_t)n_work, s);
    if (e != cudaSuccess)
        return fail(CGRB_E_CUDA, "cgrb_merge_dead_time: memset: %s", cudaGetErrorString(e));
    int64_t *splits = (int64_t *)d_workspace + n_work;
    LAUNCH("k_merge_splits", s, k_merge_splits<<<(unsigned)((n_work + NT - 1) / NT), NT, 0, s>>>(
        d_units, d_work, n_work, d_times_bkg, d_times_src, splits, d_counts));
    LAUNCH("k_merge_deadtime", s, k_merge_deadtime<<<(unsigned)n_work, NT, 0, s>>>(
        d_units, d_work, n_work, d_unit_work_begin, d_times_bkg, d_pha_bkg, d_times_src, d_pha_src,
        (unsigned long long *)d_workspace, splits, d_times_out, d_pha_out, d_counts));
    return check_launch("cgrb_merge_dead_time");
}

int64_t cgrb_gbm_trigger_workspace_bytes(int n_units, int64_t bins_per_unit)
{
    if (n_units <= 0 || bins_per_unit <= 0)
        return 0;
    const int64_t nu = n_units;
    return nu * TRIG_NARR * bins_per_unit * 4 + nu * 4 * (bins_per_unit + 1) * 8 + nu * (bins_per_unit + 1) * 8 + 256;
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
    int64_t *csum = (int64_t *)((char *)d_workspace + ((bins_bytes + 255) / 256) * 256);
    double *esum = (double *)(csum + nu * 4 * (bins_per_unit + 1));
    cudaError_t e = cudaMemsetAsync(bins, 0, bins_bytes, s);
    if (e != cudaSuccess)
        return fail(CGRB_E_CUDA, "cgrb_gbm_trigger: memset: %s", cudaGetErrorString(e));
    LAUNCH("k_trig_bin", s, k_trig_bin<<<(unsigned)n_work, NT, 0, s>>>(d_units, d_work, n_work, d_chan_bounds,
                                                                      par->base_timescale, d_times_out, d_pha_out,
                                                                      d_counts, bins_per_unit, bins));
    LAUNCH("k_trig_eval", s, k_trig_eval<<<n_units, NT, 0, s>>>(d_units, n_units, *par, d_times_out, d_counts,
                                                               bins_per_unit, bins, csum, esum, d_out));
    return check_launch("cgrb_gbm_trigger");
}

}  // extern "C"
