// cgrb_common.cuh -- block scans, thinning uniforms, the lambda(t) table shared by the kernels
// Part of the single translation unit cgrb_kernels.cu (included in order; see that file).
#pragma once

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

__device__ __forceinline__ double lambda_of(const LambdaTab *tab, const cgrb_unit_t &u, const SpecT s)
{
    double acc = 0.0;
#pragma unroll 5
    for (int m = 0; m < CGRB_NE_LAMBDA; m++)
        acc += tab->c[m] * exp(-tab->x[m] * s.inv_ec);
    return s.norm * exp(-u.alpha * s.log_ec) * acc;
}
__device__ __forceinline__ double lambda_at(const LambdaTab *tab, const cgrb_unit_t &u, double t)
{
    return lambda_of(tab, u, spec_at(u, t));
}
// counts/s per unit of energy flux at time t: lambda(t) / F(t), a smooth function of t through ep(t) only
// (time-evolving CPL; no pulse factor, hence no singularity at the pulse start)
__device__ __forceinline__ double lambda_per_flux_at(const LambdaTab *tab, const cgrb_unit_t &u, double t)
{
    return lambda_of(tab, u, spec_with(u, 1.0, u.ep_start / (1.0 + t / u.ep_tau)));
}


// ------------------------------------------------------------------------------------------
// Decoupled look-back (one warp): exclusive prefix of integer counts over the work items [wb, w) of a
// unit.  Every item publishes (flag, count) in one 64-bit word: AGG = its own count, INC = the inclusive
// prefix up to and including it; the first item of a unit publishes INC at once.  A warp looks at 256
// predecessors per round (8 loads in flight per lane): hundreds of CTAs run concurrently, so the nearest
// inclusive prefix is typically that far back.  Integer sums: the result does not depend on timing.
// The state array must be zeroed before the launch.  Call with all 32 lanes of one warp.
//
// Forward progress: an item only ever waits for items with a SMALLER launch index of the same 1-D grid
// (k_merge_deadtime: blockIdx.x is the tile's ticket in the level-major order, k_level_order), and CTAs of a 1-D
// grid are dispatched in blockIdx order, so every predecessor is running or finished -- the assumption CUB's
// decoupled look-back scan makes too.  (An atomic ticket counter claimed by persistent CTAs was built and measured:
// CTAs holding claimed tiles behind the one they process form a convoy)  The assumption is not part of
// the CUDA programming model, so no wait is unbounded: a warp only waits for the lanes at or before the nearest
// inclusive prefix, and a wait that outlasts LB_WAIT_NS (20 s) gives up and reports it through *timed_out
// (CGRB_ST_INTERNAL) instead of hanging the device.
// ------------------------------------------------------------------------------------------
#define LB_FLAG_AGG (1ull << 62)
#define LB_FLAG_INC (2ull << 62)
#define LB_VALUE_MASK ((1ull << 62) - 1ull)
#define LB_WAIT_NS 20000000000ull       // 20 s
#ifndef LB_SLEEP_NS
#define LB_SLEEP_NS 0
#endif
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ long long lookback_exclusive(unsigned long long *state, int64_t w, int64_t wb,
                                                        long long count, int lane, bool *timed_out)
{
    if (lane == 0)
        atomicExch(state + w, (unsigned long long)count | (w == wb ? LB_FLAG_INC : LB_FLAG_AGG));
    long long excl = 0;
    if (w > wb) {
        int64_t idx = w - 1;
        bool done = false;
        while (!done) {
            unsigned long long st[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int64_t my = idx - 32 * q - lane;
                st[q] = (my >= wb) ? *((volatile unsigned long long *)(state + my)) : LB_FLAG_INC;
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (done)
                    break;
                const int64_t my = idx - 32 * q - lane;
                unsigned inc;
                unsigned long long t_begin = 0;
                for (int spin = 0;; spin++) {
                    inc = __ballot_sync(0xffffffffu, (st[q] >> 62) == 2ull);
                    const unsigned pend = __ballot_sync(0xffffffffu, (st[q] >> 62) == 0ull);
                    // lanes beyond the nearest inclusive prefix do not matter (lane 0 is the nearest predecessor)
                    const unsigned upto = inc ? ((2u << (__ffs(inc) - 1)) - 1u) : 0xffffffffu;
                    if (!(pend & upto))
                        break;
                    if ((spin & 1023) == 1023) {
                        // (a predecessor in the global slow path may legitimately take seconds)
                        const unsigned long long now = __shfl_sync(0xffffffffu, global_timer_ns(), 0);
                        if (t_begin == 0)
                            t_begin = now;
                        else if (now - t_begin > LB_WAIT_NS) {
                            *timed_out = true;
                            break;
                        }
                    }
#if LB_SLEEP_NS > 0
                    __nanosleep(LB_SLEEP_NS);   // leave the issue slots to the warps that have work
#endif
                    if ((st[q] >> 62) == 0ull)
                        st[q] = *((volatile unsigned long long *)(state + my));
                }
                const int first_inc = inc ? (__ffs(inc) - 1) : 32;      // nearest item with an inclusive prefix
                long long v = (lane <= first_inc && my >= wb && (st[q] >> 62) != 0ull) ? (long long)(st[q] & LB_VALUE_MASK) : 0;
#pragma unroll
                for (int dlt = 16; dlt > 0; dlt >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, dlt);
                excl += v;
                if (inc || *timed_out)
                    done = true;
            }
            idx -= 256;
        }
        if (lane == 0)
            atomicExch(state + w, (unsigned long long)(excl + count) | LB_FLAG_INC);
    }
    return excl;
}

// ------------------------------------------------------------------------------------------
// LEVEL-MAJOR launch order of the tiles of a batch.  Tile k of a unit depends on the unit's tiles < k: the survivors
// of a dead-time tile are appended behind those of the earlier tiles (decoupled look-back above), a background tile
// starts at the time where the previous one ended.  If the tiles were launched unit by unit, the ~900 tiles in
// flight would all belong to the same unit and every one of them would wait for its in-flight predecessors
// (measured in k_merge_deadtime: 25-40 % of a tile's life).  They are launched level-major instead: tile k of every
// unit that has one, then tile k+1 of every unit, ... -- in a population batch (thousands of units) the predecessor
// of a tile finished a whole level ago and the wait is one load; with few units (the bright burst: 14) the tiles
// in flight are spread over the units.
// k_level_order (one CTA) turns the per-unit tile counts n_u (computed on the device) into
//   cum[k]    = sum_u min(n_u, k)          first ticket of level k
//   sorted[r] = units by descending n_u    (the units alive at level k are sorted[0 .. above[k]) )
//   ubeg[u]   = sum_{u' < u} n_u'          the unit's first slot in a per-tile state array
// and a plan kernel maps ticket T -> (level k by bisection of cum, unit sorted[T - cum[k]]).  A tile only ever
// waits for tiles with smaller tickets; CTAs of a 1-D grid are dispatched in blockIdx order (the assumption CUB's
// decoupled look-back scan makes too), and every wait is bounded (CGRB_ST_INTERNAL instead of a hang).
// ------------------------------------------------------------------------------------------
struct LevelOrder {
    int32_t *hist;          // [cap + 2] zeroed by the caller: hist[n] = #{u : n_u == n}
    int32_t *above;         // [cap + 2] above[k] = #{u : n_u > k}
    int64_t *cum;           // [cap + 2]
    int32_t *sorted;        // [n_units]
    int32_t *eq;            // [n_units] scratch: rank of a unit among those with the same n_u
    int64_t *ubeg;          // [n_units + 1] or NULL
    int64_t *header;        // [0] = number of live tickets, [1] = levels (max n_u)
    int64_t cap;            // bound on any n_u
};
#define BKG_TILE 2048           // candidates per tile of the one-pass background kernel

#define ORDER_NT 1024
__device__ __forceinline__ long long order_block_excl_scan(long long v, long long *s_w, long long *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
            incl += o;
    }
    if (lane == 31)
        s_w[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        long long x = s_w[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d)
                x += o;
        }
        s_w[lane] = x;
    }
    __syncthreads();
    const long long r = (wid > 0 ? s_w[wid - 1] : 0) + incl - v;
    *total = s_w[ORDER_NT / 32 - 1];
    __syncthreads();
    return r;
}

// tiles of unit u.  WHICH 0: live tiles of the one-pass combine + dead time (MD_TILE merged events each, from the
// event counts; also records n_total); WHICH 1: tiles of the one-pass background (BKG_TILE candidates each, from the
// capacity; a unit with background owns at least one: it carries the tstart event)
template <int WHICH>
__device__ __forceinline__ int level_tiles_of(const cgrb_unit_t *__restrict__ units,
                                              const int64_t *__restrict__ unit_work_begin, cgrb_counts_t *counts,
                                              int u, int64_t cap, bool first_pass)
{
    if (WHICH == 0) {
        const int64_t total = counts[u].n_bkg + counts[u].n_src;
        if (first_pass) {
            counts[u].n_total = total;
            if (total == 0)
                counts[u].n_kept = 0;           // no tile will write it
        }
        const int64_t tiles_cap = unit_work_begin[u + 1] - unit_work_begin[u];
        return (int)min(min((total + 2048 - 1) / 2048, tiles_cap), cap);
    }
    if (units[u].flags & CGRB_UNIT_NO_BACKGROUND)
        return 0;
    return (int)min(max((units[u].bkg_cap + BKG_TILE - 1) / BKG_TILE, (int64_t)1), cap);
}

template <int WHICH>
__global__ void __launch_bounds__(ORDER_NT) k_level_order(const cgrb_unit_t *__restrict__ units,
                                                          const int64_t *__restrict__ unit_work_begin, int n_units,
                                                          cgrb_counts_t *counts, LevelOrder o)
{
    __shared__ long long s_w[ORDER_NT / 32];
    __shared__ int s_max;
    if (threadIdx.x == 0)
        s_max = 0;
    __syncthreads();
    int mx = 0;
    for (int u = threadIdx.x; u < n_units; u += ORDER_NT) {
        const int n = level_tiles_of<WHICH>(units, unit_work_begin, counts, u, o.cap, true);
        o.eq[u] = atomicAdd(&o.hist[n], 1);
        mx = max(mx, n);
    }
    atomicMax(&s_max, mx);
    __threadfence();
    __syncthreads();
    const int levels = s_max;
    // above[k] = #{u : n_u > k} = sum of hist[k+1 ..]: a scan from the top, 1024 levels per round
    long long carry = 0;
    for (int top = levels; top > 0; top -= ORDER_NT) {
        const int k = top - 1 - (int)threadIdx.x;           // this round covers k = top-1 .. top-1024, descending
        const long long v = (k >= 0) ? o.hist[k + 1] : 0;
        long long tot;
        const long long ex = order_block_excl_scan(v, s_w, &tot);
        if (k >= 0)
            o.above[k] = (int32_t)(carry + ex + v);
        carry += tot;
    }
    if (threadIdx.x == 0)
        o.above[levels] = 0;
    __threadfence();
    __syncthreads();
    // position of a unit in the descending order: the units with more tiles, then its rank among its equals;
    // the unit's first state slot: exclusive sum of the tile counts in unit order
    carry = 0;
    for (int base = 0; base < n_units; base += ORDER_NT) {
        const int u = base + (int)threadIdx.x;
        const int n = (u < n_units) ? level_tiles_of<WHICH>(units, unit_work_begin, counts, u, o.cap, false) : 0;
        if (n > 0)
            o.sorted[o.above[n] + o.eq[u]] = u;
        if (o.ubeg) {
            long long tot;
            const long long ex = order_block_excl_scan(n, s_w, &tot);
            if (u < n_units)
                o.ubeg[u] = carry + ex;
            carry += tot;
        }
    }
    if (o.ubeg && threadIdx.x == 0)
        o.ubeg[n_units] = carry;
    // cum[k] = sum_{k' < k} above[k'], k = 0 .. levels
    carry = 0;
    for (int base = 0; base <= levels; base += ORDER_NT) {
        const int k = base + (int)threadIdx.x;
        const long long v = (k < levels) ? o.above[k] : 0;
        long long tot;
        const long long ex = order_block_excl_scan(v, s_w, &tot);
        if (k <= levels)
            o.cum[k] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        o.header[0] = carry;
        o.header[1] = levels;
    }
}

// ticket -> (unit, level); false beyond the last live ticket
__device__ __forceinline__ bool level_ticket(const LevelOrder &o, int64_t T, int *unit, int *level)
{
    if (T >= o.header[0])
        return false;
    int lo = 0, hi = (int)o.header[1];      // the last k with cum[k] <= T
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (o.cum[mid] <= T)
            lo = mid;
        else
            hi = mid;
    }
    *unit = o.sorted[T - o.cum[lo]];
    *level = lo;
    return true;
}
