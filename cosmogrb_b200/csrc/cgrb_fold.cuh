// cgrb_fold.cuh -- Response.digitize: guided nearest-CDF channel search, guide staging (TMA bulk copy), small math helpers
// Part of the single translation unit cgrb_kernels.cu (included in order; see that file).
#pragma once

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
                                                 const int32_t *__restrict__ trials,
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
        if (rng.mode == 0 && trials && trials[off + i] > 0 && e == e) {
            // a photon of the tabulated-envelope sampler: its channel uniform rides in the Philox call of the trial
            // that accepted it (trial_draw), so the fused fold can be replayed here from the trial counts
            int idx = lower_bound_255(s_edges, CGRB_NBINS + 1, e) - 1;
            if (idx < 0)
                idx += CGRB_NBINS;
            idx = min(idx, CGRB_NBINS - 1);
            const Philox4 p = philox_at(make_key(rng.seed), sid, ST_SRC_ENERGY, (uint64_t)i, (uint32_t)(trials[off + i] - 1));
            pha_src[off + i] = (uint8_t)nearest_cdf_guided(s_guide + idx * CGRB_NCHAN, cum + (size_t)idx * CGRB_NCHAN,
                                                           trial_draw(p).C);
        } else if (rng.mode == 0)
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

