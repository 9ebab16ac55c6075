// cgrb_spectral.cuh -- lambda(t), evolution, fmax and the squeeze table of the thinning acceptance test
// Part of the single translation unit cgrb_kernels.cu (included in order; see that file).
#pragma once

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
// nodes on a uniform time grid AND the cell midpoints, then per-cell margins on |p - linear interpolation|.
//   margin_k = 2 |p(mid_k) - (p_k + p_k+1)/2|        the interpolation error MEASURED at the cell's midpoint, x2
//            + 0.5 max |second differences| at the nodes k-1 .. k+2   (h^2/8 sup|p''| ~ d2/8, x4)
//            + rounding slack
// Cells without a squeeze (every candidate takes the exact lambda(t)) are marked in the margin slot:
//   +inf    no bound known: the majorant proposal uses P = 1 there (cells touching a node where p == 0 without
//           being identically zero; non-finite values)
//   -B < 0  the FIRST cell of a time-evolving CPL: the Norris pulse exp(-trise/t) has an essential singularity at
//           its start that no finite difference of nodes sees, so the cell is not interpolated at all; the majorant
//           proposal uses the bound  p <= B = sup K x sup G / fmax  with K the pulse (monotone up to its maximum at
//           sqrt(trise tdecay): sup over the cell is analytic) and G = lambda / K, a smooth function of ep(t) only
//           (sup from its values at the cell's ends and midpoint plus twice their second difference).
// tests/test_majorant_oracle.py restates this on the oracle's lambda(t) and checks the bounds on a finer grid
// over the population's range of trise / tdecay / duration and table sizes (worst error / margin: 0.45).
// ==========================================================================================
#ifndef THINNODES_CTAS
#define THINNODES_CTAS 2          // 128 registers: 0.246 -> 0.192 ms on the bright burst (1 CTA/SM left the fp64 chains exposed)
#endif
__global__ void __launch_bounds__(NT, THINNODES_CTAS) k_thin_nodes(const double *__restrict__ dtab,
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
    // two threads per node: the even one evaluates the node, the odd one the midpoint of the cell that starts there
    // (left in the node's margin slot for k_thin_margins)
    for (int q = threadIdx.x; q < 2 * NT; q += NT) {
        const int64_t k = wk.start + (q >> 1);
        const int mid = q & 1;
        if (k >= u.tab_n || (mid && k >= u.tab_n - 1))
            continue;
        const double h = (u.src_tstop - u.src_tstart) / (double)(u.tab_n - 1);
        if (mid && k == 0 && u.kind == 0) {
            // first cell of a pulse: -B, the bound on p over the cell (see above)
            const double a = u.src_tstart, b = u.src_tstart + h;
            const double tpk = sqrt(u.trise * u.tdecay);
            const double supK = (b > 0.0) ? norris(fmin(fmax(tpk, a), b), u.peak_flux, u.trise, u.tdecay) : 0.0;
            const double g0 = lambda_per_flux_at(&tab, u, a), gm = lambda_per_flux_at(&tab, u, a + 0.5 * h),
                         g1 = lambda_per_flux_at(&tab, u, b);
            const double supG = fmax(g0, fmax(gm, g1)) + 2.0 * fabs(g0 - 2.0 * gm + g1);
            const double B = supK * supG / u.fmax * (1.0 + 1e-6) + 1e-300;
            thin_tab[2 * u.tab_off + 1] = (B == B && B < INFINITY) ? -B : INFINITY;
            continue;
        }
        const double t = (k == u.tab_n - 1) ? u.src_tstop : u.src_tstart + ((double)k + 0.5 * (double)mid) * h;
        thin_tab[2 * (u.tab_off + k) + mid] = lambda_at(&tab, u, t) / u.fmax;
    }
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
    const int64_t k = wk.start + threadIdx.x;      // cell [k, k+1]
    if (k >= n)
        return;
    double *p = thin_tab + 2 * u->tab_off;
    // |second differences| at the nodes k-1 .. k+2 (clamped to the interior nodes 1 .. n-2); reads node values
    // (even slots) only, writes this cell's own margin slot only: no ordering between threads is needed
    double m = 0.0;
    bool bad = false;
#pragma unroll
    for (int d = -1; d <= 2; d++) {
        int64_t c = min(max(k + d, (int64_t)1), n - 2);
        double d2 = fabs(p[2 * (c - 1)] - 2.0 * p[2 * c] + p[2 * (c + 1)]);
        bad |= !(d2 == d2) || isinf(d2);
        m = fmax(m, d2);
    }
    double margin = INFINITY;
    if (k == 0 && u->kind == 0) {
        const double mb = p[1];             // -B, left by k_thin_nodes
        if (mb < 0.0)
            margin = mb;
        bad = false;
    } else if (k < n - 1) {
        const double pk = p[2 * k], pk1 = p[2 * (k + 1)], pm = p[2 * k + 1];
        const double emid = fabs(pm - 0.5 * (pk + pk1));
        const bool touches_zero = (pk == 0.0 || pk1 == 0.0) && !(pk == 0.0 && pk1 == 0.0 && pm == 0.0);
        bad |= !(emid == emid) || isinf(emid);
        if (!touches_zero)
            margin = 0.5 * m + 2.0 * emid + 1e-12 * (fabs(pk) + 1.0);
    }
    p[2 * k + 1] = bad ? INFINITY : margin;
}

