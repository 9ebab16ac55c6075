// cgrb_energy.cuh -- photon-energy sampling (sampler/cpl_functions.py:191-286,
// cpl_constant_functions.py:108-177).
//
// Two samplers with the SAME output distribution:
//
//  * literal  (k_sample_energy): the reference's rejection loop, trial for trial -- power-law
//    proposal, envelope C (x / E_idx)^alpha with C = 5 max_m f(E_m) on the 500-point grid.  Used
//    under injected uniforms (decisions replay the reference's stream) and selectable under Philox.
//    One warp per photon, 32 trials per round, lowest accepted trial wins (= the serial loop).
//
//  * envelope (k_sample_energy_env): exact rejection sampling of the literal sampler's OUTPUT
//    density  h(x) = min(f(x,t), C (x/E_idx)^alpha)  from a much tighter proposal: a per-unit,
//    time-independent piecewise-exponential envelope of s^(alpha+1) e^(-s) in l = log s, with
//    s = x (1+z) / ec(t), truncated per photon to [s(emin), s(emax)] by inverse CDF, times the
//    bound A(x) <= Amax on the effective area.  ~2-5 trials per photon instead of 30-600.
//    One thread per photon.  Philox only.
//
// f(x,t) = A(x) norm(t) (x (1+z) / ec)^alpha exp(-x (1+z) / ec),  A = interp(ea_x, ea_y, x).
#pragma once

#include "cgrb_device.cuh"

namespace cgrb {

#define CGRB_ENV_SEGS 256
#define CGRB_ENV_MAXW 256

// per-unit tables of the energy samplers, built by k_energy_tables (device memory, one per unit)
struct EnergyUnitTab {
    // upper hull of the lines y_m(w) = gl[m] - gx[m] w  (w = 1/ec): argmax_m f(E_m, t) as a
    // function of w.  hull_m[k] is the argmax for hull_w[k] <= w < hull_w[k+1].
    int n_hull;
    int n_w;                           // rows of the unit's proposal table (grid of w = (1+z)/ec)
    double hull_w[CGRB_NE_ENVELOPE];
    int hull_m[CGRB_NE_ENVELOPE];
    double hull_E[CGRB_NE_ENVELOPE];   // grid energy E_m of hull entry k (observer frame)
    double hull_lA5[CGRB_NE_ENVELOPE]; // log(5 A(E_m)): the literal envelope's constant C = 5 max f
    // proposal: CGRB_ENV_SEGS equal segments in u = log x on [log emin, log emax]
    double u0, du;
    double a;                          // alpha + 1
    double logw0, inv_dlogw;           // row k covers w in [wk[k], wk[k+1])
    double xlo[CGRB_ENV_SEGS + 1];     // segment edges in x (xlo[0] = emin, xlo[SEGS] = emax)
    double logM[CGRB_ENV_SEGS];        // log max over the segment of x^(alpha+1) max(A(x), 0)  (-inf: none)
    double invMx[CGRB_ENV_SEGS];       // xlo^(alpha+1) / M: x^(alpha+1) / M = invMx exp((alpha+1) z), x = xlo e^z
    double logAmax[CGRB_ENV_SEGS];     // log max over the segment of max(A(x), 0)
    double wk[CGRB_ENV_MAXW];          // lower edges of the w rows
    int kn[CGRB_ENV_SEGS];             // first effective-area node above the segment's lower edge
};

// interp(ea_x, ea_y, v) with a log-uniform first guess for the bracket (the bin means of
// log-spaced edges are log-uniform) and an exact fix-up, so any monotone table works.
// Same bracket as searchsorted_left - 1 clamped to [0, n-2]; linear extrapolation (extrap 0) or
// clamped value (extrap 1) outside the table.
struct EaGuess {
    double log_x0, inv_dlog;
};
__device__ __forceinline__ double interp_guess(const double *xs, const double *ys, int n, double v, double logv,
                                               EaGuess g, int extrap)
{
    if (extrap == 1) {
        if (v <= xs[0])
            return ys[0];
        if (v >= xs[n - 1])
            return ys[n - 1];
    }
    int i = (int)floor((logv - g.log_x0) * g.inv_dlog);
    i = max(0, min(i, n - 2));
    while (i < n - 2 && xs[i + 1] < v)
        i++;
    while (i > 0 && !(xs[i] < v))
        i--;
    if (extrap == 1) {
        // np.interp bracket: xs[j] <= v < xs[j+1]
        if (i < n - 2 && xs[i + 1] == v)
            i++;
        double slope = (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]);
        return slope * (v - xs[i]) + ys[i];
    }
    double lam = (v - xs[i]) / (xs[i + 1] - xs[i]);
    return (1.0 - lam) * ys[i] + lam * ys[i + 1];
}

// ------------------------------------------------------------------------------------------
// per-unit table build: one CTA per unit (256 threads = one thread per proposal segment)
// ------------------------------------------------------------------------------------------

// max over [xa, xb] of G(x) = x^a max(A(x), 0) and of max(A(x), 0), A = interp(ea_x, ea_y, .) piecewise
// linear with knots at the table nodes: endpoints of every linear piece plus the interior stationary
// point x* = -a p / ((a + 1) q) of x^a (p + q x).
__device__ __forceinline__ void segment_max(const double *eax, const double *eay, int extrap, double a, double xa,
                                            double xb, double *g_max, double *a_max)
{
    double gm = 0.0, am = 0.0;
    int i = searchsorted_right(eax, CGRB_NBINS, xa);      // first knot > xa
    double p1 = xa;
    double r1 = interp1(eax, eay, CGRB_NBINS, p1, extrap);
    for (;;) {
        const bool last = !(i < CGRB_NBINS && eax[i] < xb);
        const double p2 = last ? xb : eax[i];
        const double r2 = interp1(eax, eay, CGRB_NBINS, p2, extrap);
        const double c1 = fmax(r1, 0.0), c2 = fmax(r2, 0.0);
        am = fmax(am, fmax(c1, c2));
        gm = fmax(gm, fmax(c1 * exp(a * log(p1)), c2 * exp(a * log(p2))));
        if (p2 > p1) {
            const double q = (r2 - r1) / (p2 - p1), p = r1 - q * p1;
            const double den = (a + 1.0) * q;
            if (den != 0.0) {
                const double xs = -a * p / den;
                if (xs > p1 && xs < p2)
                    gm = fmax(gm, fmax(p + q * xs, 0.0) * exp(a * log(xs)));
            }
        }
        if (last)
            break;
        p1 = p2;
        r1 = r2;
        i++;
    }
    *g_max = gm;
    *a_max = am;
}

__global__ void __launch_bounds__(256) k_energy_tables(const double *__restrict__ dtab,
                                                       const cgrb_unit_t *__restrict__ units, int n_units,
                                                       int interp_extrap, EnergyUnitTab *__restrict__ etab)
{
    __shared__ double s_gl[CGRB_NE_ENVELOPE];
    __shared__ double s_gx[CGRB_NE_ENVELOPE];
    __shared__ double s_eax[CGRB_NBINS];
    __shared__ double s_eay[CGRB_NBINS];
    const int ui = blockIdx.x;
    if (ui >= n_units)
        return;
    const cgrb_unit_t u = units[ui];
    const double *dt = dtab + (size_t)u.det * CGRB_DT_SIZE;
    EnergyUnitTab *T = etab + ui;
    const double opz = 1.0 + u.z, alpha = u.alpha;
    for (int m = threadIdx.x; m < CGRB_NE_ENVELOPE; m += blockDim.x) {
        double x = dt[CGRB_DT_G500_E + m] * opz;
        double A = dt[CGRB_DT_G500_A + m];   // ea_scale is a constant factor: no effect on the energy law
        s_gx[m] = x;
        s_gl[m] = (A > 0.0) ? log(A) + alpha * log(x) : -INFINITY;
    }
    for (int m = threadIdx.x; m < CGRB_NBINS; m += blockDim.x) {
        s_eax[m] = dt[CGRB_DT_EA_X + m];
        s_eay[m] = dt[CGRB_DT_EA_Y + m];
    }
    __syncthreads();
    const double emin = dt[CGRB_DT_EDGES], emax = dt[CGRB_DT_EDGES + CGRB_NBINS];
    const double a = alpha + 1.0;
    const double u0 = log(emin), du = (log(emax) - u0) / CGRB_ENV_SEGS;
    if (threadIdx.x < CGRB_ENV_SEGS) {
        const int j = threadIdx.x;
        const double xa = (j == 0) ? emin : exp(u0 + j * du);
        const double xb = (j == CGRB_ENV_SEGS - 1) ? emax : exp(u0 + (j + 1) * du);
        double gm, am;
        // bounds taken on a slightly wider interval: x = exp(u) may round one ulp outside the segment
        segment_max(s_eax, s_eay, interp_extrap, a, xa * (1.0 - 1e-13), xb * (1.0 + 1e-13), &gm, &am);
        T->xlo[j] = xa;
        if (j == CGRB_ENV_SEGS - 1)
            T->xlo[CGRB_ENV_SEGS] = emax;
        T->kn[j] = searchsorted_right(s_eax, CGRB_NBINS, xa);
        T->logM[j] = (gm > 0.0) ? log(gm) + 1e-9 : -INFINITY;
        T->invMx[j] = (gm > 0.0) ? exp(a * log(xa) - (log(gm) + 1e-9)) : 0.0;
        T->logAmax[j] = (am > 0.0) ? log(am) + 1e-9 : -INFINITY;
    }
    // grid of w = (1+z)/ec(t) over the unit's photons: row k bounds exp(-x w) by exp(-xlo w_k), w_k <= w
    {
        double ep0 = u.ep_start, ep1 = u.ep_start;
        if (u.kind == 0) {
            ep0 = u.ep_start / (1.0 + u.src_tstart / u.ep_tau);
            ep1 = u.ep_start / (1.0 + u.src_tstop / u.ep_tau);
        }
        const double d = (alpha == -2.0) ? 0.0001 : (2.0 + alpha);
        const double wa = opz * d / ep0, wb = opz * d / ep1;
        double lw_lo = log(fmin(wa, wb)) - 1e-9, lw_hi = log(fmax(wa, wb)) + 1e-9;
        int n_w = (int)min((int64_t)CGRB_ENV_MAXW, max((int64_t)1, u.etab_nw));
        if (!(lw_hi > lw_lo) || isinf(lw_lo) || isinf(lw_hi)) {     // degenerate spectral evolution
            n_w = 1;
            lw_hi = lw_lo + 1.0;
        }
        const double dlw = (lw_hi - lw_lo) / n_w;
        for (int k = threadIdx.x; k < CGRB_ENV_MAXW; k += blockDim.x)
            T->wk[k] = (k < n_w) ? exp(lw_lo + k * dlw) : INFINITY;
        if (threadIdx.x == 0) {
            T->n_w = n_w;
            T->logw0 = lw_lo;
            T->inv_dlogw = 1.0 / dlw;
            T->u0 = u0;
            T->du = du;
            T->a = a;
        }
    }
    // upper hull of the 500 lines y_m(w) = gl[m] - gx[m] w, in parallel: line m is the argmax on
    // [L_m, U_m) with L_m = max over t > m of the crossing with t (m beats t beyond it; ties go to the lower
    // index, like np.argmax) and U_m = min over t < m of the crossing where t takes over.  It is on the hull
    // iff L_m < U_m; hull entries are ordered by increasing w = decreasing m.  Crossings are compared by
    // cross-multiplication (all denominators gx[t] - gx[m] are positive), one division per line at the end.
    __shared__ double s_hw[CGRB_NE_ENVELOPE];
    __shared__ int s_hm[CGRB_NE_ENVELOPE];
    __shared__ unsigned char s_on[CGRB_NE_ENVELOPE];
    __shared__ double s_L[CGRB_NE_ENVELOPE];
    __shared__ int s_nh;
    for (int m = threadIdx.x; m < CGRB_NE_ENVELOPE; m += blockDim.x) {
        bool on = s_gl[m] > -INFINITY;
        double ln = 0.0, ld = 0.0, un = 0.0, ud = 0.0;          // L = ln / ld, U = un / ud (d = 0: none yet)
        if (on) {
            const double gm = s_gl[m], xm = s_gx[m];
            for (int t = m + 1; t < CGRB_NE_ENVELOPE; t++) {
                const double gt = s_gl[t];
                if (!(gt > -INFINITY))
                    continue;
                const double num = gt - gm, den = s_gx[t] - xm;      // crossing w* = num / den, den > 0
                if (ld == 0.0 || num * ld > ln * den) {
                    ln = num;
                    ld = den;
                }
            }
            for (int t = 0; t < m; t++) {
                const double gt = s_gl[t];
                if (!(gt > -INFINITY))
                    continue;
                const double num = gm - gt, den = xm - s_gx[t];
                if (ud == 0.0 || num * ud < un * den) {
                    un = num;
                    ud = den;
                }
            }
            // nonempty interval: L < U
            if (ld != 0.0 && ud != 0.0)
                on = ln * ud < un * ld;
        }
        s_on[m] = on ? 1 : 0;
        s_L[m] = (on && ld != 0.0) ? ln / ld : -INFINITY;
    }
    __syncthreads();
    if (threadIdx.x == 0)
        s_nh = 0;
    __syncthreads();
    for (int m = threadIdx.x; m < CGRB_NE_ENVELOPE; m += blockDim.x) {
        if (!s_on[m])
            continue;
        int rank = 0;                                           // hull lines with a larger index come first
        for (int t = m + 1; t < CGRB_NE_ENVELOPE; t++)
            rank += s_on[t];
        s_hm[rank] = m;
        s_hw[rank] = s_L[m];
        atomicAdd(&s_nh, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        T->n_hull = s_nh;
    for (int k = threadIdx.x; k < s_nh; k += blockDim.x) {
        const int m = s_hm[k];
        T->hull_m[k] = m;
        T->hull_w[k] = s_hw[k];
        T->hull_E[k] = dt[CGRB_DT_G500_E + m];
        T->hull_lA5[k] = log(5.0 * dt[CGRB_DT_G500_A + m]);
    }
}

// ------------------------------------------------------------------------------------------
// the literal rejection trial in cancelled form.  The reference tests
//     y = u2 C (x/E_idx)^alpha  <=  f(x,t) = A(x) norm (x (1+z)/ec)^alpha exp(-x (1+z)/ec);
// both sides carry x^alpha > 0, so the test is  u2 K1 <= A(x) K2 exp(-x w)  with the per-photon
// constants K1 = C E_idx^-alpha, K2 = norm ((1+z)/ec)^alpha, w = (1+z)/ec, and the proposal
// x = (p_span u1 + q_lo)^(1/(alpha+1)) is evaluated as exp(log(.)/(alpha+1)): one log and two
// exps per trial instead of three pows, a log and an exp.
// ------------------------------------------------------------------------------------------
struct PropConst {
    double q_lo, p_span, inv_ap1;      // cpl_functions.py:253-258
};
struct PhotonConst {
    double K1, K2, w;
};

__device__ __forceinline__ bool literal_trial(const PropConst &pc, const PhotonConst &ph, const double *eax,
                                              const double *eay, EaGuess g, int extrap, double u1, double u2,
                                              double *x_out)
{
    const double L = log(pc.p_span * u1 + pc.q_lo) * pc.inv_ap1;
    const double x = exp(L);
    const double A = interp_guess(eax, eay, CGRB_NBINS, x, L, g, extrap);
    *x_out = x;
    return u2 * ph.K1 <= A * ph.K2 * exp(-x * ph.w);
}

// per-photon constants of the literal sampler given the grid argmax idx (cpl_functions.py:231-241)
__device__ __forceinline__ PhotonConst literal_photon(const cgrb_unit_t &u, const SpecT &s, const double *dt,
                                                      const double *gx, int *idx_io)
{
    const double opz = 1.0 + u.z, alpha = u.alpha;
    int idx = *idx_io;
    double A_idx = dt[CGRB_DT_G500_A + idx];
    double f_idx = A_idx * cpl_value(s, alpha, gx[idx], log(gx[idx]));   // literal tmp[idx]
    if (!(f_idx > 0.0)) {       // all-zero row (K = 0 before the pulse) or degenerate: np.argmax -> 0
        idx = 0;
        A_idx = dt[CGRB_DT_G500_A];
        f_idx = A_idx * cpl_value(s, alpha, gx[0], log(gx[0]));
    }
    *idx_io = idx;
    PhotonConst ph;
    const double C = f_idx * 5.0;                                        // :241
    ph.K1 = C * exp(-alpha * log(dt[CGRB_DT_G500_E + idx]));
    ph.K2 = s.norm * exp(alpha * (log(opz) - s.log_ec));
    ph.w = opz * s.inv_ec;
    return ph;
}

__device__ __forceinline__ int hull_lookup(const int *hull_m, const double *hull_w, int n_hull, double w)
{
    int lo = 0, hi = n_hull;            // last k with hull_w[k] <= w
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (hull_w[mid] <= w)
            lo = mid;
        else
            hi = mid;
    }
    return hull_m[lo];
}

}  // namespace cgrb
