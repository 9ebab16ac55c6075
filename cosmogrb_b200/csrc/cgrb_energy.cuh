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

// per-unit tables of the energy samplers, built by k_energy_tables (device memory, one per unit)
struct EnergyUnitTab {
    // upper hull of the lines y_m(w) = gl[m] - gx[m] w  (w = 1/ec): argmax_m f(E_m, t) as a
    // function of w.  hull_m[k] is the argmax for hull_w[k] <= w < hull_w[k+1].
    int n_hull;
    int pad;
    double hull_w[CGRB_NE_ENVELOPE];
    int hull_m[CGRB_NE_ENVELOPE];
    // envelope of phi(l) = a l - e^l on CGRB_ENV_SEGS equal segments of [l0, l0 + SEGS dl]
    double l0, dl, inv_dl;
    double a;                          // alpha + 1
    double amax;                       // max of max(A, 0) over [emin, emax]
    double cdf[CGRB_ENV_SEGS + 1];     // cumulative segment weights b_j dl (unnormalised)
    double inv_b[CGRB_ENV_SEGS];       // 1 / b_j  (0 where b_j = 0)
    double lstar[CGRB_ENV_SEGS];       // argmax of phi on the segment
    double sstar[CGRB_ENV_SEGS];       // e^lstar
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
// per-unit table build: one CTA per unit
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_energy_tables(const double *__restrict__ dtab,
                                                       const cgrb_unit_t *__restrict__ units, int n_units,
                                                       int interp_extrap, EnergyUnitTab *__restrict__ etab)
{
    __shared__ double s_gl[CGRB_NE_ENVELOPE];
    __shared__ double s_gx[CGRB_NE_ENVELOPE];
    __shared__ double s_b[CGRB_ENV_SEGS];
    __shared__ double s_red[256];
    const int ui = blockIdx.x;
    if (ui >= n_units)
        return;
    const cgrb_unit_t u = units[ui];
    const double *dt = dtab + (size_t)u.det * CGRB_DT_SIZE;
    EnergyUnitTab *T = etab + ui;
    const double opz = 1.0 + u.z, alpha = u.alpha;
    for (int m = threadIdx.x; m < CGRB_NE_ENVELOPE; m += blockDim.x) {
        double x = dt[CGRB_DT_G500_E + m] * opz;
        double A = dt[CGRB_DT_G500_A + m];
        s_gx[m] = x;
        s_gl[m] = (A > 0.0) ? log(A) + alpha * log(x) : -INFINITY;
    }
    // Amax over the table nodes and the two domain ends
    double am = 0.0;
    for (int m = threadIdx.x; m < CGRB_NBINS; m += blockDim.x)
        am = fmax(am, dt[CGRB_DT_EA_Y + m]);
    if (threadIdx.x == 0) {
        am = fmax(am, interp1(dt + CGRB_DT_EA_X, dt + CGRB_DT_EA_Y, CGRB_NBINS, dt[CGRB_DT_EDGES], interp_extrap));
        am = fmax(am, interp1(dt + CGRB_DT_EA_X, dt + CGRB_DT_EA_Y, CGRB_NBINS, dt[CGRB_DT_EDGES + CGRB_NBINS],
                              interp_extrap));
    }
    s_red[threadIdx.x] = am;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d)
            s_red[threadIdx.x] = fmax(s_red[threadIdx.x], s_red[threadIdx.x + d]);
        __syncthreads();
    }
    // range of s = x (1+z) / ec over the unit's photons
    const double emin = dt[CGRB_DT_EDGES], emax = dt[CGRB_DT_EDGES + CGRB_NBINS];
    double ec0, ec1;
    {
        double ep0 = u.ep_start, ep1 = u.ep_start;
        if (u.kind == 0) {
            ep0 = u.ep_start / (1.0 + u.src_tstart / u.ep_tau);
            ep1 = u.ep_start / (1.0 + u.src_tstop / u.ep_tau);
        }
        double d = (alpha == -2.0) ? 0.0001 : (2.0 + alpha);
        ec0 = ep0 / d;
        ec1 = ep1 / d;
    }
    const double w_lo = opz / fmax(ec0, ec1), w_hi = opz / fmin(ec0, ec1);
    double l_lo = log(emin * w_lo) - 1e-9, l_hi = log(emax * w_hi) + 1e-9;
    l_hi = fmin(l_hi, log(740.0));           // e^-s underflows beyond: zero weight anyway
    if (!(l_hi > l_lo + 1e-6))
        l_hi = l_lo + 1e-6;
    const double dl = (l_hi - l_lo) / CGRB_ENV_SEGS;
    const double a = alpha + 1.0;
    if (threadIdx.x < CGRB_ENV_SEGS) {
        const int j = threadIdx.x;
        double la = l_lo + j * dl, lb = l_lo + (j + 1) * dl;
        double ls;                         // argmax of phi(l) = a l - e^l on [la, lb]
        if (a > 0.0) {
            double lpk = log(a);
            ls = fmin(fmax(lpk, la), lb);
        } else
            ls = la;
        double ss = exp(ls);
        double b = exp(a * ls - ss);
        s_b[j] = b;
        T->inv_b[j] = (b > 0.0) ? 1.0 / b : 0.0;
        T->lstar[j] = ls;
        T->sstar[j] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        T->l0 = l_lo;
        T->dl = dl;
        T->inv_dl = 1.0 / dl;
        T->a = a;
        T->amax = s_red[0];
        double acc = 0.0;
        T->cdf[0] = 0.0;
        for (int j = 0; j < CGRB_ENV_SEGS; j++) {
            acc += s_b[j] * dl;
            T->cdf[j + 1] = acc;
        }
        // upper hull, lines in order of increasing slope (-gx): m = 499 .. 0
        int n = 0;
        for (int m = CGRB_NE_ENVELOPE - 1; m >= 0; m--) {
            if (!(s_gl[m] > -INFINITY))
                continue;
            double wx = -INFINITY;
            while (n > 0) {
                int t = T->hull_m[n - 1];
                // line m overtakes line t (gx[t] > gx[m]) for w > (gl[t] - gl[m]) / (gx[t] - gx[m])
                wx = (s_gl[t] - s_gl[m]) / (s_gx[t] - s_gx[m]);
                if (wx <= T->hull_w[n - 1])
                    n--;            // t never wins (ties resolve to the lower m, like np.argmax)
                else
                    break;
            }
            if (n == 0)
                wx = -INFINITY;
            T->hull_m[n] = m;
            T->hull_w[n] = wx;
            n++;
        }
        T->n_hull = n;
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
