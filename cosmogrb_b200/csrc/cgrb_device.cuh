// cgrb_device.cuh -- device-side building blocks of the photon-event engine (sm_100a).
//
// Philox4x32-10 counter RNG, fp64 special functions (regularised upper incomplete gamma),
// the CPL spectral model of cosmogrb and small warp/block primitives.  Reference lines cited
// are relative to the cosmogrb repository root.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/cosmogrb_b200.h"
#include "cgrb_logtab.cuh"

namespace cgrb {

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = (c0,c1,c2,c3), key = (k0,k1).
// ------------------------------------------------------------------------------------------
#ifndef CGRB_PHILOX_ROUNDS
#define CGRB_PHILOX_ROUNDS 10       // Philox4x32-10; other values only for tuning measurements
#endif
struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < CGRB_PHILOX_ROUNDS; r++) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// uniform double in [0,1) from 64 random bits: the top 52 bits become the mantissa of a double in
// [1,2), minus one (resolution 2^-52; two shifts, two ORs and one add -- no integer-to-double conversions)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return __hiloint2double((int)(0x3FF00000u | (a >> 12)), (int)((a << 20) | (b >> 12))) - 1.0;
}

// -ln(u) for a positive normal double (the exponential gaps under Philox: u = u53(...) is a multiple of 2^-52).
// u = 2^k z, z in [0.6875, 1.375) cut into 128 intervals by its top mantissa bits; the table holds 1 / c and ln c
// of each interval's centre (scripts/make_logtab.py), so ln u = k ln2 + ln c + log1p(r) with r = z / c - 1 from one
// fused multiply-add, |r| < 2^-7.8, and a degree-7 polynomial.  ~20 instructions against ~85 for log(); error
// <= 1.5 ulp of max(|ln u|, 1) (checked by the generator against long-double logarithms and on the device
// against log() in tests/test_gpu_guards.py).  u == 0 gives +inf like -log(0).  Injected uniforms keep log().
__device__ __forceinline__ double neglog_fast(double u)
{
    const int hi = __double2hiint(u);
    const int tmp = hi - 0x3FE60000;
    const int k = tmp >> 20;
    const double2 tc = __ldg(&CGRB_LOGTAB[(tmp >> 13) & 127]);
    const double z = __hiloint2double(hi - (tmp & (int)0xFFF00000), __double2loint(u));
    const double r = fma(z, tc.x, -1.0);
    double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    const double w = fma((double)k, 0.6931471805599453, tc.y);
    const double v = -(w + fma(r * r, p, r));
    return (u > 0.0) ? v : INFINITY;
}

// One Philox call carries a whole trial of the tabulated-envelope energy sampler: 48 bits for the proposal, 48 for
// the height, 32 for the channel draw of the photon if this trial accepts it (independent bits of one output).
// Uniforms in [0,1) with resolution 2^-48 / 2^-32, built like u53 (mantissa bits of a double in [1,2), minus one).
__device__ __forceinline__ double u48(uint32_t a, uint32_t b16)
{
    return __hiloint2double((int)(0x3FF00000u | (a >> 12)), (int)((a << 20) | (b16 << 4))) - 1.0;
}
__device__ __forceinline__ double u32d(uint32_t a)
{
    return __hiloint2double((int)(0x3FF00000u | (a >> 12)), (int)(a << 20)) - 1.0;
}
// One Philox call carries a PAIR of background events: 40 bits for each exponential gap, 24 for each channel draw.
// Gap uniforms are (k + 1/2) 2^-40, k < 2^40: never 0 (a zero would end the unit's list), resolution far below one
// ulp of an arrival time; channel uniforms are k 2^-24.
__device__ __forceinline__ double u40h(uint32_t a, uint32_t b8)
{
    return __hiloint2double((int)(0x3FF00000u | (a >> 12)), (int)((a << 20) | (b8 << 12) | 0x800u)) - 1.0;
}
__device__ __forceinline__ double u24d(uint32_t c)
{
    return __hiloint2double((int)(0x3FF00000u | (c >> 4)), (int)(c << 28)) - 1.0;
}
struct BkgDraw {
    double ga, gb;      // gap uniforms of the pair's two events
    uint32_t ca, cb;    // their channel draws, 24 bits each
};
__device__ __forceinline__ BkgDraw bkg_draw(const Philox4 p)
{
    return BkgDraw{u40h(p.x, p.y >> 24), u40h(p.z, (p.y >> 16) & 0xffu), p.w >> 8, ((p.w & 0xffu) << 16) | (p.y & 0xffffu)};
}
struct TrialDraw {
    double U, V, C;     // proposal, height, channel
};
__device__ __forceinline__ TrialDraw trial_draw(const Philox4 p)
{
    return TrialDraw{u48(p.x, p.y >> 16), u48(p.z, p.y & 0xffffu), u32d(p.w)};
}

// stages (Philox counter word 3)
enum Stage : uint32_t {
    ST_SRC_TIME = 1,   // candidate k: xy -> gap uniform, zw -> acceptance uniform
    ST_SRC_ENERGY = 2, // photon i, trial j: literal sampler xy -> proposal, zw -> height; envelope sampler: trial_draw
    ST_SRC_CHAN = 3,   // photon i: xy -> channel uniform (stand-alone fold, photons of the literal loop)
    ST_BKG = 4,        // candidate pair m: bkg_draw (two gap uniforms + two channel draws from one call); event 0: index 2^40 - 1
};

struct RngKey {
    uint32_t k0, k1;
};
__device__ __forceinline__ RngKey make_key(uint64_t seed)
{
    return RngKey{(uint32_t)seed, (uint32_t)(seed >> 32)};
}
// index < 2^40 is supported: low 32 bits in c0, next 8 bits folded into the stage word.
__device__ __forceinline__ Philox4 philox_at(RngKey key, uint32_t stream_id, uint32_t stage,
                                             uint64_t index, uint32_t sub)
{
    return philox4x32_10((uint32_t)index, sub, stream_id, stage | ((uint32_t)(index >> 32) << 8), key.k0,
                         key.k1);
}

// ------------------------------------------------------------------------------------------
// Regularised upper incomplete gamma Q(a, x) for a = 2 + alpha in (0, 2): the cephes-lineage
// igamc that scipy.special.gammaincc evaluates at cpl_functions.py:26-27.  lgamma(a) and
// lgamma(1+a) are per-GRB constants computed on the host.
// ------------------------------------------------------------------------------------------
#define CGRB_MACHEP 1.11022302462515654042E-16
#define CGRB_MAXLOG 7.09782712893383996843E2

__device__ __forceinline__ double igam_fac(double a, double x, double logx, double lgam_a)
{
    double ax = a * logx - x - lgam_a;
    if (ax < -CGRB_MAXLOG)
        return 0.0;
    return exp(ax);
}

__device__ __forceinline__ double igam_series(double a, double x, double logx, double lgam_a)
{
    double ax = igam_fac(a, x, logx, lgam_a);
    if (ax == 0.0)
        return 0.0;
    double r = a, c = 1.0, ans = 1.0;
    for (int i = 0; i < 2000; i++) {
        r += 1.0;
        c *= x / r;
        ans += c;
        if (c <= CGRB_MACHEP * ans)
            break;
    }
    return ans * ax / a;
}

__device__ __forceinline__ double igamc_series(double a, double x, double logx, double lgam_a,
                                               double lgam_1pa)
{
    double fac = 1.0, sum = 0.0, term;
    for (int n = 1; n < 2000; n++) {
        fac *= -x / (double)n;
        term = fac / (a + (double)n);
        sum += term;
        if (fabs(term) <= CGRB_MACHEP * fabs(sum))
            break;
    }
    term = -expm1(a * logx - lgam_1pa);
    return term - exp(a * logx - lgam_a) * sum;
}

__device__ __forceinline__ double igamc_cf(double a, double x, double logx, double lgam_a)
{
    double ax = igam_fac(a, x, logx, lgam_a);
    if (ax == 0.0)
        return 0.0;
    const double big = 4.503599627370496e15, biginv = 2.22044604925031308085e-16;
    double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
    double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
    double ans = pkm1 / qkm1, t;
    for (int i = 0; i < 2000; i++) {
        c += 1.0;
        y += 1.0;
        z += 2.0;
        double yc = y * c;
        double pk = pkm1 * z - pkm2 * yc;
        double qk = qkm1 * z - qkm2 * yc;
        if (qk != 0.0) {
            double r = pk / qk;
            t = fabs((ans - r) / r);
            ans = r;
        } else
            t = 1.0;
        pkm2 = pkm1;
        pkm1 = pk;
        qkm2 = qkm1;
        qkm1 = qk;
        if (fabs(pk) > big) {
            pkm2 *= biginv;
            pkm1 *= biginv;
            qkm2 *= biginv;
            qkm1 *= biginv;
        }
        if (t <= CGRB_MACHEP)
            break;
    }
    return ans * ax;
}

__device__ __forceinline__ double gammaincc(double a, double x, double lgam_a, double lgam_1pa)
{
    if (x < 0.0 || a < 0.0)
        return nan("");
    if (a == 0.0)
        return (x > 0.0) ? 0.0 : nan("");
    if (x == 0.0)
        return 1.0;
    if (isinf(x))
        return 0.0;
    double logx = log(x);
    if (x > 1.1) {
        if (x < a)
            return 1.0 - igam_series(a, x, logx, lgam_a);
        return igamc_cf(a, x, logx, lgam_a);
    } else if (x <= 0.5) {
        if (-0.4 / logx < a)
            return 1.0 - igam_series(a, x, logx, lgam_a);
        return igamc_series(a, x, logx, lgam_a, lgam_1pa);
    } else {
        if (x * 1.1 < a)
            return 1.0 - igam_series(a, x, logx, lgam_a);
        return igamc_series(a, x, logx, lgam_a, lgam_1pa);
    }
}

// ------------------------------------------------------------------------------------------
// CPL spectral model.  cpl(x) = norm(t) * exp(alpha*(log x - log ec) - x/ec)  (cpl_functions.py:11-44)
// norm depends on time only: it is evaluated once per candidate / photon instead of once per
// energy as the reference does.
// ------------------------------------------------------------------------------------------
struct SpecT {
    double norm;    // F * erg2keV / intflux
    double ec;      // cutoff energy
    double log_ec;
    double inv_ec;
};

// sampler/temporal_functions.py:5-14 with t_start = 0
__device__ __forceinline__ double norris(double x, double K, double t_rise, double t_decay)
{
    if (x > 0.0)
        return K * exp(2.0 * sqrt(t_rise / t_decay)) * exp(-t_rise / x - x / t_decay);
    return 0.0;
}

// spectrum with energy flux F and peak energy ep
__device__ __forceinline__ SpecT spec_with(const cgrb_unit_t &u, double F, double ep)
{
    const double opz = 1.0 + u.z;
    const double a = 10.0 * opz, b = 1.0e4 * opz;        // :78-79
    double ec = (u.alpha == -2.0) ? ep / 0.0001 : ep / (2.0 + u.alpha);  // :13-20
    const double ga = 2.0 + u.alpha;
    double i1 = gammaincc(ga, a / ec, u.lgamma_a, u.lgamma_1pa) * u.gamma_a;   // :26
    double i2 = gammaincc(ga, b / ec, u.lgamma_a, u.lgamma_1pa) * u.gamma_a;   // :27
    double intflux = -ec * ec * (i2 - i1);               // :29
    SpecT s;
    s.norm = F * 6.24151e8 / intflux;                    // :33-35
    s.ec = ec;
    s.log_ec = log(ec);
    s.inv_ec = 1.0 / ec;
    return s;
}

__device__ __forceinline__ SpecT spec_at(const cgrb_unit_t &u, double t)
{
    if (u.kind == 0)
        return spec_with(u, norris(t, u.peak_flux, u.trise, u.tdecay),       // cpl_functions.py:86-88
                         u.ep_start / (1.0 + t / u.ep_tau));                 // :90
    return spec_with(u, u.peak_flux, u.ep_start);                            // cpl_constant_functions.py:40-42
}

// literal cpl value at rest-frame energy x with log(x) supplied (cpl_functions.py:37-44)
__device__ __forceinline__ double cpl_value(const SpecT &s, double alpha, double x, double logx)
{
    return s.norm * exp(alpha * (logx - s.log_ec) - x / s.ec);
}

// np.searchsorted(xs, v, side='left')
__device__ __forceinline__ int searchsorted_left(const double *xs, int n, double v)
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
// np.searchsorted(xs, v, side='right')
__device__ __forceinline__ int searchsorted_right(const double *xs, int n, double v)
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

// interpolation.interp(x, y, v) at cpl_functions.py:119.  extrap 0: bracket index clamped,
// weight free (linear extrapolation); 1: np.interp (value clamped).
__device__ __forceinline__ double interp1(const double *xs, const double *ys, int n, double v, int extrap)
{
    if (extrap == 0) {
        int i = searchsorted_left(xs, n, v) - 1;
        i = max(0, min(i, n - 2));
        double lam = (v - xs[i]) / (xs[i + 1] - xs[i]);
        return (1.0 - lam) * ys[i] + lam * ys[i + 1];
    }
    if (v <= xs[0])
        return ys[0];
    if (v >= xs[n - 1])
        return ys[n - 1];
    int j = searchsorted_right(xs, n, v) - 1;
    double slope = (ys[j + 1] - ys[j]) / (xs[j + 1] - xs[j]);
    return slope * (v - xs[j]) + ys[j];
}

// ------------------------------------------------------------------------------------------
// warp / block primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_incl_scan(double v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d)
            v += n;
    }
    return v;
}

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d)
            v += n;
    }
    return v;
}

}  // namespace cgrb
