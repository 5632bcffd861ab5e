// theta_core.cuh -- shared declarations + the per-candidate n=2 device solver.
//
// The maths follows SURVEY.md Appendix A, restated on *grouped* sufficient
// statistics: a candidate's tumour column c takes at most KV distinct values,
// and every term of the reference's per-interval sums
//     dL_dMu  (python/Optimizer.py:209-221),  M2 (:223-226),  L2 (:187-196)
// depends on the interval only through c_i once rN_i cancels:
//     f(x)  = - sum_v R_v (S1 - v A) / (x S1 + v (1-x) A),   R_v = sum_{c_i=v} r_i
//     NLL   = K0 - sum_v R_v log(g_v / gbar),  g_v = tau mu + v (1-mu),  gbar = (mu S0 + (1-mu) S1)/A
// with A = sum rN, S0 = tau A, S1 = sum rN_i c_i, K0 = -sum r_i log(rN_i/A).
// All R_v, S1 are exact integers (< 2^53) held in doubles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/theta_b200.h"

namespace theta {

// candidate-independent constants of one n=2 problem
struct N2Const {
    double A;        // sum rN
    double S0;       // tau * A
    double Rtot;     // sum r
    double K0;       // -sum r_i log(rN_i / A)
    double tau;
    double max_normal;  // 1.0 => uB = 1
    double min_normal;  // 0.0 => lB = 0 (Optimizer's lower_bound; only the explicit-column solve honours it)
    int any_zero_rN;    // some rN_i == 0 -> every candidate is NaN/None
    int m;
};

struct SolveOut {
    double x;    // root in normalised coordinates (Optimizer.py:118)
    double mu;   // M2 (Optimizer.py:223-226)
    double nll;  // L2 (Optimizer.py:187-196)
    int status;
    int iters;
};

#ifndef THETA_P1_F32HALLEY
#define THETA_P1_F32HALLEY 1
#endif
#ifndef THETA_P1_FASTM2
#define THETA_P1_FASTM2 1
#endif
#define THETA_NEWTON_MAXIT 100
#define THETA_NEWTON_PRED 2e-11   // stop when the error predicted after the step, |f''/(2f')| dN^2, is below this (relative)
#define THETA_NEWTON_RTOL 1e-9    // a converged step that leaves the bracket by rounding is accepted if it is this small (relative)

// 1/a for normal, non-zero a: MUFU.RCP64H seed (about 20 bits) + two Newton-Raphson steps (< 1 ulp
// of error in practice).  The IEEE division it replaces costs ~2.5x as many instructions.
__device__ __forceinline__ double fast_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// ln(g) for positive normal g from a 256-entry table in shared memory: g = 2^e * m, m in [1,2);
// entry i holds r_i ~ 1/(1 + (i+0.5)/256) and -ln(r_i), so m*r_i = 1 + d with |d| < 2^-9 (one exact fma) and
// ln g = e ln2 - ln r_i + log1p(d), log1p by a degree-5 Taylor polynomial (truncation d^6/6 < 1e-17).
// Absolute error ~1e-16, i.e. as accurate as libm's log here, at about half the instructions.
#define THETA_LOGTAB_SIZE 256
__device__ __forceinline__ double fast_log(double g, const double2* __restrict__ LT) {
    const long long bits = __double_as_longlong(g);
    const int hi = (int)(bits >> 32);
    // zero, negative, denormal, inf, NaN: the library path (one integer compare on the exponent word)
    if ((uint32_t)(hi - 0x00100000) >= 0x7fe00000u) return log(g);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 12) & 0xff;
    const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    const double2 t = LT[idx];  // shared memory in the search kernel, L1-resident global in the tile kernels
    const double d = fma(m, t.x, -1.0);
    double q = fma(d, 0.2, -0.25);
    q = fma(d, q, 1.0 / 3.0);
    q = fma(d, q, -0.5);
    const double r = fma(d * d, q, d);
    return fma((double)e, 0.6931471805599453094, t.y + r);
}

// 1/a to ~2^-44 relative: MUFU.RCP64H seed + one Newton-Raphson step.  Enough for the Newton iterates of the
// root-finder (an error eps in every term of f moves the root by about eps * x), not used for results.
__device__ __forceinline__ double fast_rcp1(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    const double e = fma(-a, r, 1.0);
    return fma(r, e, r);
}

// Evaluate f and f' at x in the open bracket (python/Optimizer.py:209-221 and its derivative).  Every
// denominator v y + z is positive there (z = x S1 > 0), so empty groups (nv == 0) simply add zero and the
// loop needs no guard.
template <int KV>
__device__ __forceinline__ void n2_f_df(const double (&nv)[KV], const double (&ev)[KV], double S1, double A, double x,
                                        double& f, double& fp) {
    const double z = x * S1;
    const double y = (1.0 - x) * A;
    double s = 0.0, sp = 0.0;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        const double den = fma((double)v, y, z);
        const double inv = fast_rcp1(den);
        const double t = nv[v] * inv;
        s += t;
        sp = fma(t * ev[v], inv, sp);
    }
    f = -s;
    fp = sp;
}

// f, f' and f''/2 at x in the open bracket:  f = -sum n_v/den_v,  f' = sum n_v e_v/den_v^2,  f''/2 = -sum n_v e_v^2/den_v^3
// (den_v = v (1-x) A + x S1,  d den_v/dx = e_v).
template <int KV>
__device__ __forceinline__ void n2_f_df_d2(const double (&nv)[KV], const double (&ev)[KV], double S1, double A, double x,
                                           double& f, double& fp, double& fpp2) {
    const double z = x * S1;
    const double y = (1.0 - x) * A;
    double s = 0.0, sp = 0.0, spp = 0.0;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        const double den = fma((double)v, y, z);
        const double inv = fast_rcp1(den);
        const double t = nv[v] * inv;
        const double ei = ev[v] * inv;
        const double u = t * ei;
        s += t;
        sp += u;
        spp = fma(u, ei, spp);
    }
    f = -s;
    fp = sp;
    fpp2 = -spp;
}

template <int KV>
__device__ __forceinline__ double n2_f(const double (&nv)[KV], double S1, double A, double x) {
    const double z = x * S1;
    const double y = (1.0 - x) * A;
    double s = 0.0;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        if (nv[v] != 0.0) s += nv[v] / fma((double)v, y, z);
    }
    return -s;
}

// One safeguarded Halley step of the root-finder on the bracket [lo, hi] (f increasing: f' = sum R_v e_v^2 / den^2 > 0).
// With c2 = f''/(2f') the error left after a *Newton* step dN is |c2| dN^2 to leading order (and less after the
// Halley step actually taken), so the stopping test needs no history: stop once that bound is below
// THETA_NEWTON_PRED * |x|  (the reference's own brenth stops at 2e-12 absolute, Optimizer.py:118).
// Returns true when x is final.
//
// P1 (pass 1 of the search, whose start is already good to ~1e-6): the bound for the Halley step itself.  With
// w_v = e_v/den_v, f''/(2f') and f'''/(6f') are weighted means of -w_v and w_v^2 (weights R_v e_v^2/den_v^2 >= 0), w_v is
// monotone in v between 1/x and -1/(1-x), hence |w_v| <= W = 1/min(x, 1-x), and the error left after a Halley step is
// |(f''/(2f'))^2 - f'''/(6f')| dN^3 <= 2 W^2 |dN|^3 to leading order.  The test  dN^2 <= 1e-9 x min(x,1-x)  and
// |dN| <= 1e-2 min(x,1-x)  implies  2 W^2 |dN|^3 <= THETA_NEWTON_PRED |x|: the same accuracy contract as the Newton bound,
// reached one evaluation earlier.
template <int KV, bool P1 = false>
__device__ __forceinline__ bool n2_halley_step(const double (&nv)[KV], const double (&ev)[KV], double S1, double A, bool rising,
                                               double& x, double& lo, double& hi) {
    double f, fp, fpp2;
    n2_f_df_d2<KV>(nv, ev, S1, A, x, f, fp, fpp2);
    if (f == 0.0) return true;
    if ((f < 0.0) == rising) lo = x; else hi = x;
    const double rfp = fast_rcp(fp);
    const double dN = -f * rfp;          // Newton step
    const double c2 = fpp2 * rfp;         // f'' / (2 f')
    const double den = fma(dN, c2, 1.0);  // Halley: dH = dN / (1 + dN c2)
    const double dx = (den > 0.5 && den < 2.0) ? dN * fast_rcp(den) : dN;
    const double xn = x + dx;
    const bool inside = (xn > lo && xn < hi);
    bool conv;
    if (P1) {
        const double mx = fmin(x, 1.0 - x);  // (P1: the bracket is [0, 1])
        conv = (dN * dN <= (50.0 * THETA_NEWTON_PRED) * x * mx) && (fabs(dN) <= 1e-2 * mx);
    } else {
        conv = fabs(c2) * dN * dN <= THETA_NEWTON_PRED * fabs(x);
    }
    // (a converged step can round onto the bracket end it started from: then x itself is the root)
    if (conv && (inside || fabs(dN) <= THETA_NEWTON_RTOL * fabs(x))) {
        if (inside) x = xn;
        return true;
    }
    x = inside ? xn : 0.5 * (lo + hi);
    return false;
}

// Optimizer._solve_n2 (python/Optimizer.py:90-126) on grouped statistics.
//   R[v]      : sum of r_i over intervals with c_i == v
//   S1        : sum rN_i c_i
//   zero_r_c0 : some interval with c_i == 0 has r_i == 0 (0/0 at x = lB = 0 -> NaN -> None)
//   x0        : warm start for the root (<= 0: none)
// n2_root: everything up to mu (brenth + M2); returns false when the reference returns None.
// LB: honour k.min_normal (Optimizer's lower_bound -> lB = M2_Rev(C_w, lower_bound), Optimizer.py:109-110).  RunTHetA
// never sets it, so the search kernels instantiate LB = false and keep lB = 0 hard-wired.
template <int KV, bool LB = false>
__device__ __forceinline__ bool n2_root(const double (&R)[KV], double S1, bool zero_r_c0, const N2Const& k, double x0,
                                        SolveOut& o) {
    o.x = 0.0;
    o.mu = 0.0;
    o.nll = 0.0;
    o.iters = 0;
    // normalize_C (Optimizer.py:167-174): a zero tumour column gives 0/0 = NaN; brenth's
    // NaN guard raises -> None.  Same for any rN_i == 0 (both normalised entries are 0).
    if (!(S1 > 0.0) || k.any_zero_rN || zero_r_c0) {
        o.status = THETA_ST_NONE_NAN;
        return false;
    }
    double ev[KV], nv[KV];
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        ev[v] = S1 - (double)v * k.A;  // exact
        nv[v] = R[v] * ev[v];
    }
    // bracket [lB, uB] = [0, 1] or [0, M2_Rev(max_normal)] (Optimizer.py:109-112, 228-231)
    double uB = 1.0, lB = 0.0;
    // f(lB = 0) up to the positive factor 1/A: the v = 0 group has denominator 0 -> -inf when it carries reads
    double f0;
    if (LB && k.min_normal != 0.0) {
        const double numer = -k.min_normal * k.S0;
        const double denom = (k.min_normal - 1.0) * S1 + numer;
        lB = numer / denom;
        f0 = n2_f<KV>(nv, S1, k.A, lB);
    } else if (nv[0] != 0.0) {
        f0 = -INFINITY;
    } else {
        double s = 0.0;
#pragma unroll
        for (int v = 1; v < KV; ++v) s = fma(nv[v], 1.0 / (double)v, s);  // compile-time reciprocals
        f0 = -s;
    }
    double fU;
    if (k.max_normal != 1.0) {
        const double numer = -k.max_normal * k.S0;
        const double denom = (k.max_normal - 1.0) * S1 + numer;
        uB = numer / denom;
        fU = n2_f<KV>(nv, S1, k.A, uB);
    } else {
        // at x = 1 every denominator equals S1: f(1) = -(sum_v n_v) / S1; only its sign / zero matter
        double s = 0.0;
#pragma unroll
        for (int v = 0; v < KV; ++v) s += nv[v];
        fU = -s;
    }
    if (f0 != f0 || fU != fU) {
        o.status = THETA_ST_NONE_NAN;
        return false;
    }
    double x;
    int status = THETA_ST_OK;
    if (f0 == 0.0) {  // scipy brenth: f(a)==0 -> return a (flat likelihood, Q17)
        x = lB;
        status = THETA_ST_ROOT_LB;
    } else if (fU == 0.0) {
        x = uB;
        status = THETA_ST_ROOT_UB;
    } else if ((f0 < 0.0) == (fU < 0.0)) {
        o.status = THETA_ST_NONE_SIGN;
        return false;
    } else {
        // f is increasing (f' = sum R_v e_v^2 / den^2 > 0): safeguarded Halley iteration inside the bracket.
        // With c2 = f''/(2f') the error left after a *Newton* step dN is |c2| dN^2 to leading order (and less
        // after the Halley step actually taken), so the stopping test needs no history: stop once that bound is
        // below THETA_NEWTON_PRED * |x|  (the reference's own brenth stops at 2e-12 absolute, Optimizer.py:118).
        // Cubic convergence makes two evaluations enough from a warm start within ~1e-2, which keeps the lanes
        // of a warp in step.
        double lo = lB, hi = uB;
        const bool rising = f0 < 0.0;
        x = (x0 > lo && x0 < hi) ? x0 : 0.5 * (lo + hi);
        int it = 0;
        for (; it < THETA_NEWTON_MAXIT; ++it)
            if (n2_halley_step<KV>(nv, ev, S1, k.A, rising, x, lo, hi)) break;
        o.iters = it + 1;
        if (it >= THETA_NEWTON_MAXIT) {
            o.status = THETA_ST_NONE_NOCONV;
            return false;
        }
    }
    // M2 (Optimizer.py:223-226), same operation order
    const double numer = -x * S1;
    const double denom = (x - 1.0) * k.S0 + numer;
    o.x = x;
    o.mu = numer / denom;
    o.status = status;
    return true;
}

// ---- pass 1 of the search only: a cheaper way to the same root -----------------------------------------------------
// The fp64 iteration above is what decides the root (bracket, stopping test, accuracy); what it costs is the number of
// evaluations, and from the warm start of the search (linear extrapolation of the previous two roots, error ~1e-3) it
// needs two.  n2_refine32 runs the same Halley iteration in fp32 first -- FMA pipe and MUFU.RCP, which the search
// kernel otherwise leaves idle next to the saturated FP64 pipe -- and hands the fp64 loop a start that is good to
// ~1e-6, so that its first evaluation already passes the stopping test.  The fp32 numbers never reach a result: a
// poor (or non-finite) refinement only costs fp64 evaluations.
// w-form of the same sums:  w_v = e_v/den_v,  f = -sum R_v w_v,  f' = sum R_v w_v^2,  f''/2 = -sum R_v w_v^3;
// nothing overflows for any input the plan accepts (R_v, A K < 2^57; w_v <= 1/x).
__device__ __forceinline__ float rcp32(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
template <int KV>
__device__ __forceinline__ float n2_refine32(const double (&R)[KV], double S1, double A, float x0, int nit, int* used = nullptr) {
    float Rf[KV], ef[KV];
    const float S1f = (float)S1, Af = (float)A;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        Rf[v] = (float)R[v];
        ef[v] = fmaf(-(float)v, Af, S1f);
    }
    float x = x0;
    // Iterate until the step d says the next iterate is good enough for the fp64 test of n2_root_p1 to pass at its first
    // evaluation (which needs an error below sqrt(1e-9 x mx), mx = min(x, 1-x)).  Halley leaves at most 2 d^3/mx^2 (see
    // n2_halley_step), so |d| <= 2e-2 mx is enough for every x; plain Newton (THETA_P1_F32HALLEY=0) leaves d^2/mx and
    // needs |d| <= 2e-3 mx.  What matters is the slowest lane of the warp: with Halley the predicted start of nearly
    // every column passes after ONE step (level changes in the sorted read-depth ratios and the first columns after an
    // odometer carry included), with Newton 11% of the columns -- i.e. nearly every warp step -- needed a second one.
    for (int it = 0; it < nit; ++it) {
        if (used) *used = it + 1;
        const float z = x * S1f;
        const float y = (1.0f - x) * Af;
        float s = 0.0f, sp = 0.0f, spp = 0.0f;
#pragma unroll
        for (int v = 0; v < KV; ++v) {
            const float w = ef[v] * rcp32(fmaf((float)v, y, z));
            const float a = Rf[v] * w;
            s += a;
            if (THETA_P1_F32HALLEY) {
                const float b = a * w;
                sp += b;
                spp = fmaf(b, w, spp);
            } else {
                sp = fmaf(a, w, sp);
            }
        }
        const float rfp = rcp32(sp);
        float d = s * rfp;                           // -f/f'
        const float mx = fminf(x, 1.0f - x);
        const bool last = fabsf(d) <= (THETA_P1_F32HALLEY ? 2e-2f : 2e-3f) * mx;
        if (THETA_P1_F32HALLEY) {
            const float den = fmaf(-d * spp, rfp, 1.0f);  // 1 + d f''/(2f'),  f''/2 = -spp
            if (den > 0.5f && den < 2.0f) d *= rcp32(den);
        }
        float xn = x + d;
        if (!(xn > 0.0f)) xn = 0.5f * x;             // (also catches NaN)
        if (!(xn < 1.0f)) xn = 0.5f * (x + 1.0f);
        x = xn;
        if (last) break;
    }
    return x;
}

// n2_root for pass 1 of the search (lB = 0, same classification, same fp64 iteration; fp32 start x0f, refined in fp32).
// The common case -- f(0) < 0 < f(1), bracket [0, 1] -- is handled inline; everything else (None candidates, roots at
// the bracket ends, max_normal != 1) goes through n2_root itself.  xf returns the root as a float (warm-start state of the
// caller, which keeps all of it in fp32).
template <int KV>
__device__ __forceinline__ bool n2_root_p1(const double (&R)[KV], double S1, bool zero_r_c0, const N2Const& k, float x0f,
                                           SolveOut& o, float& xf, int* used32 = nullptr) {
    double ev[KV], nv[KV];
    double s = 0.0;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        ev[v] = S1 - (double)v * k.A;  // exact
        nv[v] = R[v] * ev[v];
        s += nv[v];                    // f(1) = -s / S1 (same sum, same order as n2_root)
    }
    // f(0) < 0: -inf when the copy-number-0 group carries reads, else the finite sum of n2_root (the last column of every
    // odometer row has no interval at copy number 0)
    bool f0neg = nv[0] != 0.0;
    if (!f0neg) {
        double s0 = 0.0;
#pragma unroll
        for (int v = 1; v < KV; ++v) s0 = fma(nv[v], 1.0 / (double)v, s0);
        f0neg = -s0 < 0.0;
    }
    const bool common = (S1 > 0.0) && !zero_r_c0 && f0neg && (s < 0.0) && (k.max_normal == 1.0) && !k.any_zero_rN;
    if (!common) {
        const bool ok = n2_root<KV>(R, S1, zero_r_c0, k, (double)x0f, o);
        xf = (float)o.x;
        return ok;
    }
    const bool warm = (x0f > 0.0f && x0f < 1.0f);
    xf = n2_refine32<KV>(R, S1, k.A, warm ? x0f : 0.5f, 8, used32);
    double x = (double)xf;
    // first fp64 evaluation, nearly always the only one: the Halley step and the P1 test of n2_halley_step, with the
    // bracket [0, 1] and f rising folded in (when the test passes, 1 + dN f''/(2f') lies within [0.99, 1.01])
    int iters = 1;
    {
        double f, fp, fpp2;
        n2_f_df_d2<KV>(nv, ev, S1, k.A, x, f, fp, fpp2);
        const double rfp = fast_rcp1(fp);
        const double dN = -f * rfp;
        const double xn = fma(dN, fast_rcp1(fma(dN, fpp2 * rfp, 1.0)), x);
        const double mx = fmin(x, 1.0 - x);
        const bool conv = (dN * dN <= (50.0 * THETA_NEWTON_PRED) * x * mx) && (fabs(dN) <= 1e-2 * mx) && (xn > 0.0 && xn < 1.0);
        if (conv) {
            x = xn;
        } else {
            double lo = 0.0, hi = 1.0;
            int it = 0;
            for (; it < THETA_NEWTON_MAXIT; ++it)
                if (n2_halley_step<KV, true>(nv, ev, S1, k.A, true, x, lo, hi)) break;
            iters = it + 2;
            if (it >= THETA_NEWTON_MAXIT) {
                o.x = 0.0; o.mu = 0.0; o.nll = 0.0; o.iters = iters;
                o.status = THETA_ST_NONE_NOCONV;
                return false;
            }
            xf = (float)x;
        }
    }
    o.iters = iters;
    const double numer = -x * S1;
    const double denom = (x - 1.0) * k.S0 + numer;
    o.x = x;
    o.mu = THETA_P1_FASTM2 ? numer * fast_rcp(denom) : numer / denom;  // (pass 1 feeds thresholds only; records are re-solved)
    o.nll = 0.0;
    o.status = THETA_ST_OK;
    return true;
}

// L2 (Optimizer.py:187-196) on grouped statistics at mu.
template <int KV>
__device__ __forceinline__ double n2_l2(const double (&R)[KV], double S1, double mu, const N2Const& k,
                                        const double2* __restrict__ LT) {
    const double mu1 = 1.0 - mu;
    const double gbar = (mu * k.S0 + mu1 * S1) / k.A;
    const double lg = LT ? fast_log(gbar, LT) : log(gbar);
    double acc = 0.0;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        if (R[v] != 0.0) {
            const double g = fma((double)v, mu1, k.tau * mu);
            acc = fma(R[v], (LT ? fast_log(g, LT) : log(g)) - lg, acc);
        }
    }
    return k.K0 - acc;
}

// The same NLL with fp32 MUFU logarithms: 7 special-function instructions instead of ~110 fp64 ones.
// Valid when tau*mu >= 2^-7 (then every argument of log2 lies in [2^-7, 2^4) and |log2| < 8, so that
// __log2f's 2-ulp bound is <= 2^-20): the absolute error of the result is below  THETA_APPROX_C * Rtot
//   per logarithm: 2^-20 * ln2 (MUFU.LG2) + 3 * 2^-24 (fp32 argument)  <  8.5e-7,   total weight sum_v R_v + Rtot = 2 Rtot
// and THETA_APPROX_C = 4e-6 leaves a factor 2.3 of slack.  The search uses it only to prove that a candidate
// cannot matter (NLL far above the best seen); everything that can matter is evaluated by n2_l2.
#define THETA_APPROX_C 4e-6
#define THETA_APPROX_MIN_TM 0.0078125
__device__ __forceinline__ float lg2_ftz(float a) {  // one MUFU.LG2 (no denormal scaling: the arguments are >= 2^-7)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
template <int KV>
__device__ __forceinline__ double n2_l2_approx(const double (&R)[KV], double S1, double mu, const N2Const& k, double inv_A) {
    const double mu1 = 1.0 - mu;
    const double tm = k.tau * mu;
    const float mu1f = (float)mu1, tmf = (float)tm;
    const float lgb = lg2_ftz((float)((mu * k.S0 + mu1 * S1) * inv_A));
    double acc = 0.0;
#pragma unroll
    for (int v = 0; v < KV; ++v) {
        const float lg = lg2_ftz(fmaf((float)v, mu1f, tmf));
        acc = fma(R[v], (double)lg, acc);  // R[v] == 0 adds 0 (lg is finite: arguments >= 2^-7)
    }
    acc = fma(-k.Rtot, (double)lgb, acc);
    return fma(-0.6931471805599453094, acc, k.K0);
}

template <int KV, bool LB = false>
__device__ __forceinline__ void n2_solve(const double (&R)[KV], double S1, bool zero_r_c0, const N2Const& k, double x0,
                                         SolveOut& o, const double2* __restrict__ LT = nullptr) {
    if (n2_root<KV, LB>(R, S1, zero_r_c0, k, x0, o)) o.nll = n2_l2<KV>(R, S1, o.mu, k, LT);
}

}  // namespace theta

// ---- host-side helpers shared by the translation units ----
namespace theta {
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// Caching allocators (theta_rt.cu).  Blocks are size-bucketed and kept per device until
// theta_shutdown(), so a search call allocates nothing from the driver after its first use:
// cudaMalloc / cudaFree cost 0.1-0.5 ms each and cudaFree synchronises the whole device, which
// dominated the end-to-end latency of a 1.4 ms search.  Callers free only after the work that
// uses a block has been synchronised (every host-facing entry point ends with a stream sync).
cudaError_t dmalloc_bytes(void** p, size_t bytes);
void dfree(void* p);
template <typename T>
inline cudaError_t dmalloc(T** p, size_t bytes) { return dmalloc_bytes(reinterpret_cast<void**>(p), bytes); }
// pinned host staging blocks, same policy
cudaError_t hmalloc_bytes(void** p, size_t bytes);
void hfree(void* p);
template <typename T>
inline cudaError_t hmalloc(T** p, size_t bytes) { return hmalloc_bytes(reinterpret_cast<void**>(p), bytes); }
void cache_release();  // returns every cached block to the driver
// the fast_log table of the current device (uploaded on first use, kept until theta_shutdown)
cudaError_t logtab_device(double2** out);

// scope guards: blocks and events are returned on every exit path (THETA_CUDA returns early)
struct DevBlock {
    void* p = nullptr;
    ~DevBlock() { dfree(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
    cudaError_t alloc(size_t bytes) { return dmalloc_bytes(&p, bytes); }
};
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e == cudaSuccess ? cudaEventCreate(&b) : e; }
};
}  // namespace theta

#define THETA_CUDA(expr)                                                        \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) return theta::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
