// glibc_pow2.cuh -- pow(x, 2.0) with the rounding behaviour of the reference's libm.
//
// python/Optimizer.py:296-318 (`jacobian`/`second_deriv`) squares each denominator with `(...)**2` on a numpy
// float64, which is libm pow(x, 2.0).  glibc's pow (Szabolcs Nagy's algorithm from ARM optimized-routines:
// log in double-double from a 128-entry table, exp from a 128-entry table; sysdeps/ieee754/dbl-64/e_pow.c) is
// accurate to ~0.52 ulp, i.e. NOT always the correctly rounded x*x, and MINPACK amplifies that last bit into
// O(1) differences on the singular systems of rank-deficient candidates.  This file
// restates the computation with the same operations in the same order, including which multiply-adds are
// fused: the sequence was read off the x86-64 `__pow_fma` variant (the one the ifunc resolver picks on any CPU
// with FMA+AVX2) of Ubuntu GLIBC 2.39, the libm of the container the golden vectors were made in.  Tables:
// glibc_pow_tables.inc (tools/gen_glibc_pow_tables.py).
//
// Domain: finite x with 2^-360 <= |x| <= 2^360 take the table path; everything else (0, inf, NaN, and
// magnitudes whose square leaves glibc's main exp path) returns x*x, which is what pow gives for 0/inf/NaN and
// is unreachable for the denominators of the KKT system otherwise.
#pragma once
#include <stdint.h>
#include <string.h>

#include "strict_fp.cuh"

namespace theta {
namespace glibc {

#ifdef __CUDACC__
#define THETA_POWTAB static __device__ __constant__ const
#else
#define THETA_POWTAB static const
#endif
#ifdef __CUDACC__
// host copies for the host build of the same header (nvcc compiles both sides of a __host__ __device__ function)
namespace dev {
#include "glibc_pow_tables.inc"
}
#undef THETA_POWTAB
#define THETA_POWTAB static const
namespace host {
#include "glibc_pow_tables.inc"
}
#else
namespace host {
#include "glibc_pow_tables.inc"
}
#endif
#undef THETA_POWTAB

#ifdef __CUDA_ARCH__
#define THETA_PT(x) dev::x
#else
#define THETA_PT(x) host::x
#endif

THETA_HD uint64_t as_u64(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
THETA_HD double as_f64(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

// every +, -, * below is a separately rounded IEEE operation (xf64); fma_rn marks the fused ones
THETA_HDN double pow2(double x) {
    const double ax = ::fabs(x);
    if (!(ax >= 0x1p-360 && ax <= 0x1p360)) return (xf64(x) * xf64(x)).v;
    const uint64_t ix = as_u64(ax);  // y = 2 is an even integer: pow(-x, 2) == pow(x, 2), sign_bias = 0
    // ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1) as hi + lo
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double z = as_f64(iz);
    const double kd = (double)k;
    const double invc = THETA_PT(kPowLogTab)[i][0], logc = THETA_PT(kPowLogTab)[i][1], logctail = THETA_PT(kPowLogTab)[i][2];
    const double* A = THETA_PT(kPowA);
    const double t1 = fma_rn(kd, THETA_PT(kPowLn2hi), logc);
    const double lo1 = fma_rn(kd, THETA_PT(kPowLn2lo), logctail);
    const double r = fma_rn(z, invc, -1.0);
    const xf64 R(r);
    const xf64 ar = R * xf64(A[0]);
    const double q1 = fma_rn(r, A[2], A[1]);
    const double q2 = fma_rn(r, A[4], A[3]);
    const xf64 t2 = R + xf64(t1);
    const xf64 lo2 = (xf64(t1) - t2) + R;
    const xf64 ar2 = R * ar;
    const xf64 ar3 = R * ar2;
    const double lo3 = fma_rn(ar.v, r, -ar2.v);
    const xf64 hi = t2 + ar2;
    const double q3 = fma_rn(r, A[6], A[5]);
    const double q4 = fma_rn(q3, ar2.v, q2);
    const xf64 lo4 = (t2 - hi) + ar2;
    const double q5 = fma_rn(ar2.v, q4, q1);
    const xf64 s = ((xf64(lo1) + lo2) + xf64(lo3)) + lo4;
    const double lo = fma_rn(ar3.v, q5, s.v);
    const xf64 lhi = hi + xf64(lo);
    const xf64 llo = (hi - lhi) + xf64(lo);
    // ---- pow: ehi + elo = y * (lhi + llo), y = 2
    const double y = 2.0;
    const xf64 ehi = xf64(y) * lhi;
    const double elo = fma_rn(y, llo.v, fma_rn(lhi.v, y, -ehi.v));
    const unsigned abstop = (unsigned)(as_u64(ehi.v) >> 52) & 0x7ff;
    if (abstop - 0x3c9u > 0x3eu) return (xf64(x) * xf64(x)).v;  // |ehi| < 2^-54 (x == 1) or >= 512: not on the main path
    // ---- exp_inline(ehi, elo)
    const double kz = fma_rn(ehi.v, THETA_PT(kExpInvLn2N), THETA_PT(kExpShift));
    const uint64_t ki = as_u64(kz);
    const xf64 kdd = xf64(kz) - xf64(THETA_PT(kExpShift));
    const double r0 = fma_rn(kdd.v, THETA_PT(kExpNegLn2hiN), ehi.v);
    const double r1 = fma_rn(kdd.v, THETA_PT(kExpNegLn2loN), r0);
    const xf64 rr = xf64(elo) + xf64(r1);
    const unsigned idx = 2u * (unsigned)(ki & 127);
    const uint64_t sbits = THETA_PT(kExpTab)[idx + 1] + (ki << 45);
    const double tail = as_f64(THETA_PT(kExpTab)[idx]);
    const double* C = THETA_PT(kExpC);  // C2..C5
    const double p23 = fma_rn(rr.v, C[1], C[0]);
    const xf64 rt = rr + xf64(tail);
    const xf64 r2 = rr * rr;
    const double p45 = fma_rn(rr.v, C[3], C[2]);
    const double u = fma_rn(p23, r2.v, rt.v);
    const xf64 r4 = r2 * r2;
    const double w = fma_rn(p45, r4.v, u);
    const double scale = as_f64(sbits);
    return fma_rn(w, scale, scale);
}

#undef THETA_PT

}  // namespace glibc
}  // namespace theta
