// strict_fp.cuh -- IEEE-754 double arithmetic that the compiler may not contract into fused multiply-adds.
//
// The n=3 reference-order solver (n3_ref.cuh) has to reproduce, bit for bit, arithmetic that the reference
// executes on the CPU without FMA contraction (CPython floats, numpy element-wise ops, SciPy's MINPACK build --
// `objdump -d _minpack*.so | grep -c vfmadd` = 0).  nvcc contracts a*b+c by default (-fmad=true), so that code
// is written on `xf64`, whose operators map to the never-fused intrinsics __dmul_rn / __dadd_rn / __dsub_rn /
// __ddiv_rn / __dsqrt_rn.  On the host the operators are plain IEEE operations (x86-64 has no FMA in its
// baseline ISA; the test harnesses also pass -ffp-contract=off).
#pragma once
#include <math.h>

#include "hybrj4.cuh"  // THETA_HD

namespace theta {

// keep the <math.h> overloads visible next to the xf64 ones for unqualified calls inside namespace theta
using ::fabs;
using ::fmax;
using ::fmin;
using ::sqrt;

struct xf64 {
    double v;
    xf64() = default;
    THETA_HD xf64(double x) : v(x) {}
};

THETA_HD xf64 operator+(xf64 a, xf64 b) {
#ifdef __CUDA_ARCH__
    return xf64(__dadd_rn(a.v, b.v));
#else
    return xf64(a.v + b.v);
#endif
}
THETA_HD xf64 operator-(xf64 a, xf64 b) {
#ifdef __CUDA_ARCH__
    return xf64(__dsub_rn(a.v, b.v));
#else
    return xf64(a.v - b.v);
#endif
}
THETA_HD xf64 operator*(xf64 a, xf64 b) {
#ifdef __CUDA_ARCH__
    return xf64(__dmul_rn(a.v, b.v));
#else
    return xf64(a.v * b.v);
#endif
}
THETA_HD xf64 operator/(xf64 a, xf64 b) {
#ifdef __CUDA_ARCH__
    return xf64(__ddiv_rn(a.v, b.v));
#else
    return xf64(a.v / b.v);
#endif
}
THETA_HD xf64 operator-(xf64 a) { return xf64(-a.v); }
THETA_HD xf64& operator+=(xf64& a, xf64 b) { a = a + b; return a; }
THETA_HD xf64& operator-=(xf64& a, xf64 b) { a = a - b; return a; }
THETA_HD xf64& operator*=(xf64& a, xf64 b) { a = a * b; return a; }
THETA_HD xf64& operator/=(xf64& a, xf64 b) { a = a / b; return a; }
THETA_HD bool operator<(xf64 a, xf64 b) { return a.v < b.v; }
THETA_HD bool operator>(xf64 a, xf64 b) { return a.v > b.v; }
THETA_HD bool operator<=(xf64 a, xf64 b) { return a.v <= b.v; }
THETA_HD bool operator>=(xf64 a, xf64 b) { return a.v >= b.v; }
THETA_HD bool operator==(xf64 a, xf64 b) { return a.v == b.v; }
THETA_HD bool operator!=(xf64 a, xf64 b) { return a.v != b.v; }
THETA_HD xf64 fabs(xf64 a) { return xf64(::fabs(a.v)); }
THETA_HD xf64 fmax(xf64 a, xf64 b) { return xf64(::fmax(a.v, b.v)); }  // C99: the non-NaN operand wins
THETA_HD xf64 fmin(xf64 a, xf64 b) { return xf64(::fmin(a.v, b.v)); }
THETA_HD xf64 sqrt(xf64 a) {
#ifdef __CUDA_ARCH__
    return xf64(__dsqrt_rn(a.v));
#else
    return xf64(::sqrt(a.v));
#endif
}
// fused multiply-add where the reference's libm executes one (glibc's __pow_fma)
THETA_HD double fma_rn(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return ::fma(a, b, c);
#endif
}

}  // namespace theta
