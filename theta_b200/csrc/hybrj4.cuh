// hybrj4.cuh -- MINPACK's hybrj (Powell hybrid method with an analytic Jacobian), restated for
// the fixed size N = 4 that Optimizer._solve_n3plus needs (unknowns mu_hat_0..2 and the Lagrange
// multiplier; python/Optimizer.py:147-149 calls scipy.optimize.fsolve(..., fprime=jacobian), i.e.
// MINPACK hybrj with xtol = 1.49012e-8, maxfev = 100*(N+1), factor = 100, mode = 1).
//
// The reference never checks fsolve's status and stops at a loose xtol, so mu parity at 1e-9 needs
// the same iterate sequence, not just the same fixed point.  This file follows the
// published algorithm (More, Garbow, Hillstrom: "User Guide for MINPACK-1", ANL-80-74; routines
// hybrj, qrfac, qform, dogleg, r1updt, r1mpyq, enorm) step by step, including its storage
// conventions (R packed by rows, Givens rotations encoded in a single number).
// Everything is __host__ __device__ so the same code is unit-tested on the CPU against SciPy
// (tests/test_n3_solver_host.py builds host harnesses from tests/harness/) and runs inside the CUDA kernels.
#pragma once
#include <math.h>

#ifndef __CUDACC__
#define THETA_HD inline
#define THETA_HDN inline
#define THETA_MPI inline
#else
#define THETA_HD __host__ __device__ __forceinline__
// the solver is large: out-of-line routines keep the kernels within the instruction cache ...
#define THETA_HDN __host__ __device__ __noinline__
// ... except the MINPACK routines themselves (THETA_MPI): inlined and unrolled at N = 4 their 4-vectors become
// registers and their loops straight-line code; out of line (-DTHETA_MP_OUTLINE) every vector lives in local
// memory and every loop runs with run-time indices
#ifdef THETA_MP_OUTLINE
#define THETA_MPI __host__ __device__ __noinline__
#else
#define THETA_MPI __host__ __device__ __forceinline__
#endif
#endif

namespace theta {
namespace mp {

constexpr int N = 4;
constexpr int LR = N * (N + 1) / 2;
constexpr double EPSMCH = 2.220446049250313e-16;   // dpmpar(1)
constexpr double GIANT = 1.7976931348623157e308;   // dpmpar(3)

template <class T>
THETA_HDN T enorm(int n, const T* x) {
    const T rdwarf = 3.834e-20, rgiant = 1.304e19;
    T s1 = 0.0, s2 = 0.0, s3 = 0.0, x1max = 0.0, x3max = 0.0;
    const T agiant = rgiant / (T)n;
#pragma unroll
    for (int i = 0; i < n; ++i) {
        const T xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;
        } else if (xabs <= rdwarf) {
            if (xabs > x3max) {
                const T t = x3max / xabs;
                s3 = 1.0 + s3 * (t * t);
                x3max = xabs;
            } else if (xabs != 0.0) {
                const T t = xabs / x3max;
                s3 += t * t;
            }
        } else {
            if (xabs > x1max) {
                const T t = x1max / xabs;
                s1 = 1.0 + s1 * (t * t);
                x1max = xabs;
            } else {
                const T t = xabs / x1max;
                s1 += t * t;
            }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// a is column-major N x N: a[i + N*j]
template <class T>
THETA_MPI void qrfac(T* a, T* rdiag, T* acnorm) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        acnorm[j] = enorm(N, a + N * j);
        rdiag[j] = acnorm[j];
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        T ajnorm = enorm(N - j, a + j + N * j);
        if (ajnorm != 0.0) {
            if (a[j + N * j] < 0.0) ajnorm = -ajnorm;
#pragma unroll
            for (int i = j; i < N; ++i) a[i + N * j] /= ajnorm;
            a[j + N * j] += 1.0;
#pragma unroll
            for (int k = j + 1; k < N; ++k) {
                T sum = 0.0;
#pragma unroll
                for (int i = j; i < N; ++i) sum += a[i + N * j] * a[i + N * k];
                const T temp = sum / a[j + N * j];
#pragma unroll
                for (int i = j; i < N; ++i) a[i + N * k] -= temp * a[i + N * j];
            }
        }
        rdiag[j] = -ajnorm;
    }
}

template <class T>
THETA_MPI void qform(T* q, T* wa) {
#pragma unroll
    for (int j = 1; j < N; ++j)
#pragma unroll
        for (int i = 0; i < j; ++i) q[i + N * j] = 0.0;
#pragma unroll
    for (int l = 0; l < N; ++l) {
        const int k = N - 1 - l;
#pragma unroll
        for (int i = k; i < N; ++i) {
            wa[i] = q[i + N * k];
            q[i + N * k] = 0.0;
        }
        q[k + N * k] = 1.0;
        if (wa[k] != 0.0) {
#pragma unroll
            for (int j = k; j < N; ++j) {
                T sum = 0.0;
#pragma unroll
                for (int i = k; i < N; ++i) sum += q[i + N * j] * wa[i];
                const T temp = sum / wa[k];
#pragma unroll
                for (int i = k; i < N; ++i) q[i + N * j] -= temp * wa[i];
            }
        }
    }
}

// r: upper triangle packed by rows (0-based l); x: output step
template <class T>
THETA_MPI void dogleg(const T* r, const T* diag, const T* qtb, T delta, T* x, T* wa1,
                     T* wa2) {
    // Gauss-Newton direction
    int jj = LR;  // 0-based index one past the last diagonal (1-based jj = LR + 1)
#pragma unroll
    for (int k = 1; k <= N; ++k) {
        const int j = N - k;  // 0-based
        jj -= k;              // now the 0-based index of r(j,j)
        int l = jj + 1;
        T sum = 0.0;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            sum += r[l] * x[i];
            ++l;
        }
        T temp = r[jj];
        if (temp == 0.0) {
            l = j;  // 0-based index of r(0,j)
#pragma unroll
            for (int i = 0; i <= j; ++i) {
                temp = fmax(temp, fabs(r[l]));
                l += N - 1 - i;
            }
            temp = EPSMCH * temp;
            if (temp == 0.0) temp = EPSMCH;
        }
        x[j] = (qtb[j] - sum) / temp;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        wa1[j] = 0.0;
        wa2[j] = diag[j] * x[j];
    }
    const T qnorm = enorm(N, wa2);
    if (qnorm <= delta) return;
    // scaled gradient direction
    int l = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const T temp = qtb[j];
#pragma unroll
        for (int i = j; i < N; ++i) {
            wa1[i] += r[l] * temp;
            ++l;
        }
        wa1[j] = wa1[j] / diag[j];
    }
    const T gnorm = enorm(N, wa1);
    T sgnorm = 0.0;
    T alpha = delta / qnorm;
    if (gnorm != 0.0) {
#pragma unroll
        for (int j = 0; j < N; ++j) wa1[j] = (wa1[j] / gnorm) / diag[j];
        l = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            T sum = 0.0;
#pragma unroll
            for (int i = j; i < N; ++i) {
                sum += r[l] * wa1[i];
                ++l;
            }
            wa2[j] = sum;
        }
        T temp = enorm(N, wa2);
        sgnorm = (gnorm / temp) / temp;
        alpha = 0.0;
        if (sgnorm < delta) {
            const T bnorm = enorm(N, qtb);
            temp = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
            const T dq = delta / qnorm, sd = sgnorm / delta;
            temp = temp - dq * (sd * sd) + sqrt((temp - dq) * (temp - dq) + (1.0 - dq * dq) * (1.0 - sd * sd));
            alpha = (dq * (1.0 - sd * sd)) / temp;
        }
    }
    const T temp = (1.0 - alpha) * fmin(sgnorm, delta);
#pragma unroll
    for (int j = 0; j < N; ++j) x[j] = temp * wa1[j] + alpha * x[j];
}

// s: lower triangular stored by columns (== R packed by rows); m = n = N
template <class T>
THETA_MPI void r1updt(T* s, const T* u, T* v, T* w, bool* sing) {
    const T p5 = 0.5, p25 = 0.25;
    int jj = LR - 1;  // 0-based index of s(N-1,N-1)
    w[N - 1] = s[jj];
#pragma unroll
    for (int nmj = 1; nmj <= N - 1; ++nmj) {
        const int j = N - 1 - nmj;  // 0-based
        jj -= (N - j);              // 0-based index of s(j,j)
        w[j] = 0.0;
        if (v[j] != 0.0) {
            T c, sn, tau;
            if (fabs(v[N - 1]) < fabs(v[j])) {
                const T cotan = v[N - 1] / v[j];
                sn = p5 / sqrt(p25 + p25 * (cotan * cotan));
                c = sn * cotan;
                tau = 1.0;
                if (fabs(c) * GIANT > 1.0) tau = 1.0 / c;
            } else {
                const T tn = v[j] / v[N - 1];
                c = p5 / sqrt(p25 + p25 * (tn * tn));
                sn = c * tn;
                tau = sn;
            }
            v[N - 1] = sn * v[j] + c * v[N - 1];
            v[j] = tau;
            int l = jj;
#pragma unroll
            for (int i = j; i < N; ++i) {
                const T temp = c * s[l] - sn * w[i];
                w[i] = sn * s[l] + c * w[i];
                s[l] = temp;
                ++l;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] += v[N - 1] * u[i];
    *sing = false;
#pragma unroll
    for (int j = 0; j < N - 1; ++j) {
        if (w[j] != 0.0) {
            T c, sn, tau;
            if (fabs(s[jj]) < fabs(w[j])) {
                const T cotan = s[jj] / w[j];
                sn = p5 / sqrt(p25 + p25 * (cotan * cotan));
                c = sn * cotan;
                tau = 1.0;
                if (fabs(c) * GIANT > 1.0) tau = 1.0 / c;
            } else {
                const T tn = w[j] / s[jj];
                c = p5 / sqrt(p25 + p25 * (tn * tn));
                sn = c * tn;
                tau = sn;
            }
            int l = jj;
#pragma unroll
            for (int i = j; i < N; ++i) {
                const T temp = c * s[l] + sn * w[i];
                w[i] = -sn * s[l] + c * w[i];
                s[l] = temp;
                ++l;
            }
            w[j] = tau;
        }
        if (s[jj] == 0.0) *sing = true;
        jj += (N - j);
    }
    s[jj] = w[N - 1];
    if (s[jj] == 0.0) *sing = true;
}

// a: m x N column-major with leading dimension lda
template <class T>
THETA_MPI void r1mpyq(int m, T* a, int lda, const T* v, const T* w) {
#pragma unroll
    for (int nmj = 1; nmj <= N - 1; ++nmj) {
        const int j = N - 1 - nmj;
        T c, sn;
        if (fabs(v[j]) > 1.0) {
            c = 1.0 / v[j];
            sn = sqrt(1.0 - c * c);
        } else {
            sn = v[j];
            c = sqrt(1.0 - sn * sn);
        }
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const T temp = c * a[i + lda * j] - sn * a[i + lda * (N - 1)];
            a[i + lda * (N - 1)] = sn * a[i + lda * j] + c * a[i + lda * (N - 1)];
            a[i + lda * j] = temp;
        }
    }
#pragma unroll
    for (int j = 0; j < N - 1; ++j) {
        T c, sn;
        if (fabs(w[j]) > 1.0) {
            c = 1.0 / w[j];
            sn = sqrt(1.0 - c * c);
        } else {
            sn = w[j];
            c = sqrt(1.0 - sn * sn);
        }
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const T temp = c * a[i + lda * j] + sn * a[i + lda * (N - 1)];
            a[i + lda * (N - 1)] = -sn * a[i + lda * j] + c * a[i + lda * (N - 1)];
            a[i + lda * j] = temp;
        }
    }
}

struct Result {
    int info, nfev, njev;
};

// Sys must provide:  void fcn(const T* x, T* fvec);  void jac(const T* x, T* fjac /*col-major*/);
template <class T, class Sys>
THETA_MPI Result hybrj(Sys& sys, T* x, T xtol, int maxfev, T factor) {
    const T p1 = 0.1, p5 = 0.5, p001 = 1e-3, p0001 = 1e-4;
    T fvec[N], fjac[N * N], diag[N], r[LR], qtf[N], wa1[N], wa2[N], wa3[N], wa4[N];
    Result res{0, 0, 0};
    sys.fcn(x, fvec);
    res.nfev = 1;
    T fnorm = enorm(N, fvec);
    int iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
    T delta = 0.0, xnorm = 0.0;
    for (;;) {  // outer loop: (re)compute the Jacobian
        bool jeval = true;
        res.nfev += sys.jac(x, fvec, fjac);
        ++res.njev;
        qrfac(fjac, wa1, wa2);
        if (iter == 1) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                diag[j] = wa2[j];
                if (wa2[j] == 0.0) diag[j] = 1.0;
            }
#pragma unroll
            for (int j = 0; j < N; ++j) wa3[j] = diag[j] * x[j];
            xnorm = enorm(N, wa3);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) qtf[i] = fvec[i];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (fjac[j + N * j] != 0.0) {
                T sum = 0.0;
#pragma unroll
                for (int i = j; i < N; ++i) sum += fjac[i + N * j] * qtf[i];
                const T temp = -sum / fjac[j + N * j];
#pragma unroll
                for (int i = j; i < N; ++i) qtf[i] += fjac[i + N * j] * temp;
            }
        }
        bool sing = false;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            int l = j;
#pragma unroll
            for (int i = 0; i < j; ++i) {
                r[l] = fjac[i + N * j];
                l += N - 1 - i;
            }
            r[l] = wa1[j];
            if (wa1[j] == 0.0) sing = true;
        }
        qform(fjac, wa1);
#pragma unroll
        for (int j = 0; j < N; ++j) diag[j] = fmax(diag[j], wa2[j]);
        for (;;) {  // inner loop
            dogleg(r, diag, qtf, delta, wa1, wa2, wa3);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const T pnorm = enorm(N, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            sys.fcn(wa2, wa4);
            ++res.nfev;
            const T fnorm1 = enorm(N, wa4);
            T actred = -1.0;
            if (fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            int l = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                T sum = 0.0;
#pragma unroll
                for (int j = i; j < N; ++j) {
                    sum += r[l] * wa1[j];
                    ++l;
                }
                wa3[i] = qtf[i] + sum;
            }
            const T temp = enorm(N, wa3);
            T prered = 0.0;
            if (temp < fnorm) prered = 1.0 - (temp / fnorm) * (temp / fnorm);
            T ratio = 0.0;
            if (prered > 0.0) ratio = actred / prered;
            if (ratio < p1) {
                ncsuc = 0;
                ++ncfail;
                delta = p5 * delta;
            } else {
                ncfail = 0;
                ++ncsuc;
                if (ratio >= p5 || ncsuc > 1) delta = fmax(delta, pnorm / p5);
                if (fabs(ratio - 1.0) <= p1) delta = pnorm / p5;
            }
            if (ratio >= p0001) {
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    x[j] = wa2[j];
                    wa2[j] = diag[j] * x[j];
                    fvec[j] = wa4[j];
                }
                xnorm = enorm(N, wa2);
                fnorm = fnorm1;
                ++iter;
            }
            ++nslow1;
            if (actred >= p001) nslow1 = 0;
            if (jeval) ++nslow2;
            if (actred >= p1) nslow2 = 0;
            if (delta <= xtol * xnorm || fnorm == 0.0) res.info = 1;
            if (res.info != 0) return res;
            if (res.nfev >= maxfev) res.info = 2;
            if (p1 * fmax(p1 * delta, pnorm) <= EPSMCH * xnorm) res.info = 3;
            if (nslow2 == 5) res.info = 4;
            if (nslow1 == 10) res.info = 5;
            if (res.info != 0) return res;
            if (ncfail == 2) break;  // recompute the Jacobian
#pragma unroll
            for (int j = 0; j < N; ++j) {
                T sum = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) sum += fjac[i + N * j] * wa4[i];
                wa2[j] = (sum - wa3[j]) / pnorm;
                wa1[j] = diag[j] * ((diag[j] * wa1[j]) / pnorm);
                if (ratio >= p0001) qtf[j] = sum;
            }
            r1updt(r, wa1, wa2, wa3, &sing);
            r1mpyq(N, fjac, N, wa2, wa3);
            r1mpyq(1, qtf, 1, wa2, wa3);
            jeval = false;
        }
    }
}

}  // namespace mp
}  // namespace theta
