// n3_ref.cuh -- Optimizer._solve_n3plus (python/Optimizer.py:128-165) in the REFERENCE'S OWN ARITHMETIC ORDER.
//
// For a rank-deficient candidate -- the distinct rows (tau, a, b) that carry reads lie on one line, or a tumour
// column is all zero -- the mixture is not identifiable: the KKT Jacobian (Optimizer.py:296-318) is singular at
// every point, MINPACK's hybrj walks on rounding noise, and what the reference returns (the in-range decision,
// the BFGS fallback, mu, and through a non-converged mu_hat even the NLL, off by up to 11 % from the optimum)
// is a deterministic function of every rounding on the way, not of the mathematics.
// The grouped-statistics solver (theta_n3.cuh) evaluates the same sums in a different order and so takes a
// different walk.  Those candidates are therefore solved here exactly as the reference does it, operation by
// operation, on xf64 (strict_fp.cuh: no FMA contraction):
//
//   weighted_C / normalize_C   Optimizer.py:167-182   C_w = rN (.) C, C_hat = C_w / column sums (0/0 -> NaN)
//   equations / jacobian       Optimizer.py:267-318   per-interval left-to-right sums; d**2 is libm pow (glibc_pow2.cuh)
//   scipy fsolve(fprime=...)   MINPACK hybrj          hybrj4.cuh on xf64 (bit-identical to SciPy: tests/test_n3_solver_host.py)
//   Misc.inRange               Misc.py:49-57          NaN passes
//   fmin_bfgs(L3_hat, dL3_hat) Optimizer.py:151-161, 246-265   bfgs_ls.cuh on per-interval sums
//   M3 = fsolve(M_eq)          Optimizer.py:320-330   MINPACK hybrd (forward-difference Jacobian, fdjac1) -- NOT the
//                              closed form: with a zero column it returns round-off garbage that L3 then turns into the NLL
//   L3                         Optimizer.py:236-244   per-interval, denominator summed column by column
//
// Cost: O(m) per function evaluation, ~15x the grouped solver -- taken by the search kernels only for
// rank-deficient leaves that survive screening (none on real-shaped data).
#pragma once
#include "bfgs_ls.cuh"
#include "glibc_pow2.cuh"
#include "hybrj4.cuh"
#include "strict_fp.cuh"

namespace theta {
namespace ref3 {

struct Ctx {
    int m;
    const double* r;     // [m] tumour reads
    const double* rN;    // [m] normal reads
    const uint8_t* row;  // row index of interval i at row[i * rstride]
    int rstride;
    const uint8_t* ra;   // row tables (Enumerator._create_graph order)
    const uint8_t* rb;
    double tau;
    double S0, S1, S2;   // column sums of C_w: exact integers (tau * sum rN, sum rN a, sum rN b)

    // C_hat[i][:] (normalize_C of weighted_C); a zero column sum gives 0/0 = NaN like numpy
    THETA_HD void chat(int i, xf64& c0, xf64& c1, xf64& c2) const {
        const int g = row[(size_t)i * rstride];
        const xf64 w = rN[i];
        c0 = (w * xf64(tau)) / xf64(S0);
        c1 = (w * xf64((double)ra[g])) / xf64(S1);
        c2 = (w * xf64((double)rb[g])) / xf64(S2);
    }
};

// How the per-interval sums are executed.  Every sum of the reference runs left to right over the intervals, and that
// order is part of the result; the terms themselves (divisions, pow, log) are independent.
//   SerialExec  one thread does everything (host tests, thread-per-candidate kernels).
//   WarpExec    (device) the 32 lanes of a warp compute the terms of their intervals into shared memory, then every
//               lane adds them up in interval order -- the same roundings, ~30x less latency; all lanes hold identical
//               values afterwards, so the scalar MINPACK / line-search code around the sums stays warp-uniform.
// term(i, t) writes the K addends of interval i.
struct SerialExec {
    template <int K, class F>
    THETA_HD void sums(int n, F& term, xf64* out) {
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = 0.0;
        for (int i = 0; i < n; ++i) {
            xf64 t[K];
            term(i, t);
#pragma unroll
            for (int k = 0; k < K; ++k) out[k] += t[k];
        }
    }
};
#ifdef __CUDACC__
struct WarpExec {
    double* scratch;  // [9][n_max] doubles of shared memory owned by this warp
    int stride;       // n_max
    template <int K, class F>
    __host__ __device__ __forceinline__ void sums(int n, F& term, xf64* out) {
#ifdef __CUDA_ARCH__
        const int lane = threadIdx.x & 31;
        for (int i = lane; i < n; i += 32) {
            xf64 t[K];
            term(i, t);
#pragma unroll
            for (int k = 0; k < K; ++k) scratch[k * stride + i] = t[k].v;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = 0.0;
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int k = 0; k < K; ++k) out[k] += xf64(scratch[k * stride + i]);
        }
        __syncwarp();
#endif
    }
};
#endif

// equations / jacobian (Optimizer.py:267-318) for fsolve
struct EqTerm {
    const Ctx* c;
    const xf64* x;
    THETA_HD void operator()(int i, xf64* t) const {
        xf64 c0, c1, c2;
        c->chat(i, c0, c1, c2);
        const xf64 d = (c0 * x[0] + c1 * x[1]) + c2 * x[2];  // sum([C_hat[i][j]*mu[j] ...]): 0 + t0 + t1 + t2
        const xf64 ri = c->r[i];
        t[0] = (ri * c0) / d;  // dLambda_dMu_numers[k][i] / d
        t[1] = (ri * c1) / d;
        t[2] = (ri * c2) / d;
    }
};
struct JacTerm {
    const Ctx* c;
    const xf64* x;
    THETA_HD void operator()(int i, xf64* t) const {
        xf64 ch[3];
        c->chat(i, ch[0], ch[1], ch[2]);
        const xf64 d = (ch[0] * x[0] + ch[1] * x[1]) + ch[2] * x[2];
        const xf64 d2 = glibc::pow2(d.v);  // (...)**2 on a numpy float64
        const xf64 ri = c->r[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const xf64 rk = ri * ch[k];
#pragma unroll
            for (int h = 0; h < 3; ++h) t[3 * k + h] = (rk * ch[h]) / d2;
        }
    }
};
template <class Exec>
struct Kkt {
    const Ctx* c;
    Exec* ex;
    THETA_HDN void fcn(const xf64* x, xf64* f) {
        xf64 s[3];
        EqTerm term{c, x};
        ex->template sums<3>(c->m, term, s);
        f[0] = (-s[0]) - x[3];
        f[1] = (-s[1]) - x[3];
        f[2] = (-s[2]) - x[3];
        f[3] = xf64(1.0) - ((x[0] + x[1]) + x[2]);
    }
    THETA_HDN int jac(const xf64* x, const xf64*, xf64* J) {  // column-major J[i + 4 j] = d f_i / d x_j
        xf64 s[9];
        JacTerm term{c, x};
        ex->template sums<9>(c->m, term, s);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int h = 0; h < 3; ++h) J[k + 4 * h] = s[3 * k + h];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            J[3 + 4 * i] = -1.0;
            J[i + 4 * 3] = -1.0;
        }
        J[3 + 4 * 3] = 0.0;
        return 0;
    }
};

// L3_hat and dL3_hat (Optimizer.py:246-265, sign error kept: Q7) along x0 + alpha pk, per interval
struct LhatTerm {
    const Ctx* c;
    xf64 M0, M1, M2;
    THETA_HD void operator()(int i, xf64* t) const {
        xf64 c0, c1, c2;
        c->chat(i, c0, c1, c2);
        const xf64 d = (c0 * M0 + c1 * M1) + c2 * M2;
        t[0] = xf64(c->r[i]) * xf64(log(d.v));  // log of a negative number is NaN, as in numpy
    }
};
struct DLhatTerm {
    const Ctx* c;
    xf64 M0, M1;
    THETA_HD void operator()(int i, xf64* t) const {
        xf64 c0, c1, c2;
        c->chat(i, c0, c1, c2);
        const xf64 n0 = c0 - c2, n1 = c1 - c2;
        const xf64 den = (n0 * M0 + n1 * M1) + c2;
        const xf64 ri = c->r[i];
        t[0] = ri * (n0 / den);
        t[1] = ri * (n1 / den);
    }
};
template <class Exec>
struct Ray {
    const Ctx* c;
    Exec* ex;
    double x0[2], pk[2];
    THETA_HDN double L3_hat(double m0, double m1) const {
        const xf64 M0 = m0, M1 = m1;
        LhatTerm term{c, M0, M1, xf64(1.0) - (M0 + M1)};
        xf64 tot[1];
        ex->template sums<1>(c->m, term, tot);
        return -tot[0].v;
    }
    THETA_HDN void dL3_hat(double m0, double m1, double* out) const {
        DLhatTerm term{c, xf64(m0), xf64(m1)};
        xf64 v[2];
        ex->template sums<2>(c->m, term, v);
        out[0] = v[0].v;
        out[1] = v[1].v;
    }
    THETA_HD double phi(double alpha) const {
        return L3_hat((xf64(x0[0]) + xf64(alpha) * xf64(pk[0])).v, (xf64(x0[1]) + xf64(alpha) * xf64(pk[1])).v);
    }
    THETA_HD double derphi(double alpha) const {
        double gv[2];
        dL3_hat((xf64(x0[0]) + xf64(alpha) * xf64(pk[0])).v, (xf64(x0[1]) + xf64(alpha) * xf64(pk[1])).v, gv);
        return (xf64(gv[0]) * xf64(pk[0]) + xf64(gv[1]) * xf64(pk[1])).v;
    }
};

// M_eq (Optimizer.py:320-326) with the forward-difference Jacobian of MINPACK fdjac1 (dense case:
// ml = mu = n - 1), epsfcn = machine epsilon (scipy's default when epsfcn=None)
struct MeqSys {
    xf64 cs[3];  // column sums of C_w
    xf64 mu[3];  // mu_hat handed to M3
    THETA_HDN void fcn(const xf64* x, xf64* f) {
        const xf64 temp = (x[0] * cs[0] + x[1] * cs[1]) + x[2] * cs[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) f[j] = ((mu[j] * temp) - (x[j] * cs[j])) - x[3];
        f[3] = ((x[0] + x[1]) + x[2]) - xf64(1.0);
    }
    THETA_HDN int jac(const xf64* x0, const xf64* fvec, xf64* J) {
        const xf64 eps = 1.4901161193847656e-08;  // sqrt(max(epsfcn, epsmch)) = 2^-26
        xf64 x[4], w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = x0[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const xf64 temp = x[j];
            xf64 h = eps * fabs(temp);
            if (h == xf64(0.0)) h = eps;
            x[j] = temp + h;
            fcn(x, w);
            x[j] = temp;
#pragma unroll
            for (int i = 0; i < 4; ++i) J[i + 4 * j] = (w[i] - fvec[i]) / h;
        }
        return 4;
    }
};

// L3 (Optimizer.py:236-244): denom = sum over columns (outer) and intervals (inner) of C_w[h][j] mu[j], left to right
struct DenTerm {
    const Ctx* c;
    const xf64* y;
    THETA_HD void operator()(int q, xf64* t) const {
        const int j = q / c->m, h = q - j * c->m;
        const int g = c->row[(size_t)h * c->rstride];
        const double cj = j == 0 ? c->tau : (j == 1 ? (double)c->ra[g] : (double)c->rb[g]);
        t[0] = (xf64(c->rN[h]) * xf64(cj)) * y[j];
    }
};
struct L3Term {
    const Ctx* c;
    const xf64* y;
    xf64 denom;
    double* vals;  // optional: numer_i / denom
    THETA_HD void operator()(int i, xf64* t) const {
        const int g = c->row[(size_t)i * c->rstride];
        const xf64 w = c->rN[i];
        const xf64 numer = ((w * xf64(c->tau)) * y[0] + (w * xf64((double)c->ra[g])) * y[1]) + (w * xf64((double)c->rb[g])) * y[2];
        const xf64 v = numer / denom;
        if (vals) vals[i] = v.v;
        t[0] = xf64(c->r[i]) * xf64(log(v.v));
    }
};

struct Out {
    double mu[3];     // M3's answer
    double muhat[3];  // what went into M3
    double nll;       // L3
    int status;       // THETA_ST_OK / FALLBACK_X0 / FALLBACK_MOVED / NONE_RANGE (values passed in: no dependency on theta_b200.h)
    int nfev;
};

struct StatusCodes {
    int ok, fallback_x0, fallback_moved, none_range;
};

template <class Exec>
THETA_HDN void solve(const Ctx& c, const StatusCodes& sc, Exec& ex, Out& o, double* vals = nullptr) {
    o.nll = 0.0;
    for (int j = 0; j < 3; ++j) o.mu[j] = o.muhat[j] = 0.0;
    // ---- fsolve(equations, [1/3, 1/3, 1/3, 1], fprime=jacobian)
    Kkt<Exec> sys{&c, &ex};
    const double third = 1.0 / 3.0;
    xf64 x[4] = {third, third, third, 1.0};
    const mp::Result res = mp::hybrj<xf64>(sys, x, xf64(1.49012e-8), 500, xf64(100.0));
    o.nfev = res.nfev;
    double mh[3] = {x[0].v, x[1].v, x[2].v};
    bool in_range = true;  // Misc.inRange: NaN passes
    for (int j = 0; j < 3; ++j)
        if (mh[j] < 0.0 || mh[j] > 1.0) in_range = false;
    int status = sc.ok;
    if (!in_range) {
        // ---- fmin_bfgs(L3_hat, [1/3, 1/3], fprime=dL3_hat): first iteration (see theta_n3.cuh::n3_solve)
        Ray<Exec> ray{&c, &ex, {third, third}, {0.0, 0.0}};
        double g0[2];
        const double f0 = ray.L3_hat(third, third);
        ray.dL3_hat(third, third, g0);
        mh[0] = third;
        mh[1] = third;
        mh[2] = (xf64(1.0) - (xf64(third) + xf64(third))).v;
        status = sc.fallback_x0;
        const double gmax = ::fmax(::fabs(g0[0]), ::fabs(g0[1]));
        if (gmax > 1e-5) {  // gtol; a NaN gradient compares false and BFGS returns x0
            ray.pk[0] = -g0[0];
            ray.pk[1] = -g0[1];
            const double gn = ::sqrt((xf64(g0[0]) * xf64(g0[0]) + xf64(g0[1]) * xf64(g0[1])).v);
            const double old_old = (xf64(f0) + xf64(gn) / xf64(2.0)).v;
            const double derphi0 = (xf64(g0[0]) * xf64(ray.pk[0]) + xf64(g0[1]) * xf64(ray.pk[1])).v;
            double alpha = 0.0;
            if (ls::line_search_wolfe12(ray, f0, old_old, derphi0, alpha)) {
                mh[0] = (xf64(third) + xf64(alpha) * xf64(ray.pk[0])).v;
                mh[1] = (xf64(third) + xf64(alpha) * xf64(ray.pk[1])).v;
                mh[2] = (xf64(1.0) - (xf64(mh[0]) + xf64(mh[1]))).v;
                status = sc.fallback_moved;
                for (int j = 0; j < 3; ++j)
                    if (mh[j] < 0.0 || mh[j] > 1.0) status = sc.none_range;  // Optimizer.py:159-161
            }
        }
    }
    o.status = status;
    for (int j = 0; j < 3; ++j) o.muhat[j] = mh[j];
    if (status == sc.none_range) return;
    // ---- M3: fsolve(M_eq, [.33, .33, .33, 0]) -- hybrd, maxfev = 200 * (4 + 1)
    MeqSys ms;
    ms.cs[0] = c.S0; ms.cs[1] = c.S1; ms.cs[2] = c.S2;
    for (int j = 0; j < 3; ++j) ms.mu[j] = mh[j];
    xf64 y[4] = {0.33, 0.33, 0.33, 0.0};
    mp::hybrj<xf64>(ms, y, xf64(1.49012e-8), 1000, xf64(100.0));
    // ---- L3(answer, C_w, r)
    xf64 denom[1], tot[1];
    DenTerm dt{&c, y};
    ex.template sums<1>(3 * c.m, dt, denom);
    L3Term lt{&c, y, denom[0], vals};
    ex.template sums<1>(c.m, lt, tot);
    o.nll = -tot[0].v;
    for (int j = 0; j < 3; ++j) o.mu[j] = y[j].v;
}

// Is the candidate rank-deficient?  rows g with reads (R_g > 0): fewer than three distinct (a, b), or all on one
// line; or a zero tumour column.  A handful of integer operations on the grouped statistics.
THETA_HD bool rank_deficient(int G, const double* R, const double* a, const double* b, double S1, double S2) {
    if (!(S1 > 0.0) || !(S2 > 0.0)) return true;
    int n = 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    for (int i = 0; i < G; ++i) {
        if (!(R[i] > 0.0)) continue;
        const int x = (int)a[i], y = (int)b[i];
        if (n == 0) { x0 = x; y0 = y; n = 1; }
        else if (n == 1) { if (x != x0 || y != y0) { x1 = x; y1 = y; n = 2; } }
        else if ((x1 - x0) * (y - y0) - (y1 - y0) * (x - x0) != 0) return false;
    }
    return true;
}

}  // namespace ref3
}  // namespace theta
