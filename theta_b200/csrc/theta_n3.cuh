// theta_n3.cuh -- per-candidate solver for n = 3 (normal + two tumour populations).
//
// Follows Optimizer._solve_n3plus (python/Optimizer.py:128-165) on grouped statistics.  A candidate
// assigns every interval one of NR admissible rows (a, b) (python/Enumerator.py:272-298); after
// rN_i cancels, the Lagrangian system (Optimizer.py:267-318) depends on the candidate only through
//     R_g = sum of r_i over intervals carrying row g,   S1 = sum rN_i a_i,   S2 = sum rN_i b_i:
//     d_g   = mu0/A + a_g mu1/S1 + b_g mu2/S2
//     F_k   = -t_k sum_g c_gk R_g / d_g - lambda     (c_g = (1, a_g, b_g), t = (1/A, 1/S1, 1/S2))
//     F_3   = 1 - (mu0 + mu1 + mu2)
//     J_kh  = t_k t_h sum_g c_gk c_gh R_g / d_g^2,  J_k3 = J_3k = -1,  J_33 = 0
// solved by the hybrj restatement in hybrj4.cuh from (1/3, 1/3, 1/3, 1).
#pragma once
#include "hybrj4.cuh"
#include "bfgs_ls.cuh"
#include "theta_core.cuh"
#include "n3_ref.cuh"

namespace theta {

constexpr int kMaxRows3 = 64;  // (k+1)^2 rows for k <= 7, filtered by _is_valid_row

struct N3Const {
    double Kr;     // -sum r_i log(rN_i): shifts L3_hat from grouped to the reference's per-interval scale
    double A;      // sum rN
    double Rtot;   // sum r
    double K0;     // -sum r_i log(rN_i / A): NLL of the all-normal model
    double tau;
    int m, NR;
    int any_zero_rN;
};

struct Groups3 {
    int G;
    double R[kMaxRows3];
    double a[kMaxRows3];
    double b[kMaxRows3];
};

struct Kkt3 {
    const Groups3* g;
    double t0, t1, t2;
    THETA_HDN void fcn(const double* x, double* f) {
        const double u0 = x[0] * t0, u1 = x[1] * t1, u2 = x[2] * t2;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < g->G; ++i) {
            const double d = u0 + g->a[i] * u1 + g->b[i] * u2;
            const double q = g->R[i] / d;
            s0 += q;
            s1 += g->a[i] * q;
            s2 += g->b[i] * q;
        }
        f[0] = -(t0 * s0) - x[3];
        f[1] = -(t1 * s1) - x[3];
        f[2] = -(t2 * s2) - x[3];
        f[3] = 1.0 - ((x[0] + x[1]) + x[2]);
    }
    THETA_HDN int jac(const double* x, const double*, double* J) {
        const double u0 = x[0] * t0, u1 = x[1] * t1, u2 = x[2] * t2;
        double j00 = 0, j01 = 0, j02 = 0, j11 = 0, j12 = 0, j22 = 0;
        for (int i = 0; i < g->G; ++i) {
            const double a = g->a[i], b = g->b[i];
            const double d = u0 + a * u1 + b * u2;
            const double q = g->R[i] / (d * d);
            j00 += q;
            j01 += a * q;
            j02 += b * q;
            j11 += a * a * q;
            j12 += a * b * q;
            j22 += b * b * q;
        }
        J[0 + 4 * 0] = t0 * t0 * j00;
        J[1 + 4 * 0] = J[0 + 4 * 1] = t0 * t1 * j01;
        J[2 + 4 * 0] = J[0 + 4 * 2] = t0 * t2 * j02;
        J[1 + 4 * 1] = t1 * t1 * j11;
        J[2 + 4 * 1] = J[1 + 4 * 2] = t1 * t2 * j12;
        J[2 + 4 * 2] = t2 * t2 * j22;
        for (int i = 0; i < 4; ++i) {
            J[3 + 4 * i] = -1.0;
            J[i + 4 * 3] = -1.0;
        }
        J[3 + 4 * 3] = 0.0;
        return 0;
    }
};

// L3_hat and the (sign-flipped) dL3_hat of the reference along the ray x0 + alpha * pk in (mu_hat_0, mu_hat_1)
// (python/Optimizer.py:246-265), on grouped statistics: d_i = rN_i * D_g.
struct Ray3 {
    const Groups3* g;
    double t0, t1, t2, Kr;
    double x0[2], pk[2];
    THETA_HDN double L3_hat(double m0, double m1) const {
        const double m2 = 1.0 - (m0 + m1);
        double s = 0.0;
        for (int i = 0; i < g->G; ++i) {
            const double D = m0 * t0 + g->a[i] * (m1 * t1) + g->b[i] * (m2 * t2);
            s += g->R[i] * log(D);  // log of a negative number is NaN, as in numpy
        }
        return Kr - s;
    }
    THETA_HDN void dL3_hat(double m0, double m1, double* out) const {
        double v0 = 0.0, v1 = 0.0;
        for (int i = 0; i < g->G; ++i) {
            const double c2 = g->b[i] * t2;
            const double n0 = t0 - c2, n1 = g->a[i] * t1 - c2;
            const double den = n0 * m0 + n1 * m1 + c2;
            v0 += g->R[i] * (n0 / den);
            v1 += g->R[i] * (n1 / den);
        }
        out[0] = v0;
        out[1] = v1;
    }
    THETA_HD double phi(double alpha) const { return L3_hat(x0[0] + alpha * pk[0], x0[1] + alpha * pk[1]); }
    THETA_HD double derphi(double alpha) const {
        double gv[2];
        dL3_hat(x0[0] + alpha * pk[0], x0[1] + alpha * pk[1], gv);
        return gv[0] * pk[0] + gv[1] * pk[1];
    }
};

struct Solve3Out {
    double mu[3];      // M3 (Optimizer.py:320-330), closed form
    double muhat[3];   // the normalised-coordinate solution the reference's fsolve stops at
    double nll;        // L3 (Optimizer.py:236-244)
    int status;
    int nfev;
};

// L3 on grouped statistics at normalised-coordinate solution muhat
THETA_HDN void n3_finish(const Groups3& g, double S1, double S2, const N3Const& k, const double* muhat, Solve3Out& o) {
    // M3: mu_j proportional to muhat_j / S_j (S_0 = tau A)
    const double w0 = muhat[0] / (k.tau * k.A), w1 = muhat[1] / S1, w2 = muhat[2] / S2;
    const double ws = (w0 + w1) + w2;
    const double m0 = w0 / ws, m1 = w1 / ws, m2 = w2 / ws;
    const double wbar = (k.tau * k.A * m0 + S1 * m1 + S2 * m2) / k.A;
    const double lw = log(wbar);
    double acc = 0.0;
    for (int i = 0; i < g.G; ++i) {
        if (g.R[i] != 0.0) {
            const double w = k.tau * m0 + g.a[i] * m1 + g.b[i] * m2;
            acc = fma(g.R[i], log(w) - lw, acc);
        }
    }
    o.mu[0] = m0;
    o.mu[1] = m1;
    o.mu[2] = m2;
    o.muhat[0] = muhat[0];
    o.muhat[1] = muhat[1];
    o.muhat[2] = muhat[2];
    o.nll = k.K0 - acc;
}

// ---- screening: a rigorous lower bound on anything the reference can return for this candidate ----
//
// In normalised coordinates y (y_j >= 0, sum y = 1) the NLL is convex:
//     f(y) = K0 - sum_g R_g log(A d_g(y)),     d_g = y0/A + a_g y1/S1 + b_g y2/S2
// and whatever Optimizer._solve_n3plus ends with -- the hybrj root if it is in range, the BFGS start point
// (1/3,1/3,1/3), or a point the fallback moved to inside [0,1]^3 -- is evaluated by M3 / L3 at a non-negative
// mu_hat, i.e. at f(mu_hat / sum mu_hat) >= f* = min over the simplex.  For any interior y the first-order
// (Frank-Wolfe) bound gives   f* >= f(y) - gap(y),   gap(y) = max_j t_j sum_g c_gj R_g / d_g  -  sum_g R_g >= 0
// (grad f . y = -Rtot exactly).  A few damped Newton steps in (y0, y1), warm-started from the previous leaf of the
// DFS, make the gap small; f(y) is evaluated with fp32 MUFU logarithms whose error is bounded as in
// theta_core.cuh::n2_l2_approx (arguments A d_g kept within [2^-15, 2^15], |log2| < 16, 2 ulp = 2^-19).
// Returns true when  f32(y) - gap - E  >  T + margin  (the candidate can neither be a strict prefix minimum of
// the stream nor lie within the tie window); the caller then skips the exact solve.  lb_out is that bound.
#define THETA_N3_APPROX_C 6e-6
struct Screen3 {
    double y0, y1;  // warm start (interior point of the simplex)
};

#ifdef __CUDACC__
// LT: the fast_log table in shared memory.  With it f(y) is evaluated in fp64 (E shrinks from 6e-6 Rtot to the
// rounding noise of the sums) at ~20 more instructions per group; LT == nullptr keeps the fp32 MUFU logarithms.
__device__ __noinline__ bool n3_screen(const Groups3& g, double S1, double S2, const N3Const& k, Screen3& ws,
                                       double T_plus_margin, double& lb_out, const double2* LT = nullptr) {
    lb_out = -INFINITY;
    if (k.any_zero_rN || !(S1 > 0.0) || !(S2 > 0.0)) return false;
    const double t1 = k.A / S1, t2 = k.A / S2;  // A d_g = y0 + a (A/S1) y1 + b (A/S2) y2
    const double E = (LT ? 1e-11 : THETA_N3_APPROX_C) * k.Rtot;
    double y0 = ws.y0, y1 = ws.y1;
    if (!(y0 > 1e-6 && y1 > 1e-6 && y0 + y1 < 1.0 - 1e-6)) { y0 = 1.0 / 3.0; y1 = 1.0 / 3.0; }
    for (int it = 0; it < 3; ++it) {
        const double y2 = 1.0 - (y0 + y1);
        const double u1 = y1 * t1, u2 = y2 * t2;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, h00 = 0.0, h01 = 0.0, h02 = 0.0, h11 = 0.0, h12 = 0.0, h22 = 0.0, lacc = 0.0;
        bool ok = true;
        for (int i = 0; i < g.G; ++i) {
            const double a = g.a[i], b = g.b[i], R = g.R[i];
            const double d = y0 + a * u1 + b * u2;  // A d_g > 0 in the interior
            const double r = fast_rcp1(d);
            const double q = R * r, qq = q * r, aq = a * qq, bq = b * qq;
            s0 += q; s1 = fma(a, q, s1); s2 = fma(b, q, s2);
            h00 += qq; h01 += aq; h02 += bq; h11 = fma(a, aq, h11); h12 = fma(a, bq, h12); h22 = fma(b, bq, h22);
            if (LT) {
                lacc = fma(R, fast_log(d, LT), lacc);  // natural log, |error| ~ 1e-16
            } else {
                const float df = (float)d;
                if (!(df >= 3.0517578125e-05f && df <= 32768.0f)) ok = false;
                lacc = fma(R, (double)__log2f(df), lacc);
            }
        }
        if (!ok) return false;
        // gradient components (up to sign): G_j = t_j sum c_gj q_g ; gap = max_j G_j - Rtot
        const double G0 = s0, G1 = t1 * s1, G2 = t2 * s2;
        const double gap = fmax(G0, fmax(G1, G2)) - k.Rtot;
        const double fy = LT ? k.K0 - lacc : fma(-0.6931471805599453094, lacc, k.K0);
        const double lb = fy - fmax(gap, 0.0) * (1.0 + 1e-9) - E;
        if (lb > lb_out) lb_out = lb;
        if (lb > T_plus_margin) { ws.y0 = y0; ws.y1 = y1; return true; }
        if (it == 2) break;
        // damped Newton step on the reduced problem (y2 = 1 - y0 - y1); grad f_j = -G_j
        const double g0 = -(G0 - G2), g1 = -(G1 - G2);
        const double H00 = h00, H01 = t1 * h01, H02 = t2 * h02, H11 = t1 * t1 * h11, H12 = t1 * t2 * h12, H22 = t2 * t2 * h22;
        const double r00 = H00 - 2.0 * H02 + H22, r01 = H01 - H02 - H12 + H22, r11 = H11 - 2.0 * H12 + H22;
        const double det = r00 * r11 - r01 * r01;
        if (!(det > 0.0)) break;
        const double idet = 1.0 / det;
        const double e0 = -(r11 * g0 - r01 * g1) * idet, e1 = -(r00 * g1 - r01 * g0) * idet, e2 = -(e0 + e1);
        double alpha = 1.0;
        if (e0 < 0.0) alpha = fmin(alpha, 0.9 * y0 / -e0);
        if (e1 < 0.0) alpha = fmin(alpha, 0.9 * y1 / -e1);
        if (e2 < 0.0) alpha = fmin(alpha, 0.9 * y2 / -e2);
        if (!(alpha > 0.0)) break;
        y0 = fma(alpha, e0, y0);
        y1 = fma(alpha, e1, y1);
        if (!(y0 > 0.0 && y1 > 0.0 && y0 + y1 < 1.0)) break;
    }
    ws.y0 = y0; ws.y1 = y1;
    return false;
}
#endif

THETA_HDN void n3_solve(const Groups3& g, double S1, double S2, const N3Const& k, Solve3Out& o) {
    o.nfev = 0;
    o.nll = 0.0;
    o.mu[0] = o.mu[1] = o.mu[2] = 0.0;
    o.muhat[0] = o.muhat[1] = o.muhat[2] = 0.0;
    if (k.any_zero_rN || !(S1 > 0.0) || !(S2 > 0.0)) {
        // a zero tumour column makes normalize_C (Optimizer.py:167-174) produce NaN; the reference
        // then returns NaN-tainted numbers (Q8).  Both columns zero == the all-normal model, whose
        // NLL the reference does evaluate correctly (K0); one zero column is reported as None.
        if (!k.any_zero_rN && S1 == 0.0 && S2 == 0.0) {
            o.status = THETA_ST_DEGENERATE;
            o.nll = k.K0;
            o.mu[0] = o.mu[1] = o.mu[2] = NAN;
            return;
        }
        o.status = THETA_ST_NONE_NAN;
        return;
    }
    Kkt3 sys{&g, 1.0 / k.A, 1.0 / S1, 1.0 / S2};
    double x[4] = {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0, 1.0};
    const mp::Result r = mp::hybrj(sys, x, 1.49012e-8, 500, 100.0);
    o.nfev = r.nfev;
    // Misc.inRange (python/Misc.py:49-57): NaN passes
    bool in_range = true;
    for (int j = 0; j < 3; ++j)
        if (x[j] < 0.0 || x[j] > 1.0) in_range = false;
    if (in_range) {
        o.status = THETA_ST_OK;
        n3_finish(g, S1, S2, k, x, o);
        return;
    }
    // fmin_bfgs(L3_hat, [1/3, 1/3], fprime=dL3_hat) (Optimizer.py:151-161): first (and only) iteration
    const double third = 1.0 / 3.0;
    Ray3 ray{&g, 1.0 / k.A, 1.0 / S1, 1.0 / S2, k.Kr, {third, third}, {0.0, 0.0}};
    double g0[2];
    const double f0 = ray.L3_hat(third, third);
    ray.dL3_hat(third, third, g0);
    double mh[3] = {third, third, 1.0 - (third + third)};
    int status = THETA_ST_FALLBACK_X0;
    const double gmax = fmax(fabs(g0[0]), fabs(g0[1]));
    if (gmax > 1e-5) {  // gtol; NaN compares false and BFGS returns x0
        ray.pk[0] = -g0[0];
        ray.pk[1] = -g0[1];
        const double old_old = f0 + sqrt(g0[0] * g0[0] + g0[1] * g0[1]) / 2;
        const double derphi0 = g0[0] * ray.pk[0] + g0[1] * ray.pk[1];
        double alpha = 0.0;
        if (ls::line_search_wolfe12(ray, f0, old_old, derphi0, alpha)) {
            // a step was taken: BFGS stops right after it (non-finite value) or keeps going from there;
            // either way the reference continues with this point
            mh[0] = third + alpha * ray.pk[0];
            mh[1] = third + alpha * ray.pk[1];
            mh[2] = 1.0 - (mh[0] + mh[1]);
            status = THETA_ST_FALLBACK_MOVED;
            for (int j = 0; j < 3; ++j)
                if (mh[j] < 0.0 || mh[j] > 1.0) status = THETA_ST_NONE_RANGE;  // Optimizer.py:159-161
        }
    }
    o.status = status;
    if (status == THETA_ST_NONE_RANGE) return;
    n3_finish(g, S1, S2, k, mh, o);
}

// Which solver a candidate takes (mode): 0 = grouped statistics (n3_solve) unless the candidate is rank-deficient
// or an interval has no normal reads -- then the reference's result is a function of its own rounding sequence and
// only the reference-order solver (n3_ref.cuh) reproduces it; 1 = reference-order solver always (the records a
// search returns: mu and NLL then carry the reference's roundings, not just its tolerance); 2 = grouped always.
#define THETA_N3_MODE_AUTO 0
#define THETA_N3_MODE_REF 1
#define THETA_N3_MODE_GROUPED 2
THETA_HDN void n3_solve_auto(const Groups3& g, double S1, double S2, const N3Const& k, const ref3::Ctx& c, int mode,
                             Solve3Out& o) {
    const bool ref = mode == THETA_N3_MODE_REF ||
                     (mode == THETA_N3_MODE_AUTO && (k.any_zero_rN || ref3::rank_deficient(g.G, g.R, g.a, g.b, S1, S2)));
    if (!ref) {
        n3_solve(g, S1, S2, k, o);
        return;
    }
    const ref3::StatusCodes sc{THETA_ST_OK, THETA_ST_FALLBACK_X0, THETA_ST_FALLBACK_MOVED, THETA_ST_NONE_RANGE};
    ref3::Out r;
    ref3::SerialExec ex;
    ref3::solve(c, sc, ex, r);
    for (int j = 0; j < 3; ++j) { o.mu[j] = r.mu[j]; o.muhat[j] = r.muhat[j]; }
    o.nll = r.nll;
    o.status = r.status;
    o.nfev = r.nfev;
}

}  // namespace theta
