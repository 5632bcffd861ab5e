// bfgs_ls.cuh -- what scipy.optimize.fmin_bfgs does on the first iteration of the reference's
// fallback (python/Optimizer.py:151-161).
//
// The reference hands BFGS a gradient with the wrong sign (dL3_hat, Optimizer.py:255-265 -- Q7), so the
// first search direction is the ascent direction and the first trial step (length ~1.01 in mu-hat space)
// leaves the domain of the logarithms.  Measured on the reference the run then ends in one
// of two ways: the line search fails and BFGS returns its start point bit-exactly (the candidate stays in
// the race with mu_hat = (1/3, 1/3, 1/3)), or a NaN trial point is "accepted" and BFGS stops at a point
// outside [0,1]^3, so Optimizer.solve returns None.  Which one happens is decided by the NaN / sign
// pattern of phi and phi' along the ray -- reproduced here by following the published algorithms step by
// step: MINPACK-2 dcsrch/dcstep (More & Thuente 1994; scipy _dcsrch.py), then the strong-Wolfe search of
// Nocedal & Wright, Alg. 3.5/3.6 (scipy _linesearch.py: scalar_search_wolfe2, _zoom, _cubicmin, _quadmin),
// including Python's min/max/comparison semantics for NaN.
#pragma once
#include <math.h>

#include "hybrj4.cuh"

namespace theta {
namespace ls {

THETA_HD double py_min(double a, double b) { return (b < a) ? b : a; }  // Python min(a, b)
THETA_HD double py_max(double a, double b) { return (b > a) ? b : a; }  // Python max(a, b)
THETA_HD double py_max3(double a, double b, double c) { return py_max(py_max(a, b), c); }
THETA_HD double np_sign(double x) { return (x != x) ? x : (x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0)); }
THETA_HD double np_clip(double x, double lo, double hi) {  // NaN propagates
    if (x != x) return x;
    return x < lo ? lo : (x > hi ? hi : x);
}
THETA_HD bool finite_(double x) { return x == x && fabs(x) <= 1.7976931348623157e308; }

struct DcState {
    bool brackt;
    int stage;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
};

// scipy _dcsrch.py: dcstep
THETA_HDN void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp,
                     double dp, bool& brackt, double stpmin, double stpmax) {
    const double sgnd = np_sign(dp) * np_sign(dx);
    double stpf;
    if (fp > fx) {
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = py_max3(fabs(theta), fabs(dx), fabs(dp));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma *= -1;
        const double p = (gamma - dx) + theta;
        const double q = ((gamma - dx) + gamma) + dp;
        const double r = p / q;
        const double stpc = stx + r * (stp - stx);
        const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (fabs(stpc - stx) <= fabs(stpq - stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {
        const double theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = py_max3(fabs(theta), fabs(dx), fabs(dp));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma *= -1;
        const double p = (gamma - dp) + theta;
        const double q = ((gamma - dp) + gamma) + dx;
        const double r = p / q;
        const double stpc = stp + r * (stx - stp);
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
        else stpf = stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {
        const double theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = py_max3(fabs(theta), fabs(dx), fabs(dp));
        double gamma = s * sqrt(py_max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta;
        const double q = (gamma + (dx - dp)) + gamma;
        const double r = p / q;
        double stpc;
        if (r < 0 && gamma != 0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            if (stp > stx) stpf = py_min(stp + 0.66 * (sty - stp), stpf);
            else stpf = py_max(stp + 0.66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            stpf = np_clip(stpf, stpmin, stpmax);
        }
    } else {
        if (brackt) {
            const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            const double s = py_max3(fabs(theta), fabs(dy), fabs(dp));
            double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            const double p = (gamma - dp) + theta;
            const double q = ((gamma - dp) + gamma) + dy;
            const double r = p / q;
            stpf = stp + r * (sty - stp);
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < 0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

enum Task { START = 0, FG = 1, CONV = 2, WARN = 3, ERROR = 4 };

// scipy _dcsrch.py: DCSRCH._iterate
THETA_HDN Task dcsrch_iterate(DcState& S, double& stp, double f, double g, Task task, double ftol, double gtol, double xtol,
                             double stpmin, double stpmax) {
    const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
    if (task == START) {
        if (stp < stpmin || stp > stpmax || g >= 0) return ERROR;
        S.brackt = false;
        S.stage = 1;
        S.finit = f;
        S.ginit = g;
        S.gtest = ftol * S.ginit;
        S.width = stpmax - stpmin;
        S.width1 = S.width / p5;
        S.stx = 0.0; S.fx = S.finit; S.gx = S.ginit;
        S.sty = 0.0; S.fy = S.finit; S.gy = S.ginit;
        S.stmin = 0;
        S.stmax = stp + xtrapu * stp;
        return FG;
    }
    const double ftest = S.finit + stp * S.gtest;
    if (S.stage == 1 && f <= ftest && g >= 0) S.stage = 2;
    Task t = FG;
    if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) t = WARN;
    if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) t = WARN;
    if (stp == stpmax && f <= ftest && g <= S.gtest) t = WARN;
    if (stp == stpmin && (f > ftest || g >= S.gtest)) t = WARN;
    if (f <= ftest && fabs(g) <= gtol * -S.ginit) t = CONV;
    if (t == WARN || t == CONV) return t;
    if (S.stage == 1 && f <= S.fx && f > ftest) {
        const double fm = f - stp * S.gtest;
        double fxm = S.fx - S.stx * S.gtest;
        double fym = S.fy - S.sty * S.gtest;
        const double gm = g - S.gtest;
        double gxm = S.gx - S.gtest;
        double gym = S.gy - S.gtest;
        dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
        S.fx = fxm + S.stx * S.gtest;
        S.fy = fym + S.sty * S.gtest;
        S.gx = gxm + S.gtest;
        S.gy = gym + S.gtest;
    } else {
        dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, stp, f, g, S.brackt, S.stmin, S.stmax);
    }
    if (S.brackt) {
        if (fabs(S.sty - S.stx) >= p66 * S.width1) stp = S.stx + p5 * (S.sty - S.stx);
        S.width1 = S.width;
        S.width = fabs(S.sty - S.stx);
    }
    if (S.brackt) {
        S.stmin = py_min(S.stx, S.sty);
        S.stmax = py_max(S.stx, S.sty);
    } else {
        S.stmin = stp + xtrapl * (stp - S.stx);
        S.stmax = stp + xtrapu * (stp - S.stx);
    }
    stp = np_clip(stp, stpmin, stpmax);
    if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) || (S.brackt && S.stmax - S.stmin <= xtol * S.stmax)) stp = S.stx;
    return FG;
}

// scipy _linesearch.py: _cubicmin / _quadmin (None <=> returns false)
THETA_HDN bool cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc, double& xmin) {
    const double Cc = fpa;
    const double db = b - a, dc = c - a;
    const double denom = (db * dc) * (db * dc) * (db - dc);
    const double d00 = dc * dc, d01 = -(db * db), d10 = -(dc * dc * dc), d11 = db * db * db;
    const double v0 = fb - fa - Cc * db, v1 = fc - fa - Cc * dc;
    double A = d00 * v0 + d01 * v1;
    double B = d10 * v0 + d11 * v1;
    if (denom == 0.0) return false;  // divide='raise'
    A /= denom;
    B /= denom;
    const double radical = B * B - 3 * A * Cc;
    if (radical < 0.0 || A == 0.0) return false;  // invalid / divide raise
    xmin = a + (-B + sqrt(radical)) / (3 * A);
    return finite_(xmin);
}
THETA_HDN bool quadmin(double a, double fa, double fpa, double b, double fb, double& xmin) {
    const double D = fa, Cc = fpa;
    const double db = b - a * 1.0;
    if (db * db == 0.0) return false;
    const double B = (fb - D - Cc * db) / (db * db);
    if (B == 0.0) return false;
    xmin = a - Cc / (2.0 * B);
    return finite_(xmin);
}

// Fn provides double phi(double alpha), double derphi(double alpha)
template <class Fn>
THETA_HDN bool zoom(Fn& fn, double a_lo, double a_hi, double phi_lo, double phi_hi, double derphi_lo, double phi0,
                   double derphi0, double c1, double c2, double& a_star) {
    const int maxiter = 10;
    int i = 0;
    const double delta1 = 0.2, delta2 = 0.1;
    double phi_rec = phi0, a_rec = 0;
    for (;;) {
        const double dalpha = a_hi - a_lo;
        double a, b;
        if (dalpha < 0) { a = a_hi; b = a_lo; } else { a = a_lo; b = a_hi; }
        double a_j = 0.0, cchk = 0.0;
        bool have = false;
        if (i > 0) {
            cchk = delta1 * dalpha;
            have = cubicmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, a_rec, phi_rec, a_j);
        }
        if (i == 0 || !have || (a_j > b - cchk) || (a_j < a + cchk)) {
            const double qchk = delta2 * dalpha;
            have = quadmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, a_j);
            if (!have || (a_j > b - qchk) || (a_j < a + qchk)) a_j = a_lo + 0.5 * dalpha;
        }
        const double phi_aj = fn.phi(a_j);
        if ((phi_aj > phi0 + c1 * a_j * derphi0) || (phi_aj >= phi_lo)) {
            phi_rec = phi_hi; a_rec = a_hi; a_hi = a_j; phi_hi = phi_aj;
        } else {
            const double derphi_aj = fn.derphi(a_j);
            if (fabs(derphi_aj) <= -c2 * derphi0) { a_star = a_j; return true; }
            if (derphi_aj * (a_hi - a_lo) >= 0) {
                phi_rec = phi_hi; a_rec = a_hi; a_hi = a_lo; phi_hi = phi_lo;
            } else {
                phi_rec = phi_lo; a_rec = a_lo;
            }
            a_lo = a_j; phi_lo = phi_aj; derphi_lo = derphi_aj;
        }
        ++i;
        if (i > maxiter) return false;
    }
}

// _line_search_wolfe12 as called by _minimize_bfgs (amin=1e-100, amax=1e100, c1=1e-4, c2=0.9).
// Returns true and the step length if a step is taken, false if _LineSearchError is raised.
template <class Fn>
THETA_HDN bool line_search_wolfe12(Fn& fn, double phi0, double old_phi0, double derphi0, double& alpha_out) {
    const double c1 = 1e-4, c2 = 0.9, amax = 1e100, amin = 1e-100, xtol = 1e-14;
    // ---- scalar_search_wolfe1 (DCSRCH)
    {
        double alpha1 = 1.0;
        if (derphi0 != 0) {
            alpha1 = py_min(1.0, 1.01 * 2 * (phi0 - old_phi0) / derphi0);
            if (alpha1 < 0) alpha1 = 1.0;
        }
        DcState S;
        double phi1 = phi0, derphi1 = derphi0;
        double stp = alpha1;
        Task task = START;
        bool ok = false, done = false;
        for (int i = 0; i < 100 && !done; ++i) {
            task = dcsrch_iterate(S, stp, phi1, derphi1, task, c1, c2, xtol, amin, amax);
            if (!finite_(stp)) { done = true; break; }   // task = WARN, stp = None
            if (task == FG) {
                phi1 = fn.phi(stp);
                derphi1 = fn.derphi(stp);
            } else {
                ok = (task == CONV);
                done = true;
            }
        }
        if (ok) { alpha_out = stp; return true; }
    }
    // ---- scalar_search_wolfe2
    double alpha0 = 0, alpha1 = 1.0;
    if (derphi0 != 0) alpha1 = py_min(1.0, 1.01 * 2 * (phi0 - old_phi0) / derphi0);
    if (alpha1 < 0) alpha1 = 1.0;
    alpha1 = py_min(alpha1, amax);
    double phi_a1 = fn.phi(alpha1);
    double phi_a0 = phi0, derphi_a0 = derphi0;
    for (int i = 0; i < 10; ++i) {
        if (alpha1 == 0 || alpha0 > amax) return false;
        const bool not_first = i > 0;
        if ((phi_a1 > phi0 + c1 * alpha1 * derphi0) || ((phi_a1 >= phi_a0) && not_first))
            return zoom(fn, alpha0, alpha1, phi_a0, phi_a1, derphi_a0, phi0, derphi0, c1, c2, alpha_out);
        const double derphi_a1 = fn.derphi(alpha1);
        if (fabs(derphi_a1) <= -c2 * derphi0) { alpha_out = alpha1; return true; }
        if (derphi_a1 >= 0) return zoom(fn, alpha1, alpha0, phi_a1, phi_a0, derphi_a1, phi0, derphi0, c1, c2, alpha_out);
        const double alpha2 = py_min(2 * alpha1, amax);
        alpha0 = alpha1;
        alpha1 = alpha2;
        phi_a0 = phi_a1;
        phi_a1 = fn.phi(alpha1);
        derphi_a0 = derphi_a1;
    }
    alpha_out = alpha1;  // "did not converge": the last trial step is returned, and BFGS takes it
    return true;
}

}  // namespace ls
}  // namespace theta
