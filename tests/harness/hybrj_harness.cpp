// Host harness (test infrastructure): runs theta::mp::hybrj (theta_b200/csrc/hybrj4.cuh) on the
// reference's per-interval KKT system, evaluated in the reference's arithmetic order
// (python/Optimizer.py:267-318), so the port can be compared with scipy.optimize.fsolve bit for bit.
#include <cmath>
#include "../../theta_b200/csrc/hybrj4.cuh"

struct RefSys {
    int m;
    const double* r;   // [m]
    const double* Ch;  // [m][3] normalised weighted matrix
    void fcn(const double* x, double* f) {
        double s[3] = {0, 0, 0};
        for (int i = 0; i < m; ++i) {
            const double d = (Ch[3 * i] * x[0] + Ch[3 * i + 1] * x[1]) + Ch[3 * i + 2] * x[2];
            for (int k = 0; k < 3; ++k) s[k] += (r[i] * Ch[3 * i + k]) / d;
        }
        for (int k = 0; k < 3; ++k) f[k] = (-s[k]) - x[3];
        f[3] = 1 - ((x[0] + x[1]) + x[2]);
    }
    int jac(const double* x, const double*, double* J) {  // column-major J[i + 4*j] = d f_i / d x_j
        for (int a = 0; a < 16; ++a) J[a] = 0.0;
        for (int k = 0; k < 3; ++k)
            for (int h = 0; h < 3; ++h) {
                double s = 0.0;
                for (int i = 0; i < m; ++i) {
                    const double d = (Ch[3 * i] * x[0] + Ch[3 * i + 1] * x[1]) + Ch[3 * i + 2] * x[2];
                    s += ((r[i] * Ch[3 * i + k]) * Ch[3 * i + h]) / std::pow(d, 2.0);
                }
                J[k + 4 * h] = s;
            }
        for (int i = 0; i < 4; ++i) { J[3 + 4 * i] = -1.0; J[i + 4 * 3] = -1.0; }
        J[3 + 4 * 3] = 0.0;
        return 0;
    }
};

extern "C" int hybrj_ref_system(int m, const double* r, const double* Ch, double* x, int* nfev, int* njev) {
    RefSys s{m, r, Ch};
    theta::mp::Result res = theta::mp::hybrj(s, x, 1.49012e-8, 500, 100.0);
    *nfev = res.nfev;
    *njev = res.njev;
    return res.info;
}
