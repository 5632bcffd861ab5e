// Host harness (test infrastructure): runs the reference-order n=3 solver (theta_b200/csrc/n3_ref.cuh) on the CPU so
// that it can be compared with the golden vectors of the live reference without a GPU.  Build: g++ -O2 -ffp-contract=off -mfma.
#include "../../theta_b200/csrc/n3_ref.cuh"

extern "C" int n3_ref_solve_host(int m, int tau, const double* r, const double* rN, const int* a, const int* b,
                                 double* mu, double* muhat, double* nll, int* nfev) {
    using namespace theta;
    uint8_t row[1024], ra[1024], rb[1024];
    double S0 = 0, S1 = 0, S2 = 0;
    for (int i = 0; i < m; ++i) {
        row[i] = (uint8_t)i; ra[i] = (uint8_t)a[i]; rb[i] = (uint8_t)b[i];
        S0 += rN[i] * tau; S1 += rN[i] * a[i]; S2 += rN[i] * b[i];
    }
    ref3::Ctx c{m, r, rN, row, 1, ra, rb, (double)tau, S0, S1, S2};
    ref3::StatusCodes sc{0, 3, 4, 10};
    ref3::Out o;
    ref3::SerialExec ex;
    ref3::solve(c, sc, ex, o);
    for (int j = 0; j < 3; ++j) { mu[j] = o.mu[j]; muhat[j] = o.muhat[j]; }
    *nll = o.nll; *nfev = o.nfev;
    return o.status;
}

extern "C" void pow2_batch(const double* x, double* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = theta::glibc::pow2(x[i]);
}
