// Host harness (test infrastructure): runs the device solver n3_solve (theta_b200/csrc/theta_n3.cuh) on the CPU
// (n3_solve_auto: grouped statistics, reference-order solver for rank-deficient candidates) so its arithmetic can be compared with SciPy / the golden vectors without a GPU.
#include "../../theta_b200/csrc/theta_n3.cuh"

namespace theta { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return -2; } }

extern "C" int n3_solve_host(int m, int tau, const double* r, const double* rN, const int* a, const int* b,
                             double* mu, double* muhat, double* nll, int* nfev, int mode) {
    using namespace theta;
    N3Const k{};
    long double A = 0, Rt = 0, K0 = 0, Kr = 0;
    for (int i = 0; i < m; ++i) { A += rN[i]; Rt += r[i]; }
    for (int i = 0; i < m; ++i) if (r[i] > 0) { K0 -= (long double)r[i] * logl((long double)rN[i] / A); Kr -= (long double)r[i] * logl((long double)rN[i]); }
    k.Kr = (double)Kr;
    k.A = (double)A; k.Rtot = (double)Rt; k.K0 = (double)K0; k.tau = tau; k.m = m; k.NR = 64; k.any_zero_rN = 0;
    Groups3 G; G.G = 0;
    double S1 = 0, S2 = 0;
    for (int i = 0; i < m; ++i) {
        int j = 0;
        for (; j < G.G; ++j) if (G.a[j] == a[i] && G.b[j] == b[i]) break;
        if (j == G.G) { G.a[j] = a[i]; G.b[j] = b[i]; G.R[j] = 0; ++G.G; }
        G.R[j] += r[i];
        S1 += rN[i] * a[i]; S2 += rN[i] * b[i];
    }
    // order groups by (b, a) like the device (row index order: a fastest)
    for (int x = 0; x < G.G; ++x) for (int y = x + 1; y < G.G; ++y)
        if (G.b[y] < G.b[x] || (G.b[y] == G.b[x] && G.a[y] < G.a[x])) {
            std::swap(G.a[x], G.a[y]); std::swap(G.b[x], G.b[y]); std::swap(G.R[x], G.R[y]);
        }
    Solve3Out o;
    uint8_t row[1024], ra[1024], rb[1024];
    for (int i = 0; i < m; ++i) { row[i] = (uint8_t)i; ra[i] = (uint8_t)a[i]; rb[i] = (uint8_t)b[i]; }
    const ref3::Ctx rc{m, r, rN, row, 1, ra, rb, (double)tau, (double)tau * k.A, S1, S2};
    n3_solve_auto(G, S1, S2, k, rc, mode, o);
    for (int j = 0; j < 3; ++j) { mu[j] = o.mu[j]; muhat[j] = o.muhat[j]; }
    *nll = o.nll; *nfev = o.nfev;
    return o.status;
}
