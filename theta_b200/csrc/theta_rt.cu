// theta_rt.cu -- runtime entry points: device selection, error reporting, the fp64
// roofline probe, and the single-matrix masked NLL of CalcAllC.L2 / L3.
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "theta_core.cuh"

namespace theta {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    cudaGetLastError();
    return THETA_ERR_CUDA;
}

// ---------------------------------------------------------------- caching allocators
namespace {
struct BlockCache {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_;      // (device, bucket) -> blocks
    std::unordered_map<void*, std::pair<int, size_t>> live_;         // block -> (device, bucket)
    bool pinned;
    explicit BlockCache(bool pinned_) : pinned(pinned_) {}
    static size_t bucket(size_t bytes) {
        if (bytes < 512) bytes = 512;
        if (bytes > (size_t)32 << 20) return (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);  // 2 MiB granules
        size_t b = 512;
        while (b < bytes) b <<= 1;
        return b;
    }
    cudaError_t get(void** p, size_t bytes) {
        *p = nullptr;
        int dev = 0;
        if (!pinned) { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
        const size_t b = bucket(bytes);
        std::lock_guard<std::mutex> lk(mu);
        auto& fl = free_[{dev, b}];
        if (!fl.empty()) { *p = fl.back(); fl.pop_back(); }
        else {
            cudaError_t e = pinned ? cudaHostAlloc(p, b, cudaHostAllocDefault) : cudaMalloc(p, b);
            if (e != cudaSuccess) {  // give cached blocks back and retry once
                for (auto& kv : free_) { for (void* q : kv.second) { if (pinned) cudaFreeHost(q); else cudaFree(q); } kv.second.clear(); }
                cudaGetLastError();
                e = pinned ? cudaHostAlloc(p, b, cudaHostAllocDefault) : cudaMalloc(p, b);
                if (e != cudaSuccess) return e;
            }
        }
        live_[*p] = {dev, b};
        return cudaSuccess;
    }
    void put(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        auto it = live_.find(p);
        if (it == live_.end()) { if (pinned) cudaFreeHost(p); else cudaFree(p); return; }
        free_[it->second].push_back(p);
        live_.erase(it);
    }
    void release() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv : free_) { for (void* q : kv.second) { if (pinned) cudaFreeHost(q); else cudaFree(q); } kv.second.clear(); }
    }
};
BlockCache& dcache() { static BlockCache* c = new BlockCache(false); return *c; }
BlockCache& hcache() { static BlockCache* c = new BlockCache(true); return *c; }
}  // namespace

cudaError_t dmalloc_bytes(void** p, size_t bytes) { return dcache().get(p, bytes); }
void dfree(void* p) { dcache().put(p); }
cudaError_t hmalloc_bytes(void** p, size_t bytes) { return hcache().get(p, bytes); }
void hfree(void* p) { hcache().put(p); }
void cache_release() { dcache().release(); hcache().release(); }

// fast_log table: r_i = 1/(1 + (i+0.5)/256) rounded to fp64 and -ln(r_i) of that very number
cudaError_t logtab_device(double2** out) {
    static std::mutex mu;
    static std::map<int, double2*> tabs;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    auto it = tabs.find(dev);
    if (it != tabs.end()) { *out = it->second; return cudaSuccess; }
    std::vector<double2> lt(THETA_LOGTAB_SIZE);
    for (int i = 0; i < THETA_LOGTAB_SIZE; ++i) {
        const double rinv = (double)(1.0L / (1.0L + ((long double)i + 0.5L) / 256.0L));
        lt[i].x = rinv;
        lt[i].y = (double)(-logl((long double)rinv));
    }
    double2* d = nullptr;
    e = cudaMalloc(&d, sizeof(double2) * THETA_LOGTAB_SIZE);
    if (e == cudaSuccess) e = cudaMemcpy(d, lt.data(), sizeof(double2) * THETA_LOGTAB_SIZE, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return e; }
    tabs[dev] = d;
    *out = d;
    return cudaSuccess;
}

// Dependent-free DFMA streams: 8 accumulators per thread, enough warps to saturate the fp64 pipe.
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// CalcAllC.L2 / L3 (python/CalcAllC.py:44-76): w_i = Cw[i] . mu ; D = sum_valid w_i ;
// NLL = - sum_valid r_i log(w_i / D) ; vals_i = w_i / D.  One CTA, fixed-order tree reduction.
__global__ void __launch_bounds__(256) nll_single_kernel(int n, int m, const double* __restrict__ mu,
                                                         const double* __restrict__ Cw, const double* __restrict__ r,
                                                         int l2_mu_zero, double* nll, double* vals, uint8_t* valid) {
    __shared__ double s_red[256];
    __shared__ double s_D;
    const int tid = threadIdx.x;
    double part = 0.0;
    for (int i = tid; i < m; i += blockDim.x) {
        double w = 0.0;
        for (int j = 0; j < n; ++j) w = fma(Cw[(size_t)i * n + j], mu[j], w);
        const bool ok = l2_mu_zero ? (Cw[(size_t)i * n + 1] != 0.0) : (Cw[(size_t)i * n] != 0.0);
        valid[i] = ok ? 1 : 0;
        vals[i] = w;
        if (ok) part += w;
    }
    s_red[tid] = part;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) s_red[tid] += s_red[tid + s];
        __syncthreads();
    }
    if (tid == 0) s_D = s_red[0];
    __syncthreads();
    const double D = s_D;
    part = 0.0;
    for (int i = tid; i < m; i += blockDim.x) {
        const double v = vals[i] / D;
        vals[i] = v;
        if (valid[i]) part += r[i] * log(v);
    }
    __syncthreads();
    s_red[tid] = part;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s) s_red[tid] += s_red[tid + s];
        __syncthreads();
    }
    if (tid == 0) *nll = -s_red[0];
}

}  // namespace theta

using namespace theta;

extern "C" const char* theta_last_error(void) { return g_err; }
extern "C" const char* theta_version(void) { return "theta_b200 0.1 (sm_100a)"; }

extern "C" int theta_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int theta_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); theta_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
        cudaGetLastError();
        return THETA_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return THETA_ERR_INVALID; }
    THETA_CUDA(cudaSetDevice(device));
    THETA_CUDA(cudaFree(0));
    return THETA_OK;
}

extern "C" void theta_shutdown(void) { cudaDeviceSynchronize(); theta::cache_release(); }

extern "C" int theta_fp64_peak(double* gflops, void* stream) {
    if (!gflops) { set_error("null argument"); return THETA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    THETA_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int block = 256, grid = sms * 8, iters = 1 << 14;
    theta::DevBlock blk;
    THETA_CUDA(blk.alloc(sizeof(double) * (size_t)grid * block));
    double* d = blk.as<double>();
    theta::EventPair ev;
    THETA_CUDA(ev.create());
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        THETA_CUDA(cudaEventRecord(e0, st));
        fp64_peak_kernel<<<grid, block, 0, st>>>(d, iters, 0.999999, 1e-9);
        THETA_CUDA(cudaEventRecord(e1, st));
        THETA_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * (double)iters * (double)grid * block;
        if (rep > 0 && ms > 0) best = fmax(best, fl / (ms * 1e-3) / 1e9);
    }
    THETA_CUDA(cudaGetLastError());
    *gflops = best;
    return THETA_OK;
}

extern "C" int theta_nll(int32_t n, int32_t m, const double* mu, const double* Cw, const int64_t* r, double* nll,
                         double* vals, uint8_t* valid) {
    if (!mu || !Cw || !r || !nll || !vals || !valid) { set_error("null argument"); return THETA_ERR_INVALID; }
    if ((n != 2 && n != 3) || m < 1) { set_error("bad shape n=%d m=%d", n, m); return THETA_ERR_INVALID; }
    double *d_mu = nullptr, *d_C = nullptr, *d_r = nullptr, *d_nll = nullptr, *d_vals = nullptr;
    uint8_t* d_valid = nullptr;
    std::vector<double> rd(m);
    for (int i = 0; i < m; ++i) rd[i] = (double)r[i];
    int rc = THETA_OK;
    cudaError_t e = theta::dmalloc(&d_mu, sizeof(double) * 3);
    if (e == cudaSuccess) e = theta::dmalloc(&d_C, sizeof(double) * (size_t)m * n);
    if (e == cudaSuccess) e = theta::dmalloc(&d_r, sizeof(double) * m);
    if (e == cudaSuccess) e = theta::dmalloc(&d_nll, sizeof(double));
    if (e == cudaSuccess) e = theta::dmalloc(&d_vals, sizeof(double) * m);
    if (e == cudaSuccess) e = theta::dmalloc(&d_valid, m);
    if (e == cudaSuccess) e = cudaMemcpy(d_mu, mu, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_C, Cw, sizeof(double) * (size_t)m * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_r, rd.data(), sizeof(double) * m, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        // CalcAllC.L2 picks the validity column by the truthiness of mu (CalcAllC.py:49-52)
        const int l2_mu_zero = (n == 2 && mu[0] == 0.0) ? 1 : 0;
        nll_single_kernel<<<1, 256>>>(n, m, d_mu, d_C, d_r, l2_mu_zero, d_nll, d_vals, d_valid);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(nll, d_nll, sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(vals, d_vals, sizeof(double) * m, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(valid, d_valid, m, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "theta_nll", __FILE__, __LINE__);
    theta::dfree(d_mu); theta::dfree(d_C); theta::dfree(d_r); theta::dfree(d_nll); theta::dfree(d_vals); theta::dfree(d_valid);
    return rc;
}
