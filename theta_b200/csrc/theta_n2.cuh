// theta_n2.cuh -- plan object and on-chip enumeration state for the n=2 search.
//
// Enumeration (python/Enumerator.py:119-152): the admissible tumour columns are the
// non-decreasing integer vectors lb_i <= c_i <= ub_i, visited in ascending colex
// order.  A column is held as its level counts q_v = #{i : c_i >= v}, v = 1..KT
// (q_KT <= ... <= q_1, qlo_v <= q_v <= qhi_v); colex order of c is lexicographic
// order of (q_KT, ..., q_1), so rank <-> column is a mixed-radix walk over the
// count tables of python/TimeEstimate.py:91-111 re-indexed by level.
#pragma once
#include <vector>

#include "theta_core.cuh"

namespace theta {

// Layout of the per-problem table blob (device global, staged to shared memory by each CTA):
//   double   SR [m+1]      SR[q]  = sum of the last q entries of r  (sorted order)
//   double   SRN[m+1]      SRN[q] = sum of the last q entries of rN
//   uint64   SW [KT][m+2]  SW[v-1][x] = sum_{q=x..qhi_v} W_v(q)   (W_v = completions of levels v-1..1)
//   int32    Z0 [m+1]      Z0[q]  = #{i < m-q : r_i == 0}
//   int32    qlo[KT], qhi[KT]
struct TableLayout {
    int m, KT;
    size_t off_SR, off_SRN, off_SW, off_Z0, off_qlo, off_qhi, bytes;
    __host__ __device__ TableLayout() {}
    __host__ __device__ TableLayout(int m_, int KT_) : m(m_), KT(KT_) {
        size_t o = 0;
        off_SR = o;  o += sizeof(double) * (m + 1);
        off_SRN = o; o += sizeof(double) * (m + 1);
        off_SW = o;  o += sizeof(uint64_t) * (size_t)KT * (m + 2);
        off_Z0 = o;  o += sizeof(int32_t) * (m + 1);
        off_qlo = o; o += sizeof(int32_t) * KT;
        off_qhi = o; o += sizeof(int32_t) * KT;
        bytes = (o + 15) & ~(size_t)15;
    }
};

struct Tables {  // pointers into the shared-memory copy
    const double* SR;
    const double* SRN;
    const uint64_t* SW;
    const int32_t* Z0;
    const int32_t* qlo;
    const int32_t* qhi;
    int m;
    __device__ __forceinline__ void bind(const unsigned char* base, const TableLayout& L) {
        SR = reinterpret_cast<const double*>(base + L.off_SR);
        SRN = reinterpret_cast<const double*>(base + L.off_SRN);
        SW = reinterpret_cast<const uint64_t*>(base + L.off_SW);
        Z0 = reinterpret_cast<const int32_t*>(base + L.off_Z0);
        qlo = reinterpret_cast<const int32_t*>(base + L.off_qlo);
        qhi = reinterpret_cast<const int32_t*>(base + L.off_qhi);
        m = L.m;
    }
};

// rank -> level counts.  q[v-1] holds q_v.
template <int KT>
__device__ __forceinline__ void n2_unrank(const Tables& T, uint64_t rank, int (&q)[KT]) {
    uint64_t rem = rank;
    int prev = 0;
#pragma unroll
    for (int v = KT; v >= 1; --v) {
        const uint64_t* sw = T.SW + (size_t)(v - 1) * (T.m + 2);
        int lo = max(prev, T.qlo[v - 1]);
        int hi = T.qhi[v - 1];
        const uint64_t base = sw[lo];
        int l = lo;
        while (l < hi) {  // smallest x with base - sw[x+1] > rem
            const int mid = (l + hi) >> 1;
            if (base - sw[mid + 1] > rem) hi = mid; else l = mid + 1;
        }
        rem -= base - sw[l];
        q[v - 1] = prev = l;
    }
}

// next column in enumeration order; returns false after the last one
template <int KT>
__device__ __forceinline__ bool n2_next(const Tables& T, int (&q)[KT]) {
    if (q[0] < T.qhi[0]) {  // common case: raise the last zero to one
        ++q[0];
        return true;
    }
    int vs = 0;
#pragma unroll
    for (int v = KT; v >= 2; --v)
        if (q[v - 1] < T.qhi[v - 1]) vs = v;
    if (vs == 0) return false;
#pragma unroll
    for (int v = KT; v >= 1; --v) {
        if (v == vs) ++q[v - 1];
        else if (v < vs) q[v - 1] = max(T.qlo[v - 1], q[v < KT ? v : KT - 1]);  // q[v] holds q_{v+1} (v < vs <= KT)
    }
    return true;
}

// grouped statistics of the column described by q
template <int KT>
__device__ __forceinline__ void n2_groups(const Tables& T, const int (&q)[KT], double (&R)[KT + 1], double& S1,
                                          bool& zero_r_c0) {
    double sr_prev = T.SR[T.m];  // q_0 = m
    double s1 = 0.0;
#pragma unroll
    for (int v = 1; v <= KT; ++v) {
        const double sr = T.SR[q[v - 1]];
        R[v - 1] = sr_prev - sr;
        sr_prev = sr;
        s1 += T.SRN[q[v - 1]];
    }
    R[KT] = sr_prev;  // SR[0] == 0
    S1 = s1;
    zero_r_c0 = T.Z0[q[0]] > 0;
}

}  // namespace theta

struct theta_n3_state;  // theta_n3.cu
int theta_n3_plan_init(theta_plan* pl);
void theta_n3_plan_free(theta_plan* pl);
// what the streaming NLL kernel needs from an n=3 plan: the row table and the problem constants
int theta_n3_tile_info(const theta_plan* pl, uint8_t* ra, uint8_t* rb, int* n_rows, theta::N2Const* kc);

// Host-side plan (opaque to C callers)
struct theta_plan {
    uint64_t uid = 0;  // unique per plan instance (a destroyed plan's address can be reused): keys cached CUDA graphs
    theta_n3_state* n3 = nullptr;
    int device = 0;
    int n = 2, m = 0, tau = 2, K = 0, KT = 0;
    double max_normal = 1.0;
    std::vector<int32_t> lower, upper;  // monotonised (Enumerator.py:90-113)
    std::vector<int64_t> r, rN;
    std::vector<unsigned char> h_tables;  // host copy of the table blob (host-side unranking of the winners)
    uint64_t count = 0;
    theta::N2Const kc;
    theta::TableLayout layout;
    unsigned char* d_arena = nullptr;   // one allocation: tables | r | rN | PR | PRN | ZP | counters
    unsigned char* d_tables = nullptr;
    // tile-kernel tables (sorted-order per-interval vectors as doubles / ints)
    double* d_r = nullptr;    // [m]
    double* d_rN = nullptr;   // [m]
    int64_t tile_pitch = 0;   // theta_tile_pitch(m)
    double* d_PR = nullptr;   // [tile_pitch+1] prefix sums of r (flat past m)
    double* d_PRN = nullptr;  // [tile_pitch+1] prefix sums of rN
    int32_t* d_ZP = nullptr;  // [tile_pitch+1] prefix counts of r_i == 0
    // pass-1 state
    uint64_t p1_begin = 0, p1_end = 0;
    int p1_L = 0, p1_block = 0;
    uint64_t p1_stripes = 0;
    double* d_stripe_min = nullptr;
    size_t stripe_cap = 0;
    theta_cand* h_emit = nullptr;    // pinned [h_emit_cap]
    size_t h_emit_cap = 0;
    unsigned long long* h_ctr = nullptr;   // pinned snapshot of the device status block (theta_n2.cu: kStatusWords)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    void* p1_stream = nullptr;
    bool p1_pending = false;
    unsigned long long* d_counters = nullptr;  // [0] stripe ticket, [1] solved, [2] newton iters, [3] emitted, [4] nan-nll
    theta_cand* d_emit = nullptr;
    unsigned char* d_xblk = nullptr;  // [status | filtered records] of the single-shard pass 2 (device) ...
    unsigned char* h_xblk = nullptr;  // ... and its pinned copy
    unsigned char* d_fscratch = nullptr;  // records_filter_kernel: sort area for more records than fit shared memory
    size_t emit_cap = 0;
    double2* d_logtab = nullptr;  // fast_log table (theta_core.cuh)
    uint32_t* d_list = nullptr;   // pass-2 stripe list / per-stripe prior minima
    double* d_prior = nullptr;
    size_t list_cap = 0;
    theta_search_stats stats{};
};
