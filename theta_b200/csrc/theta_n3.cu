// theta_n3.cu -- n = 3: on-chip DFS enumeration + KKT solve + the two search passes.
//
// Replaces, for n=3: Enumerator._create_graph / _generate_next_C_3 / _generate_next_C_3_recurse /
// _get_mu_bounds / _in_bounds (python/Enumerator.py:172-298), Optimizer._solve_n3plus
// (python/Optimizer.py:128-165) and the driver loop (python/RunTHetA.py:173-220).
//
// Work decomposition: the reference's DFS visits row sequences in lexicographic order of row index.
// The host cuts the DFS forest into *items* -- sub-trees given by a prefix of row indices -- using the
// count DP of TimeEstimate.count_number_matrices_3 (python/TimeEstimate.py:113-142) per sub-tree, and
// lists them in lexicographic order, so item order == enumeration order.  Two kernels alternate:
//   n3_walk_kernel         every thread takes items from a ticket counter and walks its sub-tree depth first with the
//                          reference's pruning rules, keeping the grouped statistics (R_g, S1, S2) incrementally (exact
//                          integer sums in fp64), and screens each leaf with a rigorous lower bound of its NLL against
//                          the best exact NLL known; survivors are queued;
//   n3_solve_queue_kernel  solves the queued survivors (grouped-statistics hybrj, or the reference-order solver for
//                          rank-deficient candidates) and lowers the best NLL.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "theta_n2.cuh"
#include "theta_n3.cuh"

struct theta_n3_state {
    int NR = 0, k = 0;
    std::vector<int> ra, rb;            // rows (Enumerator._create_graph order)
    std::vector<uint64_t> edges;        // edges[v] bit w: w may follow v (Enumerator._is_valid_edge)
    std::vector<uint64_t> allow;        // allow[i] bit g: row g within [lower_i, upper_i] (Enumerator._in_bounds)
    uint64_t symmask = 0, eqmask = 0;   // rows with a <= b / a == b (duplicate-column rule, Enumerator.py:180-183,199-202)
    theta::N3Const kc;
    // items
    std::vector<uint8_t> prefix;
    std::vector<uint32_t> off;          // [n_items + 1]
    std::vector<double> weight;
    // device
    uint64_t* d_allow = nullptr;
    double* d_r = nullptr;
    double* d_rN = nullptr;
    uint8_t* d_prefix = nullptr;
    uint32_t* d_off = nullptr;
    unsigned long long* d_ctr = nullptr;
    std::vector<std::vector<double>> est;  // est[i][g*2+sw]: completions of a prefix ending in row g at interval i (no mu-bounds)
    // survivor queue of the last search (kept for theta_search_n3_window)
    uint32_t* d_q_meta = nullptr;
    double* d_q_lb = nullptr;
    theta_cand* d_q_res = nullptr;
    uint8_t* d_q_path = nullptr;
    size_t q_cap = 0;
    uint64_t q_count = 0;
    double window = 0.0;
    uint32_t* d_ref_list = nullptr;  // queue slots left to the warp-cooperative reference-order solver ([q_cap])
    unsigned long long* d_ctr_rows = nullptr;  // counters of theta_solve_rows_n3 (kCtrWords)
    uint32_t* d_list = nullptr;   // wave-0 item list of the range [list_begin, list_end)
    uint64_t list_begin = 0, list_end = 0;
    uint32_t list_n0 = 0;
    uint64_t p1_begin = 0, p1_end = 0;
    unsigned long long nan_nll = 0;
    unsigned long long it_hist[3] = {0, 0, 0};
    bool truncated = false;
};

namespace theta {

static constexpr int kBlock3 = 64;
static constexpr int kScreenDefault = 2;  // fp64 table logarithms: 1% (m=500) .. 5% (m=150) of the leaves survive, against 19% .. 9% with fp32

struct N3Dev {
    const uint64_t* allow;  // [m]
    const double* r;        // [m]
    const double* rN;       // [m]
    uint64_t edges[kMaxRows3];
    uint8_t ra[kMaxRows3], rb[kMaxRows3];
    uint64_t symmask, eqmask;
    N3Const kc;
    const double2* logtab;  // fast_log table (global); staged to shared memory when the fp64 screen is on
};

// ratio codes: the reference keeps [minVal, maxVal] bounds on mu_1/mu_2 as floats yChange/(-xChange)
// (Enumerator._get_mu_bounds, Enumerator.py:225-239); x, y are differences of copy numbers in [-7, 7].
__host__ __device__ inline uint8_t ratio_code(int x, int y) { return (uint8_t)(1 + (y + 7) * 15 + (x + 7)); }
__host__ __device__ inline double ratio_val(uint8_t c) {
    if (c == 0) return -INFINITY;
    if (c == 255) return INFINITY;
    const int q = (int)c - 1;
    const int y = q / 15 - 7, x = q % 15 - 7;
    return (double)y / (double)(-x);
}

// One step of _generate_next_C_3_recurse (Enumerator.py:189-214): can `child` follow `prev` given the
// bounds (mn, mx)?  Updates the bounds.  (`if newMin:` / `elif newMax:` are always taken for a non-zero
// change: the returned ratios are never 0 when both differences are non-zero.)
__host__ __device__ inline bool step_bounds(int pa, int pb, int ca, int cb, uint8_t& mn, uint8_t& mx) {
    const int x = ca - pa, y = cb - pb;
    if (x == 0 || y == 0) return ratio_val(mn) <= ratio_val(mx);
    const uint8_t c = ratio_code(x, y);
    if (x > 0) {
        if (ratio_val(c) > ratio_val(mn)) mn = c;  // max(minVal, newMin)
    } else {
        if (ratio_val(c) < ratio_val(mx)) mx = c;  // min(maxVal, newMax)
    }
    return ratio_val(mn) <= ratio_val(mx);
}

// ======================================================================= device: walk + screen (kernel A)
//
// One thread per work item (a DFS sub-tree), items taken from a ticket counter.  Everything a thread touches while
// walking lives in registers or in shared memory laid out [slot][thread] (bank = thread: conflict-free whatever
// depth each lane is at) -- no local memory:
//   path   one byte per interval, four per 32-bit word: row index | 0x80 if that push changed (mn, mx, switch)
//   undo   the (mn, mx, switch) values replaced by such pushes (kUndo entries; deeper: recomputed by replay)
//   R[g]   reads per row type (exact integers in fp64), plus a 64-bit "present" mask in a register
// Leaves are screened against T + margin (T: the best exact NLL known when the wave started) or a fixed threshold;
// survivors go to a global queue (path + lower bound) that the solve kernel (kernel B) drains.  The 33 k-instruction
// solver is never part of this kernel, so it runs at ~100 registers with no block-level synchronisation.
static constexpr int kUndo = 12;

struct WalkParams {
    N3Dev D;
    const uint8_t* prefix;
    const uint32_t* off;
    const uint32_t* list;        // explicit item list (NULL: item = item_begin + ticket)
    uint32_t item_begin;
    uint32_t n_work;
    unsigned long long* ctr;     // kCtr* below
    double margin;               // threshold = min(fixed_thr, T + margin)
    double fixed_thr;
    uint32_t* q_meta;            // [cap][2]: item, leaf ordinal inside the item
    double* q_lb;                // [cap] screening bound of the survivor (verification)
    uint8_t* q_path;             // [cap][m]
    unsigned long long q_cap;
    int enumerate_only;          // every leaf is a survivor, nothing is screened
    int screen;                  // 0: off, 1: fp32 logarithms, 2: fp64 table logarithms
    int verify;                  // THETA_VERIFY_APPROX=1: every leaf is solved, bounds are checked against the exact NLL
    int block;
    uint32_t scout_stride;       // > 0: scout wave -- every scout_stride-th leaf of each item gets the full Newton bound,
                                 // the leaf with the lowest bound is queued (slot = ticket), nothing else
    int count_leaves;            // add the leaves walked to the candidate count (off in the scout wave)
};
enum { kCtrTicket = 0, kCtrLeaves = 1, kCtrQueue = 2, kCtrDone = 3, kCtrSolved = 4, kCtrNfev = 5, kCtrNan = 6, kCtrViol = 7,
       kCtrT = 8, kCtrIt0 = 9, kCtrIt1 = 10, kCtrIt2 = 11, kCtrRef = 12, kCtrRefQ = 13, kCtrRefDone = 14, kCtrWords = 16 };

struct Walk {
    // shared-memory views of this thread (stride B between slots)
    uint32_t* pw;
    uint32_t* undo;
    double* R;
    const uint64_t* allow;
    const double* r;
    const double* rN;
    const float* ratio;
    const N3Dev* D;
    int B;
    double S1, S2;
    uint64_t present;
    int depth, base, nundo, m;
    uint8_t mn, mx, sw;

    __device__ __forceinline__ int get(int pos) const { return (int)((pw[(pos >> 2) * B] >> ((pos & 3) * 8)) & 0xffu); }
    __device__ __forceinline__ void set(int pos, int v) {
        uint32_t& w = pw[(pos >> 2) * B];
        const int sh = (pos & 3) * 8;
        w = (w & ~(0xffu << sh)) | ((uint32_t)v << sh);
    }
    __device__ __forceinline__ int row(int pos) const { return get(pos) & 0x3f; }
    // Enumerator._generate_next_C_3_recurse's bound update (step_bounds above) on the ratio table
    __device__ __forceinline__ bool step(int pg, int cg, uint8_t& mnc, uint8_t& mxc) const {
        const int x = (int)D->ra[cg] - (int)D->ra[pg], y = (int)D->rb[cg] - (int)D->rb[pg];
        if (x != 0 && y != 0) {
            const uint8_t c = ratio_code(x, y);
            if (x > 0) { if (ratio[c] > ratio[mnc]) mnc = c; }
            else { if (ratio[c] < ratio[mxc]) mxc = c; }
        }
        return ratio[mnc] <= ratio[mxc];
    }
    __device__ __forceinline__ void push(int pos, int g, uint8_t mnc, uint8_t mxc, uint8_t s) {
        const bool changed = (mnc != mn) | (mxc != mx) | (s != sw);
        if (changed) {
            if (nundo < kUndo) undo[nundo * B] = (uint32_t)mn | ((uint32_t)mx << 8) | ((uint32_t)sw << 16);
            ++nundo;
            mn = mnc; mx = mxc; sw = s;
        }
        set(pos, g | (changed ? 0x80 : 0));
        R[g * B] += r[pos];
        present |= 1ull << g;
        S1 += rN[pos] * (double)D->ra[g];
        S2 += rN[pos] * (double)D->rb[g];
        depth = pos;
    }
    __device__ void replay(int upto) {  // (mn, mx, switch) after positions 0..upto; only when the undo log overflowed
        uint8_t mnc = 0, mxc = 255, s = 0;
        for (int i = 0; i <= upto; ++i) {
            const int g = row(i);
            if (i == 0) s = (D->ra[g] == D->rb[g]) ? 1 : 0;
            else {
                step(row(i - 1), g, mnc, mxc);
                s = (s && D->ra[g] == D->rb[g]) ? 1 : 0;
            }
        }
        mn = mnc; mx = mxc; sw = s;
    }
    __device__ __forceinline__ void pop() {
        const int pos = depth, v = get(pos), g = v & 0x3f;
        double& Rg = R[g * B];
        Rg -= r[pos];
        if (Rg == 0.0) present &= ~(1ull << g);
        S1 -= rN[pos] * (double)D->ra[g];
        S2 -= rN[pos] * (double)D->rb[g];
        if (v & 0x80) {
            --nundo;
            if (nundo < kUndo) {
                const uint32_t e = undo[nundo * B];
                mn = (uint8_t)(e & 0xff); mx = (uint8_t)((e >> 8) & 0xff); sw = (uint8_t)((e >> 16) & 1);
            } else {
                replay(pos - 1);
            }
        }
        depth = pos - 1;
    }
    __device__ __forceinline__ uint64_t children() const {
        uint64_t c = D->edges[row(depth)] & allow[depth + 1];
        if (sw) c &= D->symmask;
        return c;
    }
    __device__ void load_prefix(const uint8_t* pre, int len) {
        for (int g = 0; g < D->kc.NR; ++g) R[g * B] = 0.0;
        present = 0;
        S1 = S2 = 0.0;
        nundo = 0;
        mn = 0; mx = 255; sw = 0;
        depth = -1;
        for (int i = 0; i < len; ++i) {
            const int g = pre[i];
            uint8_t mnc = mn, mxc = mx, s;
            if (i == 0) s = (D->ra[g] == D->rb[g]) ? 1 : 0;
            else {
                step(row(i - 1), g, mnc, mxc);
                s = (sw && D->ra[g] == D->rb[g]) ? 1 : 0;
            }
            push(i, g, mnc, mxc, s);
        }
        base = len - 1;
    }
    // advance to the next leaf of the sub-tree; `fresh` = nothing visited yet in this item
    __device__ bool next_leaf(bool fresh) {
        const int last = m - 1;
        uint64_t cand;
        if (fresh) {
            if (depth == last) return true;  // the prefix itself is a complete candidate
            cand = children();
        } else {
            if (depth == base) return false;  // item was a single leaf
            const int g = row(depth);
            pop();
            cand = children() & ~((2ull << g) - 1ull);
        }
        for (;;) {
            bool pushed = false;
            const int pg = row(depth);
            while (cand) {
                const int c = __ffsll((long long)cand) - 1;
                cand &= cand - 1;
                uint8_t mnc = mn, mxc = mx;
                if (step(pg, c, mnc, mxc)) {
                    const uint8_t s = (sw && D->ra[c] == D->rb[c]) ? 1 : 0;
                    push(depth + 1, c, mnc, mxc, s);
                    pushed = true;
                    break;
                }
            }
            if (pushed) {
                if (depth == last) return true;
                cand = children();
            } else {
                if (depth == base) return false;
                const int g = row(depth);
                pop();
                cand = children() & ~((2ull << g) - 1ull);
            }
        }
    }
};

// n3_screen (theta_n3.cuh) on the walker's state: groups are the set bits of `present`.
__device__ __forceinline__ bool walk_screen(const Walk& w, const N3Const& k, Screen3& ws, double thr, double& lb_out,
                                            const double2* LT, int& iters) {
    lb_out = -INFINITY;
    iters = 0;
    const double S1 = w.S1, S2 = w.S2;
    if (k.any_zero_rN || !(S1 > 0.0) || !(S2 > 0.0)) return false;
    const double t1 = k.A / S1, t2 = k.A / S2;  // A d_g = y0 + a (A/S1) y1 + b (A/S2) y2
    const double E = (LT ? 1e-11 : THETA_N3_APPROX_C) * k.Rtot;
    double y0 = ws.y0, y1 = ws.y1;
    if (!(y0 > 1e-6 && y1 > 1e-6 && y0 + y1 < 1.0 - 1e-6)) { y0 = 1.0 / 3.0; y1 = 1.0 / 3.0; }
    for (int it = 0; it < 3; ++it) {
        iters = it + 1;
        const double y2 = 1.0 - (y0 + y1);
        const double u1 = y1 * t1, u2 = y2 * t2;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, h00 = 0.0, h01 = 0.0, h02 = 0.0, h11 = 0.0, h12 = 0.0, h22 = 0.0, lacc = 0.0;
        bool ok = true;
        for (uint64_t msk = w.present; msk; msk &= msk - 1) {
            const int g = __ffsll((long long)msk) - 1;
            const double a = (double)w.D->ra[g], b = (double)w.D->rb[g], R = w.R[g * w.B];
            const double d = y0 + a * u1 + b * u2;  // A d_g > 0 in the interior
            const double r = fast_rcp1(d);
            const double q = R * r, qq = q * r, aq = a * qq, bq = b * qq;
            s0 += q; s1 = fma(a, q, s1); s2 = fma(b, q, s2);
            h00 += qq; h01 += aq; h02 += bq; h11 = fma(a, aq, h11); h12 = fma(a, bq, h12); h22 = fma(b, bq, h22);
            if (LT) {
                lacc = fma(R, fast_log(d, LT), lacc);
            } else {
                const float df = (float)d;
                if (!(df >= 3.0517578125e-05f && df <= 32768.0f)) ok = false;
                lacc = fma(R, (double)__log2f(df), lacc);
            }
        }
        if (!ok) return false;
        const double G0 = s0, G1 = t1 * s1, G2 = t2 * s2;
        const double gap = fmax(G0, fmax(G1, G2)) - k.Rtot;
        const double fy = LT ? k.K0 - lacc : fma(-0.6931471805599453094, lacc, k.K0);
        const double lb = fy - fmax(gap, 0.0) * (1.0 + 1e-9) - E;
        if (lb > lb_out) lb_out = lb;
        if (lb > thr) { ws.y0 = y0; ws.y1 = y1; return true; }
        if (it == 2) break;
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

__host__ __device__ inline size_t walk_smem_fixed(int m) { return (size_t)((m + 1) & ~1) * 24 + sizeof(double2) * THETA_LOGTAB_SIZE + 256 * sizeof(float); }
__host__ __device__ inline size_t walk_smem_per_thread(int m, int NR) { return (size_t)NR * 8 + (size_t)((m + 3) / 4) * 4 + kUndo * 4; }

__global__ void n3_walk_kernel(const WalkParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ N3Dev D;
    const int tid = threadIdx.x, B = p.block;
    const int m = p.D.kc.m, NR = p.D.kc.NR, mp = (m + 1) & ~1, PW = (m + 3) / 4;
    uint64_t* s_allow = reinterpret_cast<uint64_t*>(smem);
    double* s_r = reinterpret_cast<double*>(s_allow + mp);
    double* s_rN = s_r + mp;
    double2* s_LT = reinterpret_cast<double2*>(s_rN + mp);
    float* s_ratio = reinterpret_cast<float*>(s_LT + THETA_LOGTAB_SIZE);
    double* s_R = reinterpret_cast<double*>(s_ratio + 256);
    uint32_t* s_path = reinterpret_cast<uint32_t*>(s_R + (size_t)NR * B);
    uint32_t* s_undo = s_path + (size_t)PW * B;
    if (tid == 0) D = p.D;
    for (int i = tid; i < m; i += B) { s_allow[i] = p.D.allow[i]; s_r[i] = p.D.r[i]; s_rN[i] = p.D.rN[i]; }
    for (int i = tid; i < THETA_LOGTAB_SIZE; i += B) s_LT[i] = p.D.logtab[i];
    for (int i = tid; i < 256; i += B) s_ratio[i] = (float)ratio_val((uint8_t)i);
    __syncthreads();
    const double2* LT = p.screen == 2 ? s_LT : nullptr;
    Walk w;
    w.pw = s_path + tid; w.undo = s_undo + tid; w.R = s_R + tid;
    w.allow = s_allow; w.r = s_r; w.rN = s_rN; w.ratio = s_ratio; w.D = &D; w.B = B; w.m = m;
    const double T = __longlong_as_double((long long)p.ctr[kCtrT]);  // fixed for the launch: the solve kernel runs between waves
    const double thr = fmin(p.fixed_thr, T + p.margin);
    Screen3 ws{1.0 / 3.0, 1.0 / 3.0};
    unsigned long long leaves = 0, it1 = 0, it2 = 0, it3 = 0;
    if (p.scout_stride) {
        // scout wave: which leaf of an item looks best?  Every scout_stride-th leaf (an unbiased sample over the item)
        // gets the full Newton bound; the exact NLL of the chosen leaves gives the first threshold T.
        for (;;) {
            const uint32_t ticket = (uint32_t)atomicAdd(&p.ctr[kCtrTicket], 1ULL);
            if (ticket >= p.n_work) break;
            const uint32_t item = p.list ? p.list[ticket] : p.item_begin + ticket;
            w.load_prefix(p.prefix + p.off[item], (int)(p.off[item + 1] - p.off[item]));
            uint32_t seq = 0;
            bool fresh = true;
            double best_lb = INFINITY;
            p.q_meta[2 * (size_t)ticket] = item;
            p.q_meta[2 * (size_t)ticket + 1] = 0xffffffffu;  // "no leaf"
            p.q_lb[ticket] = -INFINITY;
            const uint32_t phase = ticket % p.scout_stride;
            while (w.next_leaf(fresh)) {
                fresh = false;
                if (seq % p.scout_stride == phase) {
                    double lb;
                    int iters;
                    walk_screen(w, D.kc, ws, INFINITY, lb, LT, iters);
                    if (lb < best_lb || best_lb == INFINITY) {
                        best_lb = lb;
                        p.q_meta[2 * (size_t)ticket + 1] = seq;
                        uint8_t* dst = p.q_path + (size_t)ticket * m;
                        for (int i = 0; i < m; ++i) dst[i] = (uint8_t)w.row(i);
                    }
                }
                ++seq;
            }
            atomicMax(&p.ctr[kCtrQueue], (unsigned long long)ticket + 1ULL);
        }
        return;
    }
    // Warp-synchronous leaf loop: every lane first advances its own DFS to its next leaf (divergent, cheap), the warp
    // reconverges, and all lanes that have a leaf screen it together (the expensive, regular part).
    bool have_item = false, done = false, fresh = true;
    uint32_t item = 0, seq = 0;
    for (;;) {
        bool have_leaf = false;
        while (!have_leaf && !done) {
            if (!have_item) {
                const uint32_t ticket = (uint32_t)atomicAdd(&p.ctr[kCtrTicket], 1ULL);
                if (ticket >= p.n_work) { done = true; break; }
                item = p.list ? p.list[ticket] : p.item_begin + ticket;
                w.load_prefix(p.prefix + p.off[item], (int)(p.off[item + 1] - p.off[item]));
                seq = 0;
                fresh = true;
                have_item = true;
            }
            if (w.next_leaf(fresh)) { fresh = false; have_leaf = true; }
            else have_item = false;
        }
        __syncwarp();
        if (__all_sync(0xffffffffu, !have_leaf)) break;
        if (have_leaf) {
            ++leaves;
            bool keep = true;
            double lb = -INFINITY;
            if (!p.enumerate_only && p.screen) {
                int iters;
                keep = !walk_screen(w, D.kc, ws, thr, lb, LT, iters) || p.verify;
                it1 += iters <= 1; it2 += iters == 2; it3 += iters >= 3;
            }
            if (keep) {
                const unsigned long long slot = atomicAdd(&p.ctr[kCtrQueue], 1ULL);
                if (slot < p.q_cap) {
                    p.q_meta[2 * slot] = item;
                    p.q_meta[2 * slot + 1] = seq;
                    p.q_lb[slot] = lb;
                    uint8_t* dst = p.q_path + slot * (size_t)m;
                    for (int i = 0; i < m; ++i) dst[i] = (uint8_t)w.row(i);
                }
            }
            ++seq;
        }
    }
    if (leaves && p.count_leaves) atomicAdd(&p.ctr[kCtrLeaves], leaves);
    if (it1) atomicAdd(&p.ctr[kCtrIt0], it1);
    if (it2) atomicAdd(&p.ctr[kCtrIt1], it2);
    if (it3) atomicAdd(&p.ctr[kCtrIt2], it3);
}

// ======================================================================= device: solve the survivors (kernel B)
struct SolveParams {
    N3Dev D;
    unsigned long long* ctr;
    const uint32_t* q_meta;
    const double* q_lb;
    const uint8_t* q_path;
    theta_cand* q_res;           // [cap]: rank = (item << 32) | leaf ordinal, flags = queue slot
    unsigned long long q_cap;
    uint32_t* ref_list;          // [cap] slots left to the warp-cooperative reference-order solver (n3_solve_ref_kernel)
    int mode;                    // THETA_N3_MODE_*
    int enumerate_only;
    int skip_ref;                // scout wave: rank-deficient candidates are not needed for the threshold -- leave them out
};

__global__ void __launch_bounds__(kBlock3) n3_solve_queue_kernel(const SolveParams p) {
    __shared__ N3Dev D;
    if (threadIdx.x == 0) D = p.D;
    __syncthreads();
    const unsigned long long count = min(p.ctr[kCtrQueue], p.q_cap), done = p.ctr[kCtrDone];
    const int m = D.kc.m;
    unsigned long long solved = 0, nfev = 0, nnan = 0, viol = 0;
    double best = INFINITY;
    for (unsigned long long slot = done + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; slot < count;
         slot += (unsigned long long)gridDim.x * blockDim.x) {
        const uint8_t* pth = p.q_path + slot * (size_t)m;
        theta_cand c;
        c.rank = ((uint64_t)p.q_meta[2 * slot] << 32) | p.q_meta[2 * slot + 1];
        c.flags = (int32_t)slot;
        c.nll = 0.0; c.mu[0] = c.mu[1] = c.mu[2] = 0.0; c.status = THETA_ST_NONE_NAN;
        if (!p.enumerate_only && p.q_meta[2 * slot + 1] != 0xffffffffu) {
            double R[kMaxRows3];
            uint16_t cnt[kMaxRows3];
            for (int g = 0; g < D.kc.NR; ++g) { R[g] = 0.0; cnt[g] = 0; }
            double S1 = 0.0, S2 = 0.0;
            for (int j = 0; j < m; ++j) {
                const int g = pth[j];
                R[g] += D.r[j];
                cnt[g] += 1;
                S1 += D.rN[j] * (double)D.ra[g];
                S2 += D.rN[j] * (double)D.rb[g];
            }
            Groups3 G;
            int n = 0;
            for (int g = 0; g < D.kc.NR; ++g)
                if (cnt[g]) { G.R[n] = R[g]; G.a[n] = D.ra[g]; G.b[n] = D.rb[g]; ++n; }
            G.G = n;
            const bool ref = p.mode == THETA_N3_MODE_REF ||
                             (p.mode == THETA_N3_MODE_AUTO && (D.kc.any_zero_rN || ref3::rank_deficient(G.G, G.R, G.a, G.b, S1, S2)));
            if (ref) {  // one warp per candidate does these (~30x less latency than one thread)
                if (!p.skip_ref) {
                    const unsigned long long k = atomicAdd(&p.ctr[kCtrRefQ], 1ULL);
                    p.ref_list[k] = (uint32_t)slot;
                }
                c.status = -1;
            } else {
                Solve3Out o;
                n3_solve(G, S1, S2, D.kc, o);
                c.nll = o.nll; c.mu[0] = o.mu[0]; c.mu[1] = o.mu[1]; c.mu[2] = o.mu[2]; c.status = o.status;
                if (o.status < THETA_ST_NONE_SIGN) {
                    ++solved;
                    nfev += o.nfev;
                    if (o.nll != o.nll) ++nnan;
                    else {
                        best = fmin(best, o.nll);
                        if (p.q_lb[slot] > o.nll) ++viol;
                    }
                }
            }
        }
        p.q_res[slot] = c;
    }
    // NLL >= 0, so the bit pattern orders like the value; +inf (nothing solved) never lowers T
    if (best < INFINITY && best >= 0.0) atomicMin(&p.ctr[kCtrT], (unsigned long long)__double_as_longlong(best));
    if (solved) atomicAdd(&p.ctr[kCtrSolved], solved);
    if (nfev) atomicAdd(&p.ctr[kCtrNfev], nfev);
    if (nnan) atomicAdd(&p.ctr[kCtrNan], nnan);
    if (viol) atomicAdd(&p.ctr[kCtrViol], viol);
}

// Reference-order solves, one warp per candidate (n3_ref.cuh::WarpExec).  Work list: slots [RefDone, RefQ) of ref_list.
// Results go either to the search queue (q_res + counters + T) or to the arrays of theta_solve_rows_n3.
struct RefParams {
    N3Dev D;
    unsigned long long* ctr;
    const uint32_t* ref_list;
    const uint8_t* paths;        // [slot][m]
    theta_cand* q_res;           // search mode (or NULL)
    const double* q_lb;
    double* mu;                  // rows mode: [slot][3], [slot][3], [slot], [slot], [slot][m] or NULL
    double* muhat;
    double* nll;
    int32_t* status;
    double* vals;
    int warps;                   // per CTA
};

__global__ void n3_solve_ref_kernel(const RefParams p) {
    extern __shared__ __align__(16) double ref_scratch[];
    __shared__ N3Dev D;
    if (threadIdx.x == 0) D = p.D;
    __syncthreads();
    const int m = D.kc.m, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long count = p.ctr[kCtrRefQ], done = p.ctr[kCtrRefDone];
    ref3::WarpExec ex{ref_scratch + (size_t)wid * 9 * m, m};
    const ref3::StatusCodes sc{THETA_ST_OK, THETA_ST_FALLBACK_X0, THETA_ST_FALLBACK_MOVED, THETA_ST_NONE_RANGE};
    for (unsigned long long idx = done + (unsigned long long)blockIdx.x * p.warps + wid; idx < count;
         idx += (unsigned long long)gridDim.x * p.warps) {
        const uint32_t slot = p.ref_list[idx];
        const uint8_t* pth = p.paths + (size_t)slot * m;
        double S1 = 0.0, S2 = 0.0;  // exact integers: any summation order gives the same double
        for (int j = lane; j < m; j += 32) {
            const int g = pth[j];
            S1 += D.rN[j] * (double)D.ra[g];
            S2 += D.rN[j] * (double)D.rb[g];
        }
        for (int o = 16; o; o >>= 1) {
            S1 += __shfl_xor_sync(0xffffffffu, S1, o);
            S2 += __shfl_xor_sync(0xffffffffu, S2, o);
        }
        const ref3::Ctx rc{m, D.r, D.rN, pth, 1, D.ra, D.rb, D.kc.tau, D.kc.tau * D.kc.A, S1, S2};
        ref3::Out o;
        ref3::solve(rc, sc, ex, o, p.vals ? p.vals + (size_t)slot * m : nullptr);
        if (lane == 0) {
            atomicAdd(&p.ctr[kCtrRef], 1ULL);
            if (p.q_res) {
                theta_cand& c = p.q_res[slot];
                c.nll = o.nll; c.mu[0] = o.mu[0]; c.mu[1] = o.mu[1]; c.mu[2] = o.mu[2]; c.status = o.status;
                if (o.status < THETA_ST_NONE_SIGN) {
                    atomicAdd(&p.ctr[kCtrSolved], 1ULL);
                    atomicAdd(&p.ctr[kCtrNfev], (unsigned long long)o.nfev);
                    if (o.nll != o.nll) atomicAdd(&p.ctr[kCtrNan], 1ULL);
                    else {
                        if (o.nll >= 0.0 && o.nll < INFINITY) atomicMin(&p.ctr[kCtrT], (unsigned long long)__double_as_longlong(o.nll));
                        if (p.q_lb[slot] > o.nll) atomicAdd(&p.ctr[kCtrViol], 1ULL);
                    }
                }
            } else {
                for (int j = 0; j < 3; ++j) {
                    p.mu[(size_t)slot * 3 + j] = o.mu[j];
                    if (p.muhat) p.muhat[(size_t)slot * 3 + j] = o.muhat[j];
                }
                p.nll[slot] = o.nll;
                p.status[slot] = o.status;
            }
        }
        __syncwarp();
    }
}

// between waves: everything queued so far has been solved; the next wave starts a fresh ticket count
__global__ void n3_advance_kernel(unsigned long long* ctr, unsigned long long cap) {
    ctr[kCtrDone] = min(ctr[kCtrQueue], cap);
    ctr[kCtrRefDone] = ctr[kCtrRefQ];
    ctr[kCtrTicket] = 0;
}

// Optimizer.solve on explicit row-index strings
__global__ void __launch_bounds__(kBlock3) n3_solve_rows_kernel(const N3Dev D0, const uint8_t* paths, int64_t count,
                                                                double* mu, double* muhat, double* nll, int32_t* status,
                                                                double* vals, int mode, unsigned long long* ctr,
                                                                uint32_t* ref_list) {
    __shared__ N3Dev D;
    if (threadIdx.x == 0) D = D0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int m = D.kc.m;
    double R[kMaxRows3];
    uint16_t cnt[kMaxRows3];
    for (int g = 0; g < D.kc.NR; ++g) { R[g] = 0.0; cnt[g] = 0; }
    double S1 = 0.0, S2 = 0.0;
    const uint8_t* pth = paths + i * (size_t)m;
    bool bad = false;
    for (int j = 0; j < m; ++j) {
        const int g = pth[j];
        if (g >= D.kc.NR) { bad = true; continue; }
        R[g] += D.r[j];
        cnt[g] += 1;
        S1 += D.rN[j] * (double)D.ra[g];
        S2 += D.rN[j] * (double)D.rb[g];
    }
    Groups3 G;
    int n = 0;
    for (int g = 0; g < D.kc.NR; ++g)
        if (cnt[g]) { G.R[n] = R[g]; G.a[n] = D.ra[g]; G.b[n] = D.rb[g]; ++n; }
    G.G = n;
    Solve3Out o;
    if (bad) { o.status = THETA_ST_NONE_NAN; o.nll = 0.0; o.mu[0] = o.mu[1] = o.mu[2] = 0.0; o.muhat[0] = o.muhat[1] = o.muhat[2] = 0.0; }
    else {
        const bool ref = mode == THETA_N3_MODE_REF ||
                         (mode == THETA_N3_MODE_AUTO && (D.kc.any_zero_rN || ref3::rank_deficient(G.G, G.R, G.a, G.b, S1, S2)));
        if (ref) {  // left to n3_solve_ref_kernel, which writes mu / muhat / nll / status / vals of this candidate
            ref_list[atomicAdd(&ctr[kCtrRefQ], 1ULL)] = (uint32_t)i;
            return;
        }
        n3_solve(G, S1, S2, D.kc, o);
    }
    for (int j = 0; j < 3; ++j) {
        mu[i * 3 + j] = o.mu[j];
        if (muhat) muhat[i * 3 + j] = o.muhat[j];
    }
    nll[i] = o.nll;
    status[i] = o.status;
    if (vals && o.status < THETA_ST_NONE_SIGN && o.status != THETA_ST_DEGENERATE) {
        // vals_i = numer_i / denom of Optimizer.L3 (Optimizer.py:236-244)
        const double den = D.kc.tau * D.kc.A * o.mu[0] + S1 * o.mu[1] + S2 * o.mu[2];
        for (int j = 0; j < m; ++j) {
            const int g = pth[j];
            vals[i * (size_t)m + j] = D.rN[j] * (D.kc.tau * o.mu[0] + D.ra[g] * o.mu[1] + D.rb[g] * o.mu[2]) / den;
        }
    }
}

}  // namespace theta

using namespace theta;

// ======================================================================= host: plan
int theta_n3_plan_init(theta_plan* pl) {
    const int m = pl->m;
    int K = 0;
    for (int i = 0; i < m; ++i) K = std::max(K, pl->upper[i]);
    if (K > 7) { set_error("n=3 supports copy numbers up to 7 (got %d)", K); return THETA_ERR_UNSUPPORTED; }
    if (m > 1024) { set_error("n=3 supports m <= 1024"); return THETA_ERR_UNSUPPORTED; }
    theta_n3_state* s = new theta_n3_state();
    pl->n3 = s;
    s->k = K;  // Enumerator.py:58: k = max(upper_bound)
    // Enumerator._create_graph (Enumerator.py:272-298): rows with the first component fastest
    for (int b = 0; b <= K; ++b)
        for (int a = 0; a <= K; ++a)
            if ((pl->tau - a) * (pl->tau - b) >= 0) { s->ra.push_back(a); s->rb.push_back(b); }
    s->NR = (int)s->ra.size();
    s->edges.assign(s->NR, 0);
    for (int v = 0; v < s->NR; ++v)
        for (int w = 0; w < s->NR; ++w)
            if (v == w || s->ra[w] > s->ra[v] || s->rb[w] > s->rb[v]) s->edges[v] |= 1ull << w;
    for (int g = 0; g < s->NR; ++g) {
        if (s->ra[g] <= s->rb[g]) s->symmask |= 1ull << g;
        if (s->ra[g] == s->rb[g]) s->eqmask |= 1ull << g;
    }
    s->allow.assign(m, 0);
    for (int i = 0; i < m; ++i)
        for (int g = 0; g < s->NR; ++g)
            if (s->ra[g] >= pl->lower[i] && s->ra[g] <= pl->upper[i] && s->rb[g] >= pl->lower[i] && s->rb[g] <= pl->upper[i])
                s->allow[i] |= 1ull << g;
    long double A = 0, Rt = 0, K0 = 0;
    int zero_rN = 0;
    for (int i = 0; i < m; ++i) { A += pl->rN[i]; Rt += pl->r[i]; if (pl->rN[i] == 0) zero_rN = 1; }
    long double Kr = 0;
    for (int i = 0; i < m; ++i)
        if (pl->r[i] > 0 && pl->rN[i] > 0) {
            K0 -= (long double)pl->r[i] * logl((long double)pl->rN[i] / A);
            Kr -= (long double)pl->r[i] * logl((long double)pl->rN[i]);
        }
    s->kc.Kr = (double)Kr;
    s->kc.A = (double)A; s->kc.Rtot = (double)Rt; s->kc.K0 = (double)K0; s->kc.tau = pl->tau;
    s->kc.m = m; s->kc.NR = s->NR; s->kc.any_zero_rN = zero_rN;
    std::vector<double> rd(m), rnd(m);
    for (int i = 0; i < m; ++i) { rd[i] = (double)pl->r[i]; rnd[i] = (double)pl->rN[i]; }
    THETA_CUDA(theta::dmalloc(&s->d_allow, sizeof(uint64_t) * m));
    THETA_CUDA(theta::dmalloc(&s->d_r, sizeof(double) * m));
    THETA_CUDA(theta::dmalloc(&s->d_rN, sizeof(double) * m));
    THETA_CUDA(theta::dmalloc(&s->d_ctr, sizeof(unsigned long long) * kCtrWords));
    THETA_CUDA(theta::dmalloc(&s->d_ctr_rows, sizeof(unsigned long long) * kCtrWords));
    THETA_CUDA(cudaMemcpy(s->d_allow, s->allow.data(), sizeof(uint64_t) * m, cudaMemcpyHostToDevice));
    THETA_CUDA(cudaMemcpy(s->d_r, rd.data(), sizeof(double) * m, cudaMemcpyHostToDevice));
    THETA_CUDA(cudaMemcpy(s->d_rN, rnd.data(), sizeof(double) * m, cudaMemcpyHostToDevice));
    {   // prefix sums of r / rN for the streaming NLL kernel (theta_tiles.cu), flat past m
        const int64_t pitch = theta_tile_pitch(m);
        pl->tile_pitch = pitch;
        std::vector<double> P(2 * (pitch + 1), 0.0);
        int64_t sr = 0, sn = 0;
        for (int64_t i = 0; i < pitch; ++i) {
            if (i < m) { sr += pl->r[i]; sn += pl->rN[i]; }
            P[i + 1] = (double)sr; P[pitch + 1 + i + 1] = (double)sn;
        }
        THETA_CUDA(theta::dmalloc(&pl->d_arena, sizeof(double) * P.size()));
        THETA_CUDA(cudaMemcpy(pl->d_arena, P.data(), sizeof(double) * P.size(), cudaMemcpyHostToDevice));
        pl->d_PR = reinterpret_cast<double*>(pl->d_arena);
        pl->d_PRN = pl->d_PR + (pitch + 1);
        THETA_CUDA(theta::logtab_device(&pl->d_logtab));
    }
    return THETA_OK;
}

int theta_n3_tile_info(const theta_plan* pl, uint8_t* ra, uint8_t* rb, int* n_rows, theta::N2Const* kc) {
    if (!pl || pl->n != 3 || !pl->n3) { set_error("plan is not n=3"); return THETA_ERR_INVALID; }
    const theta_n3_state* s = pl->n3;
    for (int g = 0; g < 64; ++g) { ra[g] = g < s->NR ? (uint8_t)s->ra[g] : 0; rb[g] = g < s->NR ? (uint8_t)s->rb[g] : 0; }
    *n_rows = s->NR;
    *kc = theta::N2Const{};
    kc->A = s->kc.A; kc->S0 = s->kc.tau * s->kc.A; kc->Rtot = s->kc.Rtot; kc->K0 = s->kc.K0; kc->tau = s->kc.tau;
    kc->max_normal = 1.0; kc->min_normal = 0.0; kc->any_zero_rN = s->kc.any_zero_rN; kc->m = s->kc.m;
    return THETA_OK;
}

void theta_n3_plan_free(theta_plan* pl) {
    theta_n3_state* s = pl->n3;
    if (!s) return;
    theta::dfree(s->d_allow); theta::dfree(s->d_r); theta::dfree(s->d_rN); theta::dfree(s->d_prefix); theta::dfree(s->d_off);
    theta::dfree(s->d_ctr); theta::dfree(s->d_ctr_rows); theta::dfree(s->d_ref_list); theta::dfree(s->d_q_meta); theta::dfree(s->d_q_lb); theta::dfree(s->d_q_res); theta::dfree(s->d_q_path);
    theta::dfree(s->d_list);
    delete s;
    pl->n3 = nullptr;
}

static void fill_dev(const theta_plan* pl, N3Dev* D) {
    const theta_n3_state* s = pl->n3;
    std::memset(D, 0, sizeof(*D));
    D->allow = s->d_allow; D->r = s->d_r; D->rN = s->d_rN;
    for (int g = 0; g < s->NR; ++g) { D->edges[g] = s->edges[g]; D->ra[g] = (uint8_t)s->ra[g]; D->rb[g] = (uint8_t)s->rb[g]; }
    D->symmask = s->symmask; D->eqmask = s->eqmask; D->kc = s->kc;
    D->logtab = pl->d_logtab;
}

static int check_n3(const theta_plan* pl) {
    if (!pl) { set_error("null plan"); return THETA_ERR_INVALID; }
    if (pl->n != 3 || !pl->n3) { set_error("plan is not n=3"); return THETA_ERR_INVALID; }
    return THETA_OK;
}

extern "C" int theta_plan_null_nll(const theta_plan* pl, double* nll) {
    if (!pl || !nll) { set_error("null argument"); return THETA_ERR_INVALID; }
    *nll = pl->n == 3 && pl->n3 ? pl->n3->kc.K0 : pl->kc.K0;
    return THETA_OK;
}

extern "C" int theta_n3_rows(const theta_plan* pl, int32_t* a, int32_t* b, int32_t* n_rows) {
    int rc = check_n3(pl);
    if (rc) return rc;
    const theta_n3_state* s = pl->n3;
    if (n_rows) *n_rows = s->NR;
    for (int g = 0; g < s->NR; ++g) { if (a) a[g] = s->ra[g]; if (b) b[g] = s->rb[g]; }
    return THETA_OK;
}

// ----------------------------------------------------------------------- items
namespace {
struct Expander {
    const theta_plan* pl;
    theta_n3_state* s;
    double cap;
    size_t max_items;
    bool overflow = false;
    std::vector<uint8_t> path;
    // where the items go (theta_n3_prepare: the plan's item list; pass 2: a temporary list of sub-items)
    std::vector<uint8_t>* o_prefix = nullptr;
    std::vector<uint32_t>* o_off = nullptr;
    std::vector<double>* o_weight = nullptr;

    void build_est() {
        const int m = pl->m, NR = s->NR;
        std::vector<std::vector<double>>& est = s->est;
        est.assign(m, std::vector<double>(NR * 2, 0.0));
        for (int g = 0; g < NR; ++g) est[m - 1][g * 2] = est[m - 1][g * 2 + 1] = 1.0;
        for (int i = m - 2; i >= 0; --i)
            for (int g = 0; g < NR; ++g)
                for (int sw = 0; sw < 2; ++sw) {
                    uint64_t c = s->edges[g] & s->allow[i + 1];
                    if (sw) c &= s->symmask;
                    double t = 0.0;
                    while (c) {
                        const int w = __builtin_ctzll(c);
                        c &= c - 1;
                        t += est[i + 1][w * 2 + ((sw && (s->eqmask >> w & 1)) ? 1 : 0)];
                    }
                    est[i][g * 2 + sw] = t;
                }
    }
    void emit(int depth, double w) {
        if (o_off->size() - 1 >= max_items) { overflow = true; return; }
        o_prefix->insert(o_prefix->end(), path.begin(), path.begin() + depth + 1);
        o_off->push_back((uint32_t)o_prefix->size());
        o_weight->push_back(w);
    }
    void rec(int depth, uint8_t mn, uint8_t mx, int sw) {
        if (overflow) return;
        const int g = path[depth];
        const double w = s->est[depth][g * 2 + sw];
        if (w <= 0.0) return;  // no completion even without the mu-bounds
        if (depth == pl->m - 1 || w <= cap) { emit(depth, w); return; }
        uint64_t c = s->edges[g] & s->allow[depth + 1];
        if (sw) c &= s->symmask;
        while (c) {
            const int w2 = __builtin_ctzll(c);
            c &= c - 1;
            uint8_t mn2 = mn, mx2 = mx;
            if (!theta::step_bounds(s->ra[g], s->rb[g], s->ra[w2], s->rb[w2], mn2, mx2)) continue;
            path[depth + 1] = (uint8_t)w2;
            rec(depth + 1, mn2, mx2, (sw && (s->eqmask >> w2 & 1)) ? 1 : 0);
        }
    }
    void run() {
        build_est();
        path.assign(pl->m, 0);
        o_prefix = &s->prefix; o_off = &s->off; o_weight = &s->weight;
        s->prefix.clear(); s->off.assign(1, 0); s->weight.clear();
        // Enumerator._generate_next_C_3 (Enumerator.py:172-187): first rows with a <= b, within bounds
        uint64_t c = s->allow[0] & s->symmask;
        while (c) {
            const int g = __builtin_ctzll(c);
            c &= c - 1;
            path[0] = (uint8_t)g;
            rec(0, 0, 255, (s->eqmask >> g & 1) ? 1 : 0);
        }
    }
};
}  // namespace

extern "C" int theta_n3_prepare(theta_plan* pl, int64_t leaf_cap, int64_t max_items, uint64_t* n_items) {
    int rc = check_n3(pl);
    if (rc) return rc;
    theta_n3_state* s = pl->n3;
    const bool auto_cap = leaf_cap < 1;
    if (auto_cap) leaf_cap = 256;
    if (max_items < 1) max_items = 1 << 22;
    // If the space is too large for max_items sub-trees of leaf_cap candidates, the list is cut after
    // max_items items: the search then covers a lexicographic prefix of the space (a rank budget) and
    // theta_n3_truncated() reports it.
    for (int round = 0; round < 2; ++round) {
        Expander ex{pl, s};
        ex.max_items = (size_t)max_items;
        ex.cap = (double)leaf_cap;
        ex.overflow = false;
        ex.run();
        s->truncated = ex.overflow;
        if (!auto_cap || round == 1 || ex.overflow) break;
        // small spaces: one thread per item needs tens of thousands of items to fill the machine
        const size_t ni0 = s->off.size() - 1;
        double tot = 0.0;
        for (double w : s->weight) tot += w;
        if (ni0 >= 32768 || tot <= (double)ni0) break;
        leaf_cap = std::max<int64_t>(1, (int64_t)(tot / 32768.0));
        if (leaf_cap >= 256) break;
    }
    const size_t ni = s->off.size() - 1;
    theta::dfree(s->d_prefix); theta::dfree(s->d_off);
    s->d_prefix = nullptr; s->d_off = nullptr;
    s->list_begin = s->list_end = 0; s->q_count = 0;
    if (ni) {
        THETA_CUDA(theta::dmalloc(&s->d_prefix, std::max<size_t>(1, s->prefix.size())));
        THETA_CUDA(theta::dmalloc(&s->d_off, sizeof(uint32_t) * (ni + 1)));
        THETA_CUDA(cudaMemcpy(s->d_prefix, s->prefix.data(), s->prefix.size(), cudaMemcpyHostToDevice));
        THETA_CUDA(cudaMemcpy(s->d_off, s->off.data(), sizeof(uint32_t) * (ni + 1), cudaMemcpyHostToDevice));
    }
    if (n_items) *n_items = ni;
    return THETA_OK;
}

extern "C" int theta_n3_item_count(const theta_plan* pl, uint64_t* n_items) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (n_items) *n_items = pl->n3->off.empty() ? 0 : pl->n3->off.size() - 1;
    return THETA_OK;
}

extern "C" int theta_n3_truncated(const theta_plan* pl, int32_t* truncated) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (truncated) *truncated = pl->n3->truncated ? 1 : 0;
    return THETA_OK;
}

extern "C" int theta_n3_item_weights(const theta_plan* pl, double* weights) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (weights) std::memcpy(weights, pl->n3->weight.data(), sizeof(double) * pl->n3->weight.size());
    return THETA_OK;
}

// ----------------------------------------------------------------------- launches
static int n3_env_screen() {
    const char* es = getenv("THETA_N3_SCREEN");  // 0: off, 1: fp32 logarithms, 2: fp64 table logarithms
    return es ? atoi(es) : kScreenDefault;
}
static int n3_env_verify() {
    const char* ev = getenv("THETA_VERIFY_APPROX");
    return (ev && ev[0] == '1') ? 1 : 0;
}

// survivor queue: meta | lower bounds | results | paths, grown on demand, kept until the plan is freed
static int ensure_queue(theta_n3_state* s, int m, size_t cap) {
    if (cap <= s->q_cap) return THETA_OK;
    theta::dfree(s->d_q_meta); theta::dfree(s->d_q_lb); theta::dfree(s->d_q_res); theta::dfree(s->d_q_path); theta::dfree(s->d_ref_list);
    s->d_q_meta = nullptr; s->d_q_lb = nullptr; s->d_q_res = nullptr; s->d_q_path = nullptr; s->d_ref_list = nullptr;
    s->q_cap = 0;
    THETA_CUDA(theta::dmalloc(&s->d_ref_list, sizeof(uint32_t) * cap));
    THETA_CUDA(theta::dmalloc(&s->d_q_meta, sizeof(uint32_t) * 2 * cap));
    THETA_CUDA(theta::dmalloc(&s->d_q_lb, sizeof(double) * cap));
    THETA_CUDA(theta::dmalloc(&s->d_q_res, sizeof(theta_cand) * cap));
    THETA_CUDA(theta::dmalloc(&s->d_q_path, cap * (size_t)m));
    s->q_cap = cap;
    return THETA_OK;
}

// CTA size of the walk kernel: the one that keeps the most threads resident given its shared-memory state
static int walk_config(const theta_plan* pl, int* block, int* ctas_per_sm, size_t* smem) {
    const theta_n3_state* s = pl->n3;
    static bool attr_set[64] = {false};
    if (pl->device >= 0 && pl->device < 64 && !attr_set[pl->device]) {
        THETA_CUDA(cudaFuncSetAttribute(n3_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        attr_set[pl->device] = true;
    }
    int best_b = 0, best_n = 0;
    size_t best_s = 0;
    const int cands[4] = {256, 128, 64, 32};
    for (int ci = 0; ci < 4; ++ci) {
        const int b = cands[ci];
        const size_t sm = walk_smem_fixed(pl->m) + (size_t)b * walk_smem_per_thread(pl->m, s->NR);
        if (sm > 227 * 1024 - 2048) continue;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, n3_walk_kernel, b, sm) != cudaSuccess) { cudaGetLastError(); continue; }
        if (nb * b > best_n * best_b) { best_b = b; best_n = nb; best_s = sm; }
    }
    if (!best_b) { set_error("n=3 walk kernel: m=%d does not fit shared memory", pl->m); return THETA_ERR_UNSUPPORTED; }
    *block = best_b; *ctas_per_sm = best_n; *smem = best_s;
    return THETA_OK;
}

struct WalkArgs {
    const uint32_t* d_list = nullptr;
    uint32_t item_begin = 0, n_work = 0;
    const uint8_t* d_prefix = nullptr;
    const uint32_t* d_off = nullptr;
    double margin = 0.0, fixed_thr = INFINITY;
    uint32_t scout_stride = 0;
    int enumerate_only = 0, screen = 0, verify = 0, count_leaves = 1;
};

static int run_walk(const theta_plan* pl, const WalkArgs& a, cudaStream_t st) {
    theta_n3_state* s = pl->n3;
    if (a.n_work == 0) return THETA_OK;
    int block, per_sm;
    size_t smem;
    int rc = walk_config(pl, &block, &per_sm, &smem);
    if (rc) return rc;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    WalkParams p{};
    fill_dev(pl, &p.D);
    p.prefix = a.d_prefix ? a.d_prefix : s->d_prefix;
    p.off = a.d_off ? a.d_off : s->d_off;
    p.list = a.d_list; p.item_begin = a.item_begin; p.n_work = a.n_work;
    p.ctr = s->d_ctr; p.margin = a.margin; p.fixed_thr = a.fixed_thr;
    p.q_meta = s->d_q_meta; p.q_lb = s->d_q_lb; p.q_path = s->d_q_path; p.q_cap = s->q_cap;
    p.enumerate_only = a.enumerate_only; p.screen = a.screen; p.verify = a.verify; p.block = block;
    p.scout_stride = a.scout_stride; p.count_leaves = a.count_leaves;
    // small spaces: fewer threads per CTA so that the items spread over the SMs
    while (block > 32 && (int64_t)a.n_work < (int64_t)sms * block) {
        block /= 2;
        smem = walk_smem_fixed(pl->m) + (size_t)block * walk_smem_per_thread(pl->m, s->NR);
        p.block = block;
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)a.n_work + block - 1) / block, (int64_t)sms * per_sm * 2));
    n3_walk_kernel<<<grid, block, smem, st>>>(p);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

// one warp per candidate, 9 m doubles of shared memory per warp
static int launch_ref(const theta_plan* pl, RefParams& rp, int max_ctas, cudaStream_t st) {
    const size_t per_warp = (size_t)9 * pl->m * sizeof(double);
    int warps = (int)std::min<size_t>(4, (200 * 1024) / per_warp);
    if (warps < 1) { set_error("n=3 reference-order solver: m=%d does not fit shared memory", pl->m); return THETA_ERR_UNSUPPORTED; }
    const size_t smem = per_warp * warps;
    static size_t attr_smem[64] = {0};
    if (pl->device >= 0 && pl->device < 64 && smem > attr_smem[pl->device] && smem > 48 * 1024) {
        THETA_CUDA(cudaFuncSetAttribute(n3_solve_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem[pl->device] = smem;
    }
    rp.warps = warps;
    n3_solve_ref_kernel<<<max_ctas, warps * 32, smem, st>>>(rp);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

static int run_solve(const theta_plan* pl, int mode, int enumerate_only, int skip_ref, cudaStream_t st) {
    theta_n3_state* s = pl->n3;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    SolveParams p{};
    fill_dev(pl, &p.D);
    p.ctr = s->d_ctr; p.q_meta = s->d_q_meta; p.q_lb = s->d_q_lb; p.q_path = s->d_q_path; p.q_res = s->d_q_res;
    p.q_cap = s->q_cap; p.mode = mode; p.enumerate_only = enumerate_only; p.ref_list = s->d_ref_list; p.skip_ref = skip_ref;
    n3_solve_queue_kernel<<<sms * 4, kBlock3, 0, st>>>(p);
    THETA_CUDA(cudaGetLastError());
    if (!enumerate_only && !skip_ref) {
        RefParams rp{};
        fill_dev(pl, &rp.D);
        rp.ctr = s->d_ctr; rp.ref_list = s->d_ref_list; rp.paths = s->d_q_path; rp.q_res = s->d_q_res; rp.q_lb = s->d_q_lb;
        int rc = launch_ref(pl, rp, sms * 2, st);
        if (rc) return rc;
    }
    n3_advance_kernel<<<1, 1, 0, st>>>(s->d_ctr, (unsigned long long)s->q_cap);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

// Pass 1 of the n=3 search over items [item_begin, item_end): the exact minimum NLL of the shard, and -- kept in
// the plan -- every candidate whose screening bound is within `window` (+ margin) of the best NLL known when it was
// screened, solved.  Two waves: a scout over ~8192 items spread over the shard (every 8th leaf gets the full Newton
// bound, the most promising leaf of each item is solved exactly: the first T), then everything against that T.
// Order does not matter (the fold needs only the candidates within the tie cluster of the global minimum,
// DESIGN.md "rank-free record set"), so T is one number shared by every thread, wave and GPU (t_seed).
extern "C" int theta_search_n3_min(theta_plan* pl, uint64_t item_begin, uint64_t item_end, double t_seed, double window,
                                   double* local_min, void* stream) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (!local_min) { set_error("null argument"); return THETA_ERR_INVALID; }
    theta_n3_state* s = pl->n3;
    if (s->off.empty()) { set_error("theta_n3_prepare has not been called"); return THETA_ERR_INVALID; }
    const uint64_t ni = s->off.size() - 1;
    if (item_end > ni) item_end = ni;
    if (item_begin > item_end) { set_error("bad item range"); return THETA_ERR_INVALID; }
    if (!(window > 0.0)) window = 4.0 * THETA_TIE_MARGIN;
    cudaStream_t st = (cudaStream_t)stream;
    s->p1_begin = item_begin; s->p1_end = item_end; s->q_count = 0; s->window = window;
    pl->stats = theta_search_stats{};
    pl->stats.min_nll = INFINITY;
    *local_min = INFINITY;
    const uint64_t n = item_end - item_begin;
    if (n == 0) return THETA_OK;
    double wtot = 0.0;
    for (uint64_t i = item_begin; i < item_end; ++i) wtot += s->weight[i];
    const int screen = n3_env_screen(), verify = n3_env_verify();
    const bool waves = screen && !verify && wtot > 65536.0;
    // wave lists
    uint32_t n0 = 0;
    if (waves) {
        const uint64_t stride = std::max<uint64_t>(1, n / 8192);
        if (s->list_begin != item_begin || s->list_end != item_end || !s->d_list) {
            std::vector<uint32_t> l0;
            for (uint64_t i = item_begin; i < item_end; i += stride) l0.push_back((uint32_t)i);
            theta::dfree(s->d_list); s->d_list = nullptr;
            THETA_CUDA(theta::dmalloc(&s->d_list, sizeof(uint32_t) * l0.size()));
            THETA_CUDA(cudaMemcpyAsync(s->d_list, l0.data(), sizeof(uint32_t) * l0.size(), cudaMemcpyHostToDevice, st));
            THETA_CUDA(cudaStreamSynchronize(st));
            s->list_begin = item_begin; s->list_end = item_end; s->list_n0 = (uint32_t)l0.size();
        }
        n0 = s->list_n0;
    }
    size_t cap = (verify || !screen) ? (size_t)wtot + 1024 : (size_t)1 << 16;
    if (!waves) cap = std::max(cap, (size_t)std::min(wtot, 65536.0) + 1024);
    unsigned long long ctr[kCtrWords];
    float ms = 0.f;
    for (int attempt = 0;; ++attempt) {
        rc = ensure_queue(s, pl->m, cap);
        if (rc) return rc;
        unsigned long long init[kCtrWords] = {0};
        double seed = (t_seed == t_seed && t_seed >= 0.0) ? t_seed : INFINITY;
        std::memcpy(&init[kCtrT], &seed, 8);
        THETA_CUDA(cudaMemcpyAsync(s->d_ctr, init, sizeof(init), cudaMemcpyHostToDevice, st));
        cudaEvent_t e0, e1;
        THETA_CUDA(cudaEventCreate(&e0)); THETA_CUDA(cudaEventCreate(&e1));
        THETA_CUDA(cudaEventRecord(e0, st));
        WalkArgs a;
        a.margin = window + THETA_TIE_MARGIN; a.screen = screen; a.verify = verify;
        const int mode = THETA_N3_MODE_AUTO;
        if (waves) {  // scout wave: its exact solves give the threshold T of the full wave
            a.d_list = s->d_list; a.n_work = n0; a.count_leaves = 0;
            a.scout_stride = (uint32_t)std::max(4.0, wtot / (double)n / 8.0);  // ~8 sampled leaves per scout item
            rc = run_walk(pl, a, st);
            if (!rc) rc = run_solve(pl, mode, 0, 1, st);
        }
        a.d_list = nullptr; a.item_begin = (uint32_t)item_begin; a.n_work = (uint32_t)n; a.scout_stride = 0; a.count_leaves = 1;
        if (!rc) rc = run_walk(pl, a, st);
        if (!rc) rc = run_solve(pl, mode, 0, 0, st);
        if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
        THETA_CUDA(cudaEventRecord(e1, st));
        THETA_CUDA(cudaMemcpyAsync(ctr, s->d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, st));
        THETA_CUDA(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (ctr[kCtrQueue] <= cap) break;
        if (attempt >= 6) { set_error("n=3 survivor queue overflow (%llu survivors)", ctr[kCtrQueue]); return THETA_ERR_OVERFLOW; }
        cap = std::max<size_t>(2 * cap, (size_t)ctr[kCtrQueue] + 1024);  // a large tie cluster: retry with room for it
    }
    double g;
    std::memcpy(&g, &ctr[kCtrT], 8);
    s->q_count = ctr[kCtrQueue];
    s->nan_nll = ctr[kCtrNan];
    pl->stats.n_candidates = ctr[kCtrLeaves];
    pl->stats.n_solved = ctr[kCtrSolved];
    pl->stats.n_exact = ctr[kCtrQueue];
    pl->stats.n_bound_violations = ctr[kCtrViol];
    pl->stats.n_stripes = n;
    pl->stats.n_rescanned = ctr[kCtrRef];  // survivors that took the reference-order solver
    pl->stats.newton_iters = ctr[kCtrSolved] ? (double)ctr[kCtrNfev] / (double)ctr[kCtrSolved] : 0.0;
    pl->stats.min_nll = g;
    pl->stats.ms_pass1 = ms;
    s->it_hist[0] = ctr[kCtrIt0]; s->it_hist[1] = ctr[kCtrIt1]; s->it_hist[2] = ctr[kCtrIt2];
    if (getenv("THETA_DEBUG"))
        fprintf(stderr, "[theta n3] leaves %llu survivors %llu (ref-order %llu) screen iterations 1/2/3: %llu/%llu/%llu  %.2f ms\n",
                ctr[kCtrLeaves], ctr[kCtrQueue], ctr[kCtrRef], ctr[kCtrIt0], ctr[kCtrIt1], ctr[kCtrIt2], ms);
    *local_min = g;
    return THETA_OK;
}

static int solve_rows_dev(theta_plan* pl, const uint8_t* d_p, int64_t count, double* d_mu, double* d_mh, double* d_nll,
                          int32_t* d_st, double* d_vals, int mode, cudaStream_t st);

// Every solved candidate of the shard with NLL <= global_min + window, or a NaN likelihood (Misc.isClose(nan, x) is
// True: Q6), in enumeration order, re-solved by the reference-order solver (mode 1) so that mu and NLL carry the
// reference's own roundings.  `window` must not exceed the one given to theta_search_n3_min.
extern "C" int theta_search_n3_window(theta_plan* pl, double global_min, double window, theta_cand* out_host, uint64_t cap,
                                      uint64_t* n_out, uint8_t* paths_host, void* stream) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (!n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_out = 0;
    theta_n3_state* s = pl->n3;
    cudaStream_t st = (cudaStream_t)stream;
    if (!(window > 0.0)) window = s->window;
    if (window > s->window) { set_error("window %g exceeds the searched one (%g)", window, s->window); return THETA_ERR_INVALID; }
    const size_t m = (size_t)pl->m;
    const size_t cnt = (size_t)std::min<uint64_t>(s->q_count, s->q_cap);
    if (cnt == 0) return THETA_OK;
    std::vector<theta_cand> all(cnt);
    THETA_CUDA(cudaMemcpyAsync(all.data(), s->d_q_res, sizeof(theta_cand) * cnt, cudaMemcpyDeviceToHost, st));
    THETA_CUDA(cudaStreamSynchronize(st));
    const double thr = global_min + window;
    std::vector<theta_cand> sel;
    for (const theta_cand& c : all)
        if (c.status >= 0 && c.status < THETA_ST_NONE_SIGN && (c.nll != c.nll || c.nll <= thr)) sel.push_back(c);
    std::sort(sel.begin(), sel.end(), [](const theta_cand& a, const theta_cand& b) { return a.rank < b.rank; });
    sel.erase(std::unique(sel.begin(), sel.end(), [](const theta_cand& a, const theta_cand& b) { return a.rank == b.rank; }), sel.end());
    *n_out = sel.size();
    if (sel.empty()) return THETA_OK;
    if (sel.size() > cap) { set_error("window holds %zu records, buffer holds %llu", sel.size(), (unsigned long long)cap); return THETA_ERR_OVERFLOW; }
    // gather the paths on the device, re-solve in reference order
    const int64_t k = (int64_t)sel.size();
    uint8_t* d_p = nullptr; double *d_mu = nullptr, *d_mh = nullptr, *d_nll = nullptr; int32_t* d_st = nullptr;
    cudaError_t e = theta::dmalloc(&d_p, (size_t)k * m);
    if (e == cudaSuccess) e = theta::dmalloc(&d_mu, sizeof(double) * 3 * k);
    if (e == cudaSuccess) e = theta::dmalloc(&d_mh, sizeof(double) * 3 * k);
    if (e == cudaSuccess) e = theta::dmalloc(&d_nll, sizeof(double) * k);
    if (e == cudaSuccess) e = theta::dmalloc(&d_st, sizeof(int32_t) * k);
    for (int64_t i = 0; e == cudaSuccess && i < k; ++i)
        e = cudaMemcpyAsync(d_p + (size_t)i * m, s->d_q_path + (size_t)sel[i].flags * m, m, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "n3 window setup", __FILE__, __LINE__);
    if (!rc) rc = solve_rows_dev(pl, d_p, k, d_mu, d_mh, d_nll, d_st, nullptr, THETA_N3_MODE_REF, st);
    std::vector<double> mu(3 * k), nll(k);
    std::vector<int32_t> stt(k);
    std::vector<uint8_t> paths((size_t)k * m);
    if (!rc) {
        e = cudaMemcpyAsync(mu.data(), d_mu, sizeof(double) * 3 * k, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nll.data(), d_nll, sizeof(double) * k, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(stt.data(), d_st, sizeof(int32_t) * k, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(paths.data(), d_p, (size_t)k * m, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "n3 window D2H", __FILE__, __LINE__);
    }
    theta::dfree(d_p); theta::dfree(d_mu); theta::dfree(d_mh); theta::dfree(d_nll); theta::dfree(d_st);
    if (rc) return rc;
    for (int64_t i = 0; i < k; ++i) {
        sel[i].nll = nll[i]; sel[i].mu[0] = mu[3 * i]; sel[i].mu[1] = mu[3 * i + 1]; sel[i].mu[2] = mu[3 * i + 2];
        sel[i].status = stt[i]; sel[i].flags = (int32_t)i;
    }
    if (out_host) std::memcpy(out_host, sel.data(), sizeof(theta_cand) * sel.size());
    if (paths_host) std::memcpy(paths_host, paths.data(), paths.size());
    return THETA_OK;
}

extern "C" int theta_enumerate_n3(theta_plan* pl, uint64_t item_begin, uint64_t item_end, uint8_t* paths_host, uint64_t cap,
                                  uint64_t* n_out) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (!n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    theta_n3_state* s = pl->n3;
    if (s->off.empty()) { set_error("theta_n3_prepare has not been called"); return THETA_ERR_INVALID; }
    const uint64_t ni = s->off.size() - 1;
    if (item_end > ni) item_end = ni;
    *n_out = 0;
    if (item_begin >= item_end) return THETA_OK;
    double wtot = 0.0;
    for (uint64_t i = item_begin; i < item_end; ++i) wtot += s->weight[i];
    const size_t m = (size_t)pl->m;
    size_t qcap = (size_t)wtot + 1024;
    rc = ensure_queue(s, pl->m, qcap);
    if (rc) return rc;
    s->q_count = 0; s->list_begin = s->list_end = 0;
    unsigned long long init[kCtrWords] = {0};
    THETA_CUDA(cudaMemcpy(s->d_ctr, init, sizeof(init), cudaMemcpyHostToDevice));
    WalkArgs a;
    a.item_begin = (uint32_t)item_begin; a.n_work = (uint32_t)(item_end - item_begin); a.enumerate_only = 1;
    rc = run_walk(pl, a, nullptr);
    if (rc) return rc;
    unsigned long long ctr[kCtrWords];
    THETA_CUDA(cudaMemcpy(ctr, s->d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost));
    const size_t cnt = (size_t)ctr[kCtrQueue];
    if (cnt > s->q_cap) { set_error("enumeration produced more candidates than the item weights allow"); return THETA_ERR_OVERFLOW; }
    *n_out = cnt;
    if (cnt > cap) { set_error("enumeration produced %zu candidates, buffer holds %llu", cnt, (unsigned long long)cap); return THETA_ERR_OVERFLOW; }
    if (cnt == 0 || !paths_host) return THETA_OK;
    std::vector<uint32_t> meta(2 * cnt);
    std::vector<uint8_t> raw(cnt * m);
    THETA_CUDA(cudaMemcpy(meta.data(), s->d_q_meta, sizeof(uint32_t) * 2 * cnt, cudaMemcpyDeviceToHost));
    THETA_CUDA(cudaMemcpy(raw.data(), s->d_q_path, cnt * m, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> idx(cnt);
    for (size_t i = 0; i < cnt; ++i) idx[i] = (uint32_t)i;
    std::sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) {
        return meta[2 * x] != meta[2 * y] ? meta[2 * x] < meta[2 * y] : meta[2 * x + 1] < meta[2 * y + 1];
    });
    for (size_t i = 0; i < cnt; ++i) std::memcpy(paths_host + i * m, raw.data() + (size_t)idx[i] * m, m);
    return THETA_OK;
}

static int solve_rows_dev(theta_plan* pl, const uint8_t* d_p, int64_t count, double* d_mu, double* d_mh, double* d_nll,
                          int32_t* d_st, double* d_vals, int mode, cudaStream_t st) {
    theta_n3_state* s = pl->n3;
    N3Dev D;
    fill_dev(pl, &D);
    uint32_t* d_ref = nullptr;
    THETA_CUDA(theta::dmalloc(&d_ref, sizeof(uint32_t) * (size_t)count));
    cudaError_t e = cudaMemsetAsync(s->d_ctr_rows, 0, sizeof(unsigned long long) * kCtrWords, st);
    int rc = THETA_OK;
    if (e != cudaSuccess) rc = cuda_fail(e, "solve_rows counters", __FILE__, __LINE__);
    if (!rc) {
        const int grid = (int)((count + kBlock3 - 1) / kBlock3);
        n3_solve_rows_kernel<<<grid, kBlock3, 0, st>>>(D, d_p, count, d_mu, d_mh, d_nll, d_st, d_vals, mode, s->d_ctr_rows, d_ref);
        e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(e, "n3_solve_rows_kernel", __FILE__, __LINE__);
    }
    if (!rc && mode != THETA_N3_MODE_GROUPED) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
        RefParams rp{};
        rp.D = D; rp.ctr = s->d_ctr_rows; rp.ref_list = d_ref; rp.paths = d_p;
        rp.mu = d_mu; rp.muhat = d_mh; rp.nll = d_nll; rp.status = d_st; rp.vals = d_vals;
        rc = launch_ref(pl, rp, (int)std::min<int64_t>(sms * 2, count), st);
    }
    if (!rc) {
        e = cudaStreamSynchronize(st);  // d_ref is released below
        if (e != cudaSuccess) rc = cuda_fail(e, "solve_rows", __FILE__, __LINE__);
    }
    theta::dfree(d_ref);
    return rc;
}

extern "C" int theta_solve_rows_n3(theta_plan* pl, const uint8_t* paths_host, int64_t count, double* mu_host,
                                   double* muhat_host, double* nll_host, int32_t* status_host, double* vals_host, int32_t mode) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (!paths_host || !mu_host || !nll_host || !status_host) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (count <= 0) return THETA_OK;
    const int m = pl->m;
    uint8_t* d_p = nullptr; double *d_mu = nullptr, *d_mh = nullptr, *d_nll = nullptr, *d_vals = nullptr; int32_t* d_st = nullptr;
    cudaError_t e = theta::dmalloc(&d_p, (size_t)count * m);
    if (e == cudaSuccess) e = theta::dmalloc(&d_mu, sizeof(double) * 3 * count);
    if (e == cudaSuccess) e = theta::dmalloc(&d_mh, sizeof(double) * 3 * count);
    if (e == cudaSuccess) e = theta::dmalloc(&d_nll, sizeof(double) * count);
    if (e == cudaSuccess) e = theta::dmalloc(&d_st, sizeof(int32_t) * count);
    if (e == cudaSuccess && vals_host) e = theta::dmalloc(&d_vals, sizeof(double) * (size_t)count * m);
    if (e == cudaSuccess) e = cudaMemcpy(d_p, paths_host, (size_t)count * m, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && vals_host) e = cudaMemset(d_vals, 0, sizeof(double) * (size_t)count * m);
    if (e != cudaSuccess) rc = cuda_fail(e, "solve_rows_n3 setup", __FILE__, __LINE__);
    if (!rc) rc = solve_rows_dev(pl, d_p, count, d_mu, d_mh, d_nll, d_st, d_vals, (int)mode, nullptr);
    if (rc) {
        theta::dfree(d_p); theta::dfree(d_mu); theta::dfree(d_mh); theta::dfree(d_nll); theta::dfree(d_st); theta::dfree(d_vals);
        return rc;
    }
    if (e == cudaSuccess) e = cudaMemcpy(mu_host, d_mu, sizeof(double) * 3 * count, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && muhat_host) e = cudaMemcpy(muhat_host, d_mh, sizeof(double) * 3 * count, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(nll_host, d_nll, sizeof(double) * count, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(status_host, d_st, sizeof(int32_t) * count, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && vals_host) e = cudaMemcpy(vals_host, d_vals, sizeof(double) * (size_t)count * m, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "solve_rows_n3", __FILE__, __LINE__);
    theta::dfree(d_p); theta::dfree(d_mu); theta::dfree(d_mh); theta::dfree(d_nll); theta::dfree(d_st); theta::dfree(d_vals);
    return rc;
}
