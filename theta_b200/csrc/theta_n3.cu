// theta_n3.cu -- n = 3: on-chip DFS enumeration + KKT solve + the two search passes.
//
// Replaces, for n=3: Enumerator._create_graph / _generate_next_C_3 / _generate_next_C_3_recurse /
// _get_mu_bounds / _in_bounds (python/Enumerator.py:172-298), Optimizer._solve_n3plus
// (python/Optimizer.py:128-165) and the driver loop (python/RunTHetA.py:173-220).
//
// Work decomposition: the reference's DFS visits row sequences in lexicographic order of row index.
// The host cuts the DFS forest into *items* -- sub-trees given by a prefix of row indices -- using the
// count DP of TimeEstimate.count_number_matrices_3 (python/TimeEstimate.py:113-142) per sub-tree, and
// lists them in lexicographic order, so item order == enumeration order.  Every GPU thread takes items
// from a ticket counter and walks its sub-tree depth first with the reference's pruning rules,
// keeping the grouped statistics (R_g, S1, S2) incrementally (exact integer sums in fp64).
#include <algorithm>
#include <cmath>
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
    double* d_item_min = nullptr;
    unsigned long long* d_ctr = nullptr;
    std::vector<double> h_item_min;
    std::vector<std::vector<double>> est;  // est[i][g*2+sw]: completions of a prefix ending in row g at interval i (no mu-bounds)
    uint64_t p1_begin = 0, p1_end = 0;
    unsigned long long nan_nll = 0;
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

template <int MAXM>
struct Dfs {
    uint8_t path[MAXM], mn[MAXM], mx[MAXM], sw[MAXM];
    double R[kMaxRows3];
    uint16_t cnt[kMaxRows3];
    double S1, S2;
    int depth;  // last filled position
    int base;   // depth of the item prefix's last position

    __device__ void reset(int NR) {
        for (int g = 0; g < NR; ++g) { R[g] = 0.0; cnt[g] = 0; }
        S1 = S2 = 0.0;
        depth = -1;
    }
    __device__ void push(const N3Dev& D, int pos, int g, uint8_t mnc, uint8_t mxc, uint8_t s) {
        path[pos] = (uint8_t)g; mn[pos] = mnc; mx[pos] = mxc; sw[pos] = s;
        R[g] += D.r[pos];
        cnt[g] += 1;
        S1 += D.rN[pos] * (double)D.ra[g];
        S2 += D.rN[pos] * (double)D.rb[g];
        depth = pos;
    }
    __device__ void pop(const N3Dev& D) {
        const int pos = depth, g = path[pos];
        R[g] -= D.r[pos];
        cnt[g] -= 1;
        S1 -= D.rN[pos] * (double)D.ra[g];
        S2 -= D.rN[pos] * (double)D.rb[g];
        depth = pos - 1;
    }
    __device__ uint64_t children(const N3Dev& D) const {
        uint64_t c = D.edges[path[depth]] & D.allow[depth + 1];
        if (sw[depth]) c &= D.symmask;
        return c;
    }
    // replay an item prefix (already validated by the host)
    __device__ void load_prefix(const N3Dev& D, const uint8_t* pre, int len) {
        reset(D.kc.NR);
        uint8_t mnc = 0, mxc = 255, s = 0;
        for (int i = 0; i < len; ++i) {
            const int g = pre[i];
            if (i == 0) s = (D.ra[g] == D.rb[g]) ? 1 : 0;
            else {
                step_bounds(D.ra[path[i - 1]], D.rb[path[i - 1]], D.ra[g], D.rb[g], mnc, mxc);
                s = (s && D.ra[g] == D.rb[g]) ? 1 : 0;
            }
            push(D, i, g, mnc, mxc, s);
        }
        base = len - 1;
    }
    // advance to the next leaf of the sub-tree; `fresh` = nothing visited yet in this item
    __device__ bool next_leaf(const N3Dev& D, bool fresh) {
        const int last = D.kc.m - 1;
        uint64_t cand;
        if (fresh) {
            if (depth == last) return true;  // the prefix itself is a complete candidate
            cand = children(D);
        } else {
            if (depth == base) return false;  // item was a single leaf
            const int g = path[depth];
            pop(D);
            cand = children(D) & ~((2ull << g) - 1ull);
        }
        for (;;) {
            bool pushed = false;
            while (cand) {
                const int c = __ffsll((long long)cand) - 1;
                cand &= cand - 1;
                uint8_t mnc = mn[depth], mxc = mx[depth];
                const int pg = path[depth];
                if (step_bounds(D.ra[pg], D.rb[pg], D.ra[c], D.rb[c], mnc, mxc)) {
                    const uint8_t s = (sw[depth] && D.ra[c] == D.rb[c]) ? 1 : 0;
                    push(D, depth + 1, c, mnc, mxc, s);
                    pushed = true;
                    break;
                }
            }
            if (pushed) {
                if (depth == last) return true;
                cand = children(D);
            } else {
                if (depth == base) return false;
                const int g = path[depth];
                pop(D);
                cand = children(D) & ~((2ull << g) - 1ull);
            }
        }
    }
    __device__ void groups(const N3Dev& D, Groups3& G) const {
        int n = 0;
        for (int g = 0; g < D.kc.NR; ++g) {
            if (cnt[g]) {
                G.R[n] = R[g];
                G.a[n] = (double)D.ra[g];
                G.b[n] = (double)D.rb[g];
                ++n;
            }
        }
        G.G = n;
    }
};

struct Search3Params {
    N3Dev D;
    const uint8_t* prefix;
    const uint32_t* off;
    uint32_t item_begin;
    uint32_t n_items;            // pass 1: number of items from item_begin; pass 2: listed items
    double* item_min;            // [n_items] (pass 1, indexed from item_begin)
    unsigned long long* ctr;     // [0] ticket [1] solved [2] nfev [3] emitted [4] nan [5] leaves
    const uint32_t* list;        // pass 2 (NULL: item = ticket)
    const double* prior;         // pass 2
    const uint64_t* rank_base;   // pass 2 over sub-items: (parent item << 32) | (sub-item << 12); NULL: item << 32
    double near_thr;
    theta_cand* emit;
    uint8_t* emit_paths;
    unsigned long long emit_cap;
    int enumerate_only;          // emit every leaf's path, no solve
    int screen;                  // pass 1: skip the exact solve of leaves that n3_screen proves irrelevant
    int verify;                  // THETA_VERIFY_APPROX=1: solve every leaf and count leaves whose bound exceeds the exact NLL
};

// The per-candidate solver (hybrj restatement + fallback) is ~200 KB of machine code, far more than the
// instruction caches hold, so warps that run it at unrelated phases each stream it from L2 ("no instruction"
// was the top stall).  The leaf loop is therefore block-synchronous: every thread advances its own DFS to its
// next leaf (cheap, divergent), the block meets at a barrier, and all warps enter the solver together and
// move through it roughly in step, sharing the fetched instructions.
template <int MAXM, bool EMIT, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) n3_search_kernel(const Search3Params p) {
    __shared__ N3Dev D;
    __shared__ double2 s_LT[THETA_LOGTAB_SIZE];
    if (threadIdx.x == 0) D = p.D;
    if (p.screen == 2) for (int i = threadIdx.x; i < THETA_LOGTAB_SIZE; i += BLOCK) s_LT[i] = p.D.logtab[i];
    __syncthreads();
    const double2* LT = p.screen == 2 ? s_LT : nullptr;
    Dfs<MAXM> dfs;
    Groups3 G;
    unsigned long long solved = 0, nfev = 0, nan_nll = 0, leaves = 0, n_exact = 0, n_viol = 0;
    const int m = D.kc.m;
    // pass 1 screens every leaf with a rigorous lower bound (n3_screen) against T = the best exact NLL this
    // thread has seen so far -- all of it at lower ranks: a thread takes items in increasing order and walks
    // each in enumeration order -- so a skipped leaf is neither a strict prefix minimum nor in the tie window
    Screen3 ws{1.0 / 3.0, 1.0 / 3.0};
    double T = INFINITY;
    const double screen_margin = 3.0 * THETA_TIE_MARGIN;
    bool has_item = false, exhausted = false, fresh = true;
    uint32_t ticket = 0, item = 0, seq = 0;
    double prior = 0.0, tmin = INFINITY;
    for (;;) {
        bool have_leaf = false;
        while (!have_leaf && !exhausted) {
            if (!has_item) {
                ticket = (uint32_t)atomicAdd(&p.ctr[0], 1ULL);
                if (ticket >= p.n_items) { exhausted = true; break; }
                item = EMIT ? (p.list ? p.list[ticket] : ticket) : p.item_begin + ticket;
                prior = EMIT ? p.prior[ticket] : 0.0;
                dfs.load_prefix(D, p.prefix + p.off[item], (int)(p.off[item + 1] - p.off[item]));
                tmin = INFINITY;
                seq = 0;
                fresh = true;
                has_item = true;
                if (!EMIT && p.screen) {
                    // Tighten T with the minima of items that other threads have already finished: every item with
                    // a smaller ticket is at lower ranks, so its minimum (the exact NLL of one of its leaves, written
                    // once when the item completes; NaN until then) is a valid T.  Look well behind the tickets
                    // that are still in flight, 16 samples spread over everything older.
                    const long long lag = 2LL * (long long)gridDim.x * BLOCK;
                    const long long j0 = (long long)ticket - lag;
                    if (j0 >= 0) {
                        const long long stride = j0 / 16 + 1;
                        for (long long j = j0; j >= 0; j -= stride) {
                            const double v = __ldcg(p.item_min + j);
                            if (v == v) T = fmin(T, v);
                        }
                    }
                }
            }
            if (dfs.next_leaf(D, fresh)) {
                fresh = false;
                have_leaf = true;
            } else {
                if (!EMIT) p.item_min[ticket] = tmin;
                has_item = false;
            }
        }
        if (!__syncthreads_or(have_leaf ? 1 : 0)) break;
        if (have_leaf) {
            ++leaves;
            bool emit = false;
            Solve3Out o;
            o.status = THETA_ST_NONE_NAN;
            if (!p.enumerate_only) {
                dfs.groups(D, G);
                // (deferring the leaves that survive the screen so that the solver runs with full warps was
                // measured slower: surviving leaves cluster in time, and the longer divergent screening loops
                // between solver rounds cost more than they save)
                bool skip = false;
                double lb = -INFINITY;
                if (p.screen) {
                    // pass 2 emits a leaf only if NLL <= near_thr or NLL < min(prior, item minimum so far): a leaf whose
                    // lower bound exceeds both thresholds is neither emitted nor changes the item minimum
                    const double thr = EMIT ? fmax(p.near_thr, fmin(prior, tmin)) : T + screen_margin;
                    skip = n3_screen(G, dfs.S1, dfs.S2, D.kc, ws, thr, lb, LT) && !p.verify;
                }
                if (!skip) {
                    const ref3::Ctx rc{m, D.r, D.rN, dfs.path, 1, D.ra, D.rb, D.kc.tau, D.kc.tau * D.kc.A, dfs.S1, dfs.S2};
                    n3_solve_auto(G, dfs.S1, dfs.S2, D.kc, rc, EMIT ? THETA_N3_MODE_REF : THETA_N3_MODE_AUTO, o);
                    ++n_exact;
                    if (o.status < THETA_ST_NONE_SIGN) {
                        ++solved;
                        nfev += o.nfev;
                        if (o.nll != o.nll) {
                            ++nan_nll;
                            emit = true;  // Misc.isClose(nan, x) is True (Q6)
                        } else {
                            if (EMIT) emit = (o.nll <= p.near_thr) || (o.nll < prior && o.nll < tmin);
                            tmin = fmin(tmin, o.nll);
                            T = fmin(T, o.nll);
                            if (p.verify && lb > o.nll) ++n_viol;
                        }
                    }
                }
            } else {
                emit = true;
            }
            if (EMIT && emit) {
                const unsigned long long slot = atomicAdd(&p.ctr[3], 1ULL);
                if (slot < p.emit_cap) {
                    theta_cand c;
                    c.rank = (p.rank_base ? p.rank_base[ticket] : ((uint64_t)item << 32)) + seq;
                    c.nll = o.nll;
                    c.mu[0] = o.mu[0]; c.mu[1] = o.mu[1]; c.mu[2] = o.mu[2];
                    c.status = o.status;
                    c.flags = (int32_t)slot;
                    p.emit[slot] = c;
                    uint8_t* dst = p.emit_paths + slot * (size_t)m;
                    for (int i = 0; i < m; ++i) dst[i] = dfs.path[i];
                }
            }
            ++seq;
        }
    }
    if (solved) atomicAdd(&p.ctr[1], solved);
    if (nfev) atomicAdd(&p.ctr[2], nfev);
    if (nan_nll) atomicAdd(&p.ctr[4], nan_nll);
    if (leaves) atomicAdd(&p.ctr[5], leaves);
    if (n_exact) atomicAdd(&p.ctr[6], n_exact);
    if (n_viol) atomicAdd(&p.ctr[7], n_viol);
}

// Optimizer.solve on explicit row-index strings
__global__ void __launch_bounds__(kBlock3) n3_solve_rows_kernel(const N3Dev D0, const uint8_t* paths, int64_t count,
                                                                double* mu, double* muhat, double* nll, int32_t* status,
                                                                double* vals, int mode) {
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
        const ref3::Ctx rc{m, D.r, D.rN, pth, 1, D.ra, D.rb, D.kc.tau, D.kc.tau * D.kc.A, S1, S2};
        n3_solve_auto(G, S1, S2, D.kc, rc, mode, o);
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
    THETA_CUDA(theta::dmalloc(&s->d_ctr, sizeof(unsigned long long) * 8));
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
    kc->max_normal = 1.0; kc->any_zero_rN = s->kc.any_zero_rN; kc->m = s->kc.m;
    return THETA_OK;
}

void theta_n3_plan_free(theta_plan* pl) {
    theta_n3_state* s = pl->n3;
    if (!s) return;
    theta::dfree(s->d_allow); theta::dfree(s->d_r); theta::dfree(s->d_rN); theta::dfree(s->d_prefix); theta::dfree(s->d_off);
    theta::dfree(s->d_item_min); theta::dfree(s->d_ctr);
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
    // the sub-trees of one item (given by its prefix), each with at most `cap` leaves, in enumeration order
    void run_from_prefix(const uint8_t* pre, int len) {
        path.assign(pl->m, 0);
        uint8_t mn = 0, mx = 255;
        int sw = 0;
        for (int i = 0; i < len; ++i) {
            path[i] = pre[i];
            if (i == 0) sw = (s->eqmask >> pre[0] & 1) ? 1 : 0;
            else {
                theta::step_bounds(s->ra[pre[i - 1]], s->rb[pre[i - 1]], s->ra[pre[i]], s->rb[pre[i]], mn, mx);
                sw = (sw && (s->eqmask >> pre[i] & 1)) ? 1 : 0;
            }
        }
        rec(len - 1, mn, mx, sw);
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
    if (leaf_cap < 1) leaf_cap = 256;
    if (max_items < 1) max_items = 1 << 22;
    Expander ex{pl, s};
    ex.max_items = (size_t)max_items;
    ex.cap = (double)leaf_cap;
    // If the space is too large for max_items sub-trees of leaf_cap candidates, the list is cut after
    // max_items items: the search then covers a lexicographic prefix of the space (a rank budget).
    ex.overflow = false;
    ex.run();
    s->truncated = ex.overflow;
    const size_t ni = s->off.size() - 1;
    theta::dfree(s->d_prefix); theta::dfree(s->d_off); theta::dfree(s->d_item_min);
    s->d_prefix = nullptr; s->d_off = nullptr; s->d_item_min = nullptr;
    if (ni) {
        THETA_CUDA(theta::dmalloc(&s->d_prefix, std::max<size_t>(1, s->prefix.size())));
        THETA_CUDA(theta::dmalloc(&s->d_off, sizeof(uint32_t) * (ni + 1)));
        THETA_CUDA(theta::dmalloc(&s->d_item_min, sizeof(double) * ni));
        THETA_CUDA(cudaMemcpy(s->d_prefix, s->prefix.data(), s->prefix.size(), cudaMemcpyHostToDevice));
        THETA_CUDA(cudaMemcpy(s->d_off, s->off.data(), sizeof(uint32_t) * (ni + 1), cudaMemcpyHostToDevice));
    }
    if (n_items) *n_items = ni;
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
// THETA_N3_BLOCK (tuning knob): threads per CTA of the n=3 search kernel (64, 256 or 512; default: by item count)
static int g_n3_block = -1;

template <bool EMIT, int BLOCK, int MINB>
static int launch_n3_b(const theta_plan* pl, const Search3Params& sp, cudaStream_t st) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    const int m = pl->m;
    const int64_t threads_wanted = (int64_t)sp.n_items;
    int grid = (int)std::min<int64_t>((threads_wanted + BLOCK - 1) / BLOCK, (int64_t)sms * MINB);
    if (grid < 1) grid = 1;
    if (m <= 64) n3_search_kernel<64, EMIT, BLOCK, MINB><<<grid, BLOCK, 0, st>>>(sp);
    else if (m <= 256) n3_search_kernel<256, EMIT, BLOCK, MINB><<<grid, BLOCK, 0, st>>>(sp);
    else n3_search_kernel<1024, EMIT, BLOCK, MINB><<<grid, BLOCK, 0, st>>>(sp);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

template <bool EMIT>
static int launch_n3(const theta_plan* pl, const Search3Params& sp, cudaStream_t st) {
    if (g_n3_block < 0) {
        const char* e = getenv("THETA_N3_BLOCK");
        g_n3_block = e ? atoi(e) : 0;
    }
    int block = g_n3_block;
    if (block == 0) {  // the largest CTA that still gives every SM at least half a CTA of items
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
        const int64_t per_sm = (int64_t)sp.n_items / sms;
        block = per_sm >= 256 ? 512 : per_sm >= 128 ? 256 : 64;
    }
    switch (block) {
        case 64: return launch_n3_b<EMIT, 64, 8>(pl, sp, st);
        case 512: return launch_n3_b<EMIT, 512, 1>(pl, sp, st);
        default: return launch_n3_b<EMIT, 256, 2>(pl, sp, st);
    }
}

extern "C" int theta_search_n3_pass1(theta_plan* pl, uint64_t item_begin, uint64_t item_end, double* local_min, void* stream) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (!local_min) { set_error("null argument"); return THETA_ERR_INVALID; }
    theta_n3_state* s = pl->n3;
    const uint64_t ni = s->off.size() ? s->off.size() - 1 : 0;
    if (s->off.empty()) { set_error("theta_n3_prepare has not been called"); return THETA_ERR_INVALID; }
    if (item_end > ni) item_end = ni;
    if (item_begin > item_end) { set_error("bad item range"); return THETA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    s->p1_begin = item_begin; s->p1_end = item_end;
    pl->stats = theta_search_stats{};
    pl->stats.min_nll = INFINITY;
    *local_min = INFINITY;
    const uint64_t n = item_end - item_begin;
    s->h_item_min.assign(n, INFINITY);
    if (n == 0) return THETA_OK;
    THETA_CUDA(cudaMemsetAsync(s->d_ctr, 0, sizeof(unsigned long long) * 8, st));
    THETA_CUDA(cudaMemsetAsync(s->d_item_min, 0xff, sizeof(double) * n, st));  // NaN = "item not finished yet"
    Search3Params sp{};
    fill_dev(pl, &sp.D);
    sp.prefix = s->d_prefix; sp.off = s->d_off; sp.item_begin = (uint32_t)item_begin; sp.n_items = (uint32_t)n;
    sp.item_min = s->d_item_min; sp.ctr = s->d_ctr;
    {
        const char* ev = getenv("THETA_VERIFY_APPROX");
        sp.verify = (ev && ev[0] == '1') ? 1 : 0;
        const char* es = getenv("THETA_N3_SCREEN");  // 0: off, 1: fp32 logarithms, 2: fp64 table logarithms
        sp.screen = es ? atoi(es) : kScreenDefault;
    }
    cudaEvent_t e0, e1;
    THETA_CUDA(cudaEventCreate(&e0)); THETA_CUDA(cudaEventCreate(&e1));
    THETA_CUDA(cudaEventRecord(e0, st));
    rc = launch_n3<false>(pl, sp, st);
    if (rc) return rc;
    THETA_CUDA(cudaEventRecord(e1, st));
    unsigned long long ctr[8];
    THETA_CUDA(cudaMemcpyAsync(s->h_item_min.data(), s->d_item_min, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    THETA_CUDA(cudaMemcpyAsync(ctr, s->d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, st));
    THETA_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&pl->stats.ms_pass1, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    double g = INFINITY;
    for (double v : s->h_item_min) if (v < g) g = v;
    pl->stats.n_candidates = ctr[5];
    pl->stats.n_solved = ctr[1];  // of the leaves that were solved exactly (all of them with THETA_N3_SCREEN=0)
    pl->stats.n_exact = ctr[6];
    pl->stats.n_bound_violations = ctr[7];
    pl->stats.n_stripes = n;
    pl->stats.newton_iters = ctr[1] ? (double)ctr[2] / (double)ctr[1] : 0.0;
    pl->stats.min_nll = g;
    s->nan_nll = ctr[4];
    *local_min = g;
    return THETA_OK;
}

static int run_emit(theta_plan* pl, const std::vector<uint32_t>& list, const std::vector<double>& prior, double near_thr,
                    int enumerate_only, std::vector<theta_cand>* recs, std::vector<uint8_t>* paths, cudaStream_t st) {
    theta_n3_state* s = pl->n3;
    const int m = pl->m;
    recs->clear(); paths->clear();
    if (list.empty()) return THETA_OK;
    // Pass 2 walks only a handful of items, one thread each: cut every listed item into sub-trees of at most
    // kSubCap leaves (same expander, started from the item's prefix) so that the walk is spread over many threads.
    // Records keep the enumeration order through rank = (parent item << 32) | (sub-item << 12) | leaf-in-sub-item.
    constexpr double kSubCap = 16.0;
    std::vector<uint8_t> sub_prefix;
    std::vector<uint32_t> sub_off(1, 0);
    std::vector<double> sub_weight, sub_prior;
    std::vector<uint64_t> sub_rank;
    const bool split = !enumerate_only && !s->est.empty();
    if (split) {
        Expander ex{pl, s};
        ex.cap = kSubCap;
        ex.max_items = (size_t)1 << 30;
        ex.o_prefix = &sub_prefix; ex.o_off = &sub_off; ex.o_weight = &sub_weight;
        for (size_t k = 0; k < list.size(); ++k) {
            const uint32_t it = list[k];
            const uint8_t* pre = s->prefix.data() + s->off[it];
            const int len = (int)(s->off[it + 1] - s->off[it]);
            const size_t first = sub_off.size() - 1;
            if (s->weight[it] <= kSubCap * 1048576.0) ex.run_from_prefix(pre, len);
            size_t n_sub = sub_off.size() - 1 - first;
            if (n_sub == 0 || n_sub >= ((size_t)1 << 20)) {  // not split (huge item, or nothing survived): the item itself
                sub_prefix.resize(sub_off[first]); sub_off.resize(first + 1); sub_weight.resize(first);
                sub_prefix.insert(sub_prefix.end(), pre, pre + len);
                sub_off.push_back((uint32_t)sub_prefix.size());
                sub_weight.push_back(s->weight[it]);
                n_sub = 1;
            }
            for (size_t j = 0; j < n_sub; ++j) {
                sub_rank.push_back(((uint64_t)it << 32) | ((uint64_t)j << 12));
                sub_prior.push_back(prior[k]);
            }
        }
    }
    const size_t n_work = split ? sub_rank.size() : list.size();
    uint32_t* d_list = nullptr; double* d_prior = nullptr; theta_cand* d_emit = nullptr; uint8_t* d_paths = nullptr;
    uint8_t* d_sub_prefix = nullptr; uint32_t* d_sub_off = nullptr; uint64_t* d_sub_rank = nullptr;
    int rc = THETA_OK;
    cudaError_t e = cudaSuccess;
    if (split) {
        e = theta::dmalloc(&d_sub_prefix, std::max<size_t>(1, sub_prefix.size()));
        if (e == cudaSuccess) e = theta::dmalloc(&d_sub_off, sizeof(uint32_t) * sub_off.size());
        if (e == cudaSuccess) e = theta::dmalloc(&d_sub_rank, sizeof(uint64_t) * n_work);
        if (e == cudaSuccess) e = theta::dmalloc(&d_prior, sizeof(double) * n_work);
        if (e == cudaSuccess) e = cudaMemcpy(d_sub_prefix, sub_prefix.data(), sub_prefix.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_sub_off, sub_off.data(), sizeof(uint32_t) * sub_off.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_sub_rank, sub_rank.data(), sizeof(uint64_t) * n_work, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_prior, sub_prior.data(), sizeof(double) * n_work, cudaMemcpyHostToDevice);
    } else {
        e = theta::dmalloc(&d_list, sizeof(uint32_t) * list.size());
        if (e == cudaSuccess) e = theta::dmalloc(&d_prior, sizeof(double) * list.size());
        if (e == cudaSuccess) e = cudaMemcpy(d_list, list.data(), sizeof(uint32_t) * list.size(), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_prior, prior.data(), sizeof(double) * list.size(), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "n3 emit setup", __FILE__, __LINE__);
    size_t cap = 1 << 14;
    for (int attempt = 0; rc == THETA_OK && attempt < 8; ++attempt) {
        theta::dfree(d_emit); theta::dfree(d_paths); d_emit = nullptr; d_paths = nullptr;
        e = theta::dmalloc(&d_emit, sizeof(theta_cand) * cap);
        if (e == cudaSuccess) e = theta::dmalloc(&d_paths, cap * (size_t)m);
        if (e == cudaSuccess) e = cudaMemsetAsync(s->d_ctr, 0, sizeof(unsigned long long) * 8, st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "n3 emit buffers", __FILE__, __LINE__); break; }
        Search3Params sp{};
        fill_dev(pl, &sp.D);
        sp.prefix = split ? d_sub_prefix : s->d_prefix; sp.off = split ? d_sub_off : s->d_off;
        sp.n_items = (uint32_t)n_work; sp.ctr = s->d_ctr;
        sp.list = split ? nullptr : d_list; sp.rank_base = split ? d_sub_rank : nullptr;
        sp.prior = d_prior; sp.near_thr = near_thr; sp.emit = d_emit; sp.emit_paths = d_paths;
        sp.emit_cap = cap; sp.enumerate_only = enumerate_only;
        {
            const char* ev = getenv("THETA_VERIFY_APPROX");
            sp.verify = (ev && ev[0] == '1') ? 1 : 0;
            const char* es = getenv("THETA_N3_SCREEN");
            sp.screen = enumerate_only ? 0 : (es ? atoi(es) : kScreenDefault);
        }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        rc = launch_n3<true>(pl, sp, st);
        cudaEventRecord(e1, st);
        unsigned long long ctr[8] = {0};
        if (rc == THETA_OK) {
            e = cudaMemcpyAsync(ctr, s->d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = cuda_fail(e, "n3 pass2", __FILE__, __LINE__);
        }
        cudaEventElapsedTime(&pl->stats.ms_pass2, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (rc) break;
        if (ctr[3] > cap) { cap = (size_t)ctr[3] + 1024; continue; }
        recs->resize(ctr[3]);
        paths->resize((size_t)ctr[3] * m);
        if (ctr[3]) {
            e = cudaMemcpy(recs->data(), d_emit, sizeof(theta_cand) * ctr[3], cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(paths->data(), d_paths, (size_t)ctr[3] * m, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = cuda_fail(e, "n3 pass2 D2H", __FILE__, __LINE__);
        }
        break;
    }
    theta::dfree(d_list); theta::dfree(d_prior); theta::dfree(d_emit); theta::dfree(d_paths);
    theta::dfree(d_sub_prefix); theta::dfree(d_sub_off); theta::dfree(d_sub_rank);
    if (rc) return rc;
    // sort by enumeration order, carrying the paths along
    std::vector<uint32_t> idx(recs->size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
    std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (*recs)[a].rank < (*recs)[b].rank; });
    std::vector<theta_cand> r2(recs->size());
    std::vector<uint8_t> p2(paths->size());
    for (size_t i = 0; i < idx.size(); ++i) {
        r2[i] = (*recs)[idx[i]];
        std::memcpy(p2.data() + i * (size_t)m, paths->data() + (size_t)(*recs)[idx[i]].flags * m, m);
        r2[i].flags = (int32_t)i;
    }
    recs->swap(r2);
    paths->swap(p2);
    return THETA_OK;
}

extern "C" int theta_search_n3_pass2(theta_plan* pl, double global_min, double prior_min, theta_cand* out_host, uint64_t cap,
                                     uint64_t* n_out, uint8_t* paths_host, void* stream) {
    int rc = check_n3(pl);
    if (rc) return rc;
    if (!n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_out = 0;
    theta_n3_state* s = pl->n3;
    const double near_thr = global_min + 2.5 * THETA_TIE_MARGIN;
    std::vector<uint32_t> list;
    std::vector<double> prior;
    double P = prior_min;
    for (uint64_t i = 0; i < s->h_item_min.size(); ++i) {
        const double v = s->h_item_min[i];
        const bool rec = v < P, near = v <= near_thr;
        if (rec || near || s->nan_nll) {
            list.push_back((uint32_t)(s->p1_begin + i));
            prior.push_back(rec ? P : -INFINITY);
        }
        if (v < P) P = v;
    }
    pl->stats.n_rescanned = list.size();
    std::vector<theta_cand> recs;
    std::vector<uint8_t> paths;
    rc = run_emit(pl, list, prior, near_thr, 0, &recs, &paths, (cudaStream_t)stream);
    if (rc) return rc;
    {   // the device emits a superset (every sub-item restarts its running minimum): keep exactly the NaN values,
        // the strict prefix minima w.r.t. everything before this shard, and the near-minimum set
        const size_t m = (size_t)pl->m;
        size_t w = 0;
        double run = prior_min;
        for (size_t i = 0; i < recs.size(); ++i) {
            const double v = recs[i].nll;
            const bool keep = (v != v) || v < run || v <= near_thr;
            if (v < run) run = v;
            if (keep) {
                if (w != i) { recs[w] = recs[i]; std::memmove(paths.data() + w * m, paths.data() + i * m, m); }
                recs[w].flags = (int32_t)w;
                ++w;
            }
        }
        recs.resize(w);
        paths.resize(w * m);
    }
    *n_out = recs.size();
    if (recs.size() > cap) { set_error("pass2 produced %zu records, buffer holds %llu", recs.size(), (unsigned long long)cap); return THETA_ERR_OVERFLOW; }
    if (out_host && !recs.empty()) std::memcpy(out_host, recs.data(), sizeof(theta_cand) * recs.size());
    if (paths_host && !paths.empty()) std::memcpy(paths_host, paths.data(), paths.size());
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
    std::vector<uint32_t> list;
    std::vector<double> prior;
    for (uint64_t i = item_begin; i < item_end; ++i) { list.push_back((uint32_t)i); prior.push_back(-INFINITY); }
    std::vector<theta_cand> recs;
    std::vector<uint8_t> paths;
    rc = run_emit(pl, list, prior, -INFINITY, 1, &recs, &paths, nullptr);
    if (rc) return rc;
    *n_out = recs.size();
    if (recs.size() > cap) { set_error("enumeration produced %zu candidates, buffer holds %llu", recs.size(), (unsigned long long)cap); return THETA_ERR_OVERFLOW; }
    if (paths_host && !paths.empty()) std::memcpy(paths_host, paths.data(), paths.size());
    return THETA_OK;
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
    if (e == cudaSuccess) {
        N3Dev D;
        fill_dev(pl, &D);
        const int grid = (int)((count + kBlock3 - 1) / kBlock3);
        n3_solve_rows_kernel<<<grid, kBlock3>>>(D, d_p, count, d_mu, d_mh, d_nll, d_st, d_vals, (int)mode);
        e = cudaGetLastError();
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
