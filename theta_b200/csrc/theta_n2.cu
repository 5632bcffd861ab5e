// theta_n2.cu -- n=2 plans, on-chip enumeration and the fused search kernels.
//
// Replaces, for n=2: Enumerator (python/Enumerator.py:39-160), Optimizer._solve_n2
// (python/Optimizer.py:90-126) and the enumerate->solve->keep-best loop of
// do_optimization_single / process_loop (python/RunTHetA.py:53-93,173-220).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "theta_n2.cuh"

namespace theta {

static constexpr int kBlock = 128;
static constexpr int kCounters = 8;
// pass-1 tuning switches (development: -DTHETA_P1_REFINE=0 / -DTHETA_P1_INCREMENTAL=0 give the round-1 kernel back)
#ifndef THETA_P1_REFINE
#define THETA_P1_REFINE 1
#endif
#ifndef THETA_P1_DEBUGSTATS
#define THETA_P1_DEBUGSTATS 0
#endif
#ifndef THETA_P1_INCREMENTAL
#define THETA_P1_INCREMENTAL 0
#endif
// device status block (u64 words): [0..7] pass-1 counters, [8..15] pass-2 counters, [16] listed stripes,
// [17] shard minimum (double bits), [18] global minimum used by pass 2 (double bits)
static constexpr int kStatusWords = 32;
static constexpr int kStList = 16, kStLocalMin = 17, kStGlobalMin = 18;

struct SearchParams {
    const double2* logtab;
    const unsigned char* tables;
    TableLayout layout;
    N2Const kc;
    uint64_t begin, end;
    int L;               // ranks per lane in one stripe (a stripe = 32 * L consecutive ranks, one warp)
    int Lrun;            // ranks per lane in one work unit (pass 1: L; pass 2: a few, for latency)
    int units_per_item;  // ceil(L / Lrun)
    uint32_t n_items;    // stripes (pass 1); pass 2 reads the number of listed stripes from n_items_dev
    const unsigned long long* n_items_dev;
    double* stripe_min;
    unsigned long long* counters;
    const uint32_t* stripe_list;
    const double* stripe_prior;
    double near_thr;
    const unsigned long long* near_thr_base;  // pass 2: global minimum (double bits) in device memory ...
    double near_margin;                       // ... near_thr = that + near_margin
    theta_cand* emit;
    unsigned long long emit_cap;
    int verify;  // THETA_VERIFY_APPROX=1: evaluate every candidate exactly and count violations of the fp32 error bound
};

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void stage_tables(unsigned char* smem, const unsigned char* g, size_t bytes) {
    const uint4* src = reinterpret_cast<const uint4*>(g);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// Persistent warps take work units from a ticket counter (no block-level synchronisation after the
// tables are staged).  A stripe is 32 * L consecutive ranks; each lane unranks the first column of
// its run once and then walks it with the odometer, warm-starting the root-finder from the roots of
// the previous two columns.  Pass 1 (EMIT=false) writes one minimum per stripe; pass 2 (EMIT=true)
// re-walks listed stripes in short runs and emits candidate records.
template <int KT, bool EMIT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) n2_search_kernel(const SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double2 s_logtab[THETA_LOGTAB_SIZE];
    // pass 2 is launched before the host knows how many stripes were listed (no host round trip between the
    // passes): the count is read from device memory and surplus CTAs leave before staging anything
    const unsigned long long n_units = (unsigned long long)(EMIT ? (uint32_t)*p.n_items_dev : p.n_items) * p.units_per_item;
    if (EMIT && (unsigned long long)blockIdx.x * (kBlock / 32) >= n_units) return;
    for (int i = threadIdx.x; i < THETA_LOGTAB_SIZE; i += blockDim.x) s_logtab[i] = p.logtab[i];
    stage_tables(smem, p.tables, p.layout.bytes);
    Tables T_;
    T_.bind(smem, p.layout);
    const N2Const kc = p.kc;
    const int lane = threadIdx.x & 31;
    unsigned int solved = 0, iters = 0, nan_nll = 0, n_exact = 0, n_viol = 0;  // per lane (a lane sees < 2^32 candidates)
    // Pass 1 proves most candidates irrelevant with the fp32-log NLL (n2_l2_approx): a candidate whose lower
    // bound exceeds T + margin -- T = the best exact NLL this warp has seen in this or an earlier (lower-rank)
    // stripe -- is neither the stripe minimum that matters (a strict prefix minimum of the stream is below
    // everything before it, hence below T) nor within the tie window of the global minimum (<= T).
    const double approx_E = THETA_APPROX_C * kc.Rtot;
    const double approx_margin = 3.0 * THETA_TIE_MARGIN;
    const double inv_A = 1.0 / kc.A;
    const double near_thr = EMIT ? __longlong_as_double((long long)*p.near_thr_base) + p.near_margin : 0.0;
    double T = INFINITY;

    for (;;) {
        unsigned long long unit = 0;
        if (lane == 0) unit = atomicAdd(&p.counters[0], 1ULL);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        const uint32_t item = (uint32_t)(unit / p.units_per_item);
        const uint32_t sub = (uint32_t)(unit % p.units_per_item);
        const uint32_t stripe = EMIT ? p.stripe_list[item] : item;
        const double prior = EMIT ? p.stripe_prior[item] : 0.0;
        const uint64_t s_begin = p.begin + (uint64_t)stripe * 32u * (uint64_t)p.L;
        const uint64_t s_end = min(p.end, (uint64_t)(s_begin + (uint64_t)32 * (uint64_t)p.L));
        const uint64_t r0 = s_begin + ((uint64_t)sub * 32u + lane) * (uint64_t)p.Lrun;
        double tmin = INFINITY;
        const int cnt = r0 < s_end ? (int)min((uint64_t)p.Lrun, s_end - r0) : 0;
        int q[KT];
        if (cnt) n2_unrank<KT>(T_, r0, q);
        if constexpr (EMIT) {
            // Pass 2: a unit is 32 lanes x Lrun (<= 4) consecutive ranks, lane-major.  Phase 1 evaluates every candidate of
            // the unit (warm-started); phase 2 decides what to emit from the EXACT prefix minimum of the unit in rank order
            // (an exclusive prefix-min over the lanes plus the lane's own earlier columns) on top of the stripe's prior --
            // with only a per-lane running minimum every lane of a record stripe emitted a column or two, ~1 400 records
            // per shard on config #2 against the ~100 that matter; phase 3 walks the run again and solves the few emitted
            // columns from the cold start (the record must not depend on how the range was cut: SURVEY 8e).
            constexpr int kRunMax = 4;  // >= Lrun of every pass-2 launch (theta_search_n2_pass2_*)
            double nl[kRunMax];
            float xpf = -1.0f;  // previous root (phase 1 only steers the decision: the fp32-refined solver of pass 1 will do)
            double lanemin = INFINITY;
            unsigned solvedmask = 0u;  // a None candidate is never emitted (not even when nothing is solved and near_thr = +inf)
#pragma unroll
            for (int t = 0; t < kRunMax; ++t) {
                nl[t] = INFINITY;
                if (t < cnt) {
                    double R[KT + 1], S1;
                    bool z;
                    n2_groups<KT>(T_, q, R, S1, z);
                    SolveOut o;
                    float xf;
                    if (n2_root_p1<KT + 1>(R, S1, z, kc, xpf, o, xf)) {
                        xpf = xf;
                        solvedmask |= 1u << t;
                        nl[t] = n2_l2<KT + 1>(R, S1, o.mu, kc, s_logtab);
                        lanemin = fmin(lanemin, nl[t]);  // (fmin ignores a NaN NLL)
                    }
                    if (t + 1 < cnt) n2_next<KT>(T_, q);
                }
            }
            double run = lanemin;  // inclusive prefix-min over the lanes ...
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, run, o);
                if (lane >= o) run = fmin(run, y);
            }
            run = __shfl_up_sync(0xffffffffu, run, 1);  // ... made exclusive
            if (lane == 0) run = INFINITY;
            run = fmin(run, prior);
            unsigned emitmask = 0u;
#pragma unroll
            for (int t = 0; t < kRunMax; ++t) {
                const double v = nl[t];
                if ((solvedmask >> t) & 1u) {
                    if (v != v) emitmask |= 1u << t;  // isClose(nan, x) is True: a NaN NLL is always appended (Q6)
                    else if (v <= near_thr || v < run) emitmask |= 1u << t;
                    run = fmin(run, v);
                }
            }
            if (emitmask) {
                n2_unrank<KT>(T_, r0, q);
                for (int t = 0; t < cnt; ++t) {
                    if ((emitmask >> t) & 1u) {
                        double R[KT + 1], S1;
                        bool z;
                        n2_groups<KT>(T_, q, R, S1, z);
                        SolveOut oc;
                        n2_solve<KT + 1>(R, S1, z, kc, -1.0, oc, s_logtab);
                        if (oc.nll != oc.nll) ++nan_nll;
                        const unsigned long long slot = atomicAdd(&p.counters[3], 1ULL);
                        if (slot < p.emit_cap) {
                            theta_cand c;
                            c.rank = r0 + t;
                            c.nll = oc.nll;
                            c.mu[0] = oc.mu;
                            c.mu[1] = 1.0 - oc.mu;
                            c.mu[2] = 0.0;
                            c.status = oc.status;
                            c.flags = oc.iters;
                            p.emit[slot] = c;
                        }
                    }
                    if (t + 1 < cnt) n2_next<KT>(T_, q);
                }
            }
        } else if constexpr (!THETA_P1_REFINE) {
            double xprev = -1.0, xprev2 = -1.0;  // roots of the previous two columns (warm start)
            // pass 1 runs the same trip count in every lane so that the warp can pool its thresholds every 8
            // columns (candidates of the same stripe: the stripe is re-walked by pass 2 whenever it matters)
            const int trips = EMIT ? cnt : p.Lrun;
            for (int t = 0; t < trips; ++t) {
                if (!EMIT && (t & 7) == 0 && t) T = warp_min(T);
                if (t >= cnt) continue;
                double R[KT + 1], S1;
                bool z;
                n2_groups<KT>(T_, q, R, S1, z);
                SolveOut o;
                // consecutive columns differ in one breakpoint, so the root moves smoothly: start Newton
                // from the linear extrapolation of the last two roots (falls back to the last root)
                double x0 = xprev;
                if (xprev2 > 0.0) {
                    const double xe = 2.0 * xprev - xprev2;
                    if (xe > 0.0 && xe < 1.0) x0 = xe;
                }
                if (n2_root<KT + 1>(R, S1, z, kc, x0, o)) {
                    ++solved;
                    iters += o.iters;
                    xprev2 = xprev;
                    xprev = o.x;
                    bool exact = true;
                    double na = 0.0;
                    if (!EMIT && kc.tau * o.mu >= THETA_APPROX_MIN_TM) {
                        na = n2_l2_approx<KT + 1>(R, S1, o.mu, kc, inv_A);
                        if (na - approx_E > T + approx_margin && !p.verify) exact = false;  // (NaN compares false)
                    }
                    if (exact) {
                        const double nll = n2_l2<KT + 1>(R, S1, o.mu, kc, s_logtab);
                        if (!EMIT) {
                            ++n_exact;
                            if (p.verify && kc.tau * o.mu >= THETA_APPROX_MIN_TM && !(fabs(na - nll) <= approx_E)) ++n_viol;
                        }
                        bool emit = false;
                        if (nll != nll) {
                            ++nan_nll;
                            emit = true;  // isClose(nan, x) is True: a NaN NLL is always appended (Q6)
                        } else {
                            if (EMIT) emit = (nll <= near_thr) || (nll < prior && nll < tmin);
                            tmin = fmin(tmin, nll);
                            T = fmin(T, nll);
                        }
                        if (EMIT && emit) {
                            // The record must not depend on the warm start (i.e. on how the rank range was cut into
                            // shards / stripes / runs): solve the emitted candidate again from the cold start, so
                            // that G GPUs return bit-identical (mu, NLL) to one GPU.
                            SolveOut oc;
                            n2_solve<KT + 1>(R, S1, z, kc, -1.0, oc, s_logtab);
                            const unsigned long long slot = atomicAdd(&p.counters[3], 1ULL);
                            if (slot < p.emit_cap) {
                                theta_cand c;
                                c.rank = r0 + t;
                                c.nll = oc.nll;
                                c.mu[0] = oc.mu;
                                c.mu[1] = 1.0 - oc.mu;
                                c.mu[2] = 0.0;
                                c.status = oc.status;
                                c.flags = oc.iters;
                                p.emit[slot] = c;
                            }
                        }
                    }
                }
                if (t + 1 < cnt) n2_next<KT>(T_, q);
            }
        } else {
            // Pass 1.  Everything that only steers the root-finder lives in fp32: the predicted start of a column is the
            // previous root plus the previous increment scaled by the ratio of the S1 increments (moving interval i from
            // copy number 0 to 1 shifts the root by about gamma * rN_i with gamma varying slowly along the row), the first
            // two columns after an odometer carry start from the same positions of the previous row, and n2_root_p1 refines
            // the start in fp32 before its single fp64 evaluation.
            double R[KT + 1], S1 = 0.0, S1prev = 0.0, S1rest = 0.0, sr2 = 0.0;
            bool z = false, fresh = true;
            int rowpos = 0;                                  // position in the current odometer row (saturates at 2)
            float xpf = -1.0f, dxf = 0.0f, dsf_prev = 0.0f;  // previous root, its increment, the S1 increment that came with it
            float xrowf = -1.0f, drowf = 0.0f, d1rowf = 0.0f;  // first root of the row, row-to-row trend, first increment
            const double srm = T_.SR[T_.m];
            const int qhi0 = T_.qhi[0];
            // t = -1 is a probe: the root of the first column of the row the run starts in (q_1 at its lowest admissible
            // value, higher levels as unranked), so that the first odometer carry of the run already has the previous
            // row's first root to start from.  Root only -- the probe column is not a candidate of this run.
            const int q0_first = q[0];
            for (int t = -1; t < p.Lrun; ++t) {  // same trip count in every lane: the warp pools its thresholds every 8 columns
                if ((t & 7) == 0 && t) T = warp_min(T);
                if (t >= cnt || cnt == 0) continue;
                const bool probe = t < 0;
                if (probe) q[0] = max(T_.qlo[0], q[1]);
                if (THETA_P1_INCREMENTAL == 0 || fresh) {
                    n2_groups<KT>(T_, q, R, S1, z);
                    if (THETA_P1_INCREMENTAL) {
                        S1rest = S1 - T_.SRN[q[0]];  // exact integers: any order of summation gives the same S1
                        sr2 = T_.SR[q[KT >= 2 ? 1 : 0]];
                        fresh = false;
                    }
                }
                const float dsf = (float)(S1 - S1prev);
                S1prev = S1;
                float x0f;
                if (rowpos >= 2) x0f = fmaf(dxf, dsf * rcp32(dsf_prev), xpf);
                else if (rowpos == 0) x0f = xrowf + drowf;
                else x0f = xpf + d1rowf;
                if (!(x0f > 0.0f && x0f < 1.0f)) x0f = xpf;  // (xpf < 0: cold start)
                SolveOut o;
                float xf;
#if THETA_P1_DEBUGSTATS
                int used32 = 0;
                const int dcat = t == 0 ? 0 : t == 1 ? 1 : rowpos == 0 ? 2 : rowpos == 1 ? 3 : 4;
                const bool solved_now = n2_root_p1<KT + 1>(R, S1, z, kc, x0f, o, xf, &used32);
                atomicAdd(&p.counters[20 + dcat], 1ULL);
                atomicAdd(&p.counters[25 + dcat], (unsigned long long)used32);
                if (dcat == 4 && used32 >= 2) atomicAdd(&p.counters[30], 1ULL);
                if (dcat == 4 && used32 >= 3) atomicAdd(&p.counters[31], 1ULL);
                if (solved_now) {
#else
                if (n2_root_p1<KT + 1>(R, S1, z, kc, x0f, o, xf)) {
#endif
                    if (probe) {
                        xrowf = xf;
                        xpf = xf;  // (a poor but valid start for the run's first column when it is not the row's first)
                    } else {
                    ++solved;
                    iters += o.iters;
                    const float dnew = xf - xpf;
                    if (rowpos == 0) {
                        drowf = (xrowf > 0.0f && t > 0) ? xf - xrowf : 0.0f;
                        xrowf = xf;
                        rowpos = 1;
                    } else if (rowpos == 1) {
                        d1rowf = dnew;
                        rowpos = 2;
                    }
                    dxf = dnew;
                    dsf_prev = dsf;
                    xpf = xf;
                    bool exact = true;
                    double na = 0.0;
                    if (kc.tau * o.mu >= THETA_APPROX_MIN_TM) {
                        na = n2_l2_approx<KT + 1>(R, S1, o.mu, kc, inv_A);
                        if (na - approx_E > T + approx_margin && !p.verify) exact = false;  // (NaN compares false)
                    }
                    if (exact) {
                        const double nll = n2_l2<KT + 1>(R, S1, o.mu, kc, s_logtab);
                        ++n_exact;
                        if (p.verify && kc.tau * o.mu >= THETA_APPROX_MIN_TM && !(fabs(na - nll) <= approx_E)) ++n_viol;
                        if (nll != nll) {
                            ++nan_nll;
                        } else {
                            tmin = fmin(tmin, nll);
                            T = fmin(T, nll);
                        }
                    }
                    }
                }
                if (probe) {
                    // back to the run's first column: a row start itself (rowpos 0, exact start) or somewhere inside the row
                    rowpos = (q0_first == q[0]) ? 0 : 2;
                    q[0] = q0_first;
                    fresh = true;
                    continue;
                }
                if (t + 1 < cnt) {
                    if (q[0] < qhi0) {
                        // the common odometer step moves one interval from copy number 0 to 1: only R_0, R_1, S1 change
                        const int q0 = ++q[0];
                        if (THETA_P1_INCREMENTAL) {
                            const double sr1 = T_.SR[q0];
                            R[0] = srm - sr1;
                            R[1] = sr1 - sr2;
                            S1 = S1rest + T_.SRN[q0];
                            z = T_.Z0[q0] > 0;
                        }
                    } else {
                        n2_next<KT>(T_, q);
                        fresh = true;
                        rowpos = 0;
                    }
                }
            }
        }
        if (!EMIT) {
            const double w = warp_min(tmin);
            if (lane == 0) p.stripe_min[stripe] = w;
            T = fmin(warp_min(T), w);  // every lane continues with the warp's best
        }
    }
    if (!EMIT) {
        const unsigned long long w_solved = warp_sum((unsigned long long)solved), w_iters = warp_sum((unsigned long long)iters);
        const unsigned long long w_nan = warp_sum((unsigned long long)nan_nll), w_exact = warp_sum((unsigned long long)n_exact);
        const unsigned long long w_viol = warp_sum((unsigned long long)n_viol);
        if (lane == 0) {
            if (w_solved) atomicAdd(&p.counters[1], w_solved);
            if (w_iters) atomicAdd(&p.counters[2], w_iters);
            if (w_nan) atomicAdd(&p.counters[4], w_nan);
            if (w_exact) atomicAdd(&p.counters[5], w_exact);
            if (w_viol) atomicAdd(&p.counters[6], w_viol);
        }
    }
}

// Columns of consecutive ranks as int8 rows (Enumerator._C_to_array, Enumerator.py:154-160, tumour column only).
// Lane l of a warp owns row (warp_base + l); rows are written cooperatively so stores are coalesced.
template <int KT>
__global__ void __launch_bounds__(kBlock) n2_enumerate_kernel(const unsigned char* tables, TableLayout layout,
                                                              uint64_t rank_begin, int64_t count, int8_t* out,
                                                              int64_t pitch) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_tables(smem, tables, layout.bytes);
    Tables T;
    T.bind(smem, layout);
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int m = layout.m;
    const int words = (int)(pitch >> 2);
    for (int64_t wbase = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; wbase < count; wbase += warps * 32) {
        int q[KT];
        const int64_t row = wbase + lane;
        if (row < count) n2_unrank<KT>(T, rank_begin + row, q);
        else {
#pragma unroll
            for (int v = 0; v < KT; ++v) q[v] = 0;
        }
        const int nrows = (int)min((int64_t)32, count - wbase);
        for (int j = 0; j < nrows; ++j) {
            int qj[KT];
#pragma unroll
            for (int v = 0; v < KT; ++v) qj[v] = __shfl_sync(0xffffffffu, q[v], j);
            uint32_t* dst = reinterpret_cast<uint32_t*>(out + (wbase + j) * pitch);
            for (int w = lane; w < words; w += 32) {
                uint32_t word = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int i = w * 4 + b;
                    int c = 0;
#pragma unroll
                    for (int v = 0; v < KT; ++v) c += (i >= m - qj[v]) ? 1 : 0;
                    if (i >= m) c = 0;
                    word |= (uint32_t)c << (8 * b);
                }
                dst[w] = word;
            }
        }
    }
}

// Optimizer.solve for given ranks, with the per-interval `vals` the reference returns (Optimizer.py:123-126).
template <int KT>
__global__ void __launch_bounds__(kBlock) n2_eval_ranks_kernel(const unsigned char* tables, TableLayout layout,
                                                               N2Const kc, const uint64_t* ranks, int64_t count,
                                                               const double* rN, theta_cand* out, double* vals) {
    extern __shared__ __align__(16) unsigned char smem[];
    stage_tables(smem, tables, layout.bytes);
    Tables T;
    T.bind(smem, layout);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int q[KT];
    n2_unrank<KT>(T, ranks[i], q);
    double R[KT + 1], S1;
    bool z;
    n2_groups<KT>(T, q, R, S1, z);
    SolveOut o;
    n2_solve<KT + 1>(R, S1, z, kc, -1.0, o);
    theta_cand c;
    c.rank = ranks[i];
    c.nll = o.nll;
    c.mu[0] = o.mu;
    c.mu[1] = 1.0 - o.mu;
    c.mu[2] = 0.0;
    c.status = o.status;
    c.flags = o.iters;
    out[i] = c;
    if (vals != nullptr && o.status < THETA_ST_NONE_SIGN) {
        // vals_i = numer_i / denom of Optimizer.L2 (Optimizer.py:187-196)
        const double mu1 = 1.0 - o.mu;
        const double D = o.mu * kc.S0 + mu1 * S1;
        const int m = layout.m;
        for (int j = 0; j < m; ++j) {
            int cv = 0;
#pragma unroll
            for (int v = 0; v < KT; ++v) cv += (j >= m - q[v]) ? 1 : 0;
            vals[i * m + j] = rN[j] * (kc.tau * o.mu + (double)cv * mu1) / D;
        }
    }
}

// ---------------------------------------------------------------- dispatch on the level count
#define THETA_KT_DISPATCH(KT_RUNTIME, ...)                 \
    switch (KT_RUNTIME) {                                    \
        case 2: { constexpr int KT = 2; __VA_ARGS__; } break;       \
        case 3: { constexpr int KT = 3; __VA_ARGS__; } break;       \
        case 4: { constexpr int KT = 4; __VA_ARGS__; } break;       \
        case 5: { constexpr int KT = 5; __VA_ARGS__; } break;       \
        case 6: { constexpr int KT = 6; __VA_ARGS__; } break;       \
        case 7: { constexpr int KT = 7; __VA_ARGS__; } break;       \
        case 8: { constexpr int KT = 8; __VA_ARGS__; } break;       \
        case 10: { constexpr int KT = 10; __VA_ARGS__; } break;     \
        case 12: { constexpr int KT = 12; __VA_ARGS__; } break;     \
        case 15: { constexpr int KT = 15; __VA_ARGS__; } break;     \
        default: set_error("unsupported level count %d", KT_RUNTIME); return THETA_ERR_UNSUPPORTED; \
    }

static int pick_KT(int K) {
    static const int opts[] = {2, 3, 4, 5, 6, 7, 8, 10, 12, 15};
    for (int o : opts)
        if (K <= o) return o;
    return -1;
}

template <typename Kern>
static int config_kernel(Kern kern, size_t smem, int* blocks_per_sm) {
    if (smem > 48 * 1024) THETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    THETA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, kBlock, smem));
    if (*blocks_per_sm < 1) *blocks_per_sm = 1;
    return THETA_OK;
}

static int sm_count(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms;
}

template <int KT, bool EMIT, int MINB>
static int launch_search_v(const theta_plan* plan, const SearchParams& sp, cudaStream_t st) {
    int bps = 1;
    int rc = config_kernel(n2_search_kernel<KT, EMIT, MINB>, plan->layout.bytes, &bps);
    if (rc) return rc;
    const uint64_t units = (uint64_t)sp.n_items * sp.units_per_item;  // pass 2: n_items = upper bound (all stripes)
    const uint32_t grid = (uint32_t)std::min<uint64_t>((units + kBlock / 32 - 1) / (kBlock / 32), (uint64_t)sm_count(plan->device) * bps);
    if (grid == 0) return THETA_OK;
    n2_search_kernel<KT, EMIT, MINB><<<grid, kBlock, plan->layout.bytes, st>>>(sp);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

// THETA_SEARCH_MINB (tuning knob): resident CTAs per SM the pass-1 kernel is compiled for (6 -> 80 registers, 5 -> 96, 4 -> 128);
// unset: chosen by the level count
static int g_search_minb = -1;
static int search_minb(int KT) {
    if (g_search_minb < 0) {
        const char* e = getenv("THETA_SEARCH_MINB");
        g_search_minb = e ? atoi(e) : 0;
    }
    const int v = g_search_minb ? g_search_minb : (KT <= 5 ? 5 : 6);
    return v <= 4 ? 4 : (v == 5 ? 5 : 6);
}

template <int KT, bool EMIT>
static int launch_search(const theta_plan* plan, const SearchParams& sp, cudaStream_t st) {
    if (EMIT) return launch_search_v<KT, EMIT, 4>(plan, sp, st);  // pass 2 is a few hundred warps: registers before occupancy
    // measured on B200 (pass 1, m=150): K=5 0.555 ms at 5 CTAs/SM vs 0.562 at 6; K=7 14.0 ms at 5 vs 13.5 at 6
    switch (search_minb(KT)) {
        case 4: return launch_search_v<KT, false, 4>(plan, sp, st);
        case 5: return launch_search_v<KT, false, 5>(plan, sp, st);
        default: return launch_search_v<KT, false, 6>(plan, sp, st);
    }
}

}  // namespace theta

using namespace theta;

// ======================================================================= plan
extern "C" int theta_plan_create(const theta_problem* p, theta_plan** out) {
    if (!p || !out) { set_error("null argument"); return THETA_ERR_INVALID; }
    *out = nullptr;
    if (p->n != 2 && p->n != 3) { set_error("n must be 2 or 3 (FileIO.py:42), got %d", p->n); return THETA_ERR_INVALID; }
    if (p->m < 1 || !p->r || !p->rN || !p->lower || !p->upper) { set_error("bad problem"); return THETA_ERR_INVALID; }
    if (p->tau < 1) { set_error("tau must be >= 1"); return THETA_ERR_INVALID; }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError())); return THETA_ERR_CUDA; }
    theta_plan* pl = new theta_plan();
    static std::atomic<uint64_t> next_uid{1};
    pl->uid = next_uid.fetch_add(1);
    pl->device = dev;
    pl->n = p->n; pl->m = p->m; pl->tau = p->tau; pl->max_normal = p->max_normal;
    const int m = p->m;
    pl->lower.assign(p->lower, p->lower + m);
    pl->upper.assign(p->upper, p->upper + m);
    pl->r.assign(p->r, p->r + m);
    pl->rN.assign(p->rN, p->rN + m);
    // Enumerator._check_bound_order (Enumerator.py:90-113)
    for (int i = 1; i < m; ++i) if (pl->lower[i] < pl->lower[i - 1]) pl->lower[i] = pl->lower[i - 1];
    for (int i = m - 2; i >= 0; --i) if (pl->upper[i] > pl->upper[i + 1]) pl->upper[i] = pl->upper[i + 1];
    int K = 0;
    bool empty = false;
    // n=2: the reference never checks its start vector (the lower bounds) against the upper bounds --
    // Enumerator._generate_next_C_2 (Enumerator.py:119-152) just never increments a position that is already at or
    // above its upper bound -- so where lower > upper the position is pinned to its lower bound: effective upper =
    // max(upper, lower) (golden: tests/golden/bounds_edge_n2.json).  The bounds reported back stay the monotonised ones.
    std::vector<int32_t> ue(pl->upper);
    for (int i = 0; i < m; ++i) {
        if (pl->lower[i] < 0) { set_error("negative lower bound"); delete pl; return THETA_ERR_INVALID; }
        if (pl->lower[i] > pl->upper[i]) {
            if (p->n == 2) ue[i] = pl->lower[i];
            else empty = true;
        }
        K = std::max(K, ue[i]);
        if (pl->r[i] < 0 || pl->rN[i] < 0) { set_error("negative read count"); delete pl; return THETA_ERR_INVALID; }
    }
    if (K > THETA_MAX_COPY) {
        set_error("copy number bound %d exceeds THETA_MAX_COPY=%d (n=3: 7): tighten the bounds (--MAX_K, interval selection)", K, THETA_MAX_COPY);
        delete pl;
        return THETA_ERR_UNSUPPORTED;
    }
    pl->K = K;
    if (p->n == 3) {
        if (empty) { set_error("No valid Copy Number Profiles exist for these intervals within the bounds specified"); delete pl; return THETA_ERR_EMPTY; }
        int rc3 = theta_n3_plan_init(pl);
        if (rc3 != THETA_OK) { theta_plan_destroy(pl); return rc3; }
        *out = pl;
        return THETA_OK;
    }
    pl->KT = pick_KT(std::max(K, 1));
    const int KT = pl->KT;
    // constants
    long double A = 0, Rt = 0;
    int zero_rN = 0;
    for (int i = 0; i < m; ++i) { A += pl->rN[i]; Rt += pl->r[i]; if (pl->rN[i] == 0) zero_rN = 1; }
    if (A >= 9.0e15L / std::max(1, std::max(K, p->tau)) || Rt >= 9.0e15L) { set_error("read counts too large for exact fp64 sums"); delete pl; return THETA_ERR_UNSUPPORTED; }
    long double K0 = 0;
    for (int i = 0; i < m; ++i)
        if (pl->r[i] > 0 && pl->rN[i] > 0) K0 -= (long double)pl->r[i] * logl((long double)pl->rN[i] / A);
    pl->kc.A = (double)A; pl->kc.S0 = (double)(A * p->tau); pl->kc.Rtot = (double)Rt; pl->kc.K0 = (double)K0;
    pl->kc.tau = (double)p->tau; pl->kc.max_normal = p->max_normal; pl->kc.min_normal = p->min_normal; pl->kc.any_zero_rN = zero_rN; pl->kc.m = m;
    // tables
    pl->layout = TableLayout(m, KT);
    pl->h_tables.assign(pl->layout.bytes, 0);
    std::vector<unsigned char>& blob = pl->h_tables;
    double* SR = reinterpret_cast<double*>(blob.data() + pl->layout.off_SR);
    double* SRN = reinterpret_cast<double*>(blob.data() + pl->layout.off_SRN);
    uint64_t* SW = reinterpret_cast<uint64_t*>(blob.data() + pl->layout.off_SW);
    int32_t* Z0 = reinterpret_cast<int32_t*>(blob.data() + pl->layout.off_Z0);
    int32_t* qlo = reinterpret_cast<int32_t*>(blob.data() + pl->layout.off_qlo);
    int32_t* qhi = reinterpret_cast<int32_t*>(blob.data() + pl->layout.off_qhi);
    {
        int64_t s = 0, sn = 0;
        SR[0] = 0; SRN[0] = 0;
        for (int q = 1; q <= m; ++q) { s += pl->r[m - q]; sn += pl->rN[m - q]; SR[q] = (double)s; SRN[q] = (double)sn; }
        int z = 0;
        std::vector<int> zp(m + 1, 0);
        for (int i = 0; i < m; ++i) { zp[i + 1] = zp[i] + (pl->r[i] == 0 ? 1 : 0); }
        (void)z;
        for (int q = 0; q <= m; ++q) Z0[q] = zp[m - q];
    }
    for (int v = 1; v <= KT; ++v) {
        int lo = 0, hi = 0;
        for (int i = 0; i < m; ++i) { if (pl->lower[i] >= v) ++lo; if (ue[i] >= v) ++hi; }
        qlo[v - 1] = lo; qhi[v - 1] = hi;
    }
    // W_1(q) = 1;  W_v(q) = SW_{v-1}[max(q, qlo_{v-1})]   (0 beyond qhi_{v-1})
    bool overflow = false;
    const unsigned __int128 LIM = ((unsigned __int128)1) << 63;
    for (int v = 1; v <= KT; ++v) {
        uint64_t* sw = SW + (size_t)(v - 1) * (m + 2);
        const uint64_t* swp = v > 1 ? SW + (size_t)(v - 2) * (m + 2) : nullptr;
        sw[m + 1] = 0;
        unsigned __int128 acc = 0;
        for (int q = m; q >= 0; --q) {
            uint64_t w;
            if (v == 1) w = 1;
            else { const int x = std::max(q, qlo[v - 2]); w = (x <= qhi[v - 2]) ? swp[x] : 0; }
            if (q <= qhi[v - 1]) acc += w;
            if (acc >= LIM) { overflow = true; acc = LIM - 1; }
            sw[q] = (uint64_t)acc;
        }
    }
    // THETA_PLAN_TILES_ONLY: the plan serves explicit columns only (Optimizer.solve / CalcAllC): it needs r, rN, the
    // prefix tables and KT, not the count tables -- so a space too large to count (m >~ 110 with bounds [0, 15]) is fine
    const bool tiles_only = (p->reserved & THETA_PLAN_TILES_ONLY) != 0;
    if (overflow && !tiles_only) { set_error("candidate count exceeds 2^63"); delete pl; return THETA_ERR_OVERFLOW; }
    if (overflow) { pl->count = 0; }
    else {
        const uint64_t* swK = SW + (size_t)(KT - 1) * (m + 2);
        pl->count = (empty || qlo[KT - 1] > qhi[KT - 1]) ? 0 : swK[qlo[KT - 1]];
        if (empty) pl->count = 0;
    }
    // One device arena and one H2D copy (from a pinned staging block) for everything the kernels read:
    //   tables | r | rN | PR | PRN | ZP | counters.   The fast_log table is per device, uploaded once.
    const int64_t pitch = theta_tile_pitch(m);
    pl->tile_pitch = pitch;
    auto up256 = [](size_t o) { return (o + 255) & ~(size_t)255; };
    size_t o = 0;
    const size_t o_tab = o; o = up256(o + pl->layout.bytes);
    const size_t o_r = o;   o = up256(o + sizeof(double) * m);
    const size_t o_rN = o;  o = up256(o + sizeof(double) * m);
    const size_t o_PR = o;  o = up256(o + sizeof(double) * (pitch + 1));
    const size_t o_PRN = o; o = up256(o + sizeof(double) * (pitch + 1));
    const size_t o_ZP = o;  o = up256(o + sizeof(int32_t) * (pitch + 1));
    const size_t o_ctr = o; o = up256(o + sizeof(unsigned long long) * kStatusWords);
    const size_t arena_bytes = o;
    unsigned char* stage = nullptr;
    cudaError_t e = theta::hmalloc(&stage, arena_bytes);
    if (e == cudaSuccess) {
        std::memset(stage, 0, arena_bytes);
        std::memcpy(stage + o_tab, blob.data(), pl->layout.bytes);
        double* rd = reinterpret_cast<double*>(stage + o_r);
        double* rnd = reinterpret_cast<double*>(stage + o_rN);
        for (int i = 0; i < m; ++i) { rd[i] = (double)pl->r[i]; rnd[i] = (double)pl->rN[i]; }
        // prefix tables for the tile kernels (theta_tiles.cu), flat past m
        double* PR = reinterpret_cast<double*>(stage + o_PR);
        double* PRN = reinterpret_cast<double*>(stage + o_PRN);
        int32_t* ZP = reinterpret_cast<int32_t*>(stage + o_ZP);
        int64_t s = 0, sn = 0;
        int z = 0;
        for (int64_t i = 0; i < pitch; ++i) {
            if (i < m) { s += pl->r[i]; sn += pl->rN[i]; z += pl->r[i] == 0 ? 1 : 0; }
            PR[i + 1] = (double)s; PRN[i + 1] = (double)sn; ZP[i + 1] = z;
        }
        e = theta::dmalloc(&pl->d_arena, arena_bytes);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(pl->d_arena, stage, arena_bytes, cudaMemcpyHostToDevice, nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    theta::hfree(stage);
    if (e == cudaSuccess) {
        pl->d_tables = pl->d_arena + o_tab;
        pl->d_r = reinterpret_cast<double*>(pl->d_arena + o_r);
        pl->d_rN = reinterpret_cast<double*>(pl->d_arena + o_rN);
        pl->d_PR = reinterpret_cast<double*>(pl->d_arena + o_PR);
        pl->d_PRN = reinterpret_cast<double*>(pl->d_arena + o_PRN);
        pl->d_ZP = reinterpret_cast<int32_t*>(pl->d_arena + o_ZP);
        pl->d_counters = reinterpret_cast<unsigned long long*>(pl->d_arena + o_ctr);
        e = theta::logtab_device(&pl->d_logtab);
    }
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "plan allocation", __FILE__, __LINE__);
        theta_plan_destroy(pl);
        return rc;
    }
    *out = pl;
    return THETA_OK;
}

extern "C" void theta_plan_destroy(theta_plan* pl) {
    if (!pl) return;
    theta_n3_plan_free(pl);
    theta::dfree(pl->d_arena);  // tables, r, rN, PR, PRN, ZP, counters (d_logtab is the per-device table: not owned)
    theta::dfree(pl->d_stripe_min); theta::dfree(pl->d_emit);
    theta::dfree(pl->d_list); theta::dfree(pl->d_prior);
    theta::hfree(pl->h_emit); theta::hfree(pl->h_ctr);
    theta::dfree(pl->d_xblk); theta::hfree(pl->h_xblk); theta::dfree(pl->d_fscratch);
    if (pl->ev0) { cudaEventDestroy(pl->ev0); cudaEventDestroy(pl->ev1); cudaEventDestroy(pl->ev2); cudaEventDestroy(pl->ev3); }
    delete pl;
}

extern "C" int theta_plan_count(const theta_plan* pl, uint64_t* count) {
    if (!pl || !count) { set_error("null argument"); return THETA_ERR_INVALID; }
    *count = pl->count;
    return THETA_OK;
}

extern "C" int theta_plan_bounds(const theta_plan* pl, int32_t* lower, int32_t* upper) {
    if (!pl) { set_error("null plan"); return THETA_ERR_INVALID; }
    if (lower) std::memcpy(lower, pl->lower.data(), sizeof(int32_t) * pl->m);
    if (upper) std::memcpy(upper, pl->upper.data(), sizeof(int32_t) * pl->m);
    return THETA_OK;
}

extern "C" int64_t theta_tile_pitch(int32_t m) {
    int64_t p = ((int64_t)m + 7) / 8;
    if ((p & 1) == 0) ++p;
    return p * 8;
}

// ======================================================================= enumeration
extern "C" int theta_enumerate_n2_dev(theta_plan* pl, uint64_t rank_begin, int64_t count, int8_t* cols_dev,
                                      int64_t pitch, void* stream) {
    if (!pl || !cols_dev) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (pl->n != 2) { set_error("plan is not n=2"); return THETA_ERR_INVALID; }
    if (count <= 0) return THETA_OK;
    if (pitch < pl->m || (pitch & 3)) { set_error("pitch must be >= m and a multiple of 4"); return THETA_ERR_INVALID; }
    if (rank_begin + (uint64_t)count > pl->count) { set_error("rank range exceeds candidate count"); return THETA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t warps_needed = (count + 31) / 32;
    const int grid = (int)std::min<int64_t>((warps_needed + (kBlock / 32) - 1) / (kBlock / 32), (int64_t)sm_count(pl->device) * 8);
    THETA_KT_DISPATCH(pl->KT, {
        if (pl->layout.bytes > 48 * 1024)
            THETA_CUDA(cudaFuncSetAttribute(n2_enumerate_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->layout.bytes));
        n2_enumerate_kernel<KT><<<grid, kBlock, pl->layout.bytes, st>>>(pl->d_tables, pl->layout, rank_begin, count, cols_dev, pitch);
    });
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

extern "C" int theta_enumerate_n2(theta_plan* pl, uint64_t rank_begin, int64_t count, int8_t* cols_host, int64_t pitch) {
    if (!pl || !cols_host) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (count <= 0) return THETA_OK;
    int8_t* d = nullptr;
    THETA_CUDA(theta::dmalloc(&d, (size_t)count * pitch));
    int rc = theta_enumerate_n2_dev(pl, rank_begin, count, d, pitch, nullptr);
    if (rc == THETA_OK) {
        cudaError_t e = cudaMemcpy(cols_host, d, (size_t)count * pitch, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "enumerate D2H", __FILE__, __LINE__);
    }
    theta::dfree(d);
    return rc;
}

// ======================================================================= search
namespace theta {
// min over the stripe minima -> the shard minimum in the status block (and in the caller's device double: the
// operand of the NCCL all-reduce(min))
__global__ void __launch_bounds__(1024) stripe_min_reduce_kernel(const double* __restrict__ v, uint64_t n,
                                                                 unsigned long long* status, double* out1) {
    __shared__ double s_w[32];
    double m = INFINITY;
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmin(m, v[i]);
    m = warp_min(m);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = warp_min(s_w[threadIdx.x]);
        if (threadIdx.x == 0) {
            status[kStLocalMin] = (unsigned long long)__double_as_longlong(m);
            if (out1) *out1 = m;
        }
    }
}

// Which stripes can matter (RunTHetA.py:198-206): those that hold a strict prefix minimum of the NLL stream (every
// reset of the reference fold is one) or a candidate within two margins of the global minimum (every member of
// the final tie list is).  One block: exclusive prefix-min scan over the stripe minima, unordered append.
__global__ void __launch_bounds__(1024) stripe_scan_kernel(const double* __restrict__ v, uint64_t n,
                                                           const double* __restrict__ gmin_dev, double prior_min,
                                                           double margin, unsigned long long* status,
                                                           uint32_t* __restrict__ list, double* __restrict__ prior) {
    __shared__ double s_w[32];
    __shared__ unsigned int s_count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double g = gmin_dev ? *gmin_dev : __longlong_as_double((long long)status[kStLocalMin]);
    const double near_thr = g + margin;
    const bool any_nan = status[4] != 0ULL;  // a NaN NLL is always appended (Q6): every stripe is re-walked
    if (tid == 0) s_count = 0u;
    // thread t owns the contiguous chunk [t*per, (t+1)*per): chunk minimum -> one block-wide exclusive scan ->
    // a second walk over the chunk with the running prefix (two block barriers in total)
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = min(n, (uint64_t)tid * per), hi = min(n, lo + per);
    double cm = INFINITY;
    for (uint64_t i = lo; i < hi; ++i) cm = fmin(cm, v[i]);
    double incl = cm;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = fmin(incl, y);
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w = fmin(w, y);
        }
        s_w[lane] = w;
    }
    __syncthreads();
    double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev = INFINITY;
    double P = fmin(prior_min, fmin(warp > 0 ? s_w[warp - 1] : INFINITY, prev));  // min of everything before the chunk
    for (uint64_t i = lo; i < hi; ++i) {
        const double x = v[i];
        const bool rec = x < P, near = x <= near_thr;
        if (rec || near || any_nan) {
            const unsigned int slot = atomicAdd(&s_count, 1u);
            list[slot] = (uint32_t)i;
            prior[slot] = rec ? P : -INFINITY;
        }
        P = fmin(P, x);
    }
    __syncthreads();
    if (tid == 0) {
        status[kStList] = s_count;
        status[kStGlobalMin] = (unsigned long long)__double_as_longlong(g);
    }
    if (tid < kCounters) status[kCounters + tid] = 0ULL;  // pass-2 counters
}

// The records pass 2 emits are a superset of what the fold replay needs (every lane keeps its own running minimum, and
// with prior = +inf the first stripes of a shard emit nearly all their candidates: ~1 400 records per shard on config #2
// against ~100 that matter).  Sorting and filtering them on the host (theta_filter_records) was the largest fixed cost of
// a multi-GPU step: every rank sorted the union of all shards' records, 0.4 ms at 8 GPUs.  This kernel does the filter
// where the records are: one block sorts the shard's records by rank (bitonic, shared memory), takes the exclusive
// prefix minimum of the NLL in rank order and keeps   NaN  |  NLL < min(prior, everything before)  |  NLL <= near_thr
// -- exactly theta_filter_records on this shard.  A record dropped here is dropped by the filter over the union as well
// (its running minimum only gets lower with more shards in front), so the host is left with a linear pass over the
// concatenated blocks.  Output block: the status words (kStEmitted := records kept) followed by the kept records in rank
// order.  Up to kFilterMax records are sorted in shared memory; up to kFilterMaxGlobal (the whole 1.7e11-candidate space of
// config #3 emits 4.7e4 per shard) in a global scratch area of the plan, still by this one block -- a millisecond next
// to a search of seconds, and it keeps the exchange blocks small whatever the space.  More than that (or than fit
// `cap`): the block is passed through unfiltered / flagged as overflow through kStEmitted > cap, and the caller falls
// back to the host filter / a larger exchange slot.
static constexpr int kFilterMax = 4096;
static constexpr int kFilterMaxGlobal = 1 << 16;
static constexpr size_t kFilterSmem = (2 * sizeof(unsigned long long) + sizeof(unsigned int)) * kFilterMax;
static constexpr size_t kFilterScratch = (2 * sizeof(unsigned long long) + sizeof(unsigned int)) * (size_t)kFilterMaxGlobal;
static constexpr int kStEmittedWord = kCounters + 3;
__global__ void __launch_bounds__(1024) records_filter_kernel(const theta_cand* __restrict__ emit, unsigned long long emit_cap,
                                                              const unsigned long long* __restrict__ status,
                                                              const double* __restrict__ gmin_dev, double prior_min, double margin,
                                                              unsigned char* __restrict__ out, unsigned long long cap,
                                                              unsigned char* scratch /* kFilterScratch bytes or NULL */,
                                                              int smem_max /* kFilterMax; 0 forces the global path (tests) */) {
    extern __shared__ __align__(16) unsigned char fsm[];
    __shared__ double s_w[32];
    __shared__ unsigned int s_c[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* out_status = reinterpret_cast<unsigned long long*>(out);
    theta_cand* out_recs = reinterpret_cast<theta_cand*>(out + sizeof(unsigned long long) * kStatusWords);
    const unsigned long long n_raw = status[kStEmittedWord];
    const unsigned long long n_have = n_raw < emit_cap ? n_raw : emit_cap;
    const bool in_smem = n_have <= (unsigned long long)smem_max;
    if (n_raw > emit_cap || (!in_smem && (!scratch || n_have > (unsigned long long)kFilterMaxGlobal))) {
        // too many to sort here: hand over what fits, unfiltered, with the raw count (the caller's overflow / host-filter path)
        for (int i = tid; i < kStatusWords; i += 1024) out_status[i] = status[i];
        const unsigned long long ncopy = n_have < cap ? n_have : cap;
        for (unsigned long long i = tid; i < ncopy; i += 1024) out_recs[i] = emit[i];
        return;
    }
    unsigned char* area = in_smem ? fsm : scratch;
    const size_t amax = in_smem ? (size_t)kFilterMax : (size_t)kFilterMaxGlobal;
    unsigned long long* key = reinterpret_cast<unsigned long long*>(area);                        // [amax] rank
    double* val = reinterpret_cast<double*>(area + sizeof(unsigned long long) * amax);            // [amax] NLL in rank order
    unsigned int* idx = reinterpret_cast<unsigned int*>(area + 2 * sizeof(unsigned long long) * amax);  // [amax]
    const int n = (int)n_have;
    int P = 1;
    while (P < n) P <<= 1;
    for (int i = tid; i < P; i += 1024) {
        key[i] = i < n ? emit[i].rank : ~0ULL;
        idx[i] = (unsigned int)i;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += 1024) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = key[i], b = key[l];
                    if ((a > b) == up) {
                        key[i] = b; key[l] = a;
                        const unsigned int t = idx[i]; idx[i] = idx[l]; idx[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < P; i += 1024) val[i] = i < n ? emit[idx[i]].nll : INFINITY;
    __syncthreads();
    // thread t owns the contiguous chunk [t*per, (t+1)*per) of the sorted records
    const int per = (P + 1023) / 1024;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    double cm = INFINITY;
    for (int i = lo; i < hi; ++i) { const double v = val[i]; if (v == v) cm = fmin(cm, v); }
    double incl = cm;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = fmin(incl, y);
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w = fmin(w, y);
        }
        s_w[lane] = w;
    }
    __syncthreads();
    double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev = INFINITY;
    const double g = gmin_dev ? *gmin_dev : __longlong_as_double((long long)status[kStLocalMin]);
    const double near_thr = g + margin;
    const double run0 = fmin(prior_min, fmin(warp > 0 ? s_w[warp - 1] : INFINITY, prev));  // min of everything before the chunk
    // (theta_filter_records: keep = NaN || v < run || v <= near_thr; run = min(run, v) -- a NaN never lowers it)
    unsigned int kept = 0;
    {
        double run = run0;
        for (int i = lo; i < hi; ++i) {
            const double v = val[i];
            if ((v != v) || v < run || v <= near_thr) ++kept;
            if (v < run) run = v;
        }
    }
    unsigned int pos = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, pos, o);
        if (lane >= o) pos += y;
    }
    if (lane == 31) s_c[warp] = pos;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = s_c[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_c[lane] = w;
    }
    __syncthreads();
    const unsigned int total = s_c[31];
    unsigned int at = (warp > 0 ? s_c[warp - 1] : 0u) + pos - kept;
    if ((unsigned long long)total <= cap) {
        double run = run0;
        for (int i = lo; i < hi; ++i) {
            const double v = val[i];
            if ((v != v) || v < run || v <= near_thr) out_recs[at++] = emit[idx[i]];
            if (v < run) run = v;
        }
    }
    for (int i = tid; i < kStatusWords; i += 1024) out_status[i] = i == kStEmittedWord ? (unsigned long long)total : status[i];
}
}  // namespace theta

// timing events of the plan: inside a stream capture they are recorded as external event nodes, so that they can
// still be read with cudaEventElapsedTime after the graph has run
static cudaError_t record_event(cudaEvent_t ev, cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
        return cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal);
    return cudaEventRecord(ev, st);
}

// Test knobs, read at every search: THETA_FILTER_FORCE_GLOBAL=1 sorts every record block in the global scratch area;
// THETA_FILTER_HOST=1 makes theta_search_n2_pass2_dev fetch the raw records and filter them on the host (the independent
// reference of tests/test_gpu_n2.py::test_device_record_filter_equals_host_filter).
static bool env_flag(const char* name) {
    const char* e = getenv(name);
    return e && e[0] == '1';
}
// columns per lane in one pass-2 unit (<= kRunMax of the kernel): fewer = shorter dependent chain per lane (pass 2 is a few
// hundred warps, i.e. latency), more = a longer exact prefix per unit (fewer records to filter).  THETA_P2_LRUN overrides.
static int pass2_run_length() {
    const char* e = getenv("THETA_P2_LRUN");
    const int v = e ? atoi(e) : 2;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
}
static int filter_smem_max() { return env_flag("THETA_FILTER_FORCE_GLOBAL") ? 0 : kFilterMax; }

static int plan_events(theta_plan* pl) {
    if (!pl->ev0) {
        // (per device, once per plan and outside any stream capture: the first step of a plan always runs eagerly)
        THETA_CUDA(cudaFuncSetAttribute(records_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFilterSmem));
        if (!pl->d_fscratch) THETA_CUDA(theta::dmalloc(&pl->d_fscratch, kFilterScratch));
        THETA_CUDA(cudaEventCreate(&pl->ev0)); THETA_CUDA(cudaEventCreate(&pl->ev1));
        THETA_CUDA(cudaEventCreate(&pl->ev2)); THETA_CUDA(cudaEventCreate(&pl->ev3));
    }
    if (!pl->h_ctr) THETA_CUDA(theta::hmalloc(&pl->h_ctr, sizeof(unsigned long long) * kStatusWords));
    return THETA_OK;
}

// Pass 1, asynchronous: everything is enqueued on `stream` and nothing is waited for, so the caller can
// enqueue the NCCL all-reduce(min) on `local_min_dev` right behind it.
extern "C" int theta_search_n2_pass1_begin(theta_plan* pl, uint64_t rank_begin, uint64_t rank_end, double* local_min_dev,
                                           void* stream) {
    if (!pl) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (pl->n != 2) { set_error("plan is not n=2"); return THETA_ERR_INVALID; }
    if (pl->count == 0) { set_error("No valid Copy Number Profiles exist for these intervals within the bounds specified"); return THETA_ERR_EMPTY; }
    if (rank_end > pl->count) rank_end = pl->count;
    if (rank_begin > rank_end) { set_error("bad rank range"); return THETA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = plan_events(pl);
    if (rc) return rc;
    const uint64_t N = rank_end - rank_begin;
    pl->p1_begin = rank_begin; pl->p1_end = rank_end; pl->p1_stream = stream; pl->p1_pending = true;
    pl->stats = theta_search_stats{};
    pl->stats.n_candidates = N;
    pl->stats.min_nll = INFINITY;
    // run length per lane: enough stripes (one warp each) to balance the SMs, few enough to keep the stripe table small
    const uint64_t target_stripes = (uint64_t)sm_count(pl->device) * 24 * (kBlock / 32);
    uint64_t L = N / (target_stripes * 32);
    L = std::max<uint64_t>(8, std::min<uint64_t>(L, 256));
    if (L < 256) {
        // a small space is only a handful of stripes per resident warp: pick the run length near L that minimises the columns
        // the busiest warp walks, ceil(stripes / warps) * (L + 4) (config #2: L = 55 -> 14 348 stripes, 4.85 per warp; 54 -> 4.94)
        const uint64_t warps = (uint64_t)sm_count(pl->device) * (uint64_t)search_minb(pl->KT) * (kBlock / 32);
        uint64_t best_L = L, best_cost = ~0ull;
        for (uint64_t c = L + 4; c + 12 >= L && c >= 8; --c) {
            const uint64_t st_c = (N + c * 32 - 1) / (c * 32);
            const uint64_t cost = ((st_c + warps - 1) / warps) * (c + 4);  // (+4: the probe and the cold first columns of a run)
            if (cost < best_cost) { best_cost = cost; best_L = c; }
        }
        L = std::min<uint64_t>(best_L, 256);
    }
    while ((N + L * 32 - 1) / (L * 32) > (1ull << 22)) L *= 2;
    pl->p1_L = (int)L; pl->p1_block = kBlock;
    const uint64_t stripes = (N + L * 32 - 1) / (L * 32);
    pl->p1_stripes = stripes;
    if (stripes + 1 > pl->stripe_cap) {
        theta::dfree(pl->d_stripe_min); theta::dfree(pl->d_list); theta::dfree(pl->d_prior);
        pl->d_stripe_min = nullptr; pl->d_list = nullptr; pl->d_prior = nullptr; pl->stripe_cap = 0;
        THETA_CUDA(theta::dmalloc(&pl->d_stripe_min, sizeof(double) * (stripes + 1)));
        THETA_CUDA(theta::dmalloc(&pl->d_list, sizeof(uint32_t) * (stripes + 1)));
        THETA_CUDA(theta::dmalloc(&pl->d_prior, sizeof(double) * (stripes + 1)));
        pl->stripe_cap = stripes + 1;
    }
    THETA_CUDA(cudaMemsetAsync(pl->d_counters, 0, sizeof(unsigned long long) * kStatusWords, st));
    THETA_CUDA(record_event(pl->ev0, st));
    if (N) {
        SearchParams sp{};
        sp.tables = pl->d_tables; sp.layout = pl->layout; sp.kc = pl->kc; sp.logtab = pl->d_logtab;
        sp.begin = rank_begin; sp.end = rank_end; sp.L = (int)L; sp.Lrun = (int)L; sp.units_per_item = 1;
        sp.n_items = (uint32_t)stripes;
        sp.stripe_min = pl->d_stripe_min; sp.counters = pl->d_counters;
        {
            const char* ev = getenv("THETA_VERIFY_APPROX");
            sp.verify = (ev && ev[0] == '1') ? 1 : 0;
        }
        THETA_KT_DISPATCH(pl->KT, { int rc2 = launch_search<KT, false>(pl, sp, st); if (rc2) return rc2; });
    }
    THETA_CUDA(record_event(pl->ev1, st));
    stripe_min_reduce_kernel<<<1, 1024, 0, st>>>(pl->d_stripe_min, stripes, pl->d_counters, local_min_dev);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

// The host-side part of theta_search_n2_pass1_begin when the device work is replayed from a captured graph of the very
// same call (theta_mgpu.cu): per-search bookkeeping only, nothing is enqueued.
extern "C" int theta_search_n2_graph_replay(theta_plan* pl, uint64_t rank_begin, uint64_t rank_end, void* stream) {
    if (!pl || pl->n != 2) { set_error("plan is not n=2"); return THETA_ERR_INVALID; }
    if (rank_end > pl->count) rank_end = pl->count;
    pl->p1_begin = rank_begin; pl->p1_end = rank_end; pl->p1_stream = stream; pl->p1_pending = true;
    pl->stats = theta_search_stats{};
    pl->stats.n_candidates = rank_end - rank_begin;
    pl->stats.min_nll = INFINITY;
    return THETA_OK;
}

static void read_pass1_stats(theta_plan* pl) {
    const unsigned long long* ctr = pl->h_ctr;
    cudaEventElapsedTime(&pl->stats.ms_pass1, pl->ev0, pl->ev1);
    double g;
    std::memcpy(&g, &ctr[kStLocalMin], sizeof(double));
    pl->stats.n_solved = ctr[1];
    pl->stats.n_stripes = pl->p1_stripes;
    pl->stats.newton_iters = ctr[1] ? (double)ctr[2] / (double)ctr[1] : 0.0;
    pl->stats.min_nll = g;
    pl->stats.n_exact = ctr[5];
    pl->stats.n_bound_violations = ctr[6];
#if THETA_P1_DEBUGSTATS
    fprintf(stderr, "[p1 fp32 iterations] t=0: %.2f (n=%llu)  t=1: %.2f (%llu)  row start: %.2f (%llu)  second: %.2f (%llu)  rest: %.3f (%llu; >=2: %llu, >=3: %llu)\n",
            (double)ctr[25] / (double)std::max(1ULL, ctr[20]), ctr[20], (double)ctr[26] / (double)std::max(1ULL, ctr[21]), ctr[21],
            (double)ctr[27] / (double)std::max(1ULL, ctr[22]), ctr[22], (double)ctr[28] / (double)std::max(1ULL, ctr[23]), ctr[23],
            (double)ctr[29] / (double)std::max(1ULL, ctr[24]), ctr[24], ctr[30], ctr[31]);
#endif
}

extern "C" int theta_search_n2_pass1_end(theta_plan* pl, double* local_min) {
    if (!pl || !local_min) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (!pl->p1_pending) { set_error("pass1_end without pass1_begin"); return THETA_ERR_INVALID; }
    pl->p1_pending = false;
    cudaStream_t st = (cudaStream_t)pl->p1_stream;
    THETA_CUDA(cudaMemcpyAsync(pl->h_ctr, pl->d_counters, sizeof(unsigned long long) * kStatusWords, cudaMemcpyDeviceToHost, st));
    THETA_CUDA(cudaStreamSynchronize(st));
    read_pass1_stats(pl);
    *local_min = pl->stats.min_nll;
    return THETA_OK;
}

extern "C" int theta_search_n2_pass1(theta_plan* pl, uint64_t rank_begin, uint64_t rank_end, double* local_min,
                                     void* stream) {
    if (!local_min) { set_error("null argument"); return THETA_ERR_INVALID; }
    *local_min = INFINITY;
    int rc = theta_search_n2_pass1_begin(pl, rank_begin, rank_end, nullptr, stream);
    if (rc) return rc;
    return theta_search_n2_pass1_end(pl, local_min);
}

// Pass 2 without a host round trip: the stripe list is built on the device from the stripe minima and the
// global minimum in `global_min_dev` (NULL: this shard's own minimum), the emit kernel is launched behind it,
// and the host waits once, at the end.  Also completes a pending pass1_begin.
extern "C" int theta_search_n2_pass2_dev(theta_plan* pl, const double* global_min_dev, double prior_min, theta_cand* out_host,
                                         uint64_t cap, uint64_t* n_out, double* global_min_host, void* stream) {
    if (!pl || !n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_out = 0;
    if (pl->n != 2 || !pl->ev0) { set_error("pass2 without pass1"); return THETA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t stripes = pl->p1_stripes;
    const double margin = 2.5 * THETA_TIE_MARGIN;
    size_t want = std::max<size_t>(kFilterMaxGlobal, pl->emit_cap);
    bool filtered = false;  // the records came back sorted and filtered by the device
    theta_cand* host = nullptr;
    size_t n_host = 0;
    double g = INFINITY;
    int rc = THETA_OK;
    for (int attempt = 0; attempt < 8; ++attempt) {
        if (want > pl->emit_cap) {
            theta::dfree(pl->d_emit); pl->d_emit = nullptr; pl->emit_cap = 0;
            cudaError_t e = theta::dmalloc(&pl->d_emit, sizeof(theta_cand) * want);
            if (e != cudaSuccess) { rc = cuda_fail(e, "emit buffer", __FILE__, __LINE__); break; }
            pl->emit_cap = want;
        }
        record_event(pl->ev2, st);
        // (re)builds the list and zeroes the pass-2 counters
        stripe_scan_kernel<<<1, 1024, 0, st>>>(pl->d_stripe_min, stripes, global_min_dev, prior_min, margin, pl->d_counters,
                                               pl->d_list, pl->d_prior);
        if (stripes && pl->p1_begin != pl->p1_end) {
            SearchParams sp{};
            sp.tables = pl->d_tables; sp.layout = pl->layout; sp.kc = pl->kc; sp.logtab = pl->d_logtab;
            sp.begin = pl->p1_begin; sp.end = pl->p1_end; sp.L = pl->p1_L; sp.n_items = (uint32_t)stripes;
            sp.n_items_dev = pl->d_counters + kStList;
            sp.Lrun = std::min(pl->p1_L, pass2_run_length()); sp.units_per_item = (pl->p1_L + sp.Lrun - 1) / sp.Lrun;
            sp.stripe_min = pl->d_stripe_min; sp.counters = pl->d_counters + kCounters;
            sp.stripe_list = pl->d_list; sp.stripe_prior = pl->d_prior;
            sp.near_thr = 0.0; sp.near_thr_base = pl->d_counters + kStGlobalMin; sp.near_margin = margin;
            sp.emit = pl->d_emit; sp.emit_cap = pl->emit_cap;
            switch (pl->KT) {
#define THETA_P2(KTV) case KTV: rc = launch_search<KTV, true>(pl, sp, st); break;
                THETA_P2(2) THETA_P2(3) THETA_P2(4) THETA_P2(5) THETA_P2(6) THETA_P2(7) THETA_P2(8) THETA_P2(10) THETA_P2(12) THETA_P2(15)
#undef THETA_P2
                default: set_error("unsupported level count"); rc = THETA_ERR_UNSUPPORTED;
            }
            if (rc) break;
        }
        record_event(pl->ev3, st);
        // The shard's records are sorted by rank and filtered on the device (records_filter_kernel) into one small block
        // [status | records]; the status words and the first kSpecRecs records come back in ONE copy and one wait.
        constexpr size_t kXCap = 1024, kSpecRecs = 256, kHead = sizeof(unsigned long long) * kStatusWords;
        if (!pl->d_xblk) {
            cudaError_t ea = theta::dmalloc(&pl->d_xblk, kHead + sizeof(theta_cand) * kXCap);
            if (ea == cudaSuccess) ea = theta::hmalloc(&pl->h_xblk, kHead + sizeof(theta_cand) * kXCap);
            if (ea != cudaSuccess) { rc = cuda_fail(ea, "record block", __FILE__, __LINE__); break; }
        }
        records_filter_kernel<<<1, 1024, kFilterSmem, st>>>(pl->d_emit, pl->emit_cap, pl->d_counters, global_min_dev, prior_min, margin,
                                                             pl->d_xblk, kXCap, pl->d_fscratch, filter_smem_max());
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(pl->h_xblk, pl->d_xblk, kHead + sizeof(theta_cand) * kSpecRecs, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "pass2", __FILE__, __LINE__); break; }
        std::memcpy(pl->h_ctr, pl->h_xblk, kHead);
        pl->p1_pending = false;
        read_pass1_stats(pl);
        cudaEventElapsedTime(&pl->stats.ms_pass2, pl->ev2, pl->ev3);
        const unsigned long long* ctr = pl->h_ctr;
        std::memcpy(&g, &ctr[kStGlobalMin], sizeof(double));
        pl->stats.n_rescanned = ctr[kStList];
        const unsigned long long kept = ctr[kCounters + 3];
        if (kept <= kXCap && !env_flag("THETA_FILTER_HOST")) {
            if (kept > kSpecRecs) {
                e = cudaMemcpyAsync(pl->h_xblk + kHead + sizeof(theta_cand) * kSpecRecs, pl->d_xblk + kHead + sizeof(theta_cand) * kSpecRecs,
                                    sizeof(theta_cand) * (kept - kSpecRecs), cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) { rc = cuda_fail(e, "pass2 D2H", __FILE__, __LINE__); break; }
            }
            host = reinterpret_cast<theta_cand*>(pl->h_xblk + kHead);
            n_host = (size_t)kept;
            filtered = true;
            break;
        }
        // more records than the block holds, or too many for the device-side sort: the raw records come back and the
        // host sorts and filters them (below)
        e = cudaMemcpy(pl->h_ctr, pl->d_counters, sizeof(unsigned long long) * kStatusWords, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { rc = cuda_fail(e, "pass2 status", __FILE__, __LINE__); break; }
        const unsigned long long emitted = ctr[kCounters + 3];
        if (emitted > pl->emit_cap) { want = (size_t)emitted + 1024; continue; }
        if ((size_t)emitted > pl->h_emit_cap) {
            theta::hfree(pl->h_emit); pl->h_emit = nullptr; pl->h_emit_cap = 0;
            e = theta::hmalloc(&pl->h_emit, sizeof(theta_cand) * (size_t)emitted);
            if (e != cudaSuccess) { rc = cuda_fail(e, "emit staging", __FILE__, __LINE__); break; }
            pl->h_emit_cap = (size_t)emitted;
        }
        e = cudaMemcpyAsync(pl->h_emit, pl->d_emit, sizeof(theta_cand) * emitted, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "pass2 D2H", __FILE__, __LINE__); break; }
        host = pl->h_emit;
        n_host = (size_t)emitted;
        break;
    }
    if (rc) return rc;
    if (global_min_host) *global_min_host = g;
    const double near_thr = g + margin;
    if (!filtered) {
        std::sort(host, host + n_host, [](const theta_cand& a, const theta_cand& b) { return a.rank < b.rank; });
        // keep exactly: strict prefix minima of the whole stream (w.r.t. everything before this shard),
        // the near-minimum set, and NaN values -- the per-lane filter on the device emits a superset
        size_t w = 0;
        double run = prior_min;
        for (size_t i = 0; i < n_host; ++i) {
            const double v = host[i].nll;
            const bool keep = (v != v) || v < run || v <= near_thr;
            if (v < run) run = v;
            if (keep) host[w++] = host[i];
        }
        n_host = w;
    }
    *n_out = n_host;
    if (n_host > cap) {
        set_error("pass2 produced %zu records but the caller's buffer holds %llu", n_host, (unsigned long long)cap);
        return THETA_ERR_OVERFLOW;
    }
    if (out_host && n_host) std::memcpy(out_host, host, sizeof(theta_cand) * n_host);
    return THETA_OK;
}

// Fully asynchronous pass 2 for the multi-GPU exchange: list scan, emit kernel and records_filter_kernel are enqueued;
// the caller's device buffer receives  [ u64 status[32] | theta_cand records[cap] ]  -- the operand of the all-gather:
// this shard's records sorted by rank and filtered exactly as theta_filter_records would (prior_min, global minimum + 2.5
// margins), status[8+3] = their number.  Nothing is waited for.  The caller fetches the gathered buffers, hands its own
// status words to theta_search_n2_async_done and the concatenated records to theta_filter_records (a linear pass when
// every block arrives sorted).  status[8+3] > cap means this shard's records did not fit the exchange block (or were too
// many to be filtered on the device and are handed over raw once the block is large enough).
extern "C" int theta_search_n2_pass2_enqueue(theta_plan* pl, const double* global_min_dev, double prior_min,
                                             void* out_dev, uint64_t cap, void* stream) {
    if (!pl || !out_dev || cap == 0) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (pl->n != 2 || !pl->ev0) { set_error("pass2 without pass1"); return THETA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t stripes = pl->p1_stripes;
    const double margin = 2.5 * THETA_TIE_MARGIN;
    const size_t want = std::max<size_t>(std::max<size_t>(kFilterMaxGlobal, pl->emit_cap), cap);
    if (want > pl->emit_cap) {
        theta::dfree(pl->d_emit); pl->d_emit = nullptr; pl->emit_cap = 0;
        THETA_CUDA(theta::dmalloc(&pl->d_emit, sizeof(theta_cand) * want));
        pl->emit_cap = want;
    }
    THETA_CUDA(record_event(pl->ev2, st));
    stripe_scan_kernel<<<1, 1024, 0, st>>>(pl->d_stripe_min, stripes, global_min_dev, prior_min, margin, pl->d_counters,
                                           pl->d_list, pl->d_prior);
    int rc = THETA_OK;
    if (stripes && pl->p1_begin != pl->p1_end) {
        SearchParams sp{};
        sp.tables = pl->d_tables; sp.layout = pl->layout; sp.kc = pl->kc; sp.logtab = pl->d_logtab;
        sp.begin = pl->p1_begin; sp.end = pl->p1_end; sp.L = pl->p1_L; sp.n_items = (uint32_t)stripes;
        sp.n_items_dev = pl->d_counters + kStList;
        sp.Lrun = std::min(pl->p1_L, pass2_run_length()); sp.units_per_item = (pl->p1_L + sp.Lrun - 1) / sp.Lrun;
        sp.stripe_min = pl->d_stripe_min; sp.counters = pl->d_counters + kCounters;
        sp.stripe_list = pl->d_list; sp.stripe_prior = pl->d_prior;
        sp.near_thr = 0.0; sp.near_thr_base = pl->d_counters + kStGlobalMin; sp.near_margin = margin;
        sp.emit = pl->d_emit; sp.emit_cap = pl->emit_cap;
        switch (pl->KT) {
#define THETA_P2(KTV) case KTV: rc = launch_search<KTV, true>(pl, sp, st); break;
            THETA_P2(2) THETA_P2(3) THETA_P2(4) THETA_P2(5) THETA_P2(6) THETA_P2(7) THETA_P2(8) THETA_P2(10) THETA_P2(12) THETA_P2(15)
#undef THETA_P2
            default: set_error("unsupported level count"); rc = THETA_ERR_UNSUPPORTED;
        }
        if (rc) return rc;
    }
    THETA_CUDA(record_event(pl->ev3, st));
    // status block + this shard's records, sorted by rank and filtered to what the fold replay needs, into the exchange block
    records_filter_kernel<<<1, 1024, kFilterSmem, st>>>(pl->d_emit, pl->emit_cap, pl->d_counters, global_min_dev, prior_min, margin,
                                                         static_cast<unsigned char*>(out_dev), cap, pl->d_fscratch, filter_smem_max());
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

// After the caller has synchronised: record the statistics of the asynchronous search from this shard's status words.
extern "C" int theta_search_n2_async_done(theta_plan* pl, const unsigned long long* status_host) {
    if (!pl || !status_host) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (!pl->h_ctr) { set_error("no search in flight"); return THETA_ERR_INVALID; }
    std::memcpy(pl->h_ctr, status_host, sizeof(unsigned long long) * kStatusWords);
    pl->p1_pending = false;
    read_pass1_stats(pl);
    cudaEventElapsedTime(&pl->stats.ms_pass2, pl->ev2, pl->ev3);
    pl->stats.n_rescanned = status_host[kStList];
    return THETA_OK;
}

// Sort raw pass-2 records by rank and keep exactly what the fold replay needs: NaN values, strict prefix minima
// of the sequence (starting from prior_min) and everything within near_thr.  In place; returns the new count.
extern "C" int theta_filter_records(theta_cand* recs, uint64_t n, double prior_min, double near_thr, uint64_t* n_out) {
    if ((!recs && n) || !n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    // (blocks that went through records_filter_kernel arrive sorted: the concatenation of the shards' blocks then is, too)
    const auto by_rank = [](const theta_cand& a, const theta_cand& b) { return a.rank < b.rank; };
    if (!std::is_sorted(recs, recs + n, by_rank)) std::sort(recs, recs + n, by_rank);
    size_t w = 0;
    double run = prior_min;
    for (size_t i = 0; i < n; ++i) {
        const double v = recs[i].nll;
        const bool keep = (v != v) || v < run || v <= near_thr;
        if (v < run) run = v;
        if (keep) recs[w++] = recs[i];
    }
    *n_out = w;
    return THETA_OK;
}

// Host-value form: global_min is written to the status block and picked up by the device-side list builder.
extern "C" int theta_search_n2_pass2(theta_plan* pl, double global_min, double prior_min, theta_cand* out_host,
                                     uint64_t cap, uint64_t* n_out, void* stream) {
    if (!pl || !n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_out = 0;
    if (pl->n != 2 || !pl->ev0) { set_error("pass2 without pass1"); return THETA_ERR_INVALID; }
    if (pl->p1_pending) { set_error("pass2 before pass1_end"); return THETA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    double* slot = reinterpret_cast<double*>(pl->d_counters + kStatusWords - 1);  // spare word of the status block
    double* stage = reinterpret_cast<double*>(pl->h_ctr + kStatusWords - 1);
    *stage = global_min;
    THETA_CUDA(cudaMemcpyAsync(slot, stage, sizeof(double), cudaMemcpyHostToDevice, st));
    return theta_search_n2_pass2_dev(pl, slot, prior_min, out_host, cap, n_out, nullptr, stream);
}

extern "C" int theta_search_stats_get(const theta_plan* pl, theta_search_stats* s) {
    if (!pl || !s) { set_error("null argument"); return THETA_ERR_INVALID; }
    *s = pl->stats;
    return THETA_OK;
}

extern "C" int theta_search_n2(theta_plan* pl, uint64_t rank_begin, uint64_t rank_end, theta_cand* out_host,
                               uint64_t cap, uint64_t* n_out) {
    double g = 0;
    int rc = theta_search_n2_pass1(pl, rank_begin, rank_end, &g, nullptr);
    if (rc) return rc;
    return theta_search_n2_pass2(pl, g, INFINITY, out_host, cap, n_out, nullptr);
}

// Host mirror of n2_unrank (theta_n2.cuh): tumour column of `rank` from the plan's count tables.
static void host_column_of_rank(const theta_plan* pl, uint64_t rank, int8_t* col) {
    const TableLayout& L = pl->layout;
    const int m = pl->m, KT = pl->KT;
    const uint64_t* SW = reinterpret_cast<const uint64_t*>(pl->h_tables.data() + L.off_SW);
    const int32_t* qlo = reinterpret_cast<const int32_t*>(pl->h_tables.data() + L.off_qlo);
    const int32_t* qhi = reinterpret_cast<const int32_t*>(pl->h_tables.data() + L.off_qhi);
    std::memset(col, 0, (size_t)m);
    uint64_t rem = rank;
    int prev = 0;
    for (int v = KT; v >= 1; --v) {
        const uint64_t* sw = SW + (size_t)(v - 1) * (m + 2);
        int lo = std::max(prev, (int)qlo[v - 1]), hi = qhi[v - 1];
        const uint64_t base = sw[lo];
        int l = lo;
        while (l < hi) {
            const int mid = (l + hi) >> 1;
            if (base - sw[mid + 1] > rem) hi = mid; else l = mid + 1;
        }
        rem -= base - sw[l];
        prev = l;
        for (int i = m - l; i < m; ++i) col[i] += 1;  // q_v = l intervals (the last l in sorted order) have c_i >= v
    }
}

// tumour columns [count][m] of given ranks (host-side unranking from the plan's count tables; no kernel launch)
extern "C" int theta_columns_of_ranks(const theta_plan* pl, const uint64_t* ranks, int64_t count, int8_t* cols) {
    if (!pl || (count > 0 && (!ranks || !cols))) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (pl->n != 2) { set_error("plan is not n=2"); return THETA_ERR_INVALID; }
    for (int64_t i = 0; i < count; ++i) {
        if (ranks[i] >= pl->count) { set_error("rank %llu out of range", (unsigned long long)ranks[i]); return THETA_ERR_INVALID; }
        host_column_of_rank(pl, ranks[i], cols + i * (size_t)pl->m);
    }
    return THETA_OK;
}

// do_optimization_single (python/RunTHetA.py:173-220) for n=2 in ONE call: plan, both passes, the reference's
// running-min fold (RunTHetA.py:198-206 with Misc.isClose, the first matrix evaluated twice -- Q1, NaN counts as
// close -- Q6) over the emitted records, and for each survivor its tumour column and Optimizer.solve's `vals`.
// The winners come back in fold order (a rank can appear twice).  *n_best may exceed cap_best (THETA_ERR_OVERFLOW).
extern "C" int theta_optimize_n2(const theta_problem* prob, uint64_t rank_limit, int32_t* lower_out, int32_t* upper_out,
                                 uint64_t cap_best, uint64_t* n_best, theta_cand* best, int8_t* cols, double* vals,
                                 theta_search_stats* stats, uint64_t* n_candidates) {
    if (!prob || !n_best || (cap_best && (!best || !cols || !vals))) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_best = 0;
    if (prob->n != 2) { set_error("theta_optimize_n2 needs n == 2"); return THETA_ERR_INVALID; }
    theta_plan* pl = nullptr;
    int rc = theta_plan_create(prob, &pl);
    if (rc) return rc;
    const int m = pl->m;
    if (lower_out) std::memcpy(lower_out, pl->lower.data(), sizeof(int32_t) * m);
    if (upper_out) std::memcpy(upper_out, pl->upper.data(), sizeof(int32_t) * m);
    const uint64_t total = rank_limit ? std::min<uint64_t>(pl->count, rank_limit) : pl->count;
    if (n_candidates) *n_candidates = total;
    std::vector<theta_cand> recs(1024);
    uint64_t n = 0;
    rc = theta_search_n2_pass1_begin(pl, 0, total, nullptr, nullptr);
    if (rc == THETA_OK) {
        rc = theta_search_n2_pass2_dev(pl, nullptr, INFINITY, recs.data(), recs.size(), &n, nullptr, nullptr);
        if (rc == THETA_ERR_OVERFLOW && n > recs.size()) {
            recs.resize(n + 16);
            rc = theta_search_n2_pass2_dev(pl, nullptr, INFINITY, recs.data(), recs.size(), &n, nullptr, nullptr);
        }
    }
    if (rc) { theta_plan_destroy(pl); return rc; }
    if (stats) *stats = pl->stats;
    // the fold, verbatim
    std::vector<uint32_t> keep;
    double M = INFINITY;
    for (uint64_t i = 0; i < n; ++i) {
        const int reps = recs[i].rank == 0 ? 2 : 1;
        for (int k = 0; k < reps; ++k) {
            const double L = recs[i].nll;
            if (!(fabs(L - M) > THETA_TIE_MARGIN)) keep.push_back((uint32_t)i);
            else if (L < M) { keep.assign(1, (uint32_t)i); M = L; }
        }
    }
    *n_best = keep.size();
    if (keep.size() > cap_best) {
        set_error("%zu solutions tie, the caller's buffers hold %llu", keep.size(), (unsigned long long)cap_best);
        theta_plan_destroy(pl);
        return THETA_ERR_OVERFLOW;
    }
    // columns by host-side unranking, vals_i = numer_i / denom of Optimizer.L2 (Optimizer.py:187-196) at the
    // record's mu: no further kernel launch after pass 2
    for (size_t k = 0; k < keep.size(); ++k) {
        best[k] = recs[keep[k]];
        int8_t* col = cols + k * (size_t)m;
        double* vk = vals + k * (size_t)m;
        if (k && best[k].rank == best[k - 1].rank) {
            std::memcpy(col, col - m, (size_t)m);
            std::memcpy(vk, vk - m, sizeof(double) * (size_t)m);
            continue;
        }
        host_column_of_rank(pl, best[k].rank, col);
        const double mu = best[k].mu[0], mu1 = 1.0 - mu;
        double S1 = 0.0;
        for (int i = 0; i < m; ++i) S1 += (double)pl->rN[i] * (double)col[i];  // exact (< 2^53)
        const double D = mu * pl->kc.S0 + mu1 * S1;
        for (int i = 0; i < m; ++i) vk[i] = (double)pl->rN[i] * (pl->kc.tau * mu + (double)col[i] * mu1) / D;
    }
    theta_plan_destroy(pl);
    return rc;
}

extern "C" int theta_eval_ranks_n2(theta_plan* pl, const uint64_t* ranks_host, int64_t count, theta_cand* out_host,
                                   double* vals_host) {
    if (!pl || !ranks_host || !out_host) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (pl->n != 2) { set_error("plan is not n=2"); return THETA_ERR_INVALID; }
    if (count <= 0) return THETA_OK;
    for (int64_t i = 0; i < count; ++i)
        if (ranks_host[i] >= pl->count) { set_error("rank %llu out of range", (unsigned long long)ranks_host[i]); return THETA_ERR_INVALID; }
    const int m = pl->m;
    theta::DevBlock b_ranks, b_out, b_vals;
    THETA_CUDA(b_ranks.alloc(sizeof(uint64_t) * count));
    THETA_CUDA(b_out.alloc(sizeof(theta_cand) * count));
    if (vals_host) THETA_CUDA(b_vals.alloc(sizeof(double) * count * m));
    uint64_t* d_ranks = b_ranks.as<uint64_t>(); theta_cand* d_out = b_out.as<theta_cand>(); double* d_vals = b_vals.as<double>();
    THETA_CUDA(cudaMemcpy(d_ranks, ranks_host, sizeof(uint64_t) * count, cudaMemcpyHostToDevice));
    const int grid = (int)((count + kBlock - 1) / kBlock);
    THETA_KT_DISPATCH(pl->KT, {
        if (pl->layout.bytes > 48 * 1024)
            THETA_CUDA(cudaFuncSetAttribute(n2_eval_ranks_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->layout.bytes));
        n2_eval_ranks_kernel<KT><<<grid, kBlock, pl->layout.bytes>>>(pl->d_tables, pl->layout, pl->kc, d_ranks, count, pl->d_rN, d_out, d_vals);
    });
    THETA_CUDA(cudaGetLastError());
    THETA_CUDA(cudaMemcpy(out_host, d_out, sizeof(theta_cand) * count, cudaMemcpyDeviceToHost));
    if (vals_host) THETA_CUDA(cudaMemcpy(vals_host, d_vals, sizeof(double) * count * m, cudaMemcpyDeviceToHost));
    return THETA_OK;
}
