// theta_tiles.cu -- streaming kernels over explicit candidate tiles (n=2).
//
// A tile is a block of candidates stored as int8 tumour columns, one row of `pitch`
// bytes per candidate (pitch = theta_tile_pitch(m) = smallest 8*odd >= m, padding
// bytes zero), i.e. the compact form of the m x 2 float64 matrices that
// Enumerator._C_to_array (python/Enumerator.py:154-160) hands to Optimizer.solve.
//
//   nll_tiles_n2   : Optimizer.L2 (python/Optimizer.py:187-196) at a given mu -- the
//                    streaming "NLL kernel": per candidate it reads m bytes of C and
//                    8 bytes of mu and writes 8 bytes of NLL.
//   solve_tiles_n2 : Optimizer._solve_n2 (python/Optimizer.py:90-126) on explicit C.
//
// Data movement: persistent CTAs; each tile (kRows rows, contiguous in HBM) is brought
// into shared memory by one TMA bulk copy (cp.async.bulk + mbarrier complete_tx) through
// a kStages-deep ring, so HBM reads are full 128 B lines issued by the copy engine while
// the SM computes on earlier tiles.  One thread owns one row; pitch/8 is odd so the
// per-thread 8-byte row reads are bank-conflict free.
//
// Arithmetic: a row is scanned 8 bytes at a time for *change points* (c_i != c_{i-1});
// each run [s, e) of equal copy number v contributes  (PR[e]-PR[s]) * log(tau mu + v (1-mu))
// to the log-sum and v * (PRN[e]-PRN[s]) to S1, with PR / PRN the exact prefix sums of r / rN.
// THetA's candidates are monotone, so a row has at most K+1 runs and the per-byte work is
// a handful of integer instructions; arbitrary rows are handled by the same code at a
// cost proportional to their number of runs.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "theta_n2.cuh"

namespace theta {

static constexpr int kRows = 128;   // rows per tile == threads per CTA
static constexpr int kStages = 1;   // one tile in flight per CTA; 8 CTAs per SM overlap copy and compute
static constexpr int kMaxRuns = 12;  // run list per row kept in shared memory; longer rows take the direct path

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct TileParams {
    uint8_t ra[64], rb[64];  // n = 3 only: copy numbers (a, b) of row index g (Enumerator._create_graph order)
    int n_rows;              // n = 3 only: number of valid row indices
    const double2* logtab;
    const int8_t* cols;
    int64_t pitch;
    int64_t count;
    const double* PR;   // [pitch+1] prefix sums of r  (zero-extended past m)
    const double* PRN;  // [pitch+1] prefix sums of rN
    const int32_t* ZP;  // [pitch+1] prefix counts of r_i == 0
    N2Const kc;
    const double* mu_in;
    double* mu_out;
    double* nll_out;
    int32_t* status_out;
};

// Shared-memory carve-up (dynamic): ring of tiles | PR | PRN | ZP | [histogram] ; barriers are static.
struct TileSmem {
    unsigned char* ring;
    int32_t* ZP;
    double2* PRP;    // [pitch+1] (PR, PRN) interleaved: one 16-byte read per run end
    double2* LT;     // fast_log table (4 KB)
    uint16_t* runs;  // [kMaxRuns][kRows]: (value << 10) | end
    double* hist;
};

__host__ __device__ inline size_t tile_smem_bytes(int64_t pitch, int hist_levels) {
    size_t o = (size_t)kStages * kRows * pitch;
    o = (o + 15) & ~(size_t)15;
    o += sizeof(int32_t) * (pitch + 1);
    o = (o + 15) & ~(size_t)15;
    o += sizeof(double2) * (pitch + 1);
    o += sizeof(double2) * THETA_LOGTAB_SIZE;
    o += sizeof(uint16_t) * kMaxRuns * kRows;
    o += sizeof(double) * (size_t)hist_levels * kRows;
    return o;
}

__device__ __forceinline__ TileSmem tile_smem_bind(unsigned char* base, int64_t pitch) {
    TileSmem s;
    size_t o = 0;
    s.ring = base;
    o = (size_t)kStages * kRows * pitch;
    o = (o + 15) & ~(size_t)15;
    s.ZP = reinterpret_cast<int32_t*>(base + o);
    o += sizeof(int32_t) * (pitch + 1);
    o = (o + 15) & ~(size_t)15;
    s.PRP = reinterpret_cast<double2*>(base + o);
    o += sizeof(double2) * (pitch + 1);
    s.LT = reinterpret_cast<double2*>(base + o);
    o += sizeof(double2) * THETA_LOGTAB_SIZE;
    s.runs = reinterpret_cast<uint16_t*>(base + o);
    o += sizeof(uint16_t) * kMaxRuns * kRows;
    s.hist = reinterpret_cast<double*>(base + o);
    return s;
}

// Issue the bulk copy of tile `t` into ring slot `slot` (one thread).  The 16-byte-multiple
// body goes through TMA; a possible 8-byte tail of the last tile is copied by hand.
__device__ __forceinline__ void issue_tile(const TileParams& p, unsigned char* ring, uint64_t* bars, int64_t t, int slot) {
    const int64_t row0 = t * kRows;
    const int64_t rows = min((int64_t)kRows, p.count - row0);
    const uint32_t bytes = (uint32_t)(rows * p.pitch);
    const uint32_t body = bytes & ~15u;
    unsigned char* dst = ring + (size_t)slot * kRows * p.pitch;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.cols) + row0 * p.pitch;
    if (body < bytes) {
        *reinterpret_cast<uint64_t*>(dst + body) = *reinterpret_cast<const uint64_t*>(src + body);
    }
    if (body) {
        mbar_expect_tx(&bars[slot], body);
        tma_load_1d(dst, src, body, &bars[slot]);
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[slot])) : "memory");
    }
}

// Scan one row for runs of equal copy number; F(v, start, end) is called once per run.
template <typename F>
__device__ __forceinline__ void scan_runs(const unsigned char* row, int words, F&& on_run) {
    const uint64_t* rw = reinterpret_cast<const uint64_t*>(row);
    uint64_t w = rw[0];
    uint32_t cur = (uint32_t)(w & 0xff);
    uint32_t prev_last = cur;
    int start = 0;
    for (int wi = 0; wi < words; ++wi) {
        if (wi) w = rw[wi];
        const uint64_t sh = (w << 8) | prev_last;
        uint64_t x = w ^ sh;
        prev_last = (uint32_t)(w >> 56);
        while (x) {
            const int b = (__ffsll((long long)x) - 1) >> 3;
            const int pos = wi * 8 + b;
            on_run(cur, start, pos);
            start = pos;
            cur = (uint32_t)(w >> (8 * b)) & 0xff;
            x &= ~(0xffull << (8 * b));
        }
    }
    on_run(cur, start, words * 8);
}

// Fast scan: collect the runs of a row into the per-thread run list (shared memory, column `tid`).
// A word whose 8 bytes all equal the current value costs five integer instructions.  Returns the
// number of runs, or -1 if the row has more than kMaxRuns runs / a value above 63 (caller falls back
// to scan_runs).  Positions >= m (zero padding) never open a run.
__device__ __forceinline__ uint32_t nz_bytes4(uint32_t d) {
    // bit b set <=> byte b of d is non-zero (bit 7 of each byte, gathered by a carry-free multiply)
    const uint32_t t = ((((d | 0x80808080u) - 0x01010101u) | d) & 0x80808080u) >> 7;
    return (t * 0x00204081u) >> 21 & 0xfu;
}

__device__ __forceinline__ int collect_runs(const unsigned char* row, int words, int m, uint16_t* runs) {
    const uint2* rw = reinterpret_cast<const uint2*>(row);
    uint32_t cur = rw[0].x & 0xffu;  // value of the open run == last byte seen
    uint32_t curpat = cur * 0x01010101u;
    int n = 0;
    bool over = false;
    for (int wi = 0; wi < words; ++wi) {
        const uint2 w = rw[wi];
        if (((w.x ^ curpat) | (w.y ^ curpat)) != 0u) {
            // change points: byte b differs from its predecessor
            const uint32_t dlo = w.x ^ ((w.x << 8) | cur);
            const uint32_t dhi = w.y ^ __funnelshift_l(w.x, w.y, 8);
            uint32_t cm = nz_bytes4(dlo) | (nz_bytes4(dhi) << 4);
            while (cm) {
                const int b = __ffs((int)cm) - 1;
                const int pos = wi * 8 + b;
                if (pos >= m) break;  // zero padding
                if (n < kMaxRuns - 1 && cur < 64u) runs[n * kRows] = (uint16_t)((cur << 10) | (uint32_t)pos);
                else over = true;
                ++n;
                cur = __byte_perm(w.x, w.y, (uint32_t)b) & 0xffu;
                cm &= cm - 1;
            }
            if (cm == 0u) cur = w.y >> 24;  // (when we stopped inside the padding the open run keeps its value)
            curpat = (cm == 0u) ? cur * 0x01010101u : 0u;  // padding words are all zero
        }
    }
    if (over || cur >= 64u) return -1;
    runs[n * kRows] = (uint16_t)((cur << 10) | (uint32_t)m);
    return n + 1;
}

template <bool SOLVE, int KT>
__global__ void __launch_bounds__(kRows) tiles_n2_kernel(const TileParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[kStages];
    TileSmem S = tile_smem_bind(smem, p.pitch);
    const int tid = threadIdx.x;
    const int words = (int)(p.pitch >> 3);
    for (int i = tid; i <= p.pitch; i += kRows) {
        S.ZP[i] = p.ZP[i];
        S.PRP[i] = make_double2(p.PR[i], p.PRN[i]);
    }
    for (int i = tid; i < THETA_LOGTAB_SIZE; i += kRows) S.LT[i] = p.logtab[i];
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_tiles = (p.count + kRows - 1) / kRows;
    const N2Const kc = p.kc;
    // prologue: fill the ring
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < kStages - 1; ++s) {
            const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (t < n_tiles) issue_tile(p, S.ring, bars, t, s);
        }
    }
    __syncthreads();  // makes a hand-copied tail of the prologue tiles visible to every thread
    int it = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int slot = it % kStages;
        const uint32_t parity = (uint32_t)(it / kStages) & 1u;
        // refill the slot consumed in the previous iteration (every thread has passed the barrier below)
        if (tid == 0) {
            const int64_t tn = t + (int64_t)(kStages - 1) * gridDim.x;
            if (tn < n_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_tile(p, S.ring, bars, tn, (it + kStages - 1) % kStages);
            }
        }
        mbar_wait(&bars[slot], parity);
        const int64_t row = t * kRows + tid;
        if (row < p.count) {
            const unsigned char* rp = S.ring + ((size_t)slot * kRows + tid) * p.pitch;
            uint16_t* myruns = S.runs + tid;
            const int nr = collect_runs(rp, words, kc.m, myruns);
            {
                double* h = S.hist + tid;  // hist[v * kRows + tid]
#pragma unroll
                for (int v = 0; v <= KT; ++v) h[v * kRows] = 0.0;
                double s1 = 0.0;
                bool z0 = false, bad = false;
                auto on_run = [&](uint32_t v, int s, int e) {
                    const double wgt = S.PRP[e].x - S.PRP[s].x;
                    if (v > (uint32_t)KT) { if (S.PRP[e].y - S.PRP[s].y != 0.0 || wgt != 0.0) bad = true; return; }
                    s1 = fma((double)v, S.PRP[e].y - S.PRP[s].y, s1);
                    h[v * kRows] += wgt;
                    if (v == 0 && S.ZP[e] - S.ZP[s] > 0) z0 = true;
                };
                if (nr >= 0) {
                    int st = 0;
                    for (int j = 0; j < nr; ++j) {
                        const uint32_t code = myruns[j * kRows];
                        const int e = (int)(code & 1023u);
                        on_run(code >> 10, st, e);
                        st = e;
                    }
                } else {
                    scan_runs(rp, words, on_run);
                }
                double R[KT + 1];
#pragma unroll
                for (int v = 0; v <= KT; ++v) R[v] = h[v * kRows];
                SolveOut o;
                n2_solve<KT + 1, true>(R, s1, z0, kc, -1.0, o);
                if (bad) { o.status = THETA_ST_NONE_NAN; o.mu = 0.0; o.nll = 0.0; }
                p.mu_out[row] = o.mu;
                p.nll_out[row] = o.nll;
                p.status_out[row] = o.status;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming NLL kernel (Optimizer.L2, python/Optimizer.py:187-196, at a given mu per candidate).
//
// Warp-granular pipeline: every warp owns a ring of STAGES sub-tiles (32 rows x pitch bytes,
// contiguous in HBM) filled by TMA bulk copies on the warp's own mbarriers -- no block-level
// synchronisation after the tables are staged.  Lane l evaluates row l of the sub-tile in two
// phases so that the warp stays converged:
//   phase 1 (branch-free, 6 instructions per 8-byte word): mark the words that are not a
//            repetition of the last byte seen (a word inside a run never is marked);
//   phase 2: for each marked word, walk its change points; a run [s, e) of copy number v adds
//            (PR[e]-PR[s]) * log(tau mu + v (1-mu)) and v * (PRN[e]-PRN[s]) from the exact prefix
//            sums in shared memory.  Positions past m carry zero weight (the prefix tables are
//            flat there), so the zero padding needs no special case.
// Any byte values and any number of runs are handled; cost is proportional to the number of runs.
template <int WARPS, int STAGES>
struct NllCfg {
    static constexpr int kThreads = WARPS * 32;
    __host__ __device__ static size_t smem_bytes(int64_t pitch) {
        size_t o = sizeof(double2) * (size_t)(pitch + 1);
        o = (o + 127) & ~(size_t)127;
        o += sizeof(double2) * THETA_LOGTAB_SIZE;
        o += sizeof(double2) * 64;  // (a, b) per row index (n = 3)
        o += (size_t)WARPS * STAGES * 32 * (size_t)pitch;
        return o + 32;  // slack: phase 1 of the last row may read up to 24 bytes past its end
    }
};

__device__ __forceinline__ void issue_subtile(const TileParams& p, unsigned char* dst, uint64_t* bar, int64_t t) {
    const int64_t row0 = t * 32;
    const int64_t rows = min((int64_t)32, p.count - row0);
    const uint32_t bytes = (uint32_t)(rows * p.pitch);
    const uint32_t body = bytes & ~15u;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.cols) + row0 * p.pitch;
    if (body < bytes) *reinterpret_cast<uint64_t*>(dst + body) = *reinterpret_cast<const uint64_t*>(src + body);
    if (body) {
        mbar_expect_tx(bar, body);
        tma_load_1d(dst, src, body, bar);
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
}

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// prp_s: shared-space address of the (PR, PRN) table (explicit ld.shared keeps the address arithmetic to one LEA)
// N3 = false: byte = tumour copy number c_i, log argument tau mu + c (1 - mu)                (Optimizer.L2 / CalcAllC.L2)
// N3 = true : byte = row index g_i -> (a, b), log argument tau mu0 + a mu1 + b mu2 with mu = (mu, mu1, mu2)
//                                                                                (Optimizer.L3 / CalcAllC.L3)
template <bool N3>
__device__ __forceinline__ double nll_row(const unsigned char* rp, int words, uint32_t prp_s, uint32_t ab_s,
                                          const double2* __restrict__ LT, double mu, double mu1, double mu2,
                                          const N2Const& kc, double inv_A) {
    const uint2* rw = reinterpret_cast<const uint2*>(rp);
    const double tm = kc.tau * mu;
    double acc = 0.0, s1 = 0.0, s2 = 0.0, pr_s = 0.0, prn_s = 0.0;
    uint32_t cur = rw[0].x & 0xffu;     // copy number of the open run
    uint32_t pat = cur * 0x01010101u;   // phase 1: last byte seen, replicated
    auto close_run = [&](uint32_t v, int e) {
        const double2 pp = lds_f64x2(prp_s + (uint32_t)e * 16u);
        const double wgt = pp.x - pr_s;
        double g;
        if (N3) {
            const double2 ab = lds_f64x2(ab_s + (v & 63u) * 16u);
            const double dn = pp.y - prn_s;
            s1 = fma(ab.x, dn, s1);
            s2 = fma(ab.y, dn, s2);
            g = fma(ab.x, mu1, fma(ab.y, mu2, tm));
        } else {
            const double dv = (double)v;
            s1 = fma(dv, pp.y - prn_s, s1);
            g = fma(dv, mu1, tm);
        }
        if (wgt != 0.0) acc = fma(wgt, fast_log(g, LT), acc);
        pr_s = pp.x;
        prn_s = pp.y;
    };
    for (int w0 = 0; w0 < words; w0 += 32) {
        const int nw = min(32, words - w0);
        // blocks of four words with compile-time bit positions; a block may run up to three words past the
        // row (into the next row / the slack behind the ring): those bits are masked off below
        uint32_t mask = 0;
        const int nblk = (nw + 3) >> 2;
        for (int jb = 0; jb < nblk; ++jb) {
            uint32_t m4 = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint2 w = rw[w0 + jb * 4 + u];
                uint32_t diff;  // (w.x ^ pat) | (w.y ^ pat) as one LOP3 (the compiler otherwise emits two compares)
                asm("lop3.b32 %0, %1, %2, %3, 0x7e;" : "=r"(diff) : "r"(w.x), "r"(w.y), "r"(pat));
                if (diff != 0u) m4 |= 1u << u;
                pat = __byte_perm(w.y, 0u, 0x3333u);
            }
            mask |= m4 << (jb * 4);
        }
        if (nw < 32) mask &= (1u << nw) - 1u;
        uint32_t dlo = 0u, dhi = 0u;  // bytes of the current word that differ from their predecessor
        uint2 w = make_uint2(0u, 0u);
        int base = 0;
        for (;;) {
            if ((dlo | dhi) == 0u) {
                if (mask == 0u) break;
                const int wi = w0 + __ffs((int)mask) - 1;
                mask &= mask - 1u;
                w = rw[wi];
                base = wi * 8;
                dlo = w.x ^ ((w.x << 8) | cur);
                dhi = w.y ^ __funnelshift_l(w.x, w.y, 8);  // non-zero pair: the word is not a repetition of `cur`
            }
            const bool in_lo = dlo != 0u;
            const uint32_t x = in_lo ? dlo : dhi;
            const int bit8 = (__ffs((int)x) - 1) & 24;  // first byte that differs from its predecessor
            close_run(cur, base + (bit8 >> 3) + (in_lo ? 0 : 4));
            cur = ((in_lo ? w.x : w.y) >> bit8) & 0xffu;
            const uint32_t keep = ~(0xffu << bit8);
            if (in_lo) dlo &= keep; else dhi &= keep;
        }
    }
    close_run(cur, kc.m);
    const double gbar = (N3 ? fma(mu2, s2, fma(mu1, s1, mu * kc.S0)) : (mu * kc.S0 + mu1 * s1)) * inv_A;
    double nll = kc.K0 - (acc - kc.Rtot * fast_log(gbar, LT));
    if (kc.any_zero_rN) nll = INFINITY;
    return nll;
}

template <int WARPS, int STAGES, int MINB, bool N3>
__global__ void __launch_bounds__(WARPS * 32, MINB) nll_tiles_kernel(const TileParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[WARPS * STAGES];
    double2* PRP = reinterpret_cast<double2*>(smem);
    size_t o = sizeof(double2) * (size_t)(p.pitch + 1);
    o = (o + 127) & ~(size_t)127;
    double2* LT = reinterpret_cast<double2*>(smem + o);
    o += sizeof(double2) * THETA_LOGTAB_SIZE;
    double2* AB = reinterpret_cast<double2*>(smem + o);
    o += sizeof(double2) * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t slot_bytes = (size_t)32 * p.pitch;
    unsigned char* ring = smem + o + (size_t)warp * STAGES * slot_bytes;
    uint64_t* bar = bars + warp * STAGES;
    for (int i = tid; i <= p.pitch; i += WARPS * 32) PRP[i] = make_double2(p.PR[i], p.PRN[i]);
    for (int i = tid; i < THETA_LOGTAB_SIZE; i += WARPS * 32) LT[i] = p.logtab[i];
    if (N3 && tid < 64)  // an invalid row index poisons the result instead of reading garbage
        AB[tid] = tid < p.n_rows ? make_double2((double)p.ra[tid], (double)p.rb[tid]) : make_double2(NAN, NAN);
    if (tid == 0) {
        for (int s = 0; s < WARPS * STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int words = (int)(p.pitch >> 3);
    const int64_t n_sub = (p.count + 31) >> 5;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t GW = (int64_t)gridDim.x * WARPS;
    const N2Const kc = p.kc;
    const double inv_A = 1.0 / kc.A;
    uint32_t prp_s = smem_u32(PRP), ab_s = smem_u32(AB);
    asm volatile("" : "+r"(prp_s), "+r"(ab_s));  // keep them in registers (the compiler otherwise re-derives them per run)
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < (STAGES > 1 ? STAGES - 1 : 1); ++s) {
            const int64_t t = gw + (int64_t)s * GW;
            if (t < n_sub) issue_subtile(p, ring + (size_t)s * slot_bytes, &bar[s], t);
        }
    }
    int it = 0;
    for (int64_t t = gw; t < n_sub; t += GW, ++it) {
        const int slot = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        const int64_t row = t * 32 + lane;
        const bool live = row < p.count;
        double mu = 0.5, mu1 = 0.5, mu2 = 0.0;
        if (live) {
            if (N3) { mu = __ldg(p.mu_in + 3 * row); mu1 = __ldg(p.mu_in + 3 * row + 1); mu2 = __ldg(p.mu_in + 3 * row + 2); }
            else { mu = __ldg(p.mu_in + row); mu1 = 1.0 - mu; }
        }
        if (STAGES > 1) {
            __syncwarp();  // every lane is done with the slot about to be refilled
            if (lane == 0) {
                const int64_t tn = t + (int64_t)(STAGES - 1) * GW;
                if (tn < n_sub) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_subtile(p, ring + (size_t)((it + STAGES - 1) % STAGES) * slot_bytes, &bar[(it + STAGES - 1) % STAGES], tn);
                }
            }
        }
        mbar_wait(&bar[slot], parity);
        __syncwarp();  // a hand-copied 8-byte tail (last sub-tile only) becomes visible to every lane
        double nll = 0.0;
        if (live) nll = nll_row<N3>(ring + (size_t)slot * slot_bytes + (size_t)lane * p.pitch, words, prp_s, ab_s, LT, mu, mu1, mu2, kc, inv_A);
        if (STAGES == 1) {
            __syncwarp();
            if (lane == 0) {
                const int64_t tn = t + GW;
                if (tn < n_sub) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_subtile(p, ring, &bar[0], tn);
                }
            }
        }
        if (live) p.nll_out[row] = nll;
    }
}

// ---------------------------------------------------------------------------------------------
// Optimizer.solve(C) on explicit tiles (python/Optimizer.py:68-126) with the same warp-granular TMA pipeline and
// two-phase run walk as the NLL kernel: every run adds its reads to a per-lane histogram R[v] in shared memory
// (v = copy number), S1 and the zero-read flag of the v = 0 group come from the prefix tables, then n2_solve.
template <typename F>
__device__ __forceinline__ void walk_runs(const unsigned char* rp, int words, int m, F&& on_run) {
    const uint2* rw = reinterpret_cast<const uint2*>(rp);
    uint32_t cur = rw[0].x & 0xffu;
    uint32_t pat = cur * 0x01010101u;
    for (int w0 = 0; w0 < words; w0 += 32) {
        const int nw = min(32, words - w0);
        uint32_t mask = 0;
        const int nblk = (nw + 3) >> 2;
        for (int jb = 0; jb < nblk; ++jb) {
            uint32_t m4 = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint2 w = rw[w0 + jb * 4 + u];
                uint32_t diff;
                asm("lop3.b32 %0, %1, %2, %3, 0x7e;" : "=r"(diff) : "r"(w.x), "r"(w.y), "r"(pat));
                if (diff != 0u) m4 |= 1u << u;
                pat = __byte_perm(w.y, 0u, 0x3333u);
            }
            mask |= m4 << (jb * 4);
        }
        if (nw < 32) mask &= (1u << nw) - 1u;
        uint32_t dlo = 0u, dhi = 0u;
        uint2 w = make_uint2(0u, 0u);
        int base = 0;
        for (;;) {
            if ((dlo | dhi) == 0u) {
                if (mask == 0u) break;
                const int wi = w0 + __ffs((int)mask) - 1;
                mask &= mask - 1u;
                w = rw[wi];
                base = wi * 8;
                dlo = w.x ^ ((w.x << 8) | cur);
                dhi = w.y ^ __funnelshift_l(w.x, w.y, 8);
            }
            const bool in_lo = dlo != 0u;
            const uint32_t x = in_lo ? dlo : dhi;
            const int bit8 = (__ffs((int)x) - 1) & 24;
            on_run(cur, base + (bit8 >> 3) + (in_lo ? 0 : 4));
            cur = ((in_lo ? w.x : w.y) >> bit8) & 0xffu;
            const uint32_t keep = ~(0xffu << bit8);
            if (in_lo) dlo &= keep; else dhi &= keep;
        }
    }
    on_run(cur, m);
}

template <int WARPS, int KT>
__host__ __device__ inline size_t solve_smem_bytes(int64_t pitch) {
    size_t o = sizeof(double2) * (size_t)(pitch + 1);   // (PR, PRN)
    o = (o + 15) & ~(size_t)15;
    o += sizeof(int32_t) * (size_t)(pitch + 1);         // ZP
    o = (o + 127) & ~(size_t)127;
    o += sizeof(double) * (size_t)WARPS * (KT + 1) * 32;  // histograms
    o += (size_t)WARPS * 32 * (size_t)pitch;              // one sub-tile per warp
    return o + 32;
}

template <int WARPS, int MINB, int KT>
__global__ void __launch_bounds__(WARPS * 32, MINB) solve_tiles_kernel(const TileParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[WARPS];
    double2* PRP = reinterpret_cast<double2*>(smem);
    size_t o = sizeof(double2) * (size_t)(p.pitch + 1);
    o = (o + 15) & ~(size_t)15;
    int32_t* ZP = reinterpret_cast<int32_t*>(smem + o);
    o += sizeof(int32_t) * (size_t)(p.pitch + 1);
    o = (o + 127) & ~(size_t)127;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* hist = reinterpret_cast<double*>(smem + o) + (size_t)warp * (KT + 1) * 32 + lane;  // hist[v * 32]
    o += sizeof(double) * (size_t)WARPS * (KT + 1) * 32;
    const size_t slot_bytes = (size_t)32 * p.pitch;
    unsigned char* ring = smem + o + (size_t)warp * slot_bytes;
    uint64_t* bar = bars + warp;
    for (int i = tid; i <= p.pitch; i += WARPS * 32) {
        PRP[i] = make_double2(p.PR[i], p.PRN[i]);
        ZP[i] = p.ZP[i];
    }
    if (tid == 0) {
        for (int s = 0; s < WARPS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int words = (int)(p.pitch >> 3);
    const int64_t n_sub = (p.count + 31) >> 5;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t GW = (int64_t)gridDim.x * WARPS;
    const N2Const kc = p.kc;
    if (lane == 0 && gw < n_sub) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_subtile(p, ring, bar, gw);
    }
    int it = 0;
    for (int64_t t = gw; t < n_sub; t += GW, ++it) {
        const int64_t row = t * 32 + lane;
        const bool live = row < p.count;
        mbar_wait(bar, (uint32_t)it & 1u);
        __syncwarp();
        double R[KT + 1];
        double s1 = 0.0;
        bool z0 = false, bad = false;
        if (live) {
#pragma unroll
            for (int v = 0; v <= KT; ++v) hist[v * 32] = 0.0;
            double pr_s = 0.0, prn_s = 0.0;
            int zp_s = 0;
            walk_runs(ring + (size_t)lane * p.pitch, words, kc.m, [&](uint32_t v, int e) {
                const double2 pp = PRP[e];
                const int zp = ZP[e];
                const double wgt = pp.x - pr_s, dn = pp.y - prn_s;
                if (v > (uint32_t)KT) {
                    if (dn != 0.0 || wgt != 0.0) bad = true;
                } else {
                    s1 = fma((double)v, dn, s1);
                    hist[v * 32] += wgt;
                    if (v == 0u && zp - zp_s > 0) z0 = true;
                }
                pr_s = pp.x; prn_s = pp.y; zp_s = zp;
            });
#pragma unroll
            for (int v = 0; v <= KT; ++v) R[v] = hist[v * 32];
        }
        __syncwarp();  // every lane is done with the sub-tile: request the next one, solve meanwhile
        if (lane == 0) {
            const int64_t tn = t + GW;
            if (tn < n_sub) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_subtile(p, ring, bar, tn);
            }
        }
        if (live) {
            SolveOut so;
            n2_solve<KT + 1, true>(R, s1, z0, kc, -1.0, so);
            if (bad) { so.status = THETA_ST_NONE_NAN; so.mu = 0.0; so.nll = 0.0; }
            p.mu_out[row] = so.mu;
            p.nll_out[row] = so.nll;
            p.status_out[row] = so.status;
        }
    }
}

template <int KT>
static int launch_solve(const theta_plan* pl, const TileParams& tp, cudaStream_t st) {
    constexpr int WARPS = 8, MINB = 2;
    const size_t smem = solve_smem_bytes<WARPS, KT>(tp.pitch);
    if (smem > 227 * 1024) return -1000;  // caller falls back to the CTA-granular kernel
    auto kern = solve_tiles_kernel<WARPS, MINB, KT>;
    THETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 1;
    THETA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
    if (bps < 1) bps = 1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    const int64_t n_sub = (tp.count + 31) / 32;
    const int grid = (int)std::min<int64_t>((n_sub + WARPS - 1) / WARPS, (int64_t)sms * bps);
    kern<<<grid, WARPS * 32, smem, st>>>(tp);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

static int g_nll_variant = -1;  // THETA_NLL_VARIANT: 0 = 8 warps x 2 stages, 1 = 4 x 2, 2 = 8 x 1, 3 = 4 x 1, 4 = 16 x 1, 5/6 = 8 x 1 with a register cap

template <int WARPS, int STAGES, int MINB, bool N3 = false>
static int launch_nll(const theta_plan* pl, const TileParams& tp, cudaStream_t st) {
    using Cfg = NllCfg<WARPS, STAGES>;
    const size_t smem = Cfg::smem_bytes(tp.pitch);
    auto kern = nll_tiles_kernel<WARPS, STAGES, MINB, N3>;
    if (smem > 227 * 1024) { set_error("sub-tiles of 32 rows x %lld bytes do not fit shared memory", (long long)tp.pitch); return THETA_ERR_UNSUPPORTED; }
    THETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 1;
    THETA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, Cfg::kThreads, smem));
    if (bps < 1) bps = 1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    const int64_t n_sub = (tp.count + 31) / 32;
    const int grid = (int)std::min<int64_t>((n_sub + WARPS - 1) / WARPS, (int64_t)sms * bps);
    kern<<<grid, Cfg::kThreads, smem, st>>>(tp);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

int launch_nll_n3(const theta_plan* pl, const TileParams& tp, cudaStream_t st) {
    if (NllCfg<16, 1>::smem_bytes(tp.pitch) > 227 * 1024) return launch_nll<4, 1, 8, true>(pl, tp, st);
    return launch_nll<16, 1, 2, true>(pl, tp, st);
}

static int launch_nll_auto(const theta_plan* pl, const TileParams& tp, cudaStream_t st) {
    if (g_nll_variant < 0) {
        const char* e = getenv("THETA_NLL_VARIANT");
        g_nll_variant = e ? atoi(e) : 4;
    }
    int v = g_nll_variant;
    // large pitches: fall back to the leanest shared-memory footprint that fits
    if (NllCfg<8, 2>::smem_bytes(tp.pitch) > 227 * 1024 && (v == 0 || v == 4)) v = 3;
    switch (v) {
        case 1: return launch_nll<4, 2, 5>(pl, tp, st);
        case 2: return launch_nll<8, 1, 4>(pl, tp, st);
        case 3: return launch_nll<4, 1, 8>(pl, tp, st);
        case 4: return launch_nll<16, 1, 2>(pl, tp, st);
        case 5: return launch_nll<8, 1, 5>(pl, tp, st);   // <= 48 registers: 40 warps per SM
        case 6: return launch_nll<8, 1, 6>(pl, tp, st);   // <= 40 registers: 48 warps per SM
        default: return launch_nll<8, 2, 2>(pl, tp, st);
    }
}

template <bool SOLVE, int KT>
static int launch_tiles(const theta_plan* pl, const TileParams& tp, cudaStream_t st) {
    const size_t smem = tile_smem_bytes(tp.pitch, SOLVE ? KT + 1 : 0);
    auto kern = tiles_n2_kernel<SOLVE, KT>;
    if (smem > 227 * 1024) { set_error("tile of %d rows x %lld bytes does not fit shared memory", kRows, (long long)tp.pitch); return THETA_ERR_UNSUPPORTED; }
    THETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 1;
    THETA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kRows, smem));
    if (bps < 1) bps = 1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    const int64_t n_tiles = (tp.count + kRows - 1) / kRows;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)sms * bps);
    kern<<<grid, kRows, smem, st>>>(tp);
    THETA_CUDA(cudaGetLastError());
    return THETA_OK;
}

static int check_tile_args(const theta_plan* pl, const void* cols, int64_t pitch, int64_t count) {
    if (!pl || !cols) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (pl->n != 2) { set_error("plan is not n=2"); return THETA_ERR_INVALID; }
    if (count < 0) { set_error("negative count"); return THETA_ERR_INVALID; }
    if (pl->m > 1023) { set_error("tile kernels support m <= 1023"); return THETA_ERR_UNSUPPORTED; }
    if (pitch != pl->tile_pitch) {
        set_error("pitch %lld invalid: tile kernels need theta_tile_pitch(m) = %lld", (long long)pitch, (long long)pl->tile_pitch);
        return THETA_ERR_INVALID;
    }
    if (reinterpret_cast<uintptr_t>(cols) & 15) { set_error("tile buffer must be 16-byte aligned"); return THETA_ERR_INVALID; }
    return THETA_OK;
}

}  // namespace theta

using namespace theta;

extern "C" int theta_nll_tiles_n2_dev(theta_plan* pl, const int8_t* cols_dev, int64_t pitch, int64_t count,
                                      const double* mu_dev, double* nll_dev, void* stream) {
    int rc = check_tile_args(pl, cols_dev, pitch, count);
    if (rc) return rc;
    if (!mu_dev || !nll_dev) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (count == 0) return THETA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    TileParams tp{};
    tp.cols = cols_dev; tp.pitch = pitch; tp.count = count; tp.PR = pl->d_PR; tp.PRN = pl->d_PRN; tp.ZP = pl->d_ZP;
    tp.kc = pl->kc; tp.mu_in = mu_dev; tp.nll_out = nll_dev; tp.logtab = pl->d_logtab;
    return launch_nll_auto(pl, tp, st);  // asynchronous on `stream`
}

// n = 3: rows are row-index bytes (one per interval, pitch = theta_tile_pitch(m), padding 0), mu is [count][3]
extern "C" int theta_nll_rows_n3_dev(theta_plan* pl, const uint8_t* rows_dev, int64_t pitch, int64_t count,
                                     const double* mu_dev, double* nll_dev, void* stream) {
    if (!pl || !rows_dev || !mu_dev || !nll_dev) { set_error("null argument"); return THETA_ERR_INVALID; }
    TileParams tp{};
    int rc = theta_n3_tile_info(pl, tp.ra, tp.rb, &tp.n_rows, &tp.kc);
    if (rc) return rc;
    if (count < 0) { set_error("negative count"); return THETA_ERR_INVALID; }
    if (pl->m > 1023) { set_error("tile kernels support m <= 1023"); return THETA_ERR_UNSUPPORTED; }
    if (pitch != pl->tile_pitch) {
        set_error("pitch %lld invalid: tile kernels need theta_tile_pitch(m) = %lld", (long long)pitch, (long long)pl->tile_pitch);
        return THETA_ERR_INVALID;
    }
    if (reinterpret_cast<uintptr_t>(rows_dev) & 15) { set_error("tile buffer must be 16-byte aligned"); return THETA_ERR_INVALID; }
    if (count == 0) return THETA_OK;
    tp.cols = reinterpret_cast<const int8_t*>(rows_dev); tp.pitch = pitch; tp.count = count;
    tp.PR = pl->d_PR; tp.PRN = pl->d_PRN; tp.mu_in = mu_dev; tp.nll_out = nll_dev; tp.logtab = pl->d_logtab;
    return launch_nll_n3(pl, tp, (cudaStream_t)stream);  // asynchronous on `stream`
}

extern "C" int theta_nll_rows_n3(theta_plan* pl, const uint8_t* rows_host, int64_t pitch, int64_t count,
                                 const double* mu_host, double* nll_host) {
    if (!pl || !rows_host || !mu_host || !nll_host) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (count <= 0) return THETA_OK;
    uint8_t* d_c = nullptr; double *d_mu = nullptr, *d_nll = nullptr;
    int rc = THETA_OK;
    cudaError_t e = theta::dmalloc(&d_c, (size_t)count * pitch);
    if (e == cudaSuccess) e = theta::dmalloc(&d_mu, sizeof(double) * 3 * count);
    if (e == cudaSuccess) e = theta::dmalloc(&d_nll, sizeof(double) * count);
    if (e == cudaSuccess) e = cudaMemcpy(d_c, rows_host, (size_t)count * pitch, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_mu, mu_host, sizeof(double) * 3 * count, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "nll_rows_n3 H2D", __FILE__, __LINE__);
    if (rc == THETA_OK) rc = theta_nll_rows_n3_dev(pl, d_c, pitch, count, d_mu, d_nll, nullptr);
    if (rc == THETA_OK) { e = cudaMemcpy(nll_host, d_nll, sizeof(double) * count, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = cuda_fail(e, "nll_rows_n3 D2H", __FILE__, __LINE__); }
    theta::dfree(d_c); theta::dfree(d_mu); theta::dfree(d_nll);
    return rc;
}

extern "C" int theta_solve_tiles_n2_dev(theta_plan* pl, const int8_t* cols_dev, int64_t pitch, int64_t count,
                                        double* mu_dev, double* nll_dev, int32_t* status_dev, void* stream) {
    int rc = check_tile_args(pl, cols_dev, pitch, count);
    if (rc) return rc;
    if (!mu_dev || !nll_dev || !status_dev) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (count == 0) return THETA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    {
        TileParams tp{};
        tp.cols = cols_dev; tp.pitch = pitch; tp.count = count; tp.PR = pl->d_PR; tp.PRN = pl->d_PRN; tp.ZP = pl->d_ZP;
        tp.kc = pl->kc; tp.mu_out = mu_dev; tp.nll_out = nll_dev; tp.status_out = status_dev; tp.logtab = pl->d_logtab;
        switch (pl->KT) {
#define THETA_TS(KTV) case KTV: rc = launch_solve<KTV>(pl, tp, st); if (rc == -1000) rc = launch_tiles<true, KTV>(pl, tp, st); break;
            THETA_TS(2) THETA_TS(3) THETA_TS(4) THETA_TS(5) THETA_TS(6) THETA_TS(7) THETA_TS(8) THETA_TS(10) THETA_TS(12) THETA_TS(15)
#undef THETA_TS
            default: set_error("unsupported level count"); rc = THETA_ERR_UNSUPPORTED;
        }
    }
    return rc;  // asynchronous on `stream`
}

extern "C" int theta_nll_tiles_n2(theta_plan* pl, const int8_t* cols_host, int64_t pitch, int64_t count,
                                  const double* mu_host, double* nll_host) {
    if (!pl || !cols_host || !mu_host || !nll_host) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (count <= 0) return THETA_OK;
    int8_t* d_c = nullptr; double *d_mu = nullptr, *d_nll = nullptr;
    int rc = THETA_OK;
    cudaError_t e = theta::dmalloc(&d_c, (size_t)count * pitch);
    if (e == cudaSuccess) e = theta::dmalloc(&d_mu, sizeof(double) * count);
    if (e == cudaSuccess) e = theta::dmalloc(&d_nll, sizeof(double) * count);
    if (e == cudaSuccess) e = cudaMemcpy(d_c, cols_host, (size_t)count * pitch, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_mu, mu_host, sizeof(double) * count, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "nll_tiles H2D", __FILE__, __LINE__);
    if (rc == THETA_OK) rc = theta_nll_tiles_n2_dev(pl, d_c, pitch, count, d_mu, d_nll, nullptr);
    if (rc == THETA_OK) { e = cudaMemcpy(nll_host, d_nll, sizeof(double) * count, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = cuda_fail(e, "nll_tiles D2H", __FILE__, __LINE__); }
    theta::dfree(d_c); theta::dfree(d_mu); theta::dfree(d_nll);
    return rc;
}

extern "C" int theta_solve_tiles_n2(theta_plan* pl, const int8_t* cols_host, int64_t pitch, int64_t count,
                                    double* mu_host, double* nll_host, int32_t* status_host) {
    if (!pl || !cols_host || !mu_host || !nll_host || !status_host) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (count <= 0) return THETA_OK;
    int8_t* d_c = nullptr; double *d_mu = nullptr, *d_nll = nullptr; int32_t* d_st = nullptr;
    int rc = THETA_OK;
    cudaError_t e = theta::dmalloc(&d_c, (size_t)count * pitch);
    if (e == cudaSuccess) e = theta::dmalloc(&d_mu, sizeof(double) * count);
    if (e == cudaSuccess) e = theta::dmalloc(&d_nll, sizeof(double) * count);
    if (e == cudaSuccess) e = theta::dmalloc(&d_st, sizeof(int32_t) * count);
    if (e == cudaSuccess) e = cudaMemcpy(d_c, cols_host, (size_t)count * pitch, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "solve_tiles H2D", __FILE__, __LINE__);
    if (rc == THETA_OK) rc = theta_solve_tiles_n2_dev(pl, d_c, pitch, count, d_mu, d_nll, d_st, nullptr);
    if (rc == THETA_OK) {
        e = cudaMemcpy(mu_host, d_mu, sizeof(double) * count, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(nll_host, d_nll, sizeof(double) * count, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(status_host, d_st, sizeof(int32_t) * count, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "solve_tiles D2H", __FILE__, __LINE__);
    }
    theta::dfree(d_c); theta::dfree(d_mu); theta::dfree(d_nll); theta::dfree(d_st);
    return rc;
}
