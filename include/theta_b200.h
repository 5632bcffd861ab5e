/*
 * theta_b200 -- C ABI of the B200-native THetA likelihood search.
 *
 * This header is the drop-in boundary for the hot path of raphael-group/THetA
 * (SURVEY.md section 8b).  The reference has no FFI layer: the boundary is the
 * Python call site python/RunTHetA.py:442-447 and the classes it binds.  Each
 * entry point below names the reference interface it replaces (file:line,
 * relative to the reference tree).  Signatures use plain pointers and sizes
 * only; no C++ / torch types cross the boundary.  The ctypes binding that a
 * reference maintainer would add is shown in INTEGRATION.md and implemented in
 * theta_b200/_lib.py.
 *
 * Conventions
 *   - every function returns THETA_OK (0) or a negative THETA_ERR_* code; the
 *     message for the last error on the calling thread is theta_last_error().
 *   - "host" pointers are ordinary host memory owned by the caller; "dev"
 *     pointers are CUDA device pointers on the plan's device, also caller-owned.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - interval vectors (r, rN, lower, upper) are in the *sorted* interval order
 *     produced by DataTools.sort_r (python/DataTools.py:95-118).
 *   - there is no CPU fallback: without a usable CUDA device every compute
 *     entry point fails with THETA_ERR_CUDA.
 */
#ifndef THETA_B200_H
#define THETA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THETA_OK 0
#define THETA_ERR_INVALID (-1)  /* bad argument */
#define THETA_ERR_CUDA (-2)     /* CUDA runtime error / no device */
#define THETA_ERR_EMPTY (-3)    /* no admissible candidate (RunTHetA.py:154-156, 217-219) */
#define THETA_ERR_OVERFLOW (-4) /* candidate count does not fit 63 bits, or an output buffer is too small */
#define THETA_ERR_UNSUPPORTED (-5)

/* Per-candidate status.  Values < 8 mean Optimizer.solve returned a solution,
 * values >= 8 mean it returned None (python/Optimizer.py:119-121, 159-161). */
#define THETA_ST_OK 0          /* interior root / KKT point */
#define THETA_ST_ROOT_LB 1     /* brenth: f(lB)==0, returns lB (Q17) */
#define THETA_ST_ROOT_UB 2     /* brenth: f(uB)==0, returns uB */
#define THETA_ST_FALLBACK_X0 3 /* n=3: BFGS fallback returned its start point (1/3,1/3,1/3) (H2 outcome a) */
#define THETA_ST_DEGENERATE 4  /* n=3: both tumour columns zero == all-normal model; NLL valid, mu undefined (Q2/Q8) */
#define THETA_ST_FALLBACK_MOVED 5 /* n=3: BFGS fallback took a step and stayed inside [0,1]^3 (not observed on the reference) */
#define THETA_ST_NONE_SIGN 8   /* f(lB), f(uB) have the same sign */
#define THETA_ST_NONE_NAN 9    /* NaN in the objective (zero column, 0/0) */
#define THETA_ST_NONE_NOCONV 10
#define THETA_ST_NONE_RANGE 11 /* n=3: fallback left [0,1]^3 (H2 outcome b) */

#define THETA_MAX_COPY 15  /* largest copy number supported by the n=2 kernels */
#define THETA_TIE_MARGIN 1e-3 /* Misc.isClose margin 10e-4, python/Misc.py:36 */

/* Problem statement at the do_optimization* boundary (RunTHetA.py:124-125, 173-174). */
typedef struct theta_problem {
    int32_t n;            /* populations: 2 or 3 */
    int32_t m;            /* intervals */
    int32_t tau;          /* copies in the normal genome */
    int32_t reserved;     /* flags: THETA_PLAN_TILES_ONLY */
    const int64_t* r;     /* [m] tumour read counts  (host) */
    const int64_t* rN;    /* [m] normal read counts  (host) */
    const int32_t* lower; /* [m] per-interval lower copy bound (host) */
    const int32_t* upper; /* [m] per-interval upper copy bound (host) */
    double max_normal;    /* Optimizer upper_bound (Optimizer.py:43); 1.0 = unconstrained */
    double min_normal;    /* Optimizer lower_bound (Optimizer.py:43,109-110); 0.0 = unconstrained.  RunTHetA never sets it:
                             honoured by the explicit-column solve (theta_solve_tiles_n2), not by the search kernels */
} theta_problem;
/* theta_problem.reserved flag: the plan only serves explicit columns (theta_solve_tiles_* / theta_nll_tiles_*): a space
 * whose candidate count exceeds 2^63 is accepted (count reported as 0, searches refused) */
#define THETA_PLAN_TILES_ONLY 1

/* One candidate record.  `rank` is the 0-based position in the reference's
 * enumeration order (Enumerator.py:119-152: ascending colex). */
typedef struct theta_cand {
    uint64_t rank;
    double nll;
    double mu[3];
    int32_t status;
    int32_t flags;
} theta_cand;

typedef struct theta_search_stats {
    uint64_t n_candidates; /* ranks evaluated (without the duplicated first matrix, Q1) */
    uint64_t n_solved;     /* candidates for which solve() != None */
    uint64_t n_stripes;
    uint64_t n_rescanned;  /* stripes re-evaluated by pass 2 */
    double min_nll;
    double newton_iters;   /* mean root-finder iterations per solved candidate */
    float ms_pass1;
    float ms_pass2;
    uint64_t n_exact;            /* pass 1 (n=2): candidates whose NLL was evaluated in fp64 (the rest were proven
                                    irrelevant by the fp32-log bound, theta_core.cuh::n2_l2_approx) */
    uint64_t n_bound_violations; /* THETA_VERIFY_APPROX=1 only: candidates whose fp32-log NLL missed the fp64 one by
                                    more than the stated bound (must be 0) */
} theta_search_stats;

typedef struct theta_plan theta_plan; /* opaque: device tables of one problem */

/* ---- runtime ---------------------------------------------------------- */
int theta_init(int device);              /* select device, create context; THETA_ERR_CUDA if none */
int theta_device_count(void);
void theta_shutdown(void);
const char* theta_last_error(void);
const char* theta_version(void);
/* measured fp64 FMA throughput of the current device in GFLOP/s (FMA = 2 flop); roofline denominator */
int theta_fp64_peak(double* gflops, void* stream);

/* ---- plans ------------------------------------------------------------ */
/* Replaces Enumerator.__init__ + Optimizer.__init__ (Enumerator.py:39-72, Optimizer.py:43-65):
 * monotonises the bounds (Enumerator.py:90-113), builds the count tables and
 * uploads r / rN prefix sums.  n=2 only for the rank-based entry points. */
int theta_plan_create(const theta_problem* p, theta_plan** out);
void theta_plan_destroy(theta_plan* plan);
/* TimeEstimate.count_number_matrices_2 (TimeEstimate.py:91-111). */
int theta_plan_count(const theta_plan* plan, uint64_t* count);
/* NLL of the all-normal model, -sum r_i log(rN_i / sum rN): what the reference evaluates for the
 * all-[tau,0,0] matrix that opens every n=3 search (RunTHetA.py:148,188 -- Q2). */
int theta_plan_null_nll(const theta_plan* plan, double* nll);
/* Monotonised bounds as used by the search (Q16). */
int theta_plan_bounds(const theta_plan* plan, int32_t* lower, int32_t* upper);

/* ---- enumeration (Enumerator.generate_next_C, Enumerator.py:74-87,119-160) ---- */
/* Tumour columns of ranks [rank_begin, rank_begin+count) as int8 rows of `pitch` bytes
 * (pitch >= m; bytes m..pitch-1 are zero).  Host and device destinations. */
int theta_enumerate_n2(theta_plan* plan, uint64_t rank_begin, int64_t count, int8_t* cols_host, int64_t pitch);
int theta_enumerate_n2_dev(theta_plan* plan, uint64_t rank_begin, int64_t count, int8_t* cols_dev, int64_t pitch,
                           void* stream);
int64_t theta_tile_pitch(int32_t m); /* row pitch used by the tile kernels (smallest 8*odd >= m) */

/* ---- fused search (do_optimization_single, RunTHetA.py:173-220; process_loop :53-93) ---- */
/* Pass 1: evaluate ranks [rank_begin, rank_end): enumerate on chip, solve mu, NLL, reduce per stripe.
 * Returns the minimum NLL of the range (+inf if no candidate solved). */
int theta_search_n2_pass1(theta_plan* plan, uint64_t rank_begin, uint64_t rank_end, double* local_min, void* stream);
/* Pass 1 split for multi-GPU runs (replaces the merge of the workers' results, RunTHetA.py:107-122,162-171):
 * _begin enqueues the search on `stream` and, when local_min_dev is not NULL, also writes the shard minimum to
 * that DEVICE double -- so the caller can enqueue ncclAllReduce(min) on it right behind, with no host round
 * trip; _end waits for the stream and returns the shard minimum on the host. */
int theta_search_n2_pass1_begin(theta_plan* plan, uint64_t rank_begin, uint64_t rank_end, double* local_min_dev,
                                void* stream);
int theta_search_n2_pass1_end(theta_plan* plan, double* local_min);
/* Pass 2: given the global minimum (all shards) and the minimum over all ranks *before* this shard
 * (+inf for the first shard), re-evaluate only the stripes that can matter and emit, in rank order,
 * a superset of the candidates the reference's running-min fold (RunTHetA.py:198-206) can ever keep:
 * every strict prefix minimum and every candidate with NLL <= global_min + 2*margin. */
int theta_search_n2_pass2(theta_plan* plan, double global_min, double prior_min, theta_cand* out_host, uint64_t cap,
                          uint64_t* n_out, void* stream);
/* The same with the global minimum read from DEVICE memory (NULL: this shard's own minimum), so that nothing
 * returns to the host between pass 1, the all-reduce(min) and pass 2: the stripe list is built by a device-side
 * prefix-min scan, the emit kernel is launched behind it and the host waits once.  Completes a pending
 * theta_search_n2_pass1_begin (no _end needed); *global_min_host receives the value that was used. */
int theta_search_n2_pass2_dev(theta_plan* plan, const double* global_min_dev, double prior_min, theta_cand* out_host,
                              uint64_t cap, uint64_t* n_out, double* global_min_host, void* stream);
/* Fully asynchronous form for the multi-GPU exchange (replaces returnQueue / find_mins, RunTHetA.py:141,162-171):
 * pass 2 is enqueued and its status block + this shard's records -- sorted by rank and filtered ON THE DEVICE to what the
 * fold replay needs (strict prefix minima w.r.t. prior_min, everything within 2.5 margins of the global minimum, NaN) --
 * land in the caller's DEVICE buffer [uint64 status[32] | theta_cand records[cap]], the operand of the all-gather, with no
 * host synchronisation.  status[11] = records kept; > cap means they did not fit (retry with a larger block).
 * After synchronising, the caller passes its own status words to _async_done (statistics) and the concatenation of all
 * shards' blocks to theta_filter_records (the same filter over the union: a linear pass, the blocks arrive sorted). */
int theta_search_n2_pass2_enqueue(theta_plan* plan, const double* global_min_dev, double prior_min, void* out_dev,
                                  uint64_t cap, void* stream);
int theta_search_n2_async_done(theta_plan* plan, const unsigned long long* status_host);
/* Host-side bookkeeping of theta_search_n2_pass1_begin for a step whose device work is replayed from a CUDA graph
 * captured from the very same call sequence (theta_search_n2_dist does this from the second identical step on). */
int theta_search_n2_graph_replay(theta_plan* plan, uint64_t rank_begin, uint64_t rank_end, void* stream);
int theta_filter_records(theta_cand* recs, uint64_t n, double prior_min, double near_thr, uint64_t* n_out);
int theta_search_stats_get(const theta_plan* plan, theta_search_stats* stats);
/* Single-shard convenience: pass1 + pass2 with global_min = local_min, prior_min = +inf. */
int theta_search_n2(theta_plan* plan, uint64_t rank_begin, uint64_t rank_end, theta_cand* out_host, uint64_t cap,
                    uint64_t* n_out);
/* do_optimization_single (RunTHetA.py:173-220) for n=2 in one call, host buffers in and out: plan creation, both
 * passes, the reference's running-min fold over the emitted records, and for each surviving solution (fold order,
 * up to cap_best) its record, its tumour column cols[k][m] (sorted interval order) and Optimizer.solve's vals[k][m].
 * rank_limit = 0 searches the whole space.  lower_out / upper_out receive the monotonised bounds (Q16). */
int theta_optimize_n2(const theta_problem* p, uint64_t rank_limit, int32_t* lower_out, int32_t* upper_out,
                      uint64_t cap_best, uint64_t* n_best, theta_cand* best, int8_t* cols, double* vals,
                      theta_search_stats* stats, uint64_t* n_candidates);
/* Tumour columns [count][m] of given ranks (host-side unranking, no launch). */
int theta_columns_of_ranks(const theta_plan* plan, const uint64_t* ranks, int64_t count, int8_t* cols);
/* (mu, NLL, status, vals[m]) of given ranks -- Optimizer.solve's return for the winners (Optimizer.py:123-126). */
int theta_eval_ranks_n2(theta_plan* plan, const uint64_t* ranks_host, int64_t count, theta_cand* out_host,
                        double* vals_host /* [count][m] or NULL */);

/* ---- tile kernels: explicit C matrices (Optimizer.solve, Optimizer.py:68-126; CalcAllC.L2, CalcAllC.py:44-61) ---- */
/* n=2, tumour column per candidate as int8 rows of `pitch` bytes. */
int theta_solve_tiles_n2_dev(theta_plan* plan, const int8_t* cols_dev, int64_t pitch, int64_t count, double* mu_dev,
                             double* nll_dev, int32_t* status_dev, void* stream);
int theta_solve_tiles_n2(theta_plan* plan, const int8_t* cols_host, int64_t pitch, int64_t count, double* mu_host,
                         double* nll_host, int32_t* status_host);
/* NLL at a given mu (the streaming "NLL kernel": algorithmic bytes m + 8 + 8 per candidate). */
int theta_nll_tiles_n2_dev(theta_plan* plan, const int8_t* cols_dev, int64_t pitch, int64_t count,
                           const double* mu_dev, double* nll_dev, void* stream);
int theta_nll_tiles_n2(theta_plan* plan, const int8_t* cols_host, int64_t pitch, int64_t count, const double* mu_host,
                       double* nll_host);
/* n=3 (Optimizer.L3, Optimizer.py:236-244; CalcAllC.L3, CalcAllC.py:63-76): one row-index byte per interval
 * (index into theta_n3_rows(), the compact form of the m x 3 matrices of Enumerator._to_matrix), rows of
 * `pitch` = theta_tile_pitch(m) bytes with zero padding; mu is [count][3].  Algorithmic bytes m + 24 + 8. */
int theta_nll_rows_n3_dev(theta_plan* plan, const uint8_t* rows_dev, int64_t pitch, int64_t count, const double* mu_dev,
                          double* nll_dev, void* stream);
int theta_nll_rows_n3(theta_plan* plan, const uint8_t* rows_host, int64_t pitch, int64_t count, const double* mu_host,
                      double* nll_host);

/* ---- n = 3 (Enumerator._generate_next_C_3, Enumerator.py:172-298; Optimizer._solve_n3plus, Optimizer.py:128-165) ---- */
/* Row table of the plan: rows[g] = (a_g, b_g) in the reference's order (first component fastest). */
int theta_n3_rows(const theta_plan* plan, int32_t* a, int32_t* b, int32_t* n_rows);
/* Cut the DFS forest into lexicographically ordered work items (sub-trees) of at most ~leaf_cap
 * candidates (count DP of TimeEstimate.count_number_matrices_3, TimeEstimate.py:113-142, per sub-tree).
 * leaf_cap <= 0: 256, or less for small spaces so that there are enough items to fill the GPU. */
int theta_n3_prepare(theta_plan* plan, int64_t leaf_cap, int64_t max_items, uint64_t* n_items);
int theta_n3_item_count(const theta_plan* plan, uint64_t* n_items);
int theta_n3_item_weights(const theta_plan* plan, double* weights /* [n_items] estimated leaves */);
/* 1 if theta_n3_prepare stopped at max_items: the items then cover only a lexicographic prefix of the space. */
int theta_n3_truncated(const theta_plan* plan, int32_t* truncated);
/* n=3 search of items [item_begin, item_end).  The fold of RunTHetA.py:198-206 is determined by the rank-ordered
 * candidates inside the 1e-3-connected cluster of NLL values that contains the global minimum, so the
 * search needs no prefix-minimum history:
 *   theta_search_n3_min     local_min = exact minimum NLL of the shard.  Every leaf is screened by a rigorous lower
 *                           bound against T + window + 1e-3, T = best exact NLL known (t_seed: an upper bound of the
 *                           global minimum known beforehand, e.g. another shard's; +inf if none); survivors are solved
 *                           and kept in the plan.  window <= 0: 4e-3.
 *   theta_search_n3_window  every candidate of the shard with NLL <= global_min + window, or a NaN likelihood (Q6), in
 *                           enumeration order, re-solved in reference order; rank = (item << 32) | position inside the
 *                           item, flags indexes paths_host.  The caller checks that the cluster ends inside the window
 *                           (a gap > 1e-3 below global_min + window) and otherwise searches again with a wider one. */
int theta_search_n3_min(theta_plan* plan, uint64_t item_begin, uint64_t item_end, double t_seed, double window,
                        double* local_min, void* stream);
int theta_search_n3_window(theta_plan* plan, double global_min, double window, theta_cand* out_host, uint64_t cap,
                           uint64_t* n_out, uint8_t* paths_host /* [cap][m] row indices */, void* stream);
/* Every candidate of items [item_begin, item_end) in enumeration order, as row-index strings. */
int theta_enumerate_n3(theta_plan* plan, uint64_t item_begin, uint64_t item_end, uint8_t* paths_host, uint64_t cap,
                       uint64_t* n_out);
/* Optimizer.solve on explicit candidates given as row-index strings [count][m].
 * mode 0: grouped-statistics solver, except rank-deficient candidates (rows on one line / a zero tumour column),
 *         which take the reference-order solver (per-interval sums in the reference's arithmetic order, libm-exact
 *         pow, MINPACK hybrj + hybrd restated: the only way to reproduce what the reference returns for them);
 * mode 1: reference-order solver for every candidate;  mode 2: grouped solver for every candidate. */
int theta_solve_rows_n3(theta_plan* plan, const uint8_t* paths_host, int64_t count, double* mu_host /* [count][3] */,
                        double* muhat_host /* [count][3] or NULL */, double* nll_host, int32_t* status_host,
                        double* vals_host /* [count][m] or NULL */, int32_t mode);

/* ---- the search across GPUs: replaces the multiprocessing farm of do_optimization (RunTHetA.py:96-171:
 * start_processes / process_loop / returnQueue / find_mins).  Per GPU, in stream order: pass 1 on its shard, ONE
 * ncclAllReduce(min) on a device double, pass 2 / window records, ncclAllGather of the few records, one host
 * synchronisation; then the reference's fold on the host.  NCCL is bound at run time (dlopen libnccl.so.2). ---- */
typedef struct theta_comm theta_comm; /* one rank's communicator + exchange buffers (current device) */
/* One process per GPU (torchrun): rank 0 makes the id, it travels through the launcher's own transport. */
int theta_nccl_unique_id(uint8_t* id128 /* [128] */);
int theta_comm_create(int32_t rank, int32_t world, const uint8_t* id128 /* NULL iff world == 1 */, theta_comm** out);
void theta_comm_destroy(theta_comm* comm);
/* One process per GPU, optional: share the ranks' exchange areas through CUDA IPC so that the two collectives of an n=2
 * step run as remote stores over NVLink peer memory inside our own kernels instead of NCCL calls.  Every rank publishes
 * theta_comm_ipc_handle (128 bytes), the launcher all-gathers them, every rank calls theta_comm_open_peers.  Without this
 * (or with THETA_EXCHANGE=nccl, or when the GPUs cannot map each other) the step uses ncclAllReduce / ncclAllGather. */
int theta_comm_ipc_handle(theta_comm* comm, uint8_t* handle128);
int theta_comm_open_peers(theta_comm* comm, const uint8_t* handles /* [world][128] */);
/* Every rank must use the same exchange: query (enable < 0) this rank's state after theta_comm_open_peers, reduce it
 * over the ranks with MIN, and set the agreed value (enable = 0 switches this rank to the NCCL collectives).  Returns
 * the state after the call (1 = peer memory). */
int theta_comm_peer_exchange(theta_comm* comm, int32_t enable);
/* One sharded search step on this rank; every rank receives the union of the records, in enumeration order.
 * n=2: records = what the fold replay needs (theta_filter_records).  n=3: every candidate within `window` of the
 * global minimum (extra_min: NLL of the all-[tau,0,0] pseudo candidate, folded into the minimum), flags index
 * paths_host; totals = {solved, candidates} summed over all shards. */
int theta_search_n2_dist(theta_plan* plan, theta_comm* comm, uint64_t rank_begin, uint64_t rank_end, theta_cand* out_host,
                         uint64_t cap, uint64_t* n_out, double* global_min, uint64_t* n_solved_total, void* stream);
int theta_search_n3_dist(theta_plan* plan, theta_comm* comm, uint64_t item_begin, uint64_t item_end, double window,
                         double extra_min, theta_cand* out_host, uint8_t* paths_host, uint64_t cap, uint64_t* n_out,
                         double* global_min, uint64_t* totals, void* stream);
/* ONE process, n_devices GPUs: do_optimization(max_processes) (RunTHetA.py:124-171) behind one blocking call
 * (ncclCommInitAll + one persistent host thread and stream per GPU).  devices NULL: 0..n_devices-1.
 *   budget  n=2: rank limit (0 = whole space); n=3: work-item budget (0 = whole space; a space that exceeds the item
 *           limit is an error unless a budget is given)
 *   rows    [cap_best][m]: n=2 tumour column / n=3 row indices (theta_n3_rows) of each solution, in the reference's
 *           fold order; rank == UINT64_MAX marks the all-[tau,0,0] pseudo candidate of n=3 (Q2)
 *   vals    [cap_best][m]: Optimizer.solve's third return value.   *n_best may exceed cap_best (THETA_ERR_OVERFLOW). */
typedef struct theta_mgpu theta_mgpu;
int theta_mgpu_create(const theta_problem* p, int32_t n_devices, const int32_t* devices, theta_mgpu** out);
void theta_mgpu_destroy(theta_mgpu* h);
int theta_mgpu_devices(const theta_mgpu* h);
int theta_mgpu_optimize(theta_mgpu* h, uint64_t budget, int32_t* lower_out, int32_t* upper_out, uint64_t cap_best,
                        uint64_t* n_best, theta_cand* best, uint8_t* rows, double* vals, theta_search_stats* stats,
                        uint64_t* n_candidates);
int theta_mgpu_last_ms(const theta_mgpu* h, float* ms /* [n_devices] device time of the last search */);
/* One-shot: create + optimize + destroy. */
int theta_optimize(const theta_problem* p, int32_t n_devices, const int32_t* devices, uint64_t budget, int32_t* lower_out,
                   int32_t* upper_out, uint64_t cap_best, uint64_t* n_best, theta_cand* best, uint8_t* rows, double* vals,
                   theta_search_stats* stats, uint64_t* n_candidates);

/* ---- CalcAllC.L2 / L3 (CalcAllC.py:44-76): masked NLL of one weighted matrix at a given mu ---- */
/* Cw is [m][n] row-major fp64 (already weighted by rN); valid_i = Cw[i][0] != 0 (L2 with mu==0: Cw[i][1] != 0).
 * vals[i] = w_i / D (meaningful where valid[i]); does NOT modify Cw (the Python wrapper reproduces Q10). */
int theta_nll(int32_t n, int32_t m, const double* mu, const double* Cw, const int64_t* r, double* nll, double* vals,
              uint8_t* valid);

#ifdef __cplusplus
}
#endif
#endif /* THETA_B200_H */
