// theta_mgpu.cu -- the search across GPUs: replaces the multiprocessing farm of do_optimization
// (python/RunTHetA.py:96-171: start_processes / process_loop / returnQueue / find_mins).
//
// The candidate space shards embarrassingly (n=2: contiguous rank ranges; n=3: contiguous item ranges of equal
// estimated weight).  One search is, per GPU and with nothing but stream order in between:
//     pass 1 on the shard  ->  ONE ncclAllReduce(min) on a device double  ->  pass 2 / window records
//     ->  ncclAllGather of the few records  ->  one host synchronisation  ->  the reference's fold (host, exact)
// Two ways to drive it, same code underneath (theta_search_n2_dist / theta_search_n3_dist on a theta_comm):
//   * one process per GPU (torchrun): theta_nccl_unique_id on rank 0, the id travels through the launcher's own
//     transport, theta_comm_create on every rank (ncclCommInitRank);
//   * ONE process, G GPUs (theta_mgpu_*): ncclCommInitAll + one persistent host thread per GPU behind a single
//     blocking call -- what do_optimization(max_processes=G) / --NUM_PROCESSES G maps to.
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already in the process if there is one -- PyTorch
// ships its own), so the library has no link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "theta_n2.cuh"

using namespace theta;

namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the process's own copy (PyTorch's) first
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.h = h;
#define THETA_SYM(field, name) *reinterpret_cast<void**>(&api.field) = dlsym(h, name)
        THETA_SYM(GetUniqueId, "ncclGetUniqueId");
        THETA_SYM(CommInitRank, "ncclCommInitRank");
        THETA_SYM(CommInitAll, "ncclCommInitAll");
        THETA_SYM(CommDestroy, "ncclCommDestroy");
        THETA_SYM(AllReduce, "ncclAllReduce");
        THETA_SYM(AllGather, "ncclAllGather");
        THETA_SYM(GetErrorString, "ncclGetErrorString");
#undef THETA_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllReduce && api.AllGather;
    });
    return api;
}

int nccl_fail(ncclResult_t r, const char* what) {
    set_error("NCCL error %d (%s) in %s", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
    return THETA_ERR_CUDA;
}
#define THETA_NCCL(expr)                                  \
    do {                                                  \
        ncclResult_t _r = (expr);                         \
        if (_r != ncclSuccess) return nccl_fail(_r, #expr); \
    } while (0)

constexpr int kStatusWordsX = 32;  // theta_n2.cu: status block of the asynchronous pass 2
constexpr int kStEmitted = 8 + 3;
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Exchange over NVLink peer memory (the default when every GPU can map every other): the two collectives of a step are
// 8 bytes and ~12 KB per rank, i.e. pure latency, and as NCCL calls they were most of what an 8-GPU step added to a
// 1-GPU step (two kernel launches with their protocol set-up per rank, completion visible to the host only after NCCL's
// own bookkeeping).  Here each is ONE tiny kernel of ours: every rank stores its value / its record block straight into
// every peer's exchange area (remote stores through NVSwitch), publishes a step tag with release semantics, and spins on
// the tags the peers publish in its own area.  Tags are the step number (monotone), so nothing is ever reset; a spin that
// exceeds ~10 s sets an error word instead of hanging the GPU.  ncclAllReduce / ncclAllGather remain the fallback
// (THETA_EXCHANGE=nccl, no peer access, or a record block that outgrows the fixed peer slot).
constexpr int kMaxPeers = 16;
constexpr size_t kPeerRecs = 1024;  // records per rank that fit the fixed peer slot
struct XArea {                      // one per rank, device memory, zero-initialised
    struct { double val; unsigned long long tag; } mins[kMaxPeers];
    unsigned long long rtag[kMaxPeers];
    unsigned long long cur;         // step number of the exchange in flight (written by peer_min_kernel only)
    unsigned long long err;
};
struct PeerTable {
    XArea* area[kMaxPeers];
    unsigned char* recv[kMaxPeers];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until(const unsigned long long* p, unsigned long long want, unsigned long long* err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        if (clock64() - t0 > 20000000000LL) { *err = 1ULL; return false; }
    }
    return true;
}

// all-reduce(min) of one double per rank: in place on *val
__global__ void peer_min_kernel(PeerTable T, int G, int rank, double* val) {
    const int lane = threadIdx.x;
    XArea* mine = T.area[rank];
    const unsigned long long seq = mine->cur + 1ULL;
    __syncwarp();
    if (lane == 0) mine->cur = seq;
    const double v = *val;
    if (lane < G) {
        XArea* peer = T.area[lane];
        peer->mins[rank].val = v;
        __threadfence_system();
        st_release_sys(&peer->mins[rank].tag, seq);
    }
    double x = INFINITY;
    if (lane < G && spin_until(&mine->mins[lane].tag, seq, &mine->err)) x = mine->mins[lane].val;
    for (int o = 16; o; o >>= 1) x = fmin(x, __shfl_xor_sync(0xffffffffu, x, o));
    if (lane == 0) *val = x;
}

// all-gather of one block of `bytes` per rank: block b of the grid serves peer b
__global__ void peer_gather_kernel(PeerTable T, int G, int rank, const unsigned char* send, size_t bytes, size_t slot) {
    const int b = blockIdx.x;
    XArea* mine = T.area[rank];
    const unsigned long long seq = mine->cur;
    const uint4* src = reinterpret_cast<const uint4*>(send);
    uint4* dst = reinterpret_cast<uint4*>(T.recv[b] + (size_t)rank * slot);
    for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(&T.area[b]->rtag[rank], seq);
        spin_until(&mine->rtag[b], seq, &mine->err);  // ... and peer b's block has arrived here
    }
}

struct theta_comm {
    ncclComm_t comm = nullptr;  // NULL: a single shard, no collective
    int rank = 0, world = 1, device = 0;
    double* d_min = nullptr;              // operand of the all-reduce(min)
    double* h_min = nullptr;              // pinned
    unsigned char* d_send = nullptr;      // this shard's exchange block
    unsigned char* d_recv = nullptr;      // [world] blocks
    unsigned char* h_recv = nullptr;      // pinned copy of d_recv
    size_t block_bytes = 0;
    uint64_t xcap = 0;                    // n=2: records per exchange block
    cudaEvent_t dbg[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // THETA_DEBUG=1: phase times of a step
    // CUDA graph of the n=2 step (pass 1, min reduction, all-reduce, scan, pass 2, all-gather, D2H): a repeated search of
    // the same plan and range costs one graph launch instead of ~20 driver calls per rank -- with 8 host threads (or 8
    // ranks) issuing them, those calls and their jitter were half of a 1-ms step
    cudaGraphExec_t g_exec = nullptr;
    uint64_t g_key[5] = {0, 0, 0, 0, 0};
    int g_seen = 0;
    bool g_disabled = false;              // a capture failed once: stay on the eager path
    // peer-memory exchange (see above)
    XArea* x_area = nullptr;
    unsigned char* x_recv = nullptr;      // [world] slots of x_slot bytes
    size_t x_slot = 0;
    PeerTable peers{};
    bool peer_ok = false;
    unsigned long long* h_err = nullptr;  // pinned
    cudaStream_t st_own = nullptr;        // used when the caller passes the legacy default stream (which cannot be captured)
    std::vector<theta_cand> all;          // the gathered records of a step (reused)
};

static int comm_buffers(theta_comm* c, size_t block_bytes) {
    if (block_bytes <= c->block_bytes) return THETA_OK;
    theta::dfree(c->d_send); theta::dfree(c->d_recv); theta::hfree(c->h_recv);
    c->d_send = c->d_recv = c->h_recv = nullptr;
    c->block_bytes = 0;
    THETA_CUDA(theta::dmalloc(&c->d_send, block_bytes));
    THETA_CUDA(theta::dmalloc(&c->d_recv, block_bytes * c->world));
    THETA_CUDA(theta::hmalloc(&c->h_recv, block_bytes * c->world));
    c->block_bytes = block_bytes;
    return THETA_OK;
}

static int comm_init_common(theta_comm* c) {
    THETA_CUDA(cudaGetDevice(&c->device));
    THETA_CUDA(theta::dmalloc(&c->d_min, sizeof(double) * 2));
    THETA_CUDA(theta::hmalloc(&c->h_min, sizeof(double) * 2));
    c->xcap = 256;
    if (c->world > 1 && c->world <= kMaxPeers) {
        c->x_slot = sizeof(unsigned long long) * kStatusWordsX + sizeof(theta_cand) * kPeerRecs;
        THETA_CUDA(theta::dmalloc(&c->x_area, sizeof(XArea)));
        THETA_CUDA(theta::dmalloc(&c->x_recv, c->x_slot * c->world));
        THETA_CUDA(theta::hmalloc(&c->h_err, sizeof(unsigned long long)));
        THETA_CUDA(cudaMemset(c->x_area, 0, sizeof(XArea)));
    }
    return comm_buffers(c, sizeof(unsigned long long) * kStatusWordsX + sizeof(theta_cand) * c->xcap);
}

static bool peer_exchange_wanted() {
    const char* e = getenv("THETA_EXCHANGE");
    return !(e && (e[0] == 'n' || e[0] == 'N'));  // "nccl" forces the NCCL collectives
}

// One process per GPU: the exchange areas are shared through CUDA IPC.  handle128 = {area handle (64 B), recv handle (64 B)}.
extern "C" int theta_comm_ipc_handle(theta_comm* c, uint8_t* handle128) {
    if (!c || !handle128) { set_error("null argument"); return THETA_ERR_INVALID; }
    std::memset(handle128, 0, 128);
    if (!c->x_area) return THETA_OK;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t ha, hr;
    THETA_CUDA(cudaIpcGetMemHandle(&ha, c->x_area));
    THETA_CUDA(cudaIpcGetMemHandle(&hr, c->x_recv));
    std::memcpy(handle128, &ha, 64);
    std::memcpy(handle128 + 64, &hr, 64);
    return THETA_OK;
}

extern "C" int theta_comm_open_peers(theta_comm* c, const uint8_t* handles /* [world][128] */) {
    if (!c || !handles) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (!c->x_area || !peer_exchange_wanted()) return THETA_OK;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->peers.area[r] = c->x_area; c->peers.recv[r] = c->x_recv; continue; }
        cudaIpcMemHandle_t ha, hr;
        std::memcpy(&ha, handles + (size_t)r * 128, 64);
        std::memcpy(&hr, handles + (size_t)r * 128 + 64, 64);
        void *pa = nullptr, *pr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&pa, ha, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&pr, hr, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); c->peer_ok = false; return THETA_OK; }  // no peer mapping: NCCL path
        c->peers.area[r] = static_cast<XArea*>(pa);
        c->peers.recv[r] = static_cast<unsigned char*>(pr);
    }
    c->peer_ok = true;
    return THETA_OK;
}

// Query (enable < 0) or set the peer-memory exchange of a communicator.  Every rank of a communicator must use the same
// exchange (a rank on NCCL would wait for collectives its peers never post): after theta_comm_open_peers the launcher
// reduces the ranks' answers with MIN and sets the agreed value.  Returns the state after the call (1 = peer memory).
extern "C" int theta_comm_peer_exchange(theta_comm* c, int32_t enable) {
    if (!c) return 0;
    if (enable == 0) c->peer_ok = false;
    // (enabling is only possible when this rank has mapped every peer: open_peers decides that)
    return c->peer_ok ? 1 : 0;
}

extern "C" int theta_nccl_unique_id(uint8_t* id128) {
    if (!id128) { set_error("null argument"); return THETA_ERR_INVALID; }
    if (!nccl().ok) { set_error("libnccl.so.2 not found: multi-GPU search needs NCCL"); return THETA_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    THETA_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return THETA_OK;
}

extern "C" int theta_comm_create(int32_t rank, int32_t world, const uint8_t* id128, theta_comm** out) {
    if (!out || world < 1 || rank < 0 || rank >= world) { set_error("bad communicator arguments"); return THETA_ERR_INVALID; }
    *out = nullptr;
    theta_comm* c = new theta_comm();
    c->rank = rank; c->world = world;
    int rc = comm_init_common(c);
    if (!rc && world > 1) {
        if (!id128) { set_error("null NCCL id"); rc = THETA_ERR_INVALID; }
        else if (!nccl().ok) { set_error("libnccl.so.2 not found: multi-GPU search needs NCCL"); rc = THETA_ERR_UNSUPPORTED; }
        else {
            ncclUniqueId id;
            std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
            ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
            if (r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitRank");
        }
    }
    if (rc) { delete c; return rc; }
    *out = c;
    return THETA_OK;
}

extern "C" void theta_comm_destroy(theta_comm* c) {
    if (!c) return;
    if (c->g_exec) cudaGraphExecDestroy(c->g_exec);
    if (c->st_own) cudaStreamDestroy(c->st_own);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    theta::dfree(c->d_min); theta::hfree(c->h_min); theta::dfree(c->d_send); theta::dfree(c->d_recv); theta::hfree(c->h_recv);
    theta::dfree(c->x_area); theta::dfree(c->x_recv); theta::hfree(c->h_err);  // (IPC mappings of the peers end with the process)
    delete c;
}

// One sharded n=2 search step on this rank (see the file header).  Every rank receives the union of all shards'
// records, sorted by rank and filtered to what the fold replay needs (theta_filter_records).
extern "C" int theta_search_n2_dist(theta_plan* pl, theta_comm* c, uint64_t rank_begin, uint64_t rank_end,
                                    theta_cand* out_host, uint64_t cap, uint64_t* n_out, double* global_min,
                                    uint64_t* n_solved_total, void* stream) {
    if (!pl || !c || !n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_out = 0;
    if (!stream) {  // the step is self-contained and ends with a host wait: run it on the communicator's own stream
        if (!c->st_own) THETA_CUDA(cudaStreamCreateWithFlags(&c->st_own, cudaStreamNonBlocking));
        stream = c->st_own;
    }
    cudaStream_t st = (cudaStream_t)stream;
    static const bool debug = getenv("THETA_DEBUG") != nullptr;
    static const bool graph_env = !debug && !(getenv("THETA_NO_GRAPH") && getenv("THETA_NO_GRAPH")[0] == '1');
    const bool use_graph = graph_env && !c->g_disabled;
    static const bool hdbg = getenv("THETA_DEBUG_HOST") != nullptr;  // host-side timeline of the step, one line per call
    const auto th0 = std::chrono::steady_clock::now();
    auto th1 = th0, th2 = th0;
    const char* how = "eager";
    int attempts = 0;
    if (debug && !c->dbg[0]) for (int i = 0; i < 5; ++i) cudaEventCreate(&c->dbg[i]);
    const size_t head = sizeof(unsigned long long) * kStatusWordsX;
    int rc = THETA_OK;
    // everything one step enqueues on the stream (also what the graph captures)
    // (with_pass1 = false: the retry after a record block that outgrew the exchange slot -- the stripe minima of pass 1 and
    // the all-reduced minimum are still in place, so only the exchange of the minimum (idempotent), pass 2 and the gather
    // are repeated with the larger slot)
    auto enqueue = [&](size_t bytes, bool with_pass1 = true) -> int {
        if (debug) cudaEventRecord(c->dbg[0], st);
        int r = with_pass1 ? theta_search_n2_pass1_begin(pl, rank_begin, rank_end, c->d_min, stream) : THETA_OK;
        if (r) return r;
        if (debug) cudaEventRecord(c->dbg[1], st);
        const bool peer = c->world > 1 && c->peer_ok && c->xcap <= kPeerRecs;
        if (peer) peer_min_kernel<<<1, 32, 0, st>>>(c->peers, c->world, c->rank, c->d_min);
        else if (c->comm) THETA_NCCL(nccl().AllReduce(c->d_min, c->d_min, 1, ncclDouble, ncclMin, c->comm, st));
        if (debug) cudaEventRecord(c->dbg[2], st);
        r = theta_search_n2_pass2_enqueue(pl, c->d_min, INFINITY, c->d_send, c->xcap, stream);
        if (r) return r;
        if (debug) cudaEventRecord(c->dbg[3], st);
        if (peer) {
            peer_gather_kernel<<<c->world, 256, 0, st>>>(c->peers, c->world, c->rank, c->d_send, bytes, c->x_slot);
            THETA_CUDA(cudaGetLastError());
            THETA_CUDA(cudaMemcpy2DAsync(c->h_recv, bytes, c->x_recv, c->x_slot, bytes, c->world, cudaMemcpyDeviceToHost, st));
            THETA_CUDA(cudaMemcpyAsync(c->h_err, &c->x_area->err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        } else if (c->comm) {
            THETA_NCCL(nccl().AllGather(c->d_send, c->d_recv, bytes, ncclUint8, c->comm, st));
            THETA_CUDA(cudaMemcpyAsync(c->h_recv, c->d_recv, bytes * c->world, cudaMemcpyDeviceToHost, st));
        } else {
            THETA_CUDA(cudaMemcpyAsync(c->h_recv, c->d_send, bytes, cudaMemcpyDeviceToHost, st));
        }
        if (debug) cudaEventRecord(c->dbg[4], st);
        return THETA_OK;
    };
    for (int attempt = 0;; ++attempt) {
        const size_t bytes = head + sizeof(theta_cand) * c->xcap;
        rc = comm_buffers(c, bytes);
        if (rc) return rc;
        const uint64_t key[5] = {pl->uid, rank_begin, rank_end, c->xcap, (uint64_t)(uintptr_t)st};
        const bool same = std::memcmp(key, c->g_key, sizeof(key)) == 0;
        if (attempt > 0) {
            if (c->g_exec) { cudaGraphExecDestroy(c->g_exec); c->g_exec = nullptr; }
            std::memcpy(c->g_key, key, sizeof(key));
            c->g_seen = 0;
            rc = enqueue(bytes, false);
            if (rc) return rc;
        } else if (use_graph && same && c->g_exec) {
            theta_search_n2_graph_replay(pl, rank_begin, rank_end, stream);  // host-side bookkeeping of pass1_begin
            THETA_CUDA(cudaGraphLaunch(c->g_exec, st));
            how = "graph";
        } else {
            if (!same) {
                if (c->g_exec) { cudaGraphExecDestroy(c->g_exec); c->g_exec = nullptr; }
                std::memcpy(c->g_key, key, sizeof(key));
                c->g_seen = 0;
            }
            if (use_graph && same && ++c->g_seen >= 1) {
                // second identical step: capture it (the first ran eagerly and sized every buffer)
                // (anything going wrong here just means: no graph for this communicator)
                cudaGraph_t graph = nullptr;
                cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
                bool ok = e == cudaSuccess;
                if (ok) {
                    const int rcc = enqueue(bytes);
                    e = cudaStreamEndCapture(st, &graph);
                    ok = rcc == THETA_OK && e == cudaSuccess && graph;
                }
                if (ok) {
                    e = cudaGraphInstantiate(&c->g_exec, graph, 0);
                    ok = e == cudaSuccess;
                }
                if (graph) cudaGraphDestroy(graph);
                if (ok) ok = cudaGraphLaunch(c->g_exec, st) == cudaSuccess;
                how = ok ? "capture" : "capture-failed";
                if (!ok) {
                    cudaGetLastError();
                    if (c->g_exec) { cudaGraphExecDestroy(c->g_exec); c->g_exec = nullptr; }
                    c->g_disabled = true;
                    rc = enqueue(bytes);
                    if (rc) return rc;
                }
            } else {
                rc = enqueue(bytes);
                if (rc) return rc;
            }
        }
        ++attempts;
        if (hdbg) th1 = std::chrono::steady_clock::now();
        THETA_CUDA(cudaStreamSynchronize(st));
        if (hdbg) th2 = std::chrono::steady_clock::now();
        if (c->h_err && c->peer_ok && *c->h_err) { set_error("peer exchange timed out (a rank did not reach the step)"); return THETA_ERR_CUDA; }
        if (debug) {
            float t[4];
            for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], c->dbg[i], c->dbg[i + 1]);
            fprintf(stderr, "[theta dist rank %d/%d] pass1+reduce %.3f  allreduce %.3f  scan+pass2 %.3f  allgather+D2H %.3f ms\n",
                    c->rank, c->world, t[0], t[1], t[2], t[3]);
        }
        uint64_t worst = 0;
        for (int r = 0; r < c->world; ++r) {
            const unsigned long long* stw = reinterpret_cast<const unsigned long long*>(c->h_recv + (size_t)r * bytes);
            worst = std::max<uint64_t>(worst, stw[kStEmitted]);
        }
        if (worst <= c->xcap) break;
        if (attempt >= 4) { set_error("record exchange overflow (%llu records in one shard)", (unsigned long long)worst); return THETA_ERR_OVERFLOW; }
        c->xcap = worst + 64;  // every rank sees the same counts and takes the same branch
    }
    const size_t bytes = head + sizeof(theta_cand) * c->xcap;
    auto tp = std::chrono::steady_clock::now();
    double tpart[4] = {0, 0, 0, 0};
    auto lap = [&](int i) {
        if (!hdbg) return;
        const auto now = std::chrono::steady_clock::now();
        tpart[i] = std::chrono::duration<double, std::micro>(now - tp).count();
        tp = now;
    };
    rc = theta_search_n2_async_done(pl, reinterpret_cast<const unsigned long long*>(c->h_recv + (size_t)c->rank * bytes));
    if (rc) return rc;
    lap(0);
    std::vector<theta_cand>& all = c->all;
    all.clear();
    double g = INFINITY;
    uint64_t solved = 0;
    for (int r = 0; r < c->world; ++r) {
        const unsigned char* blk = c->h_recv + (size_t)r * bytes;
        const unsigned long long* stw = reinterpret_cast<const unsigned long long*>(blk);
        solved += stw[1];
        double gr;
        std::memcpy(&gr, &stw[18], sizeof(double));  // kStGlobalMin: the all-reduced minimum (same in every block)
        g = std::fmin(g, gr);
        const theta_cand* recs = reinterpret_cast<const theta_cand*>(blk + head);
        all.insert(all.end(), recs, recs + stw[kStEmitted]);
    }
    lap(1);
    uint64_t n = 0;
    rc = theta_filter_records(all.data(), all.size(), INFINITY, g + 2.5 * THETA_TIE_MARGIN, &n);
    if (rc) return rc;
    lap(2);
    if (global_min) *global_min = g;
    if (n_solved_total) *n_solved_total = solved;
    *n_out = n;
    if (n > cap) { set_error("%llu records, the caller's buffer holds %llu", (unsigned long long)n, (unsigned long long)cap); return THETA_ERR_OVERFLOW; }
    if (out_host && n) std::memcpy(out_host, all.data(), sizeof(theta_cand) * n);
    lap(3);
    if (hdbg) {
        const auto th3 = std::chrono::steady_clock::now();
        auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::micro>(b - a).count();
        };
        theta_search_stats sst{};
        theta_search_stats_get(pl, &sst);
        fprintf(stderr, "[theta dist host %d/%d] %s%s x%d: enqueue %.1f us, wait %.1f us, post %.1f us (stats %.1f, gather %.1f [%zu records], filter %.1f, copy %.1f); pass1 %.1f us, pass2 %.1f us (events)\n",
                c->rank, c->world, how, (c->world > 1 && c->peer_ok) ? "+peer" : "", attempts, us(th0, th1), us(th1, th2), us(th2, th3),
                tpart[0], tpart[1], all.size(), tpart[2], tpart[3], sst.ms_pass1 * 1e3, sst.ms_pass2 * 1e3);
    }
    return THETA_OK;
}

// One sharded n=3 search step on this rank: exact minimum of the shard, all-reduce(min) (with `extra_min`, the NLL of
// the all-[tau,0,0] pseudo candidate, folded in), window records of every shard gathered in shard (= enumeration) order.
extern "C" int theta_search_n3_dist(theta_plan* pl, theta_comm* c, uint64_t item_begin, uint64_t item_end, double window,
                                    double extra_min, theta_cand* out_host, uint8_t* paths_host, uint64_t cap,
                                    uint64_t* n_out, double* global_min, uint64_t* totals /* [2]: solved, candidates */,
                                    void* stream) {
    if (!pl || !c || !n_out) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_out = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t m = (size_t)pl->m;
    double lm = INFINITY;
    int rc = theta_search_n3_min(pl, item_begin, item_end, INFINITY, window, &lm, stream);
    if (rc) return rc;
    double g = lm;
    if (c->comm) {
        c->h_min[0] = lm;
        THETA_CUDA(cudaMemcpyAsync(c->d_min, c->h_min, sizeof(double), cudaMemcpyHostToDevice, st));
        THETA_NCCL(nccl().AllReduce(c->d_min, c->d_min, 1, ncclDouble, ncclMin, c->comm, st));
        THETA_CUDA(cudaMemcpyAsync(c->h_min, c->d_min, sizeof(double), cudaMemcpyDeviceToHost, st));
        THETA_CUDA(cudaStreamSynchronize(st));
        g = c->h_min[0];
    }
    if (extra_min == extra_min && extra_min < g) g = extra_min;
    std::vector<theta_cand> recs(256);
    std::vector<uint8_t> paths(256 * m);
    uint64_t n = 0;
    rc = theta_search_n3_window(pl, g, window, recs.data(), recs.size(), &n, paths.data(), stream);
    if (rc == THETA_ERR_OVERFLOW && n > recs.size()) {
        recs.resize(n); paths.resize(n * m);
        rc = theta_search_n3_window(pl, g, window, recs.data(), recs.size(), &n, paths.data(), stream);
    }
    if (rc) return rc;
    theta_search_stats stt;
    theta_search_stats_get(pl, &stt);
    std::vector<theta_cand> all;
    std::vector<uint8_t> allp;
    uint64_t tot[2] = {stt.n_solved, stt.n_candidates};
    if (c->comm) {
        // counts first (and the per-shard totals), then one all-gather sized for the largest shard
        unsigned long long* hh = reinterpret_cast<unsigned long long*>(c->h_recv);
        hh[0] = n; hh[1] = stt.n_solved; hh[2] = stt.n_candidates;
        THETA_CUDA(cudaMemcpyAsync(c->d_send, hh, 24, cudaMemcpyHostToDevice, st));
        THETA_NCCL(nccl().AllGather(c->d_send, c->d_recv, 24, ncclUint8, c->comm, st));
        THETA_CUDA(cudaMemcpyAsync(c->h_recv, c->d_recv, 24 * (size_t)c->world, cudaMemcpyDeviceToHost, st));
        THETA_CUDA(cudaStreamSynchronize(st));
        std::vector<uint64_t> cnt(c->world);
        uint64_t mx = 0;
        tot[0] = tot[1] = 0;
        for (int r = 0; r < c->world; ++r) {
            const unsigned long long* q = reinterpret_cast<const unsigned long long*>(c->h_recv + 24 * (size_t)r);
            cnt[r] = q[0]; tot[0] += q[1]; tot[1] += q[2];
            mx = std::max<uint64_t>(mx, q[0]);
        }
        if (mx) {
            const size_t per = sizeof(theta_cand) + m, bytes = (size_t)mx * per;
            rc = comm_buffers(c, std::max<size_t>(bytes, c->block_bytes));
            if (rc) return rc;
            std::vector<unsigned char> pack(bytes, 0);
            std::memcpy(pack.data(), recs.data(), sizeof(theta_cand) * n);
            std::memcpy(pack.data() + sizeof(theta_cand) * mx, paths.data(), n * m);
            THETA_CUDA(cudaMemcpyAsync(c->d_send, pack.data(), bytes, cudaMemcpyHostToDevice, st));
            THETA_NCCL(nccl().AllGather(c->d_send, c->d_recv, bytes, ncclUint8, c->comm, st));
            THETA_CUDA(cudaMemcpyAsync(c->h_recv, c->d_recv, bytes * c->world, cudaMemcpyDeviceToHost, st));
            THETA_CUDA(cudaStreamSynchronize(st));
            for (int r = 0; r < c->world; ++r) {
                const unsigned char* blk = c->h_recv + (size_t)r * bytes;
                const theta_cand* rr = reinterpret_cast<const theta_cand*>(blk);
                all.insert(all.end(), rr, rr + cnt[r]);
                allp.insert(allp.end(), blk + sizeof(theta_cand) * mx, blk + sizeof(theta_cand) * mx + cnt[r] * m);
            }
        }
    } else {
        all.assign(recs.begin(), recs.begin() + n);
        allp.assign(paths.begin(), paths.begin() + n * m);
    }
    for (size_t i = 0; i < all.size(); ++i) all[i].flags = (int32_t)i;  // index into the gathered paths
    if (global_min) *global_min = g;
    if (totals) { totals[0] = tot[0]; totals[1] = tot[1]; }
    *n_out = all.size();
    if (all.size() > cap) { set_error("%zu records, the caller's buffer holds %llu", all.size(), (unsigned long long)cap); return THETA_ERR_OVERFLOW; }
    if (out_host && !all.empty()) std::memcpy(out_host, all.data(), sizeof(theta_cand) * all.size());
    if (paths_host && !allp.empty()) std::memcpy(paths_host, allp.data(), allp.size());
    return THETA_OK;
}

// ======================================================================= one process, G GPUs
namespace {
struct Worker {
    int index = 0, device = 0;
    theta_plan* plan = nullptr;
    theta_comm* comm = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::thread th;
    // mailbox
    std::mutex mu;
    std::condition_variable cv;
    std::function<int(Worker&)> job;
    bool has_job = false, quit = false, done = false;
    int rc = 0;
    std::string err;
    // results of the last search
    std::vector<theta_cand> recs;
    std::vector<uint8_t> paths;
    uint64_t n_recs = 0, totals[2] = {0, 0};
    double gmin = INFINITY;
    float ms = 0.f;
    theta_search_stats stats{};
};
}  // namespace

struct theta_mgpu {
    int G = 0;
    std::vector<Worker*> w;
    // owned copy of the problem
    std::vector<int64_t> r, rN;
    std::vector<int32_t> lower, upper;
    theta_problem prob{};
    bool n3_prepared = false;
    std::vector<std::pair<uint64_t, uint64_t>> n3_ranges;
    uint64_t n3_budget = 0;
    bool truncated = false;
};

static void worker_main(Worker* w) {
    cudaSetDevice(w->device);
    for (;;) {
        std::function<int(Worker&)> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) return;
            job = w->job;
            w->has_job = false;
        }
        const int rc = job(*w);
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->rc = rc;
            w->err = rc ? theta_last_error() : "";
            w->done = true;
        }
        w->cv.notify_all();
    }
}

// run `job` on every worker, wait for all; the first failure is reported in the caller's thread
static int run_all(theta_mgpu* h, const std::function<int(Worker&)>& job) {
    for (Worker* w : h->w) {
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = job; w->has_job = true; w->done = false;
        w->cv.notify_all();
    }
    int rc = THETA_OK;
    for (Worker* w : h->w) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc && !rc) { rc = w->rc; set_error("device %d: %s", w->device, w->err.c_str()); }
    }
    return rc;
}

extern "C" void theta_mgpu_destroy(theta_mgpu* h) {
    if (!h) return;
    for (Worker* w : h->w) {
        if (w->th.joinable()) {
            run_all(h, [](Worker& x) {
                if (x.plan) theta_plan_destroy(x.plan);
                if (x.comm) theta_comm_destroy(x.comm);
                if (x.st) cudaStreamDestroy(x.st);
                if (x.e0) cudaEventDestroy(x.e0);
                if (x.e1) cudaEventDestroy(x.e1);
                x.plan = nullptr; x.comm = nullptr; x.st = nullptr; x.e0 = x.e1 = nullptr;
                return THETA_OK;
            });
            break;
        }
    }
    for (Worker* w : h->w) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    delete h;
}

extern "C" int theta_mgpu_create(const theta_problem* p, int32_t n_devices, const int32_t* devices, theta_mgpu** out) {
    if (!p || !out || n_devices < 1) { set_error("bad arguments"); return THETA_ERR_INVALID; }
    *out = nullptr;
    const int have = theta_device_count();
    if (have < 1) { set_error("no CUDA device available; theta_b200 has no CPU fallback"); return THETA_ERR_CUDA; }
    if (n_devices > have && !devices) { set_error("%d devices requested, %d present", n_devices, have); return THETA_ERR_INVALID; }
    theta_mgpu* h = new theta_mgpu();
    h->G = n_devices;
    h->r.assign(p->r, p->r + p->m); h->rN.assign(p->rN, p->rN + p->m);
    h->lower.assign(p->lower, p->lower + p->m); h->upper.assign(p->upper, p->upper + p->m);
    h->prob = *p;
    h->prob.r = h->r.data(); h->prob.rN = h->rN.data(); h->prob.lower = h->lower.data(); h->prob.upper = h->upper.data();
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        devs[i] = devices ? devices[i] : i;
        if (devs[i] < 0 || devs[i] >= have) { set_error("device %d out of range (have %d)", devs[i], have); delete h; return THETA_ERR_INVALID; }
    }
    std::vector<ncclComm_t> comms(n_devices, nullptr);
    if (n_devices > 1) {
        if (!nccl().ok) { set_error("libnccl.so.2 not found: multi-GPU search needs NCCL"); delete h; return THETA_ERR_UNSUPPORTED; }
        ncclResult_t r = nccl().CommInitAll(comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) { delete h; return nccl_fail(r, "ncclCommInitAll"); }
    }
    for (int i = 0; i < n_devices; ++i) {
        Worker* w = new Worker();
        w->index = i; w->device = devs[i];
        h->w.push_back(w);
        w->th = std::thread(worker_main, w);
    }
    const theta_problem* prob = &h->prob;
    int rc = run_all(h, [prob, &comms, n_devices](Worker& x) {
        THETA_CUDA(cudaFree(0));
        int rc2 = theta_plan_create(prob, &x.plan);
        if (rc2) return rc2;
        THETA_CUDA(cudaStreamCreateWithFlags(&x.st, cudaStreamNonBlocking));
        THETA_CUDA(cudaEventCreate(&x.e0));
        THETA_CUDA(cudaEventCreate(&x.e1));
        x.comm = new theta_comm();
        x.comm->rank = x.index; x.comm->world = n_devices; x.comm->comm = comms[x.index];
        return comm_init_common(x.comm);
    });
    if (!rc && n_devices > 1 && n_devices <= kMaxPeers && peer_exchange_wanted()) {
        // every GPU maps every other one (NVSwitch); without that the NCCL collectives stay
        std::vector<theta_comm*> cs;
        for (Worker* w : h->w) cs.push_back(w->comm);
        int rcp = run_all(h, [&cs, n_devices](Worker& x) {
            for (int j = 0; j < n_devices; ++j) {
                if (j == x.index) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, x.device, cs[j]->device) != cudaSuccess || !can) return THETA_ERR_UNSUPPORTED;
                const cudaError_t e = cudaDeviceEnablePeerAccess(cs[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return THETA_ERR_UNSUPPORTED; }
                cudaGetLastError();
            }
            return THETA_OK;
        });
        if (rcp == THETA_OK) {
            for (theta_comm* c : cs) {
                for (int j = 0; j < n_devices; ++j) { c->peers.area[j] = cs[j]->x_area; c->peers.recv[j] = cs[j]->x_recv; }
                c->peer_ok = true;
            }
        }
    }
    if (rc) { const std::string keep = theta_last_error(); theta_mgpu_destroy(h); set_error("%s", keep.c_str()); return rc; }
    *out = h;
    return THETA_OK;
}

extern "C" int theta_mgpu_devices(const theta_mgpu* h) { return h ? h->G : 0; }

// device time of the last search on each GPU (CUDA events on its stream)
extern "C" int theta_mgpu_last_ms(const theta_mgpu* h, float* ms) {
    if (!h || !ms) { set_error("null argument"); return THETA_ERR_INVALID; }
    for (int i = 0; i < h->G; ++i) ms[i] = h->w[i]->ms;
    return THETA_OK;
}

static double cluster_top(std::vector<double> v, double g) {
    std::sort(v.begin(), v.end());
    double h = g;
    for (double x : v) {
        if (x < g) continue;
        if (std::fabs(x - h) > THETA_TIE_MARGIN) break;
        h = x;
    }
    return h;
}

// do_optimization (RunTHetA.py:124-171) on all GPUs of the handle in one blocking call.
//   budget      n=2: rank limit (0 = whole space); n=3: item budget for theta_n3_prepare (0 = whole space; a truncated
//               space is an error unless a budget was given)
//   rows        [cap_best][m]: n=2 the tumour column, n=3 the row indices (theta_n3_rows) of each solution;
//               rank == UINT64_MAX marks the all-[tau,0,0] pseudo candidate of n=3 (Q2)
// Solutions come back in the reference's fold order.  *n_best may exceed cap_best (THETA_ERR_OVERFLOW).
extern "C" int theta_mgpu_optimize(theta_mgpu* h, uint64_t budget, int32_t* lower_out, int32_t* upper_out, uint64_t cap_best,
                                   uint64_t* n_best, theta_cand* best, uint8_t* rows, double* vals, theta_search_stats* stats,
                                   uint64_t* n_candidates) {
    if (!h || !n_best || (cap_best && (!best || !rows || !vals))) { set_error("null argument"); return THETA_ERR_INVALID; }
    *n_best = 0;
    const int G = h->G, m = h->prob.m;
    theta_plan* p0 = h->w[0]->plan;
    if (lower_out) std::memcpy(lower_out, p0->lower.data(), sizeof(int32_t) * m);
    if (upper_out) std::memcpy(upper_out, p0->upper.data(), sizeof(int32_t) * m);
    int rc = THETA_OK;
    if (h->prob.n == 2) {
        const uint64_t total = budget ? std::min<uint64_t>(p0->count, budget) : p0->count;
        if (n_candidates) *n_candidates = total;
        if (total == 0) { set_error("No valid Copy Number Profiles exist for these intervals within the bounds specified"); return THETA_ERR_EMPTY; }
        static const bool tdbg = getenv("THETA_DEBUG_HOST") != nullptr;
        const auto t_main0 = std::chrono::steady_clock::now();
        rc = run_all(h, [total, G, t_main0](Worker& x) {
            const auto tw0 = std::chrono::steady_clock::now();
            const uint64_t b = (uint64_t)(((unsigned __int128)total * x.index) / G), e = (uint64_t)(((unsigned __int128)total * (x.index + 1)) / G);
            if (x.recs.size() < 1024) x.recs.resize(1024);
            cudaEventRecord(x.e0, x.st);
            int rc2 = theta_search_n2_dist(x.plan, x.comm, b, e, x.recs.data(), x.recs.size(), &x.n_recs, &x.gmin, &x.totals[0], x.st);
            if (rc2 == THETA_ERR_OVERFLOW && x.n_recs > x.recs.size()) {  // a long tie list: same on every device
                x.recs.resize(x.n_recs + 16);
                rc2 = theta_search_n2_dist(x.plan, x.comm, b, e, x.recs.data(), x.recs.size(), &x.n_recs, &x.gmin, &x.totals[0], x.st);
            }
            cudaEventRecord(x.e1, x.st);
            cudaEventSynchronize(x.e1);
            cudaEventElapsedTime(&x.ms, x.e0, x.e1);
            theta_search_stats_get(x.plan, &x.stats);
            if (tdbg) {
                const auto tw1 = std::chrono::steady_clock::now();
                fprintf(stderr, "[theta mgpu worker %d] wake %.1f us after post, job %.1f us (device %.1f us)\n", x.index,
                        std::chrono::duration<double, std::micro>(tw0 - t_main0).count(),
                        std::chrono::duration<double, std::micro>(tw1 - tw0).count(), x.ms * 1e3);
            }
            return rc2;
        });
        if (tdbg) fprintf(stderr, "[theta mgpu main] run_all %.1f us\n",
                          std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_main0).count());
        if (rc) return rc;
        Worker& w0 = *h->w[0];
        if (stats) {
            *stats = w0.stats;
            stats->n_candidates = total;
            stats->n_solved = w0.totals[0];
            for (Worker* w : h->w) stats->ms_pass1 = std::max(stats->ms_pass1, w->stats.ms_pass1);
            stats->min_nll = w0.gmin;
        }
        // the fold, verbatim (RunTHetA.py:198-206; first matrix twice: Q1)
        const theta_cand* recs = w0.recs.data();
        const uint64_t n = w0.n_recs;
        std::vector<uint32_t> keep;
        double M = INFINITY;
        for (uint64_t i = 0; i < n; ++i) {
            const int reps = recs[i].rank == 0 ? 2 : 1;
            for (int k = 0; k < reps; ++k) {
                const double L = recs[i].nll;
                if (!(std::fabs(L - M) > THETA_TIE_MARGIN)) keep.push_back((uint32_t)i);
                else if (L < M) { keep.assign(1, (uint32_t)i); M = L; }
            }
        }
        *n_best = keep.size();
        if (keep.size() > cap_best) { set_error("%zu solutions tie, the caller's buffers hold %llu", keep.size(), (unsigned long long)cap_best); return THETA_ERR_OVERFLOW; }
        std::vector<uint64_t> ranks(keep.size());
        for (size_t k = 0; k < keep.size(); ++k) { best[k] = recs[keep[k]]; ranks[k] = best[k].rank; }
        if (!keep.empty()) {
            std::vector<int8_t> cols(keep.size() * (size_t)m);
            rc = theta_columns_of_ranks(p0, ranks.data(), (int64_t)ranks.size(), cols.data());
            if (rc) return rc;
            std::memcpy(rows, cols.data(), cols.size());
            for (size_t k = 0; k < keep.size(); ++k) {  // vals_i = numer_i / denom of Optimizer.L2 (Optimizer.py:187-196)
                const int8_t* col = cols.data() + k * (size_t)m;
                double* vk = vals + k * (size_t)m;
                const double mu = best[k].mu[0], mu1 = 1.0 - mu;
                double S1 = 0.0;
                for (int i = 0; i < m; ++i) S1 += (double)p0->rN[i] * (double)col[i];
                const double D = mu * p0->kc.S0 + mu1 * S1;
                for (int i = 0; i < m; ++i) vk[i] = (double)p0->rN[i] * (p0->kc.tau * mu + (double)col[i] * mu1) / D;
            }
        }
        return THETA_OK;
    }
    // ---- n = 3
    if (!h->n3_prepared || h->n3_budget != budget) {
        rc = run_all(h, [budget](Worker& x) {
            uint64_t ni = 0;
            return theta_n3_prepare(x.plan, 0, budget ? (int64_t)budget : 0, &ni);
        });
        if (rc) return rc;
        int32_t tr = 0;
        theta_n3_truncated(p0, &tr);
        h->truncated = tr != 0;
        if (tr && !budget) {
            set_error("the n=3 candidate space exceeds the work-item limit: the search would cover only a lexicographic prefix of it "
                      "(pass an explicit item budget to search a prefix)");
            return THETA_ERR_OVERFLOW;
        }
        uint64_t ni = 0;
        theta_n3_item_count(p0, &ni);
        std::vector<double> wts(ni);
        if (ni) theta_n3_item_weights(p0, wts.data());
        std::vector<double> cum(ni + 1, 0.0);
        for (uint64_t i = 0; i < ni; ++i) cum[i + 1] = cum[i] + wts[i];
        h->n3_ranges.assign(G, {0, 0});
        for (int g = 0; g < G; ++g) {
            const uint64_t lo = (uint64_t)(std::lower_bound(cum.begin(), cum.end(), cum[ni] * g / G) - cum.begin());
            const uint64_t hi = g + 1 < G ? (uint64_t)(std::lower_bound(cum.begin(), cum.end(), cum[ni] * (g + 1) / G) - cum.begin()) : ni;
            h->n3_ranges[g] = {std::min(lo, ni), std::min(std::max(hi, lo), ni)};
        }
        h->n3_prepared = true; h->n3_budget = budget;
    }
    // the all-[tau,0,0] pseudo candidate (RunTHetA.py:148,188 -- Q2): row 0 is (0, 0); reference-order solve
    std::vector<uint8_t> zero(m, 0);
    double bmu[3], bmh[3], bnll = 0.0;
    int32_t bst = 0;
    std::vector<double> bvals(m);
    int caller_dev = 0;
    cudaGetDevice(&caller_dev);
    struct DevGuard { int d; ~DevGuard() { cudaSetDevice(d); } } guard{caller_dev};
    cudaSetDevice(h->w[0]->device);
    rc = theta_solve_rows_n3(p0, zero.data(), 1, bmu, bmh, &bnll, &bst, bvals.data(), 1);
    if (rc) return rc;
    const bool has_bogus = bst < THETA_ST_NONE_SIGN;
    const double null_nll = has_bogus ? bnll : INFINITY;
    double window = 4.0 * THETA_TIE_MARGIN;
    Worker& w0 = *h->w[0];
    for (int attempt = 0;; ++attempt) {
        const auto& ranges = h->n3_ranges;
        rc = run_all(h, [&ranges, window, null_nll, m](Worker& x) {
            if (x.recs.size() < 256) { x.recs.resize(256); x.paths.resize(256 * (size_t)m); }
            cudaEventRecord(x.e0, x.st);
            int rc2 = theta_search_n3_dist(x.plan, x.comm, ranges[x.index].first, ranges[x.index].second, window, null_nll,
                                           x.recs.data(), x.paths.data(), x.recs.size(), &x.n_recs, &x.gmin, x.totals, x.st);
            if (rc2 == THETA_ERR_OVERFLOW && x.n_recs > x.recs.size()) {
                x.recs.resize(x.n_recs + 16); x.paths.resize((x.n_recs + 16) * (size_t)m);
                rc2 = theta_search_n3_dist(x.plan, x.comm, ranges[x.index].first, ranges[x.index].second, window, null_nll,
                                           x.recs.data(), x.paths.data(), x.recs.size(), &x.n_recs, &x.gmin, x.totals, x.st);
            }
            cudaEventRecord(x.e1, x.st);
            cudaEventSynchronize(x.e1);
            cudaEventElapsedTime(&x.ms, x.e0, x.e1);
            theta_search_stats_get(x.plan, &x.stats);
            return rc2;
        });
        if (rc) return rc;
        std::vector<double> fin;
        for (uint64_t i = 0; i < w0.n_recs; ++i) if (w0.recs[i].nll == w0.recs[i].nll) fin.push_back(w0.recs[i].nll);
        if (has_bogus && null_nll == null_nll) fin.push_back(null_nll);
        if (fin.empty()) break;
        const double g_ref = *std::min_element(fin.begin(), fin.end());
        if (cluster_top(fin, g_ref) + THETA_TIE_MARGIN < w0.gmin + window - 1e-6) break;
        if (attempt >= 8) { set_error("n=3 tie cluster does not end inside a window of %g", window); return THETA_ERR_OVERFLOW; }
        window *= 8.0;
    }
    if (n_candidates) *n_candidates = w0.totals[1] + 1;
    if (stats) {
        *stats = w0.stats;
        stats->n_candidates = w0.totals[1] + 1;
        stats->n_solved = w0.totals[0] + 1;
        stats->min_nll = w0.gmin;
        for (Worker* w : h->w) stats->ms_pass1 = std::max(stats->ms_pass1, w->stats.ms_pass1);
    }
    // fold over: pseudo candidate first, then the records in enumeration order
    struct Ent { double nll; int64_t idx; };
    std::vector<Ent> stream;
    if (has_bogus) stream.push_back({null_nll, -1});
    for (uint64_t i = 0; i < w0.n_recs; ++i) stream.push_back({w0.recs[i].nll, (int64_t)i});
    std::vector<int64_t> keep;
    double M = INFINITY;
    for (const Ent& e : stream) {
        if (!(std::fabs(e.nll - M) > THETA_TIE_MARGIN)) keep.push_back(e.idx);
        else if (e.nll < M) { keep.assign(1, e.idx); M = e.nll; }
    }
    *n_best = keep.size();
    if (keep.size() > cap_best) { set_error("%zu solutions tie, the caller's buffers hold %llu", keep.size(), (unsigned long long)cap_best); return THETA_ERR_OVERFLOW; }
    std::vector<uint8_t> kp;
    for (size_t k = 0; k < keep.size(); ++k) {
        if (keep[k] < 0) {
            theta_cand c{};
            c.rank = UINT64_MAX; c.nll = null_nll; c.mu[0] = bmu[0]; c.mu[1] = bmu[1]; c.mu[2] = bmu[2]; c.status = bst;
            best[k] = c;
            std::memset(rows + k * (size_t)m, 0, m);
        } else {
            best[k] = w0.recs[keep[k]];
            std::memcpy(rows + k * (size_t)m, w0.paths.data() + (size_t)w0.recs[keep[k]].flags * m, m);
        }
        kp.insert(kp.end(), rows + k * (size_t)m, rows + (k + 1) * (size_t)m);
    }
    if (!keep.empty()) {  // Optimizer.solve's vals of the winners (reference order)
        std::vector<double> mu(3 * keep.size()), nl(keep.size());
        std::vector<int32_t> stt(keep.size());
        rc = theta_solve_rows_n3(p0, kp.data(), (int64_t)keep.size(), mu.data(), nullptr, nl.data(), stt.data(), vals, 1);
        if (rc) return rc;
    }
    return THETA_OK;
}

// One-shot form: create, optimize, destroy.
extern "C" int theta_optimize(const theta_problem* p, int32_t n_devices, const int32_t* devices, uint64_t budget,
                              int32_t* lower_out, int32_t* upper_out, uint64_t cap_best, uint64_t* n_best, theta_cand* best,
                              uint8_t* rows, double* vals, theta_search_stats* stats, uint64_t* n_candidates) {
    theta_mgpu* h = nullptr;
    int rc = theta_mgpu_create(p, n_devices, devices, &h);
    if (rc) return rc;
    rc = theta_mgpu_optimize(h, budget, lower_out, upper_out, cap_best, n_best, best, rows, vals, stats, n_candidates);
    const std::string keep = rc ? theta_last_error() : "";
    theta_mgpu_destroy(h);
    if (rc) set_error("%s", keep.c_str());
    return rc;
}
