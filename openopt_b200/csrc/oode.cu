// liboode: C ABI (include/oode.h) over the sm_100a generation kernels (de_kernels.cuh).
// Host side: handle lifetime, the reduction plan (numpy's pairwise tree), draw upload for the
// replay modes, kernel launches on the handle's stream, history ring read-back.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <chrono>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>      // header-only; ranges are no-ops unless a profiler is attached

#include "de_host_kernels.cuh"
#include "de_launch.h"
#include "de_rows.cuh"
#include "mt19937_cpython.h"

using namespace oode;

static thread_local std::string g_err;

// NVTX range over a C-ABI call (shows up in ncu / nsys timelines as oode_create, oode_init, oode_step)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// OODE_TIMING=1: host-side phase timings of create / init / destroy on stderr (diagnostic)
struct PhaseTimer {
    bool on;
    std::chrono::steady_clock::time_point t;
    const char *what;
    explicit PhaseTimer(const char *w) : on(getenv("OODE_TIMING") != nullptr), t(std::chrono::steady_clock::now()), what(w) {}
    void mark(const char *label) {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[oode timing] %s: %-28s %8.3f ms\n", what, label, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(OODE_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL is loaded at run time (dlopen), and only when a handle is created with world_size > 1, so
// the single-GPU path has no dependency on it.  In a process that has imported torch this picks up
// the libnccl.so.2 torch already loaded.
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define OODE_NCCL_SYM(field, name)                                             \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, name));               \
    if (!field) { err = std::string("libnccl lacks ") + name; lib = nullptr; return false; }
        OODE_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        OODE_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        OODE_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        OODE_NCCL_SYM(AllGather, "ncclAllGather")
        OODE_NCCL_SYM(AllReduce, "ncclAllReduce")
        OODE_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef OODE_NCCL_SYM
        return true;
    }
};
static NcclApi g_nccl;

// Communicators are cached for the life of the process, keyed by the id the ranks agreed on: a
// second solve over the same ranks costs no NCCL bootstrap (seconds at 8 GPUs).  The peer-exchange
// arena (this rank's exchange buffer and the mappings of everybody else's) lives with the
// communicator for the same reason: unmapping CUDA-IPC memory costs ~30 ms per peer.
//   arena = [X_TAIL_WORDS u64: epoch flags 0..7, magic 8, wait error 9][2][G][chunk] doubles
constexpr int OODE_MAX_PEERS = 8;
constexpr int X_TAIL_WORDS = 16;
constexpr unsigned long long X_MAGIC = 0x0DE0C0DE00000000ull;
struct Arena {
    double *base = nullptr;                 // this rank's arena
    size_t doubles = 0;                     // capacity behind the tail words
    double *peer[OODE_MAX_PEERS] = {};      // every rank's arena as mapped here (own entry = base)
    unsigned long long epoch = 0;           // flags only ever grow, across handles too
    bool failed = false;                    // CUDA IPC did not work between these ranks: use NCCL
    bool in_use = false;                    // one live handle at a time
};
struct PeerImport {                         // a peer's device block mapped here through CUDA IPC
    cudaIpcMemHandle_t handle;
    int peer;
    void *ptr;
};
struct CommEntry {
    unsigned char id[128];
    int world, rank, device;
    ncclComm_t comm;
    Arena arena;
    std::vector<PeerImport> imports;        // kept for the life of the communicator (unmapping is slow)
};
static std::vector<CommEntry *> g_comms;
static std::mutex g_cache_mutex;      // guards g_comms and g_blocks (handles may live on different host threads)

#define NC(call)                                                                                     \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess)                                                                      \
            return fail(OODE_ENCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r__));         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Reduction plan: leaves and post-order fold program of numpy's pairwise sum over n terms
// (oracle/de_oracle.py::pairwise_leaves).
// ---------------------------------------------------------------------------------------------
struct ReductionPlan {
    std::vector<int2> leaves;
    std::vector<int> prog;
    void build(int n, int off) {
        if (n <= 128) {
            prog.push_back((int)leaves.size());
            leaves.push_back(make_int2(off, n));
            return;
        }
        int n2 = n / 2;
        n2 -= n2 % 8;
        build(n2, off);
        build(n - n2, off + n2);
        prog.push_back(-1);
    }
};

struct CachedBlock {
    int device;
    size_t bytes;
    void *ptr;
};

struct oode_handle {
    oode_config cfg;
    ObjEntry obj{};            // the objective's kernels (registry entry at creation time)
    int D = 0, Dl = 0, col0 = 0, nT = 0;
    int64_t NP = 0, ld = 0;
    int64_t row0 = 0, nrows = 0;
    int nslots = 1, halo = 0;
    int ring = 0;
    bool have_x0 = false;
    bool inited = false;
    int64_t gen = 0;
    ReductionPlan plan;        // leaves (local column offsets) this rank computes
    ReductionPlan gplan;       // the whole row's tree (fold program, global leaf numbering)
    PhiloxKeys keys;
    // sharding (one process per GPU).  W = ranks of the job; G = column shards (1 on a row-sharded handle)
    int W = 1;
    bool rows = false;                        // OODE_SHARD_ROWS: this rank builds the trials of rows [row0, row0+nrows)
    bool rows_fused = false;                  // ... and its trial kernel pushes the accepted rows itself (RowPush)
    double *peer_buf[2][MAX_PEERS] = {};      // row shards: every rank's slot arrays as mapped here (own = local)
    int64_t mb_doubles = 0;                   // row shards: doubles per epoch mailbox
    int ldc = 0;                              // pitch of the candidate rows in a mailbox
    int G = 1, rank = 0, shard = OODE_SHARD_NONE;
    int leaf0 = 0, Lmax = 0;   // first global leaf owned; widest shard's leaf count
    int64_t chunk = 0;         // doubles per rank in the all-gathered partials (incl. 1 counter slot)
    std::vector<int> shard_col0, shard_dl;   // every rank's column slice
    ncclComm_t comm = nullptr;   // from the process-wide cache (comm_for); never destroyed by the handle
    CommEntry *ce = nullptr;
    double *partials_all = nullptr;
    // peer exchange (leaf sums stored straight into every rank's copy over NVLink)
    bool p2p = false;
    unsigned int *xticket = nullptr;           // finished CTAs of the running trial kernel (peer_publish_tail)
    unsigned long long *wait_ns = nullptr;     // time K2b's block 0 spent waiting for the other ranks' flags
    unsigned long long wait_epoch = 0;
    Arena *arena = nullptr;                    // the communicator's arena while this handle holds it
    double *xdata = nullptr;                   // local arena data: [2][G][chunk] doubles
    unsigned long long *x_tail = nullptr;      // local tail words: [0..7] epoch flags, [8] magic, [9] wait error
    int *leaf_rank = nullptr, *leaf_li = nullptr;
    int chunks = 1;            // row chunks of the partial-sum exchange (overlapped with K2a)
    int64_t chunk_rows = 0;
    cudaStream_t comm_stream = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    cudaEvent_t comm_done = nullptr;
    double *bx_send = nullptr, *bx_all = nullptr;
    int dl_max = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool profiling = false;
    std::vector<cudaEvent_t> pev;      // 4 events per generation of a batch when profiling
    cudaEvent_t *xev = nullptr;        // where run_trial records "trial kernel done, exchange starts" (peer exchange)
    int sm_count = 148;
    int trial_ctas_per_sm = 2;
    bool uniform_bounds = false;
    int form = FORM_QUAD;              // K2a form (de_launch.h): row-walking TMA kernel, quad-per-leaf, or a round-1 TMA form
    int wpw = 1, warps = 1, rs = 2;    // FORM_ROWS: workers per warp, warps per CTA, doubles per staged input row
    bool fast = false;                 // FORM_ROWS: |F| < 1 and small cos arguments -> the FAST instantiation is valid ...
    bool trusted_box = true;           // ... as long as every row is inside [lb, ub] (false after oode_set_population)
    bool cos_checked = true;           // false: the bounds guarantee |cos argument| <= 1e6 for every gene
    double lo_s = 0, hi_s = 0;
    int cur = 0;   // index of the current selector array
    // device memory
    double *buf[2] = {nullptr, nullptr};
    uint8_t *sel[2] = {nullptr, nullptr};
    uint8_t *accept = nullptr;
    double *vals = nullptr, *tvals = nullptr;
    int32_t *ib = nullptr, *r1 = nullptr, *r2 = nullptr;      // r1, r2: [hndvi][NP]
    // strategy (de_oo.py:71-89)
    int hndvi = 1;
    bool base_best = false, dir_best = false, const_F = false;
    // general constraints (oode_set_constraints)
    bool constrained = false;
    ConstrParams cpar{};           // device matrices / right-hand sides
    double contol = 0;
    double *cvals = nullptr, *tcvals = nullptr;   // [NP] constr_val of the current / the trial population
    int *blk_cls = nullptr;
    double *tbest_x = nullptr;     // [ld] trial-best row of the last evaluated population (base 'best')
    double *dvec = nullptr;        // [2][ld] worst / best barycentre of the directed search
    int32_t *rdir = nullptr;       // [2h] directed-search indices supplied by the caller / CPython stream
    double *lb = nullptr, *ub = nullptr, *aux = nullptr, *raux = nullptr, *x0 = nullptr;
    int2 *leaves = nullptr;
    int *prog = nullptr;
    double *partials = nullptr;
    double *Uf = nullptr, *Uc = nullptr;
    double *blk_f = nullptr;
    int64_t *blk_i = nullptr;
    unsigned long long *blk_acc = nullptr;
    unsigned int *ticket = nullptr;
    unsigned long long *n_rep = nullptr;
    unsigned long long *zero_word = nullptr;
    double *best_f = nullptr, *best_x = nullptr, *bestx_ring = nullptr;
    oode_gen_record *rec = nullptr;
    // host staging
    std::vector<int32_t> h_idx;
    MT19937CPython *mt = nullptr;
    std::vector<double> h_u;
    // counters
    oode_counters ctr{};
    std::vector<CachedBlock> cached;    // device blocks; they go back to the process cache on destroy
    // scratch of the verification taps (oode_eval_points, oode_get_rows): kept on the handle, grown to the
    // largest request seen, so no call allocates or frees (and no error path can leak)
    struct Scratch { void *ptr = nullptr; size_t bytes = 0; } scratch[4];
    bool fold_fast = false;            // the reduction tree folds level-wise in registers (PartLayout::fast_leaves)
    int8_t fast_rank[8] = {}, fast_li[8] = {};
    int64_t min_gen = 0;               // oldest generation whose Best.x snapshot is valid (raised by oode_load_state)
};

// Device blocks are kept by the process when a handle is destroyed and handed to the next handle that
// asks for the same size on the same device: cudaMalloc/cudaFree (multi-GB blocks, and any block once
// peer access is enabled) cost far more than the generations of a short solve -- measured up to 100 ms
// per handle.  oode_release_cached_memory() returns them to the driver; an out-of-memory cudaMalloc
// does so by itself and retries.
static std::vector<CachedBlock> g_blocks;

static void release_cached_blocks() {     // caller holds g_cache_mutex
    int cur = 0;
    cudaGetDevice(&cur);
    for (const CachedBlock &b : g_blocks) {
        cudaSetDevice(b.device);
        cudaFree(b.ptr);
    }
    g_blocks.clear();
    cudaSetDevice(cur);
}

template <typename T> static int dalloc(oode_handle *h, T **p, size_t n) {
    void *q = nullptr;
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    {
        for (size_t i = 0; i < g_blocks.size(); ++i)
            if (g_blocks[i].device == h->cfg.device && g_blocks[i].bytes == bytes) {
                q = g_blocks[i].ptr;
                g_blocks.erase(g_blocks.begin() + i);
                break;
            }
    }
    if (!q) {
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e == cudaErrorMemoryAllocation && !g_blocks.empty()) {
            cudaGetLastError();
            release_cached_blocks();
            e = cudaMalloc(&q, bytes);
        }
        if (e != cudaSuccess) return fail(OODE_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    h->cached.push_back({h->cfg.device, bytes, q});
    *p = (T *)q;
    return OODE_OK;
}
#define DA(ptr, n)                                \
    do {                                          \
        int rc__ = dalloc(h, &(ptr), (size_t)(n)); \
        if (rc__) return rc__;                    \
    } while (0)

template <typename T> static int scratch_for(oode_handle *h, int slot, T **p, size_t n) {
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    oode_handle::Scratch &sc = h->scratch[slot];
    if (sc.bytes < bytes) {
        size_t want = std::max(bytes, sc.bytes * 2);      // the outgrown block stays with the handle until destroy
        unsigned char *q = nullptr;
        int rc = dalloc(h, &q, want);
        if (rc) return rc;
        sc.ptr = q;
        sc.bytes = want;
    }
    *p = (T *)sc.ptr;
    return OODE_OK;
}
#define SCRATCH(slot, ptr, n)                                      \
    do {                                                           \
        int rc__ = scratch_for(h, slot, &(ptr), (size_t)(n));      \
        if (rc__) return rc__;                                     \
    } while (0)

static void obj_traits(int obj, int *nslots, int *halo);

// ---------------------------------------------------------------------------------------------
// kernel dispatch: the objective picks the translation unit (de_launch.h), the handle the kernel form
// ---------------------------------------------------------------------------------------------
// Objective registry: the six built-ins, then whatever oode_register_objective has loaded.
#define OODE_BUILTIN(k, nsum, nprod, halo, aux) \
    ObjEntry{OODE_ABI_VERSION, (int)sizeof(GenParams), (int)sizeof(SelParams), nsum, nprod, halo, aux, launch_trial_obj_##k, launch_select_obj_##k, launch_finalize_obj_##k}
static std::vector<ObjEntry> g_objs = {OODE_BUILTIN(0, 1, 0, 0, 0), OODE_BUILTIN(1, 1, 0, 1, 0), OODE_BUILTIN(2, 1, 0, 0, 0),
                                       OODE_BUILTIN(3, 2, 0, 0, 0), OODE_BUILTIN(4, 1, 1, 0, 1), OODE_BUILTIN(5, 1, 0, 0, 1)};
#undef OODE_BUILTIN
static std::vector<std::string> g_obj_paths(OODE_OBJ_COUNT);
static std::mutex g_obj_mutex;

static void obj_traits(int obj, int *nslots, int *halo) {
    std::lock_guard<std::mutex> lock(g_obj_mutex);
    *nslots = g_objs[obj].nsum + g_objs[obj].nprod;
    *halo = g_objs[obj].halo;
}
static bool obj_known(int obj) {
    std::lock_guard<std::mutex> lock(g_obj_mutex);
    return obj >= 0 && obj < (int)g_objs.size();
}
static ObjEntry obj_entry(int obj) {
    std::lock_guard<std::mutex> lock(g_obj_mutex);
    return g_objs[obj];
}

// `points`: the rows are caller-supplied points (oode_eval_points), not a population inside the box
static cudaError_t launch_trial(oode_handle *h, int mode, const GenParams &p, bool points = false) {
    TrialLaunch L{};
    L.form = h->form;
    L.wpw = h->wpw;
    L.warps = h->warps;
    L.uni = h->uniform_bounds;
    L.fast = h->fast && h->trusted_box && !points;
    L.cos_checked = h->cos_checked || points;
    L.sm_count = h->sm_count;
    L.ctas_per_sm = h->trial_ctas_per_sm;
    L.stream = h->stream;
    h->ctr.gpu_launches++;
    return h->obj.trial(L, mode, p);
}

static cudaError_t launch_select(oode_handle *h, const SelParams &p) {
    h->ctr.gpu_launches++;
    return h->obj.select(h->constrained, h->stream, p);
}

// Does the post-order fold program equal the level-wise pairwise reduction over n <= 8 leaves (PartLayout::fast_leaves)?
// Both are evaluated symbolically and compared.
static bool fold_is_levelwise(const std::vector<int> &prog, int n) {
    if (n < 2 || n > 8) return false;
    std::vector<std::string> st;
    for (int v : prog) {
        if (v >= 0) st.push_back(std::to_string(v));
        else {
            if (st.size() < 2) return false;
            const std::string b = st.back(); st.pop_back();
            st.back() = "(" + st.back() + "+" + b + ")";
        }
    }
    if (st.size() != 1) return false;
    std::string v[8];
    for (int l = 0; l < n; ++l) v[l] = std::to_string(l);
    for (int w = 1; w < 8; w <<= 1)
        for (int i = 0; i + w < 8; i += 2 * w)
            if (i + w < n) v[i] = "(" + v[i] + "+" + v[i + w] + ")";
    return v[0] == st[0];
}

static PartLayout part_layout(const oode_handle *h, const double *base, int64_t nrows) {
    PartLayout L{};
    L.fast_leaves = h->fold_fast ? (int)h->gplan.leaves.size() : 0;
    for (int l = 0; l < L.fast_leaves; ++l) { L.fast_rank[l] = h->fast_rank[l]; L.fast_li[l] = h->fast_li[l]; }
    L.base = base;
    L.leaf_rank = h->leaf_rank;
    L.leaf_li = h->leaf_li;
    L.G = h->G;
    L.stride = h->Lmax * h->nslots;
    if (h->p2p && base == h->partials_all) {   // peer exchange: [G][chunk] blocks of the current epoch's buffer
        L.C = 1;
        L.R = (int)std::max<int64_t>(nrows, 1);
        L.count_full = L.count_last = h->chunk;
    } else if (base == h->partials_all) {
        L.C = h->chunks;
        L.R = (int)h->chunk_rows;
        L.count_full = h->chunk_rows * L.stride;
        L.count_last = (nrows - (int64_t)(h->chunks - 1) * h->chunk_rows) * L.stride + 1;
    } else {              // a local [nrows][Lmax][NSL] array
        L.C = 1;
        L.G = 1;
        L.R = (int)std::max<int64_t>(nrows, 1);
        L.count_full = L.count_last = nrows * L.stride;
    }
    return L;
}

// K2a over all rows.  Column shards: the rows are processed in `chunks` launches; as soon as a chunk
// is done its leaf sums are all-gathered on a second stream (the last chunk's payload also carries
// this rank's repair counter), so the exchange -- the only cross-GPU traffic of a generation --
// overlaps the next chunk's compute.  Every rank then folds the whole tree and takes the same
// selection decisions.
static int run_trial(oode_handle *h, int mode, GenParams gp) {
    if (h->G <= 1) {
        CU(launch_trial(h, mode, gp));
        return OODE_OK;
    }
    if (h->p2p) {
        // Peer exchange: ONE launch over all rows; the kernel stores every leaf sum into block `rank` of
        // every rank's copy (own: local memory, others: NVLink), so the transfer rides along with the
        // compute.  Then publish (repair counter + epoch flag on every rank) and wait for all flags.
        Arena &A = *h->arena;
        const unsigned long long e = ++A.epoch;
        const int64_t buf_off = X_TAIL_WORDS + (int64_t)(e & 1) * h->G * h->chunk;
        gp.peers.n = h->G;
        for (int g = 0; g < h->G; ++g) {
            gp.peers.part[g] = A.peer[g] + buf_off + (int64_t)h->rank * h->chunk;
            gp.peers.flags[g] = reinterpret_cast<unsigned long long *>(A.peer[g]);
        }
        {   // row chunks whose leaf sums leave as one contiguous run per peer (>= 128 bytes when a row's sums allow it)
            const int per_row = h->Lmax * h->nslots;
            gp.xrows = std::max(1, std::min(8, XBUF_DOUBLES / std::max(per_row, 1)));
            if (getenv("OODE_XCHG_UNBUFFERED")) gp.xrows = 1;
        }
        gp.peers.ticket = h->xticket;
        gp.peers.counter_slot = h->NP * (int64_t)h->Lmax * h->nslots;
        gp.peers.rank = h->rank;
        gp.peers.epoch = e;
        h->partials_all = A.base + buf_off;
        h->wait_epoch = e;                           // K2b's prologue waits for this epoch (sel_params)
        if (getenv("OODE_DEBUG_LOCAL_STORES")) {     // timing diagnostic only (results are wrong): no remote stores
            gp.peers.n = 1;
            gp.peers.part[0] = A.base + buf_off + (int64_t)h->rank * h->chunk;
            gp.peers.flags[0] = reinterpret_cast<unsigned long long *>(A.base);
        }
        // the last CTA of the trial kernel to finish publishes (repair counter + epoch flag on every rank,
        // peer_publish_tail); one warp then waits for everybody's flag in front of K2b
        CU(launch_trial(h, mode, gp));
        if (h->xev) CU(cudaEventRecord(*h->xev, h->stream));
        PeerWait w{};
        w.flags = h->x_tail;
        w.n = h->G;
        w.epoch = e;
        w.err = h->x_tail + 9;
        w.wait_ns = h->wait_ns;
        de_peer_wait_kernel<<<1, 32, 0, h->stream>>>(w);
        CU(cudaGetLastError());
        h->ctr.gpu_launches++;
        return OODE_OK;
    }
    const int64_t stride = (int64_t)h->Lmax * h->nslots;
    for (int c = 0; c < h->chunks; ++c) {
        const int64_t r0 = (int64_t)c * h->chunk_rows;
        const int64_t rows = std::min<int64_t>(h->chunk_rows, h->NP - r0);
        gp.row0 = (int)r0;
        gp.nrows = (int)rows;
        gp.partials = h->partials + r0 * stride;
        CU(launch_trial(h, mode, gp));
        size_t count = (size_t)(rows * stride);
        if (c == h->chunks - 1) {
            de_pack_counter_kernel<<<1, 1, 0, h->stream>>>(h->n_rep, h->partials + h->NP * stride);
            CU(cudaGetLastError());
            h->ctr.gpu_launches++;
            count += 1;
        }
        CU(cudaEventRecord(h->chunk_ev[c], h->stream));
        CU(cudaStreamWaitEvent(h->comm_stream, h->chunk_ev[c], 0));
        NC(g_nccl.AllGather(h->partials + r0 * stride, h->partials_all + (int64_t)h->G * r0 * stride, count, ncclFloat64,
                            h->comm, h->comm_stream));
    }
    CU(cudaEventRecord(h->comm_done, h->comm_stream));
    CU(cudaStreamWaitEvent(h->stream, h->comm_done, 0));
    return OODE_OK;
}

static GenParams gen_params(oode_handle *h, uint32_t gen) {
    GenParams p{};
    p.buf0 = h->buf[0];
    p.buf1 = h->buf[1];
    p.sel_cur = h->sel[h->cur];
    p.ld = (int)h->ld;
    p.ib = h->ib; p.r1 = h->r1; p.r2 = h->r2;
    p.lb = h->lb; p.ub = h->ub; p.aux = h->aux; p.raux = h->raux;
    p.lo_s = h->lo_s; p.hi_s = h->hi_s;
    p.Uf = h->Uf; p.Uc = h->Uc;
    p.partials = h->partials;
    p.leaves = h->leaves;
    p.nleaves = (int)h->plan.leaves.size();
    p.part_stride = h->Lmax;
    p.nT = h->nT;
    p.Dl = h->Dl;
    p.D = h->D;
    p.col0 = h->col0;
    p.row0 = (int)h->row0;
    p.nrows = (int)h->nrows;
    p.F = h->cfg.diff_factor;
    p.Cr = h->cfg.cross_rate;
    p.gen = gen;
    p.keys = h->keys;
    p.n_repaired = h->n_rep;
    p.n_repaired_zero = h->zero_word;
    p.base_vec = h->base_best ? h->tbest_x : nullptr;
    p.d1_vec = h->dir_best ? h->dvec : nullptr;
    p.d2_vec = h->dir_best ? h->dvec + (h->ld + 2) : nullptr;
    p.hndvi = h->hndvi;
    p.const_F = h->const_F ? 1 : 0;
    {   // crossover threshold on the raw Philox word (keep_parent, de_kernels.cuh)
        uint32_t t;
        int32_t all;
        oode_crossover_threshold(h->cfg.cross_rate, &t, &all);
        p.cr_thresh = t;
        p.cr_all = all;
    }
    p.NP = (int)h->NP;
    p.Fa = h->const_F ? 0.0 : h->cfg.diff_factor;
    p.Fa32 = p.Fa * (1.0 / 4294967296.0);          // exact scaling: fma(double(w), Fa32, Fb) == fma(w * 2^-32, Fa, Fb)
    p.Fb = h->const_F ? h->cfg.diff_factor : 0.0;
    p.rs = h->rs;
    p.xrows = 1;
    p.part_limit = h->chunk;                       // [NP][Lmax][NSL] (+ the counter slot)
    p.count_tail = (h->G <= 1 || h->rank == h->G - 1) ? 1 : 0;
    if (h->rows_fused) {                             // row shards, single-leaf rows: accepted rows leave from the trial kernel
        p.rpush.n = h->W;
        p.rpush.rank = h->rank;
        p.rpush.vals = h->vals;
        p.rpush.invert = h->cfg.invert;
        p.rpush.D = h->D;
        for (int g = 0; g < h->W; ++g) { p.rpush.buf0[g] = h->peer_buf[0][g]; p.rpush.buf1[g] = h->peer_buf[1][g]; }
    }
    return p;
}

extern "C" {
static RowsMailbox rows_mailbox(const oode_handle *h, double *arena_base, unsigned long long epoch);
}

static SelParams sel_params(oode_handle *h, int init, int64_t gen) {
    SelParams p{};
    p.part = part_layout(h, h->G > 1 ? h->partials_all : h->partials, h->NP);
    p.prog = h->prog;
    p.nprog = (int)h->gplan.prog.size();
    p.nleaves = (int)h->gplan.leaves.size();
    p.n_rep_parts = h->G > 1 ? h->G : 0;
    {
        const int64_t stride = (int64_t)h->Lmax * h->nslots;
        const int64_t rows_last = h->NP - (int64_t)(h->chunks - 1) * h->chunk_rows;
        p.rep_parts = h->G > 1 ? h->partials_all + (int64_t)h->G * (h->chunks - 1) * h->chunk_rows * stride + rows_last * stride
                               : nullptr;
        p.rep_stride = h->p2p ? h->chunk : rows_last * stride + 1;
    }
    p.D = h->D;
    p.Dl = h->Dl;
    p.invert = h->cfg.invert;
    p.init = init;
    p.do_indices = (h->cfg.rng == OODE_RNG_PHILOX);
    p.NP = h->NP;
    p.row0 = h->row0;
    p.nrows = h->nrows;
    p.vals = h->vals;
    p.tvals = h->tvals;
    p.sel_cur = h->sel[h->cur];
    p.sel_next = h->sel[h->cur ^ 1];
    p.accept = h->accept;
    p.blk_f = h->blk_f; p.blk_i = h->blk_i; p.blk_acc = h->blk_acc;
    p.ticket = h->ticket;
    p.n_repaired = h->n_rep;
    p.best_f = h->best_f;
    p.best_x = h->best_x;
    p.rec = h->rec + (gen % h->ring);
    p.bestx_slot = h->bestx_ring + (gen % h->ring) * (int64_t)h->Dl;
    p.buf0 = h->buf[0]; p.buf1 = h->buf[1];
    p.ld = h->ld;
    p.generation = (int32_t)gen;
    p.ib = h->ib; p.r1 = h->r1; p.r2 = h->r2;
    p.next_gen = (uint32_t)(gen + 1);
    p.keys = h->keys;
    p.hndvi = h->dir_best ? 1 : h->hndvi;
    p.tbest_x = h->base_best ? h->tbest_x : nullptr;
    p.blk_cls = h->blk_cls;
    if (h->constrained) {
        p.tcvals = h->tcvals;
        p.cvals = h->cvals;
        p.contol = h->contol;
    }
    if (h->rows) {
        p.tbest_x = nullptr;                      // set by de_rows_finish_kernel from the winning candidate
        p.rows.n = h->W;
        p.rows.rank = h->rank;
        p.rows.ldc = h->ldc;
        for (int g = 0; g < h->W; ++g) {
            const RowsMailbox m = rows_mailbox(h, h->arena->peer[g], h->arena->epoch);
            p.rows.summary[g] = m.summary;
            p.rows.cand[g] = m.cand;
        }
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int oode_abi_version(void) { return OODE_ABI_VERSION; }

const char *oode_last_error(void) { return g_err.c_str(); }

int oode_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int oode_nccl_unique_id(void *id128) {
    if (!id128) return fail(OODE_EINVAL, "oode_nccl_unique_id: NULL argument");
    if (!g_nccl.load()) return fail(OODE_ENCCL, g_nccl.err);
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return OODE_OK;
}

int oode_register_objective(const char *library) {
    if (!library) return fail(OODE_EINVAL, "oode_register_objective: NULL path");
    {
        std::lock_guard<std::mutex> lock(g_obj_mutex);
        for (size_t k = OODE_OBJ_COUNT; k < g_obj_paths.size(); ++k)
            if (g_obj_paths[k] == library) return (int)k;
    }
    void *lib = dlopen(library, RTLD_NOW | RTLD_LOCAL);
    if (!lib) return fail(OODE_EINVAL, std::string("oode_register_objective: ") + dlerror());
    auto entry = reinterpret_cast<int (*)(ObjEntry *)>(dlsym(lib, "oode_custom_objective_v1"));
    if (!entry) { dlclose(lib); return fail(OODE_EINVAL, "oode_register_objective: the library does not export oode_custom_objective_v1"); }
    ObjEntry e{};
    if (entry(&e) != 0 || e.abi != OODE_ABI_VERSION || e.size_gen != (int)sizeof(GenParams) || e.size_sel != (int)sizeof(SelParams)) {
        dlclose(lib);
        return fail(OODE_EINVAL, "oode_register_objective: the objective was compiled against other kernel headers; rebuild it");
    }
    std::lock_guard<std::mutex> lock(g_obj_mutex);
    g_objs.push_back(e);
    g_obj_paths.push_back(library);
    return (int)g_objs.size() - 1;
}

int oode_shard_layout(int32_t dim, int32_t objective, int32_t world_size, int32_t rank, int32_t *col0, int32_t *ncols,
                      int32_t *leaf0, int32_t *nleaves) {
    if (!obj_known(objective)) return fail(OODE_EINVAL, "oode_shard_layout: unknown objective id");
    if (dim < 1 || world_size < 1 || rank < 0 || rank >= world_size || !col0 || !ncols || !leaf0 || !nleaves)
        return fail(OODE_EINVAL, "oode_shard_layout: bad argument");
    int ns, halo;
    obj_traits(objective, &ns, &halo);
    ReductionPlan t;
    t.build(halo ? std::max(dim - 1, 0) : dim, 0);
    const int gl = (int)t.leaves.size();
    if (gl < world_size) return fail(OODE_EINVAL, "the row has fewer pairwise-sum leaves than GPUs");
    const int a = (int)((int64_t)rank * gl / world_size), b = (int)((int64_t)(rank + 1) * gl / world_size);
    *leaf0 = a;
    *nleaves = b - a;
    *col0 = t.leaves[a].x;
    // a neighbour-coupled objective (term j reads columns j and j+1): every shard but the last also holds -- and
    // computes, from the same draws -- the first column of the next shard
    *ncols = (b == gl ? dim : t.leaves[b].x + (halo ? 1 : 0)) - t.leaves[a].x;
    return OODE_OK;
}

void oode_destroy(oode_handle *h) {
    if (!h) return;
    PhaseTimer pt("destroy");
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
    if (h->arena) {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        h->arena->in_use = false;
    }
    pt.mark("sync");
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        for (const CachedBlock &b : h->cached) g_blocks.push_back(b);
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t e : h->pev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->chunk_ev) cudaEventDestroy(e);
    if (h->comm_done) cudaEventDestroy(h->comm_done);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h->mt;
    delete h;
}

static int comm_for(const oode_config *cfg, CommEntry **out) {
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        if (!g_nccl.load()) return fail(OODE_ENCCL, g_nccl.err);
        for (CommEntry *e : g_comms)
            if (e->world == cfg->world_size && e->rank == cfg->rank && e->device == cfg->device && !memcmp(e->id, cfg->nccl_id, 128)) {
                *out = e;
                return OODE_OK;
            }
    }
    // ncclCommInitRank is a blocking collective: the cache mutex is NOT held across it (two ranks created from
    // two threads of one process would deadlock, and every dalloc / destroy of the process would stall)
    CommEntry *e = new CommEntry();
    memcpy(e->id, cfg->nccl_id, 128);
    e->world = cfg->world_size; e->rank = cfg->rank; e->device = cfg->device;
    ncclUniqueId id;
    memcpy(&id, cfg->nccl_id, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&e->comm, cfg->world_size, id, cfg->rank);
    if (r != ncclSuccess) {
        delete e;
        return fail(OODE_ENCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    }
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    g_comms.push_back(e);
    *out = e;
    return OODE_OK;
}


// Peer exchange set-up (collective over the communicator).  If the communicator's arena is too small
// for this population it is (re)built: every rank publishes the CUDA-IPC handle of its new arena
// (all-gathered over NCCL), maps the others' and checks each mapping by reading the owner's magic
// word through it; the ranks then agree (all-reduce min) on whether everybody succeeded -- if not, this
// set of ranks uses the NCCL all-gather exchange from now on.
static int arena_for(oode_handle *h, size_t need_doubles) {
    Arena &A = h->ce->arena;
    const int G = h->W;
    unsigned char *d_send = nullptr, *d_all = nullptr;
    int *d_ok = nullptr;
    DA(d_ok, 1);
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        if (A.failed || A.in_use) return OODE_OK;          // NCCL exchange for this handle (same decision on every rank)
    }
    // meet the other ranks first: once everybody is here, every rank's previous handle on this
    // communicator is finished, so nobody is still writing into (or reading) an arena
    CU(cudaMemsetAsync(d_ok, 0, sizeof(int), h->stream));      // (and d_ok = 0 unless a later step stores 1 into it)
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    {
        // Re-agree the epoch: the flags only ever grow and every rank must resume from the same count.  If a
        // previous solve ended with the ranks on different epochs (a rank-local error or stop in the middle of a
        // batch), a rank that is behind would otherwise find the others' flags already past its epoch and read
        // stale leaf sums.  Everybody continues from the maximum.
        unsigned long long *d_epoch = nullptr;
        DA(d_epoch, 1);
        CU(cudaMemcpyAsync(d_epoch, &A.epoch, sizeof(A.epoch), cudaMemcpyHostToDevice, h->stream));
        NC(g_nccl.AllReduce(d_epoch, d_epoch, 1, ncclUint64, ncclMax, h->comm, h->stream));
        CU(cudaMemcpyAsync(&A.epoch, d_epoch, sizeof(A.epoch), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    if (!A.base || A.doubles < need_doubles) {
        if (A.base) {
            for (int g = 0; g < G; ++g)
                if (A.peer[g] && g != h->rank) cudaIpcCloseMemHandle(A.peer[g]);
            cudaFree(A.base);
            A = Arena();
        }
        int ok = 1;
        void *q = nullptr;
        const size_t bytes = (X_TAIL_WORDS + need_doubles) * sizeof(double);
        if (cudaMalloc(&q, bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof(mine));
        if (ok) {
            A.base = (double *)q;
            A.doubles = need_doubles;
            const unsigned long long magic = X_MAGIC | (unsigned)h->rank;
            // between the collectives a local CUDA failure only clears `ok` (no early return): the ranks agree on the
            // outcome in the all-reduce below, so nobody is left waiting for a rank that gave up
            if (cudaMemset(A.base, 0, X_TAIL_WORDS * sizeof(unsigned long long)) != cudaSuccess ||
                cudaMemcpy(reinterpret_cast<unsigned long long *>(A.base) + 8, &magic, sizeof(magic), cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaIpcGetMemHandle(&mine, A.base) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
        DA(d_send, sizeof(mine));
        DA(d_all, sizeof(mine) * G);
        if (cudaMemcpy(d_send, &mine, sizeof(mine), cudaMemcpyHostToDevice) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        }
        NC(g_nccl.AllGather(d_send, d_all, sizeof(mine), ncclChar, h->comm, h->stream));
        std::vector<cudaIpcMemHandle_t> all(G);
        if (cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * G, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        A.peer[h->rank] = A.base;
        for (int g = 0; g < G && ok; ++g) {
            if (g == h->rank) continue;
            void *m = nullptr;
            if (cudaIpcOpenMemHandle(&m, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            A.peer[g] = (double *)m;
            unsigned long long seen = 0;
            if (cudaMemcpy(&seen, reinterpret_cast<unsigned long long *>(A.peer[g]) + 8, sizeof(seen), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            else if (seen != (X_MAGIC | (unsigned)g)) ok = 0;
        }
        if (cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) cudaGetLastError();   // d_ok stays 0
        NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->stream));
        int all_ok = 0;
        CU(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (!all_ok) {
            for (int g = 0; g < G; ++g)
                if (A.peer[g] && g != h->rank) cudaIpcCloseMemHandle(A.peer[g]);
            if (A.base) cudaFree(A.base);
            cudaGetLastError();
            A = Arena();
            A.failed = true;
            return OODE_OK;
        }
    }
    CU(cudaMemset(reinterpret_cast<unsigned long long *>(A.base) + 9, 0, sizeof(unsigned long long)));
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        A.in_use = true;
    }
    h->arena = &A;
    h->x_tail = reinterpret_cast<unsigned long long *>(A.base);
    h->xdata = A.base + X_TAIL_WORDS;
    h->p2p = true;
    return OODE_OK;
}

// General constraints: constr_val of the rows K2a has just written (or of the initial population).
static int run_constraints(oode_handle *h, int init) {
    ConstrParams cp = h->cpar;
    cp.buf0 = h->buf[0]; cp.buf1 = h->buf[1];
    cp.sel_cur = h->sel[h->cur];
    cp.ld = h->ld; cp.NP = h->NP;
    cp.D = h->D; cp.init = init;
    cp.out = h->tcvals;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((h->NP + 7) / 8, (int64_t)h->sm_count * 8));
    de_constr_kernel<<<grid, 256, 0, h->stream>>>(cp);
    CU(cudaGetLastError());
    h->ctr.gpu_launches++;
    return OODE_OK;
}

// Row shards: map every rank's two slot arrays (collective).  Mappings are cached with the communicator
// by IPC handle: the block cache hands the same allocations to the next handle, so a repeated solve maps
// nothing.
static int import_peer_buffers(oode_handle *h) {
    const int W = h->W;
    cudaIpcMemHandle_t mine[2];
    for (int b = 0; b < 2; ++b) CU(cudaIpcGetMemHandle(&mine[b], h->buf[b]));
    unsigned char *d_send = nullptr, *d_all = nullptr;
    DA(d_send, sizeof(mine));
    DA(d_all, sizeof(mine) * W);
    CU(cudaMemcpy(d_send, mine, sizeof(mine), cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    NC(g_nccl.AllGather(d_send, d_all, sizeof(mine), ncclChar, h->comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all(2 * W);
    CU(cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * W, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int ok = 1;                                  // no early return between the collectives: the ranks agree on the outcome
    for (int g = 0; g < W && ok; ++g)
        for (int b = 0; b < 2 && ok; ++b) {
            if (g == h->rank) { h->peer_buf[b][g] = h->buf[b]; continue; }
            void *q = nullptr;
            for (const PeerImport &im : h->ce->imports)
                if (im.peer == g && !memcmp(&im.handle, &all[2 * g + b], sizeof(cudaIpcMemHandle_t))) { q = im.ptr; break; }
            if (!q) {
                // keep at most a few stale mappings per peer: a mapping pins the exporter's memory even after
                // the exporter has freed it, so old ones (never those of this handle) are closed first
                int mine_of_peer = 0;
                for (const PeerImport &im : h->ce->imports) mine_of_peer += (im.peer == g);
                for (size_t k = 0; mine_of_peer >= 6 && k < h->ce->imports.size(); ++k) {
                    PeerImport &im = h->ce->imports[k];
                    if (im.peer != g || im.ptr == (void *)h->peer_buf[0][g]) continue;
                    cudaIpcCloseMemHandle(im.ptr);
                    h->ce->imports.erase(h->ce->imports.begin() + k);
                    --mine_of_peer;
                    --k;
                }
                if (cudaIpcOpenMemHandle(&q, all[2 * g + b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = 0;
                    break;
                }
                h->ce->imports.push_back({all[2 * g + b], g, q});
            }
            h->peer_buf[b][g] = (double *)q;
        }
    // a rank that could not map a peer must not leave the others waiting in oode_init: all-reduce the outcome
    int *d_ok = nullptr;
    DA(d_ok, 1);
    CU(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->stream));
    int all_ok = 0;
    CU(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (!all_ok)
        return fail(OODE_ECUDA, "oode_create: a rank could not map a peer's population through CUDA IPC (row shards need peer access)");
    return OODE_OK;
}

static RowsMailbox rows_mailbox(const oode_handle *h, double *arena_base, unsigned long long epoch) {
    auto up4 = [](int64_t v) { return (v + 3) & ~3ll; };
    double *b = arena_base + X_TAIL_WORDS + (int64_t)(epoch & 1) * h->mb_doubles;
    RowsMailbox m;
    m.vals = b;
    m.tvals = m.vals + up4(h->NP);
    m.acc = reinterpret_cast<uint8_t *>(m.tvals + up4(h->NP));
    m.summary = m.tvals + up4(h->NP) + up4((h->NP + 7) / 8);
    m.cand = m.summary + up4(4 * h->W);
    return m;
}

// Row shards, after K2b of this rank's block: push the accepted rows and the per-row results to every
// rank, raise the epoch flag, wait for everybody, then take over the other blocks' decisions and write the
// generation record (de_rows.cuh).
static int rows_exchange(oode_handle *h, int init, int64_t gen) {
    Arena &A = *h->arena;
    const unsigned long long e = A.epoch;
    PushParams pp{};
    PeerSync ps{};
    pp.n = ps.n = h->W;
    pp.rank = ps.rank = h->rank;
    for (int g = 0; g < h->W; ++g) {
        pp.mb[g] = rows_mailbox(h, A.peer[g], e);
        pp.peer_buf0[g] = h->peer_buf[0][g];
        pp.peer_buf1[g] = h->peer_buf[1][g];
        ps.flags[g] = reinterpret_cast<unsigned long long *>(A.peer[g]);
    }
    pp.buf0 = h->buf[0]; pp.buf1 = h->buf[1];
    pp.sel_cur = h->sel[h->cur];
    pp.accept = h->accept;
    pp.vals = h->vals; pp.tvals = h->tvals;
    pp.ld = h->ld; pp.Dl = h->Dl;
    pp.row0 = h->row0; pp.nrows = h->nrows;
    pp.init = init;
    pp.rows_sent = (h->rows_fused && !init) ? 1 : 0;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((h->nrows + 7) / 8, (int64_t)h->sm_count * 8));
    de_rows_push_kernel<<<grid, 256, 0, h->stream>>>(pp);
    CU(cudaGetLastError());
    ps.counter_slot = -1;
    ps.epoch = e;
    de_publish_kernel<<<1, 32, 0, h->stream>>>(ps, h->n_rep);
    CU(cudaGetLastError());
    de_wait_kernel<<<1, 32, 0, h->stream>>>(h->x_tail, h->W, e, h->x_tail + 9);
    CU(cudaGetLastError());
    FinishParams fp{};
    fp.mb = rows_mailbox(h, A.base, e);
    fp.G = h->W; fp.rank = h->rank; fp.ldc = h->ldc;
    fp.NP = h->NP; fp.row0 = h->row0; fp.nrows = h->nrows;
    fp.Dl = h->Dl; fp.init = init;
    fp.vals = h->vals; fp.tvals = h->tvals; fp.accept = h->accept;
    fp.sel_cur = h->sel[h->cur];
    fp.sel_next = h->sel[h->cur ^ 1];
    fp.best_f = h->best_f; fp.best_x = h->best_x;
    fp.bestx_slot = h->bestx_ring + (gen % h->ring) * (int64_t)h->Dl;
    fp.tbest_x = h->base_best ? h->tbest_x : nullptr;
    fp.rec = h->rec + (gen % h->ring);
    fp.generation = (int32_t)gen;
    const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>((h->NP + 255) / 256, (int64_t)h->sm_count * 8));
    de_rows_finish_kernel<<<fgrid, 256, 0, h->stream>>>(fp);
    CU(cudaGetLastError());
    h->ctr.gpu_launches += 4;
    return OODE_OK;
}

static int create_impl(const oode_config *cfg, oode_handle *h) {
    PhaseTimer pt("create");
    h->cfg = *cfg;
    h->D = cfg->dim;
    h->NP = cfg->population;
    h->W = cfg->world_size > 1 ? cfg->world_size : 1;
    h->rank = h->W > 1 ? cfg->rank : 0;
    h->shard = h->W > 1 ? cfg->shard : OODE_SHARD_NONE;
    h->rows = h->shard == OODE_SHARD_ROWS;
    h->G = h->rows ? 1 : h->W;
    h->obj = obj_entry(cfg->objective);
    obj_traits(cfg->objective, &h->nslots, &h->halo);
    // the whole row's reduction tree; a column shard is a run of consecutive leaves of it
    h->gplan.build(h->halo ? std::max(h->D - 1, 0) : h->D, 0);
    const int gl = (int)h->gplan.leaves.size();
    h->shard_col0.assign(h->G, 0);
    h->shard_dl.assign(h->G, h->D);
    int l0 = 0, l1 = gl;
    h->Lmax = gl;
    if (h->G > 1) {
        h->Lmax = 0;
        for (int r = 0; r < h->G; ++r) {
            const int a = (int)((int64_t)r * gl / h->G), b = (int)((int64_t)(r + 1) * gl / h->G);
            h->Lmax = std::max(h->Lmax, b - a);
            h->shard_col0[r] = h->gplan.leaves[a].x;
            h->shard_dl[r] = (b == gl ? h->D : h->gplan.leaves[b].x + (h->halo ? 1 : 0)) - h->gplan.leaves[a].x;
            if (r == h->rank) { l0 = a; l1 = b; }
        }
    }
    h->leaf0 = l0;
    h->col0 = h->shard_col0[h->rows ? 0 : h->rank];
    h->Dl = h->shard_dl[h->rows ? 0 : h->rank];
    h->dl_max = *std::max_element(h->shard_dl.begin(), h->shard_dl.end());
    h->chunk = h->NP * h->Lmax * h->nslots + 1;
    h->ld = h->Dl + (h->Dl & 1);        // even, so column pairs are 16-byte aligned
    h->row0 = 0;
    h->nrows = h->NP;
    if (h->rows) {                      // contiguous row blocks
        h->row0 = (int64_t)h->rank * h->NP / h->W;
        h->nrows = (int64_t)(h->rank + 1) * h->NP / h->W - h->row0;
    }
    h->uniform_bounds = true;
    for (int j = 1; j < cfg->dim; ++j) h->uniform_bounds &= (cfg->lb[j] == cfg->lb[0] && cfg->ub[j] == cfg->ub[0]);
    h->lo_s = cfg->lb[0];
    h->hi_s = cfg->ub[0];
    {
        double m = 0;
        for (int j = 0; j < cfg->dim; ++j) m = std::max(m, std::max(std::fabs(cfg->lb[j]), std::fabs(cfg->ub[j])));
        // rastrigin/ackley: |2*pi*x|; griewank: |x / sqrt(j+1)| <= |x|
        h->cos_checked = !(m * 6.283185307179586 <= 1.0e6);
    }
    h->hndvi = cfg->hndvi > 0 ? cfg->hndvi : 1;
    h->base_best = cfg->base_strategy == OODE_BASE_BEST;
    h->dir_best = cfg->direction_strategy == OODE_DIR_BEST;
    h->const_F = cfg->factor_strategy == OODE_FACTOR_CONSTANT;
    // K2a form.  Rows of >= 32 columns walk through the TMA-fed row kernel (de_trial_rows.cuh): one worker per warp,
    // half-warp workers for narrow rows (same speed within noise on every shape measured; quarter-warp workers were
    // slower everywhere and are gone).  Tiny rows, barycentres of several
    // gathered rows per trial (hndvi > 1 without the directed search) and neighbour-coupled objectives with narrow rows
    // keep the quad-per-leaf kernel; neighbour-coupled objectives with wide rows the round-1 TMA forms.
    // OODE_TRIAL_KERNEL=quad|rows1|rows2|tma|tma2 overrides (tma / tma2 need a build with -DOODE_KEEP_OLD_TMA
    // unless the objective couples neighbours).
    {
        int maxlen = 1;
        for (int l = l0; l < l1; ++l) maxlen = std::max(maxlen, h->gplan.leaves[l].y);
        h->rs = std::max(2, (maxlen + 1) & ~1);
    }
    h->fast = std::fabs(cfg->diff_factor) <= 0.999 && !h->cos_checked;
    h->trusted_box = true;                    // pop[0] = x0 (de_oo.py:149-150) may lie outside the box
    if (cfg->x0)
        for (int j = 0; j < cfg->dim; ++j)
            if (std::isfinite(cfg->x0[j])) h->trusted_box &= (cfg->x0[j] >= cfg->lb[j] && cfg->x0[j] <= cfg->ub[j]);
    if (h->halo) {
        h->form = h->Dl >= 512 ? FORM_TMA : (h->Dl >= 96 && h->uniform_bounds ? FORM_TMA2 : FORM_QUAD);
    } else {
        h->form = h->Dl >= 32 ? FORM_ROWS : FORM_QUAD;
        h->wpw = h->Dl >= 192 ? 1 : 2;
    }
    if (const char *k = getenv("OODE_TRIAL_KERNEL")) {
        if (!strcmp(k, "quad")) h->form = FORM_QUAD;
        else if (!strcmp(k, "tma")) h->form = FORM_TMA;
        else if (!strcmp(k, "tma2")) h->form = FORM_TMA2;
        else if (!strncmp(k, "rows", 4) && !h->halo && (k[4] == '1' || k[4] == '2') && !k[5]) {
            h->form = FORM_ROWS;
            h->wpw = k[4] - '0';
        }
    }
    // barycentres of several gathered rows per trial only exist in the quad-per-leaf kernel
    if (h->hndvi > 1 && !h->dir_best) h->form = FORM_QUAD;
    // column shards of a neighbour-coupled objective overlap by one column, whose bound repairs only the last owner
    // may count (GenParams::count_tail): the quad-per-leaf kernel knows how
    if (h->halo && h->G > 1) h->form = FORM_QUAD;
    if (h->form == FORM_ROWS) {
        const size_t per_warp = rows_smem_bytes(1, h->wpw, h->rs);
        h->warps = (int)std::max<size_t>(1, std::min<size_t>(rows_maxw(h->wpw), ROWS_SMEM_MAX / per_warp));
        if (const char *e = getenv("OODE_ROWS_WARPS")) h->warps = std::max(1, std::min(h->warps, atoi(e)));
    }
    h->ring = (cfg->max_batch > 0 ? cfg->max_batch : 64) + 1;
    h->nT = h->halo ? std::max(h->Dl - 1, 0) : h->Dl;
    for (int l = l0; l < l1; ++l) h->plan.leaves.push_back(make_int2(h->gplan.leaves[l].x - h->col0, h->gplan.leaves[l].y));
    h->plan.prog = h->gplan.prog;

    // Philox round keys
    uint32_t k0 = (uint32_t)(cfg->seed & 0xFFFFFFFFull), k1 = (uint32_t)(cfg->seed >> 32);
    for (int r = 0; r < 10; ++r) {
        h->keys.rk[2 * r] = k0;
        h->keys.rk[2 * r + 1] = k1;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }

    CU(cudaSetDevice(cfg->device));
    CU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));

    const int64_t NP = h->NP;
    const int D = h->D;
    const int Dl = h->Dl, c0 = h->col0;
    const int nleaves = (int)h->plan.leaves.size();
    pt.mark("device, stream");
    if (h->rows) {
        { int rc = comm_for(cfg, &h->ce); if (rc) return rc; }
        h->comm = h->ce->comm;
        pt.mark("nccl communicator");
        auto up4 = [](int64_t v) { return (v + 3) & ~3ll; };
        h->ldc = (int)up4(h->ld + 2);
        h->mb_doubles = 2 * up4(NP) + up4((NP + 7) / 8) + up4(4 * h->W) + (int64_t)h->W * h->ldc;
        { int rc = arena_for(h, (size_t)(2 * h->mb_doubles)); if (rc) return rc; }
        pt.mark("peer arena");
        if (!h->p2p)
            return fail(OODE_ECUDA, "oode_create: row shards need the ranks to map each other's memory (CUDA IPC peer access), "
                                    "which is not available between these devices");
    }
    if (h->G > 1) {
        { int rc = comm_for(cfg, &h->ce); if (rc) return rc; }
        h->comm = h->ce->comm;
        pt.mark("nccl communicator");
        const char *xm = getenv("OODE_EXCHANGE");           // p2p (default when it can be set up) | nccl
        const bool want_p2p = h->G <= MAX_PEERS && !(xm && !strcmp(xm, "nccl"));
        const int64_t chunk_p2p = (h->chunk + 15) & ~15ll;   // rank blocks stay 128-byte aligned (contiguous runs of leaf sums)
        if (want_p2p) { int rc = arena_for(h, (size_t)(2 * chunk_p2p * h->G)); if (rc) return rc; }
        pt.mark("peer arena");
        if (h->p2p) h->chunk = chunk_p2p;
        else DA(h->partials_all, h->chunk * h->G);
        DA(h->bx_send, h->dl_max);
        DA(h->bx_all, (int64_t)h->dl_max * h->G);
        if (xm && !strcmp(xm, "p2p") && !h->p2p)
            return fail(OODE_ECUDA, "OODE_EXCHANGE=p2p, but the ranks could not map each other's memory (CUDA IPC)");
    }
    DA(h->leaf_rank, gl);
    DA(h->leaf_li, gl);
    {
        std::vector<int> lr(gl), li(gl);
        for (int r = 0, v = 0; r < h->G; ++r) {
            const int a = (int)((int64_t)r * gl / h->G), b = (int)((int64_t)(r + 1) * gl / h->G);
            for (int l = a; l < b; ++l, ++v) { lr[v] = r; li[v] = l - a; }
        }
        h->fold_fast = fold_is_levelwise(h->gplan.prog, gl) && !getenv("OODE_FOLD_INTERPRETER");
        for (int v = 0; v < gl && v < 8; ++v) { h->fast_rank[v] = (int8_t)lr[v]; h->fast_li[v] = (int8_t)li[v]; }
        CU(cudaMemcpy(h->leaf_rank, lr.data(), gl * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->leaf_li, li.data(), gl * sizeof(int), cudaMemcpyHostToDevice));
    }
    h->chunks = 1;
    if (h->G > 1 && !h->p2p) {
        h->chunks = NP >= 65536 ? 4 : 1;
        if (const char *e = getenv("OODE_CHUNKS")) h->chunks = std::max(1, atoi(e));
        h->chunks = (int)std::min<int64_t>(h->chunks, NP);
        CU(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        h->chunk_ev.resize(h->chunks);
        for (auto &e : h->chunk_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->comm_done, cudaEventDisableTiming));
    }
    h->chunk_rows = (NP + h->chunks - 1) / h->chunks;
    h->chunks = (int)((NP + h->chunk_rows - 1) / h->chunk_rows);
    DA(h->buf[0], NP * h->ld);
    DA(h->buf[1], NP * h->ld);
    if (h->rows) { int rc = import_peer_buffers(h); if (rc) return rc; }
    // single-leaf rows through the row kernel: the accepted rows ride along with the trial kernel
    h->rows_fused = h->rows && h->form == FORM_ROWS && gl == 1 && !getenv("OODE_ROWS_UNFUSED");
    pt.mark("population slot arrays");
    DA(h->sel[0], NP);
    DA(h->sel[1], NP);
    DA(h->accept, NP);
    DA(h->vals, NP);
    DA(h->tvals, NP);
    DA(h->ib, NP);
    DA(h->r1, NP * h->hndvi);
    DA(h->r2, NP * h->hndvi);
    DA(h->tbest_x, h->ld + 2);
    DA(h->dvec, 2 * (h->ld + 2));
    DA(h->rdir, 2 * OODE_MAX_HNDVI);
    DA(h->lb, Dl + 2);
    DA(h->ub, Dl + 2);
    DA(h->aux, Dl + 2);
    DA(h->raux, Dl + 2);
    DA(h->x0, Dl + 2);
    DA(h->leaves, nleaves);
    DA(h->prog, h->plan.prog.size());
    DA(h->partials, (h->p2p && !h->rows) ? 1 : h->chunk);      // [NP][Lmax][NSL] + 1 counter slot (peer exchange: unused)
    const int64_t nblk = (NP + 255) / 256;
    DA(h->blk_f, nblk);
    DA(h->blk_i, nblk);
    DA(h->blk_acc, nblk);
    DA(h->blk_cls, nblk);
    DA(h->ticket, 1);
    DA(h->xticket, 1);
    DA(h->wait_ns, 1);
    CU(cudaMemset(h->xticket, 0, sizeof(unsigned int)));
    CU(cudaMemset(h->wait_ns, 0, sizeof(unsigned long long)));
    DA(h->n_rep, 1);
    DA(h->zero_word, 1);
    DA(h->best_f, 1);
    DA(h->best_x, Dl);
    DA(h->bestx_ring, (int64_t)h->ring * Dl);
    DA(h->rec, h->ring);
    if (cfg->rng != OODE_RNG_PHILOX) {
        DA(h->Uf, NP * D);
        DA(h->Uc, NP * D);
        h->h_idx.resize((size_t)(1 + 2 * h->hndvi) * NP + 2 * OODE_MAX_HNDVI);
    }
    if (cfg->rng == OODE_RNG_CPYTHON) {
        h->mt = new MT19937CPython(cfg->seed);
        h->h_u.resize((size_t)NP * D);
    }

    h->ctr.h2d_bytes += (int64_t)(3 + (cfg->x0 ? 1 : 0)) * Dl * sizeof(double);
    CU(cudaMemcpy(h->lb, cfg->lb + c0, Dl * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->ub, cfg->ub + c0, Dl * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<double> aux(D, 0.0);
    if (cfg->obj_aux) std::copy(cfg->obj_aux, cfg->obj_aux + D, aux.begin());
    CU(cudaMemcpy(h->aux, aux.data() + c0, Dl * sizeof(double), cudaMemcpyHostToDevice));
    {
        std::vector<double> raux(D + 2, 0.0);
        for (int j = 0; j < D; ++j) raux[j] = 1.0 / aux[j];                    // one IEEE rounding
        CU(cudaMemcpy(h->raux, raux.data() + c0, (Dl + 2) * sizeof(double), cudaMemcpyHostToDevice));
    }
    h->have_x0 = false;
    if (cfg->x0) {
        for (int j = 0; j < D; ++j) h->have_x0 |= std::isfinite(cfg->x0[j]);     // de_oo.py:149
        CU(cudaMemcpy(h->x0, cfg->x0 + c0, Dl * sizeof(double), cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(h->leaves, h->plan.leaves.data(), nleaves * sizeof(int2), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->prog, h->plan.prog.data(), h->plan.prog.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemset(h->ticket, 0, sizeof(unsigned int)));
    CU(cudaMemset(h->n_rep, 0, sizeof(unsigned long long)));
    CU(cudaMemset(h->zero_word, 0, sizeof(unsigned long long)));
    CU(cudaMemset(h->accept, 0, NP));
    // the trial kernels load donor indices unconditionally (also under strategies that do not use them)
    CU(cudaMemset(h->ib, 0, NP * sizeof(int32_t)));
    CU(cudaMemset(h->r1, 0, NP * h->hndvi * sizeof(int32_t)));
    CU(cudaMemset(h->r2, 0, NP * h->hndvi * sizeof(int32_t)));
    pt.mark("other arrays, uploads");
    return OODE_OK;
}

int oode_create(const oode_config *cfg, oode_handle **out) {
    NvtxRange nvtx_range("oode_create");
    if (!cfg || !out) return fail(OODE_EINVAL, "oode_create: NULL argument");
    *out = nullptr;
    if (cfg->abi_version != OODE_ABI_VERSION) return fail(OODE_EINVAL, "oode_create: ABI version mismatch");
    if (cfg->dim < 1 || cfg->population < 1) return fail(OODE_EINVAL, "oode_create: dim and population must be >= 1");
    if (cfg->population >= (1ll << 31) - 65536) return fail(OODE_EINVAL, "oode_create: population must be < 2^31 - 65536");
    if (!cfg->lb || !cfg->ub) return fail(OODE_EINVAL, "oode_create: lb and ub are required");
    for (int j = 0; j < cfg->dim; ++j)
        if (!std::isfinite(cfg->lb[j]) || !std::isfinite(cfg->ub[j]))
            return fail(OODE_EINVAL, "this solver requires finite lb, ub: lb <= x <= ub");   // de_oo.py:122
    if (!obj_known(cfg->objective)) return fail(OODE_EINVAL, "oode_create: unknown objective id");
    if (obj_entry(cfg->objective).needs_aux && !cfg->obj_aux) return fail(OODE_EINVAL, "oode_create: this objective needs obj_aux[D]");
    if (cfg->rng < OODE_RNG_PHILOX || cfg->rng > OODE_RNG_CPYTHON) return fail(OODE_EINVAL, "oode_create: unknown rng mode");
    if (cfg->base_strategy < 0 || cfg->base_strategy > OODE_BASE_BEST || cfg->direction_strategy < 0 ||
        cfg->direction_strategy > OODE_DIR_BEST || cfg->factor_strategy < 0 || cfg->factor_strategy > OODE_FACTOR_CONSTANT)
        return fail(OODE_EINVAL, "oode_create: unknown strategy id");
    if (cfg->hndvi < 0 || cfg->hndvi > OODE_MAX_HNDVI)
        return fail(OODE_EINVAL, "oode_create: hndvi must be in [1, " + std::to_string(OODE_MAX_HNDVI) + "]");
    if ((cfg->objective == OODE_OBJ_GRIEWANK || cfg->objective == OODE_OBJ_SHIFTED_SPHERE) && !cfg->obj_aux)
        return fail(OODE_EINVAL, "oode_create: this objective needs obj_aux[D]");
    if (cfg->world_size > 1) {
        if (cfg->shard != OODE_SHARD_COLS && cfg->shard != OODE_SHARD_ROWS)
            return fail(OODE_EINVAL, "oode_create: world_size > 1 needs shard = OODE_SHARD_COLS or OODE_SHARD_ROWS");
        if (cfg->rank < 0 || cfg->rank >= cfg->world_size || !cfg->nccl_id)
            return fail(OODE_EINVAL, "oode_create: bad rank / missing nccl_id");
        // replay / CPython draws on a sharded population: every rank holds the whole draw arrays (they are indexed by
        // global row and column) and, for the CPython stream, regenerates the identical sequence itself
        if (cfg->shard == OODE_SHARD_ROWS) {
            if (cfg->world_size > OODE_MAX_PEERS) return fail(OODE_EINVAL, "oode_create: row shards support at most 8 ranks");
            if (cfg->population < cfg->world_size) return fail(OODE_EINVAL, "oode_create: fewer rows than ranks");
        } else {
        int ns, halo;
        obj_traits(cfg->objective, &ns, &halo);
        ReductionPlan t;
        t.build(halo ? std::max(cfg->dim - 1, 0) : cfg->dim, 0);
        if ((int)t.leaves.size() < cfg->world_size)
            return fail(OODE_EINVAL, "oode_create: the row has fewer pairwise-sum leaves than GPUs; it cannot be column-sharded that wide");
        }
    }
    if (oode_device_count() <= cfg->device || cfg->device < 0)
        return fail(OODE_ECUDA, "oode_create: no such CUDA device (liboode has no CPU fallback)");
    oode_handle *h = new oode_handle();
    const int rc = create_impl(cfg, h);
    if (rc) {
        const std::string keep = g_err;
        oode_destroy(h);
        g_err = keep;
        return rc;
    }
    *out = h;
    return OODE_OK;
}

static int read_records(oode_handle *h, int64_t first_gen, int n, oode_gen_record *out) {
    h->ctr.d2h_bytes += (int64_t)n * sizeof(oode_gen_record);
    for (int k = 0; k < n; ++k) {
        const int64_t g = first_gen + k;
        CU(cudaMemcpyAsync(out + k, h->rec + (g % h->ring), sizeof(oode_gen_record), cudaMemcpyDeviceToHost, h->stream));
    }
    unsigned long long werr = 0;
    if (h->p2p) CU(cudaMemcpyAsync(&werr, h->x_tail + 9, sizeof(werr), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (werr) return fail(OODE_ENCCL, "peer exchange: a rank did not deliver its leaf sums within the time limit (epoch " +
                                      std::to_string(werr) + ")");
    return OODE_OK;
}

int oode_init(oode_handle *h, const double *u0, oode_gen_record *out) {
    NvtxRange nvtx_range("oode_init");
    if (!h) return fail(OODE_EINVAL, "oode_init: NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    const int64_t NP = h->NP;
    const int D = h->D;
    InitParams ip{};
    ip.buf0 = h->buf[0];
    ip.sel0 = h->sel[0];
    ip.sel1 = h->sel[1];
    ip.ld = h->ld;
    ip.NP = NP;
    ip.Dl = h->Dl; ip.D = D; ip.col0 = h->col0;
    ip.lb = h->lb; ip.ub = h->ub;
    ip.x0 = h->have_x0 ? h->x0 : nullptr;
    ip.keys = h->keys;
    ip.use_philox = (h->cfg.rng == OODE_RNG_PHILOX);
    if (h->cfg.rng == OODE_RNG_REPLAY) {
        if (!u0) return fail(OODE_EINVAL, "oode_init: replay mode needs the initial uniforms u0[NP*D]");
        CU(cudaMemcpyAsync(h->Uf, u0, (size_t)NP * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->ctr.h2d_bytes += (int64_t)NP * D * sizeof(double);
        ip.U0 = h->Uf;
    } else if (h->cfg.rng == OODE_RNG_CPYTHON) {
        h->mt->fill_random(h->h_u.data(), (size_t)NP * D);                     // Rand(NP,D), de_oo.py:147
        CU(cudaMemcpyAsync(h->Uf, h->h_u.data(), (size_t)NP * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->ctr.h2d_bytes += (int64_t)NP * D * sizeof(double);
        ip.U0 = h->Uf;
    }
    h->cur = 0;
    h->gen = 0;
    h->min_gen = 0;
    CU(cudaEventRecord(h->ev0, h->stream));
    const int64_t n = NP * ((h->Dl + 1) / 2);          // one thread per column pair
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8));
    de_init_kernel<<<grid, 256, 0, h->stream>>>(ip);
    CU(cudaGetLastError());
    h->ctr.gpu_launches++;
    if (h->rows) ++h->arena->epoch;
    GenParams gp = gen_params(h, 0);
    { int rc = run_trial(h, MODE_EVAL, gp); if (rc) return rc; }
    if (h->constrained) { int rc = run_constraints(h, 1); if (rc) return rc; }
    SelParams sp = sel_params(h, 1, 0);
    CU(launch_select(h, sp));
    if (h->rows) { int rc = rows_exchange(h, 1, 0); if (rc) return rc; }
    CU(cudaEventRecord(h->ev1, h->stream));
    oode_gen_record rec;
    int rc = read_records(h, 0, 1, &rec);
    if (rc) return rc;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->ctr.kernel_ms += ms;
    h->ctr.trial_evals += NP;
    h->inited = true;
    if (out) *out = rec;
    return OODE_OK;
}

// Index draws of one generation, staged in h_idx as [ib: NP][r1: h*NP][r2: h*NP] (or [r: 2h] for the
// directed search) and copied to the device arrays.
static int upload_indices(oode_handle *h) {
    const int64_t NP = h->NP, hn = (int64_t)h->hndvi * NP;
    const int32_t *src = h->h_idx.data();
    if (!h->base_best) CU(cudaMemcpyAsync(h->ib, src, NP * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (h->dir_best) {
        CU(cudaMemcpyAsync(h->rdir, src + NP, 2 * h->hndvi * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    } else {
        CU(cudaMemcpyAsync(h->r1, src + NP, hn * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->r2, src + NP + hn, hn * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    }
    h->ctr.h2d_bytes += ((h->base_best ? 0 : NP) + (h->dir_best ? 2 * h->hndvi : 2 * hn)) * (int64_t)sizeof(int32_t);
    return OODE_OK;
}

static int upload_draws(oode_handle *h, const oode_replay_draws *d) {
    const int64_t NP = h->NP, hn = (int64_t)h->hndvi * NP;
    const int D = h->D;
    if ((!h->base_best && !d->ib) || !d->r1 || (!h->dir_best && !d->r2) || (!h->const_F && !d->Uf) || !d->Uc)
        return fail(OODE_EINVAL, "oode_step: a replay draw array this strategy needs is NULL");
    auto stage = [&](const int64_t *from, int64_t n, int32_t *to) {
        for (int64_t i = 0; i < n; ++i) {
            if (from[i] < 0 || from[i] >= NP) return false;
            to[i] = (int32_t)from[i];
        }
        return true;
    };
    bool ok = true;
    if (!h->base_best) ok &= stage(d->ib, NP, h->h_idx.data());
    if (h->dir_best) ok &= stage(d->r1, 2 * h->hndvi, h->h_idx.data() + NP);
    else ok &= stage(d->r1, hn, h->h_idx.data() + NP) && stage(d->r2, hn, h->h_idx.data() + NP + hn);
    if (!ok) return fail(OODE_EINVAL, "oode_step: replay donor index out of range");
    { int rc = upload_indices(h); if (rc) return rc; }
    if (!h->const_F) CU(cudaMemcpyAsync(h->Uf, d->Uf, (size_t)NP * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->Uc, d->Uc, (size_t)NP * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->ctr.h2d_bytes += (h->const_F ? 1 : 2) * (int64_t)NP * D * sizeof(double);
    // the staging vector is reused next generation
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

// CPython-stream mode: regenerate the reference's draws in its draw order on the host and upload them:
// base indices (de_oo.py:188, 'random' base only), direction indices (:200-201 two (h,NP) blocks, or
// :210 one block of 2h for the directed search), Ft uniforms (:236, 'random' factor only), crossover
// uniforms (:245).
static int generate_cpython_draws(oode_handle *h) {
    const int64_t NP = h->NP, hn = (int64_t)h->hndvi * NP;
    const int D = h->D;
    int32_t *idx = h->h_idx.data();
    if (!h->base_best) for (int64_t i = 0; i < NP; ++i) idx[i] = (int32_t)h->mt->randbelow((uint32_t)NP);
    if (h->dir_best) for (int64_t i = 0; i < 2 * h->hndvi; ++i) idx[NP + i] = (int32_t)h->mt->randbelow((uint32_t)NP);
    else for (int64_t i = 0; i < 2 * hn; ++i) idx[NP + i] = (int32_t)h->mt->randbelow((uint32_t)NP);
    { int rc = upload_indices(h); if (rc) return rc; }
    if (!h->const_F) {
        h->mt->fill_random(h->h_u.data(), (size_t)NP * D);
        CU(cudaMemcpyAsync(h->Uf, h->h_u.data(), (size_t)NP * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->ctr.h2d_bytes += (int64_t)NP * D * sizeof(double);
    }
    CU(cudaStreamSynchronize(h->stream));
    h->mt->fill_random(h->h_u.data(), (size_t)NP * D);
    CU(cudaMemcpyAsync(h->Uc, h->h_u.data(), (size_t)NP * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->ctr.h2d_bytes += (int64_t)NP * D * sizeof(double);
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

// Directed search: the generation's two barycentre vectors (de_direct_kernel), before K2a.
static int run_direct(oode_handle *h, uint32_t gen) {
    DirectParams dp{};
    dp.r = (h->cfg.rng == OODE_RNG_PHILOX) ? nullptr : h->rdir;
    dp.vals = h->vals;
    dp.cvals = h->constrained ? h->cvals : nullptr;
    dp.sel_cur = h->sel[h->cur];
    dp.buf0 = h->buf[0]; dp.buf1 = h->buf[1];
    dp.ld = h->ld; dp.NP = h->NP;
    dp.Dl = h->Dl; dp.h = h->hndvi;
    dp.gen = gen;
    dp.keys = h->keys;
    dp.d1_vec = h->dvec;
    dp.d2_vec = h->dvec + (h->ld + 2);
    de_direct_kernel<<<1, 256, 0, h->stream>>>(dp);
    CU(cudaGetLastError());
    h->ctr.gpu_launches++;
    return OODE_OK;
}

int oode_step(oode_handle *h, int32_t n_gens, const oode_replay_draws *draws, oode_gen_record *out) {
    NvtxRange nvtx_range("oode_step");
    if (!h) return fail(OODE_EINVAL, "oode_step: NULL handle");
    if (!h->inited) return fail(OODE_ESTATE, "oode_step: call oode_init first");
    if (n_gens < 1 || n_gens > h->ring - 1) return fail(OODE_EINVAL, "oode_step: n_gens must be in [1, max_batch]");
    if (h->cfg.rng == OODE_RNG_REPLAY && !draws) return fail(OODE_EINVAL, "oode_step: replay mode needs draws");
    CU(cudaSetDevice(h->cfg.device));
    const int mode = (h->cfg.rng == OODE_RNG_PHILOX) ? MODE_PHILOX : MODE_REPLAY;
    float total_ms = 0;
    const int64_t first = h->gen + 1;
    if (mode == MODE_PHILOX) CU(cudaEventRecord(h->ev0, h->stream));
    for (int k = 0; k < n_gens; ++k) {
        const int64_t g = h->gen + 1;
        if (h->cfg.rng == OODE_RNG_REPLAY) {
            int rc = upload_draws(h, draws + k);
            if (rc) return rc;
        } else if (h->cfg.rng == OODE_RNG_CPYTHON) {
            int rc = generate_cpython_draws(h);
            if (rc) return rc;
        }
        if (mode != MODE_PHILOX) CU(cudaEventRecord(h->ev0, h->stream));
        if (h->rows) ++h->arena->epoch;
        GenParams gp = gen_params(h, (uint32_t)g);
        if (h->dir_best) { int rc = run_direct(h, (uint32_t)g); if (rc) return rc; }
        if (h->profiling) CU(cudaEventRecord(h->pev[4 * k], h->stream));
        h->xev = nullptr;
        { int rc = run_trial(h, mode, gp); if (rc) return rc; }
        h->xev = nullptr;
        if (h->constrained) { int rc = run_constraints(h, 0); if (rc) return rc; }
        if (h->profiling) CU(cudaEventRecord(h->pev[4 * k + 1], h->stream));
        SelParams sp = sel_params(h, 0, g);
        CU(launch_select(h, sp));
        if (h->profiling) CU(cudaEventRecord(h->pev[4 * k + 2], h->stream));
        if (h->rows) {
            int rc = rows_exchange(h, 0, g);
            if (rc) return rc;
            if (h->profiling) CU(cudaEventRecord(h->pev[4 * k + 3], h->stream));
        }
        if (mode != MODE_PHILOX) {
            CU(cudaEventRecord(h->ev1, h->stream));
            CU(cudaEventSynchronize(h->ev1));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            total_ms += ms;
        }
        h->cur ^= 1;
        h->gen = g;
    }
    if (mode == MODE_PHILOX) {
        CU(cudaEventRecord(h->ev1, h->stream));
        CU(cudaEventSynchronize(h->ev1));
        CU(cudaEventElapsedTime(&total_ms, h->ev0, h->ev1));
    }
    if (h->profiling) {
        CU(cudaStreamSynchronize(h->stream));
        for (int k = 0; k < n_gens; ++k) {
            float a = 0, b = 0;
            CU(cudaEventElapsedTime(&a, h->pev[4 * k], h->pev[4 * k + 1]));
            CU(cudaEventElapsedTime(&b, h->pev[4 * k + 1], h->pev[4 * k + 2]));
            if (h->rows) {    // row shards: push ... finish after K2b (column shards: K2b's prologue counts its own wait)
                float x = 0;
                CU(cudaEventElapsedTime(&x, h->pev[4 * k + 2], h->pev[4 * k + 3]));
                h->ctr.exchange_wait_ms += x;
            }
            h->ctr.trial_kernel_ms += a;
            h->ctr.select_kernel_ms += b;
        }
        h->ctr.profiled_generations += n_gens;
    }
    h->ctr.kernel_ms += total_ms;
    h->ctr.generations += n_gens;
    h->ctr.trial_evals += (int64_t)n_gens * h->NP;
    if (out) return read_records(h, first, n_gens, out);
    return read_records(h, first, 0, nullptr);
}

int oode_set_constraints(oode_handle *h, const oode_constraints *c) {
    if (!h || !c) return fail(OODE_EINVAL, "oode_set_constraints: NULL argument");
    if (h->inited) return fail(OODE_ESTATE, "oode_set_constraints: call it before oode_init");
    if (h->W > 1) return fail(OODE_EINVAL, "oode_set_constraints: general constraints are not available on a sharded handle");
    const int m[4] = {c->n_lin_ineq, c->n_lin_eq, c->n_quad_ineq, c->n_quad_eq};
    const double *M[4] = {c->A, c->Aeq, c->Qc, c->Qh};
    const double *rhs[4] = {c->b, c->beq, c->rc, c->rh};
    for (int k = 0; k < 4; ++k) {
        if (m[k] < 0 || (m[k] > 0 && (!M[k] || !rhs[k]))) return fail(OODE_EINVAL, "oode_set_constraints: bad row count / NULL matrix");
    }
    CU(cudaSetDevice(h->cfg.device));
    h->constrained = (m[0] + m[1] + m[2] + m[3]) > 0;
    if (!h->constrained) return OODE_OK;
    for (int k = 0; k < 4; ++k) {
        h->cpar.m[k] = m[k];
        h->cpar.M[k] = nullptr;
        h->cpar.rhs[k] = nullptr;
        if (!m[k]) continue;
        double *dM = nullptr, *dr = nullptr;
        DA(dM, (size_t)m[k] * h->D);
        DA(dr, m[k]);
        CU(cudaMemcpy(dM, M[k], (size_t)m[k] * h->D * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dr, rhs[k], m[k] * sizeof(double), cudaMemcpyHostToDevice));
        h->ctr.h2d_bytes += (int64_t)m[k] * (h->D + 1) * sizeof(double);
        h->cpar.M[k] = dM;
        h->cpar.rhs[k] = dr;
    }
    h->contol = c->contol;
    DA(h->cvals, h->NP);
    DA(h->tcvals, h->NP);
    return OODE_OK;
}

int oode_get_constr_vals(oode_handle *h, double *cvals) {
    if (!h || !cvals) return fail(OODE_EINVAL, "oode_get_constr_vals: NULL argument");
    if (!h->inited) return fail(OODE_ESTATE, "oode_get_constr_vals: call oode_init first");
    if (!h->constrained) { memset(cvals, 0, h->NP * sizeof(double)); return OODE_OK; }
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(cvals, h->cvals, h->NP * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

int oode_set_profiling(oode_handle *h, int32_t on) {
    if (!h) return fail(OODE_EINVAL, "oode_set_profiling: NULL handle");
    CU(cudaSetDevice(h->cfg.device));
    if (on && h->pev.empty()) {
        h->pev.resize(4 * (size_t)h->ring);
        for (auto &e : h->pev) CU(cudaEventCreate(&e));
    }
    h->profiling = on != 0;
    return OODE_OK;
}

int oode_get_best_x(oode_handle *h, int64_t generation, double *x) {
    if (!h || !x) return fail(OODE_EINVAL, "oode_get_best_x: NULL argument");
    if (!h->inited) return fail(OODE_ESTATE, "oode_get_best_x: call oode_init first");
    if (generation > h->gen || generation < h->min_gen || generation <= h->gen - h->ring)
        return fail(OODE_EINVAL, "oode_get_best_x: generation is outside the history ring");
    CU(cudaSetDevice(h->cfg.device));
    const double *src = h->bestx_ring + (generation % h->ring) * (int64_t)h->Dl;
    if (h->G > 1) {
        // collective: every rank contributes its column slice (padded to the widest slice)
        CU(cudaMemcpyAsync(h->bx_send, src, h->Dl * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        NC(g_nccl.AllGather(h->bx_send, h->bx_all, (size_t)h->dl_max, ncclFloat64, h->comm, h->stream));
        for (int r = 0; r < h->G; ++r)
            CU(cudaMemcpyAsync(x + h->shard_col0[r], h->bx_all + (int64_t)r * h->dl_max, h->shard_dl[r] * sizeof(double),
                               cudaMemcpyDeviceToHost, h->stream));
        h->ctr.d2h_bytes += h->D * (int64_t)sizeof(double);
        CU(cudaStreamSynchronize(h->stream));
        return OODE_OK;
    }
    CU(cudaMemcpyAsync(x, src, h->Dl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->ctr.d2h_bytes += h->Dl * (int64_t)sizeof(double);
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

int oode_get_population(oode_handle *h, double *pop, double *vals) {
    if (!h) return fail(OODE_EINVAL, "oode_get_population: NULL handle");
    if (!h->inited) return fail(OODE_ESTATE, "oode_get_population: call oode_init first");
    CU(cudaSetDevice(h->cfg.device));
    const int64_t NP = h->NP;
    if (pop) {
        std::vector<uint8_t> sel(NP);
        CU(cudaMemcpyAsync(sel.data(), h->sel[h->cur], NP, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        // copy runs of rows that live in the same slot array
        int64_t i = 0;
        while (i < NP) {
            int64_t j = i + 1;
            while (j < NP && sel[j] == sel[i]) ++j;
            CU(cudaMemcpy2DAsync(pop + i * h->D + h->col0, h->D * sizeof(double), h->buf[sel[i]] + i * h->ld,
                                 h->ld * sizeof(double), h->Dl * sizeof(double), j - i, cudaMemcpyDeviceToHost, h->stream));
            i = j;
        }
    }
    if (vals) CU(cudaMemcpyAsync(vals, h->vals, NP * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->ctr.d2h_bytes += (pop ? NP * (int64_t)h->D * 8 + NP : 0) + (vals ? NP * 8 : 0);
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

int oode_set_population(oode_handle *h, const double *pop, const double *vals) {
    if (!h || !pop || !vals) return fail(OODE_EINVAL, "oode_set_population: NULL argument");
    if (!h->inited) return fail(OODE_ESTATE, "oode_set_population: call oode_init first");
    CU(cudaSetDevice(h->cfg.device));
    const int64_t NP = h->NP;
    CU(cudaMemcpy2DAsync(h->buf[0], h->ld * sizeof(double), pop + h->col0, h->D * sizeof(double), h->Dl * sizeof(double), NP,
                         cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->sel[0], 0, NP, h->stream));
    CU(cudaMemsetAsync(h->sel[1], 0, NP, h->stream));
    CU(cudaMemcpyAsync(h->vals, vals, NP * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->cur = 0;
    h->trusted_box = false;        // caller-supplied rows need not lie inside [lb, ub]: keep the general bound repair
    return OODE_OK;
}

// ---------------------------------------------------------------------------------------------
// Checkpoint / resume.  Blob = StateHeader, then (8-byte aligned, in this order): pop [NP][D] (current
// rows, slot-resolved), vals [NP], tvals [NP], cvals [NP] (constrained only), best_x [D], tbest_x [D],
// ib [NP] + r1, r2 [hndvi*NP] (int32), accept [NP], CPython generator state (CPYTHON stream only).
// ---------------------------------------------------------------------------------------------
struct StateHeader {
    uint64_t magic;
    int64_t NP, gen;
    int32_t D, hndvi, rng, objective, constrained, strategy_bits;
    double best_f;
};
constexpr uint64_t STATE_MAGIC = 0x4F4F444553544154ull;       // "OODESTAT"

static int64_t up8(int64_t n) { return (n + 7) & ~7ll; }

int64_t oode_state_size(oode_handle *h) {
    if (!h) return fail(OODE_EINVAL, "oode_state_size: NULL handle");
    const int64_t NP = h->NP, hn = NP * h->hndvi;
    int64_t n = up8(sizeof(StateHeader)) + NP * h->D * 8 + 2 * NP * 8 + (h->constrained ? NP * 8 : 0) + 2 * (int64_t)h->D * 8;
    n += up8((NP + 2 * hn) * 4) + up8(NP);
    if (h->cfg.rng == OODE_RNG_CPYTHON) n += up8(MT19937CPython::STATE_WORDS * 4);
    return n;
}

static int state_check(oode_handle *h, const char *who) {
    if (!h->inited) return fail(OODE_ESTATE, std::string(who) + ": call oode_init first");
    if (h->W > 1) return fail(OODE_EINVAL, std::string(who) + ": not available on a sharded handle");
    return OODE_OK;
}

int oode_save_state(oode_handle *h, void *buf) {
    if (!h || !buf) return fail(OODE_EINVAL, "oode_save_state: NULL argument");
    { int rc = state_check(h, "oode_save_state"); if (rc) return rc; }
    CU(cudaSetDevice(h->cfg.device));
    const int64_t NP = h->NP, hn = NP * h->hndvi;
    const int D = h->D;
    unsigned char *p = (unsigned char *)buf;
    StateHeader hd{};
    hd.magic = STATE_MAGIC; hd.NP = NP; hd.gen = h->gen; hd.D = D; hd.hndvi = h->hndvi; hd.rng = h->cfg.rng;
    hd.objective = h->cfg.objective; hd.constrained = h->constrained ? 1 : 0;
    hd.strategy_bits = (h->base_best ? 1 : 0) | (h->dir_best ? 2 : 0) | (h->const_F ? 4 : 0);
    double *pop = (double *)(p + up8(sizeof(StateHeader)));
    double *vals = pop + NP * D, *tvals = vals + NP, *q = tvals + NP;
    { int rc = oode_get_population(h, pop, vals); if (rc) return rc; }
    CU(cudaMemcpyAsync(tvals, h->tvals, NP * 8, cudaMemcpyDeviceToHost, h->stream));
    if (h->constrained) { CU(cudaMemcpyAsync(q, h->cvals, NP * 8, cudaMemcpyDeviceToHost, h->stream)); q += NP; }
    CU(cudaMemcpyAsync(q, h->best_x, D * 8, cudaMemcpyDeviceToHost, h->stream)); q += D;
    CU(cudaMemcpyAsync(q, h->tbest_x, D * 8, cudaMemcpyDeviceToHost, h->stream)); q += D;
    int32_t *idx = (int32_t *)q;
    CU(cudaMemcpyAsync(idx, h->ib, NP * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(idx + NP, h->r1, hn * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(idx + NP + hn, h->r2, hn * 4, cudaMemcpyDeviceToHost, h->stream));
    unsigned char *acc = (unsigned char *)idx + up8((NP + 2 * hn) * 4);
    CU(cudaMemcpyAsync(acc, h->accept, NP, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&hd.best_f, h->best_f, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->cfg.rng == OODE_RNG_CPYTHON) h->mt->save((uint32_t *)(acc + up8(NP)));
    memcpy(p, &hd, sizeof(hd));
    h->ctr.d2h_bytes += oode_state_size(h);
    return OODE_OK;
}

int oode_load_state(oode_handle *h, const void *buf) {
    if (!h || !buf) return fail(OODE_EINVAL, "oode_load_state: NULL argument");
    { int rc = state_check(h, "oode_load_state"); if (rc) return rc; }
    CU(cudaSetDevice(h->cfg.device));
    const int64_t NP = h->NP, hn = NP * h->hndvi;
    const int D = h->D;
    const unsigned char *p = (const unsigned char *)buf;
    StateHeader hd;
    memcpy(&hd, p, sizeof(hd));
    const int bits = (h->base_best ? 1 : 0) | (h->dir_best ? 2 : 0) | (h->const_F ? 4 : 0);
    if (hd.magic != STATE_MAGIC || hd.NP != NP || hd.D != D || hd.hndvi != h->hndvi || hd.rng != h->cfg.rng ||
        hd.objective != h->cfg.objective || hd.constrained != (h->constrained ? 1 : 0) || hd.strategy_bits != bits)
        return fail(OODE_EINVAL, "oode_load_state: the state was saved by a handle with another configuration");
    if (hd.gen < 0) return fail(OODE_EINVAL, "oode_load_state: corrupt state (negative generation)");
    const double *pop = (const double *)(p + up8(sizeof(StateHeader)));
    const double *vals = pop + NP * D, *tvals = vals + NP, *q = tvals + NP;
    const bool trusted = h->trusted_box;       // a saved state is a population this solver produced: still inside the box
    { int rc = oode_set_population(h, pop, vals); if (rc) return rc; }          // slot 0, selectors 0, cur = 0
    h->trusted_box = trusted;
    CU(cudaMemcpyAsync(h->tvals, tvals, NP * 8, cudaMemcpyHostToDevice, h->stream));
    if (h->constrained) { CU(cudaMemcpyAsync(h->cvals, q, NP * 8, cudaMemcpyHostToDevice, h->stream)); q += NP; }
    CU(cudaMemcpyAsync(h->best_x, q, D * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->bestx_ring + (hd.gen % h->ring) * (int64_t)h->Dl, q, D * 8, cudaMemcpyHostToDevice, h->stream));
    q += D;
    CU(cudaMemcpyAsync(h->tbest_x, q, D * 8, cudaMemcpyHostToDevice, h->stream)); q += D;
    const int32_t *idx = (const int32_t *)q;
    CU(cudaMemcpyAsync(h->ib, idx, NP * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->r1, idx + NP, hn * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->r2, idx + NP + hn, hn * 4, cudaMemcpyHostToDevice, h->stream));
    const unsigned char *acc = (const unsigned char *)idx + up8((NP + 2 * hn) * 4);
    CU(cudaMemcpyAsync(h->accept, acc, NP, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->best_f, &hd.best_f, 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->cfg.rng == OODE_RNG_CPYTHON) h->mt->load((const uint32_t *)(acc + up8(NP)));
    h->gen = hd.gen;
    h->min_gen = hd.gen;           // the history ring holds nothing older than the restored generation
    h->ctr.h2d_bytes += oode_state_size(h);
    return OODE_OK;
}

int oode_get_accept_mask(oode_handle *h, uint8_t *mask) {
    if (!h || !mask) return fail(OODE_EINVAL, "oode_get_accept_mask: NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(mask, h->accept, h->NP, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

int oode_get_trial_vals(oode_handle *h, double *tvals) {
    if (!h || !tvals) return fail(OODE_EINVAL, "oode_get_trial_vals: NULL argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(tvals, h->tvals, h->NP * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

int oode_get_rows(oode_handle *h, int32_t which, const int64_t *rows, int64_t n, double *out) {
    if (!h || !rows || !out || n < 0) return fail(OODE_EINVAL, "oode_get_rows: bad argument");
    if (which < OODE_ROWS_CURRENT || which > OODE_ROWS_PARENT) return fail(OODE_EINVAL, "oode_get_rows: unknown row version");
    if (!h->inited) return fail(OODE_ESTATE, "oode_get_rows: call oode_init first");
    if (n == 0) return OODE_OK;
    for (int64_t k = 0; k < n; ++k)
        if (rows[k] < 0 || rows[k] >= h->NP) return fail(OODE_EINVAL, "oode_get_rows: row index out of range");
    CU(cudaSetDevice(h->cfg.device));
    int64_t *d_rows = nullptr;
    double *d_out = nullptr;
    SCRATCH(0, d_rows, n);
    SCRATCH(1, d_out, n * (int64_t)h->Dl);
    CU(cudaMemcpyAsync(d_rows, rows, n * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 7) / 8, (int64_t)h->sm_count * 8));
    de_gather_rows_kernel<<<grid, 256, 0, h->stream>>>(h->buf[0], h->buf[1], h->sel[h->cur], h->accept, h->ld, h->Dl, which,
                                                       d_rows, n, d_out);
    CU(cudaGetLastError());
    h->ctr.gpu_launches++;
    CU(cudaMemcpy2DAsync(out + h->col0, h->D * sizeof(double), d_out, h->Dl * sizeof(double), h->Dl * sizeof(double), n,
                         cudaMemcpyDeviceToHost, h->stream));
    h->ctr.d2h_bytes += n * (int64_t)h->Dl * 8;
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

int oode_get_donor_indices(oode_handle *h, int64_t *ib, int64_t *r1, int64_t *r2) {
    if (!h || !ib || !r1 || !r2) return fail(OODE_EINVAL, "oode_get_donor_indices: NULL argument");
    if (!h->inited) return fail(OODE_ESTATE, "oode_get_donor_indices: call oode_init first");
    CU(cudaSetDevice(h->cfg.device));
    std::vector<int32_t> tmp(3 * (size_t)h->NP);
    CU(cudaMemcpyAsync(tmp.data(), h->ib, h->NP * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(tmp.data() + h->NP, h->r1, h->NP * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(tmp.data() + 2 * h->NP, h->r2, h->NP * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < h->NP; ++i) { ib[i] = tmp[i]; r1[i] = tmp[h->NP + i]; r2[i] = tmp[2 * h->NP + i]; }
    h->ctr.d2h_bytes += 3 * h->NP * (int64_t)sizeof(int32_t);
    return OODE_OK;
}

// f of n caller-supplied rows that already sit in device memory ([n][h->ld], 16-byte aligned rows) into df[n]: the
// trial kernel's MODE_EVAL (numpy's summation order) + the fold.  Used by oode_eval_points and by the galileo solver.
static int eval_device_rows(oode_handle *h, const double *dx, int64_t n, double *df) {
    const int nleaves = (int)h->plan.leaves.size();
    double *dpart = nullptr;
    uint8_t *dsel = nullptr;
    SCRATCH(1, dpart, (size_t)n * nleaves * h->nslots);
    SCRATCH(3, dsel, (size_t)n);
    CU(cudaMemsetAsync(dsel, 0, n, h->stream));
    GenParams gp = gen_params(h, 0);
    gp.buf0 = const_cast<double *>(dx); gp.buf1 = const_cast<double *>(dx);
    gp.sel_cur = dsel;
    gp.partials = dpart;
    gp.row0 = 0; gp.nrows = (int)n;
    gp.part_limit = n * (int64_t)nleaves * h->nslots;
    gp.NP = (int)std::max<int64_t>(n, h->NP);      // MODE_EVAL reads no donor rows; the row index checks use it
    CU(launch_trial(h, MODE_EVAL, gp, true));
    CU(h->obj.finalize(h->stream, part_layout(h, dpart, n), h->prog, (int)h->plan.prog.size(), h->D,
                                       h->cfg.invert, n, df));
    h->ctr.gpu_launches++;
    return OODE_OK;
}

int oode_eval_points(oode_handle *h, const double *x, int64_t n, double *f) {
    if (!h || !x || !f || n < 0) return fail(OODE_EINVAL, "oode_eval_points: bad argument");
    if (n == 0) return OODE_OK;
    if (h->G > 1) return fail(OODE_EINVAL, "oode_eval_points is not available on a sharded handle");
    CU(cudaSetDevice(h->cfg.device));
    double *dx = nullptr, *df = nullptr;
    SCRATCH(0, dx, (size_t)n * h->ld);
    SCRATCH(2, df, (size_t)n);
    CU(cudaMemcpy2DAsync(dx, h->ld * sizeof(double), x, h->D * sizeof(double), h->D * sizeof(double), n,
                         cudaMemcpyHostToDevice, h->stream));
    const int rc = eval_device_rows(h, dx, n, df);
    if (rc) return rc;
    CU(cudaMemcpyAsync(f, df, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

int oode_crossover_threshold(double cross_rate, uint32_t *threshold, int32_t *keep_all) {
    if (!threshold || !keep_all) return fail(OODE_EINVAL, "oode_crossover_threshold: NULL argument");
    // uc = w * 2^-32 exactly, so uc > Cr <=> w > Cr * 2^32 <=> w > floor(Cr * 2^32) for an integer w.
    // Cr >= 1 (or NaN): never true -> threshold 0xFFFFFFFF; Cr < 0: always true -> keep_all.
    *keep_all = cross_rate < 0.0 ? 1 : 0;
    *threshold = (cross_rate >= 0.0 && cross_rate < 1.0) ? (uint32_t)std::floor(cross_rate * 4294967296.0) : 0xFFFFFFFFu;
    return OODE_OK;
}

int oode_cpython_stream(uint64_t seed, uint32_t below, int64_t n_ints, int64_t *ints, int64_t n_doubles, double *doubles) {
    if ((n_ints > 0 && (!ints || below < 1)) || (n_doubles > 0 && !doubles)) return fail(OODE_EINVAL, "oode_cpython_stream: bad argument");
    MT19937CPython mt(seed);
    for (int64_t i = 0; i < n_ints; ++i) ints[i] = (int64_t)mt.randbelow(below);
    for (int64_t i = 0; i < n_doubles; ++i) doubles[i] = mt.random();
    return OODE_OK;
}

void oode_release_cached_memory(void) {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    release_cached_blocks();
}

int oode_exchange_mode(oode_handle *h) {
    if (!h) return fail(OODE_EINVAL, "oode_exchange_mode: NULL handle");
    if (h->rows) return OODE_XCHG_PEER;
    return h->G <= 1 ? OODE_XCHG_NONE : (h->p2p ? OODE_XCHG_PEER : OODE_XCHG_NCCL);
}

const char *oode_trial_kernel_name(oode_handle *h) {
    if (!h) return "";
    switch (h->form) {
        case FORM_ROWS: return h->wpw == 1 ? "de_trial_rows_kernel<wpw=1>" : "de_trial_rows_kernel<wpw=2>";
        case FORM_TMA: return "de_trial_tma_kernel";
        case FORM_TMA2: return "de_trial_tma2_kernel";
        default: return "de_trial_kernel";
    }
}

int oode_get_counters(oode_handle *h, oode_counters *out) {
    if (!h || !out) return fail(OODE_EINVAL, "oode_get_counters: NULL argument");
    *out = h->ctr;
    if (h->p2p && !h->rows && h->wait_ns) {      // column shards: time K2b's first block spent waiting for the other ranks
        unsigned long long ns = 0;
        CU(cudaSetDevice(h->cfg.device));
        CU(cudaMemcpyAsync(&ns, h->wait_ns, sizeof(ns), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        out->exchange_wait_ms = (double)ns * 1e-6;
    }
    out->bytes_algorithmic = h->ctr.trial_evals * (40ll * h->D + 16);
    return OODE_OK;
}

}  // extern "C"

// the sibling population solver behind the same boundary (include/ooga.h)
#include "ooga_host.cuh"
