// libs4mpi: C-ABI (include/s4mpi.h) and host-side orchestration of the kernels in
// s4_kernels.cuh.  One s4_ctx = one GPU + one stream; all work is stream-ordered.
// There is no CPU fallback anywhere in this file: every entry point either runs
// the CUDA path or reports an error.
#include "../../include/s4mpi.h"
#include "s4_kernels.cuh"
#include "s4_sssp_group.cuh"
#include "s4_walk_codes.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

using namespace s4;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(e__ == cudaErrorMemoryAllocation ? S4_ERR_OOM : S4_ERR_CUDA, "%s:%d %s -> %s", \
                        __FILE__, __LINE__, #expr, cudaGetErrorString(e__));                           \
    } while (0)

#define CHECK_ARG(cond, msg)                          \
    do {                                              \
        if (!(cond)) return fail(S4_ERR_ARG, "%s", msg); \
    } while (0)

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------
struct Options {
    double delta_factor = 1.5;
    int block_threads = 0;          // 0 = from the graph (mean frontier width ~ V / hop depth)
    int ctas_per_sm = 0;            // 0 = as many CTAs as keep 1024 threads per SM
    int near_smem_entries = 0;      // 0 = 6 entries per thread
    int64_t queue_entries = 0;      // 0 = auto (from V)
    int64_t batch_roots = 0;        // 0 = auto (from pool_bytes)
    double pool_bytes = 0.0;        // cap on the bytes of per-tree distance arrays; 0 = 55 % of the free device memory
    int root_mode = S4_ROOT_ORIGIN;
    int walk_tile = 0;              // lanes per walking trip; 0 = from the graph's degrees
    int lanes = 2;                  // concurrent (SSSP launch -> walk) lanes per day; 1 = strictly serial kernels
    int adaptive_delta = 1;         // widen / narrow the bucket with the wave
    int walk_ell = 1;               // the walk reads fixed-stride rows (one dependent stage less per hop); 0: CSR only
    int early_advance = 2;          // admit the next bucket when a round has fewer entries than this many eighths of a CTA
                                    // (0 = only when the near queue is empty); measured on cfg3 and a planar graph
    int preread = 2;                // 1: filtering read, then returning atomicMin;
                                    // 2: filtering read decides the push, minimum sent as RED (no wait)
    int group_lanes = 8;            // lanes that expand one node for a whole group of trees (4 | 8); 0: one tree per CTA
    int group_words = 0;            // distance words per lane (1 | 2 | 4): trees per group = group_lanes * group_words - 1;
                                    // 0 = 4 (31 trees, rows of 256 bytes) when twice as many slots as SMs fit in the pool, else 2
    double lag_factor = 1.0;        // source-batched kernel: hold a node back by this multiple of the root-to-root distance
    int group_log = 0;              // diagnostics: keep per-group cycles / rounds / entries of the last day (s4_graph_group_log)
    int min_batches = 4;            // a day is cut into at least this many (SSSP launch -> walk) batches when lanes = 2
    int walk_codes = 2;             // 1: resolve every (node, tree) predecessor once (2-bit codes in the rows' meta words), then
                                    // one-lane walks; 0: walks derive predecessors on the fly; 2 = codes when the day has at
                                    // least 150 trips per tree (the streaming pass costs about as much as 100 trips per tree)
    int l2_persist = 0;             // 1: keep the fixed-stride arc rows (64 B per node, read by every expansion and every
                                    // hop) resident in L2 with an access-policy window on both lanes' streams
    int balance_waves = 1;          // fewer trees per group when that fills the last wave of a launch (few groups per GPU)
    int cluster = 0;                // CTAs of a thread-block cluster that share one group of trees (1 | 2 | 4 | 8); 0 = from
                                    // the number of slots that fit in the pool (large graphs: fewer slots than SMs)
};

// how the trees of a launch are laid out in the distance pool
struct Layout {
    int K = 1;       // trees per slot
    int RW = 1;      // words per node row
};
static Layout layout_of(const Options &o, int words)
{
    Layout l;
    if (o.group_lanes > 0) { l.RW = o.group_lanes * words; l.K = l.RW - 1; }
    return l;
}

// Lifetime: handles are reference counted (sim -> graph -> ctx), so the host may
// release them in any order (Python finalisers run in arbitrary order).
struct s4_ctx {
    int refs = 1;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t walk_stream = nullptr;     // second lane of the day loop (see s4_sim_step)
    cudaEvent_t sssp_done[2] = {nullptr, nullptr}, walk_done[2] = {nullptr, nullptr};
    cudaDeviceProp prop;
    Options opt;
    void *nccl_comm = nullptr;      // ncclComm_t of the rank-level exchange (s4_comm_init), or null: single rank
    void *stage_host = nullptr;     // pinned staging buffer + device scratch of s4_comm_allreduce_host
    void *stage_dev = nullptr;
    size_t stage_bytes = 0;
    int comm_ranks = 1, comm_rank = 0;
    bool l2_limit_set = false, l2_window_on = false;
};

// NCCL entry points, bound at run time (see "rank-level exchange" below)
struct NcclId { char internal[128]; };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(&p, count * sizeof(T));
    }
    cudaError_t ensure(size_t count) { return count <= n && p ? cudaSuccess : alloc(count); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

struct EvPair {
    cudaEvent_t a = nullptr, b = nullptr;
};

struct s4_graph {
    int refs = 1;
    s4_ctx *ctx = nullptr;
    int32_t V = 0;
    int64_t S = 0, A = 0;
    DevBuf<u32> row_ptr;       // V+1
    DevBuf<uint2> arc_cs;      // A  {head, street}
    DevBuf<uint2> ell_cs;      // V*4 {head | overflow flag, street}; padding {self, ~0}
    DevBuf<double> length;     // S
    DevBuf<u32> orig;          // device number -> caller's dense id (empty = identity numbering)
    std::vector<int32_t> rank; // caller's dense id -> device number (empty = identity)
    std::vector<int32_t> max_speed0;   // host copy of the uploaded limits
    std::vector<u32> row_ptr_h, col_h;  // host copy of the CSR (device numbering): root clustering at simulation setup
    std::vector<u32> gorder_h, ginv_h;  // grouping order of tree roots: device number -> position / position -> device number
    double mean_ecc = 0.0;              // mean of ecc_h
    std::vector<int32_t> ecc_h;         // hops to the farther end of the graph's longest BFS path (depth of a tree rooted here)
    DevBuf<u32> gorder, ginv;
    // connected components (host): algorithmic arcs / nodes of a tree rooted at r
    std::vector<int32_t> comp_of;
    std::vector<int64_t> comp_arcs, comp_nodes;
    int auto_block = 256;      // CTA width chosen from the graph's shape (see graph upload)
    int auto_walk_tile = 8;    // lanes per walking trip
    int auto_words = 0;        // distance words per lane chosen for this graph (0 = not decided yet)
    // scratch shared by every simulation on this graph (sized lazily)
    DevBuf<u64> pool;          // slots * V
    int64_t pool_slots = 0;
    DevBuf<u32> ws_v;          // per-CTA frontier workspace (entry words / distances), two lanes
    DevBuf<u64> ws_d;
    int ws_ctas = 0;
    bool ws_auto = false;      // workspace sized by the library (not by the "queue_entries" option)
    u32 near_gcap = 0, far_cap = 0;
    DevBuf<u32> work_counters;
    DevBuf<u32> cluster_ctl;   // control words of the clusters (cluster > 1), two lanes
    DevBuf<u32> glog;          // diagnostics of the source-batched kernel: 4 words per group of the last day
    int64_t glog_groups = 0;
    // single-source API scratch
    DevBuf<Arc> arcs_tmp, ell_tmp;
    DevBuf<double> w_tmp;
    DevBuf<int32_t> pred_tmp;
    DevBuf<double> dist_tmp;
    DevBuf<int32_t> root_tmp;
    DevBuf<u64> ctr_tmp;
    DevBuf<double> wsum_tmp;
};

struct s4_sim {
    s4_graph *g = nullptr;
    int64_t T = 0, D = 0;
    int root_mode = S4_ROOT_ORIGIN;
    double jam = 0.0;
    Phys ph{};
    u32 volume = 1;
    // trips sorted by root, roots in the grouping order of the graph (canonical form, independent of the kernel) ...
    DevBuf<int32_t> trip_target0;  // T
    DevBuf<u32> trip_root0;        // T, index into roots0
    DevBuf<int32_t> roots0;        // D
    DevBuf<int64_t> root_off0;     // D (+1), first trip of every root
    std::vector<int32_t> h_roots0;     // D
    std::vector<int64_t> h_root_off0;  // D+1
    // ... and as the kernels consume them: roots clustered into groups of `grouped_K` trees (slots padded with -1),
    // trips in the order of those slots
    int grouped_K = 0;
    int grouped_fill = 0;          // trees actually put into a group (<= grouped_K)
    int grouped_fill0 = -1;        // ... as first computed from the launch geometry (before the whole-wave correction)
    bool grouped_longest_first = false;
    int64_t Dp = 0;                // slots = groups * grouped_K
    DevBuf<int32_t> trip_target;   // T
    DevBuf<u32> trip_root;         // T, index into roots
    DevBuf<int32_t> roots;         // Dp
    std::vector<int64_t> h_root_off;   // Dp+1
    int64_t alg_arcs_per_day = 0, alg_nodes_per_day = 0;
    DevBuf<u32> load, cum;
    u64 *hops_host = nullptr;      // pinned: the device's hop counter as of the end of the last completed day (heuristics only)
    u64 hops_seen = 0;             // ... and what it was a day earlier
    double hops_per_day = 0.0;     // hops walked on the last day whose counter has arrived
    DevBuf<uint2> deferred;        // T: trips the hot walk kernel hands to the slow one (zero-weight plateaus)
    DevBuf<u32> deferred_counts;   // one per batch of the day
    bool has_cum = false;
    DevBuf<int32_t> ms_vis, ms_ins;
    DevBuf<double> w_street;
    DevBuf<Arc> arcs;              // CSR records (walk, overflow arcs)
    DevBuf<Arc> ell;               // fixed-stride records (SSSP)
    DevBuf<Arc> ellw;              // fixed-stride records with street indices (walk); empty when walk_ell = 0
    DevBuf<double> wsum;           // bucket-width scale of the day: (median street weight) * S
    DevBuf<u32> w_hist;            // histogram of the day's weights (weights_street_kernel -> weight_scale_kernel)
    DevBuf<u64> counters;
    bool weights_valid = false;
    // sort scratch for K5
    DevBuf<u32> sort_keys, sort_vals, sort_keys_out, order;
    DevBuf<unsigned char> cub_tmp;
    DevBuf<uint8_t> mask;
    // timing
    std::vector<EvPair> ev_w, ev_sssp, ev_walk, ev_adapt, ev_step;
    size_t used_w = 0, used_sssp = 0, used_walk = 0, used_adapt = 0, used_step = 0;
    s4_counters acc{};
};

static int grid_for(int64_t n, int block, const s4_ctx *ctx)
{
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)ctx->prop.multiProcessorCount * 8;
    return (int)std::max<int64_t>(1, std::min(need, cap));
}

// ---------------------------------------------------------------------------
// misc entry points
// ---------------------------------------------------------------------------
extern "C" int s4_abi_version(void) { return S4_ABI_VERSION; }
extern "C" const char *s4_last_error(void) { return g_err; }

static void ctx_unref(s4_ctx *ctx);

extern "C" int s4_ctx_create(int device, void *stream, s4_ctx **out)
{
    CHECK_ARG(out, "s4_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(S4_ERR_CUDA, "no CUDA device (%s); libs4mpi has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    CHECK_ARG(device >= 0 && device < ndev, "s4_ctx_create: device out of range");
    CU(cudaSetDevice(device));
    s4_ctx *c = new (std::nothrow) s4_ctx();
    if (!c) return fail(S4_ERR_OOM, "host allocation failed");
    c->device = device;
    cudaError_t pe = cudaGetDeviceProperties(&c->prop, device);
    if (pe != cudaSuccess) { delete c; return fail(S4_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(pe)); }
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) { delete c; return fail(S4_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(se)); }
        c->own_stream = true;
    }
    bool ok = cudaStreamCreateWithFlags(&c->walk_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
        ok = cudaEventCreateWithFlags(&c->sssp_done[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->walk_done[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { ctx_unref(c); return fail(S4_ERR_CUDA, "cannot create the walk stream / events"); }
    *out = c;
    return S4_OK;
}

static void ctx_unref(s4_ctx *ctx)
{
    if (--ctx->refs > 0) return;
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm && g_nccl.CommDestroy) { cudaStreamSynchronize(ctx->stream); g_nccl.CommDestroy(ctx->nccl_comm); }
    if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
    if (ctx->stage_dev) cudaFree(ctx->stage_dev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->walk_stream) cudaStreamDestroy(ctx->walk_stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->sssp_done[i]) cudaEventDestroy(ctx->sssp_done[i]);
        if (ctx->walk_done[i]) cudaEventDestroy(ctx->walk_done[i]);
    }
    delete ctx;
}

extern "C" int s4_ctx_destroy(s4_ctx *ctx)
{
    if (!ctx) return S4_OK;
    ctx_unref(ctx);
    return S4_OK;
}

extern "C" int s4_ctx_sync(s4_ctx *ctx)
{
    CHECK_ARG(ctx, "s4_ctx_sync: ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return S4_OK;
}

static void option_slot(Options &o, const char *key, int *kind, void **raw)
{
    // kind: 0 double, 1 int, 2 int64
    struct Item { const char *k; int kind; void *p; };
    Item items[] = {
        {"delta_factor", 0, &o.delta_factor},       {"block_threads", 1, &o.block_threads},
        {"ctas_per_sm", 1, &o.ctas_per_sm},
        {"near_smem_entries", 1, &o.near_smem_entries}, {"queue_entries", 2, &o.queue_entries},
        {"batch_roots", 2, &o.batch_roots},         {"pool_bytes", 0, &o.pool_bytes},
        {"root_mode", 1, &o.root_mode},             {"walk_tile", 1, &o.walk_tile},
        {"preread", 1, &o.preread},                 {"lanes", 1, &o.lanes},
        {"adaptive_delta", 1, &o.adaptive_delta},   {"early_advance", 1, &o.early_advance},
        {"walk_ell", 1, &o.walk_ell},                {"group_lanes", 1, &o.group_lanes},
        {"group_words", 1, &o.group_words},         {"lag_factor", 0, &o.lag_factor},
        {"min_batches", 1, &o.min_batches},         {"group_log", 1, &o.group_log},
        {"cluster", 1, &o.cluster},                 {"balance_waves", 1, &o.balance_waves},
        {"l2_persist", 1, &o.l2_persist},           {"walk_codes", 1, &o.walk_codes},
    };
    for (auto &it : items)
        if (strcmp(it.k, key) == 0) { *kind = it.kind; *raw = it.p; return; }
    *raw = nullptr;
}

extern "C" int s4_ctx_set_option(s4_ctx *ctx, const char *key, double value)
{
    CHECK_ARG(ctx && key, "s4_ctx_set_option: NULL argument");
    int kind = 0;
    void *raw = nullptr;
    option_slot(ctx->opt, key, &kind, &raw);
    if (!raw) return fail(S4_ERR_ARG, "unknown option '%s'", key);
    Options trial = ctx->opt;
    option_slot(trial, key, &kind, &raw);
    if (kind == 0) *(double *)raw = value;
    else if (kind == 1) *(int *)raw = (int)value;
    else *(int64_t *)raw = (int64_t)value;
    const Options &t = trial;
    bool ok = t.delta_factor >= 0.0 && (t.block_threads == 0 || t.block_threads == 32 || t.block_threads == 64 || t.block_threads == 128 || t.block_threads == 256 ||
               t.block_threads == 512 || t.block_threads == 1024) &&
              t.ctas_per_sm >= 0 && t.ctas_per_sm <= 32 &&
              (t.near_smem_entries == 0 || (t.near_smem_entries >= 32 && t.near_smem_entries <= 9000)) && t.queue_entries >= 0 &&
              t.batch_roots >= 0 && t.pool_bytes >= 0.0 && t.root_mode >= 0 && t.root_mode <= 2 &&
              (t.walk_tile == 0 || t.walk_tile == 4 || t.walk_tile == 8 || t.walk_tile == 16 || t.walk_tile == 32) &&
              (t.preread == 1 || t.preread == 2) && (t.lanes == 1 || t.lanes == 2) &&
              t.early_advance >= 0 && t.early_advance <= 64 && (t.walk_ell == 0 || t.walk_ell == 1) &&
              (t.group_lanes == 0 || t.group_lanes == 4 || t.group_lanes == 8) &&
              (t.group_words == 0 || t.group_words == 1 || t.group_words == 2 || (t.group_words == 4 && t.group_lanes != 4)) &&
              t.lag_factor >= 0.0 && t.lag_factor <= 64.0 && t.min_batches >= 1 && t.min_batches <= 64 &&
              (t.cluster == 0 || t.cluster == 1 || t.cluster == 2 || t.cluster == 4 || t.cluster == 8) &&
              (t.l2_persist == 0 || t.l2_persist == 1) && (t.balance_waves == 0 || t.balance_waves == 1) &&
              t.walk_codes >= 0 && t.walk_codes <= 2;
    if (!ok) return fail(S4_ERR_ARG, "option '%s': value %g out of range", key, value);
    ctx->opt = trial;
    return S4_OK;
}

extern "C" int s4_ctx_get_option(s4_ctx *ctx, const char *key, double *value)
{
    CHECK_ARG(ctx && key && value, "s4_ctx_get_option: NULL argument");
    int kind = 0;
    void *raw = nullptr;
    option_slot(ctx->opt, key, &kind, &raw);
    if (!raw) return fail(S4_ERR_ARG, "unknown option '%s'", key);
    *value = kind == 0 ? *(double *)raw : kind == 1 ? (double)*(int *)raw : (double)*(int64_t *)raw;
    return S4_OK;
}

extern "C" int s4_ctx_device_info(s4_ctx *ctx, char *name_out, int name_cap, int *sm_count, uint64_t *hbm_bytes)
{
    CHECK_ARG(ctx, "s4_ctx_device_info: ctx is NULL");
    if (name_out && name_cap > 0) snprintf(name_out, (size_t)name_cap, "%s", ctx->prop.name);
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = (uint64_t)ctx->prop.totalGlobalMem;
    return S4_OK;
}

// ---------------------------------------------------------------------------
// rank-level exchange (streets4mpi.py:44-46, :97-98): NCCL, bound at run time.
// A process must hold ONE NCCL: under torch that is the libnccl torch already loaded (dlopen RTLD_NOLOAD finds
// it), otherwise $S4MPI_NCCL_LIB or the system libnccl.so.2.  Only the five entry points below are used; their
// signatures and the two enum values are NCCL's stable ABI (nccl.h: ncclUint32 = 3, ncclSum = 0, 128-byte id).
// ---------------------------------------------------------------------------
static int nccl_load()
{
    if (g_nccl.handle) return S4_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);           // the one this process already has (torch)
    const char *env = getenv("S4MPI_NCCL_LIB");
    if (!h && env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(S4_ERR_STATE, "NCCL is not available: %s", dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (int (*)(NcclId *))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclAllReduce");
    a.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString)
        return fail(S4_ERR_STATE, "libnccl lacks an entry point this library needs");
    g_nccl = a;
    return S4_OK;
}

#define NC(expr)                                                                                            \
    do {                                                                                                    \
        int r__ = (expr);                                                                                   \
        if (r__ != 0) return fail(S4_ERR_CUDA, "%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
    } while (0)

extern "C" int s4_comm_unique_id(void *id_out)
{
    CHECK_ARG(id_out, "s4_comm_unique_id: NULL argument");
    int rc = nccl_load();
    if (rc) return rc;
    NcclId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return S4_OK;
}

extern "C" int s4_comm_init(s4_ctx *ctx, int nranks, int rank, const void *unique_id)
{
    CHECK_ARG(ctx && nranks >= 1 && rank >= 0 && rank < nranks, "s4_comm_init: bad argument");
    CHECK_ARG(nranks == 1 || unique_id, "s4_comm_init: unique_id is NULL");
    if (ctx->nccl_comm) return fail(S4_ERR_STATE, "s4_comm_init: this context already has a communicator");
    ctx->comm_ranks = nranks;
    ctx->comm_rank = rank;
    if (nranks == 1) return S4_OK;                                  // the reference without mpiexec: nothing to exchange
    int rc = nccl_load();
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    NcclId id;
    memcpy(&id, unique_id, sizeof id);
    NC(g_nccl.CommInitRank(&ctx->nccl_comm, nranks, id, rank));
    return S4_OK;
}

extern "C" int s4_comm_free(s4_ctx *ctx)
{
    CHECK_ARG(ctx, "s4_comm_free: ctx is NULL");
    if (ctx->nccl_comm) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->stream));
        NC(g_nccl.CommDestroy(ctx->nccl_comm));
        ctx->nccl_comm = nullptr;
    }
    ctx->comm_ranks = 1;
    ctx->comm_rank = 0;
    return S4_OK;
}

// communicator.Allreduce(local, total, op=MPI.SUM) for a driver that keeps traffic_load on the host
// (streets4mpi.py:97-98 as written): pinned staging, H2D, ncclAllReduce, D2H.  In place on `inout`.
extern "C" int s4_comm_allreduce_host(s4_ctx *ctx, uint32_t *inout, int64_t n)
{
    CHECK_ARG(ctx && n >= 0 && (n == 0 || inout), "s4_comm_allreduce_host: bad argument");
    if (ctx->comm_ranks <= 1 || n == 0) return S4_OK;
    if (!ctx->nccl_comm) return fail(S4_ERR_STATE, "s4_comm_allreduce_host: %d ranks but no communicator (s4_comm_init)", ctx->comm_ranks);
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n * sizeof(uint32_t);
    if (bytes > ctx->stage_bytes) {
        if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
        if (ctx->stage_dev) cudaFree(ctx->stage_dev);
        ctx->stage_host = ctx->stage_dev = nullptr;
        ctx->stage_bytes = 0;
        CU(cudaHostAlloc(&ctx->stage_host, bytes, cudaHostAllocDefault));
        CU(cudaMalloc(&ctx->stage_dev, bytes));
        ctx->stage_bytes = bytes;
    }
    memcpy(ctx->stage_host, inout, bytes);
    CU(cudaMemcpyAsync(ctx->stage_dev, ctx->stage_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NC(g_nccl.AllReduce(ctx->stage_dev, ctx->stage_dev, (size_t)n, /*ncclUint32*/ 3, /*ncclSum*/ 0, ctx->nccl_comm, ctx->stream));
    CU(cudaMemcpyAsync(ctx->stage_host, ctx->stage_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(inout, ctx->stage_host, bytes);
    return S4_OK;
}

extern "C" int s4_comm_size(s4_ctx *ctx, int *nranks, int *rank)
{
    CHECK_ARG(ctx, "s4_comm_size: ctx is NULL");
    if (nranks) *nranks = ctx->comm_ranks;
    if (rank) *rank = ctx->comm_rank;
    return S4_OK;
}

// ---------------------------------------------------------------------------
// graph
// ---------------------------------------------------------------------------
static int32_t uf_find(std::vector<int32_t> &parent, int32_t x)
{
    while (parent[x] != x) {
        parent[x] = parent[parent[x]];
        x = parent[x];
    }
    return x;
}

extern "C" int s4_graph_upload(s4_ctx *ctx, int32_t V, int64_t S, const int32_t *u_in, const int32_t *v_in,
                               const double *length, const int32_t *max_speed, const int32_t *node_rank, s4_graph **out)
{
    CHECK_ARG(ctx && out, "s4_graph_upload: NULL argument");
    const int32_t *u = u_in, *v = v_in;
    *out = nullptr;
    CHECK_ARG(V > 0, "s4_graph_upload: V must be positive");
    CHECK_ARG(S >= 0 && S < (int64_t)1 << 31, "s4_graph_upload: S out of range");
    CHECK_ARG(S == 0 || (u && v && length && max_speed), "s4_graph_upload: NULL street array");
    CU(cudaSetDevice(ctx->device));
    for (int64_t s = 0; s < S; ++s)
        if (u[s] < 0 || u[s] >= V || v[s] < 0 || v[s] >= V) return fail(S4_ERR_ARG, "street %lld: node id out of range", (long long)s);
    // optional device numbering (a permutation of 0..V-1, e.g. along a space-filling curve): everything
    // below is built in device numbers; ids only matter for the tie-break, which consults `orig`
    std::vector<int32_t> uu, vv;
    std::vector<u32> orig_h;
    if (node_rank) {
        orig_h.assign((size_t)V, 0xFFFFFFFFu);
        for (int32_t i = 0; i < V; ++i) {
            const int32_t r = node_rank[i];
            if (r < 0 || r >= V || orig_h[(size_t)r] != 0xFFFFFFFFu) return fail(S4_ERR_ARG, "node_rank is not a permutation of 0..V-1");
            orig_h[(size_t)r] = (u32)i;
        }
        uu.resize((size_t)S);
        vv.resize((size_t)S);
        for (int64_t s = 0; s < S; ++s) { uu[(size_t)s] = node_rank[u[s]]; vv[(size_t)s] = node_rank[v[s]]; }
        u = uu.data();
        v = vv.data();
    }
    std::vector<u32> row_ptr((size_t)V + 1, 0u);
    for (int64_t s = 0; s < S; ++s) {
        if (!(length[s] >= 0.0) || std::isinf(length[s])) return fail(S4_ERR_ARG, "street %lld: length must be finite and >= 0", (long long)s);
        if (max_speed[s] < 1) return fail(S4_ERR_ARG, "street %lld: max_speed must be >= 1", (long long)s);
        ++row_ptr[(size_t)u[s] + 1];
        if (u[s] != v[s]) ++row_ptr[(size_t)v[s] + 1];
    }
    uint64_t total = 0;
    for (int32_t i = 0; i < V; ++i) {
        total += row_ptr[(size_t)i + 1];
        if (total >= 0xFFFFFFFFull) return fail(S4_ERR_ARG, "too many arcs for 32-bit offsets");
        row_ptr[(size_t)i + 1] = (u32)total;
    }
    const int64_t A = (int64_t)total;
    std::vector<uint2> arc_cs((size_t)A);
    {
        std::vector<u32> fill(row_ptr.begin(), row_ptr.end() - 1);
        for (int64_t s = 0; s < S; ++s) {                        // street_index order inside every row
            arc_cs[fill[(size_t)u[s]]++] = make_uint2((u32)v[s], (u32)s);
            if (u[s] != v[s]) arc_cs[fill[(size_t)v[s]]++] = make_uint2((u32)u[s], (u32)s);
        }
    }
    // fixed-stride table for the SSSP kernel: first kEllWidth arcs of every row
    std::vector<uint2> ell_cs((size_t)V * kEllWidth);
    for (int32_t n = 0; n < V; ++n) {
        const u32 rb = row_ptr[(size_t)n], deg = row_ptr[(size_t)n + 1] - rb;
        for (u32 k = 0; k < (u32)kEllWidth; ++k)
            ell_cs[(size_t)n * kEllWidth + k] = k < deg ? arc_cs[rb + k] : make_uint2((u32)n, 0xFFFFFFFFu);
        if (deg > (u32)kEllWidth) ell_cs[(size_t)n * kEllWidth].x |= kOvfBit;
    }
    // slot of the reverse arc in the head's row (7 = not among its first kEllWidth arcs), packed above the street
    if (S >= ((int64_t)1 << 29) || V >= (1 << 28)) return fail(S4_ERR_ARG, "graph too large for the packed frontier entries");
    for (int32_t n = 0; n < V; ++n) {
        for (u32 k = 0; k < (u32)kEllWidth; ++k) {
            uint2 &slot = ell_cs[(size_t)n * kEllWidth + k];
            if (slot.y == 0xFFFFFFFFu) continue;
            const u32 h = slot.x & ~kOvfBit, street = slot.y & 0x1FFFFFFFu;
            u32 rev = 7u;
            for (u32 j = 0; j < (u32)kEllWidth; ++j) {
                const uint2 &back = ell_cs[(size_t)h * kEllWidth + j];
                if (back.y != 0xFFFFFFFFu && (back.x & ~kOvfBit) == (u32)n && (back.y & 0x1FFFFFFFu) == street) { rev = j; break; }
            }
            slot.y = street | (rev << 29);
        }
    }
    s4_graph *g = new (std::nothrow) s4_graph();
    if (!g) return fail(S4_ERR_OOM, "host allocation failed");
    g->row_ptr_h = row_ptr;
    g->col_h.resize((size_t)A);
    for (int64_t e = 0; e < A; ++e) g->col_h[(size_t)e] = arc_cs[(size_t)e].x;
    g->ctx = ctx;
    ++ctx->refs;
    g->V = V;
    g->S = S;
    g->A = A;
    g->max_speed0.assign(max_speed, max_speed + S);
    if (node_rank) g->rank.assign(node_rank, node_rank + V);
    // connected components -> algorithmic work of a tree (each arc of the
    // root's component is scanned exactly once by any exact SSSP)
    {
        std::vector<int32_t> parent((size_t)V);
        std::iota(parent.begin(), parent.end(), 0);
        for (int64_t s = 0; s < S; ++s) {
            int32_t a = uf_find(parent, u[s]), b = uf_find(parent, v[s]);
            if (a != b) parent[(size_t)std::max(a, b)] = std::min(a, b);
        }
        g->comp_of.resize((size_t)V);
        g->comp_arcs.assign((size_t)V, 0);
        g->comp_nodes.assign((size_t)V, 0);
        for (int32_t i = 0; i < V; ++i) {
            int32_t r = uf_find(parent, i);
            g->comp_of[(size_t)i] = r;
            g->comp_nodes[(size_t)r] += 1;
            g->comp_arcs[(size_t)r] += row_ptr[(size_t)i + 1] - row_ptr[(size_t)i];
        }
    }
    // Shape of a shortest-path wave on this graph: two BFS sweeps give the hop depth; V / depth is the mean
    // number of nodes a tree has in flight per relaxation round.  Wide waves (grids: ~V^(1/2)) want a wide
    // CTA per tree, long thin ones (chains of degree-2 nodes) want many small CTAs in flight instead.
    {
        std::vector<int32_t> level((size_t)V, -1), queue((size_t)V);
        auto bfs = [&](int32_t src, int32_t *far_node, int64_t *reached) {
            std::fill(level.begin(), level.end(), -1);
            int64_t head = 0, tail = 0;
            queue[(size_t)tail++] = src;
            level[(size_t)src] = 0;
            int32_t last = src;
            while (head < tail) {
                const int32_t x = queue[(size_t)head++];
                last = x;
                for (u32 e = row_ptr[(size_t)x]; e < row_ptr[(size_t)x + 1]; ++e) {
                    const int32_t y = (int32_t)arc_cs[e].x;
                    if (level[(size_t)y] < 0) { level[(size_t)y] = level[(size_t)x] + 1; queue[(size_t)tail++] = y; }
                }
            }
            *far_node = last;
            *reached = tail;
            return level[(size_t)last];
        };
        int32_t start = 0;
        int64_t best = 0;
        for (int32_t i = 0; i < V; ++i)
            if (g->comp_nodes[(size_t)i] > best) { best = g->comp_nodes[(size_t)i]; start = i; }   // root of the largest component
        int32_t far1 = start, far2 = start;
        int64_t reached = 1;
        bfs(start, &far1, &reached);
        const int32_t depth = std::max<int32_t>(1, bfs(far1, &far2, &reached));
        const double width = (double)reached / (double)depth;
        int blk = 32;
        while (blk < 256 && (double)blk * 2.0 <= width) blk *= 2;
        g->auto_block = blk;
        int64_t wide = 0;
        for (int32_t i = 0; i < V; ++i) wide += (row_ptr[(size_t)i + 1] - row_ptr[(size_t)i]) > 4u;
        g->auto_walk_tile = wide * 2 > (int64_t)V ? 16 : 8;      // 8 lanes cover a road node's arcs in one step (measured: 4 is slower)

        // Grouping order of tree roots for the source-batched SSSP kernel: the trees of one group should be close
        // in the GRAPH metric (their waves then pass every node together).  Coordinates may be missing or
        // misleading (a river, a motorway without exits), so the order comes from the graph itself: four landmark
        // BFS sweeps (a, b = the two ends found above; c = the node farthest from both; d = the node farthest
        // from c) embed every node at (hops(a) - hops(b), hops(c) - hops(d)) -- on a lattice that is its position
        // rotated by 45 degrees -- and the nodes are ordered along a Hilbert curve over that plane.
        {
            std::vector<int32_t> da(level);                          // level = hops from far1 (a)
            int32_t dummy = 0;
            bfs(far2, &dummy, &reached);
            std::vector<int32_t> db(level);
            // how deep a tree rooted at a node is, in hops (the two ends of the longest BFS path bound it from below):
            // groups of trees are started longest-first when a launch has only a few waves of them
            g->ecc_h.resize((size_t)V);
            double ecc_sum = 0.0;
            for (int32_t i = 0; i < V; ++i) {
                g->ecc_h[(size_t)i] = std::max(0, std::max(da[(size_t)i], db[(size_t)i]));
                ecc_sum += (double)g->ecc_h[(size_t)i];
            }
            g->mean_ecc = ecc_sum / (double)V;
            int32_t c_node = far1;
            int32_t best_min = -1;
            for (int32_t i = 0; i < V; ++i) {
                if (da[(size_t)i] < 0 || db[(size_t)i] < 0) continue;
                const int32_t m = std::min(da[(size_t)i], db[(size_t)i]);
                if (m > best_min) { best_min = m; c_node = i; }
            }
            int32_t d_node = c_node;
            bfs(c_node, &d_node, &reached);
            std::vector<int32_t> dc(level);
            bfs(d_node, &dummy, &reached);
            const std::vector<int32_t> &dd = level;
            const int64_t span1 = 2 * (int64_t)depth + 1, span2 = 2 * (int64_t)std::max<int32_t>(1, dc[(size_t)d_node]) + 1;
            std::vector<std::pair<uint64_t, int32_t>> keyed((size_t)V);
            for (int32_t i = 0; i < V; ++i) {
                uint64_t x = 0, y = 0;
                if (da[(size_t)i] >= 0 && db[(size_t)i] >= 0 && dc[(size_t)i] >= 0 && dd[(size_t)i] >= 0) {
                    const int64_t u1 = (int64_t)da[(size_t)i] - db[(size_t)i] + depth;                       // 0 .. 2 * depth
                    const int64_t u2 = (int64_t)dc[(size_t)i] - dd[(size_t)i] + (span2 - 1) / 2;
                    x = (uint64_t)std::min<int64_t>(65535, std::max<int64_t>(0, u1 * 65535 / std::max<int64_t>(1, span1 - 1)));
                    y = (uint64_t)std::min<int64_t>(65535, std::max<int64_t>(0, u2 * 65535 / std::max<int64_t>(1, span2 - 1)));
                }
                uint64_t key = 0;                                    // Hilbert index of (x, y) on the 65536^2 lattice
                for (uint64_t sft = 32768; sft > 0; sft >>= 1) {
                    const uint64_t rx = (x & sft) ? 1 : 0, ry = (y & sft) ? 1 : 0;
                    key += sft * sft * ((3 * rx) ^ ry);
                    if (ry == 0) {                                   // rotate the quadrant (bits above sft are not read again)
                        if (rx == 1) { x = 65535 - x; y = 65535 - y; }
                        std::swap(x, y);
                    }
                }
                keyed[(size_t)i] = std::make_pair(key, i);
            }
            std::sort(keyed.begin(), keyed.end());
            g->gorder_h.resize((size_t)V);
            g->ginv_h.resize((size_t)V);
            for (int32_t pos = 0; pos < V; ++pos) {
                g->ginv_h[(size_t)pos] = (u32)keyed[(size_t)pos].second;
                g->gorder_h[(size_t)keyed[(size_t)pos].second] = (u32)pos;
            }
        }
    }
    auto bail = [&](cudaError_t e, const char *what) {
        delete g;
        --ctx->refs;
        return fail(e == cudaErrorMemoryAllocation ? S4_ERR_OOM : S4_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    cudaError_t e;
    if ((e = g->row_ptr.alloc((size_t)V + 1)) != cudaSuccess) return bail(e, "alloc row_ptr");
    if ((e = g->arc_cs.alloc((size_t)std::max<int64_t>(A, 1))) != cudaSuccess) return bail(e, "alloc arcs");
    if ((e = g->length.alloc((size_t)std::max<int64_t>(S, 1))) != cudaSuccess) return bail(e, "alloc length");
    if ((e = g->ell_cs.alloc((size_t)V * kEllWidth)) != cudaSuccess) return bail(e, "alloc ell");
    if ((e = g->gorder.alloc((size_t)V)) != cudaSuccess || (e = g->ginv.alloc((size_t)V)) != cudaSuccess) return bail(e, "alloc grouping order");
    if ((e = cudaMemcpyAsync(g->gorder.p, g->gorder_h.data(), (size_t)V * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(e, "copy grouping order");
    if ((e = cudaMemcpyAsync(g->ginv.p, g->ginv_h.data(), (size_t)V * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(e, "copy grouping order");
    if (node_rank) {
        if ((e = g->orig.alloc((size_t)V)) != cudaSuccess) return bail(e, "alloc orig");
        if ((e = cudaMemcpyAsync(g->orig.p, orig_h.data(), (size_t)V * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(e, "copy orig");
    }
    if ((e = cudaMemcpyAsync(g->ell_cs.p, ell_cs.data(), (size_t)V * kEllWidth * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(e, "copy ell");
    if ((e = cudaMemcpyAsync(g->row_ptr.p, row_ptr.data(), ((size_t)V + 1) * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(e, "copy row_ptr");
    if (A > 0 && (e = cudaMemcpyAsync(g->arc_cs.p, arc_cs.data(), (size_t)A * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(e, "copy arcs");
    if (S > 0 && (e = cudaMemcpyAsync(g->length.p, length, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(e, "copy length");
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(e, "sync");
    *out = g;
    return S4_OK;
}

static void graph_unref(s4_graph *g)
{
    if (--g->refs > 0) return;
    s4_ctx *ctx = g->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete g;
    ctx_unref(ctx);
}

extern "C" int s4_graph_free(s4_graph *g)
{
    if (!g) return S4_OK;
    graph_unref(g);
    return S4_OK;
}

extern "C" int s4_graph_group_log(s4_graph *g, uint32_t *out, int64_t cap_words, int64_t *n_words)
{
    CHECK_ARG(g && n_words, "s4_graph_group_log: NULL argument");
    CU(cudaSetDevice(g->ctx->device));
    CU(cudaDeviceSynchronize());
    *n_words = g->glog_groups * 4;
    if (out && g->glog_groups > 0) {
        CHECK_ARG(cap_words >= *n_words, "s4_graph_group_log: buffer too small");
        CU(cudaMemcpy(out, g->glog.p, (size_t)*n_words * sizeof(u32), cudaMemcpyDeviceToHost));
    }
    return S4_OK;
}

extern "C" int s4_graph_dims(s4_graph *g, int32_t *V, int64_t *S, int64_t *A)
{
    CHECK_ARG(g, "s4_graph_dims: graph is NULL");
    if (V) *V = g->V;
    if (S) *S = g->S;
    if (A) *A = g->A;
    return S4_OK;
}

// ---------------------------------------------------------------------------
// SSSP launch plumbing
// ---------------------------------------------------------------------------
typedef void (*SsspKernel)(SsspParams);
typedef void (*GroupKernel)(GroupParams);

static SsspKernel pick_sssp(int block)
{
    switch (block) {
        case 32: return sssp_nearfar_kernel<32>;
        case 64: return sssp_nearfar_kernel<64>;
        case 128: return sssp_nearfar_kernel<128>;
        case 256: return sssp_nearfar_kernel<256>;
        case 512: return sssp_nearfar_kernel<512>;
        case 1024: return sssp_nearfar_kernel<1024>;
    }
    return nullptr;
}

template <int LP, int TL>
static GroupKernel pick_group_block(int block)
{
    switch (block) {
        case 32: return sssp_group_kernel<32, LP, TL>;
        case 64: return sssp_group_kernel<64, LP, TL>;
        case 128: return sssp_group_kernel<128, LP, TL>;
        case 256: return sssp_group_kernel<256, LP, TL>;
        case 512: return sssp_group_kernel<512, LP, TL>;
        case 1024: return sssp_group_kernel<1024, LP, TL>;
    }
    return nullptr;
}

static GroupKernel pick_group(int block, int lanes, int words, int cl)
{
    if (cl > 1) {
        // clusters: the wide-row layout on wide CTAs only (the case they exist for: huge graphs, few slots)
        if (lanes != 8 || words != 2) return nullptr;
        if (block == 1024) return cl == 2 ? sssp_group_kernel<1024, 8, 2, 2> : cl == 4 ? sssp_group_kernel<1024, 8, 2, 4> : sssp_group_kernel<1024, 8, 2, 8>;
        if (block == 512) return cl == 2 ? sssp_group_kernel<512, 8, 2, 2> : cl == 4 ? sssp_group_kernel<512, 8, 2, 4> : sssp_group_kernel<512, 8, 2, 8>;
        return nullptr;
    }
    if (lanes == 8 && words == 4) return pick_group_block<8, 4>(block);
    if (lanes == 8 && words == 1) return pick_group_block<8, 1>(block);
    if (lanes == 8 && words == 2) return pick_group_block<8, 2>(block);
    if (lanes == 4 && words == 1) return pick_group_block<4, 1>(block);
    if (lanes == 4 && words == 2) return pick_group_block<4, 2>(block);
    return nullptr;
}

struct SsspLaunch {
    SsspKernel fn = nullptr;        // one tree per CTA ...
    GroupKernel gfn = nullptr;      // ... or a group of trees per CTA
    Layout lay;
    int block = 0, grid = 0;
    int cl = 1;                     // CTAs per group (thread-block cluster size)
    size_t smem = 0;
    u32 scap = 0;
    int units() const { return grid / cl; }     // frontier workspaces a launch needs: one per CTA, or per cluster
    const void *func() const { return gfn ? (const void *)gfn : (const void *)fn; }
};

// words of one slot of the distance pool (the rows of lay.K trees)
static size_t slot_words(const s4_graph *g, const Layout &lay)
{
    return lay.RW == 1 ? (size_t)((g->V + 3) & ~3) : (size_t)g->V * (size_t)lay.RW;
}

// Device memory the scratch of a graph may take: the pool of per-tree distance arrays (55 % of what is free,
// or the "pool_bytes" option if that is smaller) and the per-CTA frontier workspace (15 %).
static int scratch_budget(s4_graph *g, double *pool_budget, double *ws_budget)
{
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    double avail = (double)free_b;
    if (g->pool.p) avail += (double)(g->pool.n * sizeof(u64));
    if (g->ws_v.p) avail += (double)(g->ws_v.n * sizeof(u32));
    if (g->ws_d.p) avail += (double)(g->ws_d.n * sizeof(u64));
    const Options &o = g->ctx->opt;
    *pool_budget = o.pool_bytes > 0.0 ? std::min(o.pool_bytes, 0.55 * avail) : 0.55 * avail;
    *ws_budget = 0.15 * avail;
    return S4_OK;
}

// Trees per group.  A group of 31 (four words per lane, rows of 256 bytes) costs 1.65x the time of a group of 15 on cfg3:
// 20 % less SM time per tree.  It needs twice the pool per slot, so it is taken when both lanes can still hold a slot
// per SM; larger graphs keep 15 (and, beyond that, clusters).  Decided once per graph and option setting.
static int group_words_for(s4_graph *g, int64_t D)
{
    const Options &o = g->ctx->opt;
    if (o.group_lanes == 0) return 1;
    if (o.group_words > 0) return o.group_words;
    if (o.group_lanes != 8) return 2;
    if (g->auto_words == 0) {
        double pool_budget = 0.0, ws_budget = 0.0;
        g->auto_words = 2;
        if (scratch_budget(g, &pool_budget, &ws_budget) == S4_OK) {
            const double per_slot = 8.0 * (double)g->V * 32.0;
            const double slots = o.batch_roots > 0 ? (double)(o.batch_roots / 31) : pool_budget / per_slot;
            if (slots >= 2.0 * (double)g->ctx->prop.multiProcessorCount) g->auto_words = 4;
        }
    }
    // ... and the day must have at least 2.5 waves of such groups: with fewer (several GPUs sharing a day) the smaller
    // groups pack the SMs better (cfg3 on 8 GPUs: 96 ms against 115 ms per day and GPU)
    if (g->auto_words == 4 && (D + 30) / 31 * 2 < 5 * (int64_t)g->ctx->prop.multiProcessorCount) return 2;
    return g->auto_words;
}
static Layout layout_for(s4_graph *g, int64_t D) { return layout_of(g->ctx->opt, group_words_for(g, D)); }

static int plan_sssp(s4_graph *g, SsspLaunch *L, int64_t D)
{
    const Options &o = g->ctx->opt;
    L->lay = layout_for(g, D);
    const int words = group_words_for(g, D);
    const bool grouped = o.group_lanes > 0;
    // CTA width from the graph's wave width (graph upload): a tree keeps about `auto_block` nodes in flight per
    // relaxation round; the source-batched kernel spends group_lanes lanes on every node it expands
    int block = o.block_threads ? o.block_threads : grouped ? std::min(1024, g->auto_block * o.group_lanes / 2) : g->auto_block;
    block = std::max(block, 32);
    const int sms = g->ctx->prop.multiProcessorCount;
    int cl = 1;
    {
        double pool_budget = 0.0, ws_budget = 0.0;
        int rc = scratch_budget(g, &pool_budget, &ws_budget);
        if (rc) return rc;
        const double per_slot = 8.0 * (double)slot_words(g, L->lay);
        const int64_t slots = o.batch_roots > 0 ? std::max<int64_t>(1, o.batch_roots / L->lay.K) : (int64_t)(pool_budget / per_slot);
        const int64_t per_lane = std::max<int64_t>(1, slots / (o.lanes >= 2 ? 2 : 1));
        // narrow CTAs only pay off when that many slots fit in flight
        if (!o.block_threads)
            while (block < 1024 && (int64_t)sms * std::min(32, 1024 / block) > per_lane) block *= 2;
        // fewer slots than SMs (10^7 .. 10^8 nodes): the CTAs of a cluster share a group
        if (grouped && o.group_lanes == 8 && words == 2 && block >= 512) {
            if (o.cluster > 0) cl = o.cluster;
            else while (cl < 8 && per_lane * (cl * 2) * (1024 / block) <= (int64_t)sms) cl *= 2;
        }
    }
    const int near_entries = cl > 1 ? 0 : o.near_smem_entries ? o.near_smem_entries : grouped ? std::max(256, 8 * block) : std::max(128, 6 * block);
    const int ctas = o.ctas_per_sm ? o.ctas_per_sm : std::min(32, std::max(1, 1024 / block));
    L->cl = cl;
    if (grouped) {
        L->gfn = pick_group(block, o.group_lanes, words, cl);
        L->fn = nullptr;
        if (!L->gfn) return fail(S4_ERR_ARG, "no source-batched SSSP kernel for block=%d lanes=%d words=%d cluster=%d", block, o.group_lanes, words, cl);
    } else {
        L->fn = pick_sssp(block);
        L->gfn = nullptr;
        if (!L->fn) return fail(S4_ERR_ARG, "no SSSP kernel for block=%d", block);
    }
    L->block = block;
    L->scap = (u32)near_entries;
    L->smem = grouped ? (size_t)2 * L->scap * sizeof(u32) : (size_t)2 * L->scap * (sizeof(u64) + sizeof(u32));
    CU(cudaFuncSetAttribute(L->func(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L->smem));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, L->func(), L->block, L->smem));
    if (occ < 1) return fail(S4_ERR_ARG, "SSSP kernel does not fit on an SM (block=%d, smem=%zu)", L->block, L->smem);
    L->grid = sms * std::min(occ, ctas);
    if (cl > 1) {
        // co-scheduled clusters the device can hold (GPC boundaries cost a few SMs)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(sms / cl * cl));
        cfg.blockDim = dim3((unsigned)L->block);
        cfg.dynamicSmemBytes = L->smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        CU(cudaOccupancyMaxActiveClusters(&n_clusters, L->func(), &cfg));
        if (n_clusters < 1) return fail(S4_ERR_ARG, "no cluster of %d CTAs of %d threads fits on this device", cl, L->block);
        L->grid = n_clusters * cl;
    }
    return S4_OK;
}

// pool + per-CTA queue workspace, sized from the options; slots_wanted (in slots of lay.K trees) caps the pool
static int ensure_scratch(s4_graph *g, const SsspLaunch &L, int64_t slots_wanted)
{
    const Options &o = g->ctx->opt;
    const int64_t V = g->V;
    double pool_budget = 0.0, ws_budget = 0.0;
    int rc = scratch_budget(g, &pool_budget, &ws_budget);
    if (rc) return rc;
    const size_t sw = slot_words(g, L.lay);
    const int64_t by_mem = (int64_t)(pool_budget / (8.0 * (double)sw));
    int64_t slots = o.batch_roots > 0 ? std::min<int64_t>(std::max<int64_t>(1, o.batch_roots / L.lay.K), by_mem) : by_mem;
    slots = std::max<int64_t>(1, std::min(slots, slots_wanted));
    if ((size_t)slots * sw > g->pool.n || !g->pool.p) CU(g->pool.alloc((size_t)slots * sw));
    g->pool_slots = (int64_t)(g->pool.n / sw);
    // per-CTA spill / far-pile capacity: a quarter of the nodes for a 256-wide CTA, proportionally less for
    // the narrow CTAs used on thin-wave graphs (an overflow is survivable: Bellman-Ford safety net)
    // The source-batched kernel queues a node at most once per round (stamps), so V entries can never overflow its near
    // queues: it gets that much when the budget below allows.
    int64_t q = o.queue_entries > 0 ? o.queue_entries
                : L.gfn         ? std::max<int64_t>(8192, V)
                                : std::max<int64_t>(8192, std::max<int64_t>(65536, V / 4) * std::min(L.block, 256) / 256);
    q = std::min<int64_t>(q, std::max<int64_t>(V, 1024) * 2);
    if (o.queue_entries <= 0) {
        // many narrow CTAs on a large graph: keep the whole workspace (two lanes x CTAs x four queues x 12 B) inside
        // its share of the device memory; a wave is far thinner than V / 4 on road graphs
        const double per_entry = 2.0 * (double)L.units() * 4.0 * (double)(sizeof(u32) + sizeof(u64));
        q = std::max<int64_t>(8192, std::min<int64_t>(q, (int64_t)(ws_budget / per_entry)));
    }
    // an automatically sized workspace is kept while it is large enough (its budget moves with the free memory)
    const bool explicit_q = o.queue_entries > 0;
    if (L.cl > 1) {
        CU(g->cluster_ctl.ensure((size_t)2 * L.units() * GC_WORDS));
    }
    if (L.units() > g->ws_ctas || !g->ws_v.p || (explicit_q ? (u32)q != g->near_gcap : !g->ws_auto)) {
        g->ws_v.release();
        g->ws_d.release();
        CU(g->ws_v.alloc((size_t)2 * L.units() * 4 * (size_t)q));   // two lanes (concurrent launches)
        CU(g->ws_d.alloc((size_t)2 * L.units() * 4 * (size_t)q));
        g->ws_ctas = L.units();
        g->near_gcap = (u32)q;
        g->far_cap = (u32)q;
        g->ws_auto = !explicit_q;
    }
    return S4_OK;
}

// rows: the fixed-stride table the kernel expands from (ell for one tree per CTA, ellw for the source-batched kernel)
static int launch_sssp(s4_graph *g, const SsspLaunch &L, const Arc *rows, const Arc *arcs, const int32_t *d_roots,
                       int n_roots, u32 *work_counter, const double *wsum, u64 *counters, int64_t first_slot = 0,
                       int lane = 0, cudaStream_t st = nullptr, int64_t first_group_of_day = 0)
{
    if (!st) st = g->ctx->stream;
    const Options &o = g->ctx->opt;
    const size_t sw = slot_words(g, L.lay);
    const size_t ws_lane = (size_t)lane * g->ws_ctas * 4 * (size_t)g->near_gcap;
    if (L.gfn) {
        GroupParams p;
        p.ellw = rows;
        p.row_ptr = g->row_ptr.p;
        p.arcs = arcs;
        p.pool = g->pool.p + (size_t)first_slot * sw;
        p.slot_stride = sw;
        p.roots = d_roots;
        p.n_roots = n_roots;
        p.n_groups = (n_roots + L.lay.K - 1) / L.lay.K;
        p.V = (u32)g->V;
        p.work_counter = work_counter;
        p.wsum = wsum;
        p.delta_scale = g->S > 0 ? o.delta_factor / (double)g->S : 0.0;
        p.lag_factor = o.lag_factor;
        p.ws_v = g->ws_v.p + ws_lane;
        p.ws_k = reinterpret_cast<u32 *>(g->ws_d.p) + ws_lane;
        p.near_scap = L.scap;
        p.near_gcap = g->near_gcap;
        p.far_cap = g->far_cap;
        p.adaptive = (u32)o.adaptive_delta;
        p.early_advance = (u32)(o.early_advance * (L.block / o.group_lanes) * L.cl / 8);
        p.counters = counters;
        p.group_log = o.group_log && g->glog.p ? g->glog.p + (size_t)first_group_of_day * 4 : nullptr;
        p.cluster_ctl = nullptr;
        if (L.cl > 1) {
            p.cluster_ctl = g->cluster_ctl.p + (size_t)lane * L.units() * GC_WORDS;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(std::min(L.units(), std::max(1, p.n_groups)) * L.cl));
            cfg.blockDim = dim3((unsigned)L.block);
            cfg.dynamicSmemBytes = L.smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)L.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CU(cudaLaunchKernelEx(&cfg, L.gfn, p));
            return S4_OK;
        }
        const int grid = std::min(L.grid, std::max(1, p.n_groups));
        L.gfn<<<grid, L.block, L.smem, st>>>(p);
        CU(cudaGetLastError());
        return S4_OK;
    }
    SsspParams p;
    p.ell = rows;
    p.row_ptr = g->row_ptr.p;
    p.arcs = arcs;
    p.slot_stride = sw;
    p.dist_pool = g->pool.p + (size_t)first_slot * p.slot_stride;
    p.roots = d_roots;
    p.n_roots = n_roots;
    p.V = (u32)g->V;
    p.work_counter = work_counter;
    p.wsum = wsum;
    p.delta_scale = g->S > 0 ? o.delta_factor / (double)g->S : 0.0;
    p.ws_v = g->ws_v.p + ws_lane;
    p.ws_d = g->ws_d.p + ws_lane;
    p.near_scap = L.scap;
    p.near_gcap = g->near_gcap;
    p.far_cap = g->far_cap;
    p.preread = (u32)o.preread;
    p.adaptive = (u32)o.adaptive_delta;
    p.early_advance = (u32)(o.early_advance * L.block / 8);
    p.counters = counters;
    const int grid = std::min(L.grid, std::max(1, n_roots));
    L.fn<<<grid, L.block, L.smem, st>>>(p);
    CU(cudaGetLastError());
    return S4_OK;
}

static int launch_pred(s4_graph *g, const Arc *arcs, const u64 *dist, u32 stride, u32 root, int32_t *pred, double *dist_out)
{
    const int tile = g->ctx->opt.walk_tile ? g->ctx->opt.walk_tile : g->auto_walk_tile;
    const int block = 256;
    const int grid = grid_for((int64_t)g->V * tile, block, g->ctx);
    cudaStream_t st = g->ctx->stream;
    switch (tile) {
        case 4: pred_kernel<4><<<grid, block, 0, st>>>(g->row_ptr.p, arcs, g->orig.p, dist, stride, (u32)g->V, root, pred, dist_out); break;
        case 8: pred_kernel<8><<<grid, block, 0, st>>>(g->row_ptr.p, arcs, g->orig.p, dist, stride, (u32)g->V, root, pred, dist_out); break;
        case 16: pred_kernel<16><<<grid, block, 0, st>>>(g->row_ptr.p, arcs, g->orig.p, dist, stride, (u32)g->V, root, pred, dist_out); break;
        default: pred_kernel<32><<<grid, block, 0, st>>>(g->row_ptr.p, arcs, g->orig.p, dist, stride, (u32)g->V, root, pred, dist_out); break;
    }
    CU(cudaGetLastError());
    return S4_OK;
}

// predecessor codes for the group slots of a batch, then one-lane walks; slow nodes go to walk_slow_kernel
static int launch_walk_codes(s4_graph *g, const WalkParams &p, int n_groups, int words, cudaStream_t st)
{
    CodeParams cp;
    cp.ellw = p.ellw;
    cp.pool = const_cast<u64 *>(p.dist_pool);
    cp.slot_stride = p.slot_stride;
    cp.n_groups = n_groups;
    cp.V = p.V;
    const int sms = g->ctx->prop.multiProcessorCount;
    const int64_t subwarps = (int64_t)n_groups * (int64_t)p.V;
    const int code_grid = (int)std::max<int64_t>(1, std::min<int64_t>((subwarps * 8 + 255) / 256, (int64_t)sms * 8));
    switch (words) {
        case 1: pred_code_kernel<1><<<code_grid, 256, 0, st>>>(cp); break;
        case 2: pred_code_kernel<2><<<code_grid, 256, 0, st>>>(cp); break;
        default: pred_code_kernel<4><<<code_grid, 256, 0, st>>>(cp); break;
    }
    CU(cudaGetLastError());
    const int grid = grid_for(p.t1 - p.t0, 256, g->ctx);
    walk_code_kernel<<<grid, 256, 0, st>>>(p);
    CU(cudaGetLastError());
    walk_slow_kernel<8><<<grid_for((p.t1 - p.t0) * 8, 256, g->ctx), 256, 0, st>>>(p);      // ties / plateaus / wide nodes: may be many
    CU(cudaGetLastError());
    return S4_OK;
}

static int launch_walk(s4_graph *g, const WalkParams &p, cudaStream_t st)
{
    const int tile = g->ctx->opt.walk_tile ? g->ctx->opt.walk_tile : g->auto_walk_tile;
    const int block = 256;
    const int grid = grid_for((p.t1 - p.t0) * tile, block, g->ctx);
    // behind it: the trips it stopped on a zero-weight plateau (none on positive weights: the kernel exits at once)
    const int slow_grid = g->ctx->prop.multiProcessorCount;
    switch (tile) {
        case 4: walk_kernel<4><<<grid, block, 0, st>>>(p); walk_slow_kernel<4><<<slow_grid, block, 0, st>>>(p); break;
        case 8: walk_kernel<8><<<grid, block, 0, st>>>(p); walk_slow_kernel<8><<<slow_grid, block, 0, st>>>(p); break;
        case 16: walk_kernel<16><<<grid, block, 0, st>>>(p); walk_slow_kernel<16><<<slow_grid, block, 0, st>>>(p); break;
        default: walk_kernel<32><<<grid, block, 0, st>>>(p); walk_slow_kernel<32><<<slow_grid, block, 0, st>>>(p); break;
    }
    CU(cudaGetLastError());
    return S4_OK;
}

extern "C" int s4_graph_sssp(s4_graph *g, const double *w_street, int32_t origin, double *dist_out, int32_t *pred_out)
{
    CHECK_ARG(g && w_street, "s4_graph_sssp: NULL argument");
    CHECK_ARG(origin >= 0 && origin < g->V, "s4_graph_sssp: origin out of range");
    s4_ctx *ctx = g->ctx;
    CU(cudaSetDevice(ctx->device));
    for (int64_t s = 0; s < g->S; ++s)
        if (!(w_street[s] >= 0.0)) return fail(S4_ERR_ARG, "street %lld: weight must be >= 0", (long long)s);
    SsspLaunch L;
    int rc = plan_sssp(g, &L, 1);
    if (rc) return rc;
    rc = ensure_scratch(g, L, 1);
    if (rc) return rc;
    CU(g->arcs_tmp.ensure((size_t)std::max<int64_t>(g->A, 1)));
    CU(g->ell_tmp.ensure((size_t)g->V * kEllWidth));
    CU(g->w_tmp.ensure((size_t)std::max<int64_t>(g->S, 1)));
    CU(g->pred_tmp.ensure((size_t)g->V));
    CU(g->dist_tmp.ensure((size_t)g->V));
    const int32_t origin_dev = g->rank.empty() ? origin : g->rank[(size_t)origin];
    CU(g->root_tmp.ensure(1));
    CU(g->ctr_tmp.ensure(C_COUNT));
    CU(g->wsum_tmp.ensure(1));
    CU(g->work_counters.ensure(1024));
    cudaStream_t st = ctx->stream;
    // bucket-width scale like the simulation's: (median positive weight) * S
    double wsum = 0.0;
    {
        std::vector<double> pos;
        pos.reserve((size_t)g->S);
        for (int64_t s = 0; s < g->S; ++s)
            if (w_street[s] > 0.0 && !std::isinf(w_street[s])) pos.push_back(w_street[s]);
        if (!pos.empty()) {
            std::nth_element(pos.begin(), pos.begin() + (pos.size() - 1) / 2, pos.end());
            wsum = pos[(pos.size() - 1) / 2] * (double)g->S;
        }
    }
    if (g->S > 0) CU(cudaMemcpyAsync(g->w_tmp.p, w_street, (size_t)g->S * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(g->wsum_tmp.p, &wsum, sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(g->root_tmp.p, &origin_dev, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(g->ctr_tmp.p, 0, C_COUNT * sizeof(u64), st));
    CU(cudaMemsetAsync(g->work_counters.p, 0, sizeof(u32), st));
    if (g->A > 0) {
        weights_arc_kernel<<<grid_for(g->A, 256, ctx), 256, 0, st>>>(g->A, g->arc_cs.p, g->w_tmp.p, g->arcs_tmp.p);
        CU(cudaGetLastError());
    }
    if (L.gfn)
        weights_ellw_kernel<<<grid_for((int64_t)g->V * kEllWidth, 256, ctx), 256, 0, st>>>((int64_t)g->V * kEllWidth, g->ell_cs.p,
                                                                                          g->w_tmp.p, g->ell_tmp.p);
    else
        weights_ell_kernel<<<grid_for((int64_t)g->V * kEllWidth, 256, ctx), 256, 0, st>>>((int64_t)g->V * kEllWidth, g->ell_cs.p,
                                                                                         g->w_tmp.p, g->ell_tmp.p);
    CU(cudaGetLastError());
    rc = launch_sssp(g, L, g->ell_tmp.p, g->arcs_tmp.p, g->root_tmp.p, 1, g->work_counters.p, g->wsum_tmp.p, g->ctr_tmp.p);
    if (rc) return rc;
    // predecessors and distances come back in the caller's numbering (the tree is word 0 of the slot's rows)
    rc = launch_pred(g, g->arcs_tmp.p, g->pool.p, (u32)L.lay.RW, (u32)origin_dev, g->pred_tmp.p, g->dist_tmp.p);
    if (rc) return rc;
    if (pred_out) CU(cudaMemcpyAsync(pred_out, g->pred_tmp.p, (size_t)g->V * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (dist_out) CU(cudaMemcpyAsync(dist_out, g->dist_tmp.p, (size_t)g->V * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return S4_OK;
}

// ---------------------------------------------------------------------------
// simulation
// ---------------------------------------------------------------------------
static int cub_sort_pairs(s4_ctx *ctx, DevBuf<unsigned char> &tmp, const u32 *kin, u32 *kout, const u32 *vin, u32 *vout, int64_t n)
{
    size_t bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (int)n, 0, 32, ctx->stream));
    CU(tmp.ensure(bytes + 16));
    CU(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kin, kout, vin, vout, (int)n, 0, 32, ctx->stream));
    return S4_OK;
}

// number of distinct values in a host int32 array (used only to resolve S4_ROOT_AUTO)
static int64_t count_distinct_host(const int32_t *a, int64_t n, int32_t V)
{
    std::vector<uint8_t> seen((size_t)V, 0);
    int64_t d = 0;
    for (int64_t i = 0; i < n; ++i)
        if (!seen[(size_t)a[i]]) { seen[(size_t)a[i]] = 1; ++d; }
    return d;
}

// allocation and initial state shared by the two constructors
static int sim_alloc(s4_graph *g, s4_sim *sim, int64_t T, double jam_tolerance, const s4_physics *phys, int mode)
{
    s4_ctx *ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t S = g->S, A = g->A;
    sim->g = g;
    sim->T = T;
    sim->jam = jam_tolerance;
    sim->ph = Phys{phys->car_length, phys->min_breaking_distance, phys->braking_deceleration};
    sim->volume = phys->trip_volume;
    sim->root_mode = mode;
    CU(sim->load.alloc((size_t)std::max<int64_t>(S, 1)));
    CU(sim->cum.alloc((size_t)std::max<int64_t>(S, 1)));
    CU(sim->ms_vis.alloc((size_t)std::max<int64_t>(S, 1)));
    CU(sim->ms_ins.alloc((size_t)std::max<int64_t>(S, 1)));
    CU(sim->w_street.alloc((size_t)std::max<int64_t>(S, 1)));
    CU(sim->arcs.alloc((size_t)std::max<int64_t>(A, 1)));
    CU(sim->wsum.alloc(1));
    CU(sim->counters.alloc(C_COUNT));
    CU(cudaMemsetAsync(sim->load.p, 0, (size_t)std::max<int64_t>(S, 1) * sizeof(u32), st));   // simulation.py:43
    CU(cudaMemsetAsync(sim->cum.p, 0, (size_t)std::max<int64_t>(S, 1) * sizeof(u32), st));
    CU(cudaMemsetAsync(sim->counters.p, 0, C_COUNT * sizeof(u64), st));
    CU(cudaMemsetAsync(sim->wsum.p, 0, sizeof(double), st));
    if (S > 0) {
        CU(cudaMemcpyAsync(sim->ms_vis.p, g->max_speed0.data(), (size_t)S * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(sim->ms_ins.p, g->max_speed0.data(), (size_t)S * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    return S4_OK;
}

// group T trips (device arrays of root / target, device numbering) by tree root (tripgenerator.py:37-42)
static int sim_group(s4_sim *sim, DevBuf<u32> &kin, DevBuf<u32> &vin)
{
    s4_graph *g = sim->g;
    s4_ctx *ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t T = sim->T;
    if (T == 0) {
        sim->h_root_off0.assign(1, 0);
        sim->h_root_off.assign(1, 0);
        CU(cudaStreamSynchronize(st));
        return S4_OK;
    }
    DevBuf<u32> kout, vout, flag, rank;
    DevBuf<unsigned char> tmp;
    CU(kout.alloc((size_t)T)); CU(vout.alloc((size_t)T));
    CU(flag.alloc((size_t)T)); CU(rank.alloc((size_t)T));
    const int grid = grid_for(T, 256, ctx);
    // trips are sorted by the position of their root in the grouping order (graph upload): the roots of a day then
    // come out in that order, and consecutive roots are close in the graph
    gather_u32_kernel<<<grid, 256, 0, st>>>(T, kin.p, g->gorder.p);
    CU(cudaGetLastError());
    int rc = cub_sort_pairs(ctx, tmp, kin.p, kout.p, vin.p, vout.p, T);
    if (rc) return rc;
    head_flag_kernel<<<grid, 256, 0, st>>>(T, (const int32_t *)kout.p, flag.p);
    CU(cudaGetLastError());
    size_t bytes = 0;
    CU(cub::DeviceScan::InclusiveSum(nullptr, bytes, flag.p, rank.p, (int)T, st));
    CU(tmp.ensure(bytes + 16));
    CU(cub::DeviceScan::InclusiveSum(tmp.p, bytes, flag.p, rank.p, (int)T, st));
    u32 D = 0;
    CU(cudaMemcpyAsync(&D, rank.p + (T - 1), sizeof(u32), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    sim->D = D;
    CU(sim->roots0.alloc((size_t)D));
    CU(sim->root_off0.alloc((size_t)D + 1));
    CU(sim->trip_root0.alloc((size_t)T));
    CU(sim->trip_target0.alloc((size_t)T));
    group_scatter_kernel<<<grid, 256, 0, st>>>(T, (const int32_t *)kout.p, flag.p, rank.p, g->ginv.p, sim->trip_root0.p, sim->roots0.p, sim->root_off0.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sim->trip_target0.p, vout.p, (size_t)T * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    sim->h_root_off0.resize((size_t)D + 1);
    sim->h_roots0.resize((size_t)D);
    CU(cudaMemcpyAsync(sim->h_root_off0.data(), sim->root_off0.p, (size_t)D * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(sim->h_roots0.data(), sim->roots0.p, (size_t)D * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    sim->h_root_off0[(size_t)D] = T;
    for (u32 i = 0; i < D; ++i) {
        const int32_t c = g->comp_of[(size_t)sim->h_roots0[i]];
        sim->alg_arcs_per_day += g->comp_arcs[(size_t)c];
        sim->alg_nodes_per_day += g->comp_nodes[(size_t)c];
    }
    return S4_OK;
}

// Cluster the day's roots into groups of K trees that are close in the graph (the source-batched kernel runs a group
// on one frontier: its trees should reach every node at about the same time).  Greedy: seeds in the grouping order of
// the graph; a breadth-first search from an unassigned seed collects the unassigned roots it meets until it has K of
// them or has visited 6 * K * V / D nodes (six times the ball that holds K roots on average); smaller groups are
// padded with -1.  O(V) node visits in total.
// K_fill <= K: roots collected per group (the rest of its K slots stays empty): fewer trees per group when that
// evens out the last wave of the launch (see s4_sim_step).
static void cluster_roots(const s4_graph *g, const std::vector<int32_t> &roots, int K, int K_fill, bool longest_first,
                          std::vector<int32_t> &slots, std::vector<int64_t> &slot_of)
{
    const int64_t D = (int64_t)roots.size(), V = g->V;
    slots.clear();
    slot_of.assign((size_t)D, -1);
    if (K <= 1) {
        slots.assign(roots.begin(), roots.end());
        for (int64_t i = 0; i < D; ++i) slot_of[(size_t)i] = i;
        return;
    }
    std::vector<int32_t> level((size_t)V, -1), queue, touched;
    std::vector<int64_t> idx_of((size_t)V, -1);                     // node -> index among the roots
    std::vector<int32_t> group_of((size_t)V, -1);                   // root node -> its group once assigned
    for (int64_t i = 0; i < D; ++i) idx_of[(size_t)roots[(size_t)i]] = i;
    auto bfs_reset = [&]() { for (int32_t x : touched) level[(size_t)x] = -1; queue.clear(); touched.clear(); };
    auto bfs_push = [&](int32_t y, int32_t lv) { level[(size_t)y] = lv; queue.push_back(y); touched.push_back(y); };
    auto expand = [&](int32_t x) {
        for (u32 e = g->row_ptr_h[(size_t)x]; e < g->row_ptr_h[(size_t)x + 1]; ++e) {
            const int32_t y = (int32_t)g->col_h[e];
            if (level[(size_t)y] < 0) bfs_push(y, level[(size_t)x] + 1);
        }
    };
    // How many nodes a search visits before it has collected K_fill roots.  With every group filled to the brim
    // (spare == false) the estimate is the global density, as in a whole day on one GPU.  When the groups keep spare
    // slots (one GPU's share of a day: the roots cover part of the graph only, at the day's density) it is measured on
    // a sample of seeds, the search stops at twice that ball, and what it leaves behind -- scattered roots that would
    // otherwise be swept into wide, slow groups -- joins the nearest group that still has room.
    const bool spare = K_fill * 5 <= K * 4;
    int64_t cap = std::max<int64_t>(64, 6 * (int64_t)K_fill * V / std::max<int64_t>(1, D));
    if (spare) {
        std::vector<int64_t> balls;
        const int64_t step = std::max<int64_t>(1, D / 48);
        for (int64_t i = 0; i < D && balls.size() < 48; i += step) {
            bfs_push(roots[(size_t)i], 0);
            int got = 0;
            size_t head = 0;
            for (; head < queue.size() && got < K_fill; ++head) {
                const int32_t x = queue[head];
                if (idx_of[(size_t)x] >= 0) ++got;
                expand(x);
            }
            balls.push_back((int64_t)head);
            bfs_reset();
        }
        std::nth_element(balls.begin(), balls.begin() + balls.size() / 2, balls.end());
        cap = std::max<int64_t>(64, 2 * balls[balls.size() / 2]);
    }
    std::vector<std::vector<int32_t>> members;
    std::vector<int32_t> seed_of, radius_of;
    for (int64_t i = 0; i < D; ++i) {
        const int32_t seed = roots[(size_t)i];
        if (group_of[(size_t)seed] >= 0) continue;                  // already in a group
        const int32_t gid = (int32_t)members.size();
        members.emplace_back();
        std::vector<int32_t> &mem = members.back();
        bfs_push(seed, 0);
        int32_t radius = 0;                                         // hops from the seed to the farthest root of the group
        for (size_t head = 0; head < queue.size() && (int)mem.size() < K_fill && (int64_t)head <= cap; ++head) {
            const int32_t x = queue[head];
            if (idx_of[(size_t)x] >= 0 && group_of[(size_t)x] < 0) {
                radius = level[(size_t)x];
                group_of[(size_t)x] = gid;
                mem.push_back(x);
            }
            expand(x);
        }
        bfs_reset();
        seed_of.push_back(seed);
        radius_of.push_back(radius);
    }
    if (spare) {
        // groups less than half full are dissolved, smallest first: each of their roots joins the nearest group that is
        // at least half full and has a free slot (within four search balls), or stays
        std::vector<int32_t> order(members.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return members[(size_t)x].size() < members[(size_t)y].size(); });
        for (int32_t gid : order) {
            if ((int)members[(size_t)gid].size() * 2 > K_fill) break;
            std::vector<int32_t> keep;
            for (int32_t r : members[(size_t)gid]) {
                bfs_push(r, 0);
                int32_t target = -1;
                for (size_t head = 0; head < queue.size() && (int64_t)head <= 4 * cap; ++head) {
                    const int32_t x = queue[head];
                    const int32_t h = group_of[(size_t)x];
                    if (h >= 0 && h != gid && (int)members[(size_t)h].size() * 2 > K_fill && (int)members[(size_t)h].size() < K) {
                        target = h;
                        radius_of[(size_t)h] = std::max(radius_of[(size_t)h], level[(size_t)x] + radius_of[(size_t)h] / 2);
                        break;
                    }
                    expand(x);
                }
                bfs_reset();
                if (target >= 0) { members[(size_t)target].push_back(r); group_of[(size_t)r] = target; }
                else keep.push_back(r);
            }
            members[(size_t)gid].swap(keep);
        }
    }
    // groups in launch order: as found (neighbours along the grouping curve) or longest-first -- a tree is as deep as
    // its root is far from the far end of the graph, and a spread-out group waits for its slowest tree
    std::vector<int32_t> order;
    for (size_t gid = 0; gid < members.size(); ++gid)
        if (!members[gid].empty()) order.push_back((int32_t)gid);
    if (longest_first) {
        auto score = [&](int32_t gid) {
            return (g->ecc_h.empty() ? 0 : (int64_t)g->ecc_h[(size_t)seed_of[(size_t)gid]]) + 2 * (int64_t)radius_of[(size_t)gid];
        };
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return score(x) > score(y); });
    }
    slots.assign(order.size() * (size_t)K, -1);
    for (size_t pos = 0; pos < order.size(); ++pos) {
        const std::vector<int32_t> &mem = members[(size_t)order[pos]];
        for (size_t k = 0; k < mem.size(); ++k) {
            slots[pos * (size_t)K + k] = mem[k];
            slot_of[(size_t)idx_of[(size_t)mem[k]]] = (int64_t)(pos * (size_t)K + k);
        }
    }
}

// (re)build the arrays the kernels consume for groups of K trees
static int sim_regroup(s4_sim *sim, int K, int K_fill = 0, bool longest_first = false)
{
    if (K_fill <= 0 || K_fill > K) K_fill = K;
    sim->grouped_fill = K_fill;
    sim->grouped_fill0 = K_fill;
    sim->grouped_longest_first = longest_first;
    s4_graph *g = sim->g;
    s4_ctx *ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t T = sim->T, D = sim->D;
    sim->grouped_K = K;
    if (T == 0 || D == 0) {
        sim->Dp = 0;
        sim->h_root_off.assign(1, 0);
        return S4_OK;
    }
    std::vector<int32_t> slots;
    std::vector<int64_t> slot_of;
    cluster_roots(g, sim->h_roots0, K, K_fill, longest_first, slots, slot_of);
    const int64_t Dp = (int64_t)slots.size();
    sim->Dp = Dp;
    // trips of slot j start at h_root_off[j]; empty for padding slots
    std::vector<int64_t> count((size_t)Dp, 0), new_start((size_t)D);
    for (int64_t i = 0; i < D; ++i) count[(size_t)slot_of[(size_t)i]] = sim->h_root_off0[(size_t)i + 1] - sim->h_root_off0[(size_t)i];
    sim->h_root_off.assign((size_t)Dp + 1, 0);
    for (int64_t j = 0; j < Dp; ++j) sim->h_root_off[(size_t)j + 1] = sim->h_root_off[(size_t)j] + count[(size_t)j];
    std::vector<u32> slot_of32((size_t)D);
    for (int64_t i = 0; i < D; ++i) { new_start[(size_t)i] = sim->h_root_off[(size_t)slot_of[(size_t)i]]; slot_of32[(size_t)i] = (u32)slot_of[(size_t)i]; }
    DevBuf<int64_t> d_new_start;
    DevBuf<u32> d_slot_of;
    CU(d_new_start.alloc((size_t)D));
    CU(d_slot_of.alloc((size_t)D));
    CU(sim->roots.alloc((size_t)Dp));
    CU(sim->trip_root.ensure((size_t)T));
    CU(sim->trip_target.ensure((size_t)T));
    CU(cudaMemcpyAsync(d_new_start.p, new_start.data(), (size_t)D * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_slot_of.p, slot_of32.data(), (size_t)D * sizeof(u32), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sim->roots.p, slots.data(), (size_t)Dp * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    regroup_trips_kernel<<<grid_for(T, 256, ctx), 256, 0, st>>>(T, sim->trip_root0.p, sim->trip_target0.p, sim->root_off0.p, d_slot_of.p,
                                                               d_new_start.p, sim->trip_root.p, sim->trip_target.p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));                                  // the staging buffers go out of scope
    return S4_OK;
}

extern "C" int s4_sim_create(s4_graph *g, int64_t T, const int32_t *origin, const int32_t *goal, double jam_tolerance,
                             const s4_physics *phys, s4_sim **out)
{
    CHECK_ARG(g && out && phys, "s4_sim_create: NULL argument");
    *out = nullptr;
    CHECK_ARG(T >= 0 && T < (int64_t)1 << 31, "s4_sim_create: T out of range");
    CHECK_ARG(T == 0 || (origin && goal), "s4_sim_create: NULL trip array");
    CHECK_ARG(jam_tolerance == jam_tolerance, "s4_sim_create: jam_tolerance is NaN");
    s4_ctx *ctx = g->ctx;
    CU(cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < T; ++i)
        if (origin[i] < 0 || origin[i] >= g->V || goal[i] < 0 || goal[i] >= g->V)
            return fail(S4_ERR_ARG, "trip %lld: node id out of range", (long long)i);
    s4_sim *sim = new (std::nothrow) s4_sim();
    if (!sim) return fail(S4_ERR_OOM, "host allocation failed");
    struct Guard { s4_sim *s; ~Guard() { delete s; } } guard{sim};
    int mode = ctx->opt.root_mode;
    if (mode == S4_ROOT_AUTO)
        mode = (T > 0 && count_distinct_host(goal, T, g->V) < count_distinct_host(origin, T, g->V)) ? S4_ROOT_GOAL : S4_ROOT_ORIGIN;
    const int32_t *root_h = mode == S4_ROOT_GOAL ? goal : origin;
    const int32_t *target_h = mode == S4_ROOT_GOAL ? origin : goal;
    std::vector<int32_t> root_m, target_m;
    if (!g->rank.empty() && T > 0) {
        root_m.resize((size_t)T);
        target_m.resize((size_t)T);
        for (int64_t i = 0; i < T; ++i) { root_m[(size_t)i] = g->rank[(size_t)root_h[i]]; target_m[(size_t)i] = g->rank[(size_t)target_h[i]]; }
        root_h = root_m.data();
        target_h = target_m.data();
    }
    int rc = sim_alloc(g, sim, T, jam_tolerance, phys, mode);
    if (rc) return rc;
    DevBuf<u32> kin, vin;
    if (T > 0) {
        cudaStream_t st = ctx->stream;
        CU(kin.alloc((size_t)T));
        CU(vin.alloc((size_t)T));
        CU(cudaMemcpyAsync(kin.p, root_h, (size_t)T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(vin.p, target_h, (size_t)T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    if ((rc = sim_group(sim, kin, vin))) return rc;
    if ((rc = sim_regroup(sim, layout_for(g, sim->D).K))) return rc;
    guard.s = nullptr;
    ++g->refs;
    *out = sim;
    return S4_OK;
}

// The trips are drawn on the device (sample_trips_kernel): only the two populations cross PCIe.
extern "C" int s4_sim_create_sampled(s4_graph *g, int64_t T, int64_t n_origins, const int32_t *origins, int64_t n_goals,
                                     const int32_t *goals, uint64_t seed, double jam_tolerance, const s4_physics *phys,
                                     s4_sim **out)
{
    CHECK_ARG(g && out && phys, "s4_sim_create_sampled: NULL argument");
    *out = nullptr;
    CHECK_ARG(T >= 0 && T < (int64_t)1 << 31, "s4_sim_create_sampled: T out of range");
    CHECK_ARG(T == 0 || (origins && goals && n_origins > 0 && n_goals > 0),
              "s4_sim_create_sampled: empty population (random.sample raised ValueError there)");
    CHECK_ARG(jam_tolerance == jam_tolerance, "s4_sim_create_sampled: jam_tolerance is NaN");
    s4_ctx *ctx = g->ctx;
    CU(cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < n_origins; ++i)
        if (origins[i] < 0 || origins[i] >= g->V) return fail(S4_ERR_ARG, "origin %lld: node id out of range", (long long)i);
    for (int64_t i = 0; i < n_goals; ++i)
        if (goals[i] < 0 || goals[i] >= g->V) return fail(S4_ERR_ARG, "goal %lld: node id out of range", (long long)i);
    s4_sim *sim = new (std::nothrow) s4_sim();
    if (!sim) return fail(S4_ERR_OOM, "host allocation failed");
    struct Guard { s4_sim *s; ~Guard() { delete s; } } guard{sim};
    int mode = ctx->opt.root_mode;
    if (mode == S4_ROOT_AUTO)                       // the smaller population bounds the number of distinct endpoints
        mode = std::min<int64_t>(n_goals, T) < std::min<int64_t>(n_origins, T) ? S4_ROOT_GOAL : S4_ROOT_ORIGIN;
    int rc = sim_alloc(g, sim, T, jam_tolerance, phys, mode);
    if (rc) return rc;
    DevBuf<u32> kin, vin;
    if (T > 0) {
        cudaStream_t st = ctx->stream;
        const bool by_goal = mode == S4_ROOT_GOAL;
        const int32_t *pr = by_goal ? goals : origins, *pt = by_goal ? origins : goals;
        const int64_t nr = by_goal ? n_goals : n_origins, nt = by_goal ? n_origins : n_goals;
        std::vector<int32_t> mr(pr, pr + nr), mt(pt, pt + nt);          // populations in the device's numbering
        if (!g->rank.empty()) {
            for (auto &x : mr) x = g->rank[(size_t)x];
            for (auto &x : mt) x = g->rank[(size_t)x];
        }
        DevBuf<int32_t> d_r, d_t;
        CU(d_r.alloc((size_t)nr));
        CU(d_t.alloc((size_t)nt));
        CU(kin.alloc((size_t)T));
        CU(vin.alloc((size_t)T));
        CU(cudaMemcpyAsync(d_r.p, mr.data(), (size_t)nr * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_t.p, mt.data(), (size_t)nt * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        sample_trips_kernel<<<grid_for(T, 256, ctx), 256, 0, st>>>(T, (u64)seed, d_r.p, (u64)nr, by_goal ? 1 : 0, d_t.p, (u64)nt,
                                                                   kin.p, vin.p);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));                                   // mr / mt / d_r / d_t go out of scope
    }
    if ((rc = sim_group(sim, kin, vin))) return rc;
    if ((rc = sim_regroup(sim, layout_for(g, sim->D).K))) return rc;
    guard.s = nullptr;
    ++g->refs;
    *out = sim;
    return S4_OK;
}

// The simulation's trips in the caller's numbering, grouped by tree root (the order the device keeps them in).
extern "C" int s4_sim_get_trips(s4_sim *sim, int32_t *origin_out, int32_t *goal_out)
{
    CHECK_ARG(sim && origin_out && goal_out, "s4_sim_get_trips: NULL argument");
    s4_graph *g = sim->g;
    s4_ctx *ctx = g->ctx;
    CU(cudaSetDevice(ctx->device));
    const int64_t T = sim->T, D = sim->Dp;
    if (T == 0) return S4_OK;
    std::vector<u32> troot((size_t)T);
    std::vector<int32_t> ttarget((size_t)T), roots((size_t)D);
    CU(cudaMemcpyAsync(troot.data(), sim->trip_root.p, (size_t)T * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ttarget.data(), sim->trip_target.p, (size_t)T * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(roots.data(), sim->roots.p, (size_t)D * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> inv;                                           // device number -> caller's id
    if (!g->rank.empty()) {
        inv.resize((size_t)g->V);
        for (int32_t i = 0; i < g->V; ++i) inv[(size_t)g->rank[(size_t)i]] = i;
    }
    int32_t *root_dst = sim->root_mode == S4_ROOT_GOAL ? goal_out : origin_out;
    int32_t *target_dst = sim->root_mode == S4_ROOT_GOAL ? origin_out : goal_out;
    for (int64_t t = 0; t < T; ++t) {
        const int32_t r = roots[(size_t)troot[(size_t)t]], x = ttarget[(size_t)t];
        root_dst[t] = inv.empty() ? r : inv[(size_t)r];
        target_dst[t] = inv.empty() ? x : inv[(size_t)x];
    }
    return S4_OK;
}

// Diagnostics: the root slots of the day as the SSSP kernel consumes them (groups of K slots, -1 = empty), caller's ids.
extern "C" int s4_sim_get_root_slots(s4_sim *sim, int32_t *out, int64_t cap, int64_t *n_slots, int *group_size)
{
    CHECK_ARG(sim && n_slots, "s4_sim_get_root_slots: NULL argument");
    s4_graph *g = sim->g;
    CU(cudaSetDevice(g->ctx->device));
    *n_slots = sim->Dp;
    if (group_size) *group_size = sim->grouped_K;
    if (!out || sim->Dp == 0) return S4_OK;
    CHECK_ARG(cap >= sim->Dp, "s4_sim_get_root_slots: buffer too small");
    CU(cudaMemcpyAsync(out, sim->roots.p, (size_t)sim->Dp * sizeof(int32_t), cudaMemcpyDeviceToHost, g->ctx->stream));
    CU(cudaStreamSynchronize(g->ctx->stream));
    if (!g->rank.empty()) {
        std::vector<int32_t> inv((size_t)g->V);
        for (int32_t i = 0; i < g->V; ++i) inv[(size_t)g->rank[(size_t)i]] = i;
        for (int64_t i = 0; i < sim->Dp; ++i)
            if (out[i] >= 0) out[i] = inv[(size_t)out[i]];
    }
    return S4_OK;
}

static void free_events(std::vector<EvPair> &v)
{
    for (auto &e : v) {
        if (e.a) cudaEventDestroy(e.a);
        if (e.b) cudaEventDestroy(e.b);
    }
    v.clear();
}

extern "C" int s4_sim_free(s4_sim *sim)
{
    if (!sim) return S4_OK;
    cudaSetDevice(sim->g->ctx->device);
    cudaStreamSynchronize(sim->g->ctx->stream);
    free_events(sim->ev_w);
    free_events(sim->ev_sssp);
    free_events(sim->ev_walk);
    free_events(sim->ev_adapt);
    free_events(sim->ev_step);
    if (sim->hops_host) cudaFreeHost(sim->hops_host);
    s4_graph *g = sim->g;
    delete sim;
    graph_unref(g);
    return S4_OK;
}

extern "C" int s4_sim_roots(s4_sim *sim, int64_t *n_roots, int *root_mode)
{
    CHECK_ARG(sim, "s4_sim_roots: sim is NULL");
    if (n_roots) *n_roots = sim->D;
    if (root_mode) *root_mode = sim->root_mode;
    return S4_OK;
}

// ---- event bookkeeping: pairs are recorded around kernels and harvested lazily
static int ev_begin(std::vector<EvPair> &v, size_t &used, cudaStream_t st)
{
    if (used == v.size()) {
        EvPair p;
        CU(cudaEventCreate(&p.a));
        CU(cudaEventCreate(&p.b));
        v.push_back(p);
    }
    CU(cudaEventRecord(v[used].a, st));
    return S4_OK;
}
static int ev_end(std::vector<EvPair> &v, size_t &used, cudaStream_t st)
{
    CU(cudaEventRecord(v[used].b, st));
    ++used;
    return S4_OK;
}
// wait = false: only the pairs that have completed (a prefix; the rest moves to the front and is harvested later), so
// that s4_sim_step never waits for the previous day: the host queues day N+1 while the GPU still runs day N
static int ev_harvest(std::vector<EvPair> &v, size_t &used, double *acc, bool wait)
{
    size_t done = 0;
    for (; done < used; ++done) {
        if (wait) CU(cudaEventSynchronize(v[done].b));
        else {
            cudaError_t q = cudaEventQuery(v[done].b);
            if (q == cudaErrorNotReady) break;
            CU(q);
        }
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, v[done].a, v[done].b));
        *acc += (double)ms;
    }
    std::rotate(v.begin(), v.begin() + (ptrdiff_t)done, v.begin() + (ptrdiff_t)used);
    used -= done;
    return S4_OK;
}
static int harvest_all(s4_sim *sim, bool wait = true)
{
    int rc;
    if ((rc = ev_harvest(sim->ev_w, sim->used_w, &sim->acc.ms_weights, wait))) return rc;
    if ((rc = ev_harvest(sim->ev_sssp, sim->used_sssp, &sim->acc.ms_sssp, wait))) return rc;
    if ((rc = ev_harvest(sim->ev_walk, sim->used_walk, &sim->acc.ms_walk, wait))) return rc;
    if ((rc = ev_harvest(sim->ev_adapt, sim->used_adapt, &sim->acc.ms_adapt, wait))) return rc;
    if ((rc = ev_harvest(sim->ev_step, sim->used_step, &sim->acc.ms_step_total, wait))) return rc;
    return S4_OK;
}

// K1 on the stream: w_street, arc records, sum(w)
static int run_weights(s4_sim *sim, int reset_load)
{
    s4_graph *g = sim->g;
    s4_ctx *ctx = g->ctx;
    cudaStream_t st = ctx->stream;
    if (!sim->w_hist.p) {
        CU(sim->w_hist.alloc(kWeightBins));
        CU(cudaMemsetAsync(sim->w_hist.p, 0, kWeightBins * sizeof(u32), st));     // weight_scale_kernel clears it after use
    }
    if (g->S > 0) {
        weights_street_kernel<<<grid_for(g->S, 256, ctx), 256, 0, st>>>(g->S, g->length.p, sim->ms_vis.p, sim->load.p, sim->jam,
                                                                        sim->ph, reset_load, sim->w_street.p, sim->w_hist.p);
        CU(cudaGetLastError());
        weight_scale_kernel<<<1, 1024, 0, st>>>(sim->w_hist.p, g->S, sim->wsum.p);   // bucket-width scale: median weight * S
        CU(cudaGetLastError());
        sim->acc.kernel_launches += 2;
    } else {
        CU(cudaMemsetAsync(sim->wsum.p, 0, sizeof(double), st));
    }
    if (g->A > 0) {
        weights_arc_kernel<<<grid_for(g->A, 256, ctx), 256, 0, st>>>(g->A, g->arc_cs.p, sim->w_street.p, sim->arcs.p);
        CU(cudaGetLastError());
        ++sim->acc.kernel_launches;
    }
    // fixed-stride rows: the source-batched SSSP and the walk share one table {head, street_index, w}; the
    // one-tree-per-CTA kernel reads its own ({head, entry word, w})
    const bool grouped = ctx->opt.group_lanes > 0;
    const int64_t n_slots = (int64_t)g->V * kEllWidth;
    if (!grouped) {
        CU(sim->ell.ensure((size_t)n_slots));
        weights_ell_kernel<<<grid_for(n_slots, 256, ctx), 256, 0, st>>>(n_slots, g->ell_cs.p, sim->w_street.p, sim->ell.p);
        CU(cudaGetLastError());
        ++sim->acc.kernel_launches;
    }
    if (grouped || ctx->opt.walk_ell) {
        CU(sim->ellw.ensure((size_t)n_slots));
        weights_ellw_kernel<<<grid_for(n_slots, 256, ctx), 256, 0, st>>>(n_slots, g->ell_cs.p, sim->w_street.p, sim->ellw.p);
        CU(cudaGetLastError());
        ++sim->acc.kernel_launches;
    }
    sim->weights_valid = true;
    return S4_OK;
}

extern "C" int s4_sim_weights(s4_sim *sim, int refresh, double *w_out)
{
    CHECK_ARG(sim, "s4_sim_weights: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (!refresh && !sim->weights_valid) return fail(S4_ERR_STATE, "s4_sim_weights: no weights computed yet");
    if (refresh) {
        int rc = run_weights(sim, 0);
        if (rc) return rc;
    }
    if (w_out && sim->g->S > 0)
        CU(cudaMemcpyAsync(w_out, sim->w_street.p, (size_t)sim->g->S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return S4_OK;
}

extern "C" int s4_sim_step(s4_sim *sim)
{
    CHECK_ARG(sim, "s4_sim_step: sim is NULL");
    s4_graph *g = sim->g;
    s4_ctx *ctx = g->ctx;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc = harvest_all(sim, false);   // the timing events of earlier days that have completed; never waits
    if (rc) return rc;
    SsspLaunch L;
    if ((rc = plan_sssp(g, &L, sim->D))) return rc;
    const int64_t K = L.lay.K;
    // Trees per group.  A launch runs L.units() groups at a time; when the day has only a few waves of groups (several
    // GPUs sharing a day), a partly filled last wave costs a whole group time: then the groups take fewer trees, so
    // that their number fills whole waves (4 % head room: the greedy clustering leaves some groups short).
    int K_fill = (int)K;
    int64_t waves_wanted = 0;                                      // 0: many waves, their number does not matter
    // groups that run at the same time: one per SM (or cluster), and no more than the pool has slots for
    int64_t W_run = std::max(1, L.units());
    {
        double pool_budget = 0.0, ws_budget = 0.0;
        if ((rc = scratch_budget(g, &pool_budget, &ws_budget))) return rc;
        const int64_t by_mem = (int64_t)(pool_budget / (8.0 * (double)slot_words(g, L.lay)));
        const int64_t cap = ctx->opt.batch_roots > 0 ? std::min<int64_t>(std::max<int64_t>(1, ctx->opt.batch_roots / K), by_mem) : by_mem;
        W_run = std::max<int64_t>(1, std::min<int64_t>(W_run, cap));
    }
    if (K > 1 && sim->D > 0 && ctx->opt.balance_waves) {
        const int64_t W = W_run;
        const int64_t G0 = (sim->D + K - 1) / K, waves = (G0 + W - 1) / W;
        const double target = 0.96 * (double)(waves * W);
        K_fill = (int)std::min<int64_t>(K, std::max<int64_t>(1, (int64_t)std::ceil((double)sim->D / std::max(1.0, target))));
        if (waves <= 4) waves_wanted = waves;
    }
    // ... and with at most four waves of groups in the whole day they are started longest-first (expected depth of the
    // trees + spread of the group), all in one launch per lane-slot, so that the work queue packs the waves
    const bool few_waves = waves_wanted > 0;
    if (sim->grouped_K != (int)K || sim->grouped_fill0 != K_fill || sim->grouped_longest_first != few_waves) {
        // the clustering leaves some groups short, so their number is only known afterwards: groups beyond the
        // intended waves would cost a whole group time -- then every group takes one tree more, and again
        int fill = K_fill;
        for (int attempt = 0;; ++attempt) {
            if ((rc = sim_regroup(sim, (int)K, fill, few_waves))) return rc;
            if (!few_waves || sim->Dp / K <= waves_wanted * W_run || fill >= (int)K || attempt >= 12) break;
            ++fill;
        }
        sim->grouped_fill0 = K_fill;
    }
    const int64_t D = sim->Dp;                                     // root slots (groups padded to K)
    const int64_t G = D / K;                                       // slots of the distance pool the day needs
    if (D > 0 && (rc = ensure_scratch(g, L, G + (G & 1)))) return rc;   // an even number of slots: two lanes of ceil(G / 2)

    if ((rc = ev_begin(sim->ev_step, sim->used_step, st))) return rc;
    // ---- K1 (simulation.py:53-65) + reset of today's counters (:68)
    if ((rc = ev_begin(sim->ev_w, sim->used_w, st))) return rc;
    if ((rc = run_weights(sim, 1))) return rc;
    if ((rc = ev_end(sim->ev_w, sim->used_w, st))) return rc;

    // ---- K2 + K3 in batches (simulation.py:71-88).  The work runs in two independent lanes (own stream, own
    // half of the distance pool, own frontier workspace): batch i goes to lane i & 1, where its SSSP launch is
    // followed by its walk.  The persistent SSSP grid of one lane fills the SMs the other lane's launch frees
    // while it drains its last trees, and the walk (a latency-bound pointer chase) hides under the other lane's
    // SSSP.  A day is cut into at least `min_batches` batches even when all its trees would fit the pool at once.
    const int64_t slots = std::max<int64_t>(1, g->pool_slots);
    // a day of a single wave of groups that all fit in the pool is one launch (nothing for a second lane to overlap with)
    const bool split = ctx->opt.lanes >= 2 && slots >= 2 && G >= 2 && !(waves_wanted == 1 && slots >= G);
    int64_t B = split ? slots / 2 : slots;                         // slots per launch
    // a launch should hold at least two waves of groups (the work queue then packs them); more launches only when
    // the day is long enough for the walk of one batch to hide under the SSSP of the next
    if (split) B = std::min(B, std::max<int64_t>(std::max<int64_t>(1, (G + ctx->opt.min_batches - 1) / ctx->opt.min_batches),
                                                  ctx->opt.balance_waves ? 2 * (int64_t)std::max(1, L.units()) : 1));
    // whole waves per launch where a launch holds several (a partly filled last wave costs a whole group time)
    if (ctx->opt.balance_waves && B > (int64_t)L.units()) B -= B % (int64_t)L.units();
    const int64_t n_batches = G > 0 ? (G + B - 1) / B : 0;
    g->glog_groups = 0;
    if (ctx->opt.group_log && L.gfn && G > 0) {
        CU(g->glog.ensure((size_t)G * 4));
        CU(cudaMemsetAsync(g->glog.p, 0, (size_t)G * 4 * sizeof(u32), st));
        g->glog_groups = G;
    }
    if (n_batches > 0) {
        CU(g->work_counters.ensure((size_t)std::max<int64_t>(n_batches, 1024)));
        CU(cudaMemsetAsync(g->work_counters.p, 0, g->work_counters.n * sizeof(u32), st));
        CU(sim->deferred.ensure((size_t)std::max<int64_t>(sim->T, 1)));
        CU(sim->deferred_counts.ensure((size_t)std::max<int64_t>(n_batches, 1024)));
        CU(cudaMemsetAsync(sim->deferred_counts.p, 0, sim->deferred_counts.n * sizeof(u32), st));
    }
    // optional: the arc rows stay in L2 (persisting access-policy window; the distance rows stream through the rest)
    {
        cudaStreamAttrValue av = {};
        if (ctx->opt.l2_persist && sim->ellw.p) {
            const size_t bytes = (size_t)g->V * kEllWidth * sizeof(Arc);
            const size_t max_persist = (size_t)ctx->prop.persistingL2CacheMaxSize, max_win = (size_t)ctx->prop.accessPolicyMaxWindowSize;
            if (max_persist > 0 && max_win > 0) {
                if (!ctx->l2_limit_set) { CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, max_persist)); ctx->l2_limit_set = true; }
                av.accessPolicyWindow.base_ptr = (void *)sim->ellw.p;
                av.accessPolicyWindow.num_bytes = std::min(bytes, max_win);
                av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)std::min(bytes, max_win));
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            }
        }
        if (ctx->opt.l2_persist || ctx->l2_window_on) {
            CU(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
            CU(cudaStreamSetAttribute(ctx->walk_stream, cudaStreamAttributeAccessPolicyWindow, &av));
            ctx->l2_window_on = ctx->opt.l2_persist != 0;
        }
    }
    if (split) {                                                   // lane 1 starts after K1 and the counter reset
        CU(cudaEventRecord(ctx->sssp_done[0], st));
        CU(cudaStreamWaitEvent(ctx->walk_stream, ctx->sssp_done[0], 0));
    }
    // predecessor codes pay when the walks are long against the streaming pass over V x trees words: measured on cfg3 the
    // pass costs what 0.1 * V * D hops cost the eight-lane walk.  Hops of the last day whose counter has arrived;
    // before that, trips x half the mean hop depth of a tree (720 hops per trip on cfg3: the estimate says 770).
    if (!sim->hops_host) {
        CU(cudaHostAlloc((void **)&sim->hops_host, sizeof(u64), cudaHostAllocDefault));
        *sim->hops_host = 0;
    }
    {
        const u64 now = *(volatile u64 *)sim->hops_host;
        if (now > sim->hops_seen) { sim->hops_per_day = (double)(now - sim->hops_seen); sim->hops_seen = now; }
        else if (now < sim->hops_seen) { sim->hops_seen = now; }                    // counters were reset
    }
    const double hops_expected = sim->hops_per_day > 0.0 ? sim->hops_per_day : 0.5 * g->mean_ecc * (double)sim->T;
    const bool codes_pay = hops_expected >= 0.1 * (double)g->V * (double)std::max<int64_t>(1, sim->D);
    const size_t sw = slot_words(g, L.lay);
    for (int64_t bi = 0; bi < n_batches; ++bi) {
        const int64_t r0 = bi * B * K, r1 = std::min(D, r0 + B * K);
        const int lane = split ? (int)(bi & 1) : 0;
        cudaStream_t ls = lane ? ctx->walk_stream : st;
        const int64_t first_slot = lane * (slots / 2);
        if ((rc = ev_begin(sim->ev_sssp, sim->used_sssp, ls))) return rc;
        if ((rc = launch_sssp(g, L, L.gfn ? sim->ellw.p : sim->ell.p, sim->arcs.p, sim->roots.p + r0, (int)(r1 - r0),
                              g->work_counters.p + bi, sim->wsum.p, sim->counters.p, first_slot, lane, ls, bi * B))) return rc;
        if ((rc = ev_end(sim->ev_sssp, sim->used_sssp, ls))) return rc;
        ++sim->acc.kernel_launches;
        ++sim->acc.sssp_launches;
        WalkParams wp;
        wp.row_ptr = g->row_ptr.p;
        wp.arcs = sim->arcs.p;
        wp.ellw = ctx->opt.walk_ell ? sim->ellw.p : nullptr;
        wp.ell_cs = g->ell_cs.p;
        wp.orig = g->orig.p;
        wp.slot_stride = sw;
        wp.group = (u32)K;
        wp.node_stride = (u32)L.lay.RW;
        wp.dist_pool = g->pool.p + (size_t)first_slot * sw;
        wp.roots = sim->roots.p;
        wp.trip_target = sim->trip_target.p;
        wp.trip_root = sim->trip_root.p;
        wp.t0 = sim->h_root_off[(size_t)r0];
        wp.t1 = sim->h_root_off[(size_t)r1];
        wp.root0 = (u32)r0;
        wp.V = (u32)g->V;
        wp.volume = sim->volume;
        wp.load = sim->load.p;
        wp.counters = sim->counters.p;
        wp.t_base = wp.t0;
        wp.deferred = sim->deferred.p + wp.t0;
        wp.deferred_count = sim->deferred_counts.p + bi;
        if ((rc = ev_begin(sim->ev_walk, sim->used_walk, ls))) return rc;
        // trip-heavy days: every (node, tree) predecessor is resolved once, the walks then take one lane and one
        // dependent access per hop
        const bool codes = L.gfn && ctx->opt.group_lanes == 8 && wp.ellw && (ctx->opt.walk_codes == 1 || (ctx->opt.walk_codes == 2 && codes_pay));
        if (codes) rc = launch_walk_codes(g, wp, (int)((r1 - r0 + K - 1) / K), L.lay.RW / 8, ls);
        else rc = launch_walk(g, wp, ls);
        if (rc) return rc;
        if (codes) ++sim->acc.kernel_launches;
        if ((rc = ev_end(sim->ev_walk, sim->used_walk, ls))) return rc;
        sim->acc.kernel_launches += 2;
    }
    if (split) {                                                   // the day ends when both lanes are done
        CU(cudaEventRecord(ctx->walk_done[1], ctx->walk_stream));
        CU(cudaStreamWaitEvent(st, ctx->walk_done[1], 0));
    }
    CU(cudaMemcpyAsync(sim->hops_host, sim->counters.p + C_HOPS, sizeof(u64), cudaMemcpyDeviceToHost, st));   // for tomorrow's choice
    if ((rc = ev_end(sim->ev_step, sim->used_step, st))) return rc;
    sim->acc.days += 1;
    sim->acc.sssp_instances += (uint64_t)sim->D;
    sim->acc.arcs_algorithmic += (uint64_t)sim->alg_arcs_per_day;
    sim->acc.nodes_algorithmic += (uint64_t)sim->alg_nodes_per_day;
    return S4_OK;
}

extern "C" int s4_sim_get_load(s4_sim *sim, uint32_t *out)
{
    CHECK_ARG(sim && out, "s4_sim_get_load: NULL argument");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (sim->g->S > 0) CU(cudaMemcpyAsync(out, sim->load.p, (size_t)sim->g->S * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return S4_OK;
}

extern "C" int s4_sim_set_load(s4_sim *sim, const uint32_t *in)
{
    CHECK_ARG(sim && in, "s4_sim_set_load: NULL argument");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (sim->g->S > 0) CU(cudaMemcpyAsync(sim->load.p, in, (size_t)sim->g->S * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));     // the host buffer is only borrowed
    return S4_OK;
}

extern "C" void *s4_sim_load_device_ptr(s4_sim *sim) { return sim ? (void *)sim->load.p : nullptr; }

extern "C" int s4_sim_load_add(s4_sim *dst, s4_sim *src)
{
    CHECK_ARG(dst && src && dst->g->ctx == src->g->ctx && dst->g->S == src->g->S, "s4_sim_load_add: simulations must live on one device and have equally many streets");
    s4_ctx *ctx = dst->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (dst->g->S > 0 && dst != src) {
        add_u32_kernel<<<grid_for(dst->g->S, 256, ctx), 256, 0, ctx->stream>>>(dst->g->S, dst->load.p, src->load.p);
        CU(cudaGetLastError());
        ++dst->acc.kernel_launches;
    }
    return S4_OK;
}

extern "C" int s4_sim_load_copy(s4_sim *dst, s4_sim *src)
{
    CHECK_ARG(dst && src && dst->g->ctx == src->g->ctx && dst->g->S == src->g->S, "s4_sim_load_copy: simulations must live on one device and have equally many streets");
    s4_ctx *ctx = dst->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (dst->g->S > 0 && dst != src)
        CU(cudaMemcpyAsync(dst->load.p, src->load.p, (size_t)dst->g->S * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
    return S4_OK;
}

extern "C" int s4_sim_commit_total(s4_sim *sim)
{
    CHECK_ARG(sim, "s4_sim_commit_total: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    const int64_t S = sim->g->S;
    if (S > 0) {
        if (!sim->has_cum) {
            // merge_arrays((total, None)) == total   (utils.py:30)
            CU(cudaMemcpyAsync(sim->cum.p, sim->load.p, (size_t)S * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            add_u32_kernel<<<grid_for(S, 256, ctx), 256, 0, ctx->stream>>>(S, sim->cum.p, sim->load.p);
            CU(cudaGetLastError());
            ++sim->acc.kernel_launches;
        }
    }
    sim->has_cum = true;
    return S4_OK;
}

// streets4mpi.py:97-100 in one call: total = Allreduce(local, SUM) in place on the device-resident counters
// (stream-ordered behind the day's kernels, no host synchronisation), then cumulative = total + (cumulative or 0).
extern "C" int s4_sim_exchange(s4_sim *sim)
{
    CHECK_ARG(sim, "s4_sim_exchange: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (ctx->comm_ranks > 1) {
        if (!ctx->nccl_comm) return fail(S4_ERR_STATE, "s4_sim_exchange: %d ranks but no communicator (s4_comm_init)", ctx->comm_ranks);
        if (sim->g->S > 0)
            NC(g_nccl.AllReduce(sim->load.p, sim->load.p, (size_t)sim->g->S, /*ncclUint32*/ 3, /*ncclSum*/ 0, ctx->nccl_comm, ctx->stream));
    }
    return s4_sim_commit_total(sim);
}

// 64-bit FNV-1a over the little-endian bytes of traffic_load / cumulative_traffic_load: lets a driver compare whole
// load vectors across runs and GPU counts without moving them (tests, bench lines).
static uint64_t fnv1a(const uint32_t *v, int64_t n)
{
    uint64_t h = 1469598103934665603ull;
    const unsigned char *b = reinterpret_cast<const unsigned char *>(v);
    for (int64_t i = 0; i < n * 4; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

extern "C" int s4_sim_load_hash(s4_sim *sim, uint64_t *load_hash, uint64_t *cumulative_hash)
{
    CHECK_ARG(sim, "s4_sim_load_hash: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    const int64_t S = sim->g->S;
    std::vector<uint32_t> h((size_t)std::max<int64_t>(S, 1));
    if (load_hash) {
        if (S > 0) CU(cudaMemcpyAsync(h.data(), sim->load.p, (size_t)S * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        *load_hash = fnv1a(h.data(), S);
    }
    if (cumulative_hash) {
        *cumulative_hash = 0;
        if (sim->has_cum) {
            if (S > 0) CU(cudaMemcpyAsync(h.data(), sim->cum.p, (size_t)S * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            *cumulative_hash = fnv1a(h.data(), S);
        }
    }
    return S4_OK;
}

extern "C" int s4_sim_get_cumulative(s4_sim *sim, uint32_t *out, int *is_none)
{
    CHECK_ARG(sim, "s4_sim_get_cumulative: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (is_none) *is_none = sim->has_cum ? 0 : 1;
    if (out && sim->has_cum && sim->g->S > 0)
        CU(cudaMemcpyAsync(out, sim->cum.p, (size_t)sim->g->S * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return S4_OK;
}

extern "C" int s4_sim_set_cumulative(s4_sim *sim, const uint32_t *in_or_null)
{
    CHECK_ARG(sim, "s4_sim_set_cumulative: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    if (!in_or_null) { sim->has_cum = false; return S4_OK; }
    if (sim->g->S > 0) CU(cudaMemcpyAsync(sim->cum.p, in_or_null, (size_t)sim->g->S * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    sim->has_cum = true;
    return S4_OK;
}

// simulation.py:97-108 control flow, executed literally on the host for n streets
static void construction_counts(int64_t n, int64_t *n_dec, int64_t *n_inc)
{
    double dec_limit = 0.15 * (double)n, inc_limit = 0.95 * (double)n;
    int64_t nd = 0, ni = 0;
    for (int64_t i = 0; i < n; ++i) {
        if ((double)i <= dec_limit) { ++nd; dec_limit += 1.0; }      // :100-102 (`if not None` is always true)
        const int64_t j = n - i - 1;                                 // :103
        if ((double)j >= inc_limit) { ++ni; inc_limit -= 1.0; }      // :104-106
        if (dec_limit >= inc_limit) break;                           // :107
    }
    *n_dec = nd;
    *n_inc = ni;
}

extern "C" int s4_sim_road_construction(s4_sim *sim, const uint8_t *adaptable)
{
    CHECK_ARG(sim, "s4_sim_road_construction: sim is NULL");
    if (!sim->has_cum) return fail(S4_ERR_STATE, "road_construction: cumulative_traffic_load is None (simulation.py:94 would raise)");
    s4_graph *g = sim->g;
    s4_ctx *ctx = g->ctx;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t S = g->S;
    int rc;
    if (S > 0) {
        CU(sim->sort_vals.ensure((size_t)S));
        CU(sim->sort_keys_out.ensure((size_t)S));
        CU(sim->order.ensure((size_t)S));
        if ((rc = ev_begin(sim->ev_adapt, sim->used_adapt, st))) return rc;
        iota_kernel<<<grid_for(S, 256, ctx), 256, 0, st>>>(S, sim->sort_vals.p);
        CU(cudaGetLastError());
        // stable LSD radix sort by load; values start in street_index order => ties keep it (:96)
        if ((rc = cub_sort_pairs(ctx, sim->cub_tmp, sim->cum.p, sim->sort_keys_out.p, sim->sort_vals.p, sim->order.p, S))) return rc;
        int64_t n_dec = 0, n_inc = 0;
        construction_counts(S, &n_dec, &n_inc);
        const uint8_t *d_mask = nullptr;
        if (adaptable) {
            CU(sim->mask.ensure((size_t)S));
            CU(cudaMemcpyAsync(sim->mask.p, adaptable, (size_t)S, cudaMemcpyHostToDevice, st));
            d_mask = sim->mask.p;
        }
        adapt_kernel<<<grid_for(S, 256, ctx), 256, 0, st>>>(S, sim->order.p, n_dec, n_inc, d_mask, sim->ms_vis.p, sim->ms_ins.p);
        CU(cudaGetLastError());
        if ((rc = ev_end(sim->ev_adapt, sim->used_adapt, st))) return rc;
        sim->acc.kernel_launches += 3;
        if (adaptable) CU(cudaStreamSynchronize(st));   // borrowed host mask
    }
    sim->has_cum = false;                               // simulation.py:109
    return S4_OK;
}

extern "C" int s4_sim_get_max_speed(s4_sim *sim, int32_t *visible_out, int32_t *insertion_out)
{
    CHECK_ARG(sim, "s4_sim_get_max_speed: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)sim->g->S * sizeof(int32_t);
    if (visible_out && bytes) CU(cudaMemcpyAsync(visible_out, sim->ms_vis.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (insertion_out && bytes) CU(cudaMemcpyAsync(insertion_out, sim->ms_ins.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return S4_OK;
}

extern "C" int s4_sim_set_max_speed(s4_sim *sim, const int32_t *visible, const int32_t *insertion_or_null)
{
    CHECK_ARG(sim && visible, "s4_sim_set_max_speed: NULL argument");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    const int64_t S = sim->g->S;
    for (int64_t s = 0; s < S; ++s)
        if (visible[s] < 1) return fail(S4_ERR_ARG, "street %lld: max_speed must be >= 1", (long long)s);
    const size_t bytes = (size_t)S * sizeof(int32_t);
    if (bytes) {
        CU(cudaMemcpyAsync(sim->ms_vis.p, visible, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(sim->ms_ins.p, insertion_or_null ? insertion_or_null : visible, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return S4_OK;
}

extern "C" int s4_sim_get_counters(s4_sim *sim, s4_counters *out)
{
    CHECK_ARG(sim && out, "s4_sim_get_counters: NULL argument");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    int rc = harvest_all(sim);
    if (rc) return rc;
    u64 h[C_COUNT];
    CU(cudaMemcpy(h, sim->counters.p, sizeof h, cudaMemcpyDeviceToHost));
    *out = sim->acc;
    out->arcs_relaxed = h[C_ARCS_RELAXED];
    out->nodes_popped = h[C_NODES_POPPED];
    out->stale_pops = h[C_STALE_POPS];
    out->queue_pushes = h[C_PUSHES];
    out->bucket_advances = h[C_ADVANCES];
    out->relax_rounds = h[C_ROUNDS];
    out->fallback_sweeps = h[C_FALLBACK_SWEEPS];
    out->trips_routed = h[C_TRIPS];
    out->trips_unreachable = h[C_UNREACHABLE];
    out->trips_stuck = h[C_STUCK];
    out->hops = h[C_HOPS];
    return S4_OK;
}

extern "C" int s4_sim_trips_stuck(s4_sim *sim, uint64_t *out)
{
    CHECK_ARG(sim && out, "s4_sim_trips_stuck: NULL argument");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    u64 v = 0;
    CU(cudaMemcpyAsync(&v, sim->counters.p + C_STUCK, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out = v;
    return S4_OK;
}

extern "C" int s4_sim_reset_counters(s4_sim *sim)
{
    CHECK_ARG(sim, "s4_sim_reset_counters: sim is NULL");
    s4_ctx *ctx = sim->g->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    int rc = harvest_all(sim);
    if (rc) return rc;
    CU(cudaMemset(sim->counters.p, 0, C_COUNT * sizeof(u64)));
    sim->acc = s4_counters{};
    return S4_OK;
}

extern "C" int s4_driving_speed(s4_ctx *ctx, int64_t n, const double *length, const int32_t *max_speed,
                                const uint32_t *trips, const s4_physics *phys, double *out)
{
    CHECK_ARG(ctx && phys && n >= 0 && (n == 0 || (length && max_speed && trips && out)), "s4_driving_speed: bad argument");
    if (n == 0) return S4_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf<double> d_len, d_out;
    DevBuf<int32_t> d_ms;
    DevBuf<u32> d_tr;
    CU(d_len.alloc((size_t)n)); CU(d_out.alloc((size_t)n)); CU(d_ms.alloc((size_t)n)); CU(d_tr.alloc((size_t)n));
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(d_len.p, length, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ms.p, max_speed, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_tr.p, trips, (size_t)n * sizeof(u32), cudaMemcpyHostToDevice, st));
    Phys ph{phys->car_length, phys->min_breaking_distance, phys->braking_deceleration};
    driving_speed_kernel<<<grid_for(n, 256, ctx), 256, 0, st>>>(n, d_len.p, d_ms.p, d_tr.p, ph, d_out.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return S4_OK;
}
