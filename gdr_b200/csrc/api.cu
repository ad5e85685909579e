// C-ABI entry points (include/gdr_b200.h): handle management, argument checking, scratch sizing,
// kernel sequencing.  No torch types, no exceptions across the boundary.
#include <climits>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "gdr_common.cuh"

namespace gdr {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return GDR_ERR_CUDA;
}

static int invalid(const char *msg) {
    g_last_error = msg;
    return GDR_ERR_INVALID;
}


}  // namespace gdr

using namespace gdr;

struct gdr_store {
    const void *emb = nullptr;
    int64_t n_docs = 0;
    int32_t dim = 0, dtype = 0, n_clusters = 0, max_cluster = 0;
    const int32_t *offsets = nullptr;
    const int32_t *docid = nullptr;
    int device = 0, sm_count = 148;
    // per-cluster scratch (allocated at create)
    int32_t *cluster_ws = nullptr;   // cnt | grp_off | simt_off | umma_off | counters
    // per-batch scratch (grows)
    void *batch_ws = nullptr;
    size_t batch_ws_bytes = 0;
    bool has_tmap = false;
    CUtensorMap tmap;
    int last_launches = 0;
    int phase_launches[3] = {0, 0, 0};             // inversion, scoring, top-k kernels of the handle's current batch (phases may come in separate calls)
    int umma_min_group = 1;   // > 1 (env GDR_UMMA_MIN_GROUP) = mixed mode
    int umma_ctas_per_sm = 1; // GDR_OPT_UMMA_CTAS_PER_SM: 2 = the 4-stage kernel, two scoring CTAs per SM (score_umma_x2.cu)
    int umma_ctas = 0;        // > 0 (env GDR_UMMA_CTAS): persistent CTAs of the tcgen05 kernel (default: one per SM)
    uint32_t debug_flags = 0; // GDR_OPT_TOPK_GROUPS bits; with -DGDR_DEBUG_KNOBS also the GDR_UMMA_DEBUG / GDR_TOPK_DEBUG timing experiments
    bool topk_wide = false;   // GDR_OPT_TOPK_WIDE: the 256-thread top-k also for k <= 128
    int prio_invert = 0, prio_score = 0, prio_topk = 0;   // GDR_OPT_LAUNCH_PRIORITIES: per-launch priorities (+1000), 0 = off
    bool profiling = false;
    // what the last gdr_score_topk call on this handle set up (scratch pointers, shapes): input of gdr_score_fused (experiment)
    ScoreArgs last_args;
    bool last_valid = false, last_umma_only = false;
    int fused_groups = 5;       // GDR_OPT_FUSED_GROUPS: 5 = five 128-thread top-k groups running the lean select (default), 9 = nine 64-thread groups
    // cluster-sharded corpus (gdr_store_create_shard): this handle's emb holds global rows [row_lo, row_lo + n_local) = clusters [c_lo, c_hi)
    int32_t c_lo = 0, c_hi = 0, row_lo = 0;
    int64_t n_local = 0;
    // peer-to-peer candidate exchange (gdr_store_p2p_*): one buffer per handle = [b_own * stride] fp32 scores | [GDR_MAX_RANKS] arrival flags
    int32_t n_ranks = 1, my_rank = 0, b_own = 0, p2p_K = 0;
    void *p2p_buf = nullptr;                       // this rank's buffer (cudaMalloc: exportable with cudaIpcGetMemHandle)
    size_t p2p_score_bytes = 0;
    void *peer_buf[GDR_MAX_RANKS] = {nullptr};     // every rank's buffer as mapped into this process (own entry = p2p_buf)
    bool peer_opened[GDR_MAX_RANKS] = {false};     // mapped with cudaIpcOpenMemHandle (to be closed)
    int32_t *sig_epoch = nullptr;                  // device int: scoring launches of this handle
    long long *dbg = nullptr;   // device timeline scratch for GDR_UMMA_TRACE
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct gdr_trie {
    int32_t *first_child = nullptr, *child_tok = nullptr, *child_node = nullptr;
    int32_t n_nodes = 0, n_edges = 0;
    int32_t *child_order = nullptr;   // [n_edges] edges of every node in the reference's child insertion order (default: by token)
    std::vector<int> level_start;     // nodes are numbered breadth-first: depth d = ids [level_start[d], level_start[d+1])
    int32_t fanout = 1;          // max children of any node (>= 1: the off-tree EOS candidate)
    void *cand_ws = nullptr;     // beam-step candidates: [R, fanout] fp32 values | [R, fanout] int32 flat token ids
    size_t cand_ws_bytes = 0;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" {

int gdr_abi_version(void) { return 1; }

const char *gdr_last_error(void) { return g_last_error.c_str(); }

static int store_create_common(gdr_store_t **out, const void *emb, int64_t n_docs, int32_t dim, int32_t dtype, const int32_t *offsets,
                               int32_t n_clusters, const int32_t *docid, int32_t max_cluster_size, int64_t n_local, int32_t c_lo, int32_t c_hi,
                               int64_t row_lo);

int gdr_store_create(gdr_store_t **out, const void *emb, int64_t n_docs, int32_t dim, int32_t dtype,
                     const int32_t *offsets, int32_t n_clusters, const int32_t *docid, int32_t max_cluster_size) {
    return store_create_common(out, emb, n_docs, dim, dtype, offsets, n_clusters, docid, max_cluster_size, n_docs, 0, n_clusters, 0);
}

int gdr_store_create_shard(gdr_store_t **out, const void *emb_local, int64_t n_local_rows, int32_t dim, int32_t dtype,
                           const int32_t *offsets_global, int32_t n_clusters_global, const int32_t *docid_global, int64_t n_docs_global,
                           int32_t max_cluster_size, int32_t c_lo, int32_t c_hi, int64_t row_lo) {
    if (c_lo < 0 || c_hi < c_lo || c_hi > n_clusters_global) return invalid("gdr_store_create_shard: need 0 <= c_lo <= c_hi <= n_clusters");
    if (row_lo < 0 || n_local_rows <= 0 || row_lo + n_local_rows > n_docs_global) return invalid("gdr_store_create_shard: local rows outside the corpus");
    return store_create_common(out, emb_local, n_docs_global, dim, dtype, offsets_global, n_clusters_global, docid_global, max_cluster_size,
                               n_local_rows, c_lo, c_hi, row_lo);
}

static int store_create_common(gdr_store_t **out, const void *emb, int64_t n_docs, int32_t dim, int32_t dtype, const int32_t *offsets,
                               int32_t n_clusters, const int32_t *docid, int32_t max_cluster_size, int64_t n_local, int32_t c_lo, int32_t c_hi,
                               int64_t row_lo) {
    if (!out) return invalid("gdr_store_create: out is null");
    *out = nullptr;
    if (!emb || !offsets || !docid) return invalid("gdr_store_create: null device pointer");
    if (n_docs <= 0 || n_docs > INT_MAX) return invalid("gdr_store_create: n_docs must be in [1, 2^31)");
    if (dtype != GDR_DTYPE_F32 && dtype != GDR_DTYPE_BF16) return invalid("gdr_store_create: dtype must be F32 or BF16");
    if (dim <= 0 || dim % 8 != 0) return invalid("gdr_store_create: dim must be a positive multiple of 8");
    if (dim > MAX_DIM) {
        set_error("gdr_store_create: dim > 1024 is not supported");
        return GDR_ERR_UNSUPPORTED;
    }
    if (n_clusters <= 0) return invalid("gdr_store_create: n_clusters must be positive");
    if (max_cluster_size <= 0 || max_cluster_size > n_docs) return invalid("gdr_store_create: bad max_cluster_size");
    if (reinterpret_cast<uintptr_t>(emb) & 15) return invalid("gdr_store_create: emb must be 16-byte aligned");
    gdr_store *s = new (std::nothrow) gdr_store();
    if (!s) return GDR_ERR_NOMEM;
    s->emb = emb; s->n_docs = n_docs; s->dim = dim; s->dtype = dtype;
    s->offsets = offsets; s->n_clusters = n_clusters; s->docid = docid; s->max_cluster = max_cluster_size;
    s->n_local = n_local; s->c_lo = c_lo; s->c_hi = c_hi; s->row_lo = (int32_t)row_lo;
    cudaError_t e = cudaGetDevice(&s->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    const size_t n = (size_t)n_clusters;
    const size_t words = n + 3 * (n + 1) + CTR_COUNT + 4 * (n / 8192 + 2);
    if (e == cudaSuccess) e = cudaMalloc(&s->cluster_ws, words * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(s->cluster_ws, 0, words * sizeof(int32_t));
    if (e != cudaSuccess) {
        delete s;
        return cuda_fail(e, "gdr_store_create");
    }
    if (dtype == GDR_DTYPE_BF16 && dim % 64 == 0) s->has_tmap = umma_make_tensor_map(&s->tmap, emb, n_local, dim);   // the rows this handle holds
#ifdef GDR_DEBUG_KNOBS
    // Measurement builds only (python -m gdr_b200._build --debug-knobs): timing experiments that switch pipeline stages off or
    // cut the top-k short — results are INVALID under them, so the product build does not contain them.
    if (const char *env = getenv("GDR_UMMA_DEBUG")) s->debug_flags |= (uint32_t)atoi(env) << 27;
    if (const char *env = getenv("GDR_TOPK_DEBUG")) s->debug_flags |= ((uint32_t)atoi(env) & 15u) << 20;
    if (getenv("GDR_UMMA_TRACE")) { cudaMalloc(&s->dbg, 512 * sizeof(long long)); cudaMemset(s->dbg, 0, 512 * sizeof(long long)); }
#endif
    *out = s;
    return GDR_OK;
}

int gdr_store_destroy(gdr_store_t *s) {
    if (!s) return GDR_OK;
    cudaFree(s->cluster_ws);
    cudaFree(s->batch_ws);
    cudaFree(s->dbg);
    for (int r = 0; r < GDR_MAX_RANKS; ++r)
        if (s->peer_opened[r]) cudaIpcCloseMemHandle(s->peer_buf[r]);
    cudaFree(s->p2p_buf);
    cudaFree(s->sig_epoch);
    for (auto &e : s->ev)
        if (e) cudaEventDestroy(e);
    delete s;
    return GDR_OK;
}

}  // extern "C"

// Layout of the per-batch scratch for one (B, K, k, flags) shape: offsets into batch_ws and the total size.
namespace {
struct ScratchPlan {
    int64_t pairs, stride, q_rows;
    bool umma_possible, tile_possible, global_keys, small_topk;
    size_t o_pair, o_cand, o_cbase, o_simt, o_umma, o_score, o_qsplit, o_tmeta, o_keys, o_ghist, total;
};

ScratchPlan plan_scratch(const gdr_store *s, int32_t B, int32_t K, int32_t k, uint32_t flags) {
    ScratchPlan p;
    p.pairs = (int64_t)B * K;
    p.stride = (int64_t)align_up((size_t)K * s->max_cluster, 4);
    const int64_t simt_cap = p.pairs * ((s->max_cluster + SIMT_ROWS - 1) / SIMT_ROWS);
    const int64_t umma_cap = p.pairs * ((s->max_cluster + UMMA_ROWS - 1) / UMMA_ROWS);
    p.umma_possible = s->has_tmap && !(flags & GDR_FORCE_SIMT);
    // fp32 stores with dense groups: the shared-memory-tiled fp32 kernel (score_tile_f32.cu) walks the same tile records as the tcgen05 path
    p.tile_possible = s->dtype == GDR_DTYPE_F32 && s->dim % 32 == 0 && !(flags & GDR_FORCE_SIMT);
    // k <= 128: the fast top-k keeps no key array (its mass-tie fallback uses the global scratch); larger k: keys in smem if they fit
    p.global_keys = k <= 128 || (size_t)p.stride * 4 + 8 * 4096 + 4 * 2048 + (size_t)(3 * K + 1) * 4 > 96 * 1024;
    p.small_topk = k <= 128 && p.stride <= 65535 && !s->topk_wide;     // 128-thread top-k CTAs with 16-bit histogram bins
    p.q_rows = (flags & GDR_Q_PER_BEAM) ? p.pairs : B;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    p.o_pair = take(p.pairs * 4);
    p.o_cand = take((size_t)B * (K + 1) * 4);
    p.o_cbase = take(p.pairs * 4);
    p.o_simt = take((size_t)simt_cap * sizeof(Item));
    p.o_umma = take(p.umma_possible || p.tile_possible ? (size_t)umma_cap * sizeof(Item) : 0);
    p.o_score = take(s->n_ranks > 1 ? 0 : (size_t)B * p.stride * 4);       // sharded exchange: the scores live in the p2p buffer
    p.o_qsplit = take(p.umma_possible ? (size_t)p.q_rows * 3 * s->dim * 2 : 0);
    p.o_tmeta = take(p.umma_possible || p.tile_possible ? (size_t)umma_cap * sizeof(TileMeta) : 0);
    p.o_keys = take(p.global_keys ? (size_t)B * p.stride * 4 : 0);
    p.o_ghist = take(p.small_topk ? (size_t)B * 2048 * 4 : 0);
    p.total = off;
    return p;
}

int check_shape(const gdr_store *s, int32_t B, int32_t K, int32_t k, uint32_t flags, const char *who) {
    if (B < 0 || K <= 0 || k <= 0) return invalid((std::string(who) + ": need B >= 0, K > 0, k > 0").c_str());
    if (k > 4096) {
        set_error(std::string(who) + ": k > 4096 is not supported");
        return GDR_ERR_UNSUPPORTED;
    }
    if ((int64_t)B * K > INT_MAX / 2) return invalid((std::string(who) + ": B*K too large").c_str());
    if ((int64_t)B * K * s->max_cluster > INT_MAX - 4)
        return invalid((std::string(who) + ": B*K*max_cluster_size must be below 2^31 (score buffer index)").c_str());
    if ((flags & GDR_FORCE_UMMA) && !s->has_tmap) {
        set_error(std::string(who) + ": GDR_FORCE_UMMA needs a bf16 store with dim % 64 == 0");
        return GDR_ERR_UNSUPPORTED;
    }
    return GDR_OK;
}

// Growing the scratch synchronises the stream and reallocates: gdr_store_reserve does it ahead of the query path.
int ensure_scratch(gdr_store *s, size_t bytes, cudaStream_t st) {
    if (bytes <= s->batch_ws_bytes) return GDR_OK;
    GDR_CUDA(cudaStreamSynchronize(st));
    if (s->batch_ws) GDR_CUDA(cudaFree(s->batch_ws));
    s->batch_ws = nullptr;
    s->batch_ws_bytes = 0;
    GDR_CUDA(cudaMalloc(&s->batch_ws, bytes));
    s->batch_ws_bytes = bytes;
    s->last_valid = false;        // the work lists of the previous batch went with the old buffer
    return GDR_OK;
}
}  // namespace

extern "C" {

int gdr_store_reserve(gdr_store_t *s, int32_t B, int32_t K, int32_t k, uint32_t flags, void *stream) {
    if (!s) return invalid("gdr_store_reserve: store is null");
    if (int rc = check_shape(s, B, K, k, flags, "gdr_store_reserve")) return rc;
    if (B == 0) return GDR_OK;
    return ensure_scratch(s, plan_scratch(s, B, K, k, flags).total, (cudaStream_t)stream);
}

// ---- peer-to-peer candidate exchange of a cluster-sharded corpus (include/gdr_b200.h) --------------------------------------------
int gdr_store_p2p_init(gdr_store_t *s, int32_t n_ranks, int32_t my_rank, int32_t b_own, int32_t K, void *ipc_handle_out) {
    if (!s) return invalid("gdr_store_p2p_init: store is null");
    if (n_ranks < 2 || n_ranks > GDR_MAX_RANKS || my_rank < 0 || my_rank >= n_ranks) return invalid("gdr_store_p2p_init: need 2 <= n_ranks <= 8 and 0 <= my_rank < n_ranks");
    if (b_own <= 0 || K <= 0) return invalid("gdr_store_p2p_init: need b_own > 0, K > 0");
    if (s->p2p_buf) return invalid("gdr_store_p2p_init: already initialised");
    const size_t stride = align_up((size_t)K * s->max_cluster, 4);
    if ((int64_t)b_own * (int64_t)stride >= (1ll << OFF_OWNER_SHIFT)) return invalid("gdr_store_p2p_init: b_own * K * max_cluster_size must be below 2^28");
    const size_t score_bytes = align_up((size_t)b_own * stride * 4, 256);
    GDR_CUDA(cudaMalloc(&s->p2p_buf, score_bytes + 256));
    GDR_CUDA(cudaMemset(s->p2p_buf, 0, score_bytes + 256));
    GDR_CUDA(cudaMalloc(&s->sig_epoch, 256));
    GDR_CUDA(cudaMemset(s->sig_epoch, 0, 256));
    s->p2p_score_bytes = score_bytes;
    s->n_ranks = n_ranks; s->my_rank = my_rank; s->b_own = b_own; s->p2p_K = K;
    s->peer_buf[my_rank] = s->p2p_buf;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(h) == GDR_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
        GDR_CUDA(cudaIpcGetMemHandle(&h, s->p2p_buf));
        memcpy(ipc_handle_out, &h, sizeof(h));
    }
    GDR_CUDA(cudaDeviceSynchronize());
    s->last_valid = false;
    return GDR_OK;
}

int gdr_store_p2p_attach(gdr_store_t *s, const void *all_handles) {
    if (!s || !all_handles) return invalid("gdr_store_p2p_attach: null argument");
    if (!s->p2p_buf) return invalid("gdr_store_p2p_attach: call gdr_store_p2p_init first");
    for (int r = 0; r < s->n_ranks; ++r) {
        if (r == s->my_rank || s->peer_buf[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const char *>(all_handles) + (size_t)r * GDR_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        GDR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_buf[r] = p;
        s->peer_opened[r] = true;
    }
    return GDR_OK;
}

int gdr_store_p2p_attach_local(gdr_store_t *s, gdr_store_t *const *peers) {
    if (!s || !peers) return invalid("gdr_store_p2p_attach_local: null argument");
    if (!s->p2p_buf) return invalid("gdr_store_p2p_attach_local: call gdr_store_p2p_init first");
    for (int r = 0; r < s->n_ranks; ++r) {
        if (r == s->my_rank) continue;
        const gdr_store *o = peers[r];
        if (!o || !o->p2p_buf || o->n_ranks != s->n_ranks || o->my_rank != r || o->b_own != s->b_own || o->p2p_K != s->p2p_K ||
            o->p2p_score_bytes != s->p2p_score_bytes)
            return invalid("gdr_store_p2p_attach_local: peer handles must be initialised with the same n_ranks / b_own / K and their own rank");
        int dev_o = o->device, can = 1;
        if (dev_o != s->device) {
            GDR_CUDA(cudaDeviceCanAccessPeer(&can, s->device, dev_o));
            if (!can) {
                set_error("gdr_store_p2p_attach_local: no peer access between the two devices");
                return GDR_ERR_UNSUPPORTED;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(dev_o, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        s->peer_buf[r] = o->p2p_buf;
    }
    return GDR_OK;
}

int gdr_store_set_option(gdr_store_t *s, int32_t option, int32_t value) {
    if (!s) return invalid("gdr_store_set_option: store is null");
    switch (option) {
    case GDR_OPT_UMMA_CTAS:
        if (value < 0 || value > 1024) return invalid("gdr_store_set_option: GDR_OPT_UMMA_CTAS must be in [0, 1024]");
        s->umma_ctas = value;
        return GDR_OK;
    case GDR_OPT_UMMA_CTAS_PER_SM:
        if (value != 1 && value != 2) return invalid("gdr_store_set_option: GDR_OPT_UMMA_CTAS_PER_SM must be 1 or 2");
        s->umma_ctas_per_sm = value;
        return GDR_OK;
    case GDR_OPT_UMMA_MIN_GROUP:
        if (value < 1) return invalid("gdr_store_set_option: GDR_OPT_UMMA_MIN_GROUP must be >= 1");
        s->umma_min_group = value;
        return GDR_OK;
    case GDR_OPT_LAUNCH_PRIORITIES: {
        // inversion kernels at the greatest priority (they are tiny and otherwise queue behind the 1,024-CTA top-k grid),
        // scoring one below, top-k at the least — so a pending scoring grid takes a freed SM first
        s->prio_invert = s->prio_score = s->prio_topk = 0;
        int least = 0, greatest = 0;
        if (value > 0 && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess && greatest < least) {
            s->prio_invert = greatest + 1000;
            s->prio_score = (greatest + 1 <= least ? greatest + 1 : least) + 1000;
            s->prio_topk = least + 1000;
        }
        return GDR_OK;
    }
    case GDR_OPT_FUSED_GROUPS:
        if (value != 5 && value != 9) return invalid("gdr_store_set_option: GDR_OPT_FUSED_GROUPS must be 5 (128-thread groups) or 9 (64-thread groups)");
        s->fused_groups = value;
        return GDR_OK;
    case GDR_OPT_TOPK_GROUPS:
        if (value != 0 && value != 1 && value != 2 && value != 4) return invalid("gdr_store_set_option: GDR_OPT_TOPK_GROUPS must be 0, 1, 2 or 4");
        s->debug_flags = (s->debug_flags & ~(7u << 16)) | ((uint32_t)value << 16);
        return GDR_OK;
    case GDR_OPT_TOPK_WIDE:
        s->topk_wide = value != 0;
        return GDR_OK;
    default:
        return invalid("gdr_store_set_option: unknown option");
    }
}

int gdr_score_topk(gdr_store_t *s, const float *q, const int32_t *beams, const float *prob, const float *alphas,
                   int32_t n_alpha, int32_t B, int32_t K, int32_t act, int32_t k, uint32_t flags, float *out_scores,
                   int32_t *out_docids, void *stream) {
    if (!s) return invalid("gdr_score_topk: store is null");
    if (int rc = check_shape(s, B, K, k, flags, "gdr_score_topk")) return rc;
    if (B == 0) return GDR_OK;
    if (!q || !beams || !out_scores || !out_docids) return invalid("gdr_score_topk: null pointer");
    if (reinterpret_cast<uintptr_t>(q) & 15) return invalid("gdr_score_topk: q must be 16-byte aligned");
    if (act < GDR_ACT_NONE || act > GDR_ACT_SIGMOID) return invalid("gdr_score_topk: bad activation");
    if (n_alpha < 1 || (!alphas && n_alpha != 1)) return invalid("gdr_score_topk: n_alpha must be >= 1 (1 when alphas is null)");
    cudaStream_t st = (cudaStream_t)stream;
    const ScratchPlan p = plan_scratch(s, B, K, k, flags);
    const int64_t pairs = p.pairs;
    // the scratch grows on the first call of a larger shape (that call synchronises): gdr_store_reserve moves this off the query path
    if (int rc = ensure_scratch(s, p.total, st)) return rc;
    char *ws = reinterpret_cast<char *>(s->batch_ws);
    const size_t n = (size_t)s->n_clusters;

    ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.emb = s->emb; a.offsets = s->offsets; a.docid = s->docid;
    a.dim = s->dim; a.dtype = s->dtype; a.n_clusters = s->n_clusters; a.max_cluster = s->max_cluster; a.n_docs = s->n_docs;
    a.q = q; a.beams = beams; a.prob = prob; a.B = B; a.K = K; a.act = act; a.k = k; a.flags = flags;
    a.flags |= s->debug_flags;                                         // GDR_OPT_TOPK_GROUPS (and, in measurement builds, the timing experiments)
    a.cnt = s->cluster_ws;
    a.grp_off = s->cluster_ws + n;
    a.simt_off = a.grp_off + (n + 1);
    a.umma_off = a.simt_off + (n + 1);
    a.counters = a.umma_off + (n + 1);
    a.scan_base = a.counters + CTR_COUNT;
    a.grp_pair = reinterpret_cast<int32_t *>(ws + p.o_pair);
    a.candoff = reinterpret_cast<int32_t *>(ws + p.o_cand);
    a.cbase = reinterpret_cast<int32_t *>(ws + p.o_cbase);
    a.simt_items = reinterpret_cast<Item *>(ws + p.o_simt);
    a.umma_items = reinterpret_cast<Item *>(ws + p.o_umma);
    a.scorebuf = reinterpret_cast<float *>(ws + p.o_score);
    a.stride = p.stride;
    a.c_lo = s->c_lo; a.c_hi = s->c_hi; a.row_lo = s->row_lo;
    a.n_ranks = 1; a.my_rank = 0; a.b_own = B; a.q_base = 0; a.B_top = B;
    if (s->n_ranks > 1) {
        // sharded exchange: B is the GLOBAL batch (n_ranks x b_own queries, the same on every rank); this rank selects the top-k of its own
        // b_own queries out of a score buffer that all ranks fill; outputs are [n_alpha, b_own, k]
        if (B != s->n_ranks * s->b_own || K != s->p2p_K) return invalid("gdr_score_topk: a p2p handle takes batches of exactly n_ranks * b_own queries with the K given to gdr_store_p2p_init");
        if ((int64_t)s->b_own * p.stride >= (1ll << OFF_OWNER_SHIFT)) return invalid("gdr_score_topk: b_own * K * max_cluster_size must be below 2^28 on a p2p handle");
        a.n_ranks = s->n_ranks; a.my_rank = s->my_rank; a.b_own = s->b_own; a.q_base = s->my_rank * s->b_own; a.B_top = s->b_own;
        a.scorebuf = reinterpret_cast<float *>(s->p2p_buf);
        for (int r = 0; r < s->n_ranks; ++r) {
            a.peer_score[r] = reinterpret_cast<float *>(s->peer_buf[r]);
            a.peer_sig[r] = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(s->peer_buf[r]) + s->p2p_score_bytes);
        }
        a.sig_local = a.peer_sig[s->my_rank];
        a.sig_epoch = s->sig_epoch;
    }
    a.gkeys = p.global_keys ? reinterpret_cast<uint32_t *>(ws + p.o_keys) : nullptr;
    a.ghist = p.small_topk ? reinterpret_cast<uint32_t *>(ws + p.o_ghist) : nullptr;
    a.qsplit = nullptr;   // set below when the tcgen05 path is taken
    // One scoring path per call: a batch that names each cluster three or more times on average goes to the tcgen05
    // grouped GEMM (slab read once for the whole group); a sparse batch goes to the SIMT GEMV, which serves up to four
    // pairs per slab read and measured 89% of HBM peak at ~1 pair per cluster (cfg5 slice) against 77% for tcgen05.
    // GDR_OPT_UMMA_MIN_GROUP > 1 asks for the mixed mode instead: groups of at least that many pairs on tensor cores, the rest SIMT.
    const bool umma_possible = p.umma_possible;
    const bool mixed = umma_possible && !(flags & GDR_FORCE_UMMA) && s->umma_min_group > 1;
    bool use_umma = umma_possible && ((flags & GDR_FORCE_UMMA) || mixed || pairs >= 3 * (int64_t)s->n_clusters);
    const bool use_tile = p.tile_possible && pairs >= 3 * (int64_t)s->n_clusters;       // fp32 store, dense groups: tiled fp32 kernel, same tile queue
    const bool use_simt = (!use_umma && !use_tile) || mixed;
    a.dbg = s->dbg;
    if (use_umma) a.qsplit = reinterpret_cast<__nv_bfloat16 *>(ws + p.o_qsplit);
    if (use_umma || use_tile) a.tile_meta = reinterpret_cast<TileMeta *>(ws + p.o_tmeta);
    a.umma_min_group = (!use_umma && !use_tile) ? INT_MAX : (mixed ? s->umma_min_group : 1);

    int launches = 0;
    const bool prof = s->profiling;
    if (s->dbg && !(flags & GDR_SKIP_INVERT)) {   // kernel timeline: min-start slots to +inf, max-end slots to 0
        static const unsigned long long init[6] = {~0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull};
        GDR_CUDA(cudaMemcpyAsync(s->dbg + 500, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    if (prof) GDR_CUDA(cudaEventRecord(s->ev[0], st));
    a.launch_prio = s->prio_invert;               // per call, carried in this call's own copy of the arguments
    if (!(flags & GDR_SKIP_INVERT)) GDR_CUDA(launch_invert(a, st, &launches));
    const int l_inv = launches;
    a.launch_prio = s->prio_score;
    if (prof) GDR_CUDA(cudaEventRecord(s->ev[1], st));
    if (a.n_ranks > 1 && !(flags & GDR_SKIP_SCORE)) {     // sharded corpus: every owner's top-k of this handle's previous batch has read its scores
        GDR_CUDA(launch_wait_consumed(a, st));
        launches += 1;
    }
    if (use_umma && !(flags & GDR_SKIP_SCORE)) {
        a.signal = use_simt ? 0 : 1;              // (mixed mode: the GEMV kernel is the call's last scoring kernel and signals)
        if (s->umma_ctas_per_sm == 2 && a.n_ranks == 1)
            GDR_CUDA(launch_score_umma_x2(a, &s->tmap, st, s->umma_ctas > 0 ? s->umma_ctas : 2 * s->sm_count));
        else
            GDR_CUDA(launch_score_umma(a, &s->tmap, st, s->umma_ctas > 0 ? s->umma_ctas : s->sm_count));
        launches += 1;
    }
    if (use_tile && !(flags & GDR_SKIP_SCORE)) {
        a.signal = 1;
        GDR_CUDA(launch_score_tile_f32(a, st, s->sm_count));
        launches += 1;
    }
    if (prof) GDR_CUDA(cudaEventRecord(s->ev[2], st));
    if (use_simt && !(flags & GDR_SKIP_SCORE)) {
        a.signal = 1;
        GDR_CUDA(launch_score_simt(a, st, s->sm_count));
        launches += 1;
    }
    if (prof) GDR_CUDA(cudaEventRecord(s->ev[3], st));
    const int l_score = launches - l_inv;
    a.launch_prio = s->prio_topk;
    a.wait_in_topk = 0;
    if (a.n_ranks > 1 && !(flags & GDR_SKIP_TOPK)) {      // sharded corpus: one warp waits for every rank's scores, then the top-k grid runs
        GDR_CUDA(launch_wait_scorers(a, st));
        launches += 1;
    }
    for (int r = 0; r < n_alpha && !(flags & GDR_SKIP_TOPK); ++r) {
        const float alpha = alphas ? alphas[r] : 1.0f;
        GDR_CUDA(launch_topk_store(a, alpha, out_scores + (int64_t)r * a.B_top * k, out_docids + (int64_t)r * a.B_top * k, st));
        launches += 1;
    }
    if (a.n_ranks > 1 && !(flags & GDR_SKIP_TOPK)) {      // ... and tell every rank that this batch's scores have been read here
        GDR_CUDA(launch_signal_consumed(a, st));
        launches += 1;
    }
    if (prof) GDR_CUDA(cudaEventRecord(s->ev[4], st));
    if (!(flags & GDR_SKIP_INVERT)) s->phase_launches[0] = l_inv;
    if (!(flags & GDR_SKIP_SCORE)) s->phase_launches[1] = l_score;
    if (!(flags & GDR_SKIP_TOPK)) s->phase_launches[2] = launches - l_inv - l_score;
    if (!(flags & GDR_SKIP_INVERT)) {             // a new batch enters the handle: phases not issued yet count as zero
        if (flags & GDR_SKIP_SCORE) s->phase_launches[1] = 0;
        if (flags & GDR_SKIP_TOPK) s->phase_launches[2] = 0;
    }
    s->last_launches = s->phase_launches[0] + s->phase_launches[1] + s->phase_launches[2];
    a.launch_prio = s->prio_score;                // what gdr_score_fused launches with
    a.signal = 1;
    a.wait_in_topk = 1;                           // ... and its top-k groups wait for the arrival flags themselves
    s->last_args = a;
    s->last_valid = true;
    s->last_umma_only = use_umma && !use_simt;
    return GDR_OK;
}

// Pipelined schedule: scoring of the batch in `cur` + top-k of the batch in `prev` in one launch (csrc/score_fused.cu).
int gdr_score_fused(gdr_store_t *cur, gdr_store_t *prev, float alpha, float *prev_out_scores, int32_t *prev_out_docids, void *stream) {
    if (!cur && !prev) return invalid("gdr_score_fused: both handles are null");
    if (cur == prev) return invalid("gdr_score_fused: cur and prev must be different handles (two scratch sets)");
    cudaStream_t st = (cudaStream_t)stream;
    ScoreArgs pa;
    memset(&pa, 0, sizeof(pa));
    if (prev) {
        if (!prev->last_valid) return invalid("gdr_score_fused: prev has no batch (call gdr_score_topk on it first)");
        if (!prev_out_scores || !prev_out_docids) return invalid("gdr_score_fused: null output pointer");
        pa = prev->last_args;
        if (pa.k > 128 || !pa.gkeys || !pa.ghist || pa.stride > 65535) {
            set_error("gdr_score_fused: the fused top-k serves k <= 128 and <= 65,535 candidates per query");
            return GDR_ERR_UNSUPPORTED;
        }
    }
    if (!cur) {                                                    // flush: the last batch's top-k alone
        if (pa.n_ranks > 1) {
            pa.wait_in_topk = 0;
            GDR_CUDA(launch_wait_scorers(pa, st));
        }
        GDR_CUDA(launch_topk_store(pa, alpha, prev_out_scores, prev_out_docids, st));
        if (pa.n_ranks > 1) GDR_CUDA(launch_signal_consumed(pa, st));
        return GDR_OK;
    }
    if (!cur->last_valid) return invalid("gdr_score_fused: cur has no inversion (call gdr_score_topk with GDR_SKIP_SCORE | GDR_SKIP_TOPK first)");
    if (!cur->last_umma_only || !cur->has_tmap) {
        set_error("gdr_score_fused: the batch in cur does not take the tcgen05 path alone");
        return GDR_ERR_UNSUPPORTED;
    }
    // The fused CTA takes a whole SM (1,024 threads, all registers): the inversion kernels of the NEXT batch, which run beside it on a
    // second stream, need SMs of their own — by default the grid leaves eight (108-140 scoring CTAs measured the same speed).
    const int ctas = cur->umma_ctas > 0 ? cur->umma_ctas : (cur->sm_count > 16 ? cur->sm_count - 8 : cur->sm_count);
    GDR_CUDA(launch_score_fused(cur->last_args, &cur->tmap, pa, alpha, prev_out_scores, prev_out_docids, st, ctas, cur->fused_groups));
    return GDR_OK;
}

int gdr_store_set_profiling(gdr_store_t *s, int32_t enable) {
    if (!s) return invalid("gdr_store_set_profiling: store is null");
    if (enable && !s->ev[0])
        for (auto &e : s->ev) GDR_CUDA(cudaEventCreate(&e));
    s->profiling = enable != 0;
    return GDR_OK;
}

int gdr_store_last_phase_ms(gdr_store_t *s, float out[4]) {
    if (!s || !out) return invalid("gdr_store_last_phase_ms: null argument");
    if (!s->ev[0]) return invalid("gdr_store_last_phase_ms: profiling was never enabled");
    GDR_CUDA(cudaEventSynchronize(s->ev[4]));
    for (int i = 0; i < 4; ++i) GDR_CUDA(cudaEventElapsedTime(&out[i], s->ev[i], s->ev[i + 1]));
    return GDR_OK;
}

int gdr_store_last_stats(gdr_store_t *s, int64_t out[4], void *stream) {
    if (!s || !out) return invalid("gdr_store_last_stats: null argument");
    int32_t c[CTR_COUNT];
    GDR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (s->dbg) {   // GDR_UMMA_TRACE: dump the device timeline of CTA 0 (ns, relative to kernel start)
        long long h[512];
        GDR_CUDA(cudaMemcpy(h, s->dbg, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[umma trace] ");
        for (int i = 1; i < 200 && h[i]; ++i) fprintf(stderr, "%lld ", h[i] - h[0]);
        long long smin = 1LL << 62, smax = 0, emin = 1LL << 62, emax = 0;
        for (int c = 0; c < 148; ++c) {
            const long long st = h[200 + 2 * c] - h[0], en = h[201 + 2 * c] - h[0];
            if (!h[200 + 2 * c]) continue;
            smin = st < smin ? st : smin; smax = st > smax ? st : smax; emin = en < emin ? en : emin; emax = en > emax ? en : emax;
        }
        fprintf(stderr, "| CTA loop start min %lld max %lld, end min %lld max %lld\n", smin, smax, emin, emax);
        fprintf(stderr, "[timeline] %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", h[500], h[501], h[502], h[503], h[504], h[505],
                smin + h[0], smax + h[0], emin + h[0], emax + h[0]);
    }
    const size_t n = (size_t)s->n_clusters;
    GDR_CUDA(cudaMemcpy(c, s->cluster_ws + n + 3 * (n + 1), sizeof(c), cudaMemcpyDeviceToHost));
    out[0] = c[CTR_N_SIMT];
    out[1] = c[CTR_N_UMMA];
    out[2] = s->last_launches;
    out[3] = c[CTR_N_TOUCHED];
    return GDR_OK;
}

int gdr_cluster_centroids(gdr_store_t *s, float *out, void *stream) {
    if (!s || !out) return invalid("gdr_cluster_centroids: null argument");
    GDR_CUDA(launch_centroids(s->emb, s->dtype, s->offsets, s->n_clusters, s->dim, out, (cudaStream_t)stream));
    return GDR_OK;
}

int gdr_contrastive_loss(gdr_store_t *s, const float *q, const int32_t *pos_rows, const int32_t *cand_rows, const int32_t *cand_off,
                         int32_t B, int32_t S, int32_t act, float tau, float intra_rate, float *loss_per_query, float *loss,
                         float *grad_q, void *stream) {
    if (!s) return invalid("gdr_contrastive_loss: store is null");
    if (B <= 0 || S < 0) return invalid("gdr_contrastive_loss: need B > 0, S >= 0");
    if (!q || !pos_rows || !cand_off || (S > 0 && !cand_rows) || !loss_per_query || !loss) return invalid("gdr_contrastive_loss: null pointer");
    if (act < GDR_ACT_NONE || act > GDR_ACT_SIGMOID) return invalid("gdr_contrastive_loss: bad activation");
    if (!(tau > 0.f)) return invalid("gdr_contrastive_loss: tau must be positive");
    if ((size_t)(S + 1) * 8 > 200 * 1024) {
        set_error("gdr_contrastive_loss: more than 25,599 candidates per batch is not supported");
        return GDR_ERR_UNSUPPORTED;
    }
    GDR_CUDA(launch_contrastive(s->emb, s->dtype, s->dim, q, pos_rows, cand_rows, cand_off, B, S, act, tau, intra_rate, loss_per_query, loss,
                                grad_q, (cudaStream_t)stream));
    return GDR_OK;
}

int gdr_similarity(const float *q, int64_t Q, const void *p, int64_t P, int32_t dim, int32_t p_dtype, float *out,
                   void *stream) {
    if (Q < 0 || P < 0) return invalid("gdr_similarity: negative size");
    if (Q == 0 || P == 0) return GDR_OK;
    if (!q || !p || !out) return invalid("gdr_similarity: null pointer");
    if (p_dtype != GDR_DTYPE_F32 && p_dtype != GDR_DTYPE_BF16) return invalid("gdr_similarity: bad dtype");
    if (dim <= 0 || dim % 8 != 0) return invalid("gdr_similarity: dim must be a positive multiple of 8");
    if (dim > MAX_DIM) {
        set_error("gdr_similarity: dim > 1024 is not supported");
        return GDR_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(p) & 15) || (reinterpret_cast<uintptr_t>(q) & 15))
        return invalid("gdr_similarity: q and p must be 16-byte aligned");
    int dev = 0, sms = 148;
    GDR_CUDA(cudaGetDevice(&dev));
    GDR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // bf16 passages with dim % 64 == 0 and enough work to fill tiles: the tcgen05 grouped GEMM (similarity.cu); everything else: the GEMV
    if (p_dtype == GDR_DTYPE_BF16 && dim % 64 == 0 && P >= 64 && Q >= 8) {
        const cudaError_t e = launch_similarity_umma(q, Q, p, P, dim, out, (cudaStream_t)stream, sms);
        if (e == cudaSuccess) return GDR_OK;
        if (e != cudaErrorNotSupported) return cuda_fail(e, "gdr_similarity (tensor-core path)");
        cudaGetLastError();
    }
    GDR_CUDA(launch_similarity(q, Q, p, P, dim, p_dtype, out, (cudaStream_t)stream, sms));
    return GDR_OK;
}

int gdr_merge_topk(const float *scores, const int32_t *docids, int32_t G, int32_t B, int32_t k_in, int64_t g_stride,
                   int32_t k, float *out_scores, int32_t *out_docids, void *stream) {
    if (G <= 0 || B < 0 || k_in <= 0 || k <= 0) return invalid("gdr_merge_topk: bad sizes");
    if (g_stride < (int64_t)B * k_in) return invalid("gdr_merge_topk: g_stride smaller than one rank's block");
    if (B == 0) return GDR_OK;
    if (!scores || !docids || !out_scores || !out_docids) return invalid("gdr_merge_topk: null pointer");
    if (k > 4096 || (int64_t)G * k_in > 32768) {
        set_error("gdr_merge_topk: need k <= 4096 and G*k_in <= 32768");
        return GDR_ERR_UNSUPPORTED;
    }
    GDR_CUDA(launch_merge_topk(scores, docids, G, B, k_in, g_stride, k, out_scores, out_docids, (cudaStream_t)stream));
    return GDR_OK;
}

int gdr_trie_create(gdr_trie_t **out, const int32_t *first_child, const int32_t *child_tok, const int32_t *child_node,
                    int32_t n_nodes, int32_t n_edges) {
    if (!out) return invalid("gdr_trie_create: out is null");
    *out = nullptr;
    if (n_nodes <= 0 || n_edges < 0 || !first_child) return invalid("gdr_trie_create: bad sizes");
    if (n_edges > 0 && (!child_tok || !child_node)) return invalid("gdr_trie_create: null edge arrays");
    if (first_child[0] != 0 || first_child[n_nodes] != n_edges) return invalid("gdr_trie_create: first_child must span [0, n_edges]");
    for (int i = 0; i < n_nodes; ++i)
        if (first_child[i + 1] < first_child[i]) return invalid("gdr_trie_create: first_child must be non-decreasing");
    for (int e = 0; e < n_edges; ++e)
        if (child_node[e] <= 0 || child_node[e] >= n_nodes) return invalid("gdr_trie_create: child_node out of range");
    gdr_trie *t = new (std::nothrow) gdr_trie();
    if (!t) return GDR_ERR_NOMEM;
    t->n_nodes = n_nodes; t->n_edges = n_edges;
    for (int i = 0; i < n_nodes; ++i) t->fanout = first_child[i + 1] - first_child[i] > t->fanout ? first_child[i + 1] - first_child[i] : t->fanout;
    const size_t ne = (size_t)(n_edges > 0 ? n_edges : 1);
    cudaError_t e = cudaMalloc(&t->first_child, (size_t)(n_nodes + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t->child_tok, ne * 4);
    if (e == cudaSuccess) e = cudaMalloc(&t->child_node, ne * 4);
    if (e == cudaSuccess) e = cudaMemcpy(t->first_child, first_child, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_edges) e = cudaMemcpy(t->child_tok, child_tok, (size_t)n_edges * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_edges) e = cudaMemcpy(t->child_node, child_node, (size_t)n_edges * 4, cudaMemcpyHostToDevice);
    // depth levels (for the bottom-up node embeddings): breadth-first numbering makes every depth a contiguous id range
    {
        std::vector<int> depth((size_t)n_nodes, -1);
        depth[0] = 0;
        bool bfs = true;
        for (int n = 0; n < n_nodes && bfs; ++n) {
            if (depth[n] < 0) { bfs = false; break; }
            for (int e2 = first_child[n]; e2 < first_child[n + 1]; ++e2) {
                if (depth[child_node[e2]] >= 0) bfs = false;
                depth[child_node[e2]] = depth[n] + 1;
            }
        }
        for (int n = 1; n < n_nodes && bfs; ++n)
            if (depth[n] < depth[n - 1]) bfs = false;
        if (bfs) {
            t->level_start.push_back(0);
            for (int n = 1; n < n_nodes; ++n)
                if (depth[n] != depth[n - 1]) t->level_start.push_back(n);
            t->level_start.push_back(n_nodes);
        }                                                    // not breadth-first: masks still work, node embeddings are refused
    }
    std::vector<int32_t> ident(ne);
    for (size_t i = 0; i < ne; ++i) ident[i] = (int32_t)i;
    if (e == cudaSuccess) e = cudaMalloc(&t->child_order, ne * 4);
    if (e == cudaSuccess) e = cudaMemcpy(t->child_order, ident.data(), ne * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        gdr_trie_destroy(t);
        return cuda_fail(e, "gdr_trie_create");
    }
    *out = t;
    return GDR_OK;
}

int gdr_trie_set_child_order(gdr_trie_t *t, const int32_t *first_child, const int32_t *order) {
    if (!t || !first_child || !order) return invalid("gdr_trie_set_child_order: null argument");
    std::vector<char> seen((size_t)(t->n_edges > 0 ? t->n_edges : 1), 0);
    for (int n = 0; n < t->n_nodes; ++n)
        for (int r = first_child[n]; r < first_child[n + 1]; ++r) {
            const int e = order[r];
            if (e < first_child[n] || e >= first_child[n + 1] || seen[e]) return invalid("gdr_trie_set_child_order: order must permute every node's own edges");
            seen[e] = 1;
        }
    if (t->n_edges) GDR_CUDA(cudaMemcpy(t->child_order, order, (size_t)t->n_edges * 4, cudaMemcpyHostToDevice));
    return GDR_OK;
}

int gdr_trie_node_embeddings(gdr_trie_t *t, const int32_t *node_cluster, const float *leaf_emb, const int32_t *leaf_num, int32_t dim,
                             float *node_emb, int32_t *node_leaf_num, void *stream) {
    if (!t || !node_cluster || !leaf_emb || !leaf_num || !node_emb || !node_leaf_num) return invalid("gdr_trie_node_embeddings: null argument");
    if (dim <= 0) return invalid("gdr_trie_node_embeddings: dim must be positive");
    if (t->level_start.empty()) return invalid("gdr_trie_node_embeddings: the trie's nodes must be numbered breadth-first");
    GDR_CUDA(launch_node_embeddings(t->first_child, t->child_node, t->child_order, node_cluster, leaf_emb, leaf_num, dim, t->level_start.data(),
                                    (int)t->level_start.size() - 1, node_emb, node_leaf_num, (cudaStream_t)stream));
    return GDR_OK;
}

int gdr_tree_match(gdr_trie_t *t, const float *node_emb, const int32_t *node_leaf_num, int32_t dim, const float *docs, int32_t M,
                   int32_t max_len, int32_t *out_tokens, int32_t *out_len, void *stream) {
    if (!t) return invalid("gdr_tree_match: trie is null");
    if (M < 0 || dim <= 0 || max_len < 2) return invalid("gdr_tree_match: need M >= 0, dim > 0, max_len >= 2");
    if (M == 0) return GDR_OK;
    if (!node_emb || !node_leaf_num || !docs || !out_tokens || !out_len) return invalid("gdr_tree_match: null pointer");
    GDR_CUDA(launch_tree_match(t->first_child, t->child_tok, t->child_node, t->child_order, node_emb, node_leaf_num, dim, docs, M, max_len,
                               out_tokens, out_len, (cudaStream_t)stream));
    return GDR_OK;
}

int gdr_trie_destroy(gdr_trie_t *t) {
    if (!t) return GDR_OK;
    cudaFree(t->first_child);
    cudaFree(t->child_tok);
    cudaFree(t->child_node);
    cudaFree(t->child_order);
    cudaFree(t->cand_ws);
    delete t;
    return GDR_OK;
}

int gdr_tree_mask(gdr_trie_t *t, const int64_t *input_ids, int64_t ids_row_stride, int32_t R, int32_t cur_len,
                  float *scores, int64_t scores_row_stride, int32_t V, int32_t eos_id, int32_t strict, void *stream) {
    if (!t) return invalid("gdr_tree_mask: trie is null");
    if (R < 0 || cur_len < 1 || V <= 0) return invalid("gdr_tree_mask: need R >= 0, cur_len >= 1, V > 0");
    if (R == 0) return GDR_OK;
    if (!input_ids || !scores) return invalid("gdr_tree_mask: null pointer");
    if (ids_row_stride < cur_len || scores_row_stride < V) return invalid("gdr_tree_mask: row stride smaller than row");
    GDR_CUDA(launch_tree_mask(t->first_child, t->child_tok, t->child_node, input_ids, ids_row_stride, R, cur_len, scores,
                              scores_row_stride, V, eos_id, strict, (cudaStream_t)stream));
    return GDR_OK;
}

int gdr_beam_step(gdr_trie_t *t, const float *logits, int64_t logits_row_stride, const int64_t *input_ids,
                  int64_t ids_row_stride, const float *beam_scores, int32_t B, int32_t K, int32_t cur_len, int32_t V,
                  int32_t eos_id, float *out_scores, int32_t *out_tokens, void *stream) {
    if (!t) return invalid("gdr_beam_step: trie is null");
    if (B < 0 || K <= 0 || cur_len < 1 || V <= 0) return invalid("gdr_beam_step: need B >= 0, K > 0, cur_len >= 1, V > 0");
    if (B == 0) return GDR_OK;
    if (!logits || !input_ids || !beam_scores || !out_scores || !out_tokens) return invalid("gdr_beam_step: null pointer");
    if (logits_row_stride < V || ids_row_stride < cur_len) return invalid("gdr_beam_step: row stride smaller than row");
    if ((int64_t)K * V > INT_MAX) return invalid("gdr_beam_step: K * V must fit in int32");
    const int fanout = (t->fanout + 3) / 4 * 4;
    const int64_t R = (int64_t)B * K;
    if ((int64_t)K * fanout > 32768 || 2 * K > 4096) {
        set_error("gdr_beam_step: K * max_fanout > 32768 or 2K > 4096 is not supported (use gdr_tree_mask + torch)");
        return GDR_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t need = (size_t)R * fanout * 8;
    if (need > t->cand_ws_bytes) {
        GDR_CUDA(cudaStreamSynchronize(st));
        if (t->cand_ws) GDR_CUDA(cudaFree(t->cand_ws));
        t->cand_ws = nullptr; t->cand_ws_bytes = 0;
        GDR_CUDA(cudaMalloc(&t->cand_ws, need));
        t->cand_ws_bytes = need;
    }
    float *cand_val = reinterpret_cast<float *>(t->cand_ws);
    int32_t *cand_id = reinterpret_cast<int32_t *>(cand_val + R * fanout);
    GDR_CUDA(launch_beam_rows(t->first_child, t->child_tok, t->child_node, input_ids, ids_row_stride, cur_len, logits,
                              logits_row_stride, V, beam_scores, (int)R, K, eos_id, fanout, cand_val, cand_id, st));
    // per query: top-2K of its K * fanout candidates, (score desc, flat index asc) — the merge kernel with one "rank"
    GDR_CUDA(launch_merge_topk(cand_val, cand_id, 1, B, K * fanout, (int64_t)B * K * fanout, 2 * K, out_scores, out_tokens, st));
    return GDR_OK;
}

int gdr_position_mask(float *logits, int64_t bz, int32_t sl, int32_t V, int32_t v_out, int32_t last_eos_only,
                      void *stream) {
    if (bz < 0 || sl <= 0 || V <= 0 || v_out <= 0) return invalid("gdr_position_mask: bad sizes");
    if (bz == 0) return GDR_OK;
    if (!logits) return invalid("gdr_position_mask: null pointer");
    // the reference scatters into index (sl-1)*v_out + v_out + 1, which must exist (modeling_t5.py:1567)
    const int64_t last_t = last_eos_only ? sl - 2 : sl - 1;   // last position that keeps its digit range
    if (last_t >= 0 && last_t * v_out + v_out + 1 >= V)
        return invalid("gdr_position_mask: vocabulary too small for sl positions of v_out tokens");
    GDR_CUDA(launch_position_mask(logits, bz, sl, V, v_out, last_eos_only, (cudaStream_t)stream));
    return GDR_OK;
}

}  // extern "C"
