// Shared declarations of the gdr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <mutex>
#include <string>

#include "../../include/gdr_b200.h"

namespace gdr {

// ----- error plumbing ---------------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define GDR_CUDA(call)                                        \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return gdr::cuda_fail(_e, #call); \
    } while (0)

// ----- work items ---------------------------------------------------------------------------
// One unit of scoring work: `nrows` consecutive store rows starting at `row0` (all inside one
// cluster, `rel0` rows after the cluster start) against `nq` (query, beam) pairs taken from
// grp_pair[slot0 .. slot0+nq).  Produced on the device by the inversion kernels.
struct __align__(16) Item {
    int32_t row0;
    int32_t rel0;
    int32_t nrows_nq;  // nrows | (nq << 16)
    int32_t slot0;
};

constexpr int SIMT_ROWS = 128;   // rows per SIMT item (one warp walks them 4 at a time)
constexpr int SIMT_QT = 4;       // (query, beam) pairs per SIMT item
constexpr int UMMA_ROWS = 128;   // rows per tcgen05 tile (UMMA_M)
constexpr int UMMA_NQ = 32;      // max pairs per tcgen05 tile (UMMA_N)
constexpr int MAX_DIM = 1024;
constexpr int GDR_MAX_RANKS = 8;   // GPUs of one NVSwitch domain a sharded corpus may span
constexpr int OFF_OWNER_SHIFT = 28; // sharded score offsets: owner rank in bits 28-30, offset inside the owner's score buffer below

// counters[] layout (device int32)
// CTR_TILE_NEXT / CTR_TILE_DONE: the tcgen05 kernel's dynamic tile queue (claimed with atomicAdd; the last CTA to finish
// resets both, so they are zero between launches)
// CTR_TOPK_NEXT / CTR_TOPK_DONE: the same protocol for the query queue of the grouped top-k (experiment, GDR_TOPK_GROUPS)
enum { CTR_N_SIMT = 0, CTR_N_UMMA = 1, CTR_N_TOUCHED = 2, CTR_TILE_NEXT = 3, CTR_TILE_DONE = 4, CTR_TOPK_NEXT = 5, CTR_TOPK_DONE = 6,
       CTR_SIMT_DONE = 7, CTR_COUNT = 8 };

// Everything the tcgen05 kernel needs to know about one tile, resolved once per batch by k_tilemeta (work item ->
// pairs -> candidate offsets: three dependent loads) so that the scoring kernel fetches it with ONE bulk copy.
struct __align__(16) TileMeta {
    int32_t row0, nrows, nq, rel0;   // first store row, rows (<= 128), pairs (<= 32; -1 = no more tiles), row offset inside the cluster
    int32_t qrow[32];                // query row of each pair (B operand)
    int32_t off[32];                 // score-buffer index of each pair's first row of this tile (epilogue)
};

// Everything one gdr_score_topk call needs on the device.
struct ScoreArgs {
    // store
    const void *emb;
    const int32_t *offsets;
    const int32_t *docid;
    int32_t dim, dtype, n_clusters, max_cluster;
    int64_t n_docs;
    // batch
    const float *q;
    const int32_t *beams;
    const float *prob;
    int32_t B, K, act, k;
    uint32_t flags;
    // scratch
    int32_t *cnt;        // [C]   zero between calls (self-restoring)
    int32_t *grp_off;    // [C+1] exclusive scan of cnt
    int32_t *simt_off;   // [C+1]
    int32_t *umma_off;   // [C+1]
    int32_t *grp_pair;   // [B*K] pair ids grouped by cluster
    int32_t *candoff;    // [B, K+1] start of each beam's segment in the query's candidate list
    int32_t *cbase;      // [B, K] first store row of each beam's cluster (0 for an absent beam)
    Item *simt_items;
    Item *umma_items;
    int32_t *counters;   // [CTR_COUNT]
    int32_t *scan_base;  // [4 * (n_scan_blocks + 1)] exclusive bases (group, SIMT items, tcgen05 items, -) of each 8,192-cluster block
    float *scorebuf;     // [B, stride]
    int64_t stride;
    __nv_bfloat16 *qsplit;  // [rows(q), 3, dim] exact bf16 hi/mid/lo split of the fp32 queries (tcgen05 path), else null
    uint32_t *gkeys;     // [B, stride] keys scratch for the global-memory top-k variant
    TileMeta *tile_meta; // [umma tile capacity] (null unless the tcgen05 path is taken)
    uint32_t *ghist;     // [B, 2048] histogram scratch of the small-footprint top-k's fallback (null: variant not used)
    long long *dbg;      // [512] optional timeline scratch (GDR_UMMA_TRACE=1), else null
    int32_t umma_min_group;  // groups with at least this many pairs go to the tcgen05 path (INT_MAX = never)
    // ---- cluster-sharded corpus (all defaults = one unsharded store: c_lo = 0, c_hi = n_clusters, row_lo = 0, n_ranks = 1)
    // offsets / docid / n_clusters describe the GLOBAL corpus; emb holds only the rows of the owned clusters [c_lo, c_hi), i.e.
    // global rows [row_lo, ...).  Every rank inverts the whole global batch, scores the pairs that land in its clusters and
    // stores each score straight into the score buffer of the query's OWNER (rank b / b_own) over NVLink (peer_score[]); the
    // owner runs the top-k of its own b_own queries once all ranks have signalled (sig_*).
    int32_t c_lo, c_hi, row_lo;
    int32_t n_ranks, my_rank, b_own;   // queries per owner (= B when n_ranks == 1)
    int32_t q_base, B_top;             // the top-k of this rank covers global queries [q_base, q_base + B_top) (0, B when unsharded)
    float *peer_score[GDR_MAX_RANKS];  // score buffer [b_own, stride] of each rank (this rank's own entry = scorebuf)
    int32_t *peer_sig[GDR_MAX_RANKS];  // each rank's arrival flags [n_ranks]: peer_sig[r][my_rank] <- this handle's scoring epoch
    int32_t *sig_local;                // this rank's arrival flags [n_ranks] (= peer_sig[my_rank]); the CONSUMED flags follow GDR_MAX_RANKS ints later:
                                       // peer_sig[r][GDR_MAX_RANKS + my_rank] <- "this rank's top-k of epoch e has read its score buffer"
    int32_t *sig_epoch;                // [1] device counter: scoring launches of this handle so far
    int32_t signal;                    // 1 in the copy given to the call's LAST scoring kernel: it signals the owners when it is done
    int32_t wait_in_topk;              // 1: the top-k kernel itself waits for the arrival flags (fused launches); 0: a one-warp k_wait_scorers
                                       // launched in front of it did (stand-alone top-k: a grid of spinning CTAs could keep a scoring kernel off the SMs)
    int32_t launch_prio;     // host side only: scheduling priority class of the launches made with this copy of the arguments
                             // (0 = none; else a CUDA priority + 1000), set per phase by the C-ABI entry point — per call, not global
};

// ----- launchers (each enqueues on `s`, returns cudaGetLastError()) ---------------------------
cudaError_t launch_invert(const ScoreArgs &a, cudaStream_t s, int *n_launches);
cudaError_t launch_score_simt(const ScoreArgs &a, cudaStream_t s, int sm_count);
cudaError_t launch_score_umma(const ScoreArgs &a, const CUtensorMap *tmap, cudaStream_t s, int sm_count);
cudaError_t launch_score_umma_x2(const ScoreArgs &a, const CUtensorMap *tmap, cudaStream_t s, int ctas);   // 4-stage ring, two CTAs per SM
cudaError_t launch_score_tile_f32(const ScoreArgs &a, cudaStream_t s, int sm_count);
cudaError_t launch_topk_store(const ScoreArgs &a, float alpha, float *out_scores, int32_t *out_docids,
                              cudaStream_t s);
cudaError_t launch_wait_scorers(const ScoreArgs &a, cudaStream_t s);
cudaError_t launch_wait_consumed(const ScoreArgs &a, cudaStream_t s);
cudaError_t launch_signal_consumed(const ScoreArgs &a, cudaStream_t s);
cudaError_t launch_topk_grouped(const ScoreArgs &a, float alpha, float *out_scores, int32_t *out_docids, cudaStream_t s, int groups);
cudaError_t launch_score_fused(const ScoreArgs &a, const CUtensorMap *tmap, const ScoreArgs &prev, float alpha, float *out_scores,
                               int32_t *out_docids, cudaStream_t s, int ctas, int groups);
int fused64_groups_that_fit(int K);
cudaError_t launch_merge_topk(const float *scores, const int32_t *docids, int G, int B, int k_in, int64_t g_stride, int k,
                              float *out_scores, int32_t *out_docids, cudaStream_t s);
cudaError_t launch_similarity(const float *q, int64_t Q, const void *p, int64_t P, int dim, int p_dtype,
                              float *out, cudaStream_t s, int sm_count);
// tensor-core form for bf16 passages with dim % 64 == 0 (similarity.cu): returns cudaErrorNotSupported when it does not apply
cudaError_t launch_similarity_umma(const float *q, int64_t Q, const void *p, int64_t P, int dim, float *out, cudaStream_t s, int sm_count);
bool umma_make_tensor_map(CUtensorMap *out, const void *emb, int64_t n_docs, int dim);
cudaError_t launch_centroids(const void *emb, int dtype, const int32_t *offsets, int n_clusters, int dim, float *out, cudaStream_t s);
cudaError_t launch_tree_mask(const int32_t *first_child, const int32_t *child_tok, const int32_t *child_node,
                             const int64_t *input_ids, int64_t ids_stride, int R, int cur_len, float *scores,
                             int64_t scores_stride, int V, int eos_id, int strict, cudaStream_t s);
cudaError_t launch_beam_rows(const int32_t *first_child, const int32_t *child_tok, const int32_t *child_node,
                             const int64_t *input_ids, int64_t ids_stride, int cur_len, const float *logits,
                             int64_t logits_stride, int V, const float *beam_scores, int R, int K, int eos_id, int fanout,
                             float *cand_val, int32_t *cand_id, cudaStream_t s);
cudaError_t launch_node_embeddings(const int32_t *first_child, const int32_t *child_node, const int32_t *child_order, const int32_t *node_cluster,
                                   const float *leaf_emb, const int32_t *leaf_num, int dim, const int *level_start, int n_levels,
                                   float *node_emb, int32_t *node_leaf_num, cudaStream_t s);
cudaError_t launch_tree_match(const int32_t *first_child, const int32_t *child_tok, const int32_t *child_node, const int32_t *child_order,
                              const float *node_emb, const int32_t *node_leaf_num, int dim, const float *docs, int M, int max_len,
                              int32_t *out_tokens, int32_t *out_len, cudaStream_t s);
cudaError_t launch_contrastive(const void *emb, int dtype, int dim, const float *q, const int32_t *pos_rows, const int32_t *cand_rows,
                               const int32_t *cand_off, int B, int S, int act, float tau, float intra_rate, float *loss_per_query,
                               float *loss, float *grad_q, cudaStream_t st);
cudaError_t launch_position_mask(float *logits, int64_t bz, int sl, int V, int v_out, int last_eos_only,
                                 cudaStream_t s);

// ----- programmatic dependent launch ---------------------------------------------------------
// Every kernel of the chain (count -> scan -> fill -> score -> top-k) is launched with the PDL attribute: its
// CTAs may be scheduled, and run their prologue (shared-memory carve-up, barrier init, TMEM allocation), while
// the previous kernel is still draining.  pdl_wait() blocks until the previous kernel has completed and its
// writes are visible; it must precede the first access to anything an earlier kernel produced or still reads.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// `priority` is the per-launch scheduling class (gdr_store_set_option GDR_OPT_LAUNCH_PRIORITIES: inversion > scoring > top-k):
// 0 = no attribute, else a CUDA stream-priority number offset by +1000 (priorities can be 0 or negative).  It is an
// argument, not process state: handles driven from different host threads cannot see each other's value.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int priority, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (priority != 0) {
        attr[1].id = cudaLaunchAttributePriority;
        attr[1].val.priority = priority - 1000;
        cfg.numAttrs = 2;
    }
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute is per device and function.  One instance per launcher (function-local static): `ensure` runs the
// setter once per device under a mutex and records the device only after the setter succeeded, so a second host thread can
// neither skip the attribute nor launch before it is in place.
struct FuncAttrOnce {
    std::mutex mu;
    unsigned long long done = 0;      // one bit per device ordinal (mod 64)
    template <typename F>
    cudaError_t ensure(F setter) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lock(mu);
        if ((done >> (dev & 63)) & 1ull) return cudaSuccess;
        e = setter();
        if (e == cudaSuccess) done |= 1ull << (dev & 63);
        return e;
    }
};

// ----- optional kernel timeline (GDR_UMMA_TRACE): slots 500..511 of a.dbg hold min start / max end per kernel class
__device__ __forceinline__ unsigned long long gdr_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_start(long long *dbg, int slot) {
    if (dbg && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long *>(dbg) + 500 + slot, gdr_gtime());
}
__device__ __forceinline__ void trace_end(long long *dbg, int slot) {
    if (dbg && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(dbg) + 500 + slot, gdr_gtime());
}

// ----- sharded corpus: where a score goes, when the owner may read it ---------------------------------------------------
// `off` = index inside a score buffer; with n_ranks > 1 the owner rank sits in bits 28-30 (k_tilemeta / score_dst pack it).
__device__ __forceinline__ int64_t pack_score_off(const ScoreArgs &a, int b, int64_t within_row) {
    if (a.n_ranks <= 1) return (int64_t)b * a.stride + within_row;
    return ((int64_t)(b / a.b_own) << OFF_OWNER_SHIFT) | ((int64_t)(b % a.b_own) * a.stride + within_row);
}
__device__ __forceinline__ float *score_ptr(const ScoreArgs &a, int64_t off) {
    if (a.n_ranks <= 1) return a.scorebuf + off;
    return a.peer_score[off >> OFF_OWNER_SHIFT] + (off & ((1ll << OFF_OWNER_SHIFT) - 1));
}
// Called by ONE thread after the last CTA of a scoring kernel has finished (every CTA fenced at system scope before it
// counted itself done): bump this handle's scoring epoch and publish it to every owner.
__device__ __forceinline__ void signal_owners(const ScoreArgs &a) {
    if (a.n_ranks <= 1 || !a.signal) return;
    const int e = *a.sig_epoch + 1;
    *a.sig_epoch = e;
    __threadfence_system();
    for (int r = 0; r < a.n_ranks; ++r)
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(a.peer_sig[r] + a.my_rank), "r"(e) : "memory");
}
// Called by ONE thread of a top-k CTA / group before the first score is read: every rank's scoring launch of this handle's
// current epoch has landed in this rank's score buffer.  The local epoch is the expected value: the local scoring launch of
// the same batch precedes the top-k in stream order, and all ranks step through the handle's batches in the same order.
__device__ __forceinline__ void wait_for_scorers(const ScoreArgs &a, bool in_topk = true) {
    if (a.n_ranks <= 1 || (in_topk && !a.wait_in_topk)) return;
    const int expected = *reinterpret_cast<volatile int32_t *>(a.sig_epoch);
    for (int r = 0; r < a.n_ranks; ++r) {
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(a.sig_local + r) : "memory");
        } while (v - expected < 0);
    }
}

// The reverse handshake: "this rank's top-k has read epoch e of this handle's score buffer" (flags GDR_MAX_RANKS ints behind the arrival
// flags).  A scoring launch into a handle waits for every owner's flag of the handle's previous epoch before it overwrites their buffers.
__device__ __forceinline__ void signal_consumed(const ScoreArgs &a) {
    if (a.n_ranks <= 1) return;
    const int e = *reinterpret_cast<volatile int32_t *>(a.sig_epoch);
    __threadfence_system();
    for (int r = 0; r < a.n_ranks; ++r)
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(a.peer_sig[r] + GDR_MAX_RANKS + a.my_rank), "r"(e) : "memory");
}
__device__ __forceinline__ void wait_consumed(const ScoreArgs &a) {
    if (a.n_ranks <= 1) return;
    const int expected = *reinterpret_cast<volatile int32_t *>(a.sig_epoch);      // the epoch of this handle's PREVIOUS scoring launch
    for (int r = 0; r < a.n_ranks; ++r) {
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(a.sig_local + GDR_MAX_RANKS + r) : "memory");
        } while (v - expected < 0);
    }
}

// ----- small device helpers -----------------------------------------------------------------
// fp32 -> (hi, mid, lo) bf16 with x == hi + mid + lo exactly (8 + 8 + 8 mantissa bits)
__device__ __forceinline__ void split3(float x, __nv_bfloat16 &hi, __nv_bfloat16 &mid, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);      // negative: ~u, else u | sign bit — two instructions
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == GDR_ACT_TANH) return tanhf(x);
    if (act == GDR_ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
    return x;
}

}  // namespace gdr
