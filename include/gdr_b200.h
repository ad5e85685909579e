/*
 * gdr_b200.h — C ABI of the B200-native fine-grained stage of GDR (ypw0102/GDR).
 *
 * The reference is pure Python on torch and has no FFI layer of its own;
 * the entry points below are what a binding for this path replaces, each citing the
 * reference code it stands in for (paths relative to the reference root).  All pointers
 * marked DEV are device pointers on the current CUDA device; HOST pointers are host
 * memory.  The caller owns every buffer it passes in; the library owns only its handles
 * and their scratch.  Every function returns GDR_OK or a negative status, never throws or
 * aborts; gdr_last_error() gives the message for the calling thread.  Kernels are enqueued
 * on the `stream` argument (a cudaStream_t passed as void*; NULL = legacy default stream);
 * no entry point synchronises the device unless documented.  There is NO CPU fallback.
 *
 * Concurrency: a store / trie handle owns ONE set of scratch buffers, so calls on the same handle must be ordered
 * (same stream, or streams ordered with events).  To keep several batches in flight, create one handle per batch in
 * flight over the SAME device arrays (a handle copies nothing: gdr_b200/pipeline.py runs three per GPU).  The scratch grows on
 * the first call of a larger shape (that call synchronises the stream) unless gdr_store_reserve sized it beforehand.
 * Handles may be driven from different host threads: there is no mutable process-wide state (launch priorities travel in the
 * call's own arguments, the one-time per-device kernel attributes are set under a mutex), gdr_last_error is per thread.
 */
#ifndef GDR_B200_H
#define GDR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDR_OK 0
#define GDR_ERR_INVALID (-1)     /* bad argument (shape, dtype, null pointer)            */
#define GDR_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed                    */
#define GDR_ERR_UNSUPPORTED (-3) /* valid request this build cannot serve (e.g. D > 1024) */
#define GDR_ERR_NOMEM (-4)

#define GDR_DTYPE_F32 0
#define GDR_DTYPE_BF16 1

#define GDR_ACT_NONE 0    /* dense.py:53-54 (plain q·d)                                  */
#define GDR_ACT_TANH 1    /* main_models.py:1580-1582, --loss_func tanh (main.py:393)     */
#define GDR_ACT_SIGMOID 2 /* main_models.py:1578-1579, --loss_func sigmoid                */

/* flags for gdr_score_topk */
#define GDR_Q_PER_BEAM 1u       /* q is [B*K, D]: one query vector per (query, beam), main_models.py:1467-1571,1583-1594 */
#define GDR_FORCE_SIMT 2u       /* never use the tcgen05 grouped-GEMM path                */
#define GDR_FORCE_UMMA 4u       /* use the tcgen05 path for every non-empty group         */
/* Phase selection: a call normally runs inversion -> scoring -> top-k on `stream`.  A caller that pipelines batches
 * may issue the three phases of one batch as three calls (same arguments) on three streams, ordered with events, so
 * that the scoring kernels of consecutive batches run back to back while the inversion and top-k kernels of
 * neighbouring batches fill in around them.  (Measured at cfg2: no better than simply running whole batches on
 * three streams, DESIGN.md §7 — the flags are kept for callers that want to schedule the phases themselves.) */
#define GDR_SKIP_INVERT 256u
#define GDR_SKIP_SCORE 512u
#define GDR_SKIP_TOPK 1024u

typedef struct gdr_store gdr_store_t;
typedef struct gdr_trie gdr_trie_t;

int gdr_abi_version(void);
const char *gdr_last_error(void);

/* ---- document-embedding store ------------------------------------------------------------
 * Replaces `self.doc_embed` (int-indexable list of [D] tensors, main_models.py:806-814) and
 * `self.id_mapping` (cluster id -> list of doc indices, main_models.py:874-889) with a
 * cluster-contiguous CSR table resident in HBM:
 *   emb     DEV [n_docs, dim] row-major, GDR_DTYPE_F32 or GDR_DTYPE_BF16, 16-byte aligned,
 *           rows permuted so that cluster c is rows offsets[c] .. offsets[c+1]-1
 *   offsets DEV int32 [n_clusters + 1]
 *   docid   DEV int32 [n_docs]: the reference's global doc index of each row
 *   max_cluster_size  max_c (offsets[c+1]-offsets[c]), computed by the caller while building the CSR
 * The store keeps the pointers (no copy); they must outlive it.  dim % 8 == 0, dim <= 1024. */
int gdr_store_create(gdr_store_t **out, const void *emb, int64_t n_docs, int32_t dim, int32_t dtype,
                     const int32_t *offsets, int32_t n_clusters, const int32_t *docid,
                     int32_t max_cluster_size);
int gdr_store_destroy(gdr_store_t *store);

/* ---- cluster-sharded corpus (new design; the reference keeps a full replica per rank, main_models.py:806-814) ----
 * One handle per GPU over a corpus whose clusters are split across GPUs by contiguous ranges of a GLOBAL cluster numbering:
 *   emb_local       DEV [n_local_rows, dim]: the rows of clusters c_lo .. c_hi-1 only (global rows row_lo .. row_lo+n_local_rows-1)
 *   offsets_global  DEV int32 [n_clusters_global + 1], docid_global DEV int32 [n_docs_global]: the WHOLE corpus' CSR metadata,
 *                   replicated on every GPU (4 bytes per document)
 * gdr_score_topk on such a handle takes beams with GLOBAL cluster ids and scores only the beams that land in [c_lo, c_hi).
 * Alone it is useful only together with the peer-to-peer exchange below (candidate segments of clusters other ranks own
 * are not written by this rank). */
int gdr_store_create_shard(gdr_store_t **out, const void *emb_local, int64_t n_local_rows, int32_t dim, int32_t dtype,
                           const int32_t *offsets_global, int32_t n_clusters_global, const int32_t *docid_global,
                           int64_t n_docs_global, int32_t max_cluster_size, int32_t c_lo, int32_t c_hi, int64_t row_lo);

/* Peer-to-peer candidate exchange between the n_ranks handles (one per GPU of an NVLink / NVSwitch domain) that hold the shards
 * of one corpus.  The global batch is n_ranks * b_own queries, the same q / beams on every rank; rank r OWNS queries
 * r*b_own .. (r+1)*b_own-1.  Every rank inverts the whole batch and scores the beams of its clusters; its scoring kernel
 * stores each score straight into the OWNER's score buffer over NVLink (the exchange is fused into the scoring epilogue — no
 * collective, no candidate lists), then publishes an arrival flag to every owner; the owner's top-k (stand-alone, or the
 * top-k groups of the next fused launch) waits for the n_ranks flags and selects exactly as on one GPU.  On such a handle
 * gdr_score_topk / gdr_score_fused take B = n_ranks * b_own and write outputs [n_alpha, b_own, k] for the rank's own queries.
 * All ranks must issue the same sequence of batches on corresponding handles.
 *   gdr_store_p2p_init    allocates this handle's buffer (b_own * align4(K * max_cluster_size) scores + flags) and returns its
 *                         CUDA IPC handle (HOST, GDR_IPC_HANDLE_BYTES bytes; may be NULL for same-process use)
 *   gdr_store_p2p_attach  all_handles HOST [n_ranks][GDR_IPC_HANDLE_BYTES], rank order (one process per GPU: exchange them with
 *                         any host-side all-gather, e.g. torch.distributed); maps every peer's buffer
 *   gdr_store_p2p_attach_local  the same for handles that live in ONE process (peers[r] = rank r's handle; tests, single-process multi-GPU) */
#define GDR_IPC_HANDLE_BYTES 64
int gdr_store_p2p_init(gdr_store_t *store, int32_t n_ranks, int32_t my_rank, int32_t b_own, int32_t K, void *ipc_handle_out);
int gdr_store_p2p_attach(gdr_store_t *store, const void *all_handles);
int gdr_store_p2p_attach_local(gdr_store_t *store, gdr_store_t *const *peers);

/* Peer-to-peer all-gather of a batch's inputs (csrc/xchg.cu): in the sharded fine stage every rank needs the queries and beams of
 * the whole global batch every step; the transfer is done by the copy engines into buffers mapped with CUDA IPC and one warp
 * handles the arrival flags — no collective kernel competes with the persistent scoring CTAs for SMs.
 *   gdr_xchg_bytes        size of the buffer the caller must provide (DEV, 256-byte aligned, e.g. a torch uint8 tensor): n_slots
 *                         slot regions (part after part, each part [n_ranks x part_bytes[p]] in rank order), flags, epochs
 *   gdr_xchg_create       blob_out HOST [GDR_XCHG_BLOB_BYTES]: the buffer's CUDA IPC handle + its offset inside the allocation
 *   gdr_xchg_attach       all_blobs HOST [n_ranks][GDR_XCHG_BLOB_BYTES] in rank order (exchanged by any host-side all-gather);
 *                         gdr_xchg_attach_local: the same for objects living in one process
 *   gdr_xchg_all_gather   own DEV = this rank's parts back to back; copies them into slot `slot` of EVERY rank's buffer, raises this
 *                         rank's arrival flag on every rank and waits (one warp, stream-ordered) for all ranks' flags: work enqueued
 *                         on `stream` afterwards sees the complete slot.  All ranks must call it for the same slots in the same order.
 *   gdr_xchg_part_offset  byte offset of part `part` of slot `slot` inside the buffer */
#define GDR_XCHG_BLOB_BYTES 72
typedef struct gdr_xchg gdr_xchg_t;
int64_t gdr_xchg_bytes(int32_t n_ranks, int32_t n_slots, const int64_t *part_bytes, int32_t n_parts);
int gdr_xchg_create(gdr_xchg_t **out, void *buffer, int32_t n_ranks, int32_t my_rank, int32_t n_slots, const int64_t *part_bytes,
                    int32_t n_parts, void *blob_out);
int gdr_xchg_attach(gdr_xchg_t *x, const void *all_blobs);
int gdr_xchg_attach_local(gdr_xchg_t *x, gdr_xchg_t *const *peers);
int gdr_xchg_all_gather(gdr_xchg_t *x, int32_t slot, const void *own, void *stream);
int64_t gdr_xchg_part_offset(gdr_xchg_t *x, int32_t slot, int32_t part);
int gdr_xchg_destroy(gdr_xchg_t *x);

/* ---- fine stage: cluster-restricted scoring + top-k ---------------------------------------
 * Replaces main_models.py:1441-1462 (gather), :1577-1594 (score), :1596-1624 (rerank bias),
 * :1625 (topk) and :1628-1631 (index -> doc index) for one batch; with act = NONE, prob = NULL
 * it is dense.py:53-54 `compute_similarity` restricted to the beam clusters + `Tensor.topk`.
 *   q        DEV fp32 [B, dim]  (or [B*K, dim] with GDR_Q_PER_BEAM)
 *   beams    DEV int32 [B, K]   cluster index of each beam, -1 = absent
 *   prob     DEV fp32 [B, K] or NULL: softmax of the beam scores (main_models.py:1601)
 *   alphas   HOST fp32 [n_alpha] or NULL (then n_alpha must be 1 and alpha = 1): --score_rate
 *            (main.py:389); result r uses score + alphas[r] * prob[b, beam_of(candidate)]
 *   out_scores DEV fp32 [n_alpha, B, k], out_docids DEV int32 [n_alpha, B, k], sorted by score
 *            descending, ties by ascending docid; queries with fewer than k candidates are padded
 *            with (-inf, -1) (the reference raises there, main_models.py:1625; the Python shim
 *            restores that behaviour). */
int gdr_score_topk(gdr_store_t *store, const float *q, const int32_t *beams, const float *prob,
                   const float *alphas, int32_t n_alpha, int32_t B, int32_t K, int32_t act, int32_t k,
                   uint32_t flags, float *out_scores, int32_t *out_docids, void *stream);

/* Pipelined schedule — scoring of batch i and the top-k of batch i-1 in ONE launch (csrc/score_fused.cu; gdr_b200/pipeline.py
 * drives it).  No reference counterpart: it is a different schedule of main_models.py:1577-1631, not a different result
 * (tests/test_gpu_pipeline.py: bit-identical to gdr_score_topk).  Why: as separate grids the top-k CTAs of batch i-1 and the
 * scoring CTAs of batch i compete for SM residency; fused, the top-k has a fixed home in the scoring CTA.
 * `cur` and `prev` are two handles over the same (or different) embeddings, i.e. two scratch sets.  Protocol per batch i:
 *   gdr_score_topk(h[i % 3], q_i, beams_i, prob_i, NULL, 1, B, K, act, k, GDR_SKIP_SCORE | GDR_SKIP_TOPK, any valid out pointers, stream_inv)
 *                                             -- the inversion of batch i into h[i % 3]'s scratch (the outputs are checked, not written);
 *                                                on a second stream it runs one batch ahead of the fused launches
 *   gdr_score_fused(h[i % 3], i ? h[(i - 1) % 3] : NULL, alpha, out_scores_{i-1}, out_docids_{i-1}, stream)
 *                                             -- ONE launch: scores batch i and selects the top-k of batch i-1 in the same CTAs
 *   ... and after the last batch n-1:  gdr_score_fused(NULL, h[(n - 1) % 3], alpha, out_scores_{n-1}, out_docids_{n-1}, stream)
 * Requirements: the batch in `cur` takes the tcgen05 path alone (bf16 store, dim % 64 == 0, B*K >= 3 * n_clusters or
 * GDR_FORCE_UMMA); the batch in `prev` has k <= 128 and <= 65,535 candidates per query; prev's q / beams / prob buffers are
 * still alive; one alpha per call (result = score + alpha * prob[b, beam]); outputs DEV [B, k] of the batch in `prev`.
 * GDR_ERR_UNSUPPORTED otherwise (the caller then runs gdr_score_topk batch by batch). */
int gdr_score_fused(gdr_store_t *cur, gdr_store_t *prev, float alpha, float *prev_out_scores, int32_t *prev_out_docids,
                    void *stream);

/* Pre-size a handle's scratch for batches of up to (B, K, k) with these flags, so that no allocation (and no stream
 * synchronisation) happens on the query path and the first call of a shape can already be captured in a CUDA graph. */
int gdr_store_reserve(gdr_store_t *store, int32_t B, int32_t K, int32_t k, uint32_t flags, void *stream);

/* Launch options of a handle.  None of them changes results (tests/test_gpu_pipeline.py). */
#define GDR_OPT_UMMA_CTAS 1          /* persistent CTAs of the tcgen05 kernels; 0 = default (one per SM; fused launches: SMs - 8) */
#define GDR_OPT_UMMA_MIN_GROUP 2     /* > 1: mixed mode, groups of at least this many pairs on tensor cores, the rest on the GEMV */
#define GDR_OPT_LAUNCH_PRIORITIES 3  /* 1: per-launch scheduling priorities, inversion > scoring > top-k */
#define GDR_OPT_FUSED_GROUPS 4       /* top-k groups in the fused CTA: 5 = five 128-thread groups, lean select (default); 9 = nine 64-thread groups */
#define GDR_OPT_TOPK_GROUPS 5        /* 1, 2, 4: stand-alone top-k as persistent groups walking a query queue; 0 = one CTA per query */
#define GDR_OPT_TOPK_WIDE 6          /* 1: the 256-thread top-k also for k <= 128 */
#define GDR_OPT_UMMA_CTAS_PER_SM 7   /* 2: tcgen05 scoring kernel with a 4-stage ring, two CTAs per SM (default grid 2 x SMs; set GDR_OPT_UMMA_CTAS
                                      * to 2 x the SMs of an SM partition); for gdr_score_topk on unsharded handles, ignored elsewhere */
int gdr_store_set_option(gdr_store_t *store, int32_t option, int32_t value);

/* ---- SM partition for the pipelined schedule (csrc/partition.cu) -------------------------------------------------------------
 * Splits the SMs of the CURRENT device into two disjoint sets behind CUDA green contexts (driver 12.4+) and creates streams in
 * each: `small_sms` SMs (rounded up by the driver to its granularity, 8 on sm_90+) for the latency-bound inversion and top-k
 * kernels, the rest for the HBM-bound scoring kernel — so that the two sides stop competing for residency (the reference runs one
 * batch at a time on one stream, main_models.py:1434-1637, and has no counterpart).  The streams are ordinary CUDA streams of the
 * primary context's address space: pass them as the `stream` argument of gdr_score_topk with GDR_SKIP_* flags to issue a batch's
 * phases on either side (events order them; gdr_b200/pipeline.py, schedule "partitioned"), and set GDR_OPT_UMMA_CTAS of the
 * handles to the big side's SM count.  Never changes results.  GDR_ERR_UNSUPPORTED when the driver has no green contexts. */
typedef struct gdr_partition gdr_partition_t;
int gdr_partition_create(gdr_partition_t **out, int32_t small_sms, int32_t n_streams_big, int32_t n_streams_small);
int gdr_partition_sms(const gdr_partition_t *partition, int32_t out[2]);               /* out[0] = SMs of the big side, out[1] = small side */
void *gdr_partition_stream(gdr_partition_t *partition, int32_t small, int32_t index);  /* cudaStream_t, NULL when out of range */
int gdr_partition_destroy(gdr_partition_t *partition);                                 /* after all work on its streams has completed */

/* Counters of the most recent gdr_score_topk on this store (device-side work-list sizes):
 * out[0] = (cluster, query-chunk) items scored by the SIMT GEMV path, out[1] = tiles scored by the
 * tcgen05 grouped-GEMM path, out[2] = kernels launched for that batch (summed over the calls that issued its phases), out[3] = clusters touched.
 * Synchronises `stream`. */
int gdr_store_last_stats(gdr_store_t *store, int64_t out[4], void *stream);

/* Measurement aid (bench.py): with profiling enabled, gdr_score_topk records CUDA events on the
 * caller's stream between its phases; gdr_store_last_phase_ms synchronises and returns the device
 * time of the most recent profiled call: out[0] = pair inversion, out[1] = tcgen05 scoring (incl. the
 * query split), out[2] = SIMT scoring, out[3] = top-k (all alphas). */
int gdr_store_set_profiling(gdr_store_t *store, int32_t enable);
int gdr_store_last_phase_ms(gdr_store_t *store, float out[4]);

/* ---- index expansion --------------------------------------------------------
 * Leaf-cluster centroids of tree_embedding_calculate (main_models.py:154-158): out DEV fp32 [n_clusters, dim],
 * out[c] = mean of cluster c's rows (rows summed in store order, fp32); empty clusters give zeros.  Assigning new
 * documents to clusters (tree_embedding_insert, main_models.py:268-295: argmax_c doc . centroid_c) is then
 * gdr_score_topk with k = 1 on a store whose single cluster holds the centroids (gdr_b200/expand.py). */
int gdr_cluster_centroids(gdr_store_t *store, float *out, void *stream);

/* ---- training-time gather + contrastive loss -------------------------------------------
 * Replaces the row gather of T5FineTuner.forward (main_models.py:983-996) and encoder_cal (main_models.py:1184-1221):
 *   q DEV fp32 [B, dim]; pos_rows DEV int32 [B] = store row of each query's positive document; cand_rows DEV int32 [S] =
 *   store rows of the in-cluster candidates, query 0's first; cand_off DEV int32 [B+1] = start of each query's own
 *   candidates in cand_rows (the reference's valid_num, prefix-summed).  act = GDR_ACT_TANH | GDR_ACT_SIGMOID (--loss_func),
 *   tau (--tau), intra_rate (--intra_rate).
 *   -> loss_per_query DEV fp32 [B]; loss DEV fp32 [1] = their mean (= encoder_cal's return value); grad_q DEV fp32
 *   [B, dim] = d loss / d q, or NULL.  Document embeddings get no gradient (they are a fixed table in the reference). */
int gdr_contrastive_loss(gdr_store_t *store, const float *q, const int32_t *pos_rows, const int32_t *cand_rows,
                         const int32_t *cand_off, int32_t B, int32_t S, int32_t act, float tau, float intra_rate,
                         float *loss_per_query, float *loss, float *grad_q, void *stream);

/* ---- dense similarity (dense.py:53-54 / encoder.py:128-129): out[Q, P] = q @ p^T, fp32 out ----
 *   q DEV fp32 [Q, dim]; p DEV [P, dim] of p_dtype; out DEV fp32 [Q, P]. */
int gdr_similarity(const float *q, int64_t Q, const void *p, int64_t P, int32_t dim, int32_t p_dtype,
                   float *out, void *stream);

/* ---- merge of per-rank candidates (new in the sharded design) ----------------
 *   scores DEV fp32 [G, B, k_in], docids DEV int32 [G, B, k_in] (docid -1 = padding); rank g's
 *   block starts g * g_stride elements after the base pointer (g_stride = B * k_in when dense; a
 *   larger stride lets both arrays live interleaved in one all-gather buffer)
 *   -> out_scores DEV fp32 [B, k], out_docids DEV int32 [B, k], same ordering rule as above. */
int gdr_merge_topk(const float *scores, const int32_t *docids, int32_t G, int32_t B, int32_t k_in,
                   int64_t g_stride, int32_t k, float *out_scores, int32_t *out_docids, void *stream);

/* ---- prefix-tree docid mask ------------------------------------------------------------------
 * Device form of the `Node` trie (main_models.py:112-151).  HOST CSR arrays, copied by the
 * library: node n's children are edges first_child[n] .. first_child[n+1]-1, sorted by token;
 * node 0 is the root. */
int gdr_trie_create(gdr_trie_t **out, const int32_t *first_child, const int32_t *child_tok,
                    const int32_t *child_node, int32_t n_nodes, int32_t n_edges);
int gdr_trie_destroy(gdr_trie_t *trie);

/* ---- node embeddings and greedy descent (index expansion) ---------------------------------
 * gdr_trie_set_child_order: HOST order[n_edges] lists every node's edges (indices into child_tok / child_node, within
 *   first_child[n] .. first_child[n+1]-1) in the reference's child INSERTION order (dict order of Node.children): the
 *   order tree_embedding_calculate accumulates in and np.argmax breaks ties by.  Default: by token.
 * gdr_trie_node_embeddings replaces tree_embedding_calculate (main_models.py:154-179): node_cluster DEV int32 [n_nodes] =
 *   store cluster of a leaf-cluster node (-1 otherwise), leaf_emb DEV fp32 [C, dim] (gdr_cluster_centroids), leaf_num DEV
 *   int32 [C] (cluster sizes) -> node_emb DEV fp32 [n_nodes, dim], node_leaf_num DEV int32 [n_nodes] (0 = the node has no
 *   embedding, e.g. the EOS child of a leaf cluster).  Needs breadth-first node numbering.
 * gdr_tree_match replaces tree_match (main_models.py:232-252): docs DEV fp32 [M, dim] -> out_tokens DEV int32 [M, max_len]
 *   = [0, tok, ..., 1] and out_len DEV int32 [M]; at every node the child with the largest doc . embedding, first maximum
 *   in child order; stops at a node whose only child has no embedding. */
int gdr_trie_set_child_order(gdr_trie_t *trie, const int32_t *first_child, const int32_t *order);
int gdr_trie_node_embeddings(gdr_trie_t *trie, const int32_t *node_cluster, const float *leaf_emb, const int32_t *leaf_num,
                             int32_t dim, float *node_emb, int32_t *node_leaf_num, void *stream);
int gdr_tree_match(gdr_trie_t *trie, const float *node_emb, const int32_t *node_leaf_num, int32_t dim, const float *docs,
                   int32_t M, int32_t max_len, int32_t *out_tokens, int32_t *out_len, void *stream);

/* Replaces generation_utils_previous.py:714-729.  For each of R rows, walk the trie along
 * input_ids[r, 1:cur_len]; allowed = children of the node reached, or {eos_id} if the path
 * leaves the tree.  In place on scores DEV fp32 [R, V] (row stride in elements): allowed
 * entries become s + 0.0f, all others s + (-inf).  strict = 0 writes -inf without reading
 * the masked entries (identical unless a masked input is NaN or +inf); strict = 1 reads
 * every entry and is bit-identical for all inputs.  Child tokens >= V are ignored.
 *   input_ids DEV int64 [R, cur_len] (row stride in elements). */
int gdr_tree_mask(gdr_trie_t *trie, const int64_t *input_ids, int64_t ids_row_stride, int32_t R,
                  int32_t cur_len, float *scores, int64_t scores_row_stride, int32_t V, int32_t eos_id,
                  int32_t strict, void *stream);

/* Fused beam step ("next" row): replaces generation_utils_previous.py:694 (log_softmax), :714-729 (tree
 * mask) and :757-771 (add the beam scores, view [B, K*V], topk(2K, largest, sorted)) with ONE read of the logits.
 *   logits DEV fp32 [B*K, V] (row stride in elements, not modified), input_ids DEV int64 [B*K, cur_len],
 *   beam_scores DEV fp32 [B*K]
 *   -> out_scores DEV fp32 [B, 2K], out_tokens DEV int32 [B, 2K]: flat index beam*V + token, sorted by score
 *      descending, ties by ascending index; when fewer than 2K entries survive the mask the tail is (-inf, -1) (the
 *      reference's topk returns -inf with unspecified indices there).
 * The other postprocess_next_token_scores options (:696-708) are not applied (the reference's defaults make them no-ops). */
int gdr_beam_step(gdr_trie_t *trie, const float *logits, int64_t logits_row_stride, const int64_t *input_ids,
                  int64_t ids_row_stride, const float *beam_scores, int32_t B, int32_t K, int32_t cur_len, int32_t V,
                  int32_t eos_id, float *out_scores, int32_t *out_tokens, void *stream);

/* Replaces modeling_t5.py:1546-1571 `select_valid_embedding` (eval; last_eos_only = 0) and the
 * `logit_mask` buffer of modeling_t5.py:1279-1301 (training; last_eos_only = 1): in place on
 * logits DEV fp32 [bz, sl, V]; position t keeps tokens {t*v_out+2 .. t*v_out+v_out+1} and 1
 * (x + 0.0f), every other entry becomes x + (-1e9f). */
int gdr_position_mask(float *logits, int64_t bz, int32_t sl, int32_t V, int32_t v_out,
                      int32_t last_eos_only, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GDR_B200_H */
