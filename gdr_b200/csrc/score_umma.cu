// Gather-and-score, tensor-core path: TMA-fed tcgen05 grouped GEMM with TMEM accumulators.
//
// Replaces main_models.py:1456-1462 + 1577-1582 for clusters that many queries of the batch
// selected (a "group" = the (query, beam) pairs that name one cluster).  A GEMV would re-read the
// cluster slab once per pair; here a slab tile is read from HBM ONCE (TMA, 128-byte swizzle) and
// multiplied against all the group's queries on the 5th-gen tensor cores:
//
//     D[128 docs, N pairs] (fp32, TMEM)  +=  A[128 docs, 64 k] (bf16 smem, TMA)  x  B[N pairs, 64 k]^T (bf16 smem)
//
// Parity with the fp32 reference needs more than bf16(q): the fp32 query is split exactly into three
// bf16 terms q = hi + mid + lo (k_count writes them once per batch).  The three terms are stacked along N
// (B rows [0,TS) = hi, [TS,2TS) = mid, [2TS,3TS) = lo, TS = 16 or 32), so ONE MMA of N = 3*TS per K step
// reads the A tile from shared memory once for all three (SS-mode MMAs are bound by the 128 B/clk smem
// read of A, measured: three N=32 MMAs per K step were MIO-throttled); the epilogue adds the three TMEM
// column groups.  Every product bf16 x bf16 is exact in fp32, so the result matches an fp32 dot product to
// accumulation-order noise.  Per tile the kernel moves 128 x 768 x 2 B = 196 KB of embeddings and issues 12 x 4 MMAs of
// 128 x 96 x 16; it reaches 0.77 of the HBM peak on the whole device, but what bounds it is the per-K-block choreography and the
// per-launch fill / drain, not the byte stream (tools/probe_umma_limits.py; score_umma_x2.cu).
//
// Warp roles (320 threads, one persistent CTA per SM, tiles claimed one at a time from a global counter):
//   warp 0       TMA producer: A tiles of the store (L2 evict-first hint), 6-stage ring shared with B (96 KB of A in flight per SM)
//   warp 1       MMA issuer (one lane): tcgen05.mma cta_group::1 kind::f16; commits free the A and B stages
//   warps 2-4    B fillers: gather the group's query rows from the pre-split bf16 table (hi/mid/lo terms written once
//                per batch by k_count, L2-resident) with 16-byte cp.async straight into the 128B-swizzled K-major layout
//                the UMMA descriptor expects; wait_group -> fence.proxy.async -> arrive.  Two stages per warp.
//   warps 5-8    epilogue: tcgen05.ld the accumulator (double-buffered in TMEM so it overlaps the next
//                tile's MMAs), activation, coalesced stores into each query's candidate segment
//   warp 9       tile scheduler (one lane): atomicAdd on the tile counter, then ONE cp.async.bulk of the tile's 272-byte
//                TileMeta record (row range, query rows, score offsets: resolved once per batch by k_tilemeta) into a
//                two-slot shared-memory ring that every other role reads; a record with nq = -1 ends all loops.  No role
//                has a dependent global load on its per-tile critical path (measured: with each role walking item ->
//                pair -> candidate offset itself the empty barrier skeleton alone cost 4 us per tile, as much as the
//                tile's HBM time), and a CTA that runs slower or starts later simply claims fewer tiles.
// The dependent grid (the batch's top-k) is released at the kernel's END (see the note at pdl_launch_dependents below).
#include "gdr_common.cuh"
#include "score_umma.cuh"

namespace gdr {

#define UM_FILL_IDX (warp - 2)
#define UM_DISPATCH                                       \
    if (warp == 0) role_tma();                            \
    else if (warp == 1) role_mma();                       \
    else if (warp < 2 + UM_FILL_WARPS) role_fill();       \
    else if (warp < 2 + UM_FILL_WARPS + 4) role_epi();    \
    else role_sched();
#define UM_EXTRA_TAIL
#define UM_SCORE_PTR(o) (a.scorebuf + (o))
#define UM_P2P_FENCE
#define UM_P2P_SIGNAL
#define UM_P2P_WAIT_CONSUMED
__global__ void __launch_bounds__(UM_THREADS, 1) k_score_umma(const __grid_constant__ CUtensorMap tmap, ScoreArgs a) {
#include "score_umma_body.inc"
}
#undef UM_SCORE_PTR
#undef UM_P2P_FENCE
#undef UM_P2P_SIGNAL
#undef UM_P2P_WAIT_CONSUMED
// the same CTA for one shard of a cluster-sharded corpus: scores go straight into the score buffer of the query's owner (a peer
// GPU's memory over NVLink), the last CTA raises this rank's arrival flag on every owner (gdr_common.cuh, include/gdr_b200.h)
#define UM_SCORE_PTR(o) score_ptr(a, (o))
#define UM_P2P_FENCE __threadfence_system();
#define UM_P2P_SIGNAL signal_owners(a);
#define UM_P2P_WAIT_CONSUMED     /* the one-warp k_wait_consumed launched in front of this kernel did */
__global__ void __launch_bounds__(UM_THREADS, 1) k_score_umma_p2p(const __grid_constant__ CUtensorMap tmap, ScoreArgs a) {
#include "score_umma_body.inc"
}
#undef UM_SCORE_PTR
#undef UM_P2P_FENCE
#undef UM_P2P_SIGNAL
#undef UM_P2P_WAIT_CONSUMED
#undef UM_FILL_IDX
#undef UM_DISPATCH
#undef UM_EXTRA_TAIL

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool umma_make_tensor_map(CUtensorMap *out, const void *emb, int64_t n_docs, int dim) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)n_docs};
    const cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
    const cuuint32_t box[2] = {UM_BLOCK_K, UMMA_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn)(
        out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(emb), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

cudaError_t launch_score_umma(const ScoreArgs &a, const CUtensorMap *tmap, cudaStream_t s, int sm_count) {
    static FuncAttrOnce attr;
    cudaError_t e = attr.ensure([] {
        cudaError_t e2 = cudaFuncSetAttribute(k_score_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_BYTES);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_score_umma_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_BYTES);
        return e2;
    });
    if (e != cudaSuccess) return e;
    if (a.n_ranks > 1) return launch_pdl(k_score_umma_p2p, dim3(sm_count), dim3(UM_THREADS), UM_SMEM_BYTES, s, a.launch_prio, *tmap, a);
    return launch_pdl(k_score_umma, dim3(sm_count), dim3(UM_THREADS), UM_SMEM_BYTES, s, a.launch_prio, *tmap, a);
}

}  // namespace gdr
