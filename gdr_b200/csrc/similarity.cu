// Dense similarity on the tensor cores: out[Q, P] = q @ p^T (GDR_model/dense.py:53-54, encoder.py:128-129) for bf16 passages.
//
// The product is the grouped GEMM of score_umma.cu with ONE group: every passage row tile (128 rows) against every chunk of 32
// queries.  A tile's "score buffer offset" of query j is simply j * P + row0, so the scoring CTA (TMA-fed tcgen05, exact
// 3-term bf16 split of the fp32 queries, fp32 accumulation in TMEM) is reused unchanged — this file only writes the split query
// table and the TileMeta records (query chunk fastest: the CTAs that claim consecutive tiles share a passage tile in L2) and
// launches k_score_umma on them.  Scratch comes from the stream-ordered allocator (no handle owns this call).
#include <climits>
#include <cstring>

#include "gdr_common.cuh"

namespace gdr {

__global__ void __launch_bounds__(256) k_similarity_prepare(const float *__restrict__ q, int Q, int64_t P, int dim, __nv_bfloat16 *qsplit,
                                                            TileMeta *tiles, int n_qc, int n_tiles, int32_t *counters) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x < CTR_COUNT) counters[threadIdx.x] = threadIdx.x == CTR_N_UMMA ? n_tiles : 0;
    for (int r = warp; r < Q; r += n_warps) {                     // one warp per query row: q -> (hi, mid, lo) bf16, exact
        const float *src = q + (int64_t)r * dim;
        __nv_bfloat16 *dst = qsplit + (int64_t)r * 3 * dim;
        for (int e = lane * 4; e < dim; e += 128) {
            const float4 v = *reinterpret_cast<const float4 *>(src + e);
            const float x[4] = {v.x, v.y, v.z, v.w};
            __nv_bfloat16 t[3][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split3(x[i], t[0][i], t[1][i], t[2][i]);
#pragma unroll
            for (int term = 0; term < 3; ++term)
                *reinterpret_cast<uint2 *>(dst + term * dim + e) = *reinterpret_cast<const uint2 *>(t[term]);
        }
    }
    for (int t = warp; t < n_tiles; t += n_warps) {               // one warp per tile record
        const int rt = t / n_qc, qc = t - rt * n_qc;
        const int64_t row0 = (int64_t)rt * UMMA_ROWS;
        const int nq = min(UMMA_NQ, Q - qc * UMMA_NQ);
        TileMeta *m = tiles + t;
        if (lane == 0) {
            m->row0 = (int)row0;
            m->nrows = (int)min((int64_t)UMMA_ROWS, P - row0);
            m->nq = nq;
            m->rel0 = 0;
        }
        const int j = qc * UMMA_NQ + lane;
        m->qrow[lane] = lane < nq ? j : 0;
        m->off[lane] = lane < nq ? (int)((int64_t)j * P + row0) : 0;
    }
}

cudaError_t launch_similarity_umma(const float *q, int64_t Q, const void *p, int64_t P, int dim, float *out, cudaStream_t s, int sm_count) {
    if (dim % 64 != 0 || dim > MAX_DIM || P <= 0 || Q <= 0 || P > INT_MAX) return cudaErrorNotSupported;
    CUtensorMap tmap;
    if (!umma_make_tensor_map(&tmap, p, P, dim)) return cudaErrorNotSupported;
    // the scoring CTA indexes its output with 32 bits: queries go in chunks of at most floor((2^31 - 1) / P), a multiple of 32
    int64_t q_chunk = (((int64_t)INT_MAX) / P) / UMMA_NQ * UMMA_NQ;
    if (q_chunk <= 0) return cudaErrorNotSupported;
    if (q_chunk > Q) q_chunk = Q;
    const int64_t n_rt = (P + UMMA_ROWS - 1) / UMMA_ROWS;
    const int64_t max_tiles = n_rt * ((q_chunk + UMMA_NQ - 1) / UMMA_NQ);
    if (max_tiles > INT_MAX / 2) return cudaErrorNotSupported;
    char *ws = nullptr;
    const size_t b_split = ((size_t)q_chunk * 3 * dim * 2 + 255) / 256 * 256, b_tiles = ((size_t)max_tiles * sizeof(TileMeta) + 255) / 256 * 256;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&ws), b_split + b_tiles + 256, s);
    if (e != cudaSuccess) return e;
    __nv_bfloat16 *qsplit = reinterpret_cast<__nv_bfloat16 *>(ws);
    TileMeta *tiles = reinterpret_cast<TileMeta *>(ws + b_split);
    int32_t *counters = reinterpret_cast<int32_t *>(ws + b_split + b_tiles);
    for (int64_t q0 = 0; q0 < Q && e == cudaSuccess; q0 += q_chunk) {
        const int nq = (int)min(q_chunk, Q - q0);
        const int n_qc = (nq + UMMA_NQ - 1) / UMMA_NQ;
        const int n_tiles = (int)(n_rt * n_qc);
        k_similarity_prepare<<<sm_count * 2, 256, 0, s>>>(q + q0 * dim, nq, P, dim, qsplit, tiles, n_qc, n_tiles, counters);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
        ScoreArgs a;
        memset(&a, 0, sizeof(a));
        a.dim = dim; a.dtype = GDR_DTYPE_BF16; a.act = GDR_ACT_NONE;
        a.counters = counters; a.tile_meta = tiles; a.qsplit = qsplit;
        a.scorebuf = out + q0 * P;
        a.n_ranks = 1;
        e = launch_score_umma(a, &tmap, s, sm_count);
    }
    const cudaError_t e2 = cudaFreeAsync(ws, s);
    return e != cudaSuccess ? e : e2;
}

}  // namespace gdr
