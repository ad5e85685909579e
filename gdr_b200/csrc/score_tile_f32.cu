// Gather-and-score for fp32 stores with dense groups (BASELINE.json configs[0]: the store the reference actually holds, main_models.py:806-814,
// scored at :1582): a shared-memory-tiled fp32 kernel on the packed FMA pipe.
//
// Why not the GEMV (score_simt.cu): with ~10 pairs per cluster it re-reads every slab three times and keeps 96 registers of query fragments
// per lane (3 warps per scheduler): 0.32 of the HBM roofline at cfg1, latency-bound.  Why not tcgen05: an fp32 embedding would have to be split
// on the fly into tf32 / bf16 terms in shared memory (4x the shared-memory traffic of the tile) — not built.  This kernel keeps fp32 end to
// end: a tile = 128 rows of one cluster x up to 32 (query, beam) pairs (the SAME TileMeta records and tile queue as the tcgen05 path), K in
// chunks of 32 floats brought in by cp.async (three stages), every WARP owns four pairs and every lane four rows, and the inner product runs
// on `fma.rn.f32x2` (two fp32 FMAs per instruction and lane — the only way to the FP32 pipe's full rate on sm_100): the even and the odd k of
// a (row, pair) accumulate in the two halves of one 64-bit register and are added at the end.  Warps whose four pairs do not exist in the
// tile (groups of ~10 pairs: five of eight) only help with the loads.  Algorithmic bytes per tile = rows x dim x 4.
// Measured at cfg1 (B200): 113 us per launch = 0.46 of the HBM roofline (the GEMV: 164 us = 0.32).  It is not HBM-bound yet: per SM the
// three pipes cost about the same — HBM 52 us, packed FMAs ~30 us, shared-memory reads ~42 us (8 LDS.128 per 32 FFMA2 and warp) — and the
// barrier-separated chunk loop overlaps them poorly.  Tried without gain: five stages / four chunks in flight (the loop is not latency-
// bound), rows split over all eight warps for tiles with few pairs (fewer FMAs per LDS: more shared-memory traffic).  The next step is
// the tcgen05 kernel's structure — a producer warp, mbarriers instead of block barriers — or a tensor-core path with an on-the-fly split.
#include "gdr_common.cuh"

namespace gdr {

constexpr int TF_ROWS = UMMA_ROWS;          // 128
constexpr int TF_NQ = UMMA_NQ;              // 32
constexpr int TF_KC = 32;                   // floats per K chunk
constexpr int TF_STAGES = 3;
constexpr int TF_THREADS = 256;
constexpr int TF_ASTRIDE = TF_KC + 4;       // padded row (144 B): lanes l, l+1, ... of a quarter-warp hit distinct 16-byte bank groups
constexpr int TF_A_FLOATS = TF_ROWS * TF_ASTRIDE;
constexpr int TF_B_FLOATS = TF_NQ * TF_KC;
constexpr int TF_STAGE_FLOATS = TF_A_FLOATS + TF_B_FLOATS;
constexpr int TF_SMEM_BYTES = TF_STAGES * TF_STAGE_FLOATS * 4 + (int)sizeof(TileMeta) + 16;

__device__ __forceinline__ void ffma2(uint64_t &acc, uint64_t a, uint64_t b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

__global__ void __launch_bounds__(TF_THREADS, 2) k_score_tile_f32(ScoreArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *stage0 = reinterpret_cast<float *>(smem_raw);
    TileMeta *meta = reinterpret_cast<TileMeta *>(smem_raw + TF_STAGES * TF_STAGE_FLOATS * 4);
    int *next_tile = reinterpret_cast<int *>(meta + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *emb = reinterpret_cast<const float *>(a.emb);
    const int nkc = a.dim / TF_KC;
    pdl_wait();
    const int n_tiles = a.counters[CTR_N_UMMA];
    for (;;) {
        __syncthreads();                                           // the previous tile's metadata and stages are no longer read
        if (tid == 0) *next_tile = atomicAdd(&a.counters[CTR_TILE_NEXT], 1);
        __syncthreads();
        const int t = *next_tile;
        if (t >= n_tiles) break;
        for (int i = tid; i < (int)(sizeof(TileMeta) / 4); i += TF_THREADS)
            reinterpret_cast<int32_t *>(meta)[i] = reinterpret_cast<const int32_t *>(a.tile_meta + t)[i];
        __syncthreads();
        const int row0 = meta->row0, nrows = meta->nrows, nq = meta->nq;
        auto issue = [&](int kc) {                                 // one K chunk: 128 x 32 floats of embeddings, nq x 32 floats of queries
            float *sA = stage0 + (kc % TF_STAGES) * TF_STAGE_FLOATS, *sB = sA + TF_A_FLOATS;
#pragma unroll
            for (int i = 0; i < TF_ROWS * (TF_KC / 4) / TF_THREADS; ++i) {
                const int idx = i * TF_THREADS + tid, r = idx >> 3, c = idx & 7;
                if (r < nrows) {
                    const float *src = emb + (int64_t)(row0 + r) * a.dim + kc * TF_KC + c * 4;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sA + r * TF_ASTRIDE + c * 4)), "l"(src) : "memory");
                }
            }
            {
                const int j = tid >> 3, c = tid & 7;
                if (j < nq) {
                    const float *src = a.q + (int64_t)meta->qrow[j] * a.dim + kc * TF_KC + c * 4;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sB + j * TF_KC + c * 4)), "l"(src) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint64_t acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[r][p] = 0ull;
        const bool active = warp * 4 < nq;                         // this warp's four pairs exist (warp-uniform)
        issue(0);
        if (nkc > 1) issue(1);
        for (int kc = 0; kc < nkc; ++kc) {
            if (kc + 1 < nkc) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                       // chunk kc has landed for everyone; chunk kc-1's stage is free
            if (kc + 2 < nkc) issue(kc + 2);
            if (active) {
                const float *sA = stage0 + (kc % TF_STAGES) * TF_STAGE_FLOATS, *sB = sA + TF_A_FLOATS + warp * 4 * TF_KC;
#pragma unroll
                for (int k4 = 0; k4 < TF_KC / 4; ++k4) {
                    ulonglong2 av[4], bv[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) av[r] = *reinterpret_cast<const ulonglong2 *>(sA + (lane + 32 * r) * TF_ASTRIDE + k4 * 4);
#pragma unroll
                    for (int p = 0; p < 4; ++p) bv[p] = *reinterpret_cast<const ulonglong2 *>(sB + p * TF_KC + k4 * 4);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            ffma2(acc[r][p], av[r].x, bv[p].x);
                            ffma2(acc[r][p], av[r].y, bv[p].y);
                        }
                }
            }
        }
        if (active) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int j = warp * 4 + p;
                if (j < nq) {
                    float *dst = score_ptr(a, (int64_t)meta->off[j]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int row = lane + 32 * r;
                        const float s = __uint_as_float((uint32_t)acc[r][p]) + __uint_as_float((uint32_t)(acc[r][p] >> 32));
                        if (row < nrows) dst[row] = apply_act(s, a.act);
                    }
                }
            }
        }
    }
    // (rows of a short tile that were not loaded hold stale shared memory: their products are computed and never stored)
    __syncthreads();
    pdl_launch_dependents();
    if (tid == 0) {
        if (a.n_ranks > 1) __threadfence_system();
        if (atomicAdd(&a.counters[CTR_TILE_DONE], 1) == (int)gridDim.x - 1) {
            a.counters[CTR_TILE_NEXT] = 0;
            a.counters[CTR_TILE_DONE] = 0;
            signal_owners(a);
        }
    }
}

cudaError_t launch_score_tile_f32(const ScoreArgs &a, cudaStream_t s, int sm_count) {
    static FuncAttrOnce attr;
    cudaError_t e = attr.ensure([] { return cudaFuncSetAttribute(k_score_tile_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_BYTES); });
    if (e != cudaSuccess) return e;
    return launch_pdl(k_score_tile_f32, dim3(sm_count * 2), dim3(TF_THREADS), TF_SMEM_BYTES, s, a.launch_prio, a);
}

}  // namespace gdr
