// Gather-and-score, SIMT path: 128-bit vectorised loads + warp-shuffle transposed reduction.
//
// Replaces main_models.py:1456-1462 (per-document .cuda() + torch.cat gather) and
// main_models.py:1577-1582 (broadcast mul + sum(-1) + tanh/sigmoid) for groups too small to fill a
// tensor-core tile, and for fp32 stores.  One warp owns one work item: up to SIMT_ROWS consecutive
// rows of one cluster x up to SIMT_QT (query, beam) pairs.  The query fragments live in registers
// (lane l owns 16-byte chunks l, l+32, ... of the row), rows are streamed with LDG.128
// (L1::no_allocate), fp32 FMA accumulate, and RB x QT partial sums are reduced across the warp with
// a recursive-halving transpose (16 shuffles for 16 values instead of 80).
//
// HBM-bound by design: algorithmic bytes per item = nrows * dim * sizeof(T); with QT pairs per
// item a slab is re-read ceil(group/QT) times (from L2), which is why dense groups go to the
// tcgen05 path instead (score_umma.cu).
#include "gdr_common.cuh"

namespace gdr {

__device__ __forceinline__ uint4 ldg_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename T> struct Chunk;
template <> struct Chunk<float> {
    static constexpr int EPC = 4;
    __device__ static __forceinline__ void unpack(const uint4 &v, float *x) {
        x[0] = __uint_as_float(v.x); x[1] = __uint_as_float(v.y);
        x[2] = __uint_as_float(v.z); x[3] = __uint_as_float(v.w);
    }
};
template <> struct Chunk<__nv_bfloat16> {
    static constexpr int EPC = 8;
    __device__ static __forceinline__ void unpack(const uint4 &v, float *x) {
        // bf16 -> fp32 is a 16-bit shift: exact
        x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xffff0000u);
        x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xffff0000u);
        x[4] = __uint_as_float(v.z << 16); x[5] = __uint_as_float(v.z & 0xffff0000u);
        x[6] = __uint_as_float(v.w << 16); x[7] = __uint_as_float(v.w & 0xffff0000u);
    }
};

// Recursive-halving transpose-reduce of NV (power of two <= 32) per-lane partials.  On return
// vals[0] of lane l is the warp-wide sum of value index (l >> log2(32/NV)).
template <int NV>
__device__ __forceinline__ void transpose_reduce(float (&vals)[NV], int lane) {
    int bit = 16;
#pragma unroll
    for (int w = NV; w > 1; w >>= 1) {
        const int half = w >> 1;
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? vals[i] : vals[i + half];
            const float keep = upper ? vals[i + half] : vals[i];
            vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
        bit >>= 1;
    }
#pragma unroll
    for (; bit >= 1; bit >>= 1) vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], bit);
}

// value index held by `lane` after transpose_reduce<NV>
template <int NV> __device__ __forceinline__ int reduced_index(int lane) { return lane / (32 / NV); }
template <int NV> __device__ __forceinline__ bool reduced_writer(int lane) { return (lane % (32 / NV)) == 0; }

// Score `nrows` rows starting at `rows` against QT query rows.  qrow[j] < 0 = absent query.
// Lane's result slot: value index v = reduced_index(lane) -> row (v / QT) of the batch, query (v % QT);
// `my_dst` is that query's destination for row 0 (may be null when the lane's query is absent).
template <typename T, int CPL, int QT, int RB>
__device__ __forceinline__ void score_rows(const T *__restrict__ rows, int dim, int nrows, const float *__restrict__ q,
                                           const int (&qrow)[QT], float *my_dst, int act, int lane) {
    constexpr int EPC = Chunk<T>::EPC;
    constexpr int NV = RB * QT;
    const int nchunks = dim / EPC;
    float qv[QT][CPL * EPC];
#pragma unroll
    for (int j = 0; j < QT; ++j) {
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
            const int ch = lane + 32 * t;
#pragma unroll
            for (int e = 0; e < EPC; e += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (qrow[j] >= 0 && ch < nchunks)
                    v = __ldg(reinterpret_cast<const float4 *>(q + (int64_t)qrow[j] * dim + ch * EPC + e));
                qv[j][t * EPC + e + 0] = v.x; qv[j][t * EPC + e + 1] = v.y;
                qv[j][t * EPC + e + 2] = v.z; qv[j][t * EPC + e + 3] = v.w;
            }
        }
    }
    const int v_idx = reduced_index<NV>(lane);
    const int my_r = v_idx / QT;
    const bool writer = reduced_writer<NV>(lane) && my_dst != nullptr;
    const char *base = reinterpret_cast<const char *>(rows);
    const int64_t row_bytes = (int64_t)dim * sizeof(T);

    for (int r0 = 0; r0 < nrows; r0 += RB) {
        uint4 data[RB][CPL];
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                const int ch = lane + 32 * t;
                data[rr][t] = make_uint4(0u, 0u, 0u, 0u);
                if (r0 + rr < nrows && ch < nchunks)
                    data[rr][t] = ldg_stream(base + (int64_t)(r0 + rr) * row_bytes + (int64_t)ch * 16);
            }
        }
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
            for (int t = 0; t < CPL; ++t) {
                float x[EPC];
                Chunk<T>::unpack(data[rr][t], x);
#pragma unroll
                for (int e = 0; e < EPC; ++e) {
#pragma unroll
                    for (int j = 0; j < QT; ++j) acc[rr * QT + j] = fmaf(x[e], qv[j][t * EPC + e], acc[rr * QT + j]);
                }
            }
        }
        transpose_reduce<NV>(acc, lane);
        if (writer && r0 + my_r < nrows) my_dst[r0 + my_r] = apply_act(acc[0], act);
    }
}

// Rows kept in flight per warp, sized so loaded chunks + query fragments stay inside the register file.
template <typename T, int CPL> struct RowsInFlight {
    static constexpr int RB4 = (sizeof(T) == 2 && CPL <= 3) ? 4 : 2;
    static constexpr int RB1 = CPL <= 3 ? 8 : (CPL <= 6 ? 4 : 2);
};

template <typename T, int CPL>
__global__ void __launch_bounds__(128) k_score_simt(ScoreArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    pdl_wait();
    const int n_items = a.counters[CTR_N_SIMT];
    const bool per_beam = (a.flags & GDR_Q_PER_BEAM) != 0;
    const T *emb = reinterpret_cast<const T *>(a.emb);
    constexpr int RB4 = RowsInFlight<T, CPL>::RB4;   // rows in flight with 4 queries
    constexpr int RB1 = RowsInFlight<T, CPL>::RB1;   // rows in flight with 1 query

    for (int it = warp; it < n_items; it += n_warps) {
        const Item item = a.simt_items[it];
        const int nrows = item.nrows_nq & 0xffff;
        const int nq = item.nrows_nq >> 16;
        const T *rows = emb + (int64_t)item.row0 * a.dim;
        if (nq == 1) {
            const int p = a.grp_pair[item.slot0];
            const int b = p / a.K;
            const int qrow[1] = {per_beam ? p : b};
            float *dst = score_ptr(a, pack_score_off(a, b, a.candoff[p + b] + item.rel0));
            score_rows<T, CPL, 1, RB1>(rows, a.dim, nrows, a.q, qrow, dst, a.act, lane);
        } else {
            int qrow[SIMT_QT];
#pragma unroll
            for (int j = 0; j < SIMT_QT; ++j) {
                qrow[j] = -1;
                if (j < nq) {
                    const int p = a.grp_pair[item.slot0 + j];
                    qrow[j] = per_beam ? p : p / a.K;
                }
            }
            const int my_j = reduced_index<RB4 * SIMT_QT>(lane) % SIMT_QT;
            float *dst = nullptr;
            if (my_j < nq) {
                const int p = a.grp_pair[item.slot0 + my_j];
                const int b = p / a.K;
                dst = score_ptr(a, pack_score_off(a, b, a.candoff[p + b] + item.rel0));
            }
            score_rows<T, CPL, SIMT_QT, RB4>(rows, a.dim, nrows, a.q, qrow, dst, a.act, lane);
        }
    }
    if (a.n_ranks > 1) {              // sharded corpus: the last CTA to finish tells the owners that this rank's scores have landed
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(&a.counters[CTR_SIMT_DONE], 1) == (int)gridDim.x - 1) {
                a.counters[CTR_SIMT_DONE] = 0;
                signal_owners(a);
            }
        }
    }
    pdl_launch_dependents();      // at the end: released at entry, the top-k's CTAs would sit resident for this whole kernel (see k_score_umma)
}

// Dense similarity (dense.py:53-54): out[Q, P] = q @ p^T.  Same core; items are arithmetic:
// item -> (row tile of SIMT_ROWS rows of p, chunk of SIMT_QT queries), query chunk fastest.
template <typename T, int CPL>
__global__ void __launch_bounds__(128) k_similarity(const float *__restrict__ q, int64_t Q, const T *__restrict__ p,
                                                    int64_t P, int dim, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_qc = (Q + SIMT_QT - 1) / SIMT_QT;
    const int64_t n_rt = (P + SIMT_ROWS - 1) / SIMT_ROWS;
    constexpr int RB4 = RowsInFlight<T, CPL>::RB4;
    for (int64_t it = warp; it < n_qc * n_rt; it += n_warps) {
        const int64_t rt = it / n_qc, qc = it % n_qc;
        const int64_t row0 = rt * SIMT_ROWS;
        const int nrows = (int)min((int64_t)SIMT_ROWS, P - row0);
        int qrow[SIMT_QT];
#pragma unroll
        for (int j = 0; j < SIMT_QT; ++j) qrow[j] = (qc * SIMT_QT + j < Q) ? (int)(qc * SIMT_QT + j) : -1;
        const int my_j = reduced_index<RB4 * SIMT_QT>(lane) % SIMT_QT;
        float *dst = (qc * SIMT_QT + my_j < Q) ? out + (qc * SIMT_QT + my_j) * P + row0 : nullptr;
        score_rows<T, CPL, SIMT_QT, RB4>(p + row0 * dim, dim, nrows, q, qrow, dst, GDR_ACT_NONE, lane);
    }
}

// Leaf-cluster centroids: out[c] = mean of the rows of cluster c, fp32, rows added in store order —
// the order `sum([embedding[i] for i in embedding_index])` uses in tree_embedding_calculate (main_models.py:154-158).
// One CTA per cluster, one 16-byte chunk of the row per thread; a single pass over the store (HBM-bound, run once
// per index build / expansion).
template <typename T>
__global__ void __launch_bounds__(256) k_centroids(const T *__restrict__ emb, const int32_t *__restrict__ offsets, int dim,
                                                   float *__restrict__ out) {
    constexpr int EPC = Chunk<T>::EPC;
    const int c = blockIdx.x;
    const int lo = offsets[c], hi = offsets[c + 1];
    const int nchunks = dim / EPC;
    for (int ch = threadIdx.x; ch < nchunks; ch += blockDim.x) {
        float acc[EPC];
#pragma unroll
        for (int e = 0; e < EPC; ++e) acc[e] = 0.f;
        for (int r = lo; r < hi; ++r) {
            float x[EPC];
            Chunk<T>::unpack(ldg_stream(reinterpret_cast<const char *>(emb + (int64_t)r * dim) + (int64_t)ch * 16), x);
#pragma unroll
            for (int e = 0; e < EPC; ++e) acc[e] = __fadd_rn(acc[e], x[e]);
        }
        const float n = (float)(hi - lo);
#pragma unroll
        for (int e = 0; e < EPC; ++e) out[(int64_t)c * dim + ch * EPC + e] = hi > lo ? __fdiv_rn(acc[e], n) : 0.f;
    }
}

cudaError_t launch_centroids(const void *emb, int dtype, const int32_t *offsets, int n_clusters, int dim, float *out, cudaStream_t s) {
    if (n_clusters == 0) return cudaSuccess;
    if (dtype == GDR_DTYPE_BF16) k_centroids<__nv_bfloat16><<<n_clusters, 128, 0, s>>>((const __nv_bfloat16 *)emb, offsets, dim, out);
    else k_centroids<float><<<n_clusters, 256, 0, s>>>((const float *)emb, offsets, dim, out);
    return cudaGetLastError();
}

template <typename T> static int cpl_for(int dim) {
    const int nchunks = dim / Chunk<T>::EPC;
    const int cpl = (nchunks + 31) / 32;
    const int allowed[] = {1, 2, 3, 4, 6, 8};
    for (int v : allowed) if (cpl <= v) return v;
    return -1;
}

#define GDR_DISPATCH_CPL(T, cpl, ...)                   \
    switch (cpl) {                                      \
        case 1: { constexpr int CPL = 1; __VA_ARGS__; break; } \
        case 2: { constexpr int CPL = 2; __VA_ARGS__; break; } \
        case 3: { constexpr int CPL = 3; __VA_ARGS__; break; } \
        case 4: { constexpr int CPL = 4; __VA_ARGS__; break; } \
        case 6: { constexpr int CPL = 6; __VA_ARGS__; break; } \
        case 8: { constexpr int CPL = 8; __VA_ARGS__; break; } \
        default: return cudaErrorInvalidValue;          \
    }

cudaError_t launch_score_simt(const ScoreArgs &a, cudaStream_t s, int sm_count) {
    const int grid = sm_count * 8;   // persistent warps: 8 CTAs x 4 warps per SM, items strided
    if (a.dtype == GDR_DTYPE_BF16) {
        GDR_DISPATCH_CPL(__nv_bfloat16, cpl_for<__nv_bfloat16>(a.dim), return launch_pdl(k_score_simt<__nv_bfloat16, CPL>, dim3(grid), dim3(128), 0, s, a.launch_prio, a));
    } else {
        GDR_DISPATCH_CPL(float, cpl_for<float>(a.dim), return launch_pdl(k_score_simt<float, CPL>, dim3(grid), dim3(128), 0, s, a.launch_prio, a));
    }
    return cudaGetLastError();
}

cudaError_t launch_similarity(const float *q, int64_t Q, const void *p, int64_t P, int dim, int p_dtype, float *out,
                              cudaStream_t s, int sm_count) {
    if (Q == 0 || P == 0) return cudaSuccess;
    const int64_t items = ((Q + SIMT_QT - 1) / SIMT_QT) * ((P + SIMT_ROWS - 1) / SIMT_ROWS);
    const int grid = (int)min((int64_t)sm_count * 8, (items + 3) / 4);
    if (p_dtype == GDR_DTYPE_BF16) {
        GDR_DISPATCH_CPL(__nv_bfloat16, cpl_for<__nv_bfloat16>(dim),
                         (k_similarity<__nv_bfloat16, CPL><<<grid, 128, 0, s>>>(q, Q, (const __nv_bfloat16 *)p, P, dim, out)));
    } else {
        GDR_DISPATCH_CPL(float, cpl_for<float>(dim), (k_similarity<float, CPL><<<grid, 128, 0, s>>>(q, Q, (const float *)p, P, dim, out)));
    }
    return cudaGetLastError();
}

}  // namespace gdr
