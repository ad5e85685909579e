// Docid logit masks.
//
// k_tree_mask      replaces generation_utils_previous.py:714-729: per beam row a Python walk of the
//                  `Node` trie (input_ids[i].tolist() -> dict lookups), a full-size -inf temporary
//                  and one index_put per row, then `scores += mask`.
// k_position_mask  replaces modeling_t5.py:1546-1571 (`select_valid_embedding`) and the training
//                  buffer of modeling_t5.py:1279-1301.
//
// Both are pure HBM streams.  The tree mask CTA walks the device trie for its row (warp 0, one
// 32-wide compare per level), publishes the allowed tokens as a V-bit bitmap in shared memory,
// then streams its slice of the row with 128-bit stores.  In the default mode masked entries are
// written as -inf without being read (write-only stream, half the traffic of read-modify-write);
// strict mode reads everything and computes s + (-inf) so NaN / +inf inputs propagate exactly as in
// the reference.  Allowed entries are always read and rewritten as s + 0.0f (so -0.0 becomes
// +0.0, as `scores += mask` does).
#include "gdr_common.cuh"

namespace gdr {

constexpr int TM_THREADS = 256;

// Walk the trie along ids[1 .. cur_len).  Returns the node reached, or -1 if the path leaves the tree.
// Called by one full warp.
__device__ __forceinline__ int trie_walk(const int32_t *__restrict__ first_child, const int32_t *__restrict__ child_tok,
                                         const int32_t *__restrict__ child_node, const int64_t *__restrict__ ids,
                                         int cur_len, int lane) {
    int node = 0;
    for (int t = 1; t < cur_len && node >= 0; ++t) {
        const int64_t tok = ids[t];
        const int lo = first_child[node], hi = first_child[node + 1];
        int next = -1;
        for (int e0 = lo; e0 < hi && next < 0; e0 += 32) {
            const int e = e0 + lane;
            const bool hit = e < hi && (int64_t)child_tok[e] == tok;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) next = child_node[e0 + __ffs(m) - 1];
        }
        node = next;
    }
    return node;
}

// grid = (chunks per row, R).  Each CTA redoes the (cheap, L2-resident) walk for its row.
__global__ void __launch_bounds__(TM_THREADS) k_tree_mask(const int32_t *__restrict__ first_child,
                                                          const int32_t *__restrict__ child_tok,
                                                          const int32_t *__restrict__ child_node,
                                                          const int64_t *__restrict__ input_ids, int64_t ids_stride,
                                                          int cur_len, float *__restrict__ scores, int64_t scores_stride,
                                                          int V, int eos_id, int strict, int chunk) {
    extern __shared__ uint32_t bitmap[];   // chunk/32 words: allowed tokens inside this CTA's slice
    __shared__ int s_node;
    const int r = blockIdx.y;
    const int v0 = blockIdx.x * chunk;
    const int v1 = min(V, v0 + chunk);
    const int tid = threadIdx.x, lane = tid & 31;
    const int words = (chunk + 31) / 32;
    for (int i = tid; i < words; i += TM_THREADS) bitmap[i] = 0u;
    if (tid < 32) {
        const int node = trie_walk(first_child, child_tok, child_node, input_ids + (int64_t)r * ids_stride, cur_len, lane);
        if (lane == 0) s_node = node;
    }
    __syncthreads();
    const int node = s_node;
    if (node < 0) {
        if (tid == 0 && eos_id >= v0 && eos_id < v1) atomicOr(&bitmap[(eos_id - v0) >> 5], 1u << ((eos_id - v0) & 31));
    } else {
        const int lo = first_child[node], hi = first_child[node + 1];
        for (int e = lo + tid; e < hi; e += TM_THREADS) {
            const int tok = child_tok[e];
            if (tok >= v0 && tok < v1) atomicOr(&bitmap[(tok - v0) >> 5], 1u << ((tok - v0) & 31));
        }
    }
    __syncthreads();
    float *row = scores + (int64_t)r * scores_stride;
    const float ninf = -INFINITY;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(row + v0) & 15) == 0);
    if (vec_ok) {
        const int n4 = (v1 - v0) >> 2;
        float4 *row4 = reinterpret_cast<float4 *>(row + v0);
        for (int i = tid; i < n4; i += TM_THREADS) {
            const unsigned bits = (bitmap[i >> 3] >> ((i & 7) * 4)) & 0xfu;
            float4 v;
            if (bits == 0u && !strict) {
                v = make_float4(ninf, ninf, ninf, ninf);
            } else {
                v = row4[i];
                v.x = __fadd_rn(v.x, (bits & 1u) ? 0.0f : ninf);
                v.y = __fadd_rn(v.y, (bits & 2u) ? 0.0f : ninf);
                v.z = __fadd_rn(v.z, (bits & 4u) ? 0.0f : ninf);
                v.w = __fadd_rn(v.w, (bits & 8u) ? 0.0f : ninf);
            }
            row4[i] = v;
        }
        for (int i = v0 + (n4 << 2) + tid; i < v1; i += TM_THREADS) {
            const bool ok = (bitmap[(i - v0) >> 5] >> ((i - v0) & 31)) & 1u;
            row[i] = (ok || strict) ? __fadd_rn(row[i], ok ? 0.0f : ninf) : ninf;
        }
    } else {
        for (int i = v0 + tid; i < v1; i += TM_THREADS) {
            const bool ok = (bitmap[(i - v0) >> 5] >> ((i - v0) & 31)) & 1u;
            row[i] = (ok || strict) ? __fadd_rn(row[i], ok ? 0.0f : ninf) : ninf;
        }
    }
}

cudaError_t launch_tree_mask(const int32_t *first_child, const int32_t *child_tok, const int32_t *child_node,
                             const int64_t *input_ids, int64_t ids_stride, int R, int cur_len, float *scores,
                             int64_t scores_stride, int V, int eos_id, int strict, cudaStream_t s) {
    if (R == 0 || V == 0) return cudaSuccess;
    // 8192-float slices (32 KB per CTA, 8 x 128-bit stores per thread): 4 CTAs per 32,128-wide row
    const int chunk = 8192;
    const int chunks = (V + chunk - 1) / chunk;
    for (int r0 = 0; r0 < R; r0 += 65535) {
        const int rows = min(65535, R - r0);
        dim3 grid(chunks, rows);
        k_tree_mask<<<grid, TM_THREADS, (chunk / 32) * sizeof(uint32_t), s>>>(
            first_child, child_tok, child_node, input_ids + (int64_t)r0 * ids_stride, ids_stride, cur_len,
            scores + (int64_t)r0 * scores_stride, scores_stride, V, eos_id, strict, chunk);
    }
    return cudaGetLastError();
}

// logits [bz, sl, V]: position t keeps {t*v_out+2 .. t*v_out+v_out+1} ∪ {1}; with last_eos_only the last
// position keeps only {1} (modeling_t5.py:1296).
__global__ void __launch_bounds__(256) k_position_mask(float *__restrict__ logits, int64_t n_rows, int sl, int V,
                                                       int v_out, int last_eos_only) {
    const int64_t total = n_rows * V;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / V;
        const int v = (int)(idx - row * V);
        const int t = (int)(row % sl);
        bool ok = (v == 1);
        if (!(last_eos_only && t == sl - 1)) ok = ok || (v >= t * v_out + 2 && v < t * v_out + v_out + 2);
        logits[idx] = __fadd_rn(logits[idx], ok ? 0.0f : -1e9f);
    }
}

cudaError_t launch_position_mask(float *logits, int64_t bz, int sl, int V, int v_out, int last_eos_only, cudaStream_t s) {
    const int64_t total = bz * sl * V;
    if (total == 0) return cudaSuccess;
    const int grid = (int)min((int64_t)148 * 16, (total + 255) / 256);
    k_position_mask<<<grid, 256, 0, s>>>(logits, bz * sl, sl, V, v_out, last_eos_only);
    return cudaGetLastError();
}

}  // namespace gdr
