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
#include <type_traits>

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

// ---- fused beam step (the "next" row after the mask itself) ----------------------------------------
// Replaces the three full-size passes around the mask in the reference's beam search,
//   generation_utils_previous.py:694  scores = log_softmax(next_token_logits)          (read + write R*V)
//   generation_utils_previous.py:714-729  scores += tree mask                            (read + write R*V)
//   :757-771  next_scores = (scores + beam_scores[:, None]).view(B, K*V); topk(2K)       (read + write + read R*V)
// by ONE read of the logits: after the mask at most `fanout` entries of a row are finite, so a row only needs its
// log-sum-exp (online, one pass) and the log-probabilities of the allowed tokens.  k_beam_rows writes those candidates
// (value, beam*V + token); the per-query top-2K over the K*fanout candidates is the merge kernel of topk.cu.
// HBM-bound: algorithmic bytes = R*V*4.
__global__ void __launch_bounds__(TM_THREADS) k_beam_rows(const int32_t *__restrict__ first_child,
                                                          const int32_t *__restrict__ child_tok,
                                                          const int32_t *__restrict__ child_node,
                                                          const int64_t *__restrict__ input_ids, int64_t ids_stride, int cur_len,
                                                          const float *__restrict__ logits, int64_t logits_stride, int V,
                                                          const float *__restrict__ beam_scores, int K, int eos_id, int fanout,
                                                          float *__restrict__ cand_val, int32_t *__restrict__ cand_id) {
    __shared__ int s_node;
    __shared__ float s_m[TM_THREADS / 32], s_s[TM_THREADS / 32];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *row = logits + (int64_t)r * logits_stride;
    if (warp == 0) {
        const int node = trie_walk(first_child, child_tok, child_node, input_ids + (int64_t)r * ids_stride, cur_len, lane);
        if (lane == 0) s_node = node;
    }
    // online log-sum-exp: running maximum m and sum s of exp(x - m)
    float m = -INFINITY, s = 0.f;
    auto push = [&](float x) {
        if (x > m) { s = s * expf(m - x) + 1.f; m = x; }
        else if (x > -INFINITY) s += expf(x - m);
    };
    if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        const int n4 = V >> 2;
        for (int i = tid; i < n4; i += TM_THREADS) {
            const float4 v = __ldcs(row4 + i);      // streamed once
            push(v.x); push(v.y); push(v.z); push(v.w);
        }
        for (int i = (n4 << 2) + tid; i < V; i += TM_THREADS) push(row[i]);
    } else {
        for (int i = tid; i < V; i += TM_THREADS) push(row[i]);
    }
    auto combine = [](float &m1, float &s1, float m2, float s2) {
        const float mx = fmaxf(m1, m2);
        if (mx == -INFINITY) { s1 = 0.f; return; }
        s1 = s1 * expf(m1 - mx) + s2 * expf(m2 - mx);
        m1 = mx;
    };
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, d), s2 = __shfl_xor_sync(0xffffffffu, s, d);
        combine(m, s, m2, s2);
    }
    if (lane == 0) { s_m[warp] = m; s_s[warp] = s; }
    __syncthreads();
    m = s_m[0]; s = s_s[0];
    for (int w = 1; w < TM_THREADS / 32; ++w) combine(m, s, s_m[w], s_s[w]);
    const float log_s = logf(s);
    const int node = s_node;
    const int lo = node < 0 ? 0 : first_child[node];
    const int n_allowed = node < 0 ? 1 : first_child[node + 1] - lo;
    const float beam = beam_scores[r];
    for (int e = tid; e < fanout; e += TM_THREADS) {
        float val = -INFINITY;
        int32_t id = -1;
        if (e < n_allowed) {
            const int tok = node < 0 ? eos_id : child_tok[lo + e];
            if (tok >= 0 && tok < V) {
                // log_softmax as torch computes it, (x - max) - log(sum), then "+ 0" (mask) and "+ beam score"
                val = __fadd_rn(__fadd_rn(__fsub_rn(__fsub_rn(row[tok], m), log_s), 0.0f), beam);
                id = (r % K) * V + tok;
            }
        }
        cand_val[(int64_t)r * fanout + e] = val;
        cand_id[(int64_t)r * fanout + e] = id;
    }
}

cudaError_t launch_beam_rows(const int32_t *first_child, const int32_t *child_tok, const int32_t *child_node,
                             const int64_t *input_ids, int64_t ids_stride, int cur_len, const float *logits,
                             int64_t logits_stride, int V, const float *beam_scores, int R, int K, int eos_id, int fanout,
                             float *cand_val, int32_t *cand_id, cudaStream_t s) {
    if (R == 0) return cudaSuccess;
    k_beam_rows<<<R, TM_THREADS, 0, s>>>(first_child, child_tok, child_node, input_ids, ids_stride, cur_len, logits, logits_stride,
                                         V, beam_scores, K, eos_id, fanout, cand_val, cand_id);
    return cudaGetLastError();
}

// logits [bz, sl, V]: position t keeps {t*v_out+2 .. t*v_out+v_out+1} ∪ {1}; with last_eos_only the last
// position keeps only {1} (modeling_t5.py:1296).  One warp per row: the position and its kept token range are computed once
// per row (round 1 did a 64-bit division per ELEMENT and reached 0.39 of the HBM peak), the row is then a streamed
// read-modify-write with VEC-wide accesses (VEC = 4 when V % 4 == 0, 2 when V is even — T5's 302 — so that every row start
// keeps the alignment) and four of them in flight per lane.  HBM-bound: 2 * 4 bytes per element.
template <int VEC>
__global__ void __launch_bounds__(256) k_position_mask(float *__restrict__ logits, int64_t n_rows, int sl, int V,
                                                       int v_out, int last_eos_only) {
    using Vec = typename std::conditional<VEC == 4, float4, typename std::conditional<VEC == 2, float2, float>::type>::type;
    constexpr int U = 8;                                    // vector loads a lane has in flight before its first store (the row is
                                                            // modified in place: without the explicit batch every load waits for the store before it)
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nv = V / VEC;
    for (int64_t row = warp; row < n_rows; row += n_warps) {
        const int t = (int)(row % sl);
        const bool digits = !(last_eos_only && t == sl - 1);
        const int lo = digits ? t * v_out + 2 : V, hi = digits ? t * v_out + v_out + 2 : V;      // kept range [lo, hi), plus token 1
        Vec *p = reinterpret_cast<Vec *>(logits + row * V);
        auto fix = [&](float x, int v) { return __fadd_rn(x, (v == 1 || (v >= lo && v < hi)) ? 0.0f : -1e9f); };
        for (int i0 = 0; i0 < nv; i0 += 32 * U) {
            Vec x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * 32 + lane;
                if (i < nv) x[u] = __ldcs(p + i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * 32 + lane;
                if (i < nv) {
                    const int v = i * VEC;
                    if constexpr (VEC == 4) { x[u].x = fix(x[u].x, v); x[u].y = fix(x[u].y, v + 1); x[u].z = fix(x[u].z, v + 2); x[u].w = fix(x[u].w, v + 3); }
                    else if constexpr (VEC == 2) { x[u].x = fix(x[u].x, v); x[u].y = fix(x[u].y, v + 1); }
                    else x[u] = fix(x[u], v);
                    __stcs(p + i, x[u]);
                }
            }
        }
    }
}

cudaError_t launch_position_mask(float *logits, int64_t bz, int sl, int V, int v_out, int last_eos_only, cudaStream_t s) {
    const int64_t n_rows = bz * sl;
    if (n_rows == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)min((int64_t)sms * 8, (n_rows + 7) / 8);              // 8 warps per CTA, 8 CTAs per SM, rows strided over the warps
    const bool a16 = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
    if (V % 4 == 0 && a16) k_position_mask<4><<<grid, 256, 0, s>>>(logits, n_rows, sl, V, v_out, last_eos_only);
    else if (V % 2 == 0 && (reinterpret_cast<uintptr_t>(logits) & 7) == 0) k_position_mask<2><<<grid, 256, 0, s>>>(logits, n_rows, sl, V, v_out, last_eos_only);
    else k_position_mask<1><<<grid, 256, 0, s>>>(logits, n_rows, sl, V, v_out, last_eos_only);
    return cudaGetLastError();
}

}  // namespace gdr
