// Node embeddings of the prefix tree and the greedy descent over them (index expansion).
//
// Replaces tree_embedding_calculate (main_models.py:154-179) and tree_match (main_models.py:232-252) on the device
// trie.  Nodes are numbered breadth-first, so every depth is a contiguous id range: the embeddings are filled level by
// level from the deepest, one warp per node; a leaf cluster copies its centroid (gdr_cluster_centroids), any other node
// takes the leaf-count-weighted mean of its children accumulated in the reference's order (the children's insertion
// order, `child_order`) with the reference's operations (fp32 multiply, add, one divide — no FMA contraction), so the
// result is bit-identical to the reference's for fp32 inputs.  The descent is one warp per document.
#include "gdr_common.cuh"

namespace gdr {

__global__ void __launch_bounds__(128) k_node_embeddings(const int32_t *__restrict__ first_child, const int32_t *__restrict__ child_node,
                                                         const int32_t *__restrict__ child_order, const int32_t *__restrict__ node_cluster,
                                                         const float *__restrict__ leaf_emb, const int32_t *__restrict__ leaf_num, int dim,
                                                         int node_lo, int node_hi, float *node_emb, int32_t *node_leaf_num) {
    const int lane = threadIdx.x & 31;
    const int n = node_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (n >= node_hi) return;
    float *out = node_emb + (int64_t)n * dim;
    const int c = node_cluster[n];
    if (c >= 0) {                                           // main_models.py:156-159: a node that lists documents
        for (int d = lane; d < dim; d += 32) out[d] = leaf_emb[(int64_t)c * dim + d];
        if (lane == 0) node_leaf_num[n] = leaf_num[c];
        return;
    }
    const int e0 = first_child[n], e1 = first_child[n + 1];
    int total = 0;
    for (int e = e0; e < e1; ++e) total += node_leaf_num[child_node[e]];
    if (lane == 0) node_leaf_num[n] = total;                // 0: no embedding (the EOS child of a leaf cluster)
    if (total == 0) return;
    for (int d = lane; d < dim; d += 32) {
        float acc = 0.f;
        bool first = true;
        for (int r = e0; r < e1; ++r) {                     // :168-173, children in insertion order
            const int ch = child_node[child_order[r]];
            const int num = node_leaf_num[ch];
            if (num == 0) continue;
            const float term = __fmul_rn(node_emb[(int64_t)ch * dim + d], (float)num);
            acc = first ? term : __fadd_rn(acc, term);
            first = false;
        }
        out[d] = __fdiv_rn(acc, (float)total);              // :175
    }
}

__global__ void __launch_bounds__(128) k_tree_match(const int32_t *__restrict__ first_child, const int32_t *__restrict__ child_tok,
                                                    const int32_t *__restrict__ child_node, const int32_t *__restrict__ child_order,
                                                    const float *__restrict__ node_emb, const int32_t *__restrict__ node_leaf_num, int dim,
                                                    const float *__restrict__ docs, int M, int max_len, int32_t *out_tokens, int32_t *out_len) {
    const int lane = threadIdx.x & 31;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const float *doc = docs + (int64_t)m * dim;
    int32_t *out = out_tokens + (int64_t)m * max_len;
    int cur = 0, len = 1;
    if (lane == 0) out[0] = 0;                              // :233 match_answer = [0]
    while (len < max_len - 1) {
        const int e0 = first_child[cur], e1 = first_child[cur + 1];
        if (e1 == e0) break;
        if (e1 - e0 == 1 && node_leaf_num[child_node[e0]] == 0) break;      // :235 only child without an embedding
        float best = -INFINITY;
        int best_edge = -1;
        for (int r = e0; r < e1; ++r) {                     // :239-246, np.argmax = first maximum in insertion order
            const int e = child_order[r];
            const int ch = child_node[e];
            if (node_leaf_num[ch] == 0) continue;
            const float *ce = node_emb + (int64_t)ch * dim;
            float sim = 0.f;
            for (int d = lane; d < dim; d += 32) sim = fmaf(doc[d], ce[d], sim);
#pragma unroll
            for (int s = 16; s; s >>= 1) sim += __shfl_xor_sync(0xffffffffu, sim, s);
            if (best_edge < 0 || sim > best) { best = sim; best_edge = e; }
        }
        if (best_edge < 0) break;
        if (lane == 0) out[len] = child_tok[best_edge];
        ++len;
        cur = child_node[best_edge];
    }
    if (lane == 0) {
        out[len] = 1;                                       // :250 EOS
        out_len[m] = len + 1;
    }
}

cudaError_t launch_node_embeddings(const int32_t *first_child, const int32_t *child_node, const int32_t *child_order, const int32_t *node_cluster,
                                   const float *leaf_emb, const int32_t *leaf_num, int dim, const int *level_start, int n_levels,
                                   float *node_emb, int32_t *node_leaf_num, cudaStream_t s) {
    for (int l = n_levels - 1; l >= 0; --l) {               // deepest level first; a level only reads deeper ones
        const int lo = level_start[l], hi = level_start[l + 1];
        if (hi <= lo) continue;
        k_node_embeddings<<<(hi - lo + 3) / 4, 128, 0, s>>>(first_child, child_node, child_order, node_cluster, leaf_emb, leaf_num, dim, lo, hi,
                                                            node_emb, node_leaf_num);
    }
    return cudaGetLastError();
}

cudaError_t launch_tree_match(const int32_t *first_child, const int32_t *child_tok, const int32_t *child_node, const int32_t *child_order,
                              const float *node_emb, const int32_t *node_leaf_num, int dim, const float *docs, int M, int max_len,
                              int32_t *out_tokens, int32_t *out_len, cudaStream_t s) {
    if (M == 0) return cudaSuccess;
    k_tree_match<<<(M + 3) / 4, 128, 0, s>>>(first_child, child_tok, child_node, child_order, node_emb, node_leaf_num, dim, docs, M, max_len,
                                             out_tokens, out_len);
    return cudaGetLastError();
}

}  // namespace gdr
