// Training-time gather + contrastive loss.
//
// Replaces the candidate gather of T5FineTuner.forward (main_models.py:983-996: doc_embed[index] rows concatenated one at a
// time) and encoder_cal (main_models.py:1184-1221) with ONE kernel over the cluster store: the positives and the in-cluster
// candidates are read straight from their store rows (no [B + S, D] copy), one CTA per query computes
//     s_j = f(q . d_j)            f = tanh | sigmoid            (j = its positive, then all S candidates)
//     loss_i = -log exp(s_pos / tau) + log( intra_rate * sum_{own candidates} exp(s_j / tau) + sum_{other} exp(s_j / tau) )
// and, when asked, d loss / d q_i = (1/B) sum_j g_j f'(x_j) d_j with g_pos = -1/tau, g_j = w_j exp(s_j / tau) / (tau * denom)
// — the only gradient the reference needs (the document embeddings are a fixed table, the query encoder trains).
// The batch loss is the mean of the per-query losses, summed in query order by a second one-warp kernel (deterministic).
#include "gdr_common.cuh"

namespace gdr {

constexpr int CL_THREADS = 256;

template <typename T> __device__ __forceinline__ float row_elem(const void *emb, int64_t row, int dim, int d);
template <> __device__ __forceinline__ float row_elem<float>(const void *emb, int64_t row, int dim, int d) {
    return reinterpret_cast<const float *>(emb)[row * dim + d];
}
template <> __device__ __forceinline__ float row_elem<__nv_bfloat16>(const void *emb, int64_t row, int dim, int d) {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(emb)[row * dim + d]);
}

// dynamic shared memory: s[S + 1] | c[S + 1]  (activated similarities, then gradient coefficients)
template <typename T>
__global__ void __launch_bounds__(CL_THREADS) k_contrastive(const void *__restrict__ emb, int dim, const float *__restrict__ q,
                                                            const int32_t *__restrict__ pos_rows, const int32_t *__restrict__ cand_rows,
                                                            const int32_t *__restrict__ cand_off, int B, int S, int act, float tau,
                                                            float intra_rate, float *loss_per_query, float *grad_q) {
    extern __shared__ float cl_smem[];
    float *s = cl_smem, *coef = cl_smem + (S + 1);
    __shared__ float red[CL_THREADS / 32];
    __shared__ float denom_sh;
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *qi = q + (int64_t)i * dim;
    const int own_lo = cand_off[i], own_hi = cand_off[i + 1];
    // similarities: one warp per document (slot 0 = the positive, slot 1 + j = candidate j)
    for (int j = warp; j <= S; j += CL_THREADS / 32) {
        const int64_t row = j == 0 ? pos_rows[i] : cand_rows[j - 1];
        float x = 0.f;
        for (int d = lane; d < dim; d += 32) x = fmaf(qi[d], row_elem<T>(emb, row, dim, d), x);
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) {
            const float sv = apply_act(x, act);
            s[j] = sv;
            coef[j] = act == GDR_ACT_TANH ? 1.f - sv * sv : (act == GDR_ACT_SIGMOID ? sv * (1.f - sv) : 1.f);     // f'(x)
        }
    }
    __syncthreads();
    // denominator (:1197-1203 / :1217): exp(s / tau), the query's own candidates weighted by intra_rate
    float part = 0.f;
    for (int j = tid; j < S; j += CL_THREADS) {
        const float e = expf(s[1 + j] / tau);
        part += (j >= own_lo && j < own_hi) ? intra_rate * e : e;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
        float denom = 0.f;
        for (int w = 0; w < CL_THREADS / 32; ++w) denom += red[w];
        denom_sh = denom;
        loss_per_query[i] = -logf(expf(s[0] / tau)) + logf(denom);            // :1202-1203, the reference's own form
    }
    __syncthreads();
    if (!grad_q) return;
    const float denom = denom_sh, inv_b = 1.f / (float)B;
    for (int j = tid; j <= S; j += CL_THREADS) {
        float g;
        if (j == 0) g = -1.f / tau;
        else {
            const float e = expf(s[j] / tau);
            g = ((j - 1 >= own_lo && j - 1 < own_hi) ? intra_rate * e : e) / (tau * denom);
        }
        coef[j] = g * coef[j] * inv_b;
    }
    __syncthreads();
    for (int d = tid; d < dim; d += CL_THREADS) {             // thread per dimension: coalesced reads of every row
        float acc = 0.f;
        for (int j = 0; j <= S; ++j) {
            const int64_t row = j == 0 ? pos_rows[i] : cand_rows[j - 1];
            acc = fmaf(coef[j], row_elem<T>(emb, row, dim, d), acc);
        }
        grad_q[(int64_t)i * dim + d] = acc;
    }
}

__global__ void k_mean_in_order(const float *__restrict__ v, int n, float *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc += v[i];
        *out = acc / (float)n;
    }
}

cudaError_t launch_contrastive(const void *emb, int dtype, int dim, const float *q, const int32_t *pos_rows, const int32_t *cand_rows,
                               const int32_t *cand_off, int B, int S, int act, float tau, float intra_rate, float *loss_per_query,
                               float *loss, float *grad_q, cudaStream_t st) {
    const size_t smem = 2 * (size_t)(S + 1) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (dtype == GDR_DTYPE_BF16) {
        cudaFuncSetAttribute(k_contrastive<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        k_contrastive<__nv_bfloat16><<<B, CL_THREADS, smem, st>>>(emb, dim, q, pos_rows, cand_rows, cand_off, B, S, act, tau, intra_rate,
                                                                 loss_per_query, grad_q);
    } else {
        cudaFuncSetAttribute(k_contrastive<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        k_contrastive<float><<<B, CL_THREADS, smem, st>>>(emb, dim, q, pos_rows, cand_rows, cand_off, B, S, act, tau, intra_rate, loss_per_query,
                                                         grad_q);
    }
    k_mean_in_order<<<1, 32, 0, st>>>(loss_per_query, B, loss);
    return cudaGetLastError();
}

}  // namespace gdr
