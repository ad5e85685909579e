// k_topk_fast_grouped — EXPERIMENT (off unless GDR_TOPK_GROUPS=G is set when the store is created; written after the GPU budget
// of round 1 was spent: compiled, NOT yet run on a GPU; first run: tests/test_gpu_zz_experimental.py).  G independent 128-thread
// groups per CTA (GroupScope: own named barrier, own shared-memory slice), each claiming one query at a time from a global
// counter and running the unchanged topk_fast16 on it (topk_group_loop, topk_select.cuh).  Stand-alone it is the check that the
// group-scoped select equals k_topk_fast; the same loop is warps 10+ of the fused scoring + top-k CTA (score_fused.cu).
// In a translation unit of its own so that the product kernels of topk.cu compile exactly as they did without it.
#include "gdr_common.cuh"
#include "topk_select.cuh"

namespace gdr {

template <int G>
__global__ void __launch_bounds__(TKF_THREADS * G, 8 / G) k_topk_fast_grouped(ScoreArgs a, float alpha, float *out_scores, int32_t *out_docids) {
    extern __shared__ __align__(16) unsigned char smem[];
    pdl_wait();
    topk_group_loop<0>(a, alpha, out_scores, out_docids, smem + (size_t)GroupScope<0>::group() * tkg_slice_bytes(a.K));
    __syncthreads();                                               // every group of this CTA has made its last claim
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&a.counters[CTR_TOPK_DONE], 1) == (int)gridDim.x - 1) {      // last CTA: leave the queue ready for the next launch
            a.counters[CTR_TOPK_NEXT] = 0;
            a.counters[CTR_TOPK_DONE] = 0;
        }
    }
    pdl_launch_dependents();
}

cudaError_t launch_topk_grouped(const ScoreArgs &a, float alpha, float *out_scores, int32_t *out_docids, cudaStream_t s, int groups) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = groups >= 4 ? 2 : (groups == 2 ? 4 : 8);          // eight groups per SM in every shape
    const int ctas = min((a.B_top + groups - 1) / groups, sms * per_sm);
    const size_t smem = (size_t)groups * tkg_slice_bytes(a.K);
    if (groups >= 4) return launch_pdl(k_topk_fast_grouped<4>, dim3(ctas), dim3(TKF_THREADS * 4), smem, s, a.launch_prio, a, alpha, out_scores, out_docids);
    if (groups >= 2) return launch_pdl(k_topk_fast_grouped<2>, dim3(ctas), dim3(TKF_THREADS * 2), smem, s, a.launch_prio, a, alpha, out_scores, out_docids);
    return launch_pdl(k_topk_fast_grouped<1>, dim3(ctas), dim3(TKF_THREADS), smem, s, a.launch_prio, a, alpha, out_scores, out_docids);
}

}  // namespace gdr
