// k_score_topk_fused — EXPERIMENT (ROADMAP.md, "plan of record" for round 2).  Written after the GPU budget of round 1 was
// spent: it compiles for sm_100a and has NOT run on a GPU yet; nothing in the default path (gdr_score_topk) uses it, its first
// run is tests/test_gpu_zz_experimental.py (child process, xfail-tolerant).
//
// Why: alone, the scoring kernel takes 34.5 us and the top-k 17.5 us per 1,024-query batch, but pipelined they cost 48-50 us
// per step, and the measured reason is RESIDENCY — the top-k's ~1,000 small CTAs occupy SMs that the next batch's scoring CTA
// (174 KB of shared memory, 32 K registers) then has to wait for, and no launch-level knob reserves room (ROADMAP.md).  So the
// top-k gets a fixed home inside the scoring CTA, one batch behind:
//
//   launch i:  warps 0-9   the tcgen05 scoring CTA of batch i, unchanged (score_umma_body.inc: TMA, MMA, 3 B fillers,
//                          4 epilogue warps, tile scheduler; dynamic tile queue)
//              warps 10+   G top-k groups of 128 threads (topk_group_loop): each claims one query of batch i-1 at a time from a
//                          global counter and runs topk_fast16 on that query's row of batch i-1's score buffer (complete and
//                          L2-resident: launch i-1 wrote it, and stream order / griddepcontrol.wait separates the launches)
//
// Every wait in the kernel is on an mbarrier or a named barrier fed by the CTA's own warps — no flags, no spinning on other
// CTAs' progress.  Batches i and i-1 live in two different store handles (= two scratch sets); the handles' last ScoreArgs are
// kept by gdr_score_topk (called with GDR_SKIP_SCORE | GDR_SKIP_TOPK for the inversion) and passed in here.
//
// Budget per SM (one CTA): threads 320 + 128 G; shared memory 169 KB + G x 6.3 KB (K = 20); registers 65,536 / threads —
// G = 3: 80 used, G = 4: 72, G = 5 (960 threads, the most a CTA can hold next to the ten scoring warps): 64 — against 100 in
// k_score_umma; ptxas reports 48-68 bytes of spills in all three, so the uniform allocation is a fair first draft until the
// roles are regrouped into homogeneous warpgroups and rebalanced with setmaxnreg (ROADMAP.md).
#include "gdr_common.cuh"
#include "score_umma.cuh"
#include "topk_select.cuh"

namespace gdr {

constexpr int FU_MAX_SMEM = 227 * 1024;

#define UM_SCHED_ELSE else if (warp == 2 + UM_FILL_WARPS + 4)
#define UM_EXTRA_ROLES                                                                                                              \
    else if (prev.B > 0) {                                                                                                          \
        topk_group_loop<UM_THREADS>(prev, alpha, out_scores, out_docids,                                                            \
                                    smem + UM_SMEM_BYTES + (size_t)GroupScope<UM_THREADS>::group() * tkg_slice_bytes(prev.K));      \
    }
// after the CTA-wide barrier every group of this CTA has made its last claim: the last CTA leaves the query queue ready
#define UM_EXTRA_TAIL                                                                                                               \
    if (threadIdx.x == 32 && prev.B > 0 && atomicAdd(&prev.counters[CTR_TOPK_DONE], 1) == (int)gridDim.x - 1) {                     \
        prev.counters[CTR_TOPK_NEXT] = 0;                                                                                           \
        prev.counters[CTR_TOPK_DONE] = 0;                                                                                           \
    }

template <int G>
__global__ void __launch_bounds__(UM_THREADS + G * TKF_THREADS, 1)
k_score_topk_fused(const __grid_constant__ CUtensorMap tmap, ScoreArgs a, ScoreArgs prev, float alpha, float *out_scores, int32_t *out_docids) {
#include "score_umma_body.inc"
}
#undef UM_SCHED_ELSE
#undef UM_EXTRA_ROLES
#undef UM_EXTRA_TAIL

// prev.B == 0: no previous batch (first launch of a stream) — the top-k warps fall straight through to the final barrier.
cudaError_t launch_score_fused(const ScoreArgs &a, const CUtensorMap *tmap, const ScoreArgs &prev, float alpha, float *out_scores,
                               int32_t *out_docids, cudaStream_t s, int ctas, int groups) {
    const size_t smem = (size_t)UM_SMEM_BYTES + (size_t)groups * tkg_slice_bytes(prev.B > 0 ? prev.K : 1);
    if (smem > (size_t)FU_MAX_SMEM) return cudaErrorInvalidValue;
    static unsigned long long attr_set_mask = 0;      // one bit per device: the attribute is per device and function
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set_mask >> (dev & 63)) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k_score_topk_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_MAX_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_score_topk_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_MAX_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_score_topk_fused<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_MAX_SMEM);
        if (e != cudaSuccess) return e;
        attr_set_mask |= 1ull << (dev & 63);
    }
    if (groups == 5)
        return launch_pdl(k_score_topk_fused<5>, dim3(ctas), dim3(UM_THREADS + 5 * TKF_THREADS), smem, s, *tmap, a, prev, alpha, out_scores, out_docids);
    if (groups == 3)
        return launch_pdl(k_score_topk_fused<3>, dim3(ctas), dim3(UM_THREADS + 3 * TKF_THREADS), smem, s, *tmap, a, prev, alpha, out_scores, out_docids);
    return launch_pdl(k_score_topk_fused<4>, dim3(ctas), dim3(UM_THREADS + 4 * TKF_THREADS), smem, s, *tmap, a, prev, alpha, out_scores, out_docids);
}

}  // namespace gdr
