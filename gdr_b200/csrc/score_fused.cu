// k_score_topk_fused — scoring of batch i and the per-query top-k of batch i-1 in ONE persistent CTA per SM (gdr_score_fused).
//
// Why: alone, the scoring kernel takes 34.5 us and the top-k 17.5 us per 1,024-query batch, but as separate grids they cost
// 48-50 us per pipelined step, and the measured reason is RESIDENCY — the top-k's ~1,000 small CTAs occupy SMs that the next
// batch's scoring CTA (174 KB of shared memory, 32 K registers) then has to wait for, and no launch-level knob reserves room
// (ROADMAP.md).  So the top-k gets a fixed home inside the scoring CTA, one batch behind:
//
//   launch i:  scoring warps   the tcgen05 scoring CTA of batch i (score_umma_body.inc: TMA, MMA, 3 B fillers, 4 epilogue
//                              warps, tile scheduler; dynamic tile queue)
//              top-k warps     groups (topk_group_loop): each claims one query of batch i-1 at a time from a global counter and
//                              runs topk_fast16 on that query's row of batch i-1's score buffer (complete and L2-resident:
//                              launch i-1 wrote it, and stream order / griddepcontrol.wait separates the launches)
//
// Every wait in the kernel is on an mbarrier or a named barrier fed by the CTA's own warps — no flags, no spinning on other
// CTAs' progress.  Batches i and i-1 live in two different store handles (= two scratch sets); the handles' last ScoreArgs are
// kept by gdr_score_topk (called with GDR_SKIP_SCORE | GDR_SKIP_TOPK for the inversion) and passed in here.
//
// Two variants:
//   k_score_topk_fused<4>   (round 1) four 128-thread groups after the ten scoring warps, one uniform register allocation
//                           (72 per thread, 52 bytes of spills in the epilogue).  560 queries in flight on 140 CTAs: a batch of
//                           1,024 needs TWO rounds of ~24 us (a query's select is a chain of dependent L2 round trips and
//                           barriers) — measured 47.7 us per step, i.e. the top-k, not the scoring, sets the step time.
//   k_score_topk_fused64    (round 2) nine 64-thread groups (two warps per query): 1,260 queries in flight on 140 CTAs, so the
//                           whole previous batch is selected in ONE round that only has to finish within the scoring time.
//                           Roles are laid out in homogeneous warpgroups — {TMA, MMA, scheduler, filler 0}, {4 epilogue warps},
//                           {fillers 1-2, top-k group 0}, {top-k groups 1-2}, ... — and `setmaxnreg` moves registers from the
//                           light roles (64) to the epilogue warpgroup (120), which removes the epilogue's spills.
#include "gdr_common.cuh"
#include "score_umma.cuh"
#include "topk_select.cuh"

namespace gdr {

constexpr int FU_MAX_SMEM = 227 * 1024;

// ------------------------------------------------------------------------------------------------------------------------
// variant A (default): five 128-thread top-k groups running the lean select (topk_lean128), homogeneous warpgroups + setmaxnreg
//   warp 0 TMA | 1 MMA | 2 tile scheduler | 3 filler 0 || 4-7 epilogue || 8-9 fillers 1-2 | 10-11 idle || 12-31 top-k groups 0-4
//   1,024 threads = the launch holds 64 registers per thread; the two light scoring warpgroups drop to 48, the epilogue rises to
//   96, the five top-k warpgroups keep their 64
// ------------------------------------------------------------------------------------------------------------------------
constexpr int F128_GROUPS = 5;
constexpr int F128_FIRST = 12 * 32;                                   // first top-k thread
constexpr int F128_THREADS = F128_FIRST + F128_GROUPS * TKF_THREADS;  // 1,024 = eight warpgroups
constexpr int F128_REGS_LIGHT = 48, F128_REGS_TOPK = 64, F128_REGS_EPI = 96;    // warpgroups 0 and 2 | 3-7 | 1
static_assert(F128_THREADS == 1024, "eight whole warpgroups");
static_assert(2 * 128 * F128_REGS_LIGHT + F128_GROUPS * 128 * F128_REGS_TOPK + 128 * F128_REGS_EPI <= 65536, "register file");

#define UM_FILL_IDX (warp == 3 ? 0 : warp - 7)
// setmaxnreg is warpgroup-wide: it sits at the head of each WARPGROUP's branch (all four warps of a warpgroup reach the same
// instruction), and the roles of a warpgroup are dispatched inside that branch, so that ptxas sizes each role's registers by it
#define UM_DISPATCH                                                                                                                 \
    if ((warp >> 2) == 1) {                                                                                                         \
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F128_REGS_EPI));                                                   \
        role_epi();                                                                                                                 \
    } else if ((warp >> 2) >= 3) {                                                                                                  \
        /* the top-k warpgroups keep the launch allocation (F128_REGS_TOPK = 64) */                                                 \
        if (prev.B > 0)                                                                                                             \
            topk_group_loop<F128_FIRST, TKF_THREADS>(prev, alpha, out_scores, out_docids,                                           \
                smem + UM_SMEM_BYTES + (size_t)GroupScope<F128_FIRST, TKF_THREADS>::group() * tkg_slice_bytes(prev.K));             \
    } else {                                                                                                                        \
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F128_REGS_LIGHT));                                                 \
        if (warp == 0) role_tma();                                                                                                  \
        else if (warp == 1) role_mma();                                                                                             \
        else if (warp == 2) role_sched();                                                                                           \
        else if (warp == 3 || warp == 8 || warp == 9) role_fill();                                                                  \
    }
// after the CTA-wide barrier every group of this CTA has made its last claim: the last CTA leaves the query queue ready
#define UM_EXTRA_TAIL                                                                                                               \
    if (threadIdx.x == 32 && prev.B > 0 && atomicAdd(&prev.counters[CTR_TOPK_DONE], 1) == (int)gridDim.x - 1) {                     \
        prev.counters[CTR_TOPK_NEXT] = 0;                                                                                           \
        prev.counters[CTR_TOPK_DONE] = 0;                                                                                           \
    }

#define UM_SCORE_PTR(o) (a.scorebuf + (o))
#define UM_P2P_FENCE
#define UM_P2P_SIGNAL
#define UM_P2P_WAIT_CONSUMED
__global__ void __launch_bounds__(F128_THREADS, 1)
k_score_topk_fused(const __grid_constant__ CUtensorMap tmap, ScoreArgs a, ScoreArgs prev, float alpha, float *out_scores, int32_t *out_docids) {
#include "score_umma_body.inc"
}
#undef UM_SCORE_PTR
#undef UM_P2P_FENCE
#undef UM_P2P_SIGNAL
#undef UM_P2P_WAIT_CONSUMED
// ... and for one shard of a cluster-sharded corpus (scores into the owners' buffers over NVLink, arrival flags; the top-k groups
// wait for every rank's flag before they read the previous batch's scores: topk_group_loop)
#define UM_SCORE_PTR(o) score_ptr(a, (o))
#define UM_P2P_FENCE __threadfence_system();
#define UM_P2P_SIGNAL { signal_owners(a); if (prev.B > 0) signal_consumed(prev); }
#define UM_P2P_WAIT_CONSUMED { if (lane == 0) wait_consumed(a); __syncwarp(); }
__global__ void __launch_bounds__(F128_THREADS, 1)
k_score_topk_fused_p2p(const __grid_constant__ CUtensorMap tmap, ScoreArgs a, ScoreArgs prev, float alpha, float *out_scores, int32_t *out_docids) {
#include "score_umma_body.inc"
}
#undef UM_SCORE_PTR
#undef UM_P2P_FENCE
#undef UM_P2P_SIGNAL
#undef UM_P2P_WAIT_CONSUMED
#undef UM_FILL_IDX
#undef UM_DISPATCH

// ------------------------------------------------------------------------------------------------------------------------
// variant B: nine 64-thread top-k groups (two warps per query, no keys in registers)
//   warp 0 TMA | 1 MMA | 2 tile scheduler | 3 filler 0 || 4-7 epilogue || 8-9 fillers 1-2 | 10-27 top-k groups 0-8 (two warps each)
// ------------------------------------------------------------------------------------------------------------------------
constexpr int F64_GROUPS = 9;
constexpr int F64_FIRST = 10 * 32;                                   // first top-k thread
constexpr int F64_THREADS = F64_FIRST + F64_GROUPS * TKF64_THREADS;  // 896 = seven warpgroups
constexpr int F64_REGS_SMALL = 64, F64_REGS_EPI = 120;               // launch: 896 x 72 = 64,512; then 768 x 64 + 128 x 120 = 64,512
static_assert(F64_THREADS % 128 == 0, "setmaxnreg is a warpgroup-wide instruction: no partial warpgroup");
static_assert((F64_THREADS - 128) * F64_REGS_SMALL + 128 * F64_REGS_EPI <= F64_THREADS * 72, "registers the launch allocation holds");

#define UM_FILL_IDX (warp == 3 ? 0 : warp - 7)
// warpgroup 0 = {TMA, MMA, scheduler, filler 0}, 1 = epilogue, 2 = {fillers 1-2, top-k group 0}, 3-6 = top-k groups 1-8
#define UM_DISPATCH                                                                                                                 \
    if ((warp >> 2) == 1) {                                                                                                         \
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F64_REGS_EPI));                                                    \
        role_epi();                                                                                                                 \
    } else {                                                                                                                        \
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(F64_REGS_SMALL));                                                  \
        if (warp == 0) role_tma();                                                                                                  \
        else if (warp == 1) role_mma();                                                                                             \
        else if (warp == 2) role_sched();                                                                                           \
        else if (warp == 3 || warp == 8 || warp == 9) role_fill();                                                                  \
        else if (prev.B > 0)                                                                                                        \
            topk_group_loop<F64_FIRST, TKF64_THREADS>(prev, alpha, out_scores, out_docids,                                          \
                smem + UM_SMEM_BYTES + (size_t)GroupScope<F64_FIRST, TKF64_THREADS>::group() * tkg_slice_bytes(prev.K));            \
    }
#define UM_SCORE_PTR(o) (a.scorebuf + (o))
#define UM_P2P_FENCE
#define UM_P2P_SIGNAL
#define UM_P2P_WAIT_CONSUMED
__global__ void __launch_bounds__(F64_THREADS, 1)
k_score_topk_fused64(const __grid_constant__ CUtensorMap tmap, ScoreArgs a, ScoreArgs prev, float alpha, float *out_scores, int32_t *out_docids) {
#include "score_umma_body.inc"
}
#undef UM_SCORE_PTR
#undef UM_P2P_FENCE
#undef UM_P2P_SIGNAL
#undef UM_P2P_WAIT_CONSUMED
// ... and for one shard of a cluster-sharded corpus (scores into the owners' buffers over NVLink, arrival flags; the top-k groups
// wait for every rank's flag before they read the previous batch's scores: topk_group_loop)
#define UM_SCORE_PTR(o) score_ptr(a, (o))
#define UM_P2P_FENCE __threadfence_system();
#define UM_P2P_SIGNAL { signal_owners(a); if (prev.B > 0) signal_consumed(prev); }
#define UM_P2P_WAIT_CONSUMED { if (lane == 0) wait_consumed(a); __syncwarp(); }
__global__ void __launch_bounds__(F64_THREADS, 1)
k_score_topk_fused64_p2p(const __grid_constant__ CUtensorMap tmap, ScoreArgs a, ScoreArgs prev, float alpha, float *out_scores, int32_t *out_docids) {
#include "score_umma_body.inc"
}
#undef UM_SCORE_PTR
#undef UM_P2P_FENCE
#undef UM_P2P_SIGNAL
#undef UM_P2P_WAIT_CONSUMED
#undef UM_FILL_IDX
#undef UM_DISPATCH
#undef UM_EXTRA_TAIL

// How many top-k groups of the 64-thread variant fit beside the scoring ring for beam width K (0 = it does not fit at all)
int fused64_groups_that_fit(int K) {
    const int room = FU_MAX_SMEM - UM_SMEM_BYTES;
    const int g = room / tkg_slice_bytes(K);
    return g >= F64_GROUPS ? F64_GROUPS : 0;
}

// prev.B == 0: no previous batch (first launch of a stream) — the top-k warps fall straight through to the final barrier.
// groups: 9 = variant B (nine 64-thread groups), anything else = variant A (five 128-thread groups, the default).
cudaError_t launch_score_fused(const ScoreArgs &a, const CUtensorMap *tmap, const ScoreArgs &prev, float alpha, float *out_scores,
                               int32_t *out_docids, cudaStream_t s, int ctas, int groups) {
    const int K = prev.B > 0 ? prev.K : 1;
    const bool p2p = a.n_ranks > 1;       // a shard of a cluster-sharded corpus
    const bool v64 = groups == F64_GROUPS && fused64_groups_that_fit(K) == F64_GROUPS;
    const int g = v64 ? F64_GROUPS : F128_GROUPS;
    const size_t smem = (size_t)UM_SMEM_BYTES + (size_t)g * tkg_slice_bytes(K);
    if (smem > (size_t)FU_MAX_SMEM) return cudaErrorInvalidValue;
    static FuncAttrOnce attr;
    cudaError_t e = attr.ensure([] {
        cudaError_t e2 = cudaFuncSetAttribute(k_score_topk_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_MAX_SMEM);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_score_topk_fused_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_MAX_SMEM);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_score_topk_fused64, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_MAX_SMEM);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_score_topk_fused64_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_MAX_SMEM);
        return e2;
    });
    if (e != cudaSuccess) return e;
    auto kernel = v64 ? (p2p ? k_score_topk_fused64_p2p : k_score_topk_fused64) : (p2p ? k_score_topk_fused_p2p : k_score_topk_fused);
    return launch_pdl(kernel, dim3(ctas), dim3(v64 ? F64_THREADS : F128_THREADS), smem, s, a.launch_prio, *tmap, a, prev, alpha, out_scores, out_docids);
}

}  // namespace gdr
