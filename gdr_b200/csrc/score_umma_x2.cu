// The tcgen05 scoring CTA of score_umma.cu with a 4-stage ring (112 KB) so that TWO CTAs are resident per SM (GDR_OPT_UMMA_CTAS_PER_SM = 2).
//
// Why a second instantiation: a scoring launch costs ~10 us that do not scale with its tiles — CTA prologue (barrier init, TMEM
// allocation), the first tile's claim + metadata copy + first TMA / B fill (~3 us to the first MMA), the last-tile spread and the
// tear-down — and with one 174 KB CTA per SM the next launch's CTAs cannot become resident before the previous launch's leave, so
// those phases are exposed between launches (measured with tools/probe_umma_limits.py: the barrier / epilogue skeleton alone, no
// TMA, no MMA, no B fill, takes 26 of the kernel's 35 us on 148 SMs and 10.4 us + 2.2 us per tile per CTA across CTA counts).  Two
// co-resident CTAs with half the ring each keep the same bytes in flight per SM (8 stages) and let one CTA's fill / drain / tile
// boundary overlap the other's steady state, across launches too.  On the whole device that buys nothing (35.5 vs 34.8 us: HBM-
// bound there); confined to 84 - 100 SMs by the partitioned schedule (gdr_b200/pipeline.py) it is worth 3 - 7 us per 1,024-query
// step (cfg2: 50.6 -> 44.3 us with 56 SMs on the top-k side).  Same body, same results bit for bit (tests/test_gpu_pipeline.py).
#define GDR_UM_STAGES 4
#define GDR_UM_MIN_CTAS 2
#include "gdr_common.cuh"
#include "score_umma.cuh"

namespace gdr {

static_assert(UM_SA == 4 && UM_FILL_WARPS == 2 && UM_THREADS == 288, "this translation unit is the 4-stage variant");
static_assert(2 * (UM_SMEM_BYTES + 127) / 128 * 128 + 2 * 1024 <= 228 * 1024, "two CTAs per SM (1 KB reserved per CTA, 128-byte allocation granularity)");

#define UM_FILL_IDX (warp - 2)
#define UM_DISPATCH                                       \
    if (warp == 0) role_tma();                            \
    else if (warp == 1) role_mma();                       \
    else if (warp < 2 + UM_FILL_WARPS) role_fill();       \
    else if (warp < 2 + UM_FILL_WARPS + 4) role_epi();    \
    else role_sched();
#define UM_EXTRA_TAIL
#define UM_SCORE_PTR(o) (a.scorebuf + (o))
#define UM_P2P_FENCE
#define UM_P2P_SIGNAL
#define UM_P2P_WAIT_CONSUMED
__global__ void __launch_bounds__(UM_THREADS, 2) k_score_umma_x2(const __grid_constant__ CUtensorMap tmap, ScoreArgs a) {
#include "score_umma_body.inc"
}
#undef UM_SCORE_PTR
#undef UM_P2P_FENCE
#undef UM_P2P_SIGNAL
#undef UM_P2P_WAIT_CONSUMED
#undef UM_FILL_IDX
#undef UM_DISPATCH
#undef UM_EXTRA_TAIL

cudaError_t launch_score_umma_x2(const ScoreArgs &a, const CUtensorMap *tmap, cudaStream_t s, int ctas) {
    static FuncAttrOnce attr;
    cudaError_t e = attr.ensure([] {
        cudaError_t e2 = cudaFuncSetAttribute(k_score_umma_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_BYTES);
        // two CTAs per SM need the whole shared-memory carve-out
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_score_umma_x2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return e2;
    });
    if (e != cudaSuccess) return e;
    return launch_pdl(k_score_umma_x2, dim3(ctas), dim3(UM_THREADS), UM_SMEM_BYTES, s, a.launch_prio, *tmap, a);
}

}  // namespace gdr
