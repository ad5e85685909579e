// Constants and PTX wrappers of the tcgen05 scoring CTA, shared by k_score_umma (score_umma.cu) and the fused scoring + top-k
// kernel (score_fused.cu).  The CTA body itself is score_umma_body.inc, included into each kernel.
#pragma once
#include "gdr_common.cuh"

namespace gdr {

constexpr int UM_BLOCK_K = 64;                 // bf16 elements per K block = one 128-byte swizzle atom
#ifndef GDR_UM_STAGES
#define GDR_UM_STAGES 6                        // score_umma_x2.cu defines 4 (two CTAs per SM); GDR_BUILD_UM_STAGES=8 in developer builds
#endif
constexpr int UM_SA = GDR_UM_STAGES;           // ONE ring of 6 stages, each = A tile (TMA) + B tile (cp.async): one full and one empty
constexpr int UM_SB = GDR_UM_STAGES;           // barrier per stage, so the MMA warp pays one wait + one commit per K block.  Every stage
                                               // is owned by exactly one filler warp, which keeps each waiter at most one mbarrier phase
                                               // ahead (parity waits stay unambiguous).  6 x 28 KB = 168 KB leaves ~58 KB of the SM for
                                               // co-resident top-k / inversion CTAs of neighbouring batches (8 stages: same speed alone; confined to 100 / 92 / 84 SMs by
                                               // the partitioned schedule, 8 stages gain 1 us of a 48 - 53 us step: the CTA is not bound by bytes in flight).
constexpr int UM_FILL_WARPS = UM_SB / 2;       // each filler warp owns two stages (one cp.async group in flight per stage)
constexpr int UM_A_BYTES = UMMA_ROWS * 128;    // 16 KB
constexpr int UM_BT_BYTES = UMMA_NQ * 128;     // 4 KB per query term
constexpr int UM_B_BYTES = 3 * UM_BT_BYTES;    // 12 KB
constexpr int UM_THREADS = 64 + 32 * UM_FILL_WARPS + 128 + 32;   // 320
constexpr int UM_ACC_COLS = 128;               // TMEM columns reserved per accumulator (3 * 32 used)
constexpr int UM_TMEM_COLS = 2 * UM_ACC_COLS;  // double-buffered accumulator
constexpr int UM_RING_BYTES = UM_SA * UM_A_BYTES + UM_SB * UM_B_BYTES;   // 168 KB: leaves room for co-resident top-k / inversion CTAs
static_assert(UM_SA == UM_SB, "A and B share one ring");
constexpr int UM_MD = 2;                       // tile-metadata ring depth = how far ahead of its slowest role a CTA claims tiles
constexpr int UM_META_CONSUMERS = 2 + UM_FILL_WARPS + 4;    // TMA, MMA, fillers, epilogue warps
constexpr int UM_BAR_BYTES = GDR_UM_STAGES <= 4 ? 272 : 512;          // mbarriers (8 bytes each) from 0, the TMEM base slot at 256
constexpr int UM_SMEM_BYTES = UM_RING_BYTES + UM_BAR_BYTES + UM_MD * (int)sizeof(TileMeta);
static_assert(UM_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(sizeof(TileMeta) % 16 == 0 && UM_BAR_BYTES % 16 == 0 && UM_RING_BYTES % 16 == 0, "bulk-copy alignment of the metadata ring");
static_assert(UM_SB == 2 * UM_FILL_WARPS, "two B stages per filler warp");
static_assert(UMMA_NQ == 32, "epilogue and filler lane maps assume 32 pairs per tile");

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {      // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// same load with an L2 eviction-priority hint: the store is streamed once per batch (and is larger than L2), so its lines
// are marked evict-first and do not push the score buffer, the split-query table and the work lists out of L2
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap *tmap, int c0, int c1, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// One lane of a converged warp; the guarded code stays warp-uniform for ptxas, so descriptors and barrier
// addresses are computed on the uniform datapath instead of per-lane registers + R2UR moves.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (sm_100 format): rows are 128 B apart,
// 8-row swizzle atoms 1024 B apart (SBO), LBO unused for swizzled K-major (encoded 1), version 1.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N runtime
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(UMMA_ROWS >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// the grouped GEMM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// timeline trace (GDR_UMMA_TRACE=1): CTA 0's MMA warp stamps slot i with the global timer
#define UM_TRACE(i) do { if (a.dbg && blockIdx.x == 0 && lane == 0 && (i) < 200) a.dbg[(i)] = gtime(); } while (0)

}  // namespace gdr
