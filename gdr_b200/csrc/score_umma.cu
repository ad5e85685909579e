// Gather-and-score, tensor-core path: TMA-fed tcgen05 grouped GEMM with TMEM accumulators.
//
// Replaces main_models.py:1456-1462 + 1577-1582 for clusters that many queries of the batch
// selected (a "group" = the (query, beam) pairs that name one cluster).  A GEMV would re-read the
// cluster slab once per pair; here a slab tile is read from HBM ONCE (TMA, 128-byte swizzle) and
// multiplied against all the group's queries on the 5th-gen tensor cores:
//
//     D[128 docs, N pairs] (fp32, TMEM)  +=  A[128 docs, 64 k] (bf16 smem, TMA)  x  B[N pairs, 64 k]^T (bf16 smem)
//
// Parity with the fp32 reference needs more than bf16(q): the fp32 query is split exactly into three
// bf16 terms q = hi + mid + lo (k_count writes them once per batch).  The three terms are stacked along N
// (B rows [0,TS) = hi, [TS,2TS) = mid, [2TS,3TS) = lo, TS = 16 or 32), so ONE MMA of N = 3*TS per K step
// reads the A tile from shared memory once for all three (SS-mode MMAs are bound by the 128 B/clk smem
// read of A, measured: three N=32 MMAs per K step were MIO-throttled); the epilogue adds the three TMEM
// column groups.  Every product bf16 x bf16 is exact in fp32, so the result matches an fp32 dot product to
// accumulation-order noise.  The kernel is still HBM-bound: per tile it moves 128 x 768 x 2 B = 196 KB of
// embeddings and issues 12 x 4 MMAs of 128 x 96 x 16.
//
// Warp roles (320 threads, one persistent CTA per SM, tiles claimed one at a time from a global counter):
//   warp 0       TMA producer: A tiles of the store (L2 evict-first hint), 6-stage ring shared with B (96 KB of A in flight per SM)
//   warp 1       MMA issuer (one lane): tcgen05.mma cta_group::1 kind::f16; commits free the A and B stages
//   warps 2-4    B fillers: gather the group's query rows from the pre-split bf16 table (hi/mid/lo terms written once
//                per batch by k_count, L2-resident) with 16-byte cp.async straight into the 128B-swizzled K-major layout
//                the UMMA descriptor expects; wait_group -> fence.proxy.async -> arrive.  Two stages per warp.
//   warps 5-8    epilogue: tcgen05.ld the accumulator (double-buffered in TMEM so it overlaps the next
//                tile's MMAs), activation, coalesced stores into each query's candidate segment
//   warp 9       tile scheduler (one lane): atomicAdd on the tile counter, then ONE cp.async.bulk of the tile's 272-byte
//                TileMeta record (row range, query rows, score offsets: resolved once per batch by k_tilemeta) into a
//                two-slot shared-memory ring that every other role reads; a record with nq = -1 ends all loops.  No role
//                has a dependent global load on its per-tile critical path (measured: with each role walking item ->
//                pair -> candidate offset itself the empty barrier skeleton alone cost 4 us per tile, as much as the
//                tile's HBM time), and a CTA that runs slower or starts later simply claims fewer tiles.
// The dependent grid (the batch's top-k) is released at the kernel's END (see the note at pdl_launch_dependents below).
#include "gdr_common.cuh"

namespace gdr {

constexpr int UM_BLOCK_K = 64;                 // bf16 elements per K block = one 128-byte swizzle atom
constexpr int UM_SA = 6;                       // ONE ring of 6 stages, each = A tile (TMA) + B tile (cp.async): one full and one empty
constexpr int UM_SB = 6;                       // barrier per stage, so the MMA warp pays one wait + one commit per K block.  Every stage
                                               // is owned by exactly one filler warp, which keeps each waiter at most one mbarrier phase
                                               // ahead (parity waits stay unambiguous).  6 x 28 KB = 168 KB leaves ~58 KB of the SM for
                                               // co-resident top-k / inversion CTAs of neighbouring batches (8 stages: same speed alone).
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
constexpr int UM_BAR_BYTES = 512;
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

__global__ void __launch_bounds__(UM_THREADS, 1) k_score_umma(const __grid_constant__ CUtensorMap tmap, ScoreArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];             // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + UM_RING_BYTES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + UM_RING_BYTES + 256);   // after the 28 barriers
    TileMeta *meta = reinterpret_cast<TileMeta *>(smem + UM_RING_BYTES + UM_BAR_BYTES);

    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (UM_SA + s); };
    auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * UM_SA + i); };
    auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * UM_SA + 2 + i); };
    auto mfull_bar = [&](int i) { return bar_base + 8u * (2 * UM_SA + 4 + i); };
    auto mempty_bar = [&](int i) { return bar_base + 8u * (2 * UM_SA + 4 + UM_MD + i); };
    // consumer side of the metadata ring: wait for tile `it`'s slot, copy what the role needs, release the slot
    auto meta_acquire = [&](int it) -> const TileMeta * {
        mbar_wait(mfull_bar(it % UM_MD), (uint32_t)(it / UM_MD) & 1u);
        return meta + it % UM_MD;
    };
    auto meta_release = [&](int it) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(mempty_bar(it % UM_MD));
    };
    auto a_smem = [&](int s) { return smem_base + (uint32_t)s * UM_A_BYTES; };
    auto b_smem = [&](int s, int t) { return smem_base + UM_SA * UM_A_BYTES + (uint32_t)s * UM_B_BYTES + (uint32_t)t * UM_BT_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1) UM_TRACE(0);
    const int nkb = a.dim / UM_BLOCK_K;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
        for (int s = 0; s < UM_SA; ++s) {
            mbar_init(full_bar(s), 2);         // TMA expect_tx arrive + the filler warp that owns the stage
            mbar_init(empty_bar(s), 1);        // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar(i), 1);        // tcgen05.commit
            mbar_init(tempty_bar(i), 4);       // 4 epilogue warps
        }
        for (int i = 0; i < UM_MD; ++i) {
            mbar_init(mfull_bar(i), 1);        // metadata warp
            mbar_init(mempty_bar(i), UM_META_CONSUMERS);
        }
        if (smem_base & 1023u) __trap();       // the dynamic shared window must start 1024-byte aligned
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(UM_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation) overlapped the inversion kernels; its outputs are read from here on
    pdl_wait();
    trace_start(a.dbg, 2);
    const int n_tiles = (a.flags & (1u << 31)) ? 0 : a.counters[CTR_N_UMMA];

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        int s = 0;
        uint32_t ph = 0;
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        const bool hint = !(a.flags & (1u << 27));
        for (int it = 0;; ++it) {
            const TileMeta *m = meta_acquire(it);
            const int nq = m->nq, row0 = m->row0;
            if (nq < 0) break;
            meta_release(it);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (elect_one()) {
                    if (a.flags & (1u << 28)) mbar_arrive(full_bar(s));
                    else {
                    mbar_arrive_expect_tx(full_bar(s), UM_A_BYTES);
                    if (hint) tma_load_2d_hint(a_smem(s), &tmap, kb * UM_BLOCK_K, row0, full_bar(s), policy);
                    else tma_load_2d(a_smem(s), &tmap, kb * UM_BLOCK_K, row0, full_bar(s));
                    }
                }
                __syncwarp();
                if (++s == UM_SA) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        int sa = 0;
        uint32_t pha = 0;
        UM_TRACE(1);
        if (a.dbg && lane == 0) a.dbg[200 + 2 * blockIdx.x] = gtime();      // every CTA: loop start / end
        int tr = 2;
        for (int it = 0;; ++it) {
            const int nq = meta_acquire(it)->nq;
            if (nq < 0) break;
            meta_release(it);
            UM_TRACE(tr); ++tr;           // metadata of tile `it` in hand
            const uint32_t idesc = umma_idesc(nq <= 16 ? 48 : 96);     // N = 3 terms x TS rows
            const int acc = it & 1;
            const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
            tc_fence_after();
            UM_TRACE(tr); ++tr;           // accumulator free
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * UM_ACC_COLS;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_bar(sa), pha);
                tc_fence_after();
                if (it < 3) { UM_TRACE(tr); ++tr; }   // K block data in hand (first three tiles only)
                if (elect_one()) {
                    if (!(a.flags & (1u << 29))) {
                    const uint64_t adesc = umma_smem_desc(a_smem(sa));
                    const uint64_t bdesc = umma_smem_desc(b_smem(sa, 0));
#pragma unroll
                    for (int k = 0; k < UM_BLOCK_K / 16; ++k) {
                        // +32 bytes per UMMA_K = 16 bf16 inside the swizzle atom: +2 in the (addr >> 4) field
                        umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, k != 0 ? 1u : (kb != 0 ? 1u : 0u));
                    }
                    }
                    umma_commit(empty_bar(sa));
                    if (kb == nkb - 1) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++sa == UM_SA) { sa = 0; pha ^= 1u; }
            }
        }
        if (a.dbg && lane == 0) a.dbg[201 + 2 * blockIdx.x] = gtime();
    } else if (warp < 2 + UM_FILL_WARPS) {
        // ===================== B fillers: warp fw owns stages fw and fw + S/2, i.e. K blocks g = fw, fw + S/2, fw + S, ... =====================
        // Pure copies: the three bf16 terms of every query were written once per batch by k_count (a.qsplit, L2-resident);
        // each lane issues 16-byte cp.async (LDGSTS) straight into the swizzled B tile, commits the group and moves on to
        // its next K block; the PREVIOUS block's group is then complete (wait_group 1), gets its generic->async proxy
        // fence and is published to the MMA warp.  No register staging, no L2 latency on the warp's critical path.
        const int fw = warp - 2;
        const uint32_t b_ring = smem_base + UM_SA * UM_A_BYTES;
        // every filler warp consumes every tile's metadata slot (lane l caches the query row of pair l), including
        // tiles in which it owns no K block, so the slot's consumer count is the same for all tiles
        int cur_it = -1, nq = 0, qrow = 0, prev_g = -1;
        auto publish = [&](int g_done, int pending_groups) {
            if (pending_groups) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(g_done % UM_SB));
        };
        auto advance_to = [&](int it_target) -> bool {      // false once the end-of-work record has been reached
            while (cur_it < it_target) {
                if (nq < 0) return false;
                ++cur_it;
                // The K block this warp filled last is published only after the next one has been issued.  If the next
                // tile's record is not there yet, publish first: the record may be waiting for this very K block (it is
                // claimed when the epilogue moves on, i.e. after the MMAs that need the block; one-K-block tiles, dim 64).
                if (prev_g >= 0 && !mbar_test(mfull_bar(cur_it % UM_MD), (uint32_t)(cur_it / UM_MD) & 1u)) {
                    publish(prev_g, 0);
                    prev_g = -1;
                }
                const TileMeta *m = meta_acquire(cur_it);
                nq = m->nq;
                qrow = m->qrow[lane];
                if (nq < 0) return false;
                meta_release(cur_it);
            }
            return true;
        };
        const int64_t row_stride = 3 * (int64_t)a.dim;                 // bf16 elements per query row of the split table
        for (int g = fw;; g += UM_FILL_WARPS) {
            const int it = g / nkb, kb = g - it * nkb;
            if (!advance_to(it)) break;
            const int sb = g % UM_SB;
            const uint32_t phb = (uint32_t)(g / UM_SB) & 1u;
            const uint32_t bst = b_ring + (uint32_t)sb * UM_B_BYTES;
            const int nq_eff = (a.flags & (1u << 30)) ? 0 : nq;         // 3 terms x 8 16-byte chunks per pair and K block
            const int ts = nq <= 16 ? 16 : 32;                            // rows of term t start t * TS rows into the B tile
            mbar_wait(empty_bar(sb), phb ^ 1u);
            // lane -> (pair jb + lane/8, chunk lane%8): four pairs x eight 16-byte chunks per instruction, one term at a time
            const int jl = lane >> 3, c = lane & 7;
            for (int jb = 0; jb < nq_eff; jb += 4) {
                const int j = jb + jl;
                const int qr = __shfl_sync(0xffffffffu, qrow, j & 31);
                if (j < nq_eff) {
                    const __nv_bfloat16 *src = a.qsplit + (int64_t)qr * row_stride + kb * UM_BLOCK_K + c * 8;
                    const uint32_t dst = bst + (uint32_t)(j * 128 + ((c ^ (j & 7)) << 4));
#pragma unroll
                    for (int t = 0; t < 3; ++t)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(t * ts * 128)), "l"(src + t * a.dim) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (prev_g >= 0) publish(prev_g, 1);
            prev_g = g;
        }
        if (prev_g >= 0) publish(prev_g, 0);
    } else if (warp < 2 + UM_FILL_WARPS + 4) {
        // ===================== epilogue (128 threads) =====================
        const int wq = warp & 3;                    // TMEM lane quarter this warp may read
        const int row = wq * 32 + lane;
        for (int it = 0;; ++it) {
            const TileMeta *m = meta_acquire(it);
            const int nrows = m->nrows, nq = m->nq;
            const int64_t off = m->off[lane];      // lane l owns column l: where pair l's scores of this tile start
            if (nq < 0) break;
            meta_release(it);
            const int acc = it & 1;
            const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)acc * UM_ACC_COLS;
            // score[c] = D[c] + D[TS + c] + D[2*TS + c]  (hi + mid + lo terms)
            float sum[UMMA_NQ];
            if (nq <= 16) {
                uint32_t r[3][16];
#pragma unroll
                for (int t = 0; t < 3; ++t) tmem_ld16(taddr + t * 16, r[t]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    sum[j] = (__uint_as_float(r[0][j]) + __uint_as_float(r[1][j])) + __uint_as_float(r[2][j]);
            } else {
                uint32_t r[2][16];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    tmem_ld16(taddr + t * 32, r[0]);
                    tmem_ld16(taddr + t * 32 + 16, r[1]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(r[j >> 4][j & 15]);
                        sum[j] = t == 0 ? v : sum[j] + v;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));     // accumulator is in registers: release it to the MMA warp
#pragma unroll
            for (int c = 0; c < UMMA_NQ; ++c) {
                if (c < nq) {                                 // warp-uniform
                    const int64_t o = __shfl_sync(0xffffffffu, off, c);
                    if (row < nrows) a.scorebuf[o + row] = apply_act(sum[c], a.act);
                }
            }
        }
    }

    else {
        // ===================== tile scheduler (one lane): claim the next tile, bulk-copy its TileMeta record into the ring =====================
        // Tiles are claimed one at a time from a global counter, so CTAs that run slower (an SM shared with the previous
        // batch's top-k CTAs) or start later (an SM still busy with the previous batch's scoring CTA) simply take fewer.
        for (int it = 0;; ++it) {
            const int slot = it % UM_MD;
            mbar_wait(mempty_bar(slot), ((uint32_t)(it / UM_MD) & 1u) ^ 1u);
            int t = 0;
            if (lane == 0) {
                t = atomicAdd(&a.counters[CTR_TILE_NEXT], 1);
                if (t < n_tiles) {
                    mbar_arrive_expect_tx(mfull_bar(slot), (uint32_t)sizeof(TileMeta));
                    bulk_load(smem_u32(meta + slot), a.tile_meta + t, (uint32_t)sizeof(TileMeta), mfull_bar(slot));
                } else {
                    meta[slot].nq = -1;                      // end of work: every role leaves its loop on this record
                    mbar_arrive(mfull_bar(slot));
                }
            }
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= n_tiles) break;
        }
    }

    tc_fence_before();
    __syncthreads();
    // The dependent grid (this batch's top-k: one CTA per query) is released only now, at the CTA's very end.  Triggered
    // at kernel entry, its ~1,000 CTAs became resident at once and sat in griddepcontrol.wait for the whole scoring
    // kernel, holding the registers, shared memory and thread slots that the previous batch's top-k and the next
    // batch's inversion needed to run beside this kernel (measured: 17% slower scoring, 3x slower top-k when pipelined).
    pdl_launch_dependents();
    trace_end(a.dbg, 3);
    if (threadIdx.x == 0 && atomicAdd(&a.counters[CTR_TILE_DONE], 1) == (int)gridDim.x - 1) {
        a.counters[CTR_TILE_NEXT] = 0;                       // every CTA has made its last claim: leave the queue ready for the next launch
        a.counters[CTR_TILE_DONE] = 0;
    }
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(UM_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool umma_make_tensor_map(CUtensorMap *out, const void *emb, int64_t n_docs, int dim) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)n_docs};
    const cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
    const cuuint32_t box[2] = {UM_BLOCK_K, UMMA_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn)(
        out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(emb), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

cudaError_t launch_score_umma(const ScoreArgs &a, const CUtensorMap *tmap, cudaStream_t s, int sm_count) {
    static unsigned long long attr_set_mask = 0;      // one bit per device: the attribute is per device and function
    int dev = 0;
    cudaGetDevice(&dev);
    const bool attr_set = (attr_set_mask >> (dev & 63)) & 1ull;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_score_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set_mask |= 1ull << (dev & 63);
    }
    return launch_pdl(k_score_umma, dim3(sm_count), dim3(UM_THREADS), UM_SMEM_BYTES, s, *tmap, a);
}

}  // namespace gdr
