// Per-query top-k: radix select in shared memory + bitonic sort of the k survivors.
//
// Replaces main_models.py:1619-1626 (per-alpha segment bias `score[seg_i] += alpha * p[b][i]`,
// then `Tensor.topk(k, largest=True, sorted=True)`) and :1628-1631 (candidate index -> doc index).
// One CTA per query.  Candidate scores are read once from the score buffer (written by the scoring
// kernels, normally still L2-resident), biased, mapped to order-preserving uint32 keys and kept in
// shared memory (global scratch when a query has too many candidates).  The k-th largest key is
// found with three histogram passes (11 + 11 + 10 bits, early exit when a bucket is taken whole);
// exact score ties at the threshold are broken by ascending docid with a second select over the
// tied candidates, so the result does not depend on candidate order (and therefore not on how the
// corpus is sharded across GPUs).  The same kernel, with an explicit candidate list as source,
// is the merge step after the cross-rank exchange.
//
// Kernels: k_topk_fast (k <= 128 and <= 65,535 candidates per query: 128 threads, 16-bit histogram bins, 5.4 KB of shared
// memory, keys of up to 2,560 candidates held in registers) is what the reference's top-100 runs; k_topk_store (256 threads)
// covers larger k / more candidates with the three-pass radix select; k_topk_merge merges per-rank lists.
#include "gdr_common.cuh"
#include "topk_select.cuh"

namespace gdr {

// dynamic shared memory: sel[128] u64 | hist[2048] u16 | co[K+1] | cbase[K] | bias[K]
// (eight CTAs per SM = 64 registers: 1,024 queries on 148 SMs are seven CTAs per SM)
__global__ void __launch_bounds__(TKF_THREADS, 8) k_topk_fast(ScoreArgs a, float alpha, float *out_scores, int32_t *out_docids) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TkShared sh;
    uint64_t *sel = reinterpret_cast<uint64_t *>(smem);
    uint32_t *hist_words = reinterpret_cast<uint32_t *>(sel + 128);
    int32_t *co = reinterpret_cast<int32_t *>(hist_words + TK_BINS / 2);
    int32_t *cbase = co + a.K + 1;
    float *bias = reinterpret_cast<float *>(cbase + a.K);
    const int b = blockIdx.x;                 // row of the score buffer / of the outputs
    const int bg = b + a.q_base;              // the query's index in the (global) batch: candoff / cbase / prob rows
    pdl_wait();
    trace_start(a.dbg, 4);
    if (a.n_ranks > 1) {                      // sharded corpus: the other ranks' scores of this batch have landed
        if (threadIdx.x == 0) wait_for_scorers(a);
        __syncthreads();
    }
    for (int i = threadIdx.x; i <= a.K; i += TKF_THREADS) {          // one round trip: k_count left both arrays per query
        co[i] = a.candoff[(int64_t)bg * (a.K + 1) + i];
        if (i < a.K) {
            cbase[i] = a.cbase[(int64_t)bg * a.K + i];
            if (a.prob) bias[i] = __fmul_rn(alpha, a.prob[(int64_t)bg * a.K + i]);
        }
    }
    const int n = a.candoff[(int64_t)bg * (a.K + 1) + a.K];          // every thread reads it itself: no barrier before the score loads
    const uint32_t dbg = (a.flags >> 20) & 15u;
    if (dbg == 9u) {                          // timing experiment: stay resident ~15 us without doing anything, then stop
        for (int i = 0; i < 150; ++i) __nanosleep(100);
        return;
    }
    if (dbg & 1u) return;                     // stop after the prologue
    StoreSrc src{a.scorebuf + (int64_t)b * a.stride, co, cbase, a.prob ? bias : nullptr, a.docid, a.K};
#ifdef GDR_DEBUG_KNOBS
    if (dbg) {                                // measurement builds: the phase-by-phase cut-offs live in topk_fast16
        topk_fast16<TKF_THREADS, TKF_R4>(src, n, a.k, a.gkeys + (int64_t)b * a.stride, a.ghist + (int64_t)b * TK_BINS, sel, hist_words, &sh,
                                         out_scores + (int64_t)b * a.k, out_docids + (int64_t)b * a.k, dbg);
        return;
    }
#endif
    if (topk_lean_eligible<StoreSrc, CtaScope>(n, a.k))                  // the reference's regime: k <= 128 out of <= 2,560 candidates
        topk_lean128<StoreSrc>(src, n, a.k, a.gkeys + (int64_t)b * a.stride, a.ghist + (int64_t)b * TK_BINS, sel, hist_words, &sh,
                               out_scores + (int64_t)b * a.k, out_docids + (int64_t)b * a.k);
    else                                                                 // more candidates (streamed twice), n <= k: out of line
        topk_fast16_cold<TKF_THREADS, TKF_R4, StoreSrc, CtaScope>(src, n, a.k, a.gkeys + (int64_t)b * a.stride, a.ghist + (int64_t)b * TK_BINS, sel,
                                                                  hist_words, &sh, out_scores + (int64_t)b * a.k, out_docids + (int64_t)b * a.k);
    // dependents (the next batch's k_count on this stream) are released at the end: released at entry, their CTAs would hold
    // registers and thread slots beside this kernel for its whole duration (see k_score_umma)
    pdl_launch_dependents();
    trace_end(a.dbg, 5);
}

// dynamic shared memory layout: sel[cap] u64 | hist[TK_BINS] u32 | keys[stride] u32 (smem variant) | co[K+1] i32 | cbase[K] i32 | bias[K] f32
// KEYS: 0 = key array in global scratch, 1 = key array in shared memory, 2 = fast path (k <= 128) without a key array
// (its mass-tie fallback uses the global scratch)
template <int KEYS>
__global__ void __launch_bounds__(TK_THREADS, 4) k_topk_store(ScoreArgs a, float alpha, int cap, float *out_scores,
                                                              int32_t *out_docids) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TkShared sh;
    uint64_t *sel = reinterpret_cast<uint64_t *>(smem);
    uint32_t *hist = reinterpret_cast<uint32_t *>(sel + cap);
    uint32_t *skeys = hist + TK_BINS;                              // 16-byte aligned: cap * 8 + 8 KB
    int32_t *co = reinterpret_cast<int32_t *>(skeys + (KEYS == 1 ? a.stride : 0));
    int32_t *cbase = co + a.K + 1;
    float *bias = reinterpret_cast<float *>(cbase + a.K);
    pdl_wait();
    trace_start(a.dbg, 4);
    const int b = blockIdx.x;
    const int bg = b + a.q_base;
    if (a.n_ranks > 1 && threadIdx.x == 0) wait_for_scorers(a);      // (the barrier below orders the other threads behind it)
    for (int i = threadIdx.x; i <= a.K; i += TK_THREADS) {
        co[i] = a.candoff[(int64_t)bg * (a.K + 1) + i];
        if (i < a.K) {
            cbase[i] = a.cbase[(int64_t)bg * a.K + i];
            if (a.prob) bias[i] = __fmul_rn(alpha, a.prob[(int64_t)bg * a.K + i]);
        }
    }
    __syncthreads();
    uint32_t *keys = KEYS == 1 ? skeys : a.gkeys + (int64_t)b * a.stride;
    StoreSrc src{a.scorebuf + (int64_t)b * a.stride, co, cbase, a.prob ? bias : nullptr, a.docid, a.K};
    topk_body<TK_THREADS, KEYS != 2>(src, co[a.K], a.k, cap, keys, sel, hist, &sh, out_scores + (int64_t)b * a.k,
                         out_docids + (int64_t)b * a.k);
    trace_end(a.dbg, 5);
}

__global__ void __launch_bounds__(TK_THREADS) k_topk_merge(ListSrc src0, int n, int k, int cap, float *out_scores,
                                                           int32_t *out_docids) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TkShared sh;
    uint64_t *sel = reinterpret_cast<uint64_t *>(smem);
    uint32_t *hist = reinterpret_cast<uint32_t *>(sel + cap);
    uint32_t *keys = hist + TK_BINS;
    const int b = blockIdx.x;
    ListSrc src = src0;
    src.scores += (int64_t)b * src.k_in;
    src.docids += (int64_t)b * src.k_in;
    topk_body<TK_THREADS, true>(src, n, k, cap, keys, sel, hist, &sh, out_scores + (int64_t)b * k, out_docids + (int64_t)b * k);
}

// Sharded corpus, stand-alone top-k: ONE warp waits for every rank's arrival flag; the top-k grid behind it (same stream) then starts
// with all scores in place.  The wait is not done by the top-k CTAs themselves: with several batches in flight a grid of ~1,000 spinning
// CTAs could occupy the SMs that the scoring kernel another rank is waiting for needs — a cross-GPU deadlock; one spinning warp cannot.
__global__ void __launch_bounds__(32) k_wait_scorers(ScoreArgs a) {
    pdl_wait();
    if (threadIdx.x == 0) wait_for_scorers(a, false);
    __syncwarp();
    __threadfence_system();
}

// The other direction of the handshake (stand-alone schedule, several batches in flight): a rank may overwrite an owner's score buffer with
// the NEXT batch of this handle only after the owner's top-k of the previous batch has read it.  k_signal_consumed runs behind the top-k
// launches of a call and tells every rank "epoch e of this handle is consumed here"; k_wait_consumed runs in front of the scoring kernel of
// the handle's next call and waits for every owner's flag.  (The fused launches need neither: a launch raises its arrival flags only after
// its own top-k groups are done, and the next launch's groups wait for those flags — DESIGN.md §3.10.)
__global__ void __launch_bounds__(32) k_signal_consumed(ScoreArgs a) {
    pdl_wait();
    if (threadIdx.x == 0) signal_consumed(a);
}

__global__ void __launch_bounds__(32) k_wait_consumed(ScoreArgs a) {
    pdl_wait();
    if (threadIdx.x == 0) wait_consumed(a);
    __syncwarp();
}

cudaError_t launch_signal_consumed(const ScoreArgs &a, cudaStream_t s) { return launch_pdl(k_signal_consumed, dim3(1), dim3(32), 0, s, a.launch_prio, a); }
cudaError_t launch_wait_consumed(const ScoreArgs &a, cudaStream_t s) { return launch_pdl(k_wait_consumed, dim3(1), dim3(32), 0, s, a.launch_prio, a); }

cudaError_t launch_wait_scorers(const ScoreArgs &a, cudaStream_t s) {
    return launch_pdl(k_wait_scorers, dim3(1), dim3(32), 0, s, a.launch_prio, a);
}

static int pow2_at_least(int x) { int p = 2; while (p < x) p <<= 1; return p; }

cudaError_t launch_topk_store(const ScoreArgs &a, float alpha, float *out_scores, int32_t *out_docids, cudaStream_t s) {
    if (a.B_top == 0) return cudaSuccess;
    const int cap = pow2_at_least(a.k);
    const size_t fixed = (size_t)cap * 8 + (size_t)TK_BINS * 4 + (size_t)(3 * a.K + 1) * 4;
    const size_t with_keys = fixed + (size_t)a.stride * 4;
    static FuncAttrOnce attr;
    cudaError_t e = attr.ensure([] {
        cudaError_t e2 = cudaFuncSetAttribute(k_topk_store<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_topk_store<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        return e2;
    });
    if (e != cudaSuccess) return e;
    const int grid = a.B_top;  // (a few persistent CTAs per SM walking the queries measured slower in the pipelined step: 62 vs 57 us)
    const int groups = (int)((a.flags >> 16) & 7u);                // GDR_OPT_TOPK_GROUPS (topk_grouped.cu)
    const int pr = a.launch_prio;
    if (groups && cap <= 128 && a.gkeys && a.ghist && a.stride <= 65535) return launch_topk_grouped(a, alpha, out_scores, out_docids, s, groups);
    if (cap <= 128 && a.gkeys && a.ghist && a.stride <= 65535)
        return launch_pdl(k_topk_fast, dim3(grid), dim3(TKF_THREADS), (size_t)128 * 8 + TK_BINS * 2 + (size_t)(3 * a.K + 1) * 4, s, pr, a, alpha,
                          out_scores, out_docids);
    if (cap <= 128 && a.gkeys)
        return launch_pdl(k_topk_store<2>, dim3(grid), dim3(TK_THREADS), fixed, s, pr, a, alpha, cap, out_scores, out_docids);
    if (with_keys <= 96 * 1024)
        return launch_pdl(k_topk_store<1>, dim3(grid), dim3(TK_THREADS), with_keys, s, pr, a, alpha, cap, out_scores, out_docids);
    return launch_pdl(k_topk_store<0>, dim3(grid), dim3(TK_THREADS), fixed, s, pr, a, alpha, cap, out_scores, out_docids);
}

cudaError_t launch_merge_topk(const float *scores, const int32_t *docids, int G, int B, int k_in, int64_t g_stride, int k,
                              float *out_scores, int32_t *out_docids, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    const int cap = pow2_at_least(k);
    const int n = G * k_in;
    const size_t smem = (size_t)cap * 8 + TK_BINS * 4 + (size_t)((n + 3) / 4 * 4) * 4;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static FuncAttrOnce attr;
    cudaError_t e = attr.ensure([] { return cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (e != cudaSuccess) return e;
    ListSrc src{scores, docids, g_stride, k_in};
    k_topk_merge<<<B, TK_THREADS, smem, s>>>(src, n, k, cap, out_scores, out_docids);
    return cudaGetLastError();
}

}  // namespace gdr
