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
#include "gdr_common.cuh"

namespace gdr {

constexpr int TK_THREADS = 256;
constexpr int TK_BINS = 2048;      // 11-bit digits (sign + exponent + 2 mantissa bits in the first pass): 8 KB of shared memory

struct TkShared {
    int sel_count;
    int eq2_count;
    int found_bin, found_gt, found_eq;
    int warp_tot[TK_THREADS / 32];
};

struct Threshold {
    uint32_t prefix;  // selected high bits (low `shift` bits are zero)
    int shift;        // bits of the key NOT yet decided; 32 = take everything
    int need;         // how many of the boundary class are still needed
    int n_eq;         // size of the boundary class
};

// Select the kk largest keys among the active candidates.  key_at(j, key) returns false for
// inactive candidates.  All threads of the CTA call this with identical arguments.
template <typename KeyAt>
__device__ Threshold radix_select(int n, int kk, uint32_t *hist, TkShared *sh, KeyAt key_at) {
    Threshold th{0u, 32, kk, 0};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);      // 11 + 11 + 10 bits
        const int nb = pass == 2 ? 1024 : 2048;
        for (int i = tid; i < nb; i += TK_THREADS) hist[i] = 0;
        __syncthreads();
        for (int j = tid; j < n; j += TK_THREADS) {
            uint32_t key;
            if (!key_at(j, key)) continue;
            if (th.shift == 32 || (key >> th.shift) == (th.prefix >> th.shift))
                atomicAdd(&hist[(key >> shift) & (nb - 1)], 1u);
        }
        __syncthreads();
        // thread t owns `per` bins counted from the top: [nb - (t+1)*per, nb - t*per)
        const int per = nb / TK_THREADS;
        const int top = nb - tid * per;
        int local = 0;
        for (int i = 1; i <= per; ++i) local += hist[top - i];
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sh->warp_tot[warp] = incl;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += sh->warp_tot[w];
        incl += before;
        const int excl = incl - local;          // candidates in bins above this thread's range
        if (excl < th.need && th.need <= incl) {
            int running = excl;
            for (int i = 1; i <= per; ++i) {
                const int h = hist[top - i];
                if (running + h >= th.need) {
                    sh->found_bin = top - i;
                    sh->found_gt = running;
                    sh->found_eq = h;
                    break;
                }
                running += h;
            }
        }
        __syncthreads();
        th.prefix |= (uint32_t)sh->found_bin << shift;
        th.need -= sh->found_gt;
        th.n_eq = sh->found_eq;
        th.shift = shift;
        __syncthreads();
        if (th.n_eq == th.need) break;
    }
    return th;
}

__device__ __forceinline__ void bitonic_sort_desc(uint64_t *v, int cap) {
    for (int size = 2; size <= cap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < cap; i += TK_THREADS) {
                const int p = i ^ stride;
                if (p > i) {
                    const uint64_t a = v[i], b = v[p];
                    const bool desc = (i & size) == 0;
                    if (desc ? (a < b) : (a > b)) { v[i] = b; v[p] = a; }
                }
            }
        }
    }
    __syncthreads();
}

// ---- candidate sources ------------------------------------------------------------------------
struct StoreSrc {   // a query's candidates = its K beam segments of the score buffer
    const float *sb;        // score buffer row of this query
    const int32_t *co;      // [K+1] segment starts (shared memory)
    const int32_t *cbase;   // [K] first store row of each beam's cluster (shared memory)
    const float *prob;      // [K] or null
    const int32_t *docid;
    int K;
    float alpha;
    __device__ int seg(int j) const {
        int lo = 0, hi = K - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (co[mid + 1] <= j) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    __device__ float score(int j) const {
        float s = sb[j];
        // main_models.py:1623-1624: score + alpha * p[b][i], two roundings (no FMA contraction)
        if (prob) s = __fadd_rn(s, __fmul_rn(alpha, prob[seg(j)]));
        return s;
    }
    __device__ int32_t doc(int j) const {
        const int i = seg(j);
        return docid[cbase[i] + (j - co[i])];
    }
};

struct ListSrc {    // explicit candidate lists from G ranks: [G, B, k_in]
    const float *scores;
    const int32_t *docids;
    int64_t g_stride;   // elements between consecutive ranks' blocks
    int k_in;
    __device__ int64_t at(int j) const { return (int64_t)(j / k_in) * g_stride + (j % k_in); }
    __device__ float score(int j) const { return scores[at(j)]; }
    __device__ int32_t doc(int j) const { return docids[at(j)]; }
};

template <typename Src>
__device__ void topk_general(const Src &src, int n, int k, int cap, uint32_t *keys, uint64_t *sel, uint32_t *hist,
                          TkShared *sh, float *out_s, int32_t *out_d) {
    const int tid = threadIdx.x;
    if (tid == 0) { sh->sel_count = 0; sh->eq2_count = 0; }
    for (int j = tid; j < n; j += TK_THREADS) keys[j] = float_to_ordered(src.score(j));
    __syncthreads();
    const int kk = min(k, n);
    Threshold t1{0u, 32, kk, n};
    if (n > k) t1 = radix_select(n, kk, hist, sh, [&](int j, uint32_t &key) { key = keys[j]; return true; });
    const bool tie = t1.shift == 0 && t1.n_eq > t1.need;   // exact-score ties straddle the cut
    Threshold t2{0u, 32, t1.need, t1.n_eq};
    if (tie) {
        const uint32_t T = t1.prefix;
        t2 = radix_select(n, t1.need, hist, sh, [&](int j, uint32_t &key) {
            if (keys[j] != T) return false;
            key = ~(uint32_t)src.doc(j);
            return true;
        });
    }
    for (int j = tid; j < n; j += TK_THREADS) {
        const uint32_t key = keys[j];
        bool take = true;
        uint32_t nd = 0;
        bool have_doc = false;
        if (t1.shift < 32) {
            const uint32_t hi = key >> t1.shift, thi = t1.prefix >> t1.shift;
            if (hi < thi) take = false;
            else if (hi == thi && tie) {
                nd = ~(uint32_t)src.doc(j);
                have_doc = true;
                if (t2.shift < 32) {
                    const uint32_t hi2 = nd >> t2.shift, thi2 = t2.prefix >> t2.shift;
                    if (hi2 < thi2) take = false;
                    else if (hi2 == thi2 && t2.n_eq > t2.need) take = atomicAdd(&sh->eq2_count, 1) < t2.need;
                }
            }
        }
        if (take) {
            if (!have_doc) nd = ~(uint32_t)src.doc(j);
            const int slot = atomicAdd(&sh->sel_count, 1);
            if (slot < cap) sel[slot] = ((uint64_t)key << 32) | nd;
        }
    }
    __syncthreads();
    for (int i = kk + tid; i < cap; i += TK_THREADS) sel[i] = 0ull;
    bitonic_sort_desc(sel, cap);
    for (int r = tid; r < k; r += TK_THREADS) {
        float s = -INFINITY;
        int32_t d = -1;
        if (r < kk) {
            const uint64_t v = sel[r];
            s = ordered_to_float((uint32_t)(v >> 32));
            d = (int32_t)(~(uint32_t)v);
        }
        out_s[r] = s;
        out_d[r] = d;
    }
}


// ---- fast path ---------------------------------------------------------------------------------
// Two passes over the n keys instead of five: (1) build keys + 2048-bin histogram of the top 11 key bits,
// (2) classify against the boundary bin: keys above it are selected outright, keys inside it (~n/32 on
// spread-out scores) go to a small boundary list that is resolved by rank counting on (key, ~docid).
// The k survivors are ordered by rank counting as well (k <= 256) — no bitonic network, ~7 barriers in all.
// Falls back to the general radix select when the boundary bin holds more than TK_BND keys (mass ties).
constexpr int TK_BND = 256;

template <typename Src>
__device__ void topk_body(const Src &src, int n, int k, int cap, uint32_t *keys, uint64_t *sel, uint32_t *hist,
                          TkShared *sh, float *out_s, int32_t *out_d) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= k || cap > 256) {     // few candidates (take all) or large k: general path
        topk_general(src, n, k, cap, keys, sel, hist, sh, out_s, out_d);
        return;
    }
    for (int i = tid; i < TK_BINS; i += TK_THREADS) hist[i] = 0;
    if (tid == 0) { sh->sel_count = 0; sh->eq2_count = 0; }
    __syncthreads();
    for (int j = tid; j < n; j += TK_THREADS) {
        const uint32_t key = float_to_ordered(src.score(j));
        keys[j] = key;
        atomicAdd(&hist[key >> 21], 1u);
    }
    __syncthreads();
    // boundary bin: warp w sums bins [256w, 256w + 256) with conflict-free strided reads ...
    {
        int part = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) part += hist[warp * 256 + i * 32 + lane];
#pragma unroll
        for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        if (lane == 0) sh->warp_tot[warp] = part;
    }
    __syncthreads();
    // ... then every warp redundantly walks down from the top to the 256-bin range and the bin where the
    // cumulative count reaches k (no further barrier needed: all threads end up with the same d / gt / eq)
    int above = 0, range = TK_THREADS / 32 - 1;
    for (; range > 0; --range) {
        const int t = sh->warp_tot[range];
        if (above + t >= k) break;
        above += t;
    }
    int d_bin, gt, eq;
    {
        const int top = range * 256 + 256 - lane * 8;        // lane owns bins [top-8, top), lane 0 the highest
        int local = 0;
#pragma unroll
        for (int i = 1; i <= 8; ++i) local += hist[top - i];
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int need_here = k - above;
        const bool mine = (incl - local) < need_here && need_here <= incl;
        int fb = 0, fg = 0, fe = 0;
        if (mine) {
            int running = above + incl - local;
            for (int i = 1; i <= 8; ++i) {
                const int h = hist[top - i];
                if (running + h >= k) { fb = top - i; fg = running; fe = h; break; }
                running += h;
            }
        }
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        const int srcl = __ffs(who) - 1;
        d_bin = __shfl_sync(0xffffffffu, fb, srcl);
        gt = __shfl_sync(0xffffffffu, fg, srcl);
        eq = __shfl_sync(0xffffffffu, fe, srcl);
    }
    if (eq > TK_BND) {             // mass ties in the boundary bin: general path (uniform decision)
        __syncthreads();
        topk_general(src, n, k, cap, keys, sel, hist, sh, out_s, out_d);
        return;
    }
    // classify: sel[0, gt) <- keys above the boundary bin; bnd[0, eq) <- keys inside it   (bnd aliases hist)
    uint64_t *bnd = reinterpret_cast<uint64_t *>(hist);           // 2 x TK_BND x 8 B = 4 KB <= the 8 KB histogram
    uint64_t *bnd2 = bnd + TK_BND;
    __syncthreads();                                              // everyone is done reading hist
    for (int j = tid; j < n; j += TK_THREADS) {
        const uint32_t key = keys[j];
        const int bin = (int)(key >> 21);
        if (bin > d_bin) sel[atomicAdd(&sh->sel_count, 1)] = ((uint64_t)key << 32) | (uint32_t)j;
        else if (bin == d_bin) bnd[atomicAdd(&sh->eq2_count, 1)] = ((uint64_t)key << 32) | (uint32_t)j;
    }
    __syncthreads();
    const int need = k - gt;                                      // 1 <= need <= eq
    if (need == eq) {
        if (tid < eq) sel[gt + tid] = bnd[tid];
    } else {
        // order the boundary keys by (key desc, docid asc); duplicates of the same (key, docid) by list position
        if (tid < eq) {
            const uint64_t e = bnd[tid];
            bnd2[tid] = (e & 0xffffffff00000000ull) | (uint32_t)~(uint32_t)src.doc((int)(uint32_t)e);
        }
        __syncthreads();
        if (tid < eq) {
            const uint64_t mine = bnd2[tid];
            int rank = 0;
            for (int u = 0; u < eq; ++u) {
                const uint64_t o = bnd2[u];
                rank += (o > mine) || (o == mine && u < tid);
            }
            if (rank < need) sel[gt + rank] = bnd[tid];
        }
    }
    __syncthreads();
    // (key, candidate index) -> (key, ~docid), then order the k survivors by rank counting and write them out
    if (tid < k) {
        const uint64_t e = sel[tid];
        sel[tid] = (e & 0xffffffff00000000ull) | (uint32_t)~(uint32_t)src.doc((int)(uint32_t)e);
    }
    __syncthreads();
    if (tid < k) {
        const uint64_t mine = sel[tid];
        int rank = 0;
        for (int u = 0; u < k; ++u) {
            const uint64_t o = sel[u];
            rank += (o > mine) || (o == mine && u < tid);
        }
        out_s[rank] = ordered_to_float((uint32_t)(mine >> 32));
        out_d[rank] = (int32_t)(~(uint32_t)mine);
    }
}

// dynamic shared memory layout: sel[cap] u64 | hist[TK_BINS] u32 | co[K+1] i32 | cbase[K] i32 | keys[...] u32 (smem variant)
template <bool KEYS_IN_SMEM>
__global__ void __launch_bounds__(TK_THREADS, 8) k_topk_store(ScoreArgs a, float alpha, int cap, float *out_scores,
                                                              int32_t *out_docids) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TkShared sh;
    uint64_t *sel = reinterpret_cast<uint64_t *>(smem);
    uint32_t *hist = reinterpret_cast<uint32_t *>(sel + cap);
    int32_t *co = reinterpret_cast<int32_t *>(hist + TK_BINS);
    const int b = blockIdx.x;
    int32_t *cbase = co + a.K + 1;
    for (int i = threadIdx.x; i <= a.K; i += TK_THREADS) {
        co[i] = a.candoff[(int64_t)b * (a.K + 1) + i];
        if (i < a.K) {
            const int c = a.beams[(int64_t)b * a.K + i];
            cbase[i] = (c >= 0 && c < a.n_clusters) ? a.offsets[c] : 0;
        }
    }
    __syncthreads();
    uint32_t *keys = KEYS_IN_SMEM ? reinterpret_cast<uint32_t *>(cbase + a.K) : a.gkeys + (int64_t)b * a.stride;
    StoreSrc src{a.scorebuf + (int64_t)b * a.stride, co, cbase, a.prob ? a.prob + (int64_t)b * a.K : nullptr,
                 a.docid, a.K, alpha};
    topk_body(src, co[a.K], a.k, cap, keys, sel, hist, &sh, out_scores + (int64_t)b * a.k,
              out_docids + (int64_t)b * a.k);
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
    topk_body(src, n, k, cap, keys, sel, hist, &sh, out_scores + (int64_t)b * k, out_docids + (int64_t)b * k);
}

static int pow2_at_least(int x) { int p = 2; while (p < x) p <<= 1; return p; }

cudaError_t launch_topk_store(const ScoreArgs &a, float alpha, float *out_scores, int32_t *out_docids, cudaStream_t s) {
    if (a.B == 0) return cudaSuccess;
    const int cap = pow2_at_least(a.k);
    const size_t fixed = (size_t)cap * 8 + (size_t)TK_BINS * 4 + (size_t)(2 * a.K + 1) * 4;
    const size_t with_keys = fixed + (size_t)a.stride * 4;
    if (with_keys <= 96 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(k_topk_store<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            attr_set = true;
        }
        k_topk_store<true><<<a.B, TK_THREADS, with_keys, s>>>(a, alpha, cap, out_scores, out_docids);
    } else {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(k_topk_store<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            attr_set = true;
        }
        k_topk_store<false><<<a.B, TK_THREADS, fixed, s>>>(a, alpha, cap, out_scores, out_docids);
    }
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(const float *scores, const int32_t *docids, int G, int B, int k_in, int64_t g_stride, int k,
                              float *out_scores, int32_t *out_docids, cudaStream_t s) {
    if (B == 0) return cudaSuccess;
    const int cap = pow2_at_least(k);
    const int n = G * k_in;
    const size_t smem = (size_t)cap * 8 + TK_BINS * 4 + (size_t)n * 4;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    ListSrc src{scores, docids, g_stride, k_in};
    k_topk_merge<<<B, TK_THREADS, smem, s>>>(src, n, k, cap, out_scores, out_docids);
    return cudaGetLastError();
}

}  // namespace gdr
