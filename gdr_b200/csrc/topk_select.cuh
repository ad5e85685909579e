// Per-query top-k selection: the device code shared by the stand-alone kernels (topk.cu) and the fused scoring + top-k CTA
// (score_fused.cu).  See topk.cu for what the kernels are and where they are used.
#pragma once
#include <type_traits>
#include "gdr_common.cuh"

namespace gdr {

constexpr int TK_THREADS = 256;
constexpr int TK_BINS = 2048;      // 11-bit digits (sign + exponent + 2 mantissa bits in the first pass): 8 KB of shared memory

struct TkShared {
    int sel_count;
    int eq2_count;
    int bnd_count;
    uint32_t hist2[256];
    int found_bin, found_gt, found_eq;
    int warp_tot[TK_THREADS / 32];
};

struct Threshold {
    uint32_t prefix;  // selected high bits (low `shift` bits are zero)
    int shift;        // bits of the key NOT yet decided; 32 = take everything
    int need;         // how many of the boundary class are still needed
    int n_eq;         // size of the boundary class
};

// Who is "the group" that runs one top-k: in the stand-alone kernels it is the whole CTA (CtaScope: threadIdx.x and
// __syncthreads() — the kernels compile to the SASS they had before the policy existed); GroupScope<FIRST> is one of the
// NT-thread groups (128 or 64) that follow the first FIRST threads of a larger CTA (group g = threads [FIRST + NT g, FIRST + NT g + NT),
// named barrier 1 + g) — the building block of the fused scoring + top-k CTA of ROADMAP.md.  Purely static: no argument is
// added to any function, so the default instantiation is the code it was.
struct CtaScope {
    __device__ __forceinline__ static int tid() { return (int)threadIdx.x; }
    __device__ __forceinline__ static void sync() { __syncthreads(); }
};
template <int FIRST, int NT = 128>
struct GroupScope {
    static_assert(NT == 64 || NT == 128, "group sizes: two or four warps");
    __device__ __forceinline__ static int tid() { return ((int)threadIdx.x - FIRST) & (NT - 1); }
    __device__ __forceinline__ static int group() { return ((int)threadIdx.x - FIRST) / NT; }
    __device__ __forceinline__ static void sync() { asm volatile("bar.sync %0, %1;" ::"r"(1 + group()), "n"(NT) : "memory"); }
};

// Select the kk largest keys among the active candidates.  key_at(j, key) returns false for
// inactive candidates.  All threads of the CTA call this with identical arguments.
template <int NT, typename Scope = CtaScope, typename KeyAt>
__device__ Threshold radix_select(int n, int kk, uint32_t *hist, TkShared *sh, KeyAt key_at) {
    Threshold th{0u, 32, kk, 0};
    const int tid = Scope::tid(), lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);      // 11 + 11 + 10 bits
        const int nb = pass == 2 ? 1024 : 2048;
        for (int i = tid; i < nb; i += NT) hist[i] = 0;
        Scope::sync();
        for (int j = tid; j < n; j += NT) {
            uint32_t key;
            if (!key_at(j, key)) continue;
            if (th.shift == 32 || (key >> th.shift) == (th.prefix >> th.shift))
                atomicAdd(&hist[(key >> shift) & (nb - 1)], 1u);
        }
        Scope::sync();
        // thread t owns `per` bins counted from the top: [nb - (t+1)*per, nb - t*per)
        const int per = nb / NT;
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
        Scope::sync();
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
        Scope::sync();
        th.prefix |= (uint32_t)sh->found_bin << shift;
        th.need -= sh->found_gt;
        th.n_eq = sh->found_eq;
        th.shift = shift;
        Scope::sync();
        if (th.n_eq == th.need) break;
    }
    return th;
}

template <int NT, typename Scope = CtaScope>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t *v, int cap) {
    for (int size = 2; size <= cap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            Scope::sync();
            for (int i = Scope::tid(); i < cap; i += NT) {
                const int p = i ^ stride;
                if (p > i) {
                    const uint64_t a = v[i], b = v[p];
                    const bool desc = (i & size) == 0;
                    if (desc ? (a < b) : (a > b)) { v[i] = b; v[p] = a; }
                }
            }
        }
    }
    Scope::sync();
}

// ---- candidate sources ------------------------------------------------------------------------
struct StoreSrc {   // a query's candidates = its K beam segments of the score buffer
    const float *sb;        // score buffer row of this query
    const int32_t *co;      // [K+1] segment starts (shared memory)
    const int32_t *cbase;   // [K] first store row of each beam's cluster (shared memory)
    const float *bias;      // [K] alpha * p[b][i] (shared memory) or null
    const int32_t *docid;
    int K;
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
        if (bias) s = __fadd_rn(s, bias[seg(j)]);
        return s;
    }
    __device__ int32_t doc(int j) const {
        const int i = seg(j);
        return docid[cbase[i] + (j - co[i])];
    }
    // four consecutive candidates starting at j4 (multiple of 4; the score row is 16-byte aligned and padded).
    // One segment search per four: they almost always share a beam segment; otherwise walk forward from it.
    __device__ float4 load4(int j4) const { return *reinterpret_cast<const float4 *>(sb + j4); }
    __device__ void score4(int j4, int n, float (&s)[4]) const { bias4(load4(j4), j4, n, s); }
    __device__ void bias4(const float4 v, int j4, int n, float (&s)[4]) const {
        s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
        if (bias) {
            int i = seg(j4);
            if (j4 + 3 < co[i + 1]) {
                const float bv = bias[i];
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = __fadd_rn(s[e], bv);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    while (i < K - 1 && co[i + 1] <= j4 + e) ++i;
                    if (j4 + e < n) s[e] = __fadd_rn(s[e], bias[i]);
                }
            }
        }
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
    __device__ void score4(int j4, int n, float (&s)[4]) const {
#pragma unroll
        for (int e = 0; e < 4; ++e) s[e] = j4 + e < n ? scores[at(j4 + e)] : 0.f;
    }
};

template <int NT, typename Src, typename Scope = CtaScope>
__device__ void topk_general(const Src &src, int n, int k, int cap, uint32_t *keys, uint64_t *sel, uint32_t *hist,
                          TkShared *sh, float *out_s, int32_t *out_d) {
    const int tid = Scope::tid();
    if (tid == 0) { sh->sel_count = 0; sh->eq2_count = 0; }
    for (int j4 = tid * 4; j4 < n; j4 += NT * 4) {            // keys[] is padded to a multiple of 4
        float s4[4];
        src.score4(j4, n, s4);
        *reinterpret_cast<uint4 *>(keys + j4) = make_uint4(float_to_ordered(s4[0]), float_to_ordered(s4[1]), float_to_ordered(s4[2]),
                                                           float_to_ordered(s4[3]));
    }
    Scope::sync();
    const int kk = min(k, n);
    Threshold t1{0u, 32, kk, n};
    if (n > k) t1 = radix_select<NT, Scope>(n, kk, hist, sh, [&](int j, uint32_t &key) { key = keys[j]; return true; });
    const bool tie = t1.shift == 0 && t1.n_eq > t1.need;   // exact-score ties straddle the cut
    Threshold t2{0u, 32, t1.need, t1.n_eq};
    if (tie) {
        const uint32_t T = t1.prefix;
        t2 = radix_select<NT, Scope>(n, t1.need, hist, sh, [&](int j, uint32_t &key) {
            if (keys[j] != T) return false;
            key = ~(uint32_t)src.doc(j);
            return true;
        });
    }
    for (int j = tid; j < n; j += NT) {
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
    Scope::sync();
    for (int i = kk + tid; i < cap; i += NT) sel[i] = 0ull;
    bitonic_sort_desc<NT, Scope>(sel, cap);
    for (int r = tid; r < k; r += NT) {
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


// One warp, 256 bins, lane owns bins [256 - 8*lane - 8, 256 - 8*lane) (lane 0 the highest): find the bin d where
// the count accumulated from the top reaches `need`; gt = count strictly above d, eq = count in d.
template <typename BinT>
__device__ __forceinline__ void scan_down8(const BinT *bins, int lane, int need, int &d, int &gt, int &eq) {
    const int top = 256 - lane * 8;
    int local = 0;
#pragma unroll
    for (int i = 1; i <= 8; ++i) local += (int)bins[top - i];
    int incl = local;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += t;
    }
    const bool mine = (incl - local) < need && need <= incl;
    int fb = 0, fg = 0, fe = 0;
    if (mine) {
        int running = incl - local;
        for (int i = 1; i <= 8; ++i) {
            const int h = (int)bins[top - i];
            if (running + h >= need) { fb = top - i; fg = running; fe = h; break; }
            running += h;
        }
    }
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const int srcl = __ffs(who) - 1;
    d = __shfl_sync(0xffffffffu, fb, srcl);
    gt = __shfl_sync(0xffffffffu, fg, srcl);
    eq = __shfl_sync(0xffffffffu, fe, srcl);
}

// ---- fast path ---------------------------------------------------------------------------------
// For k <= 128 (the reference's top-100) and more candidates than k.  Two passes over the n keys:
//   (1) build keys (128-bit loads) + 2048-bin histogram of the top 11 key bits;
//   (2) classify against the boundary bin: keys above it are selected outright, keys inside it (~n/32 on
//       spread-out scores) go to a short boundary list and into a 256-bin histogram of the next 8 bits;
// the boundary list is then cut by that second histogram, what is left tied after 19 bits (normally 1-2
// keys) is ordered by rank counting on (key, ~docid), and the k survivors are sorted by ONE warp with a
// register bitonic network (4 keys per lane, shuffles only).  ~8 block barriers in all.  Falls back to the
// general radix select when the boundary bin holds more than TK_BND keys (mass ties).
constexpr int TK_BND = 256;

// descending bitonic sort of 128 u64 keys held 4 per lane (element index = lane * 4 + r).  The (size, stride) loops are
// NOT unrolled: fully unrolled the network is ~1,300 instructions (21 KB of SASS) executed once per query by one warp,
// and the top-k kernel was losing a quarter of its issue slots to instruction-cache misses (ncu: no_inst 26%).
__device__ __forceinline__ void warp_bitonic128_desc(uint64_t (&v)[4], int lane) {
    auto local_stage = [&](int size, int stride) {                 // both partners in this lane's registers
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if ((r & stride) == 0) {
                const int p = r | stride;
                const bool desc = (((lane * 4 + r) & size) == 0);
                const uint64_t mx = v[r] > v[p] ? v[r] : v[p], mn = v[r] > v[p] ? v[p] : v[r];
                v[r] = desc ? mx : mn;
                v[p] = desc ? mn : mx;
            }
        }
    };
#pragma unroll 1
    for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride >= 4; stride >>= 1) {  // partner in lane ^ (stride / 4)
            const int lm = stride >> 2;
            const bool lower = (lane & lm) == 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint64_t other = __shfl_xor_sync(0xffffffffu, v[r], lm);
                const bool desc = (((lane * 4 + r) & size) == 0);
                const bool want_max = (lower == desc);
                const uint64_t mx = v[r] > other ? v[r] : other, mn = v[r] > other ? other : v[r];
                v[r] = want_max ? mx : mn;
            }
        }
        if (size >= 4) local_stage(size, 2);
        local_stage(size, 1);
    }
}

// STORE_KEYS = false: the keys are not kept between the two passes (pass 2 re-reads the L2-resident scores), which cuts
// the CTA's shared memory from ~20 KB to ~9 KB so that more of these CTAs co-reside with the scoring kernel of the next
// batch on every SM; `keys` then points to global scratch used only by the general fallback.
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

// ---- large k (128 < k <= 1,024: the reference's default beam = top-k = 100 ... 1,000 regime, BASELINE.json configs[2]) ----------------
// The general select above makes three histogram passes over ALL n keys and sorts with a shared-memory bitonic network of 55
// barrier-separated stages (measured 170 us per 1,024-query batch at k = 1,000).  Here: ONE pass over the keys (2,048-bin histogram
// of the top 11 bits), one pass that classifies them against the boundary bin — keys above it are selected, keys inside it go to a
// boundary list that aliases the histogram — then the list alone is cut by its next 8 bits and, if still tied, ranked on
// (key desc, docid asc).  The k survivors are sorted by a BLOCKED bitonic network: every thread holds EPT = cap / 256 entries in
// registers; strides below EPT are register swaps, strides inside a warp are shuffles, only the strides that cross warps (three
// sizes, six stages at cap = 1,024) go through shared memory.  Falls back to the general select when a list would overflow.
template <int EPT, typename Scope>
__device__ __forceinline__ void blocked_bitonic_desc(uint64_t *sel, int tid) {
    constexpr int NT = 256, CAP = NT * EPT;
    const int lane = tid & 31;
    uint64_t v[EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) v[r] = sel[tid * EPT + r];
#pragma unroll 1
    for (int size = 2; size <= CAP; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32 * EPT) {                              // partner in another warp: through shared memory
                Scope::sync();
#pragma unroll
                for (int r = 0; r < EPT; ++r) sel[tid * EPT + r] = v[r];
                Scope::sync();
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int e = tid * EPT + r;
                    const uint64_t o = sel[e ^ stride];
                    const bool keep_max = (((e & stride) == 0) == ((e & size) == 0));
                    v[r] = keep_max ? (v[r] > o ? v[r] : o) : (v[r] > o ? o : v[r]);
                }
            } else if (stride >= EPT) {                            // partner in another lane of this warp
                const int lm = stride / EPT;
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int e = tid * EPT + r;
                    const uint64_t o = shfl_xor_u64(v[r], lm);
                    const bool keep_max = (((e & stride) == 0) == ((e & size) == 0));
                    v[r] = keep_max ? (v[r] > o ? v[r] : o) : (v[r] > o ? o : v[r]);
                }
            } else {                                               // both entries in this thread's registers (compile-time indices)
                auto cx = [&](int a, int b) {
                    const bool desc = ((tid * EPT + a) & size) == 0;
                    const uint64_t mx = v[a] > v[b] ? v[a] : v[b], mn = v[a] > v[b] ? v[b] : v[a];
                    v[a] = desc ? mx : mn;
                    v[b] = desc ? mn : mx;
                };
                if constexpr (EPT == 4) {
                    if (stride == 2) { cx(0, 2); cx(1, 3); }
                    else { cx(0, 1); cx(2, 3); }
                } else if constexpr (EPT == 2) {
                    cx(0, 1);
                }
            }
        }
    }
    Scope::sync();
#pragma unroll
    for (int r = 0; r < EPT; ++r) sel[tid * EPT + r] = v[r];
    Scope::sync();
    (void)lane;
}

// returns false (nothing written) when a list would overflow: the caller then runs the general select
template <int NT, typename Src, typename Scope = CtaScope>
__device__ bool topk_big(const Src &src, int n, int k, int cap, uint32_t *keys, uint64_t *sel, uint32_t *hist, TkShared *sh,
                         float *out_s, int32_t *out_d) {
    static_assert(NT == 256, "eight warps: 8 bins per thread in the threshold scan, cap / 256 entries per thread in the sort");
    const int tid = Scope::tid(), lane = tid & 31, warp = tid >> 5;
    constexpr int BND_CAP = TK_BINS * 4 / 8;                       // the boundary list aliases the 8 KB histogram: 1,024 entries
    for (int i = tid; i < TK_BINS; i += NT) hist[i] = 0;
    sh->hist2[tid] = 0;
    if (tid == 0) { sh->sel_count = 0; sh->eq2_count = 0; sh->bnd_count = 0; }
    Scope::sync();
    for (int j4 = tid * 4; j4 < n; j4 += NT * 4) {                 // keys[] is padded to a multiple of 4
        float s4[4];
        src.score4(j4, n, s4);
        uint32_t k4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            k4[e] = float_to_ordered(s4[e]);
            if (j4 + e < n) atomicAdd(&hist[k4[e] >> 21], 1u);
        }
        *reinterpret_cast<uint4 *>(keys + j4) = make_uint4(k4[0], k4[1], k4[2], k4[3]);
    }
    Scope::sync();
    {   // threshold: thread t owns bins [2048 - 8(t+1), 2048 - 8t), counted from the top
        const uint4 *h4 = reinterpret_cast<const uint4 *>(hist + TK_BINS - 8 * (tid + 1));
        const uint4 lo4 = h4[0], hi4 = h4[1];
        const int local = (int)(lo4.x + lo4.y + lo4.z + lo4.w + hi4.x + hi4.y + hi4.z + hi4.w);
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sh->warp_tot[warp] = incl;
        Scope::sync();
        int before = 0;
#pragma unroll
        for (int w = 0; w < NT / 32 - 1; ++w) before += w < warp ? sh->warp_tot[w] : 0;
        incl += before;
        const int excl = incl - local;
        if (excl < k && k <= incl) {
            const uint32_t b8[8] = {hi4.w, hi4.z, hi4.y, hi4.x, lo4.w, lo4.z, lo4.y, lo4.x};      // top-down
            int running = excl, found = 0, fg = 0, fe = 0;
            bool done = false;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int hcnt = (int)b8[i];
                if (!done && running + hcnt >= k) { found = TK_BINS - 8 * tid - 1 - i; fg = running; fe = hcnt; done = true; }
                running += hcnt;
            }
            sh->found_bin = found; sh->found_gt = fg; sh->found_eq = fe;
        }
        Scope::sync();
    }
    const int d_bin = sh->found_bin, gt = sh->found_gt, eq = sh->found_eq;
    if (eq > BND_CAP) return false;                                // uniform
    uint64_t *bnd = reinterpret_cast<uint64_t *>(hist);
    for (int j4 = tid * 4; j4 < n; j4 += NT * 4) {
        const uint4 kv = *reinterpret_cast<const uint4 *>(keys + j4);
        const uint32_t k4[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (j4 + e < n) {
                const int bin = (int)(k4[e] >> 21);
                if (bin > d_bin) sel[atomicAdd(&sh->sel_count, 1)] = ((uint64_t)k4[e] << 32) | (uint32_t)(j4 + e);
                else if (bin == d_bin) bnd[atomicAdd(&sh->bnd_count, 1)] = ((uint64_t)k4[e] << 32) | (uint32_t)(j4 + e);
            }
        }
    }
    Scope::sync();
    // the boundary list: the next 8 key bits, then (rarely) a rank count on (key desc, docid asc)
    for (int t = tid; t < eq; t += NT) atomicAdd(&sh->hist2[(uint32_t)(bnd[t] >> 45) & 255u], 1u);
    Scope::sync();
    int d2, gt2, eq2;
    scan_down8(sh->hist2, lane, k - gt, d2, gt2, eq2);
    const int need2 = k - gt - gt2;                                // 1 <= need2 <= eq2
    if (need2 < eq2 && eq2 > 512) return false;                    // a long exact tie: the general select's docid radix pass (uniform)
    auto with_doc = [&](uint64_t e) { return (e & 0xffffffff00000000ull) | (uint32_t)~(uint32_t)src.doc((int)(uint32_t)e); };
    // entries above the second-level bin join sel; the tied ones are compacted to the FRONT of a second list.  That list lives in the
    // upper half of sel's capacity when it fits there (k <= cap), else right behind the boundary list: keep it simple — it reuses
    // bnd in place through a per-thread staging register (at most BND_CAP / NT = 4 entries per thread).
    uint64_t mine[BND_CAP / NT];
    int n_mine = 0;
#pragma unroll
    for (int u = 0; u < BND_CAP / NT; ++u) {
        const int t = tid + u * NT;
        mine[u] = 0ull;
        if (t < eq) {
            const uint64_t e = bnd[t];
            const int sub = (int)((uint32_t)(e >> 45) & 255u);
            if (sub > d2) sel[atomicAdd(&sh->sel_count, 1)] = e;
            else if (sub == d2) { mine[u] = e; n_mine |= 1 << u; }
        }
    }
    Scope::sync();                                                 // everyone has read its entries of bnd: it can be overwritten
    uint64_t *bnd2 = bnd;
#pragma unroll
    for (int u = 0; u < BND_CAP / NT; ++u)
        if (n_mine & (1 << u)) bnd2[atomicAdd(&sh->eq2_count, 1)] = need2 < eq2 ? with_doc(mine[u]) : mine[u];
    Scope::sync();
    if (need2 == eq2) {
        for (int t = tid; t < eq2; t += NT) sel[gt + gt2 + t] = bnd2[t];
    } else {
        for (int t = tid; t < eq2; t += NT) {                      // entries carry ~docid: (key desc, docid asc), identical pairs by list position
            const uint64_t me = bnd2[t];
            int rank = 0;
            for (int u = 0; u < eq2; ++u) {
                const uint64_t o = bnd2[u];
                rank += (o > me) || (o == me && u < t);
            }
            if (rank < need2) sel[gt + gt2 + rank] = me;
        }
    }
    Scope::sync();
    // (key, candidate index) -> (key, ~docid) for the entries that do not carry the docid yet, zero padding up to cap
    const bool tied = need2 < eq2;
    for (int i = tid; i < cap; i += NT) {
        uint64_t e = 0ull;
        if (i < k) {
            e = sel[i];
            if (!(tied && i >= gt + gt2)) e = with_doc(e);
        }
        sel[i] = e;
    }
    Scope::sync();
    if (cap == 1024) blocked_bitonic_desc<4, Scope>(sel, tid);
    else if (cap == 512) blocked_bitonic_desc<2, Scope>(sel, tid);
    else blocked_bitonic_desc<1, Scope>(sel, tid);
    for (int r = tid; r < k; r += NT) {
        const uint64_t v = sel[r];
        out_s[r] = ordered_to_float((uint32_t)(v >> 32));
        out_d[r] = (int32_t)(~(uint32_t)v);
    }
    return true;
}

template <int NT, bool STORE_KEYS, typename Src>
__device__ void topk_body(const Src &src, int n, int k, int cap, uint32_t *keys, uint64_t *sel, uint32_t *hist,
                          TkShared *sh, float *out_s, int32_t *out_d) {
    static_assert(NT == 256, "the bin ranges below assume eight warps");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n > k && cap > 128 && cap <= 1024 && STORE_KEYS) {         // large k: one-pass select + blocked bitonic sort
        if (topk_big<NT, Src>(src, n, k, cap, keys, sel, hist, sh, out_s, out_d)) return;
        __syncthreads();                                           // (a list would overflow: the general select, from scratch)
    }
    if (n <= k || cap > 128) {     // few candidates (take all) or large k: general path
        topk_general<NT>(src, n, k, cap, keys, sel, hist, sh, out_s, out_d);
        return;
    }
    for (int i = tid; i < TK_BINS; i += NT) hist[i] = 0;
    sh->hist2[tid] = 0;                                           // NT == 256 bins
    if (tid == 0) { sh->sel_count = 0; sh->eq2_count = 0; sh->bnd_count = 0; }
    __syncthreads();
    // pass 1: keys + histogram, four candidates per thread and iteration
    for (int j4 = tid * 4; j4 < n; j4 += NT * 4) {
        float s4[4];
        src.score4(j4, n, s4);
        uint32_t k4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            k4[e] = float_to_ordered(s4[e]);
            if (j4 + e < n) atomicAdd(&hist[k4[e] >> 21], 1u);
        }
        if (STORE_KEYS) *reinterpret_cast<uint4 *>(keys + j4) = make_uint4(k4[0], k4[1], k4[2], k4[3]);
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
    int above = 0, range = NT / 32 - 1;
    for (; range > 0; --range) {
        const int t = sh->warp_tot[range];
        if (above + t >= k) break;
        above += t;
    }
    int d_bin, gt, eq;
    scan_down8(hist + range * 256, lane, k - above, d_bin, gt, eq);
    d_bin += range * 256;
    gt += above;
    if (eq > TK_BND) {             // mass ties in the boundary bin: general path (uniform decision)
        __syncthreads();
        topk_general<NT>(src, n, k, cap, keys, sel, hist, sh, out_s, out_d);
        return;
    }
    // pass 2: classify.  sel[0, gt) <- keys above the boundary bin; bnd[0, eq) <- keys inside it (bnd aliases hist)
    uint64_t *bnd = reinterpret_cast<uint64_t *>(hist);           // 2 x TK_BND x 8 B = 4 KB <= the 8 KB histogram
    uint64_t *bnd2 = bnd + TK_BND;
    __syncthreads();                                              // everyone is done reading hist
    for (int j4 = tid * 4; j4 < n; j4 += NT * 4) {
        uint32_t k4[4];
        if (STORE_KEYS) {
            const uint4 kv = *reinterpret_cast<const uint4 *>(keys + j4);
            k4[0] = kv.x; k4[1] = kv.y; k4[2] = kv.z; k4[3] = kv.w;
        } else {
            float s4[4];
            src.score4(j4, n, s4);
#pragma unroll
            for (int e = 0; e < 4; ++e) k4[e] = float_to_ordered(s4[e]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (j4 + e < n) {
                const uint32_t key = k4[e];
                const int bin = (int)(key >> 21);
                if (bin > d_bin) sel[atomicAdd(&sh->sel_count, 1)] = ((uint64_t)key << 32) | (uint32_t)(j4 + e);
                else if (bin == d_bin) {
                    bnd[atomicAdd(&sh->bnd_count, 1)] = ((uint64_t)key << 32) | (uint32_t)(j4 + e);
                    atomicAdd(&sh->hist2[(key >> 13) & 255u], 1u);
                }
            }
        }
    }
    __syncthreads();
    // cut the boundary list with the next 8 key bits
    int d2, gt2, eq2;
    scan_down8(sh->hist2, lane, k - gt, d2, gt2, eq2);
    const int need2 = k - gt - gt2;                               // 1 <= need2 <= eq2
    if (tid < eq) {
        const uint64_t e = bnd[tid];
        const int sub = (int)((e >> 45) & 255u);
        if (sub > d2) sel[atomicAdd(&sh->sel_count, 1)] = e;
        else if (sub == d2) bnd2[atomicAdd(&sh->eq2_count, 1)] = e;
    }
    __syncthreads();
    if (need2 == eq2) {
        if (tid < eq2) sel[gt + gt2 + tid] = bnd2[tid];
    } else {
        // still tied after 19 key bits: order by (key desc, docid asc); identical (key, docid) pairs by list position
        uint64_t mine = 0, orig = 0;
        if (tid < eq2) {
            orig = bnd2[tid];
            mine = (orig & 0xffffffff00000000ull) | (uint32_t)~(uint32_t)src.doc((int)(uint32_t)orig);
            bnd2[tid] = mine;                                     // own slot only; others read it after the barrier
        }
        __syncthreads();
        if (tid < eq2) {
            int rank = 0;
            for (int u = 0; u < eq2; ++u) {
                const uint64_t o = bnd2[u];
                rank += (o > mine) || (o == mine && u < tid);
            }
            if (rank < need2) sel[gt + gt2 + rank] = orig;
        }
    }
    __syncthreads();
    // one warp: (key, candidate index) -> (key, ~docid), sort the k survivors, write them out
    if (warp == 0) {
        uint64_t v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane * 4 + r;
            v[r] = 0;
            if (i < k) {
                const uint64_t e = sel[i];
                v[r] = (e & 0xffffffff00000000ull) | (uint32_t)~(uint32_t)src.doc((int)(uint32_t)e);
            }
        }
        warp_bitonic128_desc(v, lane);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane * 4 + r;
            if (i < k) {
                out_s[i] = ordered_to_float((uint32_t)(v[r] >> 32));
                out_d[i] = (int32_t)(~(uint32_t)v[r]);
            }
        }
    }
}


// The general select as the COLD fallback of the small-footprint fast path: kept out of line so that its two call sites
// do not triple the hot kernel's instruction-cache footprint.
template <int NT, typename Src, typename Scope = CtaScope>
__device__ __noinline__ void topk_general_cold(const Src &src, int n, int k, int cap, uint32_t *keys, uint64_t *sel, uint32_t *hist,
                                               TkShared *sh, float *out_s, int32_t *out_d) {
    topk_general<NT, Src, Scope>(src, n, k, cap, keys, sel, hist, sh, out_s, out_d);
}

constexpr int TKF_THREADS = 128;
constexpr int TKF_R4 = 5;            // 128 threads x 5 x 4 = 2,560 candidates held in registers
// topk_fast16 as the COLD path of kernels whose hot path is the lean select: out of line, so that its registers and spills do not
// shape the hot code
template <int NT, int R4, typename Src, typename Scope>
__device__ __noinline__ void topk_fast16_cold(const Src &src, int n, int k, uint32_t *gkeys, uint32_t *ghist, uint64_t *sel, uint32_t *hist_words,
                                              TkShared *sh, float *out_s, int32_t *out_d);
constexpr int TKF64_THREADS = 64;    // the fused kernel's groups: two warps per query, no keys in registers, 1,024 first-level bins
constexpr int TKF64_BITS = 10;

// ---- lean select (round 2) -----------------------------------------------------------------------------------------------
// The same result as topk_fast16 for the reference's regime (k <= 128, k < n <= 2,560 candidates, 128 threads) with half the
// instructions and a third of the dependent chain (ncu per-line profile of k_topk_fast, profiles/r02_topk_lines.txt: 8,000 warp
// instructions per query, 1,450 of them the final sort on ONE warp, 2,500 the two classify loops, 800 the threshold walk):
//   * 1,024 first-level bins of 32 bits (10 key bits): one shift and one RED per candidate, the threshold is found by ONE
//     block-wide top-down scan (thread t owns bins [1024 - 8(t+1), 1024 - 8t)): two barriers, no 2,048-bin walk
//   * candidates are classified from the keys in registers with predicated code only (count, two atomics per thread, store)
//   * the second-level histogram is built from the boundary list (<= 256 entries), not inside the classify loop
//   * the k survivors are sorted by ALL FOUR warps: each warp sorts its 32 entries with a shuffle bitonic network (15 stages,
//     one entry per lane), the four sorted runs are merged by rank — every thread finds its entry's position in the other three
//     runs with a 5-step binary search — and written straight to the output row
// Everything else (order-preserving keys, (key desc, docid asc) ties, fallbacks for n <= k, mass ties, n > 2,560) is shared
// with / delegated to topk_fast16, so the two produce identical bits (tests/test_gpu_pipeline.py, tests/test_gpu_parity.py).
constexpr int TKL_BITS = 10, TKL_BINS = 1 << TKL_BITS, TKL_SH1 = 32 - TKL_BITS, TKL_SH2 = TKL_SH1 - 8;
static_assert(TKL_BINS * 4 == TK_BINS * 2, "the lean select's 1,024 32-bit bins live in the 4 KB of the 2,048 16-bit bins");

template <typename Src, typename Scope>
__device__ __forceinline__ bool topk_lean_eligible(int n, int k) { return n > k && n <= TKF_THREADS * 4 * TKF_R4 && k <= 128; }

template <typename Src, typename Scope = CtaScope>
__device__ void topk_lean128(const Src &src, int n, int k, uint32_t *gkeys, uint32_t *ghist, uint64_t *sel, uint32_t *hist,
                             TkShared *sh, float *out_s, int32_t *out_d) {
    constexpr int NT = TKF_THREADS, R4 = TKF_R4;
    int tid_;
    if constexpr (std::is_same<Scope, CtaScope>::value) tid_ = threadIdx.x; else tid_ = Scope::tid();
    const int tid = tid_, lane = tid & 31, warp = tid >> 5;
    // (the caller has STARTED filling co / cbase / bias in shared memory and has not synchronised: the first barrier covers it)
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4 *>(hist)[tid] = z;                  // 1,024 words = 256 x 16 bytes
        reinterpret_cast<uint4 *>(hist)[tid + NT] = z;
        sh->hist2[tid] = 0u;
        sh->hist2[tid + NT] = 0u;
        if (tid == 0) { sh->sel_count = 0; sh->eq2_count = 0; sh->bnd_count = 0; }
    }
    float4 v[R4];
#pragma unroll
    for (int r = 0; r < R4; ++r) {
        const int j4 = (r * NT + tid) * 4;
        v[r] = j4 < n ? src.load4(j4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    Scope::sync();                                                 // bins are zero, the caller's segment tables are in place
    uint32_t key[R4][4];
#pragma unroll
    for (int r = 0; r < R4; ++r) {
        const int j4 = (r * NT + tid) * 4;
        float s4[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
        if (j4 < n) src.bias4(v[r], j4, n, s4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            key[r][e] = float_to_ordered(s4[e]);
            if (j4 + e < n) atomicAdd(&hist[key[r][e] >> TKL_SH1], 1u);
        }
    }
    Scope::sync();
    // ---- threshold: the bin where the count accumulated from the top reaches k
    {
        const uint4 *h4 = reinterpret_cast<const uint4 *>(hist + TKL_BINS - 8 * (tid + 1));
        const uint4 lo4 = h4[0], hi4 = h4[1];                      // bins ascending: lo4.x is the thread's LOWEST bin
        const int local = (int)(lo4.x + lo4.y + lo4.z + lo4.w + hi4.x + hi4.y + hi4.z + hi4.w);
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sh->warp_tot[warp] = incl;
        Scope::sync();
        int before = 0;
#pragma unroll
        for (int w = 0; w < NT / 32 - 1; ++w) before += w < warp ? sh->warp_tot[w] : 0;
        incl += before;
        const int excl = incl - local;                             // candidates in bins above this thread's eight
        if (excl < k && k <= incl) {
            const uint32_t b8[8] = {hi4.w, hi4.z, hi4.y, hi4.x, lo4.w, lo4.z, lo4.y, lo4.x};      // top-down
            int running = excl, found = 0, fg = 0, fe = 0;
            bool done = false;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int hcnt = (int)b8[i];
                if (!done && running + hcnt >= k) { found = TKL_BINS - 8 * tid - 1 - i; fg = running; fe = hcnt; done = true; }
                running += hcnt;
            }
            sh->found_bin = found; sh->found_gt = fg; sh->found_eq = fe;
        }
        Scope::sync();
    }
    const int d_bin = sh->found_bin, gt = sh->found_gt, eq = sh->found_eq;
    if (eq > TK_BND) {                        // mass ties in the boundary bin: the general select (uniform decision)
        Scope::sync();
        topk_general_cold<NT, Src, Scope>(src, n, k, 128, gkeys, sel, ghist, sh, out_s, out_d);
        return;
    }
    // ---- classify from the registers: sel[0, gt) <- keys above the boundary bin, bnd[0, eq) <- keys inside it (bnd aliases the bins:
    // nobody reads them after the barrier above)
    uint64_t *bnd = reinterpret_cast<uint64_t *>(hist);            // 2 x TK_BND x 8 B = the 4 KB of the bins
    uint64_t *bnd2 = bnd + TK_BND;
    {
        int c_sel = 0, c_bnd = 0;
#pragma unroll
        for (int r = 0; r < R4; ++r) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool valid = (r * NT + tid) * 4 + e < n;
                const int bin = (int)(key[r][e] >> TKL_SH1);
                c_sel += (valid && bin > d_bin) ? 1 : 0;
                c_bnd += (valid && bin == d_bin) ? 1 : 0;
            }
        }
        int at_sel = c_sel ? atomicAdd(&sh->sel_count, c_sel) : 0;
        int at_bnd = c_bnd ? atomicAdd(&sh->bnd_count, c_bnd) : 0;
        if (c_sel | c_bnd) {
#pragma unroll
            for (int r = 0; r < R4; ++r) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = (r * NT + tid) * 4 + e;
                    const int bin = j < n ? (int)(key[r][e] >> TKL_SH1) : -1;
                    const uint64_t ent = ((uint64_t)key[r][e] << 32) | (uint32_t)j;
                    if (bin > d_bin) sel[at_sel++] = ent;
                    else if (bin == d_bin) bnd[at_bnd++] = ent;
                }
            }
        }
    }
    Scope::sync();
    // ---- second level: the next 8 key bits of the boundary list; every listed candidate becomes (key, ~docid) on the way (the docid
    // reads — segment search + one global load each — are issued together, one or two per thread)
    auto with_doc = [&](uint64_t e) { return (e & 0xffffffff00000000ull) | (uint32_t)~(uint32_t)src.doc((int)(uint32_t)e); };
    for (int t = tid; t < eq; t += NT) atomicAdd(&sh->hist2[(uint32_t)(bnd[t] >> (32 + TKL_SH2)) & 255u], 1u);
    if (tid < gt) sel[tid] = with_doc(sel[tid]);                   // gt < k <= 128 = NT; slots >= gt are appended below
    uint64_t mine_b[TK_BND / NT];
#pragma unroll
    for (int u = 0; u < TK_BND / NT; ++u) mine_b[u] = tid + u * NT < eq ? with_doc(bnd[tid + u * NT]) : 0ull;
    Scope::sync();
    int d2, gt2, eq2;
    scan_down8(sh->hist2, lane, k - gt, d2, gt2, eq2);
    const int need2 = k - gt - gt2;                                // 1 <= need2 <= eq2
#pragma unroll
    for (int u = 0; u < TK_BND / NT; ++u) {
        if (tid + u * NT < eq) {
            const int sub = (int)((uint32_t)(mine_b[u] >> (32 + TKL_SH2)) & 255u);
            if (sub > d2) sel[atomicAdd(&sh->sel_count, 1)] = mine_b[u];
            else if (sub == d2) bnd2[atomicAdd(&sh->eq2_count, 1)] = mine_b[u];
        }
    }
    Scope::sync();
    if (need2 == eq2) {
        for (int t = tid; t < eq2; t += NT) sel[gt + gt2 + t] = bnd2[t];
    } else {
        // still tied after 18 key bits: order by (key desc, docid asc); identical (key, docid) pairs by list position
        for (int t = tid; t < eq2; t += NT) {
            const uint64_t mine = bnd2[t];
            int rank = 0;
            for (int u = 0; u < eq2; ++u) {
                const uint64_t o = bnd2[u];
                rank += (o > mine) || (o == mine && u < t);
            }
            if (rank < need2) sel[gt + gt2 + rank] = mine;
        }
    }
    Scope::sync();
    // ---- sort: every warp sorts 32 entries (shuffle bitonic network, descending, one entry per lane), then the four runs are merged by rank
    uint64_t x = tid < k ? sel[tid] : 0ull;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const uint64_t other = shfl_xor_u64(x, stride);
            const bool desc = (lane & size) == 0;                  // size == 32: every lane
            const bool lower = (lane & stride) == 0;
            const uint64_t mx = x > other ? x : other, mn = x > other ? other : x;
            x = (lower == desc) ? mx : mn;
        }
    }
    Scope::sync();                                                 // everyone holds its entry: sel can be overwritten
    sel[tid] = x;
    Scope::sync();
    int rank = lane;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        if (w == warp) continue;
        const uint64_t *run = sel + w * 32;                        // descending
        int lo = 0, hi = 32;                                       // number of entries of `run` that come before x
#pragma unroll
        for (int it = 0; it < 6; ++it) {
            if (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const uint64_t y = run[mid];
                const bool before = (y > x) || (y == x && w < warp);
                lo = before ? mid + 1 : lo;
                hi = before ? hi : mid;
            }
        }
        rank += lo;
    }
    if (rank < k) {
        out_s[rank] = ordered_to_float((uint32_t)(x >> 32));
        out_d[rank] = (int32_t)(~(uint32_t)x);
    }
}

// ---- fast path, small-footprint variant --------------------------------------------------------------------------
// Same algorithm as topk_body<false> for CTAs of NT = 128 threads with 16-bit histogram bins (n <= 65,535): 6.4 KB of
// shared memory and 4 K registers per query.  A batch's top-k overlaps the NEXT batch's scoring kernel, whose CTA leaves
// ~53 KB of shared memory and ~32 K registers per SM; every query's top-k is a chain of dependent L2 round trips and
// block barriers (latency-bound: ~25 us per query under a saturated memory system), so what matters is how many
// queries are in flight per SM: eight of these CTAs fit beside the scoring CTA against four of the 256-thread ones.
// The rare fallbacks (n <= k, mass ties in the boundary bin) run the general select with its histogram in global scratch.
// BITS = width of the first-level digit (11: the 2,048 bins of the stand-alone kernels; 10: 1,024 bins = 2 KB per group for the
// 64-thread groups of the fused kernel, whose eleven slices must fit beside the scoring ring).  The result does not depend on it.
template <int NT, int R4, typename Src, typename Scope = CtaScope, int BITS = 11>
__device__ void topk_fast16(const Src &src, int n, int k, uint32_t *gkeys, uint32_t *ghist, uint64_t *sel, uint32_t *hist_words,
                            TkShared *sh, float *out_s, int32_t *out_d, uint32_t dbg = 0) {
    constexpr int NW = NT / 32;
    constexpr int NBINS = 1 << BITS;
    constexpr int SH1 = 32 - BITS;                                 // key >> SH1 = first-level bin
    constexpr int SH2 = SH1 - 8;                                   // (key >> SH2) & 255 = second-level bin
    constexpr int RANGE = NBINS / NW;                              // bins summed by one warp (512 for four warps)
    static_assert(RANGE % 256 == 0, "a warp's range is walked in 256-bin blocks");
    // 128-bit score loads a thread keeps in flight in the passes that do not hold keys in registers (two-warp groups: every
    // exposed L2 round trip counts; the 128-thread kernel keeps its one-load loop and its register allocation)
    constexpr int TK_U = NT == 64 ? 4 : 1;
    int tid_;
    if constexpr (std::is_same<Scope, CtaScope>::value) tid_ = threadIdx.x; else tid_ = Scope::tid();
    const int tid = tid_, lane = tid & 31, warp = tid >> 5;
    // The caller has STARTED filling co / cbase / bias in shared memory and has not synchronised: the first barrier below
    // covers that fill too, so the score loads (which need only n) are in flight together with the caller's loads.
    if (n <= k) {
        Scope::sync();
        topk_general_cold<NT, Src, Scope>(src, n, k, 128, gkeys, sel, ghist, sh, out_s, out_d);
        return;
    }
    uint16_t *hist = reinterpret_cast<uint16_t *>(hist_words);     // bin b = half (b & 1) of word b >> 1
    for (int i = tid; i < NBINS / 2; i += NT) hist_words[i] = 0;
    for (int i = tid; i < 256; i += NT) sh->hist2[i] = 0;
    if (tid == 0) { sh->sel_count = 0; sh->eq2_count = 0; sh->bnd_count = 0; }
    // Up to NT * 4 * R4 candidates (2,560: the reference's beam 20 x ~107-doc clusters) are read ONCE, all loads of a
    // thread in flight together, and their keys stay in registers for the second pass; a query's top-k is a chain of
    // dependent steps on few warps, so every exposed L2 round trip (one per loop iteration otherwise) is what it costs.
    const bool in_regs = R4 > 0 && n <= NT * 4 * R4;
    uint32_t kreg[R4 > 0 ? R4 : 1][4];
    if (in_regs) {
        float4 v[R4 > 0 ? R4 : 1];
#pragma unroll
        for (int r = 0; r < R4; ++r) {
            const int j4 = (r * NT + tid) * 4;
            v[r] = j4 < n ? src.load4(j4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        Scope::sync();                                           // histogram is zeroed
#pragma unroll
        for (int r = 0; r < R4; ++r) {
            const int j4 = (r * NT + tid) * 4;
            float s4[4];
            if (j4 < n) src.bias4(v[r], j4, n, s4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                kreg[r][e] = float_to_ordered(s4[e]);
                const uint32_t bin = kreg[r][e] >> SH1;
                if (j4 + e < n) atomicAdd(&hist_words[bin >> 1], 1u << ((bin & 1u) * 16u));
            }
        }
    } else {
        // candidates are not kept: both passes read the (L2-resident) scores, TK_U 128-bit loads of a thread in flight together
        Scope::sync();
        for (int j0 = 0; j0 < n; j0 += NT * 4 * TK_U) {
            float4 v[TK_U];
#pragma unroll
            for (int u = 0; u < TK_U; ++u) {
                const int j4 = j0 + (u * NT + tid) * 4;
                v[u] = j4 < n ? src.load4(j4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < TK_U; ++u) {
                const int j4 = j0 + (u * NT + tid) * 4;
                float s4[4];
                if (j4 < n) src.bias4(v[u], j4, n, s4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t bin = float_to_ordered(s4[e]) >> SH1;
                    if (j4 + e < n) atomicAdd(&hist_words[bin >> 1], 1u << ((bin & 1u) * 16u));
                }
            }
        }
    }
    Scope::sync();
    if (dbg & 2u) return;                     // GDR_TOPK_DEBUG timing experiments: stop after pass 1
    {
        int part = 0;
#pragma unroll
        for (int i = 0; i < RANGE / 64; ++i) {
            const uint32_t w = hist_words[warp * (RANGE / 2) + i * 32 + lane];
            part += (int)(w & 0xffffu) + (int)(w >> 16);
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        if (lane == 0) sh->warp_tot[warp] = part;
    }
    Scope::sync();
    int above = 0, range = NW - 1;
    for (; range > 0; --range) {
        const int t = sh->warp_tot[range];
        if (above + t >= k) break;
        above += t;
    }
    int base = range * RANGE + RANGE - 256;                        // walk the range's 256-bin blocks from the top
    for (; base > range * RANGE; base -= 256) {
        const uint32_t *w = hist_words + base / 2 + lane * 4;
        int t = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) t += (int)(w[i] & 0xffffu) + (int)(w[i] >> 16);
#pragma unroll
        for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (above + t >= k) break;
        above += t;
    }
    int d_bin, gt, eq;
    scan_down8(hist + base, lane, k - above, d_bin, gt, eq);
    d_bin += base;
    gt += above;
    if (eq > TK_BND) {
        Scope::sync();
        topk_general_cold<NT, Src, Scope>(src, n, k, 128, gkeys, sel, ghist, sh, out_s, out_d);
        return;
    }
    uint64_t *bnd = reinterpret_cast<uint64_t *>(hist_words);      // 2 x TK_BND x 8 B = the 4 KB histogram
    uint64_t *bnd2 = bnd + TK_BND;
    Scope::sync();
    auto classify = [&](uint32_t key, int j) {
        const int bin = (int)(key >> SH1);
        if (bin > d_bin) sel[atomicAdd(&sh->sel_count, 1)] = ((uint64_t)key << 32) | (uint32_t)j;
        else if (bin == d_bin) {
            bnd[atomicAdd(&sh->bnd_count, 1)] = ((uint64_t)key << 32) | (uint32_t)j;
            atomicAdd(&sh->hist2[(key >> SH2) & 255u], 1u);
        }
    };
    if (in_regs) {
        // count this thread's hits first, reserve their slots with ONE atomic per list, then store: per-candidate atomics
        // on the two list counters cost ~20 instructions each after the compiler's warp aggregation, 20 times per thread
        int c_sel = 0, c_bnd = 0;
#pragma unroll
        for (int r = 0; r < R4; ++r) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int bin = (r * NT + tid) * 4 + e < n ? (int)(kreg[r][e] >> SH1) : -1;
                c_sel += bin > d_bin;
                c_bnd += bin == d_bin;
            }
        }
        int at_sel = c_sel ? atomicAdd(&sh->sel_count, c_sel) : 0;
        int at_bnd = c_bnd ? atomicAdd(&sh->bnd_count, c_bnd) : 0;
        if (c_sel | c_bnd) {
#pragma unroll
            for (int r = 0; r < R4; ++r) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = (r * NT + tid) * 4 + e;
                    const uint32_t key = kreg[r][e];
                    const int bin = j < n ? (int)(key >> SH1) : -1;
                    if (bin > d_bin) sel[at_sel++] = ((uint64_t)key << 32) | (uint32_t)j;
                    else if (bin == d_bin) {
                        bnd[at_bnd++] = ((uint64_t)key << 32) | (uint32_t)j;
                        atomicAdd(&sh->hist2[(key >> SH2) & 255u], 1u);
                    }
                }
            }
        }
    } else {
        for (int j0 = 0; j0 < n; j0 += NT * 4 * TK_U) {
            float4 v[TK_U];
#pragma unroll
            for (int u = 0; u < TK_U; ++u) {
                const int j4 = j0 + (u * NT + tid) * 4;
                v[u] = j4 < n ? src.load4(j4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < TK_U; ++u) {
                const int j4 = j0 + (u * NT + tid) * 4;
                float s4[4];
                if (j4 < n) src.bias4(v[u], j4, n, s4);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (j4 + e < n) classify(float_to_ordered(s4[e]), j4 + e);
            }
        }
    }
    Scope::sync();
    if (dbg & 4u) return;                     // stop after pass 2
    // Every listed candidate becomes (key, ~docid) HERE, one or two entries per thread with all lanes busy: the docid
    // reads (segment search + one global load each) are issued together and overlap the second-level scan, instead of one
    // exposed round trip in front of the final sort.
    auto with_doc = [&](uint64_t e) { return (e & 0xffffffff00000000ull) | (uint32_t)~(uint32_t)src.doc((int)(uint32_t)e); };
    for (int t = tid; t < gt; t += NT) sel[t] = with_doc(sel[t]);  // gt < k <= 128; slots >= gt are appended below
    int d2, gt2, eq2;
    scan_down8(sh->hist2, lane, k - gt, d2, gt2, eq2);
    const int need2 = k - gt - gt2;
    for (int t = tid; t < eq; t += NT) {
        const uint64_t e = with_doc(bnd[t]);
        const int sub = (int)((e >> (32 + SH2)) & 255u);
        if (sub > d2) sel[atomicAdd(&sh->sel_count, 1)] = e;
        else if (sub == d2) bnd2[atomicAdd(&sh->eq2_count, 1)] = e;
    }
    Scope::sync();
    if (need2 == eq2) {
        for (int t = tid; t < eq2; t += NT) sel[gt + gt2 + t] = bnd2[t];
    } else {
        // still tied after 19 key bits: order by (key desc, docid asc); identical (key, docid) pairs by list position
        for (int t = tid; t < eq2; t += NT) {
            const uint64_t mine = bnd2[t];
            int rank = 0;
            for (int v = 0; v < eq2; ++v) {
                const uint64_t o = bnd2[v];
                rank += (o > mine) || (o == mine && v < t);
            }
            if (rank < need2) sel[gt + gt2 + rank] = mine;
        }
    }
    Scope::sync();
    if (dbg & 8u) return;                     // stop before the docid gather + sort + output
    if (warp == 0) {
        uint64_t v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane * 4 + r;
            v[r] = 0;
            if (i < k) v[r] = sel[i];
        }
        warp_bitonic128_desc(v, lane);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane * 4 + r;
            if (i < k) {
                out_s[i] = ordered_to_float((uint32_t)(v[r] >> 32));
                out_d[i] = (int32_t)(~(uint32_t)v[r]);
            }
        }
    }
}


template <int NT, int R4, typename Src, typename Scope>
__device__ __noinline__ void topk_fast16_cold(const Src &src, int n, int k, uint32_t *gkeys, uint32_t *ghist, uint64_t *sel, uint32_t *hist_words,
                                              TkShared *sh, float *out_s, int32_t *out_d) {
    topk_fast16<NT, R4, Src, Scope>(src, n, k, gkeys, ghist, sel, hist_words, sh, out_s, out_d, 0u);
}

// ---- grouped variant -----------------------------------------------------------------------------------------------------
// G independent NT-thread groups per CTA (GroupScope: own named barrier, own shared-memory slice), each claiming one query
// at a time from a global counter and running topk_fast16 on it.  This loop is what the top-k warps of the fused scoring +
// top-k CTA execute (score_fused.cu); stand-alone it is k_topk_fast_grouped (topk_grouped.cu), the check that the group-scoped
// select equals k_topk_fast.  NT = 128: four warps per query, keys of up to 2,560 candidates in registers; NT = 64: two warps
// per query and no keys in registers (both passes read the L2-resident scores) — half the threads and registers per query in
// flight, which is what lets one fused CTA hold nine queries at a time.
// slice: sel[128] u64 | hist[2048] u16 | co[K+1] | cbase[K] | bias[K] | TkShared | next query (int), rounded up to 16 bytes
__host__ __device__ inline int tkg_slice_bytes(int K) {
    return (128 * 8 + TK_BINS * 2 + (3 * K + 1) * 4 + (int)sizeof(TkShared) + 4 + 15) / 16 * 16;
}

template <int FIRST, int NT = TKF_THREADS>
__device__ void topk_group_loop(const ScoreArgs &a, float alpha, float *out_scores, int32_t *out_docids, unsigned char *slice) {
    using S = GroupScope<FIRST, NT>;
    constexpr int R4 = NT == TKF_THREADS ? TKF_R4 : 0;
    const int tid = S::tid();
    uint64_t *sel = reinterpret_cast<uint64_t *>(slice);
    uint32_t *hist_words = reinterpret_cast<uint32_t *>(sel + 128);
    int32_t *co = reinterpret_cast<int32_t *>(hist_words + TK_BINS / 2);
    int32_t *cbase = co + a.K + 1;
    float *bias = reinterpret_cast<float *>(cbase + a.K);
    TkShared *sh = reinterpret_cast<TkShared *>(bias + a.K);
    volatile int *next = reinterpret_cast<int *>(sh + 1);
    if (a.n_ranks > 1 && tid == 0) wait_for_scorers(a);            // sharded corpus (the first barrier below orders the group behind it)
    for (;;) {
        if (tid == 0) *next = atomicAdd(&a.counters[CTR_TOPK_NEXT], 1);
        S::sync();
        const int b = *next;
        if (b >= a.B_top) break;                                   // group-uniform
        const int bg = b + a.q_base;                               // row of candoff / cbase / prob (global batch index)
        for (int i = tid; i <= a.K; i += NT) {
            co[i] = a.candoff[(int64_t)bg * (a.K + 1) + i];
            if (i < a.K) {
                cbase[i] = a.cbase[(int64_t)bg * a.K + i];
                if (a.prob) bias[i] = __fmul_rn(alpha, a.prob[(int64_t)bg * a.K + i]);
            }
        }
        const int n = a.candoff[(int64_t)bg * (a.K + 1) + a.K];
        StoreSrc src{a.scorebuf + (int64_t)b * a.stride, co, cbase, a.prob ? bias : nullptr, a.docid, a.K};
        bool lean = false;
        if constexpr (NT == TKF_THREADS) lean = topk_lean_eligible<StoreSrc, S>(n, a.k);
        if constexpr (NT == TKF_THREADS) {
            if (lean) topk_lean128<StoreSrc, S>(src, n, a.k, a.gkeys + (int64_t)b * a.stride, a.ghist + (int64_t)b * TK_BINS, sel, hist_words, sh,
                                                out_scores + (int64_t)b * a.k, out_docids + (int64_t)b * a.k);
        }
        if (!lean) {
            if constexpr (NT == TKF_THREADS)
                topk_fast16_cold<NT, R4, StoreSrc, S>(src, n, a.k, a.gkeys + (int64_t)b * a.stride, a.ghist + (int64_t)b * TK_BINS, sel, hist_words,
                                                      sh, out_scores + (int64_t)b * a.k, out_docids + (int64_t)b * a.k);
            else
                topk_fast16<NT, R4, StoreSrc, S>(src, n, a.k, a.gkeys + (int64_t)b * a.stride, a.ghist + (int64_t)b * TK_BINS, sel, hist_words,
                                                 sh, out_scores + (int64_t)b * a.k, out_docids + (int64_t)b * a.k, 0u);
        }
        S::sync();                                                 // the slice (and *next) is free for the next query
    }
}

}  // namespace gdr
