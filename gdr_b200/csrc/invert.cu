// Pair inversion: turn the [B, K] beam table (query-major) into cluster -> (query, beam) groups
// and emit the work lists the scoring kernels walk.  Replaces the host-side dict walk of
// main_models.py:1435-1443 (flatten clusters, id_mapping lookup per (query, beam)).
//
//   k_count : one warp per query.  cnt[c] += 1 per valid beam; candoff[b, i] = start of beam i's
//             segment in query b's candidate list (beam-major order, as main_models.py:1441-1443).
//   k_scan  : one CTA.  Exclusive scans over clusters of the group sizes and of the number of
//             SIMT / tcgen05 work items each cluster produces.
//   k_fill  : scatter pair ids into their groups (atomicSub on cnt, which returns it to all-zero
//             for the next call) and write the work items.
#include "gdr_common.cuh"

namespace gdr {

// One warp: rows [row0, row0 + nrows) of q -> a.qsplit[row][term][dim] (the B operand of the tcgen05 path)
__device__ __forceinline__ void split_rows(const ScoreArgs &a, int64_t row0, int nrows, int lane) {
    for (int r = 0; r < nrows; ++r) {
        const float *src = a.q + (row0 + r) * a.dim;
        __nv_bfloat16 *dst = a.qsplit + (row0 + r) * 3 * a.dim;
        for (int e = lane * 4; e < a.dim; e += 128) {
            const float4 v = *reinterpret_cast<const float4 *>(src + e);
            const float x[4] = {v.x, v.y, v.z, v.w};
            __nv_bfloat16 t[3][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split3(x[i], t[0][i], t[1][i], t[2][i]);
#pragma unroll
            for (int term = 0; term < 3; ++term)
                *reinterpret_cast<uint2 *>(dst + term * a.dim + e) = *reinterpret_cast<const uint2 *>(t[term]);
        }
    }
}

// One warp: count query b's beams into cnt[] and write its candidate-segment starts.
__device__ __forceinline__ void count_query(const ScoreArgs &a, int b, int lane, int32_t *cnt) {
    if (a.qsplit) {
        if (a.flags & GDR_Q_PER_BEAM) split_rows(a, (int64_t)b * a.K, a.K, lane);
        else split_rows(a, b, 1, lane);
    }
    const int32_t *beams = a.beams + (int64_t)b * a.K;
    int32_t *co = a.candoff + (int64_t)b * (a.K + 1);
    int32_t *cb = a.cbase + (int64_t)b * a.K;
    int carry = 0;
    for (int i0 = 0; i0 < a.K; i0 += 32) {
        const int i = i0 + lane;
        int sz = 0;
        if (i < a.K) {
            const int c = beams[i];
            int row_lo = 0;
            if (c >= 0 && c < a.n_clusters) {
                row_lo = a.offsets[c];
                sz = a.offsets[c + 1] - row_lo;
                if (c >= a.c_lo && c < a.c_hi) atomicAdd(&cnt[c], 1);     // sharded corpus: only the owned clusters form groups here
            }
            cb[i] = row_lo;               // the top-k maps a winning candidate back to its store row without touching beams/offsets
        }
        int incl = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (i < a.K) co[i] = carry + incl - sz;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) co[a.K] = carry;
}

__global__ void __launch_bounds__(128) k_count(ScoreArgs a) {
    pdl_launch_dependents();
    pdl_wait();                 // candoff is still read by the previous call's top-k
    trace_start(a.dbg, 0);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp < a.B) count_query(a, warp, threadIdx.x & 31, a.cnt);
}

__device__ __forceinline__ int ceil_div(int x, int y) { return (x + y - 1) / y; }

__device__ __forceinline__ void item_counts(const ScoreArgs &a, int g, int size, int &ns, int &nu) {
    if (g >= a.umma_min_group) nu = ceil_div(size, UMMA_ROWS) * ceil_div(g, UMMA_NQ);
    else ns = ceil_div(size, SIMT_ROWS) * ceil_div(g, SIMT_QT);
}

// Work items of cluster c: row tile outer, query chunk inner, so consecutive items re-read the same rows.
__device__ __forceinline__ void write_items(const ScoreArgs &a, int c, int g0, int g, int simt_at, int umma_at) {
    const int row_lo = a.offsets[c];
    const int size = a.offsets[c + 1] - row_lo;
    if (g <= 0 || size <= 0) return;
    const bool umma = g >= a.umma_min_group;
    const int rows_per = umma ? UMMA_ROWS : SIMT_ROWS;
    const int q_per = umma ? UMMA_NQ : SIMT_QT;
    Item *dst = umma ? a.umma_items + umma_at : a.simt_items + simt_at;
    for (int r = 0; r < size; r += rows_per) {
        const int nrows = min(rows_per, size - r);
        for (int s = 0; s < g; s += q_per) {
            Item it;
            it.row0 = row_lo + r - a.row_lo;        // row inside this handle's emb (a shard holds global rows [row_lo, ...))
            it.rel0 = r;
            it.nrows_nq = nrows | (min(q_per, g - s) << 16);
            it.slot0 = g0 + s;
            *dst++ = it;
        }
    }
}

// Block-wide exclusive scan of three ints per thread (up to 1024 threads), returns totals via `tot`.
__device__ __forceinline__ void block_scan3(int &x, int &y, int &z, int tot[3], int (*wsum)[3]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ix = x, iy = y, iz = z;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int tx = __shfl_up_sync(0xffffffffu, ix, d);
        int ty = __shfl_up_sync(0xffffffffu, iy, d);
        int tz = __shfl_up_sync(0xffffffffu, iz, d);
        if (lane >= d) { ix += tx; iy += ty; iz += tz; }
    }
    if (lane == 31) { wsum[warp][0] = ix; wsum[warp][1] = iy; wsum[warp][2] = iz; }
    __syncthreads();
    const int nwarps = (blockDim.x + 31) >> 5;
    if (warp == 0) {
        int sx = lane < nwarps ? wsum[lane][0] : 0, sy = lane < nwarps ? wsum[lane][1] : 0, sz = lane < nwarps ? wsum[lane][2] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int tx = __shfl_up_sync(0xffffffffu, sx, d);
            int ty = __shfl_up_sync(0xffffffffu, sy, d);
            int tz = __shfl_up_sync(0xffffffffu, sz, d);
            if (lane >= d) { sx += tx; sy += ty; sz += tz; }
        }
        wsum[lane][0] = sx; wsum[lane][1] = sy; wsum[lane][2] = sz;   // inclusive over warps
    }
    __syncthreads();
    const int bx = warp ? wsum[warp - 1][0] : 0, by = warp ? wsum[warp - 1][1] : 0, bz = warp ? wsum[warp - 1][2] : 0;
    tot[0] = wsum[31][0]; tot[1] = wsum[31][1]; tot[2] = wsum[31][2];
    x = bx + ix - x; y = by + iy - y; z = bz + iz - z;
    __syncthreads();
}

constexpr int SCAN_PER_THREAD = 8;   // clusters per thread and round: 8,192 clusters per block-wide scan
constexpr int SCAN_BLOCK = 1024 * SCAN_PER_THREAD;

// One block-wide round over clusters [c0, c0 + SCAN_BLOCK): exclusive prefixes of (group size, SIMT items, tcgen05 items)
// written relative to `carry`; returns the round's totals in tot[] and the number of touched clusters of this thread.
__device__ __forceinline__ int scan_round(const ScoreArgs &a, int c0, const int carry[3], int tot[3], int (*wsum)[3]) {
    const int C = a.n_clusters;
    const int cb = c0 + threadIdx.x * SCAN_PER_THREAD;      // this thread's consecutive clusters
    int g[SCAN_PER_THREAD], ns[SCAN_PER_THREAD], nu[SCAN_PER_THREAD];
    int sg = 0, ss = 0, su = 0, touched = 0;
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        g[i] = ns[i] = nu[i] = 0;
        const int c = cb + i;
        if (c < C) {
            g[i] = a.cnt[c];
            const int size = a.offsets[c + 1] - a.offsets[c];
            if (g[i] > 0 && size > 0) {
                touched++;
                item_counts(a, g[i], size, ns[i], nu[i]);
            }
        }
        sg += g[i]; ss += ns[i]; su += nu[i];
    }
    block_scan3(sg, ss, su, tot, wsum);                      // exclusive prefix of the per-thread sums
    sg += carry[0]; ss += carry[1]; su += carry[2];
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) {
        const int c = cb + i;
        if (c < C) {
            a.grp_off[c] = sg;
            a.simt_off[c] = ss;
            a.umma_off[c] = su;
        }
        sg += g[i]; ss += ns[i]; su += nu[i];
    }
    return touched;
}

// block reduce of `touched` (reuses wsum); result valid in thread 0
__device__ __forceinline__ int reduce_touched(int touched, int (*wsum)[3]) {
    for (int d = 16; d; d >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][0] = touched;
    __syncthreads();
    int t = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += wsum[w][0];
    return t;
}

// C <= SCAN_BLOCK: one CTA, one round, final offsets directly
__global__ void __launch_bounds__(1024) k_scan(ScoreArgs a) {
    __shared__ int wsum[32][3];
    pdl_launch_dependents();
    pdl_wait();
    const int carry[3] = {0, 0, 0};
    int tot[3];
    const int touched = reduce_touched(scan_round(a, 0, carry, tot, wsum), wsum);
    if (threadIdx.x == 0) {
        const int C = a.n_clusters;
        a.grp_off[C] = tot[0];
        a.simt_off[C] = tot[1];
        a.umma_off[C] = tot[2];
        a.counters[CTR_N_SIMT] = tot[1];
        a.counters[CTR_N_UMMA] = tot[2];
        a.counters[CTR_N_TOUCHED] = touched;
        a.scan_base[0] = a.scan_base[1] = a.scan_base[2] = 0;
    }
}

// C > SCAN_BLOCK (cfg5: 131,072 clusters per GPU): one CTA per 8,192 clusters writes block-local offsets and its totals
// (scan_base[4 * (blk + 1) ...] as raw totals), then one small CTA turns the totals into exclusive bases; k_fill adds
// the base of a cluster's block.  (The single-CTA scan took ~200 us of the 2.4 ms cfg5 step.)
__global__ void __launch_bounds__(1024) k_scan_part(ScoreArgs a) {
    __shared__ int wsum[32][3];
    pdl_launch_dependents();
    pdl_wait();
    const int carry[3] = {0, 0, 0};
    int tot[3];
    const int touched = reduce_touched(scan_round(a, blockIdx.x * SCAN_BLOCK, carry, tot, wsum), wsum);
    if (threadIdx.x == 0) {
        int32_t *t = a.scan_base + 4 * blockIdx.x;
        t[0] = tot[0]; t[1] = tot[1]; t[2] = tot[2]; t[3] = touched;
    }
}

__global__ void __launch_bounds__(32) k_scan_bases(ScoreArgs a, int n_blocks) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) {
        int run[4] = {0, 0, 0, 0}, last[3] = {0, 0, 0};
        for (int b = 0; b < n_blocks; ++b) {
            int32_t *t = a.scan_base + 4 * b;
            const int v[4] = {t[0], t[1], t[2], t[3]};
            t[0] = run[0]; t[1] = run[1]; t[2] = run[2];
            for (int i = 0; i < 4; ++i) run[i] += v[i];
            last[0] = v[0]; last[1] = v[1]; last[2] = v[2];
        }
        const int C = a.n_clusters;
        a.grp_off[C] = last[0];       // block-LOCAL end offsets of the last block: k_fill takes differences inside a block
        a.simt_off[C] = last[1];
        a.umma_off[C] = last[2];
        a.counters[CTR_N_SIMT] = run[1];
        a.counters[CTR_N_UMMA] = run[2];
        a.counters[CTR_N_TOUCHED] = run[3];
    }
}

__global__ void __launch_bounds__(256) k_fill(ScoreArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_pairs = (int64_t)a.B * a.K;
    // offsets are block-local when the scan ran in several CTAs: add the base of the cluster's 8,192-cluster block
    if (t < n_pairs) {
        const int c = a.beams[t];
        if (c >= a.c_lo && c < a.c_hi && c >= 0 && c < a.n_clusters) {
            const int slot = a.scan_base[4 * (c / SCAN_BLOCK)] + a.grp_off[c] + atomicSub(&a.cnt[c], 1) - 1;
            a.grp_pair[slot] = (int32_t)t;
        }
    }
    if (t < a.n_clusters) {
        const int c = (int)t;
        const int32_t *base = a.scan_base + 4 * (c / SCAN_BLOCK);
        // group size: the next cluster's local offset, unless it starts a new block (then this block's total is needed:
        // recompute from the counter-free item count — cnt is being drained, so use the scan's own prefix differences)
        const bool last_in_block = (c + 1) % SCAN_BLOCK == 0 && c + 1 < a.n_clusters;
        int g;
        if (!last_in_block) g = a.grp_off[c + 1] - a.grp_off[c];
        else g = (a.scan_base[4 * (c / SCAN_BLOCK + 1)] - base[0]) - a.grp_off[c];
        write_items(a, c, base[0] + a.grp_off[c], g, base[1] + a.simt_off[c], base[2] + a.umma_off[c]);
    }
    trace_end(a.dbg, 1);
}

// One warp per tcgen05 tile: work item -> pairs -> candidate offsets, written as the tile's TileMeta record.
__device__ __forceinline__ void write_tile_meta(const ScoreArgs &a, int t, int lane) {
    const Item item = a.umma_items[t];
    const int nq = item.nrows_nq >> 16;
    int qrow = 0, off = 0;
    if (lane < nq) {
        const int p = a.grp_pair[item.slot0 + lane];
        const int b = p / a.K;
        qrow = (a.flags & GDR_Q_PER_BEAM) ? p : b;
        off = (int)pack_score_off(a, b, a.candoff[p + b] + item.rel0);
    }
    TileMeta *m = a.tile_meta + t;
    if (lane == 0) {
        m->row0 = item.row0;
        m->nrows = item.nrows_nq & 0xffff;
        m->nq = nq;
        m->rel0 = item.rel0;
    }
    m->qrow[lane] = qrow;
    m->off[lane] = off;
}

__global__ void __launch_bounds__(256) k_tilemeta(ScoreArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int n_tiles = a.counters[CTR_N_UMMA];
    const int lane = threadIdx.x & 31;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles; t += (gridDim.x * blockDim.x) >> 5) write_tile_meta(a, t, lane);
}

// Small batches (B <= 64, C <= 2048): the whole inversion in ONE CTA with the per-cluster arrays in shared
// memory — one launch instead of three dependent ones.  (Larger batches need the per-query walk spread over many
// SMs: measured 47 us in one CTA vs 14 us as three kernels for B = 1,024.)
constexpr int INV_SMALL_C = 2048;   // 4 arrays x 2048 x 4 B = 32 KB of static shared memory
__global__ void __launch_bounds__(1024) k_invert_small(ScoreArgs a) {
    __shared__ int32_t s_cnt[INV_SMALL_C];
    __shared__ int32_t s_off[INV_SMALL_C + 1];
    __shared__ int32_t s_simt[INV_SMALL_C + 1];
    __shared__ int32_t s_umma[INV_SMALL_C + 1];
    __shared__ int wsum[32][3];
    const int C = a.n_clusters;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    for (int c = tid; c < C; c += 1024) s_cnt[c] = 0;
    __syncthreads();
    pdl_wait();
    for (int b = warp; b < a.B; b += 32) count_query(a, b, lane, s_cnt);
    __syncthreads();
    int carry[3] = {0, 0, 0};
    int touched = 0;
    for (int c0 = 0; c0 < C; c0 += 1024) {
        const int c = c0 + tid;
        int g = 0, ns = 0, nu = 0;
        if (c < C) {
            g = s_cnt[c];
            const int size = a.offsets[c + 1] - a.offsets[c];
            if (g > 0 && size > 0) {
                touched++;
                item_counts(a, g, size, ns, nu);
            }
        }
        int tot[3];
        block_scan3(g, ns, nu, tot, wsum);
        if (c < C) { s_off[c] = carry[0] + g; s_simt[c] = carry[1] + ns; s_umma[c] = carry[2] + nu; }
        carry[0] += tot[0]; carry[1] += tot[1]; carry[2] += tot[2];
    }
    for (int d = 16; d; d >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, d);
    __syncthreads();
    if (lane == 0) wsum[warp][0] = touched;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += wsum[w][0];
        a.counters[CTR_N_SIMT] = carry[1];
        a.counters[CTR_N_UMMA] = carry[2];
        a.counters[CTR_N_TOUCHED] = t;
    }
    const int n_pairs = a.B * a.K;
    for (int p = tid; p < n_pairs; p += 1024) {
        const int c = a.beams[p];
        if (c >= a.c_lo && c < a.c_hi && c >= 0 && c < C) a.grp_pair[s_off[c] + atomicSub(&s_cnt[c], 1) - 1] = p;
    }
    // s_cnt is being decremented above; group sizes for the items come from the scan
    for (int c = tid; c < C; c += 1024) {
        const int g = (c + 1 < C ? s_off[c + 1] : carry[0]) - s_off[c];
        write_items(a, c, s_off[c], g, s_simt[c], s_umma[c]);
    }
    if (a.tile_meta) {
        __syncthreads();                       // items and grp_pair were written by this CTA
        for (int t = warp; t < carry[2]; t += 32) write_tile_meta(a, t, lane);
    }
}

cudaError_t launch_invert(const ScoreArgs &a, cudaStream_t s, int *n_launches) {
    if (a.n_clusters <= INV_SMALL_C && a.B <= 64 && (int64_t)a.B * a.K <= 65536) {   // small batches (the reference's eval_batch_size 1-64)
        *n_launches += 1;
        return launch_pdl(k_invert_small, dim3(1), dim3(1024), 0, s, a.launch_prio, a);
    }
    cudaError_t e = launch_pdl(k_count, dim3((a.B + 3) / 4), dim3(128), 0, s, a.launch_prio, a);
    const int n_blocks = (a.n_clusters + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (n_blocks <= 1) {
        // as few threads as the cluster count needs (8 clusters per thread): a 1,024-thread CTA cannot be placed on an SM
        // that already hosts the scoring kernel and top-k CTAs of neighbouring batches, which stalled the whole pipeline
        const int threads = min(1024, max(32, ((a.n_clusters + SCAN_PER_THREAD - 1) / SCAN_PER_THREAD + 31) / 32 * 32));
        if (e == cudaSuccess) e = launch_pdl(k_scan, dim3(1), dim3(threads), 0, s, a.launch_prio, a);
    } else {
        if (e == cudaSuccess) e = launch_pdl(k_scan_part, dim3(n_blocks), dim3(1024), 0, s, a.launch_prio, a);
        if (e == cudaSuccess) e = launch_pdl(k_scan_bases, dim3(1), dim3(32), 0, s, a.launch_prio, a, n_blocks);
        *n_launches += 1;
    }
    const int64_t n = max((int64_t)a.B * a.K, (int64_t)a.n_clusters);
    if (e == cudaSuccess) e = launch_pdl(k_fill, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, a.launch_prio, a);
    *n_launches += 3;
    if (a.tile_meta) {
        if (e == cudaSuccess) e = launch_pdl(k_tilemeta, dim3(148), dim3(256), 0, s, a.launch_prio, a);
        *n_launches += 1;
    }
    return e;
}

}  // namespace gdr
