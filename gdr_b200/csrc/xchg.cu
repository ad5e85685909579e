// Peer-to-peer all-gather of small per-rank inputs (the queries and beams of a batch) over NVLink, without a collective kernel.
//
// Why not NCCL here: in the sharded fine stage every rank needs the whole global batch (its beams may land in any rank's
// clusters), and that all-gather sits in the data path of EVERY step.  A NCCL all-gather is a kernel of several 640-thread CTAs
// that spin until the peers arrive; beside the persistent scoring CTAs (one per SM, 170 KB of shared memory) it can only run on
// SMs the scoring grid has not yet taken, and while it spins it keeps the scoring grid off those SMs — measured, the end-to-end
// step of the sharded pipeline was the SUM of its stages (upload 62 us + all-gather 25-60 us + compute) instead of their maximum.
// Here the transfer is done by the COPY ENGINES (cudaMemcpyAsync into the peers' buffers, mapped with CUDA IPC) and the
// synchronisation by one warp: after its copies a rank raises an arrival flag on every peer, and waits for the peers' flags.
//
// One exchange object per rank; the memory comes from the caller (a torch tensor): [n_slots][slot region] | flags | epochs.  A slot
// region holds, part after part, [n_ranks x part_bytes[p]] — so part p of a slot is the rank-order concatenation the consumer wants
// (q [n_ranks * B, D], beams [n_ranks * B, K]).  Slots are reused round-robin by the caller; a slot may be refilled once the batch that
// used it has been fully processed on every rank (gdr_b200/sharded.py: an upload waits for the download of the batch that last
// used the slot, which in turn waited for every rank's scores of that batch).
#include <cstring>
#include <new>
#include <vector>

#include "gdr_common.cuh"

struct gdr_xchg {
    int n_ranks = 1, my_rank = 0, n_slots = 1, n_parts = 0;
    int64_t part_bytes[4] = {0, 0, 0, 0}, part_off[4] = {0, 0, 0, 0};
    int64_t slot_stride = 0, flags_off = 0, epoch_off = 0, total = 0;
    char *base = nullptr;                          // this rank's buffer (owned by the caller)
    char *peer[gdr::GDR_MAX_RANKS] = {nullptr};    // every rank's buffer as mapped here (own entry = base)
    void *opened[gdr::GDR_MAX_RANKS] = {nullptr};  // cudaIpcOpenMemHandle results to close
};

namespace gdr {

struct XchgArgs {
    int32_t *peer_flags[GDR_MAX_RANKS];   // each rank's flag array [n_slots][n_ranks]
    int32_t *local_flags;
    int32_t *epoch;                       // [n_slots] device counters
    int32_t n_ranks, my_rank, slot;
};

__global__ void __launch_bounds__(32) k_xchg_signal_wait(XchgArgs x) {
    if (threadIdx.x == 0) {
        const int e = x.epoch[x.slot] + 1;
        x.epoch[x.slot] = e;
        __threadfence_system();
        for (int r = 0; r < x.n_ranks; ++r)
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(x.peer_flags[r] + x.slot * x.n_ranks + x.my_rank), "r"(e) : "memory");
        for (int r = 0; r < x.n_ranks; ++r) {
            int v;
            do {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(x.local_flags + x.slot * x.n_ranks + r) : "memory");
            } while (v - e < 0);
        }
        __threadfence_system();
    }
}

}  // namespace gdr

using namespace gdr;

static int xinvalid(const char *msg) {
    set_error(msg);
    return GDR_ERR_INVALID;
}

static void xchg_layout(gdr_xchg *x) {
    int64_t off = 0;
    for (int p = 0; p < x->n_parts; ++p) {
        x->part_off[p] = off;
        off += (x->n_ranks * x->part_bytes[p] + 255) / 256 * 256;
    }
    x->slot_stride = off;
    x->flags_off = x->slot_stride * x->n_slots;
    x->epoch_off = x->flags_off + ((int64_t)x->n_slots * x->n_ranks * 4 + 255) / 256 * 256;
    x->total = x->epoch_off + ((int64_t)x->n_slots * 4 + 255) / 256 * 256;
}

extern "C" {

int64_t gdr_xchg_bytes(int32_t n_ranks, int32_t n_slots, const int64_t *part_bytes, int32_t n_parts) {
    if (n_ranks < 1 || n_ranks > GDR_MAX_RANKS || n_slots < 1 || n_parts < 1 || n_parts > 4 || !part_bytes) return -1;
    gdr_xchg x;
    x.n_ranks = n_ranks; x.n_slots = n_slots; x.n_parts = n_parts;
    for (int p = 0; p < n_parts; ++p) {
        if (part_bytes[p] <= 0 || part_bytes[p] % 16) return -1;
        x.part_bytes[p] = part_bytes[p];
    }
    xchg_layout(&x);
    return x.total;
}

int gdr_xchg_create(gdr_xchg_t **out, void *buffer, int32_t n_ranks, int32_t my_rank, int32_t n_slots, const int64_t *part_bytes,
                    int32_t n_parts, void *blob_out) {
    if (!out) return xinvalid("gdr_xchg_create: out is null");
    *out = nullptr;
    if (gdr_xchg_bytes(n_ranks, n_slots, part_bytes, n_parts) < 0) return xinvalid("gdr_xchg_create: need 1 <= n_ranks <= 8, n_slots >= 1, 1-4 parts of positive multiples of 16 bytes");
    if (!buffer || (reinterpret_cast<uintptr_t>(buffer) & 255)) return xinvalid("gdr_xchg_create: buffer must be a 256-byte aligned device pointer");
    if (my_rank < 0 || my_rank >= n_ranks) return xinvalid("gdr_xchg_create: bad rank");
    gdr_xchg *x = new (std::nothrow) gdr_xchg();
    if (!x) return GDR_ERR_NOMEM;
    x->n_ranks = n_ranks; x->my_rank = my_rank; x->n_slots = n_slots; x->n_parts = n_parts;
    for (int p = 0; p < n_parts; ++p) x->part_bytes[p] = part_bytes[p];
    xchg_layout(x);
    x->base = reinterpret_cast<char *>(buffer);
    x->peer[my_rank] = x->base;
    cudaError_t e = cudaMemset(x->base + x->flags_off, 0, (size_t)(x->total - x->flags_off));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && blob_out) {
        // the buffer may sit inside a larger allocation (torch's caching allocator): the IPC handle names the allocation, the blob adds the offset
        CUdeviceptr abase = 0;
        size_t asize = 0;
        typedef CUresult (*PFN_range)(CUdeviceptr *, size_t *, CUdeviceptr);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
            reinterpret_cast<PFN_range>(fn)(&abase, &asize, (CUdeviceptr)(uintptr_t)buffer) != CUDA_SUCCESS) {
            delete x;
            set_error("gdr_xchg_create: cuMemGetAddressRange failed");
            return GDR_ERR_CUDA;
        }
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, reinterpret_cast<void *>((uintptr_t)abase));
        if (e == cudaSuccess) {
            memcpy(blob_out, &h, sizeof(h));
            const int64_t off = (int64_t)((uintptr_t)buffer - (uintptr_t)abase);
            memcpy(reinterpret_cast<char *>(blob_out) + sizeof(h), &off, 8);
        }
    }
    if (e != cudaSuccess) {
        delete x;
        return cuda_fail(e, "gdr_xchg_create");
    }
    *out = x;
    return GDR_OK;
}

int gdr_xchg_attach(gdr_xchg_t *x, const void *all_blobs) {
    if (!x || !all_blobs) return xinvalid("gdr_xchg_attach: null argument");
    for (int r = 0; r < x->n_ranks; ++r) {
        if (r == x->my_rank || x->peer[r]) continue;
        const char *blob = reinterpret_cast<const char *>(all_blobs) + (size_t)r * GDR_XCHG_BLOB_BYTES;
        cudaIpcMemHandle_t h;
        int64_t off = 0;
        memcpy(&h, blob, sizeof(h));
        memcpy(&off, blob + sizeof(h), 8);
        void *p = nullptr;
        GDR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->opened[r] = p;
        x->peer[r] = reinterpret_cast<char *>(p) + off;
    }
    return GDR_OK;
}

int gdr_xchg_attach_local(gdr_xchg_t *x, gdr_xchg_t *const *peers) {
    if (!x || !peers) return xinvalid("gdr_xchg_attach_local: null argument");
    for (int r = 0; r < x->n_ranks; ++r) {
        if (r == x->my_rank) continue;
        if (!peers[r] || peers[r]->total != x->total || peers[r]->my_rank != r) return xinvalid("gdr_xchg_attach_local: peers must be created with the same layout and their own rank");
        x->peer[r] = peers[r]->base;
    }
    return GDR_OK;
}

int gdr_xchg_all_gather(gdr_xchg_t *x, int32_t slot, const void *own, void *stream) {
    if (!x || !own) return xinvalid("gdr_xchg_all_gather: null argument");
    if (slot < 0 || slot >= x->n_slots) return xinvalid("gdr_xchg_all_gather: bad slot");
    cudaStream_t st = (cudaStream_t)stream;
    XchgArgs a;
    memset(&a, 0, sizeof(a));
    int64_t src_off = 0;
    for (int p = 0; p < x->n_parts; ++p) {
        for (int r = 0; r < x->n_ranks; ++r) {
            if (!x->peer[r]) return xinvalid("gdr_xchg_all_gather: peers are not attached");
            char *dst = x->peer[r] + (int64_t)slot * x->slot_stride + x->part_off[p] + (int64_t)x->my_rank * x->part_bytes[p];
            GDR_CUDA(cudaMemcpyAsync(dst, reinterpret_cast<const char *>(own) + src_off, (size_t)x->part_bytes[p], cudaMemcpyDeviceToDevice, st));
        }
        src_off += x->part_bytes[p];
    }
    for (int r = 0; r < x->n_ranks; ++r) a.peer_flags[r] = reinterpret_cast<int32_t *>(x->peer[r] + x->flags_off);
    a.local_flags = reinterpret_cast<int32_t *>(x->base + x->flags_off);
    a.epoch = reinterpret_cast<int32_t *>(x->base + x->epoch_off);
    a.n_ranks = x->n_ranks; a.my_rank = x->my_rank; a.slot = slot;
    k_xchg_signal_wait<<<1, 32, 0, st>>>(a);
    GDR_CUDA(cudaGetLastError());
    return GDR_OK;
}

int64_t gdr_xchg_part_offset(gdr_xchg_t *x, int32_t slot, int32_t part) {
    if (!x || slot < 0 || slot >= x->n_slots || part < 0 || part >= x->n_parts) return -1;
    return (int64_t)slot * x->slot_stride + x->part_off[part];
}

int gdr_xchg_destroy(gdr_xchg_t *x) {
    if (!x) return GDR_OK;
    for (int r = 0; r < GDR_MAX_RANKS; ++r)
        if (x->opened[r]) cudaIpcCloseMemHandle(x->opened[r]);
    delete x;
    return GDR_OK;
}

}  // extern "C"
