// Which stream-capture topologies work with CUDA green contexts on this driver?  (tools/greenctx/README in scripts/README.md)
// Stand-alone: nvcc -arch=sm_100a capture_probe.cu -lcuda -o capture_probe.  Prints one line per API call that fails and a verdict per case.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <vector>

__global__ void probe_kernel(unsigned *mask, int spin_us) {
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicOr(&mask[smid >> 5], 1u << (smid & 31));
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)spin_us * 1900) { }
    }
}

static int count_sms(const unsigned *mask_dev) {
    unsigned h[8];
    cudaMemcpy(h, mask_dev, sizeof(h), cudaMemcpyDeviceToHost);
    int n = 0;
    for (unsigned w : h) n += __builtin_popcount(w);
    return n;
}

#define RT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { printf("    %s -> %s\n", #call, cudaGetErrorName(e_)); ok = false; } } while (0)
#define DR(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { const char *m_ = nullptr; cuGetErrorName(r_, &m_); printf("    %s -> %s\n", #call, m_ ? m_ : "?"); ok = false; } } while (0)

static cudaError_t launch(cudaStream_t s, unsigned *mask, bool ex_pdl) {
    if (!ex_pdl) {
        probe_kernel<<<592, 128, 0, s>>>(mask, 20);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(592); cfg.blockDim = dim3(128); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, probe_kernel, mask, 20);
}

// origin: stream the capture begins on; sides: streams forked from it by events (each launches one kernel on its own mask)
static void run_case(const char *name, cudaStream_t origin, std::vector<cudaStream_t> sides, std::vector<unsigned *> masks, cudaStreamCaptureMode mode,
                     bool ex_pdl, cudaStream_t replay_on, bool origin_kernel, unsigned *origin_mask) {
    bool ok = true;
    printf("case %s (mode %d, %s)\n", name, (int)mode, ex_pdl ? "cudaLaunchKernelEx + PDL attribute" : "<<<>>>");
    for (unsigned *m : masks) cudaMemset(m, 0, 32);
    if (origin_mask) cudaMemset(origin_mask, 0, 32);
    cudaDeviceSynchronize();
    cudaEvent_t fork, join[8];
    RT(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    for (size_t i = 0; i < sides.size(); ++i) RT(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
    cudaGraph_t g = nullptr;
    RT(cudaStreamBeginCapture(origin, mode));
    if (origin_kernel) RT(launch(origin, origin_mask, ex_pdl));
    RT(cudaEventRecord(fork, origin));
    for (size_t i = 0; i < sides.size(); ++i) {
        RT(cudaStreamWaitEvent(sides[i], fork, 0));
        RT(launch(sides[i], masks[i], ex_pdl));
        RT(launch(sides[i], masks[i], ex_pdl));
        RT(cudaEventRecord(join[i], sides[i]));
        RT(cudaStreamWaitEvent(origin, join[i], 0));
    }
    RT(cudaStreamEndCapture(origin, &g));
    (void)cudaGetLastError();
    if (ok && g) {
        cudaGraphExec_t ge = nullptr;
        RT(cudaGraphInstantiate(&ge, g, 0));
        if (ok) {
            RT(cudaGraphLaunch(ge, replay_on));
            RT(cudaStreamSynchronize(replay_on));
        }
        if (ok) {
            printf("    replay ok; SMs used:");
            if (origin_mask) printf(" origin %d", count_sms(origin_mask));
            for (size_t i = 0; i < masks.size(); ++i) printf(" side%zu %d", i, count_sms(masks[i]));
            printf("\n");
        }
        if (ge) cudaGraphExecDestroy(ge);
    }
    if (g) cudaGraphDestroy(g);
    (void)cudaGetLastError();
    cudaDeviceSynchronize();
    (void)cudaGetLastError();
    printf("    => %s\n", ok ? "WORKS" : "FAILS");
}

int main() {
    bool ok = true;
    RT(cudaFree(nullptr));
    CUdevice dev;
    DR(cuDeviceGet(&dev, 0));
    CUdevResource all, grp, rest;
    DR(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned n = 1;
    DR(cuDevSmResourceSplitByCount(&grp, &n, &all, &rest, 0, 64));
    printf("device SMs %u -> group %u + rest %u\n", all.sm.smCount, grp.sm.smCount, rest.sm.smCount);
    CUdevResourceDesc d_small, d_big;
    DR(cuDevResourceGenerateDesc(&d_small, &grp, 1));
    DR(cuDevResourceGenerateDesc(&d_big, &rest, 1));
    CUgreenCtx g_small, g_big;
    DR(cuGreenCtxCreate(&g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    DR(cuGreenCtxCreate(&g_big, d_big, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream a1, a2, b1;
    DR(cuGreenCtxStreamCreate(&a1, g_small, CU_STREAM_NON_BLOCKING, 0));
    DR(cuGreenCtxStreamCreate(&a2, g_small, CU_STREAM_NON_BLOCKING, 0));
    DR(cuGreenCtxStreamCreate(&b1, g_big, CU_STREAM_NON_BLOCKING, 0));
    cudaStream_t p1, p2;
    RT(cudaStreamCreateWithFlags(&p1, cudaStreamNonBlocking));
    RT(cudaStreamCreateWithFlags(&p2, cudaStreamNonBlocking));
    unsigned *m[4];
    for (auto &x : m) RT(cudaMalloc(&x, 32));
    if (!ok) { printf("setup failed\n"); return 1; }

    // eager sanity: the partitions hold
    for (auto &x : m) cudaMemset(x, 0, 32);
    launch(a1, m[0], false); launch(b1, m[1], false); launch(p1, m[2], false);
    cudaDeviceSynchronize();
    printf("eager: small-set stream used %d SMs, big-set stream %d, primary stream %d\n", count_sms(m[0]), count_sms(m[1]), count_sms(m[2]));

    for (int pdl = 0; pdl < 2; ++pdl) {
        const bool x = pdl != 0;
        run_case("T0 primary origin, primary side (control)", p1, {p2}, {m[0]}, cudaStreamCaptureModeGlobal, x, p1, true, m[3]);
        run_case("T1 primary origin -> small-set side", p1, {(cudaStream_t)a1}, {m[0]}, cudaStreamCaptureModeGlobal, x, p1, true, m[3]);
        run_case("T1r same, relaxed mode", p1, {(cudaStream_t)a1}, {m[0]}, cudaStreamCaptureModeRelaxed, x, p1, true, m[3]);
        run_case("T1n primary origin without a kernel of its own -> small-set side", p1, {(cudaStream_t)a1}, {m[0]}, cudaStreamCaptureModeGlobal, x, p1, false, nullptr);
        run_case("T2 small-set origin -> big-set side", (cudaStream_t)a1, {(cudaStream_t)b1}, {m[0]}, cudaStreamCaptureModeGlobal, x, (cudaStream_t)a1, true, m[3]);
        run_case("T2p same, replayed on a primary stream", (cudaStream_t)a1, {(cudaStream_t)b1}, {m[0]}, cudaStreamCaptureModeGlobal, x, p1, true, m[3]);
        run_case("T3 small-set origin -> small-set side (one green context)", (cudaStream_t)a1, {(cudaStream_t)a2}, {m[0]}, cudaStreamCaptureModeGlobal, x, (cudaStream_t)a1, true, m[3]);
        run_case("T3p same, replayed on a primary stream", (cudaStream_t)a1, {(cudaStream_t)a2}, {m[0]}, cudaStreamCaptureModeGlobal, x, p1, true, m[3]);
        run_case("T4 primary origin -> small-set and big-set sides", p1, {(cudaStream_t)a1, (cudaStream_t)b1}, {m[0], m[1]}, cudaStreamCaptureModeGlobal, x, p1, false, nullptr);
        run_case("T5 small-set origin alone (no fork)", (cudaStream_t)a1, {}, {}, cudaStreamCaptureModeGlobal, x, (cudaStream_t)a1, true, m[3]);
        run_case("T5p same, replayed on a primary stream", (cudaStream_t)a1, {}, {}, cudaStreamCaptureModeGlobal, x, p1, true, m[3]);
    }
    return 0;
}
