// SM partition for the pipelined fine stage (include/gdr_b200.h, gdr_partition_*): two disjoint sets of SMs of the current device,
// each behind a CUDA green context with streams of its own.
//
// Why: at cfg2 the HBM-bound scoring kernel (one 174 KB CTA per SM) and the latency-bound top-k / inversion kernels of the
// neighbouring batches cost more together than alone because they compete for RESIDENCY: top-k CTAs fill an SM whenever no scoring
// CTA is pending, and the next scoring CTA then waits for several of them to leave (ROADMAP.md).  Launch
// priorities only reorder pending CTAs; a partition gives each side SMs the other cannot take.  The scoring kernel is HBM-bound from
// about 100 CTAs on, so it does not need all 148 SMs.
//
// The reference has no counterpart (one batch at a time on one stream, GDR_model/main_models.py:1434-1637); this is launch
// plumbing and never changes a result.  Driver entry points are resolved through the runtime (no link against libcuda).
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/gdr_b200.h"
#include "gdr_common.cuh"

struct gdr_partition {
    CUgreenCtx ctx[2] = {nullptr, nullptr};          // 0 = big (scoring), 1 = small (inversion + top-k)
    std::vector<CUstream> streams[2];
    int sms[2] = {0, 0};
};

namespace {

struct Driver {
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
    CUresult (*StreamDestroy)(CUstream) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
    const char *missing = nullptr;
};

template <typename F>
void resolve(Driver &d, F &fn, const char *name) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) {
        if (!d.missing) d.missing = name;
        (void)cudaGetLastError();
        return;
    }
    fn = reinterpret_cast<F>(p);
}

Driver load_driver() {
    Driver d;
    resolve(d, d.DeviceGet, "cuDeviceGet");
    resolve(d, d.DeviceGetDevResource, "cuDeviceGetDevResource");
    resolve(d, d.DevSmResourceSplitByCount, "cuDevSmResourceSplitByCount");
    resolve(d, d.DevResourceGenerateDesc, "cuDevResourceGenerateDesc");
    resolve(d, d.GreenCtxCreate, "cuGreenCtxCreate");
    resolve(d, d.GreenCtxDestroy, "cuGreenCtxDestroy");
    resolve(d, d.GreenCtxStreamCreate, "cuGreenCtxStreamCreate");
    resolve(d, d.StreamDestroy, "cuStreamDestroy");
    resolve(d, d.GetErrorString, "cuGetErrorString");
    return d;
}

int fail(int code, const std::string &msg) {
    gdr::set_error(msg);
    return code;
}

int drv_fail(const Driver &d, CUresult r, const char *what) {
    const char *msg = nullptr;
    if (d.GetErrorString) d.GetErrorString(r, &msg);
    return fail(GDR_ERR_CUDA, std::string("gdr_partition_create: ") + what + " failed: " + (msg ? msg : "?") + " (" + std::to_string((int)r) + ")");
}

void destroy(const Driver &d, gdr_partition *p) {
    for (int side = 0; side < 2; ++side) {
        for (CUstream s : p->streams[side])
            if (s && d.StreamDestroy) d.StreamDestroy(s);
        if (p->ctx[side] && d.GreenCtxDestroy) d.GreenCtxDestroy(p->ctx[side]);
    }
    delete p;
}

}  // namespace

extern "C" {

int gdr_partition_create(gdr_partition_t **out, int32_t small_sms, int32_t n_streams_big, int32_t n_streams_small) {
    if (!out) return fail(GDR_ERR_INVALID, "gdr_partition_create: out is null");
    *out = nullptr;
    if (small_sms < 8 || n_streams_big < 1 || n_streams_small < 1 || n_streams_big > 64 || n_streams_small > 64)
        return fail(GDR_ERR_INVALID, "gdr_partition_create: small_sms must be >= 8 and both stream counts in [1, 64]");
    int dev_ordinal = 0;
    cudaError_t e = cudaFree(nullptr);                       // the primary context exists from here on
    if (e == cudaSuccess) e = cudaGetDevice(&dev_ordinal);
    if (e != cudaSuccess) return gdr::cuda_fail(e, "gdr_partition_create");
    const Driver d = load_driver();
    if (d.missing) return fail(GDR_ERR_UNSUPPORTED, std::string("gdr_partition_create: this driver does not export ") + d.missing + " (green contexts need CUDA 12.4+)");

    CUdevice dev;
    CUresult r = d.DeviceGet(&dev, dev_ordinal);
    if (r != CUDA_SUCCESS) return drv_fail(d, r, "cuDeviceGet");
    CUdevResource all, group, rest;
    r = d.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
    if (r != CUDA_SUCCESS) return drv_fail(d, r, "cuDeviceGetDevResource");
    if ((unsigned)small_sms + 8 > all.sm.smCount) return fail(GDR_ERR_INVALID, "gdr_partition_create: small_sms leaves fewer than 8 of the device's " + std::to_string(all.sm.smCount) + " SMs");
    unsigned int n_groups = 1;
    r = d.DevSmResourceSplitByCount(&group, &n_groups, &all, &rest, 0, (unsigned)small_sms);    // the driver rounds the group up to its granularity (8 SMs on sm_90+)
    if (r != CUDA_SUCCESS || n_groups != 1) return drv_fail(d, r, "cuDevSmResourceSplitByCount");
    if (rest.sm.smCount < 8) return fail(GDR_ERR_INVALID, "gdr_partition_create: the split left " + std::to_string(rest.sm.smCount) + " SMs for the scoring side");

    gdr_partition *p = new gdr_partition();
    CUdevResource *res[2] = {&rest, &group};
    const int n_streams[2] = {n_streams_big, n_streams_small};
    for (int side = 0; side < 2; ++side) {
        CUdevResourceDesc desc;
        r = d.DevResourceGenerateDesc(&desc, res[side], 1);
        if (r != CUDA_SUCCESS) { destroy(d, p); return drv_fail(d, r, "cuDevResourceGenerateDesc"); }
        r = d.GreenCtxCreate(&p->ctx[side], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM);
        if (r != CUDA_SUCCESS) { destroy(d, p); return drv_fail(d, r, "cuGreenCtxCreate"); }
        p->sms[side] = (int)res[side]->sm.smCount;
        for (int i = 0; i < n_streams[side]; ++i) {
            CUstream s = nullptr;
            r = d.GreenCtxStreamCreate(&s, p->ctx[side], CU_STREAM_NON_BLOCKING, 0);
            if (r != CUDA_SUCCESS) { destroy(d, p); return drv_fail(d, r, "cuGreenCtxStreamCreate"); }
            p->streams[side].push_back(s);
        }
    }
    *out = p;
    return GDR_OK;
}

int gdr_partition_sms(const gdr_partition_t *p, int32_t out[2]) {
    if (!p || !out) return fail(GDR_ERR_INVALID, "gdr_partition_sms: null argument");
    out[0] = p->sms[0];
    out[1] = p->sms[1];
    return GDR_OK;
}

void *gdr_partition_stream(gdr_partition_t *p, int32_t small, int32_t index) {
    if (!p || index < 0) return nullptr;
    const std::vector<CUstream> &v = p->streams[small ? 1 : 0];
    return index < (int)v.size() ? (void *)v[index] : nullptr;
}

int gdr_partition_destroy(gdr_partition_t *p) {
    if (!p) return GDR_OK;
    const Driver d = load_driver();
    destroy(d, p);
    return GDR_OK;
}

}  // extern "C"
