// sm_partition.cu -- mx_sm_partition: two disjoint sets of SMs on one GPU (CUDA green contexts), one stream each.
//
// memex's worker embeds and appends (reference lib/worker/src/tasks.rs:15-59) while its API searches
// (lib/api/src/endpoints/collections/handlers.rs:61-81).  Here both roles are persistent kernels that fill every SM when
// alone -- the HBM-bound scan and the tensor-bound forward pass -- so on one GPU they could only take turns.  A partition
// gives the scan `sms_first` SMs (rounded up to the hardware's granularity of 8; the scan stays HBM-bound far below the full
// chip) and the embedder the rest; EVERY kernel launched into a partition's stream, persistent or not, runs on its SMs only.
// The grids still have to be sized for their share: mx_store_set_sm_limit / mx_embedder_set_sm_limit.
//
// Driver entry points are looked up at run time (cudaGetDriverEntryPoint), so the library keeps no link-time
// dependency on libcuda and answers MX_ERR_UNSUPPORTED where the driver has no green contexts.
#include <cuda.h>

#include "common.cuh"

using namespace mx;

struct mx_sm_partition : HandleBase {
    int32_t device = 0;
    CUgreenCtx ctx[2] = {nullptr, nullptr};
    CUstream stream[2] = {nullptr, nullptr};
    uint32_t sms[2] = {0, 0};
};

namespace {
constexpr uint32_t kPartitionMagic = 0x4d585350;   // "MXSP"

template <typename Fn>
Fn entry(const char *name)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return reinterpret_cast<Fn>(p);
}
}  // namespace

extern "C" {

void mx_sm_partition_destroy(mx_sm_partition *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    auto stream_destroy = entry<CUresult (*)(CUstream)>("cuStreamDestroy");
    auto ctx_destroy = entry<CUresult (*)(CUgreenCtx)>("cuGreenCtxDestroy");
    for (int i = 0; i < 2; ++i) {
        if (p->stream[i]) {
            cudaStreamSynchronize((cudaStream_t)p->stream[i]);
            if (stream_destroy) stream_destroy(p->stream[i]);
        }
    }
    for (int i = 0; i < 2; ++i)
        if (p->ctx[i] && ctx_destroy) ctx_destroy(p->ctx[i]);
    p->magic = 0;
    delete p;
}

int32_t mx_sm_partition_create(int32_t device, uint32_t sms_first, mx_sm_partition **out)
{
    if (!out) return fail(nullptr, MX_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MX_ERR_CONNECTION, "no CUDA device (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, MX_ERR_CONNECTION, "device %d out of range [0, %d)", device, ndev);
    MX_CUDA(nullptr, MX_ERR_CONNECTION, cudaSetDevice(device));
    MX_CUDA(nullptr, MX_ERR_CONNECTION, cudaFree(nullptr));   // the primary context exists
    auto dev_get = entry<CUresult (*)(CUdevice *, int)>("cuDeviceGet");
    auto get_res = entry<CUresult (*)(CUdevice, CUdevResource *, CUdevResourceType)>("cuDeviceGetDevResource");
    auto split = entry<CUresult (*)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int,
                                    unsigned int)>("cuDevSmResourceSplitByCount");
    auto gen_desc = entry<CUresult (*)(CUdevResourceDesc *, CUdevResource *, unsigned int)>("cuDevResourceGenerateDesc");
    auto ctx_create = entry<CUresult (*)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int)>("cuGreenCtxCreate");
    auto stream_create = entry<CUresult (*)(CUstream *, CUgreenCtx, unsigned int, int)>("cuGreenCtxStreamCreate");
    if (!dev_get || !get_res || !split || !gen_desc || !ctx_create || !stream_create)
        return fail(nullptr, MX_ERR_UNSUPPORTED, "this driver has no green contexts (SM partitioning)");
    CUdevice dev;
    CUdevResource all{}, first{}, rest{};
    if (dev_get(&dev, device) != CUDA_SUCCESS || get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS)
        return fail(nullptr, MX_ERR_UNSUPPORTED, "cuDeviceGetDevResource failed");
    const uint32_t total = all.sm.smCount;
    if (sms_first == 0 || sms_first + 8 > total)
        return fail(nullptr, MX_ERR_INVALID, "the first partition must have 1 .. %u SMs (both parts need at least 8)", total - 8);
    unsigned int groups = 1;
    CUresult r = split(&first, &groups, &all, &rest, 0, sms_first);
    if (r != CUDA_SUCCESS || groups != 1 || rest.sm.smCount == 0)
        return fail(nullptr, MX_ERR_UNSUPPORTED, "cuDevSmResourceSplitByCount(%u of %u) failed (%d)", sms_first, total, (int)r);
    mx_sm_partition *p = new mx_sm_partition();
    p->magic = kPartitionMagic;
    p->device = device;
    CUdevResource parts[2] = {first, rest};
    for (int i = 0; i < 2; ++i) {
        CUdevResourceDesc desc;
        r = gen_desc(&desc, &parts[i], 1);
        if (r == CUDA_SUCCESS) r = ctx_create(&p->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM);
        if (r == CUDA_SUCCESS) r = stream_create(&p->stream[i], p->ctx[i], CU_STREAM_NON_BLOCKING, 0);
        if (r != CUDA_SUCCESS) {
            int32_t rc = fail(nullptr, MX_ERR_UNSUPPORTED, "green context %d (%u SMs) could not be created (%d)", i,
                              parts[i].sm.smCount, (int)r);
            mx_sm_partition_destroy(p);
            return rc;
        }
        p->sms[i] = parts[i].sm.smCount;
    }
    *out = p;
    return MX_OK;
}

void *mx_sm_partition_stream(mx_sm_partition *p, uint32_t which) { return p && which < 2 ? (void *)p->stream[which] : nullptr; }
uint32_t mx_sm_partition_sms(mx_sm_partition *p, uint32_t which) { return p && which < 2 ? p->sms[which] : 0; }

}  // extern "C"
