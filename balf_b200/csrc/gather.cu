// (e) Multi-GPU: gather of the per-rank keypoint records over NCCL (NVLink 5 / NVSwitch).
//
// The reference is single-device and batch-1 (demo/demo_match.py:29): there is nothing to replace --
// this is the one collective of the sharded path (image batches shard by rank, no exchange inside the
// network, one all-gather of fixed-size records per step).  The entry takes an `ncclComm_t`; NCCL is bound at run time
// (dlopen of the libnccl.so.2 already loaded by the host process, e.g. the one PyTorch ships), so the library has no link-time
// dependency on it and every other entry point works on a box without NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include "common.cuh"
#include "../../include/balf_b200.h"

namespace balf {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};
static NcclApi g_nccl{};
static const NcclApi* nccl_api() {
    if (g_nccl.ok) return &g_nccl;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);            // the copy the process already uses (PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(h, "ncclAllGather"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.GetErrorString;
    return g_nccl.ok ? &g_nccl : nullptr;
}
#define BALF_NCCL_OK(api, expr)                                                                       \
    do {                                                                                              \
        ncclResult_t r_ = (expr);                                                                     \
        if (r_ != ncclSuccess) return set_error(1000 + (int)r_, "%s failed: %s", #expr, (api)->GetErrorString(r_)); \
    } while (0)

// record row of image b: [xy (2K int32) | score bits (K) | count]  (balf_b200/sharding.py pack_records)
__global__ void pack_records_kernel(const int32_t* __restrict__ xy, const float* __restrict__ score, const int32_t* __restrict__ count,
                                    int K, int32_t* __restrict__ out) {
    const int b = blockIdx.y, row = 3 * K + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row; i += gridDim.x * blockDim.x) {
        int32_t v;
        if (i < 2 * K) v = xy[(size_t)b * 2 * K + i];
        else if (i < 3 * K) v = __float_as_int(score[(size_t)b * K + i - 2 * K]);
        else v = count[b];
        out[(size_t)b * row + i] = v;
    }
}
__global__ void unpack_records_kernel(const int32_t* __restrict__ in, int K, int32_t* __restrict__ xy, float* __restrict__ score,
                                      int32_t* __restrict__ count) {
    const int b = blockIdx.y, row = 3 * K + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row; i += gridDim.x * blockDim.x) {
        const int32_t v = in[(size_t)b * row + i];
        if (i < 2 * K) xy[(size_t)b * 2 * K + i] = v;
        else if (i < 3 * K) score[(size_t)b * K + i - 2 * K] = __int_as_float(v);
        else count[b] = v;
    }
}

}  // namespace balf

using namespace balf;

extern "C" int balf_nccl_unique_id(void* id_host_128) {
    BALF_REQUIRE(id_host_128, "null pointer argument");
    const NcclApi* api = nccl_api();
    BALF_REQUIRE(api, "NCCL (libnccl.so.2) is not available in this process");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    BALF_NCCL_OK(api, api->GetUniqueId(static_cast<ncclUniqueId*>(id_host_128)));
    return 0;
}

extern "C" int balf_nccl_comm_create(const void* id_host_128, int world, int rank, void** comm_out) {
    BALF_REQUIRE(id_host_128 && comm_out && world > 0 && rank >= 0 && rank < world, "bad communicator arguments");
    const NcclApi* api = nccl_api();
    BALF_REQUIRE(api, "NCCL (libnccl.so.2) is not available in this process");
    ncclUniqueId id;
    memcpy(&id, id_host_128, sizeof(id));
    ncclComm_t comm = nullptr;
    BALF_NCCL_OK(api, api->CommInitRank(&comm, world, id, rank));      // on the calling thread's current CUDA device
    *comm_out = comm;
    return 0;
}

extern "C" int balf_nccl_comm_destroy(void* comm) {
    if (!comm) return 0;
    const NcclApi* api = nccl_api();
    BALF_REQUIRE(api, "NCCL (libnccl.so.2) is not available in this process");
    BALF_NCCL_OK(api, api->CommDestroy(static_cast<ncclComm_t>(comm)));
    return 0;
}

extern "C" size_t balf_gather_workspace_bytes(int world, int B_local, int K) {
    if (world <= 0 || B_local <= 0 || K <= 0) return 0;
    return align_up((size_t)B_local * (3 * K + 1) * 4, 256) + align_up((size_t)world * B_local * (3 * K + 1) * 4, 256);
}

extern "C" int balf_gather_keypoints(void* comm, int world, const int32_t* xy, const float* score, const int32_t* count, int B_local,
                                     int K, int32_t* xy_all, float* score_all, int32_t* count_all, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    BALF_REQUIRE(comm && xy && score && count && xy_all && score_all && count_all && workspace, "null pointer argument");
    BALF_REQUIRE(world > 0 && B_local > 0 && K > 0, "world, B_local, K must be positive");
    BALF_REQUIRE(workspace_bytes >= balf_gather_workspace_bytes(world, B_local, K), "workspace too small");
    const NcclApi* api = nccl_api();
    BALF_REQUIRE(api, "NCCL (libnccl.so.2) is not available in this process");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t row = (size_t)3 * K + 1;
    int32_t* local = static_cast<int32_t*>(workspace);
    int32_t* all = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + align_up((size_t)B_local * row * 4, 256));
    pack_records_kernel<<<dim3(cdiv((int)row, 256 * 4), B_local), 256, 0, st>>>(xy, score, count, K, local);
    BALF_LAUNCH_OK();
    {
        ProfScope p("gather_allgather", st);
        BALF_NCCL_OK(api, api->AllGather(local, all, (size_t)B_local * row, ncclInt32, static_cast<ncclComm_t>(comm), st));
    }
    unpack_records_kernel<<<dim3(cdiv((int)row, 256 * 4), world * B_local), 256, 0, st>>>(all, K, xy_all, score_all, count_all);
    BALF_COUNT_LAUNCH(2);
    BALF_LAUNCH_OK();
    return 0;
}
