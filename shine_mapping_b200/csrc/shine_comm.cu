// shine_comm.cu — the multi-GPU exchange of the training step (the reference is single-GPU).
//
//   * boundary rows: a spatially partitioned map duplicates the corner rows that lie on a face between two blocks
//     (model/feature_octree.py:131-137 shares corners between neighbouring voxels).  shine_boundary_pack copies the
//     gradients of those rows into a compact exchange buffer at globally agreed slots, shine_boundary_unpack writes
//     the reduced values back.
//   * the collective: shine_allreduce_decoder_grads = ncclAllReduce(sum, in place) over NVLink on the caller's
//     stream.  The exchange buffer is laid out [decoder grads (1 377 + pad) | boundary rows], so ONE collective
//     per step carries both.  NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process — the one
//     PyTorch ships — or the system one), so the library has no link-time dependency on it.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "shine_device.cuh"

namespace {

// ---- boundary pack / unpack -------------------------------------------------------------------------------

template <bool PACK>
__global__ void __launch_bounds__(256) boundary_kernel(const __grid_constant__ shine_boundary plan, int feature_dim,
                                                       float* __restrict__ buf) {
    const shine_boundary_level& b = plan.lv[blockIdx.y];
    const int lp = feature_dim >> 2;
    for (int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gt < (int64_t)b.count * lp;
         gt += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(gt / lp), part = (int)(gt % lp);
        float4* row = reinterpret_cast<float4*>(b.table + (int64_t)b.rows[r] * feature_dim) + part;
        float4* slot = reinterpret_cast<float4*>(buf + b.offset + (int64_t)b.slots[r] * feature_dim) + part;
        if (PACK) *slot = *row; else *row = *slot;
    }
}

int launch_boundary(bool pack, const shine_boundary* plan, int32_t num_levels, int32_t feature_dim, float* buf,
                    cudaStream_t st) {
    if (!plan || !buf || num_levels < 1 || num_levels > SHINE_MAX_LEVELS || feature_dim < 4 || (feature_dim & 3))
        return SHINE_ERR_INVALID_ARG;
    int64_t most = 0;
    for (int i = 0; i < num_levels; ++i) {
        const shine_boundary_level& b = plan->lv[i];
        if (b.count < 0 || (b.count > 0 && (!b.table || !b.rows || !b.slots)) || (b.offset & 3)) return SHINE_ERR_INVALID_ARG;
        if (b.count > most) most = b.count;
    }
    if (most == 0) return SHINE_OK;
    DeviceGuard guard(buf);
    int64_t blocks = (most * (feature_dim / 4) + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)num_levels);
    if (pack) boundary_kernel<true><<<grid, 256, 0, st>>>(*plan, feature_dim, buf);
    else boundary_kernel<false><<<grid, 256, 0, st>>>(*plan, feature_dim, buf);
    return (int)cudaGetLastError();
}

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------

typedef struct { char internal[128]; } nccl_unique_id;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(void** comm, int nranks, nccl_unique_id id, int rank);
typedef int (*fn_all_reduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st);
typedef int (*fn_comm_destroy)(void* comm);
typedef const char* (*fn_error_string)(int);
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_error_string error_string = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy already in the process (PyTorch's)
        if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW);
        if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW);
        if (!a.handle) return a;
        a.get_unique_id = (fn_get_unique_id)dlsym(a.handle, "ncclGetUniqueId");
        a.comm_init_rank = (fn_comm_init_rank)dlsym(a.handle, "ncclCommInitRank");
        a.all_reduce = (fn_all_reduce)dlsym(a.handle, "ncclAllReduce");
        a.comm_destroy = (fn_comm_destroy)dlsym(a.handle, "ncclCommDestroy");
        a.error_string = (fn_error_string)dlsym(a.handle, "ncclGetErrorString");
        a.ok = a.get_unique_id && a.comm_init_rank && a.all_reduce && a.comm_destroy;
        return a;
    }();
    return api;
}

constexpr int kErrNcclBase = -1000;     // -1000 - ncclResult_t
thread_local char g_last_error[256] = {0};

int nccl_rc(int r, const char* what) {
    if (r == 0) return SHINE_OK;
    const char* msg = nccl().error_string ? nccl().error_string(r) : "?";
    snprintf(g_last_error, sizeof(g_last_error), "shine_b200: %s failed: NCCL error %d (%s)", what, r, msg);
    return kErrNcclBase - r;
}

}  // namespace

extern "C" {

int shine_boundary_pack(const shine_boundary* plan, int32_t num_levels, int32_t feature_dim, float* buf, void* stream) {
    return launch_boundary(true, plan, num_levels, feature_dim, buf, (cudaStream_t)stream);
}

int shine_boundary_unpack(const shine_boundary* plan, int32_t num_levels, int32_t feature_dim, float* buf, void* stream) {
    return launch_boundary(false, plan, num_levels, feature_dim, buf, (cudaStream_t)stream);
}

int shine_nccl_unique_id(void* out_id128) {
    if (!out_id128) return SHINE_ERR_INVALID_ARG;
    if (!nccl().ok) return SHINE_ERR_UNSUPPORTED;
    nccl_unique_id id;
    const int rc = nccl_rc(nccl().get_unique_id(&id), "ncclGetUniqueId");
    if (rc == SHINE_OK) memcpy(out_id128, &id, sizeof(id));
    return rc;
}

int shine_nccl_comm_create(const void* id128, int32_t nranks, int32_t rank, int32_t device, void** out_comm) {
    if (!id128 || !out_comm || nranks < 1 || rank < 0 || rank >= nranks) return SHINE_ERR_INVALID_ARG;
    if (!nccl().ok) return SHINE_ERR_UNSUPPORTED;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    nccl_unique_id id;
    memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    const int rc = nccl_rc(nccl().comm_init_rank(&comm, nranks, id, rank), "ncclCommInitRank");
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (rc == SHINE_OK) *out_comm = comm;
    return rc;
}

int shine_allreduce_decoder_grads(void* comm, float* buf, int64_t count, void* stream) {
    if (!comm || count < 0 || (count > 0 && !buf)) return SHINE_ERR_INVALID_ARG;
    if (!nccl().ok) return SHINE_ERR_UNSUPPORTED;
    if (count == 0) return SHINE_OK;
    DeviceGuard guard(buf);
    return nccl_rc(nccl().all_reduce(buf, buf, (size_t)count, kNcclFloat32, kNcclSum, comm, (cudaStream_t)stream),
                   "ncclAllReduce");
}

int shine_nccl_comm_destroy(void* comm) {
    if (!comm) return SHINE_OK;
    if (!nccl().ok) return SHINE_ERR_UNSUPPORTED;
    return nccl_rc(nccl().comm_destroy(comm), "ncclCommDestroy");
}

const char* shine_comm_last_error(void) { return g_last_error; }

}  // extern "C"
