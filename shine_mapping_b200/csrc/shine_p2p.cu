// shine_p2p.cu — the step's multi-GPU exchange as ONE kernel over NVLink peer memory.
//
// What is exchanged every step: the 1 377 decoder gradients and the gradients of the corner rows that a spatial
// partition duplicates on range faces — a few KB to a few hundred KB, i.e. latency-bound.  The NCCL route
// (shine_boundary_pack -> ncclAllReduce -> shine_boundary_unpack) is three launches and ~20 us of collective latency; here
// every rank runs one small kernel:
//     1. pack      own decoder segment + own boundary rows -> own exchange buffer (parity = step & 1)
//     2. publish   last block to finish: release-store (system scope) of the step number into its flag slot in EVERY
//                  peer's buffer, through the peers' IPC-mapped pointers
//     3. wait      until every peer's flag in the local buffer has reached the step number (bounded spin)
//     4. reduce    out[i] = sum over the ranks that hold a row of the corner (fixed order: bitwise identical on every rank)
//                  of their buffers, read straight over NVLink with all of a thread's loads in flight together, written
//                  in place into the decoder gradients / the table-gradient rows
// Double buffering by step parity replaces the trailing barrier: a rank can only overwrite parity p two steps later, and
// it cannot get there before every peer has published the step in between, i.e. has finished reading parity p.
// Buffers are cudaMalloc'ed here and shared with cudaIpc*MemHandle (one process per GPU).
#include <string.h>

#include "shine_device.cuh"

struct shine_p2p {
    int32_t nranks, rank, device;
    int64_t max_floats;
    unsigned char* local;              // [flags 4 KB | control 256 B | data 2 x max_floats x 4]
    unsigned char* peer[16];           // IPC-mapped bases (peer[rank] == local)
    bool opened[16];
    uint32_t step;
};

namespace {

constexpr int kFlagStride = 128;        // one line per publishing rank
constexpr int kCtrlOff = 4096;          // {uint32 arrive; uint32 timeouts; uint32 steps done}
constexpr int kDataOff = 4096 + 256;
constexpr int kMaxRanks = 16;
constexpr long long kSpinLimit = 1ll << 24;     // ~0.3 s of polling: a missing peer becomes an error flag, not a hang

struct P2PParams {
    unsigned char* peer[kMaxRanks];
    int32_t nranks, rank;
    uint32_t step;
    int64_t max_floats, dec_floats;
    float* dec;                         // decoder segment (in place)
    shine_boundary plan;                // table-gradient rows shared with other ranks
    shine_boundary_inverse inv;         // slot -> local row (or -1) per level: lets the reduce run as ONE flat loop
    int64_t total_floats;               // decoder segment + every level's slots
    int32_t num_levels, feature_dim;
};

__device__ __forceinline__ float* data_of(unsigned char* base, uint32_t step, int64_t max_floats) {
    return reinterpret_cast<float*>(base + kDataOff) + (int64_t)(step & 1u) * max_floats;
}
// peers' buffers are read past L1 (they were written by another GPU)
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

constexpr int kP2PThreads = 512;

__global__ void __launch_bounds__(kP2PThreads) p2p_exchange_kernel(const __grid_constant__ P2PParams P) {
    // the step number lives in device memory (control word 2, bumped by the last block of every launch): the launch has no
    // per-call host state, so a step that contains it can be captured once into a CUDA graph and replayed
    const uint32_t step = reinterpret_cast<volatile uint32_t*>(P.peer[P.rank] + kCtrlOff)[2] + 1u;
    unsigned char* mine_base = P.peer[P.rank];
    float* mine = data_of(mine_base, step, P.max_floats);
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gs = (int64_t)gridDim.x * blockDim.x;
    const int lp = P.feature_dim >> 2;

    // 1. pack
    for (int64_t i = gt; i < P.dec_floats / 4; i += gs)
        reinterpret_cast<float4*>(mine)[i] = reinterpret_cast<const float4*>(P.dec)[i];
    for (int l = 0; l < P.num_levels; ++l) {
        const shine_boundary_level& b = P.plan.lv[l];
        for (int64_t i = gt; i < (int64_t)b.count * lp; i += gs) {
            const int r = (int)(i / lp), part = (int)(i % lp);
            reinterpret_cast<float4*>(mine + b.offset + (int64_t)b.slots[r] * P.feature_dim)[part] =
                reinterpret_cast<const float4*>(b.table + (int64_t)b.rows[r] * P.feature_dim)[part];
        }
    }
    // 2. publish: one system-scope fence by the publishing threads after the block(s) have finished packing (the
    //    barrier / the device-scope fence + counter make the other threads' stores cumulative with it)
    volatile uint32_t* ctrl = reinterpret_cast<volatile uint32_t*>(mine_base + kCtrlOff);
    __shared__ int is_last;
    __syncthreads();
    if (gridDim.x == 1) {
        is_last = 1;
        if (threadIdx.x == 0) ctrl[2] = step;                           // every thread of the block has read it (barrier above)
    } else {
        if (threadIdx.x == 0) {
            __threadfence();
            const uint32_t prev = atomicAdd(const_cast<uint32_t*>(ctrl), 1u);
            __threadfence();                                              // acquire side of the hand-shake for the last block
            is_last = prev == gridDim.x - 1;
            if (is_last) { ctrl[0] = 0u; ctrl[2] = step; }              // every block has arrived, i.e. has read the step
        }
        __syncthreads();
    }
    if (is_last && threadIdx.x < P.nranks) {        // st.release.sys orders everything the barrier made visible to this thread
        uint32_t* flag = reinterpret_cast<uint32_t*>(P.peer[threadIdx.x] + (size_t)P.rank * kFlagStride);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
    }
    // while the publication travels: which of this thread's exchange-buffer quads does the rank hold, and which ranks
    // contribute to them (local reads only).  Up to kItems quads per thread; the launch sizes the grid so that this covers
    // the buffer, the tail loop below catches anything beyond.
    constexpr int kItems = 2;
    float* dst[kItems];
    uint32_t from[kItems];
    const int64_t n4 = P.total_floats / 4;
    const uint32_t everyone = P.nranks >= 32 ? 0xFFFFFFFFu : ((1u << P.nranks) - 1u);
    auto locate = [&](int64_t i, float*& d, uint32_t& m) {
        d = nullptr; m = 0u;
        if (i >= n4) return;
        const int64_t fo = 4 * i;                                    // float offset in the exchange buffer
        if (fo < P.dec_floats) { d = P.dec + fo; m = everyone; return; }
        int l = 0;
#pragma unroll
        for (int q = 1; q < SHINE_MAX_LEVELS; ++q)
            if (q < P.num_levels && fo >= P.plan.lv[q].offset) l = q;
        const int64_t rel = fo - P.plan.lv[l].offset;
        const int slot = (int)(rel / P.feature_dim), within = (int)(rel % P.feature_dim);
        const int row = P.inv.row_of_slot[l][slot];
        if (row < 0) return;                                         // a corner this rank does not hold
        d = P.plan.lv[l].table + (int64_t)row * P.feature_dim + within;
        m = P.inv.holders[l] ? (uint32_t)P.inv.holders[l][slot] : everyone;
    };
#pragma unroll
    for (int k = 0; k < kItems; ++k) locate(gt + k * gs, dst[k], from[k]);
    // 3. wait for every rank's publication of this step (acquire loads of the local flags the peers write)
    if (threadIdx.x < P.nranks) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine_base + (size_t)threadIdx.x * kFlagStride);
        long long spin = 0;
        uint32_t seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
            if (++spin > kSpinLimit) { atomicAdd(const_cast<uint32_t*>(ctrl) + 1, 1u); break; }
        } while ((int32_t)(seen - step) < 0);
    }
    __syncthreads();
    // 4. reduce in place, fixed rank order (bitwise identical on every rank).  Only the ranks that hold a row of the corner
    //    are read, and all of a thread's NVLink loads are in flight together: one round trip, not one per rank.
    auto reduce_item = [&](int64_t i, float* d, uint32_t m) {
        const int64_t fo = 4 * i;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int base = 0; base < kMaxRanks; base += 8) {
            if (base >= P.nranks) break;
            float4 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + r < P.nranks && ((m >> (base + r)) & 1u))
                    v[r] = ld_peer_f4(data_of(P.peer[base + r], step, P.max_floats) + fo);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        }
        *reinterpret_cast<float4*>(d) = acc;
    };
    if (kItems == 2 && dst[0] && dst[1] && P.nranks <= 8) {
        // both quads of this thread: sixteen loads in flight
        float4 va[8], vb[8];
        const int64_t fa = 4 * gt, fb = 4 * (gt + gs);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            va[r] = make_float4(0.f, 0.f, 0.f, 0.f); vb[r] = va[r];
            if (r < P.nranks) {
                const float* base = data_of(P.peer[r], step, P.max_floats);
                if ((from[0] >> r) & 1u) va[r] = ld_peer_f4(base + fa);
                if ((from[1] >> r) & 1u) vb[r] = ld_peer_f4(base + fb);
            }
        }
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a.x += va[r].x; a.y += va[r].y; a.z += va[r].z; a.w += va[r].w;
            b.x += vb[r].x; b.y += vb[r].y; b.z += vb[r].z; b.w += vb[r].w;
        }
        *reinterpret_cast<float4*>(dst[0]) = a;
        *reinterpret_cast<float4*>(dst[1]) = b;
    } else {
#pragma unroll
        for (int k = 0; k < kItems; ++k)
            if (dst[k]) reduce_item(gt + k * gs, dst[k], from[k]);
    }
    for (int64_t i = gt + kItems * gs; i < n4; i += gs) {             // buffers beyond kItems quads per thread
        float* d; uint32_t m;
        locate(i, d, m);
        if (d) reduce_item(i, d, m);
    }
}

// the slots a rank does not hold must read as zero on it: cleared once per parity and kept clean by construction (a rank
// only ever writes the slots of its own rows, always the same ones)

}  // namespace

extern "C" {

int shine_p2p_create(int32_t nranks, int32_t rank, int32_t device, int64_t max_floats, void* out_handle64, shine_p2p** out) {
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks || max_floats < 4 || !out_handle64 || !out)
        return SHINE_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    shine_p2p* ctx = new shine_p2p();
    ctx->nranks = nranks; ctx->rank = rank; ctx->device = device; ctx->step = 0;
    ctx->max_floats = (max_floats + 3) & ~(int64_t)3;
    for (int i = 0; i < kMaxRanks; ++i) { ctx->peer[i] = nullptr; ctx->opened[i] = false; }
    const size_t bytes = (size_t)kDataOff + 2 * (size_t)ctx->max_floats * sizeof(float);
    e = cudaMalloc(reinterpret_cast<void**>(&ctx->local), bytes);
    if (e == cudaSuccess) e = cudaMemset(ctx->local, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ctx->local);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) { if (ctx->local) cudaFree(ctx->local); delete ctx; return (int)e; }
    memcpy(out_handle64, &h, sizeof(h));
    ctx->peer[rank] = ctx->local;
    *out = ctx;
    return SHINE_OK;
}

int shine_p2p_connect(shine_p2p* ctx, const void* handles) {
    if (!ctx || !handles) return SHINE_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->local);
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const unsigned char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return (int)e;
        ctx->peer[r] = reinterpret_cast<unsigned char*>(p);
        ctx->opened[r] = true;
    }
    return SHINE_OK;
}

int shine_p2p_exchange(shine_p2p* ctx, float* dec_grads, int64_t dec_floats, const shine_boundary* plan,
                       const shine_boundary_inverse* inverse, int32_t num_levels, int32_t feature_dim, void* stream) {
    if (!ctx || !dec_grads || dec_floats < 0 || (dec_floats & 3) || num_levels < 0 || num_levels > SHINE_MAX_LEVELS)
        return SHINE_ERR_INVALID_ARG;
    if (num_levels > 0 && (!plan || !inverse || feature_dim < 4 || (feature_dim & 3))) return SHINE_ERR_INVALID_ARG;
    P2PParams P;
    int64_t end = dec_floats;
    for (int l = 0; l < num_levels; ++l) {
        const shine_boundary_level& b = plan->lv[l];
        if (b.count < 0 || (b.count > 0 && (!b.table || !b.rows || !b.slots)) || (b.offset & 3) || b.offset != end)
            return SHINE_ERR_INVALID_ARG;                              // levels are laid out back to back
        if (inverse->slots[l] < 0 || (inverse->slots[l] > 0 && !inverse->row_of_slot[l])) return SHINE_ERR_INVALID_ARG;
        end += (int64_t)inverse->slots[l] * feature_dim;
        P.plan.lv[l] = b;
        P.inv.row_of_slot[l] = inverse->row_of_slot[l]; P.inv.slots[l] = inverse->slots[l];
        P.inv.holders[l] = inverse->holders[l];
    }
    if (end > ctx->max_floats) return SHINE_ERR_INVALID_ARG;
    P.total_floats = end;
    const int64_t most = end / 4;
    for (int r = 0; r < kMaxRanks; ++r) P.peer[r] = r < ctx->nranks ? ctx->peer[r] : nullptr;
    for (int r = 0; r < ctx->nranks; ++r) if (!P.peer[r]) return SHINE_ERR_INVALID_ARG;      // connect() first
    ctx->step += 1;                     // bookkeeping only (launches issued); the protocol's step number is on the device
    P.nranks = ctx->nranks; P.rank = ctx->rank; P.step = 0; P.max_floats = ctx->max_floats; P.dec_floats = dec_floats;
    P.dec = dec_grads; P.num_levels = num_levels; P.feature_dim = feature_dim;
    DeviceGuard guard(ctx->local);
    // latency-bound: every thread should own at most two quads of the buffer, so that the reduce is ONE NVLink round trip
    // (the step kernel has retired: all SMs are free).  One block (no grid hand-shake) covers 1 024 quads = 16 KB.
    int64_t blocks = (most + 2 * kP2PThreads - 1) / (2 * kP2PThreads);
    if (blocks < 1) blocks = 1;
    if (blocks > 128) blocks = 128;
    p2p_exchange_kernel<<<(unsigned)blocks, kP2PThreads, 0, (cudaStream_t)stream>>>(P);
    return (int)cudaGetLastError();
}

int shine_p2p_timeouts(shine_p2p* ctx, int32_t* out_count) {
    if (!ctx || !out_count) return SHINE_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->local);
    uint32_t v = 0;
    cudaError_t e = cudaMemcpy(&v, ctx->local + kCtrlOff + 4, sizeof(v), cudaMemcpyDeviceToHost);
    *out_count = (int32_t)v;
    return (int)e;
}

int shine_p2p_destroy(shine_p2p* ctx) {
    if (!ctx) return SHINE_OK;
    DeviceGuard guard(ctx->local);
    cudaDeviceSynchronize();
    for (int r = 0; r < ctx->nranks; ++r)
        if (ctx->opened[r] && ctx->peer[r]) cudaIpcCloseMemHandle(ctx->peer[r]);
    cudaFree(ctx->local);
    delete ctx;
    return SHINE_OK;
}

}  // extern "C"
