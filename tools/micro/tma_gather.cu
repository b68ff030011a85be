// microbenchmark: GATHER of random 32-byte table rows, the step's forward access pattern:
//   (a) paired LDG.256: two adjacent lanes fetch rows 2j / 2j+1 (z-neighbours) with one ld.global.nc.v8.f32 each
//   (b) cp.async.bulk (TMA, global -> shared, 32 B per row, mbarrier complete_tx), then LDS of the row
// on a table that lives in L2 (C2-sized, 2.75 MB) and on one that does not (512 MB).  Prints rows/clk/SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tma_gather tools/micro/tma_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t rnd(uint32_t x) { x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16; return x; }
constexpr int kRowsPerIter = 4;      // rows in flight per lane and iteration (the step fetches 4 per level)

__global__ void k_ldg(const float* __restrict__ table, uint32_t row_pairs, int iters, float* out) {
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pair = gt >> 1, half = gt & 1;
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        float v[kRowsPerIter][8];
#pragma unroll
        for (int k = 0; k < kRowsPerIter; ++k) {
            const uint32_t r = 2u * (rnd(pair * 977u + i * kRowsPerIter + k) % row_pairs) + half;
            const float* p = table + (size_t)r * 8;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[k][0]), "=f"(v[k][1]), "=f"(v[k][2]), "=f"(v[k][3]), "=f"(v[k][4]), "=f"(v[k][5]), "=f"(v[k][6]), "=f"(v[k][7]) : "l"(p));
        }
#pragma unroll
        for (int k = 0; k < kRowsPerIter; ++k)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += v[k][q];
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void k_tma(const float* __restrict__ table, uint32_t row_pairs, int iters, float* out) {
    extern __shared__ __align__(128) float sm[];                 // [threads][kRowsPerIter][8] + one mbarrier per warp
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pair = gt >> 1, half = gt & 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* my = sm + threadIdx.x * (kRowsPerIter * 8);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + blockDim.x * kRowsPerIter * 8);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars + warp);
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(my);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    float acc = 0.f;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
        if (lane == 0)
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(32u * kRowsPerIter * 32u) : "memory");
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kRowsPerIter; ++k) {
            const uint32_t r = 2u * (rnd(pair * 977u + i * kRowsPerIter + k) % row_pairs) + half;
            const float* p = table + (size_t)r * 8;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                         ::"r"(dst0 + 32u * k), "l"(p), "r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        phase ^= 1u;
#pragma unroll
        for (int k = 0; k < kRowsPerIter; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(my + 8 * k), b = *reinterpret_cast<const float4*>(my + 8 * k + 4);
            acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
        }
        __syncwarp();
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    float* out; cudaMalloc(&out, 4);
    int dev = 0, clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double ghz = clk * 1e-6;
    const size_t sizes[2] = {(size_t)86016 * 32, (size_t)512 << 20};
    const char* names[2] = {"2.75 MB table (L2 resident)", "512 MB table (HBM)"};
    for (int s = 0; s < 2; ++s) {
        float* t; cudaMalloc(&t, sizes[s]); cudaMemset(t, 0, sizes[s]);
        const uint32_t row_pairs = (uint32_t)(sizes[s] / 64);
        const int blocks = 148 * 4, thr = 256, iters = 128;
        const double nrows = (double)blocks * thr * iters * kRowsPerIter;
        printf("== %s, %d blocks x %d threads, %d rows in flight per lane\n", names[s], blocks, thr, kRowsPerIter);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); k_ldg<<<blocks, thr>>>(t, row_pairs, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("paired LDG.256      : %.3f ms  %.1f Grows/s  %.3f rows/clk/SM  %.0f GB/s  %s\n", ms, nrows / ms / 1e6,
                   nrows / (ms * 1e-3) / 148 / (ghz * 1e9), nrows * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            const size_t smem = (size_t)thr * kRowsPerIter * 32 + 8 * (thr / 32);
            cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaEventRecord(e0); k_tma<<<blocks, thr, smem>>>(t, row_pairs, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("cp.async.bulk 32 B  : %.3f ms  %.1f Grows/s  %.3f rows/clk/SM  %.0f GB/s  %s\n", ms, nrows / ms / 1e6,
                   nrows / (ms * 1e-3) / 148 / (ghz * 1e9), nrows * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(t);
    }
    printf("(the C2 step needs ~%.2f gathered rows/clk/SM to run in 0.20 ms: 776616 points x 32 rows x hit fraction 0.56)\n",
           776616.0 * 32 * 0.56 / (0.20e-3 * ghz * 1e9 * 148));
    return 0;
}
