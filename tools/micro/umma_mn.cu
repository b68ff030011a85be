// descriptor-semantics probe for tcgen05.mma kind::tf32 with MN-major operands (no swizzle).
// A is [M=128][K=8], B is [N=48][K=8]; element (mn, k) is stored at  (mn>>2)*CH + k*16 + (mn&3)*4  (MN chunks of 4 columns,
// 8 K rows of 16 bytes per core matrix).  Tries the two assignments of the chunk stride to the descriptor's LBO / SBO fields.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t umma_desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void probe(float* out, int variant) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int tid = threadIdx.x;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);
    float* A = reinterpret_cast<float*>(sm);            // 32 chunks x 128 B = 4096 B
    float* B = reinterpret_cast<float*>(sm + 4096);     // 12 chunks x 128 B = 1536 B
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 8192);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8208);
    for (int i = tid; i < 128 * 8; i += 128) { const int m = i / 8, k = i % 8; A[(m >> 2) * 32 + k * 4 + (m & 3)] = (float)((m % 7) + 1) * (float)(k + 1); }
    for (int i = tid; i < 48 * 8; i += 128) { const int n = i / 8, k = i % 8; B[(n >> 2) * 32 + k * 4 + (n & 3)] = (float)((n % 5) + 1) * (k % 2 ? 1.f : 2.f); }
    const uint32_t barx = sb + 8208;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barx)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(sb + 8192)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((48u >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t da, db;
        if (variant == 0) { da = umma_desc(sb, 4096, 128); db = umma_desc(sb + 4096, 1536, 128); }        // LBO = K-group stride, SBO = chunk stride
        else if (variant == 1) { da = umma_desc(sb, 128, 4096); db = umma_desc(sb + 4096, 128, 1536); }   // LBO = chunk stride
        else { da = umma_desc(sb, 128, 128); db = umma_desc(sb + 4096, 128, 128); }
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barx) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; !done && spin < (1 << 22); ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(barx), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float v[48];
    const uint32_t trow = tmem + ((uint32_t)(32 * (tid >> 5)) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=f"(v[0]),"=f"(v[1]),"=f"(v[2]),"=f"(v[3]),"=f"(v[4]),"=f"(v[5]),"=f"(v[6]),"=f"(v[7]),"=f"(v[8]),"=f"(v[9]),"=f"(v[10]),"=f"(v[11]),"=f"(v[12]),"=f"(v[13]),"=f"(v[14]),"=f"(v[15]),"=f"(v[16]),"=f"(v[17]),"=f"(v[18]),"=f"(v[19]),"=f"(v[20]),"=f"(v[21]),"=f"(v[22]),"=f"(v[23]),"=f"(v[24]),"=f"(v[25]),"=f"(v[26]),"=f"(v[27]),"=f"(v[28]),"=f"(v[29]),"=f"(v[30]),"=f"(v[31]) : "r"(trow));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[32]),"=f"(v[33]),"=f"(v[34]),"=f"(v[35]),"=f"(v[36]),"=f"(v[37]),"=f"(v[38]),"=f"(v[39]),"=f"(v[40]),"=f"(v[41]),"=f"(v[42]),"=f"(v[43]),"=f"(v[44]),"=f"(v[45]),"=f"(v[46]),"=f"(v[47]) : "r"(trow + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 48; ++n) out[tid * 48 + n] = v[n];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}
int main() {
    float* d; cudaMalloc(&d, 128 * 48 * 4);
    static float h[128 * 48];
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 9216);
    for (int variant = 0; variant < 3; ++variant) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 128, 9216>>>(d, variant);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0; double maxv = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 48; ++n) {
            double want = 0; for (int k = 0; k < 8; ++k) want += (double)((m % 7) + 1) * (k + 1) * ((n % 5) + 1) * (k % 2 ? 1.0 : 2.0);
            if (fabs(h[m * 48 + n] - want) > 1e-3 * fabs(want)) ++bad; if (fabs(h[m * 48 + n]) > maxv) maxv = fabs(h[m * 48 + n]);
        }
        printf("variant %d: %s, mismatches %d / %d, max|D| %.1f, D[0][0..3] = %.1f %.1f %.1f %.1f, D[5][7] = %.1f (want %.1f)\n", variant,
               cudaGetErrorString(e), bad, 128 * 48, maxv, h[0], h[1], h[2], h[3], h[5 * 48 + 7], 6.0 * 3 * (2 + 2 + 6 + 4 + 10 + 6 + 14 + 8));
    }
    return 0;
}
