// shine_train_tc.cu — the training step with the decoder on the 5th-generation tensor cores (tcgen05.mma, accumulators
// in Tensor Memory), warp-specialised.  Same contract as sdf_fused_kernel<TRAIN> (reference shine_batch.py:123-209);
// selected with SHINE_FLAG_TCGEN05 on shine_sdf_bce_step.
//
// Why: in the mma.sync kernel every warp re-reads the decoder weights and stages the weight-gradient operands through
// the LSU for each 16-point tile, and its time tracks its instruction count (ncu: 115 M warp instructions, 0.39 IPC per
// scheduler whatever the variant).  Here the 2 lanes-per-point gather / scatter warps do nothing but gather and scatter,
// and the decoder runs as 128-point tcgen05 tiles whose operands the tensor core reads from shared memory itself.
//
// One CTA per SM, 20 warps:
//   warps 0-15  gather/scatter (GS), two groups of 8.  Group g owns the rounds r = g, g+2, ... of this CTA; a round is a
//               tile of 128 points, warp w of the group owns rows 16w..16w+15 (lane layout of the mma.sync kernel:
//               2 lanes per point, z-split corners, LDG.256 rows).  Per round: hash walk + gather + blend -> X rows
//               (hi/lo tf32) into the round's operand buffer, corner rows + blend factors parked in TMEM, arrive on
//               x_full; later wait dx_full, read dL/dfeature rows, scatter-add (red.v4) into the tables.
//   warps 16-19 decoder epilogue (EP), one thread per row of the tile; thread 0 also issues the MMAs.  Per round:
//               D1 = X W1^T            -> +b1, ReLU (mask kept), H1 hi/lo -> smem
//               D2 = H1 W2^T           -> +b2, ReLU, pred, BCE loss, dL/dpred, dH2 hi/lo -> smem, dW3/db3 in registers
//               D3 = dH2 W2            -> ReLU mask, dH1 hi/lo -> smem
//               D4 = dH1 W1  and  Dw += [dH2|dH1]^T [H1|X|1]   (weight + bias gradients of layers 1, 2: ONE accumulator
//                                        in TMEM for the whole kernel, operands read MN-major from the same buffers)
//               D4 -> dX rows -> smem, arrive dx_full.
//   All contractions are 3xTF32 (hi*hi + lo*hi + hi*lo, fp32 accumulate) like the mma.sync kernel.
//
// Shared-memory operand layout (no swizzle): an activation matrix [128 points][C columns] is stored as core matrices of
// 8 points x 16 bytes (4 columns): offset(pt, c) = (pt>>3)*S_pt + (c>>2)*128 + (pt&7)*16 + (c&3)*4.  As A operand of the
// forward / dgrad MMAs (M = points, K = columns) this is the canonical K-major layout (LBO = 128, SBO = S_pt); as operand
// of the weight-gradient MMA (M or N = columns, K = points) the very same bytes are the canonical MN-major layout
// (SBO = 128, one 8-point group per K = 8 instruction, start address advanced by S_pt).
#include "shine_device.cuh"

using shine_internal::StepParams;

namespace {

constexpr int kGSWarps = 16, kEPWarps = 4;
constexpr int kTcThreads = 32 * (kGSWarps + kEPWarps);     // 640
constexpr int kRound = 128;                                 // points per round (MMA M)

struct TP {                                               // byte offsets in dynamic shared memory
    static constexpr int HX_PT = 12 * 128;                // [H1 32 | X 8 | 1,0,0,0 | 0 x4] = 48 columns = 12 chunks
    static constexpr int HX_BYTES = 16 * HX_PT;           // 24 576
    static constexpr int HX = 0;                          // [slot 2][hi/lo 2]
    static constexpr int DH_PT = 16 * 128;                // [dH2 32 | dH1 32] = 64 columns = 16 chunks
    static constexpr int DH_BYTES = 16 * DH_PT;           // 32 768
    static constexpr int DH = HX + 4 * HX_BYTES;          // [hi/lo 2]   (the M = 128 weight-gradient A operand reads
                                                          //  2 KB past each half: what follows must stay mapped)
    static constexpr int W1H = DH + 2 * DH_BYTES;         // W1  [N 32][K 8]   K-major: 4 groups x 2 chunks x 128 B
    static constexpr int W1L = W1H + 1024;
    static constexpr int W2H = W1L + 1024;                // W2  [N 32][K 32]  (forward layer 2: B[n2][k1])
    static constexpr int W2L = W2H + 4096;
    static constexpr int W2TH = W2L + 4096;               // W2^T [N = k1][K = n2]  (dgrad layer 2)
    static constexpr int W2TL = W2TH + 4096;
    static constexpr int W1TH = W2TL + 4096;              // W1^T [N = 16 (8 used)][K = n1 32]  (dgrad layer 1)
    static constexpr int W1TL = W1TH + 2048;
    static constexpr int VEC = W1TL + 2048;               // b1[32] b2[32] w3[32] b3 + pad            400 B
    static constexpr int DX = VEC + 400;                  // [slot 2][128][8] fp32                     8 192 B
    static constexpr int RED = DX + 8192;                 // dw3[32] + db3 (block reduction)            144 B
    static constexpr int BAR = RED + 144;                 // x_full[2], dx_full[2], mma_done : 5 x 8 B; tmem base 4 B
    static constexpr int BYTES = BAR + 48;
};
static_assert(TP::DX % 16 == 0 && TP::BAR % 8 == 0, "alignment");

// instruction descriptor (kind::tf32, fp32 accumulate): N >> 3 at bit 17, M >> 4 at bit 24, A / B major at bits 15 / 16
__host__ __device__ constexpr uint32_t tc_idesc(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRound >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ep_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// 3xTF32 product D (+)= A B over `ksteps` K = 8 steps; operand start addresses advance by a_step / b_step per K step
__device__ __forceinline__ void mma3x(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int ksteps,
                                      uint32_t a_step, uint32_t b_step, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo,
                                      uint32_t b_sbo, uint32_t idesc, uint32_t first_accumulate) {
    uint32_t acc = first_accumulate;
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t ah = umma_desc(a_hi + k * a_step, a_lbo, a_sbo), al = umma_desc(a_lo + k * a_step, a_lbo, a_sbo);
        const uint64_t bh = umma_desc(b_hi + k * b_step, b_lbo, b_sbo), bl = umma_desc(b_lo + k * b_step, b_lbo, b_sbo);
        umma_tf32(d, al, bh, idesc, acc);
        umma_tf32(d, ah, bl, idesc, 1u);
        umma_tf32(d, ah, bh, idesc, 1u);
        acc = 1u;
    }
}

template <bool DEC_GRAD>
__global__ void __launch_bounds__(kTcThreads, 1) sdf_train_tc_kernel(const __grid_constant__ StepParams P) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    float* vec = reinterpret_cast<float*>(sm + TP::VEC);
    const uint32_t bar_x = sbase + TP::BAR, bar_dx = sbase + TP::BAR + 16, bar_mma = sbase + TP::BAR + 32;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + TP::BAR + 40);

    // ---- prologue: zero the operand buffers (the padded MN extents read other rows: keep them finite), the constant
    //      1-column, weights in UMMA layouts (hi/lo), barriers, TMEM -------------------------------------------------------
    for (int i = tid; i < (TP::W1H - TP::HX) / 16; i += kTcThreads) reinterpret_cast<uint4*>(sm + TP::HX)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < 2 * kRound; i += kTcThreads) {          // column 40 (chunk 10, element 0) = 1.0 in both slots' hi copy
        const int slot = i / kRound, pt = i % kRound;
        *reinterpret_cast<float*>(sm + TP::HX + (2 * slot) * TP::HX_BYTES + (pt >> 3) * TP::HX_PT + 10 * 128 + (pt & 7) * 16) = 1.0f;
    }
    for (int i = tid; i < kH * kF; i += kTcThreads) {
        const int n = i / kF, k = i % kF;
        uint32_t hi, lo; split_tf32(P.dec.w1[i], hi, lo);
        const int off = (n >> 3) * 256 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;               // W1 [n][k]
        *reinterpret_cast<uint32_t*>(sm + TP::W1H + off) = hi; *reinterpret_cast<uint32_t*>(sm + TP::W1L + off) = lo;
        const int offt = (k >> 3) * 1024 + (n >> 2) * 128 + (k & 7) * 16 + (n & 3) * 4;              // W1^T [k][n] (rows 8..15 zero)
        *reinterpret_cast<uint32_t*>(sm + TP::W1TH + offt) = hi; *reinterpret_cast<uint32_t*>(sm + TP::W1TL + offt) = lo;
    }
    for (int i = tid; i < 8 * kH; i += kTcThreads) {               // zero rows 8..15 of W1^T
        const int k = 8 + i / kH, n = i % kH;
        const int offt = (k >> 3) * 1024 + (n >> 2) * 128 + (k & 7) * 16 + (n & 3) * 4;
        *reinterpret_cast<uint32_t*>(sm + TP::W1TH + offt) = 0u; *reinterpret_cast<uint32_t*>(sm + TP::W1TL + offt) = 0u;
    }
    for (int i = tid; i < kH * kH; i += kTcThreads) {
        const int n = i / kH, k = i % kH;
        uint32_t hi, lo; split_tf32(P.dec.w2[i], hi, lo);
        const int off = (n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4;              // W2 [n2][k1]
        *reinterpret_cast<uint32_t*>(sm + TP::W2H + off) = hi; *reinterpret_cast<uint32_t*>(sm + TP::W2L + off) = lo;
        const int offt = (k >> 3) * 1024 + (n >> 2) * 128 + (k & 7) * 16 + (n & 3) * 4;             // W2^T [k1][n2]
        *reinterpret_cast<uint32_t*>(sm + TP::W2TH + offt) = hi; *reinterpret_cast<uint32_t*>(sm + TP::W2TL + offt) = lo;
    }
    if (tid < kH) {
        vec[tid] = P.dec.b1 ? P.dec.b1[tid] : 0.f;
        vec[32 + tid] = P.dec.b2 ? P.dec.b2[tid] : 0.f;
        vec[64 + tid] = P.dec.w3[tid];
    }
    if (tid < 36) reinterpret_cast<float*>(sm + TP::RED)[tid] = 0.f;
    if (tid == 0) {
        vec[96] = P.dec.b3 ? P.dec.b3[0] : 0.f;
        mbar_init(bar_x, 8); mbar_init(bar_x + 8, 8);            // one arrival per gather warp of the group
        mbar_init(bar_dx, 1); mbar_init(bar_dx + 8, 1);          // one arrival by the issuing epilogue thread
        mbar_init(bar_mma, 1);                                   // tcgen05.commit
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kGSWarps) {       // the first epilogue warp allocates TMEM: 256 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(sbase + TP::BAR + 40) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // TMEM columns: D1/D3 0..31 | D2 32..63 | D4 64..79 | Dw 80..127 | gather-warp parking 128..255
    constexpr uint32_t cD1 = 0, cD2 = 32, cD4 = 64, cDw = 80, cPark = 128;

    const int64_t tiles_total = (P.n + kRound - 1) / kRound;
    const int rounds = (int)((tiles_total > blockIdx.x) ? (tiles_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
    const float up = P.d_loss ? __ldg(P.d_loss) : 1.0f;
    const float gscale = P.loss_scale * up;

    if (warp < kGSWarps) {
        // ============================== gather / scatter warps =====================================================
        const int grp = warp >> 3, wg = warp & 7;
        const int g = lane >> 2, t = lane & 3, odd = t & 1, half = t >> 1;
        const int row = 16 * wg + g + 8 * odd;                                   // this lane pair's row of the tile
        const uint32_t tpark = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + cPark + 32u * (uint32_t)(warp >> 2);
        const bool poly = P.oct.poly_interp != 0;
        const int L = P.oct.num_levels;
        bool consecutive = true;
#pragma unroll
        for (int i = 1; i < 4; ++i)
            if (i < L && P.oct.lv[i].level != P.oct.lv[0].level - i) consecutive = false;
        unsigned char* hx_hi = sm + TP::HX + (2 * grp) * TP::HX_BYTES;
        unsigned char* hx_lo = hx_hi + TP::HX_BYTES;
        const float* dxt = reinterpret_cast<const float*>(sm + TP::DX + grp * 4096);
        const int xoff = (row >> 3) * TP::HX_PT + (8 + half) * 128 + (row & 7) * 16;

        for (int r = grp; r < rounds; r += 2) {
            const int64_t base = ((int64_t)blockIdx.x + (int64_t)r * gridDim.x) * kRound;
            const int64_t myp = base + row;
            const bool valid = myp < P.n;
            float x = 0.f, y = 0.f, z = 0.f;
            if (valid) { x = __ldg(P.coord + 3 * myp); y = __ldg(P.coord + 3 * myp + 1); z = __ldg(P.coord + 3 * myp + 2); }

            // ---- hash walk (model/feature_octree.py:199-218): the pair splits the LEVELS for the first probe ------------
            int slot[4];
            {
                const unsigned long long key0 = valid ? morton_of(x, y, z, P.oct.lv[0].level) : 0ull;
                unsigned long long kq[2], kf[2];
                int mine[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = 2 * j + half;
                    mine[j] = -1;
                    if (i < L && valid) {
                        const shine_level& lv = P.oct.lv[i];
                        kq[j] = consecutive ? (key0 >> (3 * i)) : morton_of(x, y, z, lv.level);
                        mine[j] = (int)(hash_key(kq[j]) & (lv.hash_capacity - 1));
                        kf[j] = __ldg(&reinterpret_cast<const HashSlot*>(lv.hash_slots)[mine[j]].key);
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = 2 * j + half;
                    if (i < L && valid && kf[j] != kq[j]) {
                        if (kf[j] == kEmptyKey) mine[j] = -1;
                        else {
                            const shine_level& lv = P.oct.lv[i];
                            mine[j] = probe_slot_from(reinterpret_cast<const HashSlot*>(lv.hash_slots), lv.hash_capacity - 1,
                                                      kq[j], (uint32_t)mine[j], 1u);
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int other = __shfl_xor_sync(kFull, mine[j], 2);
                    slot[2 * j] = half ? other : mine[j];
                    slot[2 * j + 1] = half ? mine[j] : other;
                }
            }
            // ---- 8-corner gather + blend (model/feature_octree.py:222-234): the pair splits the CORNERS by z bit ----------
            float pk[16], idp[16];
            float feat[4];
            {
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 16; ++i) { pk[i] = 0.f; idp[i] = __int_as_float(-1); }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < L && slot[i] >= 0) {
                        const shine_level& lv = P.oct.lv[i];
                        const int4 id4 = ldg_i4(slot_ids(reinterpret_cast<const HashSlot*>(lv.hash_slots), slot[i], half));
                        idp[4 * i] = __int_as_float(id4.x); idp[4 * i + 1] = __int_as_float(id4.y);
                        idp[4 * i + 2] = __int_as_float(id4.z); idp[4 * i + 3] = __int_as_float(id4.w);
                        float r0[8], r1[8], r2[8], r3[8];
                        ldg_row8(lv.features + (int64_t)id4.x * kF, r0);
                        ldg_row8(lv.features + (int64_t)id4.y * kF, r1);
                        ldg_row8(lv.features + (int64_t)id4.z * kF, r2);
                        ldg_row8(lv.features + (int64_t)id4.w * kF, r3);
                        Blend b; b.init(x, y, z, lv.level, poly);
                        pk[3 * i] = b.tx; pk[3 * i + 1] = b.ty; pk[3 * i + 2] = b.tz;
                        const float wz = half ? b.tz : b.uz;
                        const float w0 = __fmul_rn(__fmul_rn(b.ux, b.uy), wz), w1 = __fmul_rn(__fmul_rn(b.ux, b.ty), wz);
                        const float w2 = __fmul_rn(__fmul_rn(b.tx, b.uy), wz), w3 = __fmul_rn(__fmul_rn(b.tx, b.ty), wz);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float a = acc[q];
                            a = fmaf(w0, r0[q], a); a = fmaf(w1, r1[q], a); a = fmaf(w2, r2[q], a); a = fmaf(w3, r3[q], a);
                            acc[q] = a;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float send = half ? acc[q] : acc[4 + q];
                    const float recv = __shfl_xor_sync(kFull, send, 2);
                    feat[q] = (half ? acc[4 + q] : acc[q]) + recv;
                }
            }
            tmem_st16(tpark, pk); tmem_st16(tpark + 16, idp);
            // X rows (this lane: the 4 channels of its half = one 16-byte K chunk), hi / lo
            {
                uint32_t h4[4], l4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) split_fast(feat[q], h4[q], l4[q]);
                *reinterpret_cast<uint4*>(hx_hi + xoff) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
                *reinterpret_cast<uint4*>(hx_lo + xoff) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
            }
            tmem_wait_st();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_x + 8 * grp);

            // ---- wait for dL/dfeature of this round, then scatter-add (index_put_ accumulate) ------------------------------
            mbar_wait(bar_dx + 8 * grp, (uint32_t)((r >> 1) & 1));
            const float4 dxv = *reinterpret_cast<const float4*>(dxt + row * 8 + 4 * half);
            const float dx[4] = {dxv.x, dxv.y, dxv.z, dxv.w};
            float qk[16], qid[16];
            tmem_ld16(tpark, qk); tmem_ld16(tpark + 16, qid);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int ids[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int mine = __float_as_int(qid[4 * i + k]);
                    const int other = __shfl_xor_sync(kFull, mine, 2);
                    ids[2 * k] = half ? other : mine;
                    ids[2 * k + 1] = half ? mine : other;
                }
                if (i < L && ids[0] >= 0) {
                    const shine_level& lv = P.oct.lv[i];
                    Blend b;
                    b.tx = qk[3 * i]; b.ty = qk[3 * i + 1]; b.tz = qk[3 * i + 2];
                    b.ux = __fsub_rn(1.0f, b.tx); b.uy = __fsub_rn(1.0f, b.ty); b.uz = __fsub_rn(1.0f, b.tz);
                    float* gb = grad_base(lv, (uint32_t)(blockIdx.x * kGSWarps + warp + r), kF) + 4 * half;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float w = b.w(c);
                        red_add_f4(gb + (int64_t)ids[c] * kF, w * dx[0], w * dx[1], w * dx[2], w * dx[3]);
                    }
                }
            }
            __syncwarp();
        }
    } else {
        // ============================== decoder epilogue warps (+ MMA issue by thread 0) ==============================
        const int et = tid - 32 * kGSWarps;                     // row of the tile owned by this thread
        const int eq = et >> 5;                                 // TMEM lane quadrant == warp % 4
        const uint32_t trow = tmem + ((uint32_t)(32 * eq) << 16);
        const uint32_t dh_hi = sbase + TP::DH, dh_lo = dh_hi + TP::DH_BYTES;
        constexpr uint32_t idN32 = tc_idesc(32, false, false), idN16 = tc_idesc(16, false, false), idW = tc_idesc(48, true, true);
        uint32_t mph = 0;                                       // parity of the next mma_done completion
        float dw3acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) dw3acc[j] = 0.f;
        float db3acc = 0.f, loss_acc = 0.f;
        const int hoff = (et >> 3) * TP::HX_PT + (et & 7) * 16;           // + chunk * 128
        const int doff = (et >> 3) * TP::DH_PT + (et & 7) * 16;

        for (int r = 0; r < rounds; ++r) {
            const int slot = r & 1;
            const int64_t p = ((int64_t)blockIdx.x + (int64_t)r * gridDim.x) * kRound + et;
            const bool valid = p < P.n;
            float lab = 0.f, wgt = 1.f;
            if (valid) {
                lab = __ldg(P.label + p);
                if (P.weighted) wgt = fabsf(__ldg(P.weight + p));                 // shine_batch.py:172 abs()
            }
            const uint32_t hx_hi = sbase + TP::HX + (2 * slot) * TP::HX_BYTES, hx_lo = hx_hi + TP::HX_BYTES;
            unsigned char* hxp_hi = sm + TP::HX + (2 * slot) * TP::HX_BYTES;
            unsigned char* hxp_lo = hxp_hi + TP::HX_BYTES;

            // ---- layer 1: D1 = X W1^T --------------------------------------------------------------------------------
            mbar_wait(bar_x + 8 * slot, (uint32_t)((r >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (et == 0) {
                mma3x(tmem + cD1, hx_hi + 8 * 128, hx_lo + 8 * 128, sbase + TP::W1H, sbase + TP::W1L, 1, 256, 256,
                      128, TP::HX_PT, 128, 256, idN32, 0u);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float hv[32];
            tmem_ld32(trow + cD1, hv);
            tmem_wait_ld();
            uint32_t m1 = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 bb = *reinterpret_cast<const float4*>(vec + 4 * c);
                float v0 = hv[4 * c] + bb.x, v1 = hv[4 * c + 1] + bb.y, v2 = hv[4 * c + 2] + bb.z, v3 = hv[4 * c + 3] + bb.w;
                m1 |= (v0 > 0.f ? 1u : 0u) << (4 * c) | (v1 > 0.f ? 1u : 0u) << (4 * c + 1) | (v2 > 0.f ? 1u : 0u) << (4 * c + 2) |
                      (v3 > 0.f ? 1u : 0u) << (4 * c + 3);
                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_fast(fmaxf(v0, 0.f), h0, l0); split_fast(fmaxf(v1, 0.f), h1, l1);
                split_fast(fmaxf(v2, 0.f), h2, l2); split_fast(fmaxf(v3, 0.f), h3, l3);
                *reinterpret_cast<uint4*>(hxp_hi + hoff + 128 * c) = make_uint4(h0, h1, h2, h3);
                *reinterpret_cast<uint4*>(hxp_lo + hoff + 128 * c) = make_uint4(l0, l1, l2, l3);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ep_bar();

            // ---- layer 2: D2 = H1 W2^T, output layer, loss, dL/dpred ------------------------------------------------------
            if (et == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mma3x(tmem + cD2, hx_hi, hx_lo, sbase + TP::W2H, sbase + TP::W2L, 4, 256, 256, 128, TP::HX_PT, 128, 1024,
                      idN32, 0u);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem_ld32(trow + cD2, hv);
            tmem_wait_ld();
            float pr = vec[96];
            uint32_t m2 = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float v = hv[j] + vec[32 + j];
                m2 |= (v > 0.f ? 1u : 0u) << j;
                hv[j] = fmaxf(v, 0.f);
                pr = fmaf(hv[j], vec[64 + j], pr);
            }
            if (P.pred && valid) P.pred[p] = pr;
            float dp = 0.f;
            if (valid) {      // sdf_bce_loss (utils/loss.py:17-24), MUFU-based like the mma.sync kernel
                const float zt = __fdividef(1.0f, 1.0f + __expf(-__fdividef(lab, P.sigma)));
                const float e = __expf(-fabsf(pr));
                loss_acc += wgt * (fmaxf(pr, 0.f) - pr * zt + __logf(1.0f + e));
                const float rs = __fdividef(1.0f, 1.0f + e);
                dp = ((pr >= 0.f ? rs : e * rs) - zt) * wgt * gscale;
            }
            db3acc += dp;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float d[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = 4 * c + q;
                    dw3acc[j] = fmaf(dp, hv[j], dw3acc[j]);
                    d[q] = ((m2 >> j) & 1u) ? dp * vec[64 + j] : 0.f;
                }
                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_fast(d[0], h0, l0); split_fast(d[1], h1, l1); split_fast(d[2], h2, l2); split_fast(d[3], h3, l3);
                *reinterpret_cast<uint4*>(sm + TP::DH + doff + 128 * c) = make_uint4(h0, h1, h2, h3);
                *reinterpret_cast<uint4*>(sm + TP::DH + TP::DH_BYTES + doff + 128 * c) = make_uint4(l0, l1, l2, l3);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ep_bar();

            // ---- dgrad layer 2: D3 = dH2 W2 ------------------------------------------------------------------------------
            if (et == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mma3x(tmem + cD1, dh_hi, dh_lo, sbase + TP::W2TH, sbase + TP::W2TL, 4, 256, 256, 128, TP::DH_PT, 128, 1024,
                      idN32, 0u);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem_ld32(trow + cD1, hv);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_fast(((m1 >> (4 * c)) & 1u) ? hv[4 * c] : 0.f, h0, l0);
                split_fast(((m1 >> (4 * c + 1)) & 1u) ? hv[4 * c + 1] : 0.f, h1, l1);
                split_fast(((m1 >> (4 * c + 2)) & 1u) ? hv[4 * c + 2] : 0.f, h2, l2);
                split_fast(((m1 >> (4 * c + 3)) & 1u) ? hv[4 * c + 3] : 0.f, h3, l3);
                *reinterpret_cast<uint4*>(sm + TP::DH + doff + 128 * (8 + c)) = make_uint4(h0, h1, h2, h3);
                *reinterpret_cast<uint4*>(sm + TP::DH + TP::DH_BYTES + doff + 128 * (8 + c)) = make_uint4(l0, l1, l2, l3);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ep_bar();

            // ---- dgrad layer 1: D4 = dH1 W1; weight gradients: Dw += [dH2|dH1]^T [H1|X|1] -----------------------------------
            if (et == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mma3x(tmem + cD4, dh_hi + 8 * 128, dh_lo + 8 * 128, sbase + TP::W1TH, sbase + TP::W1TL, 4, 256, 256, 128,
                      TP::DH_PT, 128, 1024, idN16, 0u);
                if (DEC_GRAD)
                    mma3x(tmem + cDw, dh_hi, dh_lo, hx_hi, hx_lo, 16, TP::DH_PT, TP::HX_PT, TP::DH_PT, 128, TP::HX_PT, 128, idW,
                          r > 0 ? 1u : 0u);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, mph); mph ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float dxr[8];
            tmem_ld8(trow + cD4, dxr);
            tmem_wait_ld();
            float* dxo = reinterpret_cast<float*>(sm + TP::DX + slot * 4096) + et * 8;
            *reinterpret_cast<float4*>(dxo) = make_float4(dxr[0], dxr[1], dxr[2], dxr[3]);
            *reinterpret_cast<float4*>(dxo + 4) = make_float4(dxr[4], dxr[5], dxr[6], dxr[7]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            ep_bar();
            if (et == 0) mbar_arrive(bar_dx + 8 * slot);
        }

        // ---- epilogue: loss, decoder gradients ---------------------------------------------------------------------------
        if (P.loss) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(kFull, loss_acc, o);
            if (lane == 0 && loss_acc != 0.f) atomicAdd(P.loss, loss_acc * P.loss_scale);
        }
        if (DEC_GRAD && rounds > 0) {
            float* red = reinterpret_cast<float*>(sm + TP::RED);
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float s = dw3acc[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
                if (lane == j) mine = s;
            }
            atomicAdd(red + lane, mine);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) db3acc += __shfl_xor_sync(kFull, db3acc, o);
            if (lane == 0) atomicAdd(red + 32, db3acc);
            // Dw rows 0..31 = [dW2[n2][0..31] | dW1 junk | db2[n2] at column 40]; rows 32..63 = [junk | dW1[n1][0..7] | db1[n1]]
            if (eq < 2) {
                float w[48];
                tmem_ld32(trow + cDw, w); tmem_ld16(trow + cDw + 32, w + 32);
                tmem_wait_ld();
                if (eq == 0) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) if (w[k] != 0.f) atomicAdd(P.dec.gw2 + lane * kH + k, w[k]);
                    if (P.dec.gb2 && w[40] != 0.f) atomicAdd(P.dec.gb2 + lane, w[40]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (w[32 + k] != 0.f) atomicAdd(P.dec.gw1 + lane * kF + k, w[32 + k]);
                    if (P.dec.gb1 && w[40] != 0.f) atomicAdd(P.dec.gb1 + lane, w[40]);
                }
            }
            ep_bar();
            if (et < 32) { const float v = red[et]; if (v != 0.f) atomicAdd(P.dec.gw3 + et, v); }
            if (et == 32 && P.dec.gb3) { const float v = red[32]; if (v != 0.f) atomicAdd(P.dec.gb3, v); }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kGSWarps) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}

}  // namespace

namespace shine_internal {

int launch_train_tc(const StepParams& P, bool dec_grad, cudaStream_t st) {
    if (P.oct.num_levels > 4 || P.oct.feature_dim != kF) return SHINE_ERR_UNSUPPORTED;
    auto kern = dec_grad ? sdf_train_tc_kernel<true> : sdf_train_tc_kernel<false>;
    static int ready[2][kMaxDevices] = {{0}};
    int& done = ready[dec_grad ? 1 : 0][current_device()];
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TP::BYTES);
        if (e != cudaSuccess) return (int)e;
        done = 1;
    }
    const int64_t tiles = (P.n + kRound - 1) / kRound;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kTcThreads, TP::BYTES, st>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace shine_internal
