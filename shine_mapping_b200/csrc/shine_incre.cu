// shine_incre.cu — continual-learning terms of the incremental mapping loop (BASELINE config 4) as per-touched-row
// kernels.  Reference: FeatureOctree.cal_regularization (model/feature_octree.py:246-255) and cal_feature_importance
// (utils/incre_learning.py:8-40).  The reference finds "the rows this batch touched" with torch.unique over the
// [N*8] index tensor of every level (a sort) and then works on dense [rows, F] tensors; here the touched rows are
// collected with one atomicOr per corner into a bitmap (first setter appends the row to a compact list), and both
// terms run over that list only.
#include "shine_device.cuh"

namespace {

// one thread per (point, level): hash walk, then mark the 8 corner rows
__global__ void __launch_bounds__(256) mark_touched_kernel(const __grid_constant__ shine_octree oct,
                                                           const __grid_constant__ shine_touched tch,
                                                           const float* __restrict__ coord, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lvl = blockIdx.y;
    if (i >= n) return;
    const shine_level& lv = oct.lv[lvl];
    const shine_touched_level& t = tch.lv[lvl];
    const HashSlot* slots = reinterpret_cast<const HashSlot*>(lv.hash_slots);
    const int s = probe_slot(slots, lv.hash_capacity - 1, morton_of(coord[3 * i], coord[3 * i + 1], coord[3 * i + 2], lv.level));
    if (s < 0) return;                       // miss: the reference's -1 row carries zero importance (utils/incre_learning.py:40)
    const int4 a = ldg_i4(slots[s].ids0), b = ldg_i4(slots[s].ids1);
    const int ids[8] = {a.x, b.x, a.y, b.y, a.z, b.z, a.w, b.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t bit = 1u << (ids[c] & 31);
        uint32_t* word = t.bitmap + (ids[c] >> 5);
        if (__ldg(word) & bit) continue;     // already marked by an earlier point (plain read: may be stale, then the atomic decides)
        const uint32_t old = atomicOr(word, bit);
        if (!(old & bit)) {
            const int pos = atomicAdd(t.count, 1);
            if (pos < t.capacity) t.rows[pos] = ids[c];
        }
    }
}

// MODE 0: regularisation   grads[u] += scale * Omega[u] * (f[u] - f_last[u]);  reg += sum Omega * (f - f_last)^2
// MODE 1: importance       Omega[u] += |grads[u]|  (optionally grads[u] = 0)
// F/4 adjacent lanes share a row (one float4 each).
template <int MODE>
__global__ void __launch_bounds__(256) touched_rows_kernel(const __grid_constant__ shine_octree oct,
                                                           const __grid_constant__ shine_touched tch,
                                                           const __grid_constant__ shine_row_tables aux, float scale,
                                                           float* __restrict__ out_reg, int clear, int zero_grads) {
    const int lvl = blockIdx.y;
    const shine_level& lv = oct.lv[lvl];
    const shine_touched_level& t = tch.lv[lvl];
    const int lp = oct.feature_dim >> 2;
    int count = *t.count;
    if (count > t.capacity) count = t.capacity;
    float local = 0.f;
    for (int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gt < (int64_t)count * lp;
         gt += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(gt / lp), part = (int)(gt % lp);
        const int u = t.rows[r];
        const int64_t off = (int64_t)u * oct.feature_dim + 4 * part;
        if (MODE == 0) {
            const float4 f = *reinterpret_cast<const float4*>(lv.features + off);
            const float4 fl = *reinterpret_cast<const float4*>(aux.last[lvl] + off);
            const float4 w = *reinterpret_cast<const float4*>(aux.importance[lvl] + off);
            const float4 d = make_float4(f.x - fl.x, f.y - fl.y, f.z - fl.z, f.w - fl.w);
            local += w.x * d.x * d.x + w.y * d.y * d.y + w.z * d.z * d.z + w.w * d.w * d.w;
            float4* g = reinterpret_cast<float4*>(lv.feature_grads + off);       // every row appears once: plain RMW
            float4 gv = *g;
            gv.x += scale * w.x * d.x; gv.y += scale * w.y * d.y; gv.z += scale * w.z * d.z; gv.w += scale * w.w * d.w;
            *g = gv;
        } else {
            float4* g = reinterpret_cast<float4*>(lv.feature_grads + off);
            float4* w = reinterpret_cast<float4*>(aux.importance_rw[lvl] + off);
            const float4 gv = *g;
            float4 wv = *w;
            wv.x += fabsf(gv.x); wv.y += fabsf(gv.y); wv.z += fabsf(gv.z); wv.w += fabsf(gv.w);
            *w = wv;
            if (zero_grads) *g = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (clear && part == 0) atomicAnd(t.bitmap + (u >> 5), ~(1u << (u & 31)));
    }
    if (MODE == 0 && out_reg) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(kFull, local, o);
        if ((threadIdx.x & 31) == 0 && local != 0.f) atomicAdd(out_reg, local);
    }
}

int check_touched(const shine_octree* oct, const shine_touched* t) {
    if (!t) return SHINE_ERR_INVALID_ARG;
    for (int i = 0; i < oct->num_levels; ++i) {
        const shine_touched_level& l = t->lv[i];
        if (!l.bitmap || !l.rows || !l.count || l.capacity < 1) return SHINE_ERR_INVALID_ARG;
    }
    return SHINE_OK;
}

unsigned list_blocks(const shine_octree* oct, const shine_touched* t) {
    int64_t most = 1;
    for (int i = 0; i < oct->num_levels; ++i)
        if (t->lv[i].capacity > most) most = t->lv[i].capacity;
    int64_t blocks = (most * (oct->feature_dim / 4) + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int shine_mark_touched(const shine_octree* oct, const float* coord, int64_t n, const shine_touched* touched, void* stream) {
    int rc = check_octree(oct, false);
    if (rc) return rc;
    if ((rc = check_touched(oct, touched))) return rc;
    if (n < 0 || (n > 0 && !coord)) return SHINE_ERR_INVALID_ARG;
    if (n == 0) return SHINE_OK;
    if ((rc = check_same_device(oct, coord))) return rc;
    DeviceGuard guard(oct->lv[0].features);
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)oct->num_levels);
    mark_touched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*oct, *touched, coord, n);
    return (int)cudaGetLastError();
}

int shine_regularization_apply(const shine_octree* oct, const shine_touched* touched, const shine_row_tables* aux,
                               float grad_scale, float* out_reg, int32_t clear_marks, void* stream) {
    int rc = check_octree(oct, true);
    if (rc) return rc;
    if ((rc = check_touched(oct, touched))) return rc;
    if (!aux) return SHINE_ERR_INVALID_ARG;
    for (int i = 0; i < oct->num_levels; ++i)
        if (!aux->last[i] || !aux->importance[i]) return SHINE_ERR_INVALID_ARG;
    DeviceGuard guard(oct->lv[0].features);
    dim3 grid(list_blocks(oct, touched), (unsigned)oct->num_levels);
    touched_rows_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(*oct, *touched, *aux, grad_scale, out_reg, clear_marks, 0);
    return (int)cudaGetLastError();
}

int shine_importance_accumulate(const shine_octree* oct, const shine_touched* touched, const shine_row_tables* aux,
                                int32_t zero_grads, int32_t clear_marks, void* stream) {
    int rc = check_octree(oct, true);
    if (rc) return rc;
    if ((rc = check_touched(oct, touched))) return rc;
    if (!aux) return SHINE_ERR_INVALID_ARG;
    for (int i = 0; i < oct->num_levels; ++i)
        if (!aux->importance_rw[i]) return SHINE_ERR_INVALID_ARG;
    DeviceGuard guard(oct->lv[0].features);
    dim3 grid(list_blocks(oct, touched), (unsigned)oct->num_levels);
    touched_rows_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(*oct, *touched, *aux, 0.f, nullptr, clear_marks, zero_grads);
    return (int)cudaGetLastError();
}

}  // extern "C"
