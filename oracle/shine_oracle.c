/* Plain-C restatement of the FORWARD half of SHINE's per-point SDF step.  *** TEST INFRASTRUCTURE ***
 *
 * Independent of oracle/shine_oracle.py (numpy + torch) and of the CUDA code: written from the reference's behaviour,
 * each function citing the reference file:line (PRBonn/SHINE_mapping @ 0fbaf8a).  Only tests/ and
 * __graft_entry__.build()/smoke() may use it; the product never links it.
 *
 * The node table (reference: Python dict nodes_lookup_tables[level], model/feature_octree.py:46-52) is passed as a
 * SORTED array of Morton keys + the 8 corner rows of each, looked up by binary search.
 *
 *   gcc -O2 -shared -fPIC -o oracle/_build/libshine_oracle_c.so oracle/shine_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>

/* kal.ops.spc.quantize_points (call site model/feature_octree.py:203): floor(clamp(res*(x+1)/2, 0, res-1)) in fp32 */
static int32_t quantize1(float x, int level) {
    const float res = (float)(1 << level);
    volatile float a = x + 1.0f;            /* volatile: keep the reference's fp32 rounding steps, no contraction */
    volatile float b = res * a;
    volatile float v = b / 2.0f;
    if (v < 0.0f) v = 0.0f;
    if (v > res - 1.0f) v = res - 1.0f;
    return (int32_t)floorf(v);
}

/* kal.ops.spc.points_to_morton (call site model/feature_octree.py:204): x -> bit 3i+2, y -> 3i+1, z -> 3i */
static int64_t morton3(int32_t x, int32_t y, int32_t z) {
    int64_t m = 0;
    for (int i = 0; i < 16; ++i) {
        m |= ((int64_t)((x >> i) & 1)) << (3 * i + 2);
        m |= ((int64_t)((y >> i) & 1)) << (3 * i + 1);
        m |= ((int64_t)((z >> i) & 1)) << (3 * i);
    }
    return m;
}

void orc_points_to_morton(const float* xyz, int64_t n, int32_t level, int64_t* out) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = morton3(quantize1(xyz[3 * i], level), quantize1(xyz[3 * i + 1], level), quantize1(xyz[3 * i + 2], level));
}

static int64_t find_key(const int64_t* keys, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int64_t mid = (lo + hi) / 2;
        if (keys[mid] == key) return mid;
        if (keys[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* FeatureOctree.interpolat (model/feature_octree.py:172-196): t per axis, then the 8 products in (x,y,z) bit order */
static void blend_weights(const float* p, int level, int poly, float w[8]) {
    float t[3], u[3];
    for (int a = 0; a < 3; ++a) {
        volatile float h = p[a] * 0.5f;
        volatile float s = h + 0.5f;
        volatile float c = (float)(1 << level) * s;
        volatile float d = c - truncf(c);                       /* torch.frac */
        if (poly) {
            volatile float d2 = d * d;
            volatile float d3 = d2 * d;
            volatile float a3 = 3.0f * d2;
            volatile float b2 = 2.0f * d3;
            t[a] = a3 - b2;
        } else {
            t[a] = d;
        }
        u[a] = 1.0f - t[a];
    }
    for (int c = 0; c < 8; ++c) {
        volatile float xy = ((c & 4) ? t[0] : u[0]) * ((c & 2) ? t[1] : u[1]);
        w[c] = xy * ((c & 1) ? t[2] : u[2]);
    }
}

/* get_indices (model/feature_octree.py:199-218) + query_feature_with_indices (:222-234) + Decoder.sdf
 * (model/decoder.py:49-63) + sdf_bce_loss (utils/loss.py:17-24, unweighted).
 * Levels bottom-up: level i = max_level - i uses keys[i] (sorted, nkeys[i]) / ids[i] ([nkeys,8]) / tables[i] ([rows,F]).
 * Outputs: idx [L,n,8] (-1 on miss), feat [n,F], pred [n], returns the summed loss (caller divides for "mean"). */
double orc_forward(const float* xyz, const float* label, int64_t n, int32_t max_level, int32_t L, int32_t F, int32_t poly,
                   const int64_t* const* keys, const int64_t* nkeys, const int32_t* const* ids, const float* const* tables,
                   const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                   int32_t H, float sigma, int64_t* idx, float* feat, float* pred) {
    double loss = 0.0;
    for (int64_t p = 0; p < n; ++p) {
        float f[64];
        for (int k = 0; k < F; ++k) f[k] = 0.0f;
        for (int i = 0; i < L; ++i) {
            const int level = max_level - i;
            const int64_t key = morton3(quantize1(xyz[3 * p], level), quantize1(xyz[3 * p + 1], level),
                                        quantize1(xyz[3 * p + 2], level));
            const int64_t pos = find_key(keys[i], nkeys[i], key);
            int64_t* out = idx + ((int64_t)i * n + p) * 8;
            if (pos < 0) { for (int c = 0; c < 8; ++c) out[c] = -1; continue; }     /* trash-bin row == zeros */
            float w[8];
            blend_weights(xyz + 3 * p, level, poly, w);
            float ls[64];
            for (int k = 0; k < F; ++k) ls[k] = 0.0f;
            for (int c = 0; c < 8; ++c) {
                const int32_t row = ids[i][pos * 8 + c];
                out[c] = row;
                for (int k = 0; k < F; ++k) ls[k] += w[c] * tables[i][(int64_t)row * F + k];
            }
            for (int k = 0; k < F; ++k) f[k] += ls[k];
        }
        for (int k = 0; k < F; ++k) feat[p * F + k] = f[k];
        float h1[256], h2[256];
        for (int j = 0; j < H; ++j) {
            float a = b1 ? b1[j] : 0.0f;
            for (int k = 0; k < F; ++k) a += w1[j * F + k] * f[k];
            h1[j] = a > 0.0f ? a : 0.0f;
        }
        for (int j = 0; j < H; ++j) {
            float a = b2 ? b2[j] : 0.0f;
            for (int k = 0; k < H; ++k) a += w2[j * H + k] * h1[k];
            h2[j] = a > 0.0f ? a : 0.0f;
        }
        float o = b3 ? b3[0] : 0.0f;
        for (int k = 0; k < H; ++k) o += w3[k] * h2[k];
        pred[p] = o;
        const double z = 1.0 / (1.0 + exp(-(double)label[p] / (double)sigma));
        loss += fmax((double)o, 0.0) - (double)o * z + log1p(exp(-fabs((double)o)));
    }
    return loss;
}
