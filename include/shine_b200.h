/*
 * shine_b200.h — C ABI of the B200-native (sm_100a) implementation of SHINE-mapping's per-point SDF
 * training step.  Plain C: raw device pointers + sizes + a cudaStream_t passed as void*; no torch types.
 *
 * The reference (PRBonn/SHINE_mapping @ 0fbaf8a) is 100 % Python and has NO FFI boundary for this path:
 * the path sits behind three Python call sites of the training loop (shine_batch.py:123 `query_feature`,
 * :128 `Decoder.sdf`, :174 `sdf_bce_loss`, :209 `backward`).  Each entry point below names the reference
 * interface it replaces (file:line, relative to the reference root).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - return 0 on success, a positive cudaError_t on a CUDA failure, a negative SHINE_ERR_* on a bad
 *     argument / unsupported configuration.  shine_error_string() decodes either.
 *   - all buffers are caller-owned device memory; nothing is allocated, freed or synchronised inside;
 *     every call is asynchronous on `stream` and re-entrant.  Kernels are launched on the device that owns the
 *     buffers (looked up from the pointers), whatever the calling thread's current device is; a batch on another
 *     device than the tables is SHINE_ERR_INVALID_ARG (pinned host batch pointers are accepted).
 *   - levels are described BOTTOM-UP like `FeatureOctree.hierarchical_indices`
 *     (model/feature_octree.py:201-202): lv[0] is the leaf level `tree_level_world`.
 *   - a voxel that is not in the level's node table is a MISS: its eight corner ids are -1, its feature
 *     contribution is exactly 0 (the reference's zeroed "trash-bin" last row, model/feature_octree.py:76-81,
 *     205-213,232-233) and it receives no gradient.
 */
#ifndef SHINE_B200_H_
#define SHINE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHINE_ABI_VERSION 2
#define SHINE_MAX_LEVELS 8
#define SHINE_HASH_SLOT_BYTES 64

#define SHINE_OK 0
#define SHINE_ERR_INVALID_ARG (-1)
#define SHINE_ERR_UNSUPPORTED (-2)

/* flags of shine_sdf_* calls */
#define SHINE_FLAG_REDUCTION_SUM 1u   /* loss_reduction == "sum" (shine_incre.py:77-78); default mean   */
#define SHINE_FLAG_WEIGHTED 2u        /* loss_weight_on (utils/loss.py:18-19): per-sample weight applied */
#define SHINE_FLAG_TF32X1 4u          /* decoder contractions in plain TF32 (default: 3xTF32 ~ fp32)     */
#define SHINE_FLAG_TCGEN05 8u         /* shine_sdf_infer / shine_sdf_bce_step: decoder on tcgen05.mma with TMEM accumulators
                                         (128-point tiles, warp-specialised gather / epilogue warps) instead of
                                         warp-level mma.sync                                                  */
#define SHINE_FLAG_MORTON_ORDERED 16u  /* shine_sdf_bce_step: the batch is in Morton order of its coordinates (the order the
                                         Morton-sorted sample pool hands batches out in).  A hint, never a requirement:
                                         neighbouring points then share nodes, and the step sums the gradients of each run of
                                         equal node on the tensor cores before ONE red per corner row (results equal up to
                                         fp32 summation order); on an unordered batch the flag only costs time           */

/* One featured level of the FeatureOctree (model/feature_octree.py:46-63). */
typedef struct shine_level {
    const void* hash_slots;   /* capacity x 64-byte slots {u64 morton key | i32 node | pad | i32 ids[8]}:
                                 replaces nodes_lookup_tables[level] (dict morton -> 8 corner rows)      */
    const float* features;    /* hier_features[k]: [rows, F] fp32, last row = trash-bin                  */
    float* feature_grads;     /* same shape, accumulated into (+=); may be NULL when no grads are asked  */
    float* grad_replicas;     /* optional scratch [num_replicas-1, rows, F], all-zero on entry: gradient
                                 privatisation for small (hot) levels — warps spread their red.adds over the
                                 replicas to avoid same-address serialisation in L2; fold them back with
                                 shine_reduce_grad_replicas (which also re-zeroes the scratch)           */
    uint32_t hash_capacity;   /* power of two                                                            */
    int32_t rows;             /* N_l + 1                                                                 */
    int32_t level;            /* octree level in world numbering (leaf = tree_level_world)               */
    int32_t num_replicas;     /* 0/1 = none, else a power of two <= 64                                   */
} shine_level;

typedef struct shine_octree {
    int32_t num_levels;       /* L = tree_level_feat, 1..SHINE_MAX_LEVELS                                */
    int32_t feature_dim;      /* F = feature_dim (multiple of 4; fused sdf_* kernels need 8)             */
    int32_t poly_interp;      /* poly_int_on: smoothstep 3d^2-2d^3 weights (model/feature_octree.py:176) */
    int32_t reserved;
    shine_level lv[SHINE_MAX_LEVELS];
} shine_octree;

/* Geometry decoder (model/decoder.py:29-36): Linear(F,H)+ReLU, Linear(H,H)+ReLU, Linear(H,1);
 * PyTorch (out,in) row-major weights.  Bias pointers may be NULL (geo_mlp_bias_on False).
 * g* are gradient buffers (+=); all NULL == frozen decoder (utils/tools.py:188-191). */
typedef struct shine_decoder {
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    float *gw1, *gb1, *gw2, *gb2, *gw3, *gb3;
    int32_t in_dim;           /* F  (8)  */
    int32_t hidden;           /* H  (32) */
    int32_t mlp_level;        /* 2       */
    int32_t reserved;
} shine_decoder;

int shine_abi_version(void);
const char* shine_error_string(int code);

/* Build / extend the device node table of one level.  Replaces the Python dict fill at
 * model/feature_octree.py:162-166.  `slots` must have been memset to 0xFF (empty) before the first
 * insert.  keys: [n] int64 Morton codes, corner_ids: [n,8] int32 rows, node_base: ordinal of keys[0].
 * overflow_count (device int32, may be NULL): incremented once per key that could NOT be stored because the
 * table is full — a Python dict never drops a key, so the caller must grow the table and re-insert. */
int shine_hash_insert(void* slots, uint32_t capacity, const int64_t* keys, const int32_t* corner_ids,
                      int64_t n, int32_t node_base, int32_t* overflow_count, void* stream);

/* kal.ops.spc.quantize_points + points_to_morton (call sites model/feature_octree.py:203-204):
 * coord [n,3] fp32 -> morton [n] int64 at `level`. */
int shine_points_to_morton(const float* coord, int64_t n, int32_t level, int64_t* morton, void* stream);

/* FeatureOctree.get_indices (model/feature_octree.py:199-218): out_idx [L, n, 8] int64, level-major
 * bottom-up, -1 x8 on a miss. */
int shine_get_indices(const shine_octree* oct, const float* coord, int64_t n, int64_t* out_idx, void* stream);

/* FeatureOctree.query_feature (model/feature_octree.py:237-244; interpolat :172-196, blend :222-234):
 * out_feat [n, F] fp32 = sum over levels of the 8-corner blend. */
int shine_query_fwd(const shine_octree* oct, const float* coord, int64_t n, float* out_feat, void* stream);

/* autograd of the above (the index_put_(accumulate=True) of shine_batch.py:209):
 * lv[i].feature_grads[id] += w_c * dfeat[p]  for every hit corner. */
int shine_query_bwd(const shine_octree* oct, const float* coord, int64_t n, const float* dfeat, void* stream);

/* Coordinate derivatives of query_feature, for the eikonal / normal terms that the reference obtains with
 * torch.autograd.grad(pred, coord, create_graph=True) (utils/tools.py:175-185, shine_batch.py:141-142,183-185).
 * With dw_c/da the derivative of the interpolation weight of corner c w.r.t. axis a (incl. the smoothstep and the
 * 2^level/2 scaling of model/feature_octree.py:173-178):
 *   coord_grad : out_dcoord[p][a]  = sum_levels sum_c dw_c/da * <features[id_c], dfeat[p]>            ([n,3])
 *   tangent_fwd: out[p][:]         = sum_levels sum_c (sum_a tangent[p][a] dw_c/da) * features[id_c]   ([n,F])
 *   tangent_bwd: feature_grads[id_c] += (sum_a tangent[p][a] dw_c/da) * dfeat[p]
 * coord_grad is the backward of query_fwd w.r.t. coord; tangent_fwd / tangent_bwd are coord_grad's own backward
 * w.r.t. dfeat / the tables (double backward). */
int shine_query_coord_grad(const shine_octree* oct, const float* coord, int64_t n, const float* dfeat,
                           float* out_dcoord, void* stream);
int shine_query_tangent_fwd(const shine_octree* oct, const float* coord, int64_t n, const float* tangent,
                            float* out_feat, void* stream);
int shine_query_tangent_bwd(const shine_octree* oct, const float* coord, int64_t n, const float* tangent,
                            const float* dfeat, void* stream);

/* query_feature -> Decoder.sdf (model/decoder.py:49-63) fused, forward only (the mesher's query,
 * utils/mesher.py:60-72).  out_pred [n].  out_mask (optional, may be NULL) [n] uint8 = voxel present at
 * lv[mask_level] (utils/mesher.py:82-89). */
int shine_sdf_infer(const shine_octree* oct, const shine_decoder* dec, const float* coord, int64_t n,
                    float* out_pred, uint8_t* out_mask, int32_t mask_level, uint32_t flags, void* stream);

/* query_feature -> Decoder.sdf -> sdf_bce_loss (utils/loss.py:17-24), forward only.
 * label [n], weight [n] or NULL, sigma = sigma_sigmoid (shine_batch.py:87).
 * out_loss [1] fp32 is ACCUMULATED (+=): caller zeroes it.  loss_scale multiplies every per-point term
 * (1/N_global for "mean", 1 for "sum"). */
int shine_sdf_bce_fwd(const shine_octree* oct, const shine_decoder* dec, const float* coord,
                      const float* label, const float* weight, int64_t n, float sigma, float loss_scale,
                      float* out_pred, float* out_loss, uint32_t flags, void* stream);

/* The whole training step shine_batch.py:123-209 in ONE pass: forward, loss and the backward that
 * scatter-adds into lv[i].feature_grads and dec->g*.  d_loss: device scalar dL/dloss (NULL == 1).
 * out_pred / out_loss may be NULL (pure backward == "recompute" mode for a separate autograd backward). */
int shine_sdf_bce_step(const shine_octree* oct, const shine_decoder* dec, const float* coord,
                       const float* label, const float* weight, int64_t n, float sigma, float loss_scale,
                       const float* d_loss, float* out_pred, float* out_loss, uint32_t flags, void* stream);

/* feature_grads[l] += sum of grad_replicas[l][r]; grad_replicas[l] = 0.  No-op for levels without
 * replicas.  Part of the backward (the reference's index_put_ is one pass, shine_batch.py:209). */
int shine_reduce_grad_replicas(const shine_octree* oct, void* stream);

/* Dense Adam (utils/tools.py:78-79: betas (0.9,0.99), eps 1e-15, weight decay as L2 on grads) over up to
 * SHINE_ADAM_MAX_TENSORS tensors in one launch — shine_batch.py:210 `opt.step()`. */
#define SHINE_ADAM_MAX_TENSORS 16
typedef struct shine_adam_tensor {
    float* param; float* grad; float* exp_avg; float* exp_avg_sq;
    int64_t numel; float lr; float weight_decay;
} shine_adam_tensor;
int shine_adam_step(const shine_adam_tensor* tensors, int32_t count, float beta1, float beta2, float eps,
                    int32_t step, int32_t zero_grad, void* stream);

/* Same, CUDA-graph friendly: the step number lives on the device.  state = {int32 step, f32 bc1, f32 bc2_sqrt}
 * (12 bytes, zero-initialised by the caller once); every call increments step and refreshes the bias
 * corrections in a 1-thread kernel, then runs the Adam kernel reading them — so a captured graph can be replayed. */
int shine_adam_step_dev(const shine_adam_tensor* tensors, int32_t count, float beta1, float beta2, float eps,
                        void* state, int32_t zero_grad, void* stream);

/* ---- FeatureOctree.update on the GPU (model/feature_octree.py:114-166) ---------------------------------------------
 * All featured levels of one scan at once; lv[] here is coarse -> fine.  Scratch (frame-local key sets, lists) is
 * caller-owned; sets are arrays of u64 filled with 0xFF, capacity a power of two >= 2x the insertions.
 * Sequence: frame_nodes -> (host reads new_node_count, sizes tables) -> frame_corners -> (host reads new_corner_count)
 * -> sort_corners -> assign_rows -> fill_nodes.  Row numbering is the reference's: unseen corners of the new nodes,
 * unique, in lexicographic (x, y, z) order, appended after the existing rows (:131-151). */
typedef struct shine_build_level {
    void* node_slots;               /* the level's node table (64-byte slots), NULL while the level is empty          */
    void* corner_slots;             /* the level's corner table: 16-byte slots {u64 lexicographic key, i32 row, pad} */
    void* frame_node_set;           /* scratch u64 set                                                                */
    void* frame_corner_set;         /* scratch u64 set (frame_corners onwards)                                        */
    int64_t* new_node_keys;         /* out [<= n points] Morton keys of this scan's unseen nodes (unordered)         */
    int32_t* node_ids_out;          /* out [new nodes, 8] corner rows of those nodes (fill_nodes)                     */
    int64_t* corner_morton_out;     /* out [new corners] Morton code of each new row, in row order (assign_rows)      */
    uint32_t node_capacity, corner_capacity, frame_node_set_capacity, frame_corner_set_capacity;
    int32_t level;                  /* world level                                                                    */
    int32_t nodes_before, rows_before;   /* nodes / corner rows (without the trash row) the level already holds       */
    int32_t reserved;
} shine_build_level;
typedef struct shine_build {
    int32_t num_levels, max_level;
    int32_t* new_node_count;        /* device [L], zero on entry                                                      */
    int32_t* new_corner_count;      /* device [L], zero on entry                                                      */
    int32_t* new_corner_total;      /* device [1], zero on entry                                                      */
    uint64_t* new_corner_keys;      /* scratch [8 * sum new nodes]: (level index << 51) | lexicographic key           */
    shine_build_level lv[SHINE_MAX_LEVELS];
} shine_build;

int shine_octree_frame_nodes(const shine_build* plan, const float* points, int64_t n, void* stream);
int shine_octree_frame_corners(const shine_build* plan, int32_t max_new_nodes, void* stream);
int64_t shine_octree_sort_scratch_bytes(int32_t n);
int shine_octree_sort_corners(const void* keys_in, void* keys_out, int32_t n, void* scratch, int64_t scratch_bytes,
                              void* stream);
int shine_octree_assign_rows(const shine_build* plan, const void* sorted_keys, int32_t total, void* stream);
int shine_octree_fill_nodes(const shine_build* plan, int32_t max_new_nodes, int32_t* overflow_count, void* stream);
/* rebuild a corner table from the per-row Morton codes (after growing it, moving devices or unpickling) */
int shine_octree_corner_rehash(void* corner_slots, uint32_t capacity, const int64_t* corner_morton_by_row, int64_t rows,
                               void* stream);

/* ---- the step with `ekional_loss_on` (config/kitti/kitti_batch.yaml:46) as one kernel ---------------------------------
 * Replaces shine_batch.py:119-142,172-185,208-209 + utils/tools.py:175-185 (autograd.grad(pred, coord, create_graph=True)
 * and the double backward through gather, decoder and loss):
 *   g = sigma * d pred / d coord;   L = sdf_bce_loss(pred, label) + weight_e * mean_{weight > 0} (1 - |g|)^2
 * Accumulates dL/d(tables) into lv[i].feature_grads and dL/d(decoder) into dec->g* (both terms).
 *   weight     [n]: sign marks surface (+) / free-space (-) samples (shine_batch.py:137); |weight| multiplies the BCE
 *              term only with SHINE_FLAG_WEIGHTED
 *   n_surface  device int32: number of samples with weight > 0 (shine_count_positive); 0 -> no eikonal contribution
 *   out_pred [n] / out_grad [n,3] (g) may be NULL; out_loss (+=) BCE part; out_eikonal (+=) the mean, without weight_e */
int shine_count_positive(const float* values, int64_t n, int32_t* out_count, void* stream);
int shine_sdf_bce_eikonal_step(const shine_octree* oct, const shine_decoder* dec, const float* coord,
                               const float* label, const float* weight, int64_t n, float sigma, float loss_scale,
                               float weight_e, const int32_t* n_surface, float* out_pred, float* out_grad,
                               float* out_loss, float* out_eikonal, uint32_t flags, void* stream);

/* ---- continual-learning terms of the incremental loop (BASELINE config 4) ---------------------------------------
 * The reference finds the rows a batch touched with `hierarchical_indices[i].flatten().unique()` (a sort per level
 * per step, model/feature_octree.py:251) and then works on dense [rows, F] tensors.  Here the touched rows of a batch
 * are a compact per-level list: one bit per row in `bitmap` (all zero between uses), first setter appends the row. */
typedef struct shine_touched_level {
    uint32_t* bitmap;         /* [(rows + 31) / 32] words, zero on entry                                       */
    int32_t* rows;            /* [capacity] touched row ids, unordered, each row once                           */
    int32_t* count;           /* device scalar, zero on entry                                                   */
    int32_t capacity;         /* >= min(rows, 8 n)                                                              */
    int32_t reserved;
} shine_touched_level;
typedef struct shine_touched { shine_touched_level lv[SHINE_MAX_LEVELS]; } shine_touched;   /* bottom-up like shine_octree */

/* per-level row-aligned side tables of the incremental loop, bottom-up: features_last_frame / importance_weight
 * (model/feature_octree.py:70-72) */
typedef struct shine_row_tables {
    const float* last[SHINE_MAX_LEVELS];        /* features_last_frame (regularisation)                       */
    const float* importance[SHINE_MAX_LEVELS];  /* importance_weight, read  (regularisation)                  */
    float* importance_rw[SHINE_MAX_LEVELS];     /* importance_weight, updated (importance pass)               */
} shine_row_tables;

/* Collect the rows touched by `coord` (hit voxels only: the reference's -1 row has zero importance,
 * utils/incre_learning.py:40).  Replaces the unique() of model/feature_octree.py:251. */
int shine_mark_touched(const shine_octree* oct, const float* coord, int64_t n, const shine_touched* touched,
                       void* stream);

/* FeatureOctree.cal_regularization (model/feature_octree.py:246-255) and its gradient, over the touched rows only:
 *   *out_reg (+=)         sum_u Omega[u] * (f[u] - f_last[u])^2
 *   feature_grads[u] +=   grad_scale * Omega[u] * (f[u] - f_last[u])        (grad_scale = 2 * lambda_forget)
 * clear_marks != 0 also clears the bitmap bits of the processed rows (leave it 0 when another pass follows). */
int shine_regularization_apply(const shine_octree* oct, const shine_touched* touched, const shine_row_tables* aux,
                               float grad_scale, float* out_reg, int32_t clear_marks, void* stream);

/* cal_feature_importance's accumulation (utils/incre_learning.py:36-40) over the touched rows only:
 *   importance_rw[u] += |feature_grads[u]|;  zero_grads != 0: feature_grads[u] = 0 afterwards (:38). */
int shine_importance_accumulate(const shine_octree* oct, const shine_touched* touched, const shine_row_tables* aux,
                                int32_t zero_grads, int32_t clear_marks, void* stream);

/* ---- multi-GPU exchange (SURVEY.md 8e, 8b export (6); the reference is single-GPU) --------------------------------
 * One process per GPU.  The map is partitioned by Morton prefix at the coarsest featured level; every rank owns the
 * rows reachable from its blocks, so corner rows on a face between two blocks exist on both ranks.  Their gradients
 * are summed through a compact exchange buffer laid out [decoder grads | boundary rows of lv[0] | lv[1] | ...]:
 * pack -> ONE all-reduce (decoder + boundary) -> unpack. */
typedef struct shine_boundary_level {
    float* table;             /* this rank's [rows, F] gradient (or feature) table of the level                 */
    const int32_t* rows;      /* [count] local rows that are shared with another rank                           */
    const int32_t* slots;     /* [count] their positions in the level's globally agreed boundary list           */
    int64_t offset;           /* float offset of the level's segment in the exchange buffer (multiple of 4)     */
    int32_t count;
    int32_t reserved;
} shine_boundary_level;
typedef struct shine_boundary { shine_boundary_level lv[SHINE_MAX_LEVELS]; } shine_boundary;

/* buf[offset_l + slot * F ..] = table_l[row]   /   table_l[row] = buf[offset_l + slot * F ..] */
int shine_boundary_pack(const shine_boundary* plan, int32_t num_levels, int32_t feature_dim, float* buf, void* stream);
int shine_boundary_unpack(const shine_boundary* plan, int32_t num_levels, int32_t feature_dim, float* buf, void* stream);

/* NCCL communicator owned by this library (NCCL is bound with dlopen at run time).  Rank 0 makes the 128-byte
 * unique id, the caller distributes it (any transport), every rank creates its communicator on `device`. */
int shine_nccl_unique_id(void* out_id128);
int shine_nccl_comm_create(const void* id128, int32_t nranks, int32_t rank, int32_t device, void** out_comm);
int shine_nccl_comm_destroy(void* comm);
/* In-place sum all-reduce of `count` floats over NVLink, asynchronous on `stream`: the decoder-gradient exchange
 * that follows the backward (comm is the ncclComm_t made above, or any ncclComm_t of the same NCCL).
 * NCCL failures return -1000 - ncclResult_t; shine_comm_last_error() has the text. */
int shine_allreduce_decoder_grads(void* comm, float* buf, int64_t count, void* stream);
const char* shine_comm_last_error(void);

/* The same exchange as ONE kernel over NVLink peer memory (no NCCL): pack own [decoder | boundary rows] into an
 * IPC-shared buffer, publish a step flag into every peer's buffer, wait for the peers' flags, sum all ranks' buffers
 * in fixed rank order straight over NVLink, in place into dec_grads / the plan's table rows.  One process per GPU:
 * create (returns the 64-byte cudaIpcMemHandle_t of this rank's buffer) -> the caller all-gathers the handles ->
 * connect -> exchange every step (all ranks, same order).  plan offsets are relative to the exchange buffer whose first
 * dec_floats floats are the decoder segment, exactly as for shine_boundary_pack.  A peer that never shows up is a
 * counted timeout (shine_p2p_timeouts), not a hang. */
typedef struct shine_boundary_inverse {
    const int32_t* row_of_slot[SHINE_MAX_LEVELS];   /* [slots of the level] local row holding that shared corner, -1 if none */
    int32_t slots[SHINE_MAX_LEVELS];                /* length of the level's globally agreed boundary list           */
    const int32_t* holders[SHINE_MAX_LEVELS];       /* [slots] bit r set: rank r holds a row of that corner (its buffer is
                                                       read for the sum); NULL: every rank's buffer is read            */
} shine_boundary_inverse;
typedef struct shine_p2p shine_p2p;
int shine_p2p_create(int32_t nranks, int32_t rank, int32_t device, int64_t max_floats, void* out_handle64, shine_p2p** out);
int shine_p2p_connect(shine_p2p* ctx, const void* handles /* nranks x 64 B, rank order */);
int shine_p2p_exchange(shine_p2p* ctx, float* dec_grads, int64_t dec_floats, const shine_boundary* plan,
                       const shine_boundary_inverse* inverse, int32_t num_levels, int32_t feature_dim, void* stream);
int shine_p2p_timeouts(shine_p2p* ctx, int32_t* out_count);
int shine_p2p_destroy(shine_p2p* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SHINE_B200_H_ */
