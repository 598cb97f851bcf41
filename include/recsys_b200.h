/*
 * recsys_b200.h — C ABI of the B200-native CTR embedding hot path.
 *
 * The reference (neoyinyao/Recommender) has no FFI / plugin interface: the path sits behind
 * Keras layer calls and TensorFlow-internal ops.  Each entry point below names the reference
 * call site (file:line under /root/reference) or the TensorFlow op
 * whose work it replaces.  The Python mirror of the reference's call surface
 * (recommender_b200/layers.py, model.py) binds these symbols with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the library never
 *    allocates, frees or retains caller memory — scratch comes in through (ws, ws_bytes),
 *    sized by the matching *_workspace_bytes() query.
 *  - every call takes the CUDA stream to launch on (a cudaStream_t passed as void*), is
 *    asynchronous and stream-ordered; one process per GPU.
 *  - return value: RB_OK (0) or a negative rb_status; rb_last_error() gives the message
 *    (thread-local).  Nothing throws.
 *  - tables are fp32, row-major [rows, D] with row stride == D; rows must be aligned to the
 *    vector width the kernel uses (16 B if D % 4 == 0, 8 B if D % 2 == 0, else 4 B) and
 *    D / vector-width <= 32 (D <= 128 for D % 4 == 0).
 *  - indices are int32 or int64 (both occur: ctr/tfrecord_io.py:82 vs dien/train.py:101).
 */
#ifndef RECSYS_B200_H_
#define RECSYS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_VERSION 100  /* 0.1.0 */
#define RB_MAX_GRAD_SOURCES 16
#define RB_MAX_LOOKUP_GROUPS 8
#define RB_MAX_DENSE_TENSORS 32
#define RB_MAX_RANKS 8          /* GPUs of one NVSwitch box */
#define RB_IPC_HANDLE_BYTES 64  /* sizeof(cudaIpcMemHandle_t) */

typedef enum rb_status {
  RB_OK = 0,
  RB_ERR_ARG = -1,        /* null pointer, bad enum, negative size */
  RB_ERR_SHAPE = -2,      /* unsupported D / F / L combination */
  RB_ERR_ALIGN = -3,      /* pointer not aligned for the vector width */
  RB_ERR_WORKSPACE = -4,  /* workspace missing or too small */
  RB_ERR_CUDA = -5        /* a CUDA runtime call or launch failed */
} rb_status;

typedef enum rb_index_type { RB_I32 = 0, RB_I64 = 1 } rb_index_type;

/* element type of the interaction output / its incoming gradient */
typedef enum rb_float_type {
  RB_F32 = 0,
  RB_BF16 = 1,
  RB_BF16_ONES = 2  /* output rows only: bf16 whose FIRST pad column holds 1.0 (the others 0): a consumer Dense layer
                       whose padded kernel has a zero row there reads its bias gradient off its weight-gradient GEMM */
} rb_float_type;

/* how the fused lookup of rb_dot_interaction_fwd / _bwd copies table rows: past L1 (best for ids that rarely repeat), through L1
 * (best when few rows take most of the lookups: Zipf ids, the OOV row of a shared table), or decided on the device from a flag
 * that rb_sparse_bwd_prepare derives from the sorted ids of the step */
typedef enum rb_row_cache { RB_ROW_CACHE_L2 = 0, RB_ROW_CACHE_L1 = 1, RB_ROW_CACHE_AUTO = 2 } rb_row_cache;

/* pooling over the L positions of a bag */
typedef enum rb_pool_mode {
  RB_POOL_SUM = 1,          /* tf.reduce_sum(E, axis=1)              ctr/model.py:21        */
  RB_POOL_MEAN = 2,         /* sum / L                                                       */
  RB_POOL_MASKED_MEAN = 3   /* compute_his_average                   dien/layers.py:5-17    */
} rb_pool_mode;

typedef enum rb_optimizer {
  RB_OPT_SGD = 0,            /* var[r] -= lr*g            (the commented option, ctr/train.py:79) */
  RB_OPT_ADAGRAD = 1,        /* Keras Adagrad sparse apply (north star)             */
  RB_OPT_ADAM_LAZY = 2,      /* Keras Adam formula on touched rows only        */
  RB_OPT_ADAM_TF_DENSE = 3   /* exact Keras Adam._resource_apply_sparse: every row decays/moves */
} rb_optimizer;

/* Optimizer hyper-parameters.  Defaults of the reference: Adam() at ctr/train.py:80,84. */
typedef struct rb_opt_params {
  int32_t optimizer;   /* rb_optimizer */
  int32_t step;        /* t = iterations + 1 (>= 1); Adam only */
  float lr;            /* 1e-3 */
  float beta_1;        /* 0.9 */
  float beta_2;        /* 0.999 */
  float epsilon;       /* 1e-7 (outside the square root, Keras form) */
  const float* alpha_t_dev;  /* optional DEVICE f32[1]: when non-NULL, Adam reads alpha_t (rb_adam_alpha_t of the current
                                step) from here instead of deriving it from `step` — lets a captured CUDA graph be
                                replayed for successive steps (the host refreshes the value before each replay) */
} rb_opt_params;

/*
 * Where the gradient row of lookup position p comes from.
 * With bag b = p / L and l = p % L, the row is
 *     sum_k  src[k][ b*bag_stride[k] + l*pos_stride[k]  ..  + D )        (added left to right)
 * then scaled per `scale_mode`, then (optionally) the FM term g_fm[b]*(s[b,:] - W[row,:]) is
 * added.  Examples: dE[B,F,D] of the un-pooled lookup: L=F, bag_stride=F*D, pos_stride=D;
 * a pooled bag [B,ld]: pos_stride=0, bag_stride=ld; the slice dX[:, :26] of a [B,27,D] tensor:
 * bag_stride=27*D; ESMM/MMOE consumers of the concat [B,sum D_f]: one src per consumer.
 */
typedef enum rb_grad_scale {
  RB_SCALE_NONE = 0,
  RB_SCALE_MEAN = 1,         /* g / L                                  */
  RB_SCALE_MASKED_MEAN = 2,  /* mask ? g / count[b] : 0                dien/layers.py:13-16 */
  RB_SCALE_MASKED = 3        /* mask ? g : 0 (no division)             dien/layers.py:53-55: gradient rows of the masked
                                positions are exactly zero and are never read — the producer need not write them */
} rb_grad_scale;

typedef struct rb_grad_source {
  int32_t num_src;                           /* 1..RB_MAX_GRAD_SOURCES */
  int32_t scale_mode;                        /* rb_grad_scale */
  const float* src[RB_MAX_GRAD_SOURCES];
  int64_t bag_stride[RB_MAX_GRAD_SOURCES];   /* in elements */
  int64_t pos_stride[RB_MAX_GRAD_SOURCES];   /* in elements */
  const void* mask_idx;                      /* RB_SCALE_MASKED_MEAN: index array [n] whose != 0 is the mask
                                                (same index type as idx); NULL = idx itself */
  const float* count;                        /* RB_SCALE_MASKED_MEAN: f32[B] valid counts from the forward */
  const float* fm_g;                         /* optional f32[B]: dL/d(fm)                 ctr/model.py:21-23 */
  const float* fm_s;                         /* optional f32[B,D]: sum_f E[b,f,:] from rb_gather_fm_fwd */
} rb_grad_source;

/*
 * One use of a table inside a training step.  A table that is looked up several times per step
 * (dien/model.py:26-30: item_embedding serves the target item AND the history) contributes one
 * group per use; TF concatenates their IndexedSlices in use order before the duplicate-row sum
 * (SURVEY A.1), which is exactly what passing several groups to one call does.
 */
typedef struct rb_lookup_group {
  const void* idx;                  /* [n] lookup ids of this use */
  int32_t idx_type;                 /* rb_index_type */
  int32_t L;                        /* positions per bag */
  int64_t n;                        /* number of lookup positions (bags * L) */
  const int64_t* field_row_offset;  /* optional [L] */
  int64_t hash_mod;                 /* 0 = off */
  rb_grad_source grad;
} rb_lookup_group;

/* ---- library ---------------------------------------------------------------------------- */

int rb_version(void);
const char* rb_last_error(void);
/* Kernels of this library launched by this process so far (cub's sort/scan passes are not counted). */
uint64_t rb_kernel_launches(void);
/* Programmatic dependent launch of this library's kernels (a kernel is placed on the SMs while the one before it in the stream
 * drains, and orders itself with griddepcontrol.wait): 1 = on, 0 = off, -1 = the RB_PDL environment variable decides (default
 * on).  Read at every launch; a captured CUDA graph keeps what was set when it was captured. */
void rb_set_pdl(int32_t mode);
/* alpha_t = lr*sqrt(1-b2^t)/(1-b1^t) in fp32, as Keras evaluates it. Host only. */
float rb_adam_alpha_t(float lr, float beta_1, float beta_2, int32_t step);

/* ---- K1: embedding lookup ---------------------------------------------------------------- */

/*
 * out[p,:] = table[row(p),:],  p in [0,n).   Replaces keras.layers.Embedding.__call__ /
 * ResourceGather at ctr/model.py:19, :49; dien/model.py:16-17; esmm/esmm.py:16.
 *   row(p) = idx[p] (+ field_row_offset[p % L] when field_row_offset != NULL; L = positions per
 *   sample, used for T>1 tables stored back to back).  hash_mod != 0 first folds
 *   idx <- uint64(idx) mod hash_mod (the build-defined id->row map).
 *   out row p is written at out + p*out_stride (elements); out_stride >= D.
 *   oob_flag (optional int32*): set to 1 if an index falls outside [0, rows); such lookups
 *   produce zeros (TF's GPU kernel behaviour).
 */
int rb_gather_fwd(const float* table, int64_t rows, int32_t D,
                  const void* idx, int32_t idx_type, int64_t n, int32_t L,
                  const int64_t* field_row_offset, int64_t hash_mod,
                  float* out, int64_t out_stride, int32_t* oob_flag, void* stream);

/*
 * Per-table id validation for T tables stored back to back: raises *oob_flag when an id (after hash_mod) of field f = p % L lies
 * outside [0, field_rows[f]).  The lookups themselves range-check the FINAL row against the total row count only, so such an
 * id would read and update the next table; keras.layers.Embedding rejects it per table on CPU (InvalidArgument,
 * ctr/model.py:19, :49).  One pass over the ids; n % L == 0; field_rows is a device array of L row counts.
 */
int rb_check_indices(const void* idx, int32_t idx_type, int64_t n, int32_t L, const int64_t* field_rows,
                     int64_t hash_mod, int32_t* oob_flag, void* stream);

/*
 * Pooled lookup: out[b, :D] = pool_l table[row(b,l),:]   (written at out + b*out_stride).
 * Replaces tf.reduce_sum(E,axis=1) (ctr/model.py:21) and compute_flat_embedding +
 * compute_his_average (dien/model.py:14-19,25-31; dien/layers.py:5-17).
 *   mask_idx: for RB_POOL_MASKED_MEAN the array whose != 0 is the mask (NULL = idx itself: the
 *   item-derived mask also gates the cat table, dien/model.py:25).  Masked positions are not read.
 *   count_out (optional f32[B]) receives the number of valid positions (L for sum/mean).
 *   No guard for count == 0: the result is NaN, as in the reference (dien/layers.py:16).
 */
int rb_bag_pool_fwd(const float* table, int64_t rows, int32_t D,
                    const void* idx, int32_t idx_type, int64_t B, int32_t L,
                    const int64_t* field_row_offset, int64_t hash_mod,
                    int32_t pool_mode, const void* mask_idx,
                    float* out, int64_t out_stride, float* count_out,
                    int32_t* oob_flag, void* stream);

/*
 * DeepFM front end, one pass over the rows (ctr/model.py:19-23):
 *   E[b,f,:] = table[idx[b,f],:]  (optional, may be NULL),
 *   s[b,:] = sum_f E[b,f,:] (optional),  fm[b] = 0.5*sum_d(s^2 - sum_f E^2).
 */
int rb_gather_fm_fwd(const float* table, int64_t rows, int32_t D,
                     const void* idx, int32_t idx_type, int64_t B, int32_t F,
                     const int64_t* field_row_offset, int64_t hash_mod,
                     float* E, float* s, float* fm, int32_t* oob_flag, void* stream);

/*
 * The same pass also writing DeepFM's deep input (ctr/model.py:25-26: tf.reshape(cat_embedding, (-1, F*D)) and the
 * tf.concat with int_features) as the bf16 K operand of the MLP's first Dense layer:
 *   deep[b, :] = [ E[b,0,:] ... E[b,F-1,:] | dense[b, 0..num_dense) | 1.0 | 0 ... ]   (ld_deep columns; the 1.0 only if a
 * column is left: it turns that layer's bias gradient into a row of its weight-gradient GEMM).  E, s may be NULL.
 */
int rb_gather_fm_deep_fwd(const float* table, int64_t rows, int32_t D,
                          const void* idx, int32_t idx_type, int64_t B, int32_t F,
                          const int64_t* field_row_offset, int64_t hash_mod,
                          const float* dense, int32_t num_dense, int64_t dense_ld, void* deep_bf16, int32_t ld_deep,
                          float* E, float* s, float* fm, int32_t* oob_flag, void* stream);

/* ---- K3..K6, K10: DotInteraction ------------------------------------------------------------ */

/*
 * DotInteraction.call (ctr/layers.py:23-43) on bf16 tensor cores, fp32 accumulate, with the
 * DLRM concats fused (ctr/model.py:51-55):
 *   X[b] = [ E[b,0..F-1,:] ; dense_vec[b,:] ]  (dense_vec optional -> F' = F or F+1 features)
 *   Z = X X^T;  kept set: self_interaction ? (j <= i) : (j > i)   (ctr/layers.py:27-33)
 *   skip_gather: out[b, i*F'+j] = kept ? Z[i,j] : 0   (F'^2 columns)        (ctr/layers.py:36-38)
 *   else       : kept entries in row-major order     (F'(F'-+1)/2 columns) (ctr/layers.py:39-42)
 *   tail != 0 : dense_vec[b,:] is also copied to out[b, ncols .. ncols+D)   (ctr/model.py:54)
 * E == NULL selects the fused gather: rows are read straight from `table` through idx[B,F]
 * (+field_row_offset), so [B,F,D] never exists in HBM.
 * out row b starts at out + b*out_stride (elements of out_dtype).  out_dtype = RB_F32 is the
 * reference layout (tf.float32, ctr/layers.py:38).  out_dtype = RB_BF16 rounds the row to bf16
 * for a bf16 top MLP and ALSO zero-fills the pad columns [ncols(+D), out_stride), so the padded
 * row can be the K operand of a GEMM (out_stride - width < 64).
 * Requires F' <= 32, D in {16, 32, 64, 128}.
 */
int rb_dot_interaction_fwd(const float* E, const float* table, int64_t rows,
                           const void* idx, int32_t idx_type, const int64_t* field_row_offset,
                           const float* dense_vec, int64_t B, int32_t F, int32_t D,
                           int32_t self_interaction, int32_t skip_gather, int32_t tail,
                           void* out, int32_t out_dtype, int64_t out_stride,
                           int32_t row_cache, const int32_t* row_cache_hint, void* stream);
/* row_cache (rb_row_cache) applies to the fused gather (E == NULL); row_cache_hint: the DEVICE flag of RB_ROW_CACHE_AUTO
 * (NULL reads as 0).  The choice changes only which cache level serves the row copies, never a result. */

/*
 * Backward of the above (TF autodiff of ctr/layers.py:25-42 and ctr/model.py:51-55):
 *   G = mask (.) dOut,  dX = (G + G^T) X;  dE[b,f,:] = dX[b,f,:] (f < F),
 *   d_dense[b,:] = dX[b,F,:] (+ dOut[b, ncols .. ncols+D) when tail != 0).
 * dE row (b,f) is written at dE + (b*F + f)*D.  dOut is RB_F32 or RB_BF16 (dout_dtype), row b at
 * dOut + b*dout_stride elements; dE and d_dense are always fp32.
 */
int rb_dot_interaction_bwd(const float* E, const float* table, int64_t rows,
                           const void* idx, int32_t idx_type, const int64_t* field_row_offset,
                           const float* dense_vec, int64_t B, int32_t F, int32_t D,
                           int32_t self_interaction, int32_t skip_gather, int32_t tail,
                           const void* dOut, int32_t dout_dtype, int64_t dout_stride,
                           float* dE, float* d_dense, int32_t row_cache, const int32_t* row_cache_hint, void* stream);

/* ---- K7..K9: backward scatter + sparse optimizer row update ------------------------------- */

size_t rb_sparse_bwd_update_workspace_bytes(int64_t n, int32_t D, int64_t rows);

/*
 * Gradient of the lookup applied straight to the table: IndexedSlices (A.1) -> duplicate-row
 * sum (A.2: Keras _deduplicate_indexed_slices = Unique + UnsortedSegmentSum) -> optimizer row
 * update (A.3/A.4: Adam._resource_apply_sparse at ctr/train.py:80,84,97).
 * Deterministic: (row, position) pairs are radix-sorted (stable), duplicate rows are summed in
 * position order by a segmented reduction, no floating-point atomics anywhere.
 *   state0/state1: Adam m, v; Adagrad acc (state1 unused); SGD: both unused.
 *   n = number of lookup positions, L = positions per bag (see rb_grad_source).
 *   RB_OPT_ADAM_TF_DENSE additionally makes the two dense passes over all `rows`.
 */
int rb_sparse_bwd_update(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                         const void* idx, int32_t idx_type, int64_t n, int32_t L,
                         const int64_t* field_row_offset, int64_t hash_mod,
                         const rb_grad_source* grad, const rb_opt_params* opt,
                         void* ws, size_t ws_bytes, int32_t* oob_flag, void* stream);

/* The general form: the concatenation of `num_groups` (1..RB_MAX_LOOKUP_GROUPS) uses of one table. */
int rb_sparse_bwd_update_groups(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                                const rb_lookup_group* groups, int32_t num_groups,
                                const rb_opt_params* opt,
                                void* ws, size_t ws_bytes, int32_t* oob_flag, void* stream);

/*
 * The two phases of rb_sparse_bwd_update_groups as separate calls.  Phase 1 (keys + stable radix sort)
 * depends only on the ids, which are known when the forward starts: run it on a side stream during
 * the forward / MLPs and the backward pays for phase 2 only.  `groups[k].grad` is ignored by prepare.
 * The workspace carries the sorted pairs from prepare to apply and must not be touched in between;
 * *sorted_sel (host) tells apply which half of the double buffers holds them.
 */
int rb_sparse_bwd_prepare(int64_t rows, int32_t D, const rb_lookup_group* groups, int32_t num_groups,
                          void* ws, size_t ws_bytes, int32_t* oob_flag, int32_t* sorted_sel, int32_t* hot_rows_flag,
                          void* stream);
/* hot_rows_flag (optional DEVICE int32[1]): set to 1 when rows that take >= 4096 lookups each hold more than 1/16 of the step's
 * lookups beyond their first 4095 (Zipf ids on a shared table, an OOV row), else 0 — the RB_ROW_CACHE_AUTO input of rb_dot_interaction_fwd / _bwd. */
int rb_sparse_bwd_apply(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                        const rb_lookup_group* groups, int32_t num_groups, const rb_opt_params* opt,
                        void* ws, size_t ws_bytes, int32_t sorted_sel, void* stream);

/*
 * Rows a step touches exactly once — fused into the backward.  With uniform Criteo-shaped ids (BASELINE config 2) 94 % of a
 * step's lookups hit a row no other lookup of the step hits: its summed gradient is the single gradient row, so the
 * IndexedSlices -> _deduplicate_indexed_slices -> _resource_apply_sparse chain (ctr/train.py:80,97) reduces to
 * "apply the optimizer to that row with that gradient".  The fused interaction backward does this where the gradient row is
 * produced, while the table row is still in shared memory: the gradient row never goes to HBM and back (2 x N x D x 4 bytes)
 * and the table row is not read a second time.  Results are bit-identical to the unfused chain.  Protocol per step:
 *   rb_sparse_bwd_prepare            (side stream, as before)
 *   rb_sparse_bwd_mark_singletons    single[p] = 1 iff the row of lookup position p occurs once among the sorted pairs;
 *                                    the pairs of the other rows are compacted for the reduction below
 *   rb_dot_interaction_bwd_update    = rb_dot_interaction_bwd; positions with single[p] != 0 get their optimizer row update
 *                                    instead of a dE row (their dE rows are left unwritten)
 *   rb_sparse_bwd_apply_ex(flags = RB_APPLY_SKIP_SINGLETONS)   the sorted reduction over the remaining (duplicate) rows
 * One lookup group per table and step, plain fp32 gradient rows, D % 4 == 0; SGD / Adagrad / lazy Adam.
 */
#define RB_APPLY_SKIP_SINGLETONS 1
int rb_sparse_bwd_mark_singletons(int64_t rows, int32_t D, int64_t n, void* ws, size_t ws_bytes, int32_t* sorted_sel,
                                  uint8_t* single, void* stream);
/* ... and compacts the sorted pairs of all OTHER rows (order kept) into the free half of the workspace's double buffers:
 * *sorted_sel (host, in/out) is flipped to that half and the survivors' count stays on the device; hand the new selector to
 * rb_sparse_bwd_apply_ex together with RB_APPLY_SKIP_SINGLETONS. */
int rb_sparse_bwd_apply_ex(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                           const rb_lookup_group* groups, int32_t num_groups, const rb_opt_params* opt,
                           void* ws, size_t ws_bytes, int32_t sorted_sel, int32_t flags, void* stream);
/* `table` must be the fp32 table itself (E == NULL form only, no hash_mod); state0 / state1 as in rb_sparse_bwd_update. */
int rb_dot_interaction_bwd_update(float* table, int64_t rows, const void* idx, int32_t idx_type,
                                  const int64_t* field_row_offset, const float* dense_vec, int64_t B, int32_t F, int32_t D,
                                  int32_t self_interaction, int32_t skip_gather, int32_t tail,
                                  const void* dOut, int32_t dout_dtype, int64_t dout_stride, float* dE, float* d_dense,
                                  const uint8_t* single, float* state0, float* state1, const rb_opt_params* opt,
                                  int32_t row_cache, const int32_t* row_cache_hint, void* stream);

/*
 * Update of tables REPLICATED on every rank of a sharded job from `num_lists` (one per rank, all-gathered) sorted
 * lists of (row, summed gradient): list k = cap records of D + 1 floats — D gradient values, then the row id as int32 bits —
 * rows ascending, padded with ids >= rows.  A row held by any list gets ONE optimizer update with the lists' gradients added
 * in list order (the IndexedSlices of all replicas concatenated, SURVEY A.7, then A.2-A.3); other rows do not move.
 * shadow_bf16 (optional): bf16 copy of the table kept in step.
 */
int rb_replicated_rows_update(float* table, float* state0, float* state1, int64_t rows, int32_t D, const float* lists,
                              int32_t num_lists, int64_t cap, const rb_opt_params* opt, void* shadow_bf16, void* stream);

/*
 * The same sort + segmented reduction WITHOUT the optimizer: writes the unique rows (ascending)
 * and their summed gradients — the deduplicated IndexedSlices itself.  Used by the parity tests
 * and by callers that own their optimizer.
 *   uniq_rows int64[n], uniq_grad f32[n,D], num_unique int64[1] (device).
 */
int rb_sparse_bwd_dedup(int64_t rows, int32_t D,
                        const void* idx, int32_t idx_type, int64_t n, int32_t L,
                        const int64_t* field_row_offset, int64_t hash_mod,
                        const rb_grad_source* grad,
                        int64_t* uniq_rows, float* uniq_grad, int64_t* num_unique,
                        void* ws, size_t ws_bytes, int32_t* oob_flag, void* stream);

/* ---- dense side of the step ------------------------------------------------- */

/* One dense parameter tensor of the MLPs (ctr/layers.py:5-14) with its gradient and optimizer state. */
typedef struct rb_dense_slot {
  float* param;        /* f32[n], updated in place */
  float* state0;       /* Adam m / Adagrad accumulator / NULL for SGD */
  float* state1;       /* Adam v / NULL */
  const float* grad;   /* f32[n] */
  int64_t n;
  void* shadow_bf16;   /* optional bf16[n]: receives the updated parameter rounded to bf16 (the operand of bf16 GEMMs) */
} rb_dense_slot;

/*
 * Keras `_resource_apply_dense` (Adam: SURVEY A.3, the same formula as the sparse rows; Adagrad: A.4;
 * SGD) for up to RB_MAX_DENSE_TENSORS tensors in ONE launch — replaces the per-variable
 * ResourceApplyAdam ops of `optimizer.apply_gradients` (ctr/train.py:80,97; dien/train.py:22).
 * `slots` is a HOST array.  RB_OPT_ADAM_LAZY and RB_OPT_ADAM_TF_DENSE are the same thing for dense tensors.
 */
int rb_dense_opt_step(const rb_dense_slot* slots, int32_t num, const rb_opt_params* opt, void* stream);

size_t rb_colsum_workspace_bytes(int64_t rows, int32_t cols);

/*
 * out[c] = sum_r x[r, c]: the bias gradient of a Dense layer (TF BiasAddGrad under ctr/layers.py:8-9).
 * x is RB_F32 or RB_BF16 [rows, cols] with row stride `row_stride` elements; cols % 8 == 0, cols <= 2048.
 * Deterministic two-stage reduction, fp32 accumulation.
 */
int rb_colsum(const void* x, int32_t dtype, int64_t rows, int32_t cols, int64_t row_stride, float* out,
              void* ws, size_t ws_bytes, void* stream);

size_t rb_bce_workspace_bytes(int64_t n);

/*
 * Loss head of ctr/train.py:85-87 for DLRM: Keras binary_crossentropy on PROBABILITIES (clipped form, SURVEY
 * Appendix A.5), batch mean, and d loss / d prob in the same pass (dprob_out optional).  prob f32[n];
 * label f32[n] (label_type 0) or int64[n] (label_type 1).  Deterministic two-stage mean.
 */
int rb_bce_clipped(const float* prob, const void* label, int32_t label_type, int64_t n, float* loss_out, float* dprob_out,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- Dense layers of the towers (ctr/layers.py:5-14) on tcgen05 tensor cores ------------------
 *
 * `MLP.call` (ctr/layers.py:11-14) is a chain of keras.layers.Dense; DLRM builds 13 -> 512 -> 256 -> 64 (relu) and
 * 793 -> 512 -> 256 -> 1 (sigmoid), DeepFM 429 -> 512 -> 256 -> 1 (ctr/train.py:74-75,82; ctr/model.py:27,50,56).  These
 * entry points are the three products of one Dense layer and its Dense(1) head, replacing TF's MatMul / BiasAdd /
 * Relu / Sigmoid ops and their gradients.  Operands are bf16 (fp32 accumulation in tensor memory); every matrix is
 * row-major with a leading dimension in ELEMENTS; base pointers and row strides must be 16-byte aligned; `in_dim` and
 * `units` must be multiples of 8 (the callers zero-pad the feature axis: 13 -> 16, 793 -> 800).
 */
typedef enum rb_activation { RB_ACT_NONE = 0, RB_ACT_RELU = 1, RB_ACT_SIGMOID = 2 } rb_activation;

/* y = activation(x . W + bias)   — Dense.call (ctr/layers.py:8-9,12-13).
 * x bf16 [rows, in_dim] (ldx), w bf16 [in_dim, units] (ldw), bias f32 [units] or NULL, y RB_BF16 or RB_F32 [rows, units] (ldy). */
int rb_dense_fwd(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* w, int32_t units, int64_t ldw,
                 const float* bias, int32_t activation, void* y, int32_t y_type, int64_t ldy, void* stream);

/* dx = dy . W^T   — the MatMul gradient w.r.t. the layer input.  dy bf16 [rows, units], w bf16 [in_dim, units], dx bf16 [rows, in_dim]. */
int rb_dense_bwd_input(const void* dy, int64_t rows, int32_t units, int64_t lddy, const void* w, int32_t in_dim, int64_t ldw,
                       void* dx, int64_t lddx, void* stream);

size_t rb_dense_bwd_weight_workspace_bytes(int64_t rows, int32_t in_dim, int32_t units);

/* dW = x^T . dy (f32 [in_dim, units], lddw)   — the MatMul gradient w.r.t. the kernel.  The batch axis is split over the
 * SMs; partial tiles are summed in split order (deterministic).  When x carries a ones column (rb_dense_pack_input,
 * RB_BF16_ONES rows of rb_dot_interaction_fwd) the matching row of dW is the bias gradient (BiasAddGrad). */
int rb_dense_bwd_weight(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* dy, int32_t units, int64_t lddy,
                        float* dw, int64_t lddw, void* ws, size_t ws_bytes, void* stream);

/* Dense(1) head: out[r] = activation(sum_k x[r,k] * w[k] + bias[0])   (the last unit of [512, 256, 1]).  x, w bf16; out f32 [rows]. */
int rb_dense_head_fwd(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* w, const float* bias, int32_t activation,
                      float* out, void* stream);

size_t rb_dense_head_bwd_workspace_bytes(int64_t rows, int32_t in_dim);

/* Backward of the head: dz = dout * activation'(out);  dx[r,:] = bf16(dz[r] * w[:]) (optional);  dw[k] = sum_r x[r,k] * dz[r];
 * db[0] = sum_r dz[r];  dx_colsum[k] = sum_r dx[r,k] (optional, f32 [in_dim]: the bias gradient of the layer below, which
 * would otherwise re-read dx).  Deterministic two-stage sums. */
int rb_dense_head_bwd(const float* dout, const float* out, int32_t activation, const void* x, int64_t rows, int32_t in_dim, int64_t ldx,
                      const void* w, void* dx, int64_t lddx, float* dw, float* db, float* dx_colsum, void* ws, size_t ws_bytes,
                      void* stream);

/* Dense(1) + sigmoid, the clipped binary cross-entropy of its output (rb_bce_clipped's formulas, batch mean) and the head's
 * backward seeded with d loss = 1, in ONE pass over x: prob f32[rows], loss f32[1], dx bf16 [rows, in_dim] (optional), dw f32[in_dim],
 * db f32[1], dx_colsum (optional) — bit-identical to rb_dense_head_fwd -> rb_bce_clipped -> rb_dense_head_bwd except for the order
 * the loss terms are added.  ctr/model.py:56-57 + ctr/train.py:85-87,97 for models ending in Dense(1, sigmoid).
 * in_dim in {8, 16, 32, 64, 128, 256}; label_type: 0 = f32, 1 = i64. */
size_t rb_dense_head_bce_workspace_bytes(int64_t rows, int32_t in_dim);
int rb_dense_head_bce(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* w, const float* bias, const void* label,
                      int32_t label_type, float* prob, float* loss, void* dx, int64_t lddx, float* dw, float* db, float* dx_colsum,
                      void* ws, size_t ws_bytes, void* stream);

/* out_bf16[i] = bf16(dy[i] * activation'(y[i])) over n contiguous f32 elements: the ReluGrad / SigmoidGrad in front of the
 * last layer's gradient GEMMs. */
int rb_dense_act_bwd(const float* dy, const float* y, int32_t activation, int64_t n, void* out_bf16, void* stream);
/* the same with bf16 dy and y: the activation of a HIDDEN layer (dien/layers.py:37-38: Dense(80, sigmoid), Dense(40, sigmoid)),
 * whose output the forward kept in bf16.  dz = bf16(f32(dy) * act'(f32(y))). */
int rb_dense_act_bwd_bf16(const void* dy_bf16, const void* y_bf16, int32_t activation, int64_t n, void* out_bf16, void* stream);

/* out_bf16[rows, ld_out] = [x (f32 [rows, in_dim], ldx) | 1.0 if ones_col | 0 ...]: the padded bf16 K operand of a first Dense layer. */
int rb_dense_pack_input(const float* x, int64_t rows, int32_t in_dim, int64_t ldx, void* out_bf16, int32_t ld_out, int32_t ones_col,
                        void* stream);

/* Diagnostic: when `device_buf` (u64 [8 * 148], zeroed by the caller) is non-NULL, every following Dense GEMM launch of this
 * process records per CTA the cycles its roles spent waiting: [0] MMA issuer on operand stages, [1] MMA issuer on a drained
 * accumulator, [2] MMA issuer total, [3] TMA producer on free stages, [4] producer total, [5] epilogue on a finished
 * accumulator, [6] epilogue total.  NULL switches it off.  Tuning aid (scripts/mlp_check.py --stats); not thread-safe. */
int rb_dense_debug_stats(void* device_buf);

/* ---- id -> row map ------------------------------------------------------------------------ */

/* rows_out[p] = uint64(ids[p]) mod vocab  (SURVEY §8c "index hashing"; bit-exact with the oracle).
 * Row-wise sharding: owner = row mod world, local = row div world (world <= 1: no sharding). */
int rb_hash_ids(const void* ids, int32_t idx_type, int64_t n, int64_t vocab, int32_t world,
                int64_t* rows_out, int32_t* owner_out, int64_t* local_out, void* stream);

/* ---- row-wise sharding: route lookups to their owner rank -------------------------------------- */

size_t rb_bucket_by_owner_workspace_bytes(int64_t n, int32_t world);

/*
 * Stable partition of n lookups by owner = row mod world (row-wise sharding), the
 * send side of the index all-to-all.  row(p) as in rb_gather_fwd (hash_mod, field_row_offset).
 *   local_rows_out int64[n]: row div world, in bucket order (rank 0's lookups first, ...)
 *   perm_out       int32[n]: bucket slot -> original lookup position p
 *   inv_perm_out   int32[n]: original position p -> bucket slot (the index array with which the
 *                            interaction / gather kernels read the returned rows in place)
 *   counts_out     int64[world]: lookups per owner
 */
/* rb_bucket_by_owner that leaves lookups whose row is >= skip_from_row out of every bucket (rows of replicated tables):
 * not counted, no slot, inv_perm_out[p] = -1.  world <= 8. */
int rb_bucket_by_owner_skip(const void* idx, int32_t idx_type, int64_t n, int32_t L, const int64_t* field_row_offset,
                            int64_t hash_mod, int32_t world, int64_t skip_from_row, int64_t* local_rows_out, int32_t* perm_out,
                            int32_t* inv_perm_out, int64_t* counts_out, void* ws, size_t ws_bytes, void* stream);
int rb_bucket_by_owner(const void* idx, int32_t idx_type, int64_t n, int32_t L,
                       const int64_t* field_row_offset, int64_t hash_mod, int32_t world,
                       int64_t* local_rows_out, int32_t* perm_out, int32_t* inv_perm_out,
                       int64_t* counts_out, void* ws, size_t ws_bytes, void* stream);

/* ---- sharded tables over peer memory --------------------------------------------------- */
/*
 * One process per GPU; the reference only replicates (MirroredStrategy, ctr/train.py:71-84).  Here the
 * table is split row-wise (row r -> rank r mod world, local row r div world) and the other GPUs'
 * shards are read IN the kernels over NVLink/NVSwitch peer pointers — the gather is the collective:
 * no all-to-all of rows or gradients, no staging copies.
 *
 * rb_shared_alloc: cudaMalloc + cudaIpcGetMemHandle (handle_out: RB_IPC_HANDLE_BYTES bytes to ship to the
 * peers, e.g. with torch.distributed.all_gather_object).  rb_ipc_open maps a peer's buffer into this process.
 */
int rb_shared_alloc(size_t bytes, void** ptr_out, void* handle_out);
int rb_shared_free(void* ptr);
int rb_ipc_open(const void* handle, void** ptr_out);
int rb_ipc_close(void* ptr);
/* single process driving several GPUs: let kernels on the current device dereference `peer_device` memory */
int rb_enable_peer_access(int32_t peer_device);

/* rb_dot_interaction_fwd / _bwd with the rows fetched from `world` shards in peer memory.
 * shard_ptrs_dev: DEVICE array of `world` device pointers (own shard included); rows: GLOBAL row count.
 * x_save (optional, bf16 [B, F', D]): the forward also stores the operand rows rounded to bf16 — exactly
 * what its MMAs consume — and the backward given the same buffer as x_saved reads them locally instead
 * of gathering the rows over NVLink a second time (bit-identical results).
 * shadow_ptrs_dev (optional, DEVICE array of `world` bf16 shard pointers): bf16 copies of the shards kept in
 * step by rb_sparse_bwd_apply_p2p(shadow_bf16); the forward then moves 2*D bytes per row over NVLink
 * instead of 4*D, and since the MMA operands are the rows rounded to bf16 either way, results are unchanged. */
int rb_dot_interaction_fwd_sharded(const void* const* shard_ptrs_dev, int32_t world, int64_t rows,
                                   const void* idx, int32_t idx_type, const int64_t* field_row_offset,
                                   const float* dense_vec, int64_t B, int32_t F, int32_t D,
                                   int32_t self_interaction, int32_t skip_gather, int32_t tail,
                                   void* out, int32_t out_dtype, int64_t out_stride, void* x_save,
                                   const void* const* shadow_ptrs_dev, void* stream);
/* ... with tables REPLICATED on every rank behind the sharded ones: rows in [small_base, rows) are read from this rank's own copy
 * (small_rep f32 / small_rep_bf16 [rows - small_base, D], the bf16 one when shadow shards are given) instead of a shard.  The
 * Criteo-Terabyte cardinalities hold 11 tables of <= 2208 rows that take 42 % of all lookups: sharding them row-wise sends every
 * one of those lookups to the same handful of owners. */
int rb_dot_interaction_fwd_sharded_rep(const void* const* shard_ptrs_dev, int32_t world, int64_t rows,
                                       const void* idx, int32_t idx_type, const int64_t* field_row_offset,
                                       const float* dense_vec, int64_t B, int32_t F, int32_t D,
                                       int32_t self_interaction, int32_t skip_gather, int32_t tail,
                                       void* out, int32_t out_dtype, int64_t out_stride, void* x_save,
                                       const void* const* shadow_ptrs_dev, int64_t small_base, const float* small_rep,
                                       const void* small_rep_bf16, void* stream);
int rb_dot_interaction_bwd_sharded(const void* const* shard_ptrs_dev, int32_t world, int64_t rows,
                                   const void* idx, int32_t idx_type, const int64_t* field_row_offset,
                                   const float* dense_vec, int64_t B, int32_t F, int32_t D,
                                   int32_t self_interaction, int32_t skip_gather, int32_t tail,
                                   const void* dOut, int32_t dout_dtype, int64_t dout_stride,
                                   float* dE, float* d_dense, const void* x_saved, void* stream);

/* rb_dot_interaction_bwd_sharded that writes the gradient rows of the fields with de_slot[f] >= 0 (DEVICE int32[F]; the fields of
 * replicated tables) to the compact tensor dE_small[B, num_small, D] at column de_slot[f] instead of dE: their duplicate-row sum
 * (rb_sparse_bwd_dedup) then reads one contiguous tensor.  Needs the saved operand rows (x_saved). */
int rb_dot_interaction_bwd_sharded_split(const void* const* shard_ptrs_dev, int32_t world, int64_t rows,
                                         const void* idx, int32_t idx_type, const int64_t* field_row_offset,
                                         const float* dense_vec, int64_t B, int32_t F, int32_t D,
                                         int32_t self_interaction, int32_t skip_gather, int32_t tail,
                                         const void* dOut, int32_t dout_dtype, int64_t dout_stride,
                                         float* dE, float* d_dense, const void* x_saved, const int32_t* de_slot, float* dE_small,
                                         int32_t num_small, void* stream);

/*
 * Owner side of the sharded backward.  Every rank k publishes its rb_bucket_by_owner outputs (local
 * rows, perm, counts) in peer memory; owner `me` collects the slices addressed to it as
 * (key = local row, value = k*n_local + position) pairs into the sparse-backward workspace, padded
 * to the static `capacity` with a key that sorts last (static shapes: the step can be a CUDA graph).
 * *n_valid_dev receives the real pair count, *overflow_flag is raised if it exceeds capacity.
 * The *_ptrs arrays are HOST arrays of `world` device (peer) pointers.  ws: rb_sparse_bwd_update_workspace_bytes(
 * capacity, D, local_rows + 1).
 */
int rb_p2p_collect_keys(int32_t world, int32_t me, int64_t n_local, const void* const* rows_ptrs,
                        const void* const* perm_ptrs, const void* const* counts_ptrs, int64_t local_rows,
                        int64_t capacity, void* ws, size_t ws_bytes, int32_t D, int32_t* n_valid_dev,
                        int32_t* overflow_flag, void* stream);
/* stable radix sort of the collected pairs (the phase-1 counterpart of rb_sparse_bwd_prepare) */
int rb_sparse_bwd_prepare_collected(int64_t local_rows, int32_t D, int64_t capacity, void* ws, size_t ws_bytes,
                                    int32_t* sorted_sel, void* stream);
/* phase 2: segmented reduction + optimizer row update on the local shard; gradient rows are pulled from
 * the ranks' dE[B_local, L, D] tensors in peer memory (dE_ptrs: HOST array of `world` device pointers). */
int rb_sparse_bwd_apply_p2p(float* table, float* state0, float* state1, int64_t local_rows, int32_t D, int32_t world,
                            int64_t n_local, int32_t L, const void* const* dE_ptrs, int64_t capacity,
                            const int32_t* n_valid_dev, const rb_opt_params* opt, void* ws, size_t ws_bytes,
                            int32_t sorted_sel, void* shadow_bf16, void* stream);

/* ---- input side: Criteo TSV -> batch ------------------------------------------------ */
/*
 * Replaces the per-line Python of ctr/tfrecord_io.py: build_vocab (:15-35), the body of write_tfrecord (:43-66)
 * and what read_tfrecord (:78-96) hands to the model — ({'int_features' f32[13], 'cat_features' int64[26]}, label).
 * The TFRecord container itself is not reproduced: the text goes from HBM straight to the batch.
 *
 * text: the file (or a chunk of whole lines, < 2 GiB) resident in HBM, 16-byte aligned, its allocation readable up to
 * the next multiple of 16 bytes.  One line = label \t I1..I13 \t C1..C26, '\n'-terminated (the last line may lack it).
 *
 * Dictionary keys.  The reference keys ONE dictionary for all 26 columns by Python strings (:24-33); this library keys
 * by the token's bytes packed little-endian into a uint64 — the same identity for tokens of 1..8 ASCII bytes (Criteo's
 * are 8 hex digits).  Two quirks are kept: (1) str.split leaves the line's newline attached to C26, so its tokens are
 * different keys from the same bytes in another column — carried as bit 63; (2) an empty column ('' or '\n', :21/:55)
 * is imputed with a per-column token — key (column + 1) << 8, whose low byte no real token has.
 *
 * error_flag (optional int32*, OR-ed): 1 = a line with fewer than 40 columns (the reference raises IndexError),
 * 2 = a label / integer column int() would reject or of more than 18 digits (int() accepts blanks and underscores;
 * this parser takes an optional sign and digits), 4 = a categorical token longer than 8 bytes, 8 = a non-ASCII byte
 * in a token, 16 = a line longer than 1024 bytes.  Offending values are written as 0 / truncated; outputs of other
 * lines are unaffected.
 */
size_t rb_criteo_index_workspace_bytes(int64_t nbytes);
/* line_start[i] = byte offset of line i (i < min(*num_lines_dev, max_lines)); `for line in f` semantics: a final
 * newline does not open an empty last line.  *num_lines_dev is the true count even when it exceeds max_lines. */
int rb_criteo_index_lines(const uint8_t* text, int64_t nbytes, int64_t max_lines, int64_t* line_start,
                          int64_t* num_lines_dev, void* ws, size_t ws_bytes, void* stream);
/*
 * One warp per line.  label int64[n] (:66), int_features f32[n,13] = log(float32(max(x, 0)) + 1), '' -> 0 (:45-53),
 * cat_tokens (optional) uint64[n,26] the dictionary keys in scan order (input of rb_vocab_build),
 * cat_features (optional) int64[n,26] = vocabulary id, 0 when absent (:58-65) — needs the table of
 * rb_vocab_table_build.  num_lines is a HOST value (read *num_lines_dev back, or know the chunk).
 */
int rb_criteo_parse(const uint8_t* text, int64_t nbytes, const int64_t* line_start, int64_t num_lines,
                    int64_t* label, float* int_features, uint64_t* cat_tokens, int64_t* cat_features,
                    const uint64_t* vocab_table_keys, const int32_t* vocab_table_vals, int64_t vocab_capacity,
                    int32_t* error_flag, void* stream);
/*
 * build_vocab (:15-35) over n tokens in scan order (line-major, column-minor): a token is kept when it occurs more
 * than min_count times (10 in the reference) and ids are dense from 0 in first-seen order, which is what iterating
 * the reference's insertion-ordered dict yields.  vocab_keys_out[id] = key for id < min(*num_vocab_dev, max_vocab);
 * *num_vocab_dev is the true size.  Stream-ordered, no host round trip.  n < 2^31 tokens per call.
 */
size_t rb_vocab_build_workspace_bytes(int64_t n);
int rb_vocab_build(const uint64_t* tokens, int64_t n, int32_t min_count, uint64_t* vocab_keys_out, int64_t max_vocab,
                   int64_t* num_vocab_dev, void* ws, size_t ws_bytes, void* stream);
/* key -> id table in HBM: open addressing, linear probing, slot = splitmix64(key) & (capacity - 1);
 * capacity: a power of two > num_vocab (2x or more keeps probes short).  vocab_keys must be distinct. */
int rb_vocab_table_build(const uint64_t* vocab_keys, int64_t num_vocab, uint64_t* table_keys, int32_t* table_vals,
                         int64_t capacity, void* stream);
/* ids_out[i] = id of tokens[i], 0 when absent (:61-64: OOV shares id 0 with the first vocabulary entry) */
int rb_vocab_lookup(const uint64_t* tokens, int64_t n, const uint64_t* table_keys, const int32_t* table_vals,
                    int64_t capacity, int64_t* ids_out, void* stream);

/* ---- DIN attention pooling ----------------------------------------------------------------
 * dien/layers.py:34-59 LocalActivationUnit at its call site dien/model.py:42-53: the history rows are the concatenation
 * of an item-table row and a category-table row (compute_flat_embedding, dien/model.py:14-19), mask = item id != 0
 * (keras mask_zero, dien/model.py:11,43).  Only VALID positions get a feature row; they are numbered
 * p = offsets[b] + (rank of l among sample b's valid positions).  The three Dense layers between rb_din_build_features
 * and rb_din_pool_fwd run on rb_dense_fwd / rb_dense_head_fwd (sigmoid, sigmoid, none), their backward on
 * rb_dense_head_bwd / rb_dense_act_bwd_bf16 / rb_dense_bwd_input / rb_dense_bwd_weight. */
typedef struct rb_din_history {
  const float* table0;  int64_t rows0;  int32_t D0;  const void* idx0;   /* item table, ids [B, L] */
  const float* table1;  int64_t rows1;  int32_t D1;  const void* idx1;   /* category table or NULL / 0 / 0 / NULL */
  int32_t idx_type;                                                     /* RB_I32 | RB_I64, both id arrays */
  const void* mask;     int32_t mask_type;                              /* [B, L], != 0 = valid; NULL: idx0 != 0 */
  int64_t B;            int32_t L;
} rb_din_history;

size_t rb_din_workspace_bytes(int64_t B);
/* offsets int32[B + 1] (device): exclusive scan of the per-sample valid counts; offsets[B] = P, the number of feature rows */
int rb_din_offsets(const rb_din_history* h, int32_t* offsets, void* ws, size_t ws_bytes, void* stream);
/* x[p, :] = bf16([t_b | h_p | t_b - h_p | t_b * h_p | 0...]) for every valid position (dien/layers.py:48-49); target f32[B, E],
 * E = D0 + D1; x bf16 [P, ldx], 4E <= ldx < 4E + 32 */
int rb_din_build_features(const rb_din_history* h, const float* target, const int32_t* offsets, void* x_bf16, int64_t ldx, void* stream);
/* rep[b, :] = sum over sample b's valid positions, in order of l, of w[p] * h_p   (dien/layers.py:53-57) */
int rb_din_pool_fwd(const rb_din_history* h, const int32_t* offsets, const float* w, float* rep, void* stream);
/* dw[p] = <d_rep[b], h_p> */
int rb_din_pool_bwd_weights(const rb_din_history* h, const int32_t* offsets, const float* d_rep, float* dw, void* stream);
/* dh[b, l, :] (f32 [B, L, E], written for valid positions; masked ones are zero-filled when zero_masked != 0, else left
 * untouched: read them through RB_SCALE_MASKED) = dx[p, E:2E] - dx[p, 2E:3E] + dx[p, 3E:4E] * t_b + w[p] * d_rep[b];
 * d_target[b, :] = sum_p dx[p, 0:E] + dx[p, 2E:3E] + dx[p, 3E:4E] * h_p */
int rb_din_feature_bwd(const rb_din_history* h, const float* target, const int32_t* offsets, const void* dx_bf16, int64_t ldx,
                       const float* w, const float* d_rep, float* dh, float* d_target, int32_t zero_masked, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* RECSYS_B200_H_ */
