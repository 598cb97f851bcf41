// K3..K6 / K10: DotInteraction forward and backward on bf16 tensor cores (sm_100a).
//
// HBM-bound (9 FLOP/B against a ridge of ~214), so the kernels are built around the
// memory pipeline, not the MMA:
//   * persistent CTAs (one per SM), every warp owns a private 2-stage ring in shared memory and
//     walks its samples independently — no CTA-wide barrier anywhere;
//   * the F' <= 32 feature rows of the NEXT sample (F embedding rows read straight from the
//     table through idx in the fused-gather form, or from E; plus the optional dense vector)
//     are fetched with 16-byte cp.async (LDGSTS, L2-only) while the current sample is computed;
//     the row ids of the sample after that are already in flight (one coalesced load, lane f
//     owns field f), so no dependent global load sits on the critical path;
//   * MMA fragments are built directly from the fp32 rows in shared memory (conflict-free
//     padded stride, cvt.rn.bf16x2 in registers): Z = X X^T needs no separate B operand (the B
//     fragment of n-tile t is half of the A fragment of m-tile t/2), dX = (G+G^T) X takes its A
//     operand straight from the staged dOut row through per-lane offsets computed once;
//   * the triangular mask, zero fill / compaction, the DLRM concat of the dense vector and the
//     "|| bmlp" tail are applied in the epilogue; the output row is staged
//     in shared memory and written with aligned 16-byte stores.  fp32 [B, ncols(+D)] is the
//     reference layout; the bf16 form (row padded to out_stride with zeros) is what the top
//     MLP's first GEMM consumes, and the backward accepts dOut in the same two forms.
// The per-sample problem (27x27x64) is far below a tcgen05 128-row tile; m16n8k16 bf16 MMAs keep
// the tensor work at a few percent of the kernel.
#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace rb {

#ifndef RB_PEER_CA
#define RB_PEER_CA 1
#endif
#ifndef RB_IX_STAGES
#define RB_IX_STAGES 2
#endif
#ifndef RB_IX_WARPS
#define RB_IX_WARPS 12
#endif
#ifndef RB_IX_FWD_PAD
#define RB_IX_FWD_PAD 8
#endif
constexpr int kIxStages = RB_IX_STAGES;   // ring depth per warp: samples in flight = kIxStages - 1
// upper bound of warps per CTA (each with its own ring); the launch picks as many as fit in 227 KiB of shared memory
template <int D>
struct IxWarps {
  static constexpr int value = RB_IX_WARPS;
};
constexpr size_t kIxSmemLimit = 227 * 1024 - 1024;
static int warps_that_fit(size_t per_warp_bytes) {
  int w = static_cast<int>(kIxSmemLimit / per_warp_bytes);
  return w > RB_IX_WARPS ? RB_IX_WARPS : w;
}

struct IxArgs {
  const float* E;          // [B,F,D] or null (fused gather)
  const float* table;      // used when E == null
  const float* const* shards;  // row-wise sharded table: device array of `world` shard pointers (peer memory), or null
  int world;               // row r lives in shards[r % world] at local row r / world
  __nv_bfloat16* x_save;        // forward, optional: X[B,F',D] rounded to bf16 (exactly the MMA operands) for the backward
  const __nv_bfloat16* x_load;  // backward, optional: read X from there instead of gathering the rows again
  IndexMap map;            // idx[B,F]
  const float* dense_vec;  // [B,D] or null
  int64_t B;
  int F;                   // embedding features
  int Fp;                  // F + (dense_vec != null)
  int self_interaction;
  int skip_gather;
  int tail;
  int ncols;               // interaction columns (without tail)
  int pad_one;             // bf16 output rows: first pad column = 1.0 (RB_BF16_ONES)
  const int32_t* de_slot;  // backward, optional [F]: de_slot[f] >= 0 sends the gradient rows of field f to the compact tensor
  float* dE_small;         //   dE_small[B, num_small, D] at column de_slot[f] instead of dE[B, F, D] (fields of replicated tables)
  int num_small;
  uint32_t small_base;     // sharded forms: rows >= small_base belong to tables REPLICATED on every rank and are read from the
  const void* small_rep;   //   local replica small_rep (row - small_base; fp32 rows, or bf16 rows in the shadow form) instead of a shard
  int row_cache;           // rb_row_cache: copy table rows through L1 (hot rows: Zipf ids, the OOV row) or past it (uniform ids)
  const int32_t* row_cache_hint;   // RB_ROW_CACHE_AUTO: device flag written by rb_sparse_bwd_prepare of this / the previous step
};

// Fused optimizer update of the rows a step touches once (rb_dot_interaction_bwd_update): the backward kernel applies the
// optimizer to such a row where its gradient row is produced instead of writing the gradient to dE.
struct IxUpdate {
  const uint8_t* single;   // [B*F]: 1 = the row of this lookup position occurs once among the step's lookups
  float* table;            // the table, writable
  float* s0;               // optimizer state rows (Adam m / Adagrad accumulator) or null
  float* s1;               // Adam v or null
  OptMath math;
  const float* alpha_dev;  // optional device alpha_t (graph replays)
};

// Hot rows.  Under Zipf ids on ONE shared table (the reference's layout, ctr/model.py:42) a tenth of all lookups read the
// same row: through L2 only (cp.async.cg) every SM queues on the one L2 slice that holds the line and the forward runs 3.3x
// slower than with uniform ids (r2_00: 376 us against 132 us); through L1 (cp.async.ca) each SM keeps the handful of hot
// rows itself (r2_10: 105 us).  With uniform ids the L1 pass costs ~20 % (107 -> 130 us): nothing is re-used and every
// row allocates two lines.  So the choice is made per step from the ids themselves (RB_ROW_CACHE_AUTO).
__device__ __forceinline__ bool rows_through_l1(const IxArgs& a) {
  if (a.row_cache == RB_ROW_CACHE_AUTO) return a.row_cache_hint != nullptr && __ldg(a.row_cache_hint) != 0;
  return a.row_cache == RB_ROW_CACHE_L1;
}

// ---- small PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return r;
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// kept(i,j) and its column in the output row   (ctr/layers.py:27-42)
__device__ __forceinline__ int compact_pos(int i, int j, int Fp, int self_interaction) {
  return self_interaction ? (i * (i + 1) / 2 + j) : (i * Fp - i * (i + 1) / 2 + (j - i - 1));
}
__device__ __forceinline__ int out_pos(int i, int j, const IxArgs& a) {
  return a.skip_gather ? (i * a.Fp + j) : compact_pos(i, j, a.Fp, a.self_interaction);
}

// ---- output / dOut element types ---------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float from(float v) { return v; }
  static __device__ __forceinline__ float to_f32(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ __nv_bfloat16 from(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
};

// ---- the row ids of one sample: lane f holds the source row of field f -----------------------------
// Fused gather: row = idx[b,f] (+ field offset), kInvalidRow when out of range (TF's GPU gather
// writes zeros).  Materialised E: row = b*F + f of E viewed as [B*F, D].
constexpr uint32_t kInvalidRow = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t load_sample_row(const IxArgs& a, int64_t b, int lane, int64_t lane_off) {
  if (lane >= a.F || b >= a.B) return kInvalidRow;
  if (a.E != nullptr) return static_cast<uint32_t>(b * a.F + lane);
  const int64_t id = load_raw_index(a.map.idx, a.map.is64, b * a.F + lane) + lane_off;
  return (id >= 0 && id < a.map.rows) ? static_cast<uint32_t>(id) : kInvalidRow;
}

// Issue the async copies of one sample's F rows (+ the dense vector as row F) into xs[32][STRIDE].
// src_lane = (E or table) + this lane's column offset; dst_lane = xs + sub*STRIDE + column offset.
// Sharded table (shard_base != null, smem copy of the peer shard pointers): row r is read from the owner's HBM
// over NVLink, shard_base[r % world] + (r / world) * D — the gather IS the collective.
template <int D, int STRIDE, bool SHARDED>
__device__ __forceinline__ void issue_rows(const float* __restrict__ src_lane, uint32_t my_row, int F, float* dst_lane,
                                           int sub, const float* __restrict__ dense_lane, float* dense_dst,
                                           const float* const* shard_base, uint32_t world, int col, bool l1,
                                           uint32_t small_base = 0xFFFFFFFFu, const float* small_rep = nullptr) {
  constexpr int kRowsPerIter = 32 / (D / 4);
#pragma unroll 4
  for (int r0 = 0; r0 < F; r0 += kRowsPerIter) {
    const int r = r0 + sub;                                        // <= 31
    const uint32_t row = __shfl_sync(0xffffffffu, my_row, r);
    const bool ok = row != kInvalidRow;
    const float* src = src_lane;
    if (ok) {
      if constexpr (SHARDED) {
        if (row >= small_base) {
          src = small_rep + static_cast<size_t>(row - small_base) * D + col;       // replicated table: this rank's own copy
        } else {
          const uint32_t local = row / world;
          src = shard_base[row - local * world] + static_cast<size_t>(local) * D + col;
        }
      } else {
        src = src_lane + static_cast<size_t>(row) * D;
      }
    }
    if (r < F) {   // invalid row: zero-filled by the copy
      if ((SHARDED && RB_PEER_CA) || l1) cp_async16_ca(dst_lane + r0 * STRIDE, src, ok ? 16 : 0);
      else cp_async16(dst_lane + r0 * STRIDE, src, ok ? 16 : 0);
    }
  }
  if (dense_lane != nullptr) cp_async16(dense_dst, dense_lane, 16);
}

template <typename T>
__device__ __forceinline__ int misalign_elems(const T* p) {
  return static_cast<int>((reinterpret_cast<uintptr_t>(p) & 15) / sizeof(T));
}

// Write `width` staged elements os[0..width) to the global row `grow` with aligned 16-byte stores.
// mis = misalignment (elements) of grow from a 16-byte boundary; chunk c covers row elements
// [c*EPC - mis, (c+1)*EPC - mis).
template <typename T>
__device__ __forceinline__ void store_row(T* __restrict__ grow, const T* os, int mis, int width, int lane) {
  constexpr int EPC = 16 / static_cast<int>(sizeof(T));
  const int nchunks = (mis + width + EPC - 1) / EPC;
  if (mis == 0) {
    const int full = width / EPC;
    for (int c = lane; c < full; c += 32) __stcs(reinterpret_cast<float4*>(grow) + c, reinterpret_cast<const float4*>(os)[c]);
    for (int e = full * EPC + lane; e < width; e += 32) grow[e] = os[e];
    return;
  }
  for (int c = lane; c < nchunks; c += 32) {
    const int e0 = c * EPC - mis;
    if (e0 >= 0 && e0 + EPC <= width) {
      if constexpr (sizeof(T) == 4) {
        __stcs(reinterpret_cast<float4*>(grow + e0), make_float4(os[e0], os[e0 + 1], os[e0 + 2], os[e0 + 3]));
      } else {
#pragma unroll
        for (int k = 0; k < EPC; ++k) grow[e0 + k] = os[e0 + k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < EPC; ++k)
        if (e0 + k >= 0 && e0 + k < width) grow[e0 + k] = os[e0 + k];
    }
  }
}

// Async read of `width` elements of the global row `grow` into gs[mis .. mis+width) (16-byte chunks
// line up on both sides; ragged ends go element-wise).
template <typename T>
__device__ __forceinline__ void load_row_async(const T* __restrict__ grow, T* gs, int mis, int width, int lane) {
  constexpr int EPC = 16 / static_cast<int>(sizeof(T));
  const int nchunks = (mis + width + EPC - 1) / EPC;
  const T* gbase = grow - mis;
  for (int c = lane; c < nchunks; c += 32) {
    const int e0 = c * EPC;
    if (e0 >= mis && e0 + EPC <= mis + width) {
      cp_async16(gs + e0, gbase + e0, 16);
    } else {
#pragma unroll
      for (int k = 0; k < EPC; ++k) {
        const int e = e0 + k;
        if (e >= mis && e < mis + width) {
          if constexpr (sizeof(T) == 4) cp_async4(gs + e, gbase + e);
          else gs[e] = gbase[e];   // bf16 rows are 16 B aligned and padded in practice; this keeps odd layouts correct
        }
      }
    }
  }
}

// ---- forward -----------------------------------------------------------------------------------------------
// Z = X X^T is symmetric: only the 6 tiles (m-tile, n-tile) that touch the upper triangle are
// computed; the self-interaction form (kept j <= i) reads element (j,i) instead.
// smem per warp: kIxStages x xs[32][D+8] fp32 | 16 B trash slot | staged output row.
template <int D, typename OUT, bool SHARDED>
__global__ void __launch_bounds__(IxWarps<D>::value * 32, 1)
dot_interaction_fwd_kernel(IxArgs a, OUT* __restrict__ out, int64_t out_stride, int write_width, int os_bytes) {
  griddep_wait();      // launched programmatically behind the kernel before it (common.cuh)
  constexpr int STRIDE = D + RB_IX_FWD_PAD;  // floats; (D+8) % 32 == 8 -> conflict-free 64-bit fragment loads
  constexpr int kXsFloats = 32 * STRIDE;
  const int kIxWarps = blockDim.x / 32;
  constexpr int kLanesPerRow = D / 4;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ const float* s_shards[RB_MAX_RANKS];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if constexpr (SHARDED) {
    if (threadIdx.x < a.world) s_shards[threadIdx.x] = a.shards[threadIdx.x];
    __syncthreads();
  }
  const float* const* shard_base = SHARDED ? s_shards : nullptr;
  unsigned char* my = smem + static_cast<size_t>(warp) * (kIxStages * kXsFloats * 4 + os_bytes);
  float* xs_base = reinterpret_cast<float*>(my);
  OUT* os = reinterpret_cast<OUT*>(my + kIxStages * kXsFloats * 4 + 16);   // row origin; [-16 B, 0) is the trash slot

  // one-time init: rows >= F' of every stage stay zero; masked / pad columns of the staged row stay zero
  for (int i = lane; i < (kIxStages * kXsFloats * 4 + os_bytes) / 16; i += 32) reinterpret_cast<float4*>(my)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();

  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kIxWarps;
  int64_t b = static_cast<int64_t>(blockIdx.x) * kIxWarps + warp;
  if (b >= a.B) return;

  const int g = lane / 4, t2 = (lane % 4) * 2;
  const int F = a.F;
  const int total = a.ncols + (a.tail ? D : 0);
  const int width = write_width > total ? write_width : total;
  if (a.pad_one && width > total && lane == 0) os[total] = Elem<OUT>::from(1.0f);   // pad columns are written once: they never change

  // where each accumulator element goes in the staged row (sample-independent):
  // tile T = (mt, nt) in {(0,0),(0,1),(0,2),(0,3),(1,2),(1,3)}, element k: i = mt*16+g+(k/2)*8, j = nt*8+t2+(k%2)
  constexpr int kTileM[6] = {0, 0, 0, 0, 1, 1};
  constexpr int kTileN[6] = {0, 1, 2, 3, 2, 3};
  int opos[6][4];
#pragma unroll
  for (int T = 0; T < 6; ++T)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = kTileM[T] * 16 + g + (k / 2) * 8;
      const int j = kTileN[T] * 8 + t2 + (k % 2);
      int pos = -static_cast<int>(16 / sizeof(OUT));             // trash slot
      if (i < a.Fp && j < a.Fp) {
        if (!a.self_interaction) {
          if (j > i) pos = out_pos(i, j, a);
        } else if (j >= i) {
          pos = out_pos(j, i, a);                                // kept (row j, col i): Z[j][i] == Z[i][j]
        }
      }
      opos[T][k] = pos;
    }

  const float* src_lane = (a.E != nullptr ? a.E : (SHARDED ? s_shards[0] : a.table)) + (lane % kLanesPerRow) * 4;
  const int sub = lane / kLanesPerRow;
  const int dst_off = sub * STRIDE + (lane % kLanesPerRow) * 4;
  const bool dense_lane_on = a.dense_vec != nullptr && lane < kLanesPerRow;
  const int64_t lane_off = (a.E == nullptr && a.map.field_row_offset != nullptr && lane < F) ? __ldg(a.map.field_row_offset + lane) : 0;

  uint32_t row_next = load_sample_row(a, b, lane, lane_off);
  const int col = (lane % kLanesPerRow) * 4;
  const bool l1 = rows_through_l1(a);
  auto issue = [&](int64_t bb, uint32_t rows, int st) {
    float* xn = xs_base + st * kXsFloats;
    issue_rows<D, STRIDE, SHARDED>(src_lane, rows, F, xn + dst_off, sub, dense_lane_on ? a.dense_vec + bb * D + lane * 4 : nullptr,
                                   xn + F * STRIDE + lane * 4, shard_base, a.world, col, l1, a.small_base,
                                   static_cast<const float*>(a.small_rep));
  };
  // prologue: the first kIxStages-1 samples of this warp are put in flight
#pragma unroll
  for (int s0 = 0; s0 < kIxStages - 1; ++s0) {
    const int64_t bs = b + s0 * nwarps;
    if (bs < a.B) issue(bs, row_next, s0);
    cp_async_commit();
    row_next = load_sample_row(a, bs + nwarps, lane, lane_off);
  }
  int stage = 0;

  for (; b < a.B; b += nwarps) {
    const int64_t bn = b + (kIxStages - 1) * nwarps;
    float* xs = xs_base + stage * kXsFloats;
    if (bn < a.B) {
      int sa = stage + kIxStages - 1;
      if (sa >= kIxStages) sa -= kIxStages;
      issue(bn, row_next, sa);
      row_next = load_sample_row(a, bn + nwarps, lane, lane_off);   // in flight during this sample's math
    }
    cp_async_commit();
    cp_async_wait<kIxStages - 1>();
    __syncwarp();

    if (SHARDED && a.x_save != nullptr) {   // the sample's operand rows as bf16, contiguous [F', D]: the backward reads them locally
      constexpr int kChunksPerRow = D / 8;
      __nv_bfloat16* xrow = a.x_save + b * a.Fp * D;
      for (int cidx = lane; cidx < a.Fp * kChunksPerRow; cidx += 32) {
        const int r = cidx / kChunksPerRow, c8 = (cidx % kChunksPerRow) * 8;
        const float4 lo = *reinterpret_cast<const float4*>(xs + r * STRIDE + c8);
        const float4 hi = *reinterpret_cast<const float4*>(xs + r * STRIDE + c8 + 4);
        uint4 pk;
        pk.x = pack_bf16(lo.x, lo.y);
        pk.y = pack_bf16(lo.z, lo.w);
        pk.z = pack_bf16(hi.x, hi.y);
        pk.w = pack_bf16(hi.z, hi.w);
        *reinterpret_cast<uint4*>(xrow + r * D + c8) = pk;
      }
    }

    float acc[6][4];
#pragma unroll
    for (int T = 0; T < 6; ++T)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[T][k] = 0.f;

#pragma unroll
    for (int k0 = 0; k0 < D; k0 += 16) {
      uint32_t af[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* p = xs + (mt * 16 + g) * STRIDE + k0 + t2;
        const float2 v0 = *reinterpret_cast<const float2*>(p);
        const float2 v1 = *reinterpret_cast<const float2*>(p + 8 * STRIDE);
        const float2 v2 = *reinterpret_cast<const float2*>(p + 8);
        const float2 v3 = *reinterpret_cast<const float2*>(p + 8 * STRIDE + 8);
        af[mt][0] = pack_bf16(v0.x, v0.y);
        af[mt][1] = pack_bf16(v1.x, v1.y);
        af[mt][2] = pack_bf16(v2.x, v2.y);
        af[mt][3] = pack_bf16(v3.x, v3.y);
      }
#pragma unroll
      for (int T = 0; T < 6; ++T) {
        // B[k][n] = X[n][k]: rows nt*8+g of X are rows (nt&1)*8+g of m-tile nt/2
        const int nt = kTileN[T];
        mma_bf16_16816(acc[T], af[kTileM[T]], af[nt / 2][(nt & 1) ? 1 : 0], af[nt / 2][(nt & 1) ? 3 : 2]);
      }
    }

    // epilogue: kept elements to their columns of the staged row; everything else to the trash slot
#pragma unroll
    for (int T = 0; T < 6; ++T)
#pragma unroll
      for (int k = 0; k < 4; ++k) os[opos[T][k]] = Elem<OUT>::from(acc[T][k]);
    if (a.tail) {
      for (int d = lane; d < D; d += 32) os[a.ncols + d] = Elem<OUT>::from(xs[F * STRIDE + d]);
    }
    __syncwarp();
    OUT* grow = out + b * out_stride;
    store_row<OUT>(grow, os, misalign_elems(grow), width, lane);
    __syncwarp();   // os and xs[stage] are free again
    if (++stage == kIxStages) stage = 0;
  }
}

// ---- forward, sharded table with a bf16 shadow ---------------------------------------------------------------------
// The owners keep a bf16 copy of their shard (refreshed by the optimizer row update): the rows cross
// NVLink as 2*D bytes instead of 4*D and are ALREADY the MMA operands (same values as rounding the fp32
// row on arrival, so results are bit-identical to the fp32-source kernel).  smem per warp:
// kIxStages x { x16[32][D+8] bf16 | dense vector fp32[D] } | 16 B trash slot | staged output row.
template <int D, typename OUT>
__global__ void __launch_bounds__(IxWarps<D>::value * 32, 1)
dot_interaction_fwd16_kernel(IxArgs a, const __nv_bfloat16* const* __restrict__ shadows, OUT* __restrict__ out, int64_t out_stride,
                             int write_width, int os_bytes) {
  constexpr int S16 = D + 8;                    // bf16 elements per tile row: 4*g + t bank pattern, conflict-free
  constexpr int kTileBytes = 32 * S16 * 2;
  constexpr int kStageBytes = kTileBytes + D * 4;
  const int kIxWarps = blockDim.x / 32;
  constexpr int kLanesPerRow = D / 8;           // 16-byte chunks per bf16 row
  constexpr int kRowsPerIter = 32 / kLanesPerRow;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ const __nv_bfloat16* s_shards[RB_MAX_RANKS];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x < a.world) s_shards[threadIdx.x] = shadows[threadIdx.x];
  __syncthreads();
  unsigned char* my = smem + static_cast<size_t>(warp) * (kIxStages * kStageBytes + os_bytes);
  OUT* os = reinterpret_cast<OUT*>(my + kIxStages * kStageBytes + 16);

  for (int i = lane; i < (kIxStages * kStageBytes + os_bytes) / 16; i += 32) reinterpret_cast<float4*>(my)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();

  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kIxWarps;
  int64_t b = static_cast<int64_t>(blockIdx.x) * kIxWarps + warp;
  if (b >= a.B) return;

  const int g = lane / 4, t2 = (lane % 4) * 2;
  const int F = a.F;
  const int total = a.ncols + (a.tail ? D : 0);
  const int width = write_width > total ? write_width : total;
  if (a.pad_one && width > total && lane == 0) os[total] = Elem<OUT>::from(1.0f);

  constexpr int kTileM[6] = {0, 0, 0, 0, 1, 1};
  constexpr int kTileN[6] = {0, 1, 2, 3, 2, 3};
  int opos[6][4];
#pragma unroll
  for (int T = 0; T < 6; ++T)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = kTileM[T] * 16 + g + (k / 2) * 8;
      const int j = kTileN[T] * 8 + t2 + (k % 2);
      int pos = -static_cast<int>(16 / sizeof(OUT));
      if (i < a.Fp && j < a.Fp) {
        if (!a.self_interaction) {
          if (j > i) pos = out_pos(i, j, a);
        } else if (j >= i) {
          pos = out_pos(j, i, a);
        }
      }
      opos[T][k] = pos;
    }

  const int sub = lane / kLanesPerRow;
  const int col = (lane % kLanesPerRow) * 8;
  const uint32_t world = static_cast<uint32_t>(a.world);
  const int64_t lane_off = (a.map.field_row_offset != nullptr && lane < F) ? __ldg(a.map.field_row_offset + lane) : 0;
  const bool dense_lane_on = a.dense_vec != nullptr && lane < D / 4;

  auto issue = [&](int64_t bb, uint32_t my_row, int st) {
    unsigned char* sp = my + st * kStageBytes;
    __nv_bfloat16* x16 = reinterpret_cast<__nv_bfloat16*>(sp);
#pragma unroll 2
    for (int r0 = 0; r0 < F; r0 += kRowsPerIter) {
      const int r = r0 + sub;
      const uint32_t row = __shfl_sync(0xffffffffu, my_row, r & 31);
      const bool ok = row != kInvalidRow;
      const __nv_bfloat16* src = s_shards[0];
      if (ok) {
        if (row >= a.small_base) {
          src = static_cast<const __nv_bfloat16*>(a.small_rep) + static_cast<size_t>(row - a.small_base) * D + col;
        } else {
          const uint32_t local = row / world;
          src = s_shards[row - local * world] + static_cast<size_t>(local) * D + col;
        }
      }
      if (r < F) cp_async16_ca(x16 + r * S16 + col, src, ok ? 16 : 0);
    }
    if (dense_lane_on) cp_async16(sp + kTileBytes + lane * 16, a.dense_vec + bb * D + lane * 4, 16);
  };

  uint32_t row_next = load_sample_row(a, b, lane, lane_off);
#pragma unroll
  for (int s0 = 0; s0 < kIxStages - 1; ++s0) {
    const int64_t bs = b + s0 * nwarps;
    if (bs < a.B) issue(bs, row_next, s0);
    cp_async_commit();
    row_next = load_sample_row(a, bs + nwarps, lane, lane_off);
  }
  int stage = 0;

  for (; b < a.B; b += nwarps) {
    const int64_t bn = b + (kIxStages - 1) * nwarps;
    unsigned char* sp = my + stage * kStageBytes;
    __nv_bfloat16* x16 = reinterpret_cast<__nv_bfloat16*>(sp);
    const float* dvec = reinterpret_cast<const float*>(sp + kTileBytes);
    if (bn < a.B) {
      int sa = stage + kIxStages - 1;
      if (sa >= kIxStages) sa -= kIxStages;
      issue(bn, row_next, sa);
      row_next = load_sample_row(a, bn + nwarps, lane, lane_off);
    }
    cp_async_commit();
    cp_async_wait<kIxStages - 1>();
    __syncwarp();

    if (a.dense_vec != nullptr) {   // the dense vector joins the tile as row F (bf16 operand)
      for (int d = lane * 2; d < D; d += 64) *reinterpret_cast<uint32_t*>(x16 + F * S16 + d) = pack_bf16(dvec[d], dvec[d + 1]);
      __syncwarp();
    }
    if (a.x_save != nullptr) {
      __nv_bfloat16* xrow = a.x_save + b * a.Fp * D;
      for (int cidx = lane; cidx < a.Fp * kLanesPerRow; cidx += 32) {
        const int r = cidx / kLanesPerRow, c8 = (cidx % kLanesPerRow) * 8;
        *reinterpret_cast<uint4*>(xrow + r * D + c8) = *reinterpret_cast<const uint4*>(x16 + r * S16 + c8);
      }
    }

    float acc[6][4];
#pragma unroll
    for (int T = 0; T < 6; ++T)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[T][k] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < D; k0 += 16) {
      uint32_t af[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const __nv_bfloat16* p = x16 + (mt * 16 + g) * S16 + k0 + t2;
        af[mt][0] = *reinterpret_cast<const uint32_t*>(p);
        af[mt][1] = *reinterpret_cast<const uint32_t*>(p + 8 * S16);
        af[mt][2] = *reinterpret_cast<const uint32_t*>(p + 8);
        af[mt][3] = *reinterpret_cast<const uint32_t*>(p + 8 * S16 + 8);
      }
#pragma unroll
      for (int T = 0; T < 6; ++T) {
        const int nt = kTileN[T];
        mma_bf16_16816(acc[T], af[kTileM[T]], af[nt / 2][(nt & 1) ? 1 : 0], af[nt / 2][(nt & 1) ? 3 : 2]);
      }
    }
#pragma unroll
    for (int T = 0; T < 6; ++T)
#pragma unroll
      for (int k = 0; k < 4; ++k) os[opos[T][k]] = Elem<OUT>::from(acc[T][k]);
    if (a.tail) {
      for (int d = lane; d < D; d += 32) os[a.ncols + d] = Elem<OUT>::from(dvec[d]);
    }
    __syncwarp();
    OUT* grow = out + b * out_stride;
    store_row<OUT>(grow, os, misalign_elems(grow), width, lane);
    __syncwarp();
    if (++stage == kIxStages) stage = 0;
  }
}

// ---- backward ------------------------------------------------------------------------------------------------
// dX = (G + G^T) X.  smem per warp: kIxStages x { xs[32][D+4] fp32 | 16 B of zeros | staged dOut row }.
// SRC: where X comes from — 0: rows gathered from E / the local table, 1: rows gathered from the shards in peer
// memory, 2: the bf16 operand rows the forward saved (x_load)
// UPD (SRC == 0 only): rows flagged in u.single get their optimizer update here.  Per warp, next to the ring: the m / v rows of
// the CURRENT sample's flagged rows (cp.async, issued while the sample's X rows are still landing) and, per stage, the row ids
// and flags of the sample staged there.  The epilogue updates W (the fp32 row already in the stage), m and v in place in
// shared memory, column block by column block — columns of X are dead as MMA operands once their block is done — and the
// finished rows go back to the table / state with full-row 16-byte stores.
constexpr int kIxUpdWarps = 8;     // the fused-update form: more shared memory and more registers per warp
template <int D, typename DOUT, bool SELF, int SRC, bool UPD = false>
__global__ void __launch_bounds__(UPD ? kIxUpdWarps * 32 : IxWarps<D>::value * 32, 1)
dot_interaction_bwd_kernel(IxArgs a, const DOUT* __restrict__ dOut, int64_t dout_stride, float* __restrict__ dE,
                           float* __restrict__ d_dense, int copy_width, int gs_bytes, IxUpdate u = IxUpdate{}) {
  griddep_wait();      // launched programmatically behind the kernel before it (common.cuh)
  static_assert(!UPD || SRC == 0, "the fused row update reads fp32 rows of the local table");
  constexpr int STRIDE = D + 4;  // floats; 2*(D+4) % 32 == 8 -> conflict-free 32-bit B-fragment loads
  constexpr int MS = D + 8;      // floats per staged m / v row: conflict-free 64-bit accesses of the epilogue
  constexpr int kXsFloats = 32 * STRIDE;
  const int kIxWarps = blockDim.x / 32;
  constexpr int kLanesPerRow = D / 4;
  constexpr int EPC = 16 / static_cast<int>(sizeof(DOUT));
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ const float* s_shards[RB_MAX_RANKS];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if constexpr (SRC == 1) {
    if (threadIdx.x < a.world) s_shards[threadIdx.x] = a.shards[threadIdx.x];
    __syncthreads();
  }
  const float* const* shard_base = SRC == 1 ? s_shards : nullptr;
  const int stage_bytes = kXsFloats * 4 + gs_bytes;
  const int upd_bytes = UPD ? (2 * a.F * MS * 4 + kIxStages * 256) : 0;
  unsigned char* my = smem + static_cast<size_t>(warp) * (kIxStages * stage_bytes + upd_bytes);
  float* mvm = reinterpret_cast<float*>(my + kIxStages * stage_bytes);            // [F][MS] state0 rows of the current sample
  float* mvv = mvm + a.F * MS;                                                     // [F][MS] state1 rows
  uint32_t* st_rows = reinterpret_cast<uint32_t*>(mvv + a.F * MS);                 // [kIxStages][32] row ids of the staged sample
  uint32_t* st_single = st_rows + kIxStages * 32;                                  // [kIxStages][32] its flags
  if constexpr (UPD) {
    if (u.alpha_dev != nullptr) u.math.alpha = __ldg(u.alpha_dev);
  }

  for (int i = lane; i < kIxStages * stage_bytes / 16; i += 32) reinterpret_cast<float4*>(my)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();

  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kIxWarps;
  int64_t b = static_cast<int64_t>(blockIdx.x) * kIxWarps + warp;
  if (b >= a.B) return;

  const int g = lane / 4, t2 = (lane % 4) * 2;
  const int F = a.F;

  // Offsets (elements, relative to the staged row's element 0) of the 32 S = G + G^T values this lane
  // feeds into the A fragments: element e = ((mt*2 + ks)*4 + reg)*2 + c  <->  S[i][j],
  //   i = mt*16 + g + (reg & 1)*8,  j = ks*16 + t2 + c + (reg >> 1)*8.
  // Zero entries point at the zero slot (-EPC: 16 B below the row in every alignment).
  int aoff[32];
  uint32_t diag = 0;   // SELF only: S[i][i] = 2*G[i][i]
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const int c = e & 1, reg = (e >> 1) & 3, ks = (e >> 3) & 1, mt = e >> 4;
    const int i = mt * 16 + g + (reg & 1) * 8;
    const int j = ks * 16 + t2 + c + (reg >> 1) * 8;
    int off = -EPC;
    if (i < a.Fp && j < a.Fp) {
      if (i == j) {
        if (SELF) {
          off = out_pos(i, i, a);
          diag |= (1u << e);
        }
      } else {
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        off = SELF ? out_pos(hi, lo, a) : out_pos(lo, hi, a);
      }
    }
    aoff[e] = off;
  }
  // which of this lane's 4 output rows are embedding rows / the dense row
  bool is_emb[2][2], is_dense[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = mt * 16 + g + h * 8;
      is_emb[mt][h] = (i < F) && dE != nullptr;
      is_dense[mt][h] = (i == F) && a.dense_vec != nullptr && d_dense != nullptr;
    }
  // fields whose gradient rows go to the compact tensor of the replicated tables
  int de_slot[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = mt * 16 + g + h * 8;
      de_slot[mt][h] = (a.de_slot != nullptr && i < F) ? __ldg(a.de_slot + i) : -1;
    }

  const float* src_lane = (a.E != nullptr ? a.E : (SRC == 1 ? s_shards[0] : a.table)) + (lane % kLanesPerRow) * 4;
  const int sub = lane / kLanesPerRow;
  const int dst_off = sub * STRIDE + (lane % kLanesPerRow) * 4;
  const bool dense_lane_on = a.dense_vec != nullptr && lane < kLanesPerRow;
  const int64_t lane_off = (SRC != 2 && a.E == nullptr && a.map.field_row_offset != nullptr && lane < F) ? __ldg(a.map.field_row_offset + lane) : 0;

  constexpr int S16 = D + 8;   // bf16 elements per row of the staged X when it comes from x_load
  constexpr bool xl = (SRC == 2);
  const bool l1 = rows_through_l1(a);
  uint32_t single_next = 0;
  auto load_single = [&](int64_t bb) -> uint32_t {
    if constexpr (UPD) return (lane < F && bb < a.B) ? static_cast<uint32_t>(__ldg(u.single + bb * F + lane)) : 0u;
    return 0u;
  };
  auto issue = [&](int64_t bb, uint32_t row, int st) {
    unsigned char* sp = my + st * stage_bytes;
    float* xn = reinterpret_cast<float*>(sp);
    if constexpr (UPD) {
      st_rows[st * 32 + lane] = row;
      st_single[st * 32 + lane] = (row != kInvalidRow) ? single_next : 0u;
    }
    if (xl) {
      constexpr int kChunksPerRow = D / 8;
      const __nv_bfloat16* xrow = a.x_load + bb * a.Fp * D;
      __nv_bfloat16* x16 = reinterpret_cast<__nv_bfloat16*>(sp);
      for (int cidx = lane; cidx < a.Fp * kChunksPerRow; cidx += 32) {
        const int r = cidx / kChunksPerRow, c8 = (cidx % kChunksPerRow) * 8;
        cp_async16(x16 + r * S16 + c8, xrow + r * D + c8, 16);
      }
    } else
    issue_rows<D, STRIDE, SRC == 1>(src_lane, row, F, xn + dst_off, sub, dense_lane_on ? a.dense_vec + bb * D + lane * 4 : nullptr,
                          xn + F * STRIDE + lane * 4, shard_base, a.world, (lane % kLanesPerRow) * 4, l1, a.small_base,
                          static_cast<const float*>(a.small_rep));
    const DOUT* grow = dOut + bb * dout_stride;
    load_row_async<DOUT>(grow, reinterpret_cast<DOUT*>(sp + kXsFloats * 4 + 16), misalign_elems(grow), copy_width, lane);
  };

  uint32_t row_next = xl ? kInvalidRow : load_sample_row(a, b, lane, lane_off);
  single_next = load_single(b);
#pragma unroll
  for (int s0 = 0; s0 < kIxStages - 1; ++s0) {
    const int64_t bs = b + s0 * nwarps;
    if (bs < a.B) issue(bs, row_next, s0);
    cp_async_commit();
    if (!xl) row_next = load_sample_row(a, bs + nwarps, lane, lane_off);
    single_next = load_single(bs + nwarps);
  }
  int stage = 0;

  for (; b < a.B; b += nwarps) {
    const int64_t bn = b + (kIxStages - 1) * nwarps;
    unsigned char* sp = my + stage * stage_bytes;
    const float* xs = reinterpret_cast<const float*>(sp);
    uint32_t rowc = kInvalidRow, smask = 0;
    if constexpr (UPD) {
      // this sample's rows and flags (stored when its stage was issued); the state rows of its flagged rows start
      // moving now, ahead of the stage issued below
      __syncwarp();
      rowc = st_rows[stage * 32 + lane];
      smask = __ballot_sync(0xffffffffu, st_single[stage * 32 + lane] != 0);
      constexpr int kRowsPerIter = 32 / kLanesPerRow;
      const int colq = (lane % kLanesPerRow) * 4;
#pragma unroll 4
      for (int r0 = 0; r0 < F; r0 += kRowsPerIter) {
        const int r = r0 + sub;
        const uint32_t row = __shfl_sync(0xffffffffu, rowc, r);
        if (r < F && ((smask >> r) & 1u)) {
          const size_t o = static_cast<size_t>(row) * D + colq;
          if (u.s0 != nullptr) cp_async16(mvm + r * MS + colq, u.s0 + o, 16);
          if (u.s1 != nullptr) cp_async16(mvv + r * MS + colq, u.s1 + o, 16);
        }
      }
      cp_async_commit();
    }
    if (bn < a.B) {
      int sa = stage + kIxStages - 1;
      if (sa >= kIxStages) sa -= kIxStages;
      issue(bn, row_next, sa);
      if (!xl) row_next = load_sample_row(a, bn + nwarps, lane, lane_off);
      single_next = load_single(bn + nwarps);
    }
    cp_async_commit();
    cp_async_wait<UPD ? kIxStages : kIxStages - 1>();     // UPD: the state-row group and the next stage may still be in flight
    __syncwarp();

    // staged row element e lives at gs[e]
    const DOUT* gs = reinterpret_cast<const DOUT*>(sp + kXsFloats * 4 + 16) + misalign_elems(dOut + b * dout_stride);

    // A = S as bf16 fragments [mt][ks][4]
    uint32_t af[2][2][4];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float v0 = Elem<DOUT>::to_f32(gs[aoff[2 * q]]);
      float v1 = Elem<DOUT>::to_f32(gs[aoff[2 * q + 1]]);
      if (SELF) {
        if (diag & (1u << (2 * q))) v0 += v0;
        if (diag & (1u << (2 * q + 1))) v1 += v1;
      }
      af[q >> 3][(q >> 2) & 1][q & 3] = pack_bf16(v0, v1);
    }

    float* de_lane = dE != nullptr ? dE + (b * F + g) * D + t2 : nullptr;
    if constexpr (UPD) {
      cp_async_wait<1>();       // the state rows have landed (the next stage may still be in flight)
      __syncwarp();
    }
    float* xw = reinterpret_cast<float*>(sp);
#pragma unroll 4
    for (int nt = 0; nt < D / 8; ++nt) {
      float acc[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[mt][k] = 0.f;
      const float* p = xs + t2 * STRIDE + nt * 8 + g;   // B[k][n] = X[k][n]: two consecutive feature rows per register
      const unsigned short* p16 = reinterpret_cast<const unsigned short*>(sp) + t2 * S16 + nt * 8 + g;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t b0, b1;
        if (xl) {
          b0 = static_cast<uint32_t>(p16[ks * 16 * S16]) | (static_cast<uint32_t>(p16[(ks * 16 + 1) * S16]) << 16);
          b1 = static_cast<uint32_t>(p16[(ks * 16 + 8) * S16]) | (static_cast<uint32_t>(p16[(ks * 16 + 9) * S16]) << 16);
        } else {
          b0 = pack_bf16(p[ks * 16 * STRIDE], p[(ks * 16 + 1) * STRIDE]);
          b1 = pack_bf16(p[(ks * 16 + 8) * STRIDE], p[(ks * 16 + 9) * STRIDE]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(acc[mt], af[mt][ks], b0, b1);
      }
      if constexpr (UPD) {
        // The four rows this lane holds two columns of, all at once and branch-free: eight independent chains of the
        // optimizer arithmetic (its IEEE divisions and square roots are long dependent sequences) instead of four
        // divergent branches of two.  Rows that are not flagged compute on whatever the buffers hold and store nothing here.
        Row<8> gr, wr, mr, vr;
        float2* wp[4];
        float2* mp[4];
        float2* vp[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = (q >> 1) * 16 + g + (q & 1) * 8;
          const int ic = i < F ? i : F - 1;                       // the m / v buffers hold F rows
          wp[q] = reinterpret_cast<float2*>(xw + i * STRIDE + nt * 8 + t2);
          mp[q] = reinterpret_cast<float2*>(mvm + ic * MS + nt * 8 + t2);
          vp[q] = reinterpret_cast<float2*>(mvv + ic * MS + nt * 8 + t2);
          const float2 w2 = *wp[q], m2 = *mp[q], v2 = *vp[q];
          // the row's only gradient this step: its summed gradient is 0 + g, exactly as the segmented reduction forms it
          gr.v[2 * q] = __fadd_rn(0.f, acc[q >> 1][(q & 1) * 2]);
          gr.v[2 * q + 1] = __fadd_rn(0.f, acc[q >> 1][(q & 1) * 2 + 1]);
          wr.v[2 * q] = w2.x; wr.v[2 * q + 1] = w2.y;
          mr.v[2 * q] = m2.x; mr.v[2 * q + 1] = m2.y;
          vr.v[2 * q] = v2.x; vr.v[2 * q + 1] = v2.y;
        }
        opt_row_math<8>(u.math, gr, wr, mr, vr);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int mt = q >> 1, h = q & 1;
          const int i = mt * 16 + g + h * 8;
          if (is_emb[mt][h] && ((smask >> i) & 1u)) {
            *wp[q] = make_float2(wr.v[2 * q], wr.v[2 * q + 1]);
            *mp[q] = make_float2(mr.v[2 * q], mr.v[2 * q + 1]);
            *vp[q] = make_float2(vr.v[2 * q], vr.v[2 * q + 1]);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float2 v = make_float2(acc[mt][h * 2], acc[mt][h * 2 + 1]);
          if (UPD && is_emb[mt][h] && ((smask >> (mt * 16 + g + h * 8)) & 1u)) {
            // updated above
          } else if (is_emb[mt][h]) {
            if (de_slot[mt][h] >= 0)
              __stcs(reinterpret_cast<float2*>(a.dE_small + (b * a.num_small + de_slot[mt][h]) * D + nt * 8 + t2), v);
            else
              __stcs(reinterpret_cast<float2*>(de_lane + (mt * 16 + h * 8) * D + nt * 8), v);
          } else if (is_dense[mt][h]) {
            const int col = nt * 8 + t2;
            if (a.tail) {
              v.x += Elem<DOUT>::to_f32(gs[a.ncols + col]);
              v.y += Elem<DOUT>::to_f32(gs[a.ncols + col + 1]);
            }
            *reinterpret_cast<float2*>(d_dense + b * D + col) = v;
          }
        }
    }
    if constexpr (UPD) {
      // finished rows back to the table and its optimizer state: 16 bytes per lane, a full row per kLanesPerRow lanes
      __syncwarp();
      constexpr int kRowsPerIter = 32 / kLanesPerRow;
      const int colq = (lane % kLanesPerRow) * 4;
#pragma unroll 4
      for (int r0 = 0; r0 < F; r0 += kRowsPerIter) {
        const int r = r0 + sub;
        const uint32_t row = __shfl_sync(0xffffffffu, rowc, r);
        if (r < F && ((smask >> r) & 1u)) {
          const size_t o = static_cast<size_t>(row) * D + colq;
          *reinterpret_cast<float4*>(u.table + o) = *reinterpret_cast<const float4*>(xw + r * STRIDE + colq);
          if (u.s0 != nullptr) *reinterpret_cast<float4*>(u.s0 + o) = *reinterpret_cast<const float4*>(mvm + r * MS + colq);
          if (u.s1 != nullptr) *reinterpret_cast<float4*>(u.s1 + o) = *reinterpret_cast<const float4*>(mvv + r * MS + colq);
        }
      }
    }
    __syncwarp();   // the stage may be overwritten by the next iteration's copies
    if (++stage == kIxStages) stage = 0;
  }
}

static int fill_args(IxArgs* a, const float* E, const float* table, int64_t rows, const void* idx, int idx_type,
                     const int64_t* off, const float* dense_vec, int64_t B, int F, int D, int self_interaction,
                     int skip_gather, int tail, const float* const* shards = nullptr, int world = 1) {
  a->row_cache = RB_ROW_CACHE_L2;
  a->row_cache_hint = nullptr;
  a->small_base = 0xFFFFFFFFu;
  a->small_rep = nullptr;
  a->de_slot = nullptr;
  a->dE_small = nullptr;
  a->num_small = 0;
  RB_CHECK_ARG(B >= 0 && F > 0, RB_ERR_ARG, "bad B/F");
  RB_CHECK_ARG(D == 16 || D == 32 || D == 64 || D == 128, RB_ERR_SHAPE, "dot interaction needs D in {16,32,64,128}, got %d", D);
  const int Fp = F + (dense_vec != nullptr ? 1 : 0);
  RB_CHECK_ARG(Fp <= 32, RB_ERR_SHAPE, "dot interaction supports at most 32 features, got %d", Fp);
  RB_CHECK_ARG(!tail || dense_vec != nullptr, RB_ERR_ARG, "tail requires dense_vec");
  RB_CHECK_ARG(shards == nullptr || (E == nullptr && world >= 1 && world <= RB_MAX_RANKS), RB_ERR_ARG,
               "sharded form: no E, world in [1, %d]", RB_MAX_RANKS);
  if (E == nullptr && shards != nullptr) {
    RB_CHECK_ARG(idx != nullptr && rows > 0 && rows < 0xFFFFFFFFll, RB_ERR_ARG, "sharded gather needs idx and rows < 2^32-1");
    RB_CHECK_ARG(idx_type == RB_I32 || idx_type == RB_I64, RB_ERR_ARG, "bad index type");
  } else if (E == nullptr) {
    RB_CHECK_ARG(table != nullptr && idx != nullptr && rows > 0, RB_ERR_ARG, "fused gather needs table and idx");
    RB_CHECK_ARG(idx_type == RB_I32 || idx_type == RB_I64, RB_ERR_ARG, "bad index type");
    RB_CHECK_ARG(aligned_for(table, 4), RB_ERR_ALIGN, "table not 16 B aligned");
    RB_CHECK_ARG(rows < 0xFFFFFFFFll, RB_ERR_ARG, "dot interaction addresses at most 2^32-2 table rows");
  } else {
    RB_CHECK_ARG(aligned_for(E, 4), RB_ERR_ALIGN, "E not 16 B aligned");
    RB_CHECK_ARG(B * F < 0xFFFFFFFFll, RB_ERR_ARG, "B*F must be below 2^32-1");
  }
  RB_CHECK_ARG(dense_vec == nullptr || aligned_for(dense_vec, 4), RB_ERR_ALIGN, "dense_vec not 16 B aligned");
  a->E = E;
  a->table = table;
  a->shards = shards;
  a->world = world;
  a->x_save = nullptr;
  a->x_load = nullptr;
  a->map = make_index_map(idx, idx_type, off, 0, rows, F);
  a->dense_vec = dense_vec;
  a->B = B;
  a->F = F;
  a->Fp = Fp;
  a->self_interaction = self_interaction ? 1 : 0;
  a->skip_gather = skip_gather ? 1 : 0;
  a->tail = tail ? 1 : 0;
  a->ncols = skip_gather ? Fp * Fp : (self_interaction ? Fp * (Fp + 1) / 2 : Fp * (Fp - 1) / 2);
  a->pad_one = 0;
  return RB_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) {
    set_error("dot interaction needs %zu bytes of shared memory per CTA (> 227 KiB)", bytes);
    return RB_ERR_SHAPE;
  }
  if (bytes > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return RB_OK;
}

static unsigned int persistent_grid(int64_t B, int warps) {
  const int64_t need = (B + warps - 1) / warps;
  return static_cast<unsigned int>(need < kNumSMs ? (need < 1 ? 1 : need) : kNumSMs);
}

template <typename OUT>
static int launch_fwd(const IxArgs& a, int D, OUT* out, int64_t out_stride, int write_width, cudaStream_t st) {
  // staged row: up to 15 bytes of misalignment + the widest row written, rounded to 16 B
  const int total = a.ncols + (a.tail ? D : 0);
  const int width = write_width > total ? write_width : total;
  const int os_bytes = 16 + ((width * static_cast<int>(sizeof(OUT)) + 15) / 16) * 16;   // trash slot + staged row
  int rc = RB_OK;
#define LAUNCH(DD)                                                                                                      \
  {                                                                                                                     \
    const size_t per_warp = static_cast<size_t>(kIxStages) * 32 * (DD + RB_IX_FWD_PAD) * 4 + os_bytes;                  \
    const int W = warps_that_fit(per_warp);                                                                             \
    RB_CHECK_ARG(W >= 1, RB_ERR_SHAPE, "interaction row does not fit in shared memory");                                \
    size_t smem = static_cast<size_t>(W) * per_warp;                                                                    \
    if (a.shards != nullptr) {                                                                                          \
      rc = set_smem(dot_interaction_fwd_kernel<DD, OUT, true>, smem);                                                   \
      if (rc != RB_OK) return rc;                                                                                       \
      dot_interaction_fwd_kernel<DD, OUT, true><<<persistent_grid(a.B, W), W * 32, smem, st>>>(a, out, out_stride, write_width, \
                                                                                                 os_bytes);             \
    } else {                                                                                                            \
      rc = set_smem(dot_interaction_fwd_kernel<DD, OUT, false>, smem);                                                  \
      if (rc != RB_OK) return rc;                                                                                       \
      RB_CUDA(launch_dependent(dot_interaction_fwd_kernel<DD, OUT, false>, persistent_grid(a.B, W), W * 32, smem, st, true, a, out, \
                               out_stride, write_width, os_bytes));                                                     \
    }                                                                                                                   \
  }
  if (D == 16) LAUNCH(16) else if (D == 32) LAUNCH(32) else if (D == 64) LAUNCH(64) else LAUNCH(128)
#undef LAUNCH
  RB_LAUNCH_CHECK("dot_interaction_fwd_kernel");
  return RB_OK;
}

// The sharded forward waits on NVLink, not on its SM: with fewer warps (less shared memory) per CTA the radix sort of the
// owner's pairs, which runs beside it on the side stream, finds room on the same SMs instead of queueing behind it.
// RB_IX16_WARPS caps the warps per CTA (tuning; default: as many as fit).
static int sharded_fwd_warp_cap() {
  const char* e = getenv("RB_IX16_WARPS");
  const int v = e != nullptr ? atoi(e) : 0;
  return v > 0 ? v : 64;
}

template <typename OUT>
static int launch_fwd16(const IxArgs& a, int D, const __nv_bfloat16* const* shadows, OUT* out, int64_t out_stride, int write_width,
                        cudaStream_t st) {
  const int total = a.ncols + (a.tail ? D : 0);
  const int width = write_width > total ? write_width : total;
  const int os_bytes = 16 + ((width * static_cast<int>(sizeof(OUT)) + 15) / 16) * 16;
  int rc = RB_OK;
#define LAUNCH(DD)                                                                                                      \
  {                                                                                                                     \
    const size_t per_warp = static_cast<size_t>(kIxStages) * (32 * (DD + 8) * 2 + DD * 4) + os_bytes;                   \
    const int W = std::min(warps_that_fit(per_warp), sharded_fwd_warp_cap());                                           \
    RB_CHECK_ARG(W >= 1, RB_ERR_SHAPE, "interaction row does not fit in shared memory");                                \
    size_t smem = static_cast<size_t>(W) * per_warp;                                                                    \
    rc = set_smem(dot_interaction_fwd16_kernel<DD, OUT>, smem);                                                         \
    if (rc != RB_OK) return rc;                                                                                         \
    dot_interaction_fwd16_kernel<DD, OUT><<<persistent_grid(a.B, W), W * 32, smem, st>>>(a, shadows, out, out_stride,   \
                                                                                           write_width, os_bytes);      \
  }
  if (D == 16) LAUNCH(16) else if (D == 32) LAUNCH(32) else if (D == 64) LAUNCH(64) else LAUNCH(128)
#undef LAUNCH
  RB_LAUNCH_CHECK("dot_interaction_fwd16_kernel");
  return RB_OK;
}

#define BWD_ONE(DD, SELFV, SRCV)                                                                                        \
  {                                                                                                                     \
    rc = set_smem(dot_interaction_bwd_kernel<DD, DOUT, SELFV, SRCV>, smem);                                             \
    if (rc != RB_OK) return rc;                                                                                         \
    RB_CUDA(launch_dependent(dot_interaction_bwd_kernel<DD, DOUT, SELFV, SRCV>, persistent_grid(a.B, W), W * 32, smem, st, true, a, dOut, \
                             dout_stride, dE, d_dense, copy_width, gs_bytes, IxUpdate{}));                              \
  }
#define BWD_DISPATCH(DD, MODE)                                                                                          \
  if (a.self_interaction) {                                                                                             \
    if (MODE == 2) BWD_ONE(DD, true, 2) else if (MODE == 1) BWD_ONE(DD, true, 1) else BWD_ONE(DD, true, 0)              \
  } else {                                                                                                              \
    if (MODE == 2) BWD_ONE(DD, false, 2) else if (MODE == 1) BWD_ONE(DD, false, 1) else BWD_ONE(DD, false, 0)           \
  }

template <typename DOUT>
static int launch_bwd(const IxArgs& a, int D, const DOUT* dOut, int64_t dout_stride, float* dE, float* d_dense, cudaStream_t st) {
  const int total = a.ncols + (a.tail ? D : 0);
  // copy whole 16-byte chunks when the row's padding allows it (no element-wise tail)
  constexpr int EPC = 16 / static_cast<int>(sizeof(DOUT));
  const int rounded = (total + EPC - 1) / EPC * EPC;
  const int copy_width = (dout_stride >= rounded) ? rounded : total;
  const int gs_bytes = 16 + ((copy_width * static_cast<int>(sizeof(DOUT)) + 15 + 15) / 16) * 16;   // zero slot + row at any alignment
  int rc = RB_OK;
#define LAUNCH(DD)                                                                                                      \
  {                                                                                                                     \
    const size_t per_warp = static_cast<size_t>(kIxStages) * (32 * (DD + 4) * 4 + gs_bytes);                            \
    const int W = warps_that_fit(per_warp);                                                                             \
    RB_CHECK_ARG(W >= 1, RB_ERR_SHAPE, "interaction row does not fit in shared memory");                                \
    size_t smem = static_cast<size_t>(W) * per_warp;                                                                    \
    const int src_mode = a.x_load != nullptr ? 2 : (a.shards != nullptr ? 1 : 0);                                       \
    BWD_DISPATCH(DD, src_mode)                                                                                          \
  }
  if (D == 16) LAUNCH(16) else if (D == 32) LAUNCH(32) else if (D == 64) LAUNCH(64) else LAUNCH(128)
#undef LAUNCH
  RB_LAUNCH_CHECK("dot_interaction_bwd_kernel");
  return RB_OK;
}

// the fused-update form: local table, fp32 rows (SRC == 0)
template <typename DOUT>
static int launch_bwd_update(const IxArgs& a, int D, const DOUT* dOut, int64_t dout_stride, float* dE, float* d_dense, const IxUpdate& u,
                             cudaStream_t st) {
  const int total = a.ncols + (a.tail ? D : 0);
  constexpr int EPC = 16 / static_cast<int>(sizeof(DOUT));
  const int rounded = (total + EPC - 1) / EPC * EPC;
  const int copy_width = (dout_stride >= rounded) ? rounded : total;
  const int gs_bytes = 16 + ((copy_width * static_cast<int>(sizeof(DOUT)) + 15 + 15) / 16) * 16;
  int rc = RB_OK;
#define LAUNCH_U(DD, SELFV)                                                                                             \
  {                                                                                                                     \
    const size_t per_warp = static_cast<size_t>(kIxStages) * (32 * (DD + 4) * 4 + gs_bytes) + 2 * a.F * (DD + 8) * 4 + kIxStages * 256; \
    const int W = std::min(warps_that_fit(per_warp), kIxUpdWarps);                                                      \
    RB_CHECK_ARG(W >= 1, RB_ERR_SHAPE, "interaction row does not fit in shared memory");                                \
    const size_t smem = static_cast<size_t>(W) * per_warp;                                                              \
    rc = set_smem(dot_interaction_bwd_kernel<DD, DOUT, SELFV, 0, true>, smem);                                          \
    if (rc != RB_OK) return rc;                                                                                         \
    dot_interaction_bwd_kernel<DD, DOUT, SELFV, 0, true><<<persistent_grid(a.B, W), W * 32, smem, st>>>(a, dOut, dout_stride, dE, \
                                                                                                          d_dense, copy_width, \
                                                                                                          gs_bytes, u); \
  }
#define LAUNCH_UD(DD)                                                                                                   \
  if (a.self_interaction) LAUNCH_U(DD, true) else LAUNCH_U(DD, false)
  if (D == 16) LAUNCH_UD(16) else if (D == 32) LAUNCH_UD(32) else if (D == 64) LAUNCH_UD(64) else LAUNCH_UD(128)
#undef LAUNCH_UD
#undef LAUNCH_U
  RB_LAUNCH_CHECK("dot_interaction_bwd_kernel (fused row update)");
  return RB_OK;
}

}  // namespace rb

using namespace rb;

extern "C" int rb_dot_interaction_bwd_update(float* table, int64_t rows, const void* idx, int32_t idx_type,
                                             const int64_t* field_row_offset, const float* dense_vec, int64_t B, int32_t F, int32_t D,
                                             int32_t self_interaction, int32_t skip_gather, int32_t tail, const void* dOut,
                                             int32_t dout_dtype, int64_t dout_stride, float* dE, float* d_dense,
                                             const uint8_t* single, float* state0, float* state1, const rb_opt_params* opt,
                                             int32_t row_cache, const int32_t* row_cache_hint, void* stream) {
  IxArgs a;
  int rc = fill_args(&a, nullptr, table, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D, self_interaction, skip_gather, tail);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(row_cache >= RB_ROW_CACHE_L2 && row_cache <= RB_ROW_CACHE_AUTO, RB_ERR_ARG, "bad row_cache %d", row_cache);
  a.row_cache = row_cache;
  a.row_cache_hint = row_cache_hint;
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(dOut != nullptr && dout_stride >= a.ncols + (a.tail ? D : 0), RB_ERR_ARG, "dOut is null or stride too small");
  RB_CHECK_ARG(dout_dtype == RB_F32 || dout_dtype == RB_BF16, RB_ERR_ARG, "bad dout_dtype %d", dout_dtype);
  RB_CHECK_ARG(dE != nullptr && aligned_for(dE, 4) && (d_dense == nullptr || aligned_for(d_dense, 4)), RB_ERR_ALIGN,
               "dE (required: rows the step touches more than once still go through it) / d_dense not 16 B aligned");
  RB_CHECK_ARG(single != nullptr && opt != nullptr, RB_ERR_ARG, "single / opt is null");
  const int o = opt->optimizer;
  RB_CHECK_ARG(o == RB_OPT_SGD || o == RB_OPT_ADAGRAD || o == RB_OPT_ADAM_LAZY, RB_ERR_ARG,
               "the fused row update serves the row-sparse optimizers (SGD, Adagrad, lazy Adam), got %d", o);
  RB_CHECK_ARG(o != RB_OPT_ADAM_LAZY || (state0 != nullptr && state1 != nullptr && opt->step >= 1), RB_ERR_ARG, "Adam needs m, v and step >= 1");
  RB_CHECK_ARG(o != RB_OPT_ADAGRAD || state0 != nullptr, RB_ERR_ARG, "Adagrad needs its accumulator");
  RB_CHECK_ARG((state0 == nullptr || aligned_for(state0, 4)) && (state1 == nullptr || aligned_for(state1, 4)), RB_ERR_ALIGN,
               "optimizer state not 16 B aligned");
  IxUpdate u;
  u.single = single;
  u.table = table;
  u.s0 = o == RB_OPT_SGD ? nullptr : state0;
  u.s1 = o == RB_OPT_ADAM_LAZY ? state1 : nullptr;
  u.math.opt = o;
  u.math.lr = opt->lr;
  u.math.b1 = opt->beta_1;
  u.math.b2 = opt->beta_2;
  u.math.omb1 = 1.0f - opt->beta_1;
  u.math.omb2 = 1.0f - opt->beta_2;
  u.math.eps = opt->epsilon;
  u.math.alpha = o == RB_OPT_ADAM_LAZY ? rb_adam_alpha_t(opt->lr, opt->beta_1, opt->beta_2, opt->step) : 0.f;
  u.alpha_dev = o == RB_OPT_ADAM_LAZY ? opt->alpha_t_dev : nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dout_dtype == RB_F32) {
    RB_CHECK_ARG((reinterpret_cast<uintptr_t>(dOut) & 3) == 0, RB_ERR_ALIGN, "dOut not 4 B aligned");
    return launch_bwd_update<float>(a, D, static_cast<const float*>(dOut), dout_stride, dE, d_dense, u, st);
  }
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(dOut) & 1) == 0, RB_ERR_ALIGN, "dOut not 2 B aligned");
  return launch_bwd_update<__nv_bfloat16>(a, D, static_cast<const __nv_bfloat16*>(dOut), dout_stride, dE, d_dense, u, st);
}

extern "C" int rb_dot_interaction_fwd(const float* E, const float* table, int64_t rows, const void* idx,
                                      int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                      int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                      int32_t tail, void* out, int32_t out_dtype, int64_t out_stride, int32_t row_cache,
                                      const int32_t* row_cache_hint, void* stream) {
  IxArgs a;
  int rc = fill_args(&a, E, table, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D, self_interaction,
                     skip_gather, tail);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(row_cache >= RB_ROW_CACHE_L2 && row_cache <= RB_ROW_CACHE_AUTO, RB_ERR_ARG, "bad row_cache %d", row_cache);
  a.row_cache = row_cache;
  a.row_cache_hint = row_cache_hint;
  if (B == 0) return RB_OK;
  const int total = a.ncols + (a.tail ? D : 0);
  RB_CHECK_ARG(out != nullptr && out_stride >= total, RB_ERR_ARG, "out is null or out_stride too small");
  RB_CHECK_ARG(out_dtype == RB_F32 || out_dtype == RB_BF16 || out_dtype == RB_BF16_ONES, RB_ERR_ARG, "bad out_dtype %d", out_dtype);
  RB_CHECK_ARG(out_dtype != RB_BF16_ONES || out_stride > total, RB_ERR_ARG, "RB_BF16_ONES needs at least one pad column");
  a.pad_one = (out_dtype == RB_BF16_ONES);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == RB_F32) {
    RB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 3) == 0, RB_ERR_ALIGN, "out not 4 B aligned");
    return launch_fwd<float>(a, D, static_cast<float*>(out), out_stride, total, st);
  }
  // bf16: the pad columns [total, out_stride) are written as zeros so a GEMM can consume the padded row
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 1) == 0, RB_ERR_ALIGN, "out not 2 B aligned");
  RB_CHECK_ARG(out_stride - total < 64, RB_ERR_ARG, "bf16 out_stride pads more than 63 columns");
  return launch_fwd<__nv_bfloat16>(a, D, static_cast<__nv_bfloat16*>(out), out_stride, static_cast<int>(out_stride), st);
}

extern "C" int rb_dot_interaction_bwd(const float* E, const float* table, int64_t rows, const void* idx,
                                      int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                      int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                      int32_t tail, const void* dOut, int32_t dout_dtype, int64_t dout_stride, float* dE,
                                      float* d_dense, int32_t row_cache, const int32_t* row_cache_hint, void* stream) {
  IxArgs a;
  int rc = fill_args(&a, E, table, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D, self_interaction,
                     skip_gather, tail);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(row_cache >= RB_ROW_CACHE_L2 && row_cache <= RB_ROW_CACHE_AUTO, RB_ERR_ARG, "bad row_cache %d", row_cache);
  a.row_cache = row_cache;
  a.row_cache_hint = row_cache_hint;
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(dOut != nullptr && dout_stride >= a.ncols + (a.tail ? D : 0), RB_ERR_ARG, "dOut is null or stride too small");
  RB_CHECK_ARG(dout_dtype == RB_F32 || dout_dtype == RB_BF16, RB_ERR_ARG, "bad dout_dtype %d", dout_dtype);
  RB_CHECK_ARG((dE == nullptr || aligned_for(dE, 4)) && (d_dense == nullptr || aligned_for(d_dense, 4)), RB_ERR_ALIGN,
               "dE/d_dense not 16 B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dout_dtype == RB_F32) {
    RB_CHECK_ARG((reinterpret_cast<uintptr_t>(dOut) & 3) == 0, RB_ERR_ALIGN, "dOut not 4 B aligned");
    return launch_bwd<float>(a, D, static_cast<const float*>(dOut), dout_stride, dE, d_dense, st);
  }
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(dOut) & 1) == 0, RB_ERR_ALIGN, "dOut not 2 B aligned");
  return launch_bwd<__nv_bfloat16>(a, D, static_cast<const __nv_bfloat16*>(dOut), dout_stride, dE, d_dense, st);
}

// Row-wise sharded forms: the table is `world` shards in peer memory (shard_ptrs_dev: DEVICE array of
// `world` device pointers, this rank's own shard included); `rows` is the GLOBAL row count.
extern "C" int rb_dot_interaction_fwd_sharded_rep(const void* const* shard_ptrs_dev, int32_t world, int64_t rows, const void* idx,
                                                  int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                                  int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                                  int32_t tail, void* out, int32_t out_dtype, int64_t out_stride, void* x_save,
                                                  const void* const* shadow_ptrs_dev, int64_t small_base, const float* small_rep,
                                                  const void* small_rep_bf16, void* stream);

extern "C" int rb_dot_interaction_fwd_sharded(const void* const* shard_ptrs_dev, int32_t world, int64_t rows, const void* idx,
                                              int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                              int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                              int32_t tail, void* out, int32_t out_dtype, int64_t out_stride, void* x_save,
                                              const void* const* shadow_ptrs_dev, void* stream) {
  return rb_dot_interaction_fwd_sharded_rep(shard_ptrs_dev, world, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D,
                                            self_interaction, skip_gather, tail, out, out_dtype, out_stride, x_save, shadow_ptrs_dev, rows,
                                            nullptr, nullptr, stream);
}

extern "C" int rb_dot_interaction_fwd_sharded_rep(const void* const* shard_ptrs_dev, int32_t world, int64_t rows, const void* idx,
                                                  int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                                  int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                                  int32_t tail, void* out, int32_t out_dtype, int64_t out_stride, void* x_save,
                                                  const void* const* shadow_ptrs_dev, int64_t small_base, const float* small_rep,
                                                  const void* small_rep_bf16, void* stream) {
  RB_CHECK_ARG(shard_ptrs_dev != nullptr, RB_ERR_ARG, "shard_ptrs_dev is null");
  IxArgs a;
  int rc = fill_args(&a, nullptr, nullptr, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D, self_interaction, skip_gather,
                     tail, reinterpret_cast<const float* const*>(shard_ptrs_dev), world);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(small_base >= 0 && small_base <= rows, RB_ERR_ARG, "small_base must lie in [0, rows]");
  if (small_base < rows) {
    const void* rep = shadow_ptrs_dev != nullptr ? small_rep_bf16 : static_cast<const void*>(small_rep);
    RB_CHECK_ARG(rep != nullptr && (reinterpret_cast<uintptr_t>(rep) & 15) == 0, RB_ERR_ARG,
                 "rows >= small_base need the local replica (bf16 with shadow shards, fp32 otherwise), 16 B aligned");
    a.small_base = static_cast<uint32_t>(small_base);
    a.small_rep = rep;
  }
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(x_save == nullptr || (reinterpret_cast<uintptr_t>(x_save) & 15) == 0, RB_ERR_ALIGN, "x_save not 16 B aligned");
  a.x_save = static_cast<__nv_bfloat16*>(x_save);
  const int total = a.ncols + (a.tail ? D : 0);
  RB_CHECK_ARG(out != nullptr && out_stride >= total, RB_ERR_ARG, "out is null or out_stride too small");
  RB_CHECK_ARG(out_dtype == RB_F32 || out_dtype == RB_BF16 || out_dtype == RB_BF16_ONES, RB_ERR_ARG, "bad out_dtype %d", out_dtype);
  RB_CHECK_ARG(out_dtype != RB_BF16_ONES || out_stride > total, RB_ERR_ARG, "RB_BF16_ONES needs at least one pad column");
  a.pad_one = (out_dtype == RB_BF16_ONES);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* const* shadows = reinterpret_cast<const __nv_bfloat16* const*>(shadow_ptrs_dev);
  if (out_dtype == RB_F32) {
    if (shadows != nullptr) return launch_fwd16<float>(a, D, shadows, static_cast<float*>(out), out_stride, total, st);
    return launch_fwd<float>(a, D, static_cast<float*>(out), out_stride, total, st);
  }
  RB_CHECK_ARG(out_stride - total < 64, RB_ERR_ARG, "bf16 out_stride pads more than 63 columns");
  if (shadows != nullptr)
    return launch_fwd16<__nv_bfloat16>(a, D, shadows, static_cast<__nv_bfloat16*>(out), out_stride, static_cast<int>(out_stride), st);
  return launch_fwd<__nv_bfloat16>(a, D, static_cast<__nv_bfloat16*>(out), out_stride, static_cast<int>(out_stride), st);
}

extern "C" int rb_dot_interaction_bwd_sharded_split(const void* const* shard_ptrs_dev, int32_t world, int64_t rows, const void* idx,
                                                    int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                                    int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                                    int32_t tail, const void* dOut, int32_t dout_dtype, int64_t dout_stride, float* dE,
                                                    float* d_dense, const void* x_saved, const int32_t* de_slot, float* dE_small,
                                                    int32_t num_small, void* stream);

extern "C" int rb_dot_interaction_bwd_sharded(const void* const* shard_ptrs_dev, int32_t world, int64_t rows, const void* idx,
                                              int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                              int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                              int32_t tail, const void* dOut, int32_t dout_dtype, int64_t dout_stride, float* dE,
                                              float* d_dense, const void* x_saved, void* stream) {
  return rb_dot_interaction_bwd_sharded_split(shard_ptrs_dev, world, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D,
                                              self_interaction, skip_gather, tail, dOut, dout_dtype, dout_stride, dE, d_dense, x_saved,
                                              nullptr, nullptr, 0, stream);
}

extern "C" int rb_dot_interaction_bwd_sharded_split(const void* const* shard_ptrs_dev, int32_t world, int64_t rows, const void* idx,
                                                    int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                                    int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                                    int32_t tail, const void* dOut, int32_t dout_dtype, int64_t dout_stride, float* dE,
                                                    float* d_dense, const void* x_saved, const int32_t* de_slot, float* dE_small,
                                                    int32_t num_small, void* stream) {
  RB_CHECK_ARG(shard_ptrs_dev != nullptr, RB_ERR_ARG, "shard_ptrs_dev is null");
  RB_CHECK_ARG((de_slot == nullptr) == (dE_small == nullptr) && (de_slot == nullptr || (num_small >= 1 && num_small <= F && aligned_for(dE_small, 4) &&
                                                                                       x_saved != nullptr && dE != nullptr)),
               RB_ERR_ARG, "de_slot and dE_small come together, with 1 <= num_small <= F, saved operand rows and dE");
  IxArgs a;
  int rc = fill_args(&a, nullptr, nullptr, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D, self_interaction, skip_gather,
                     tail, reinterpret_cast<const float* const*>(shard_ptrs_dev), world);
  if (rc != RB_OK) return rc;
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(x_saved == nullptr || (reinterpret_cast<uintptr_t>(x_saved) & 15) == 0, RB_ERR_ALIGN, "x_saved not 16 B aligned");
  a.x_load = static_cast<const __nv_bfloat16*>(x_saved);
  a.de_slot = de_slot;
  a.dE_small = dE_small;
  a.num_small = num_small;
  RB_CHECK_ARG(dOut != nullptr && dout_stride >= a.ncols + (a.tail ? D : 0), RB_ERR_ARG, "dOut is null or stride too small");
  RB_CHECK_ARG(dout_dtype == RB_F32 || dout_dtype == RB_BF16, RB_ERR_ARG, "bad dout_dtype %d", dout_dtype);
  RB_CHECK_ARG((dE == nullptr || aligned_for(dE, 4)) && (d_dense == nullptr || aligned_for(d_dense, 4)), RB_ERR_ALIGN,
               "dE/d_dense not 16 B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dout_dtype == RB_F32) return launch_bwd<float>(a, D, static_cast<const float*>(dOut), dout_stride, dE, d_dense, st);
  return launch_bwd<__nv_bfloat16>(a, D, static_cast<const __nv_bfloat16*>(dOut), dout_stride, dE, d_dense, st);
}
