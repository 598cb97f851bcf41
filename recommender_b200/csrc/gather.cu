// K1 / K2 / K11: embedding lookup, bag pooling and the DeepFM front end (sm_100a).
//
// HBM-bound kernels.  A row of D floats is moved by a GS-lane group with 128-bit (or 64/32-bit)
// accesses, so a warp request covers whole 32 B sectors of each row.  The un-pooled gather
// stages its index block in shared memory with one bulk async copy (cp.async.bulk + mbarrier,
// the 1-D TMA path: SASS UBLKCP) so that the dependent row loads are issued from on-chip data.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rb {

// The DeepFM deep input row (ctr/model.py:25-26) as the bf16 K operand of the MLP's first Dense layer:
// [flatten(E) | int_features | 1.0 | 0 ...] of `ld` columns, written by the kernel that gathers E.
struct DeepRow {
  __nv_bfloat16* out;      // [B, ld] or null
  const float* dense;      // [B, num_dense] (row stride dense_ld)
  int num_dense, dense_ld, ld;
};

// ---- mbarrier / bulk-copy PTX (1-D TMA) ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ int64_t map_raw(const IndexMap& m, int64_t id, int64_t p) {
  if (m.hash_mod > 0) id = static_cast<int64_t>(static_cast<uint64_t>(id) % static_cast<uint64_t>(m.hash_mod));
  if (m.field_row_offset != nullptr) id += __ldg(m.field_row_offset + (p % m.L));
  return (id >= 0 && id < m.rows) ? id : -1;
}

// ---- un-pooled gather ------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kGatherChunk = 1024;  // lookups per CTA; index block = 4 or 8 KiB
constexpr int kGatherUnroll = 4;    // independent row loads in flight per group

template <int VEC, int GS>
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(const float* __restrict__ table, IndexMap m, int64_t n, int D, float* __restrict__ out,
                   int64_t out_stride, int* __restrict__ oob_flag) {
  __shared__ __align__(16) unsigned char s_idx[kGatherChunk * 8];
  __shared__ __align__(8) uint64_t s_bar;

  const int64_t base = static_cast<int64_t>(blockIdx.x) * kGatherChunk;
  const int64_t left = n - base;
  const int cnt = left < kGatherChunk ? static_cast<int>(left) : kGatherChunk;
  const int esz = m.is64 ? 8 : 4;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(m.idx) + base * esz;
  const uint32_t bytes = static_cast<uint32_t>(cnt) * esz;
  // bulk copy needs 16 B aligned source and a multiple of 16 B; otherwise (ragged tail, sliced
  // tensor) fall back to plain coalesced loads of the index block.
  const bool bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15) == 0);
  if (bulk) {
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&s_bar, bytes);
      bulk_g2s(s_idx, src, bytes, &s_bar);
    }
    mbar_wait(&s_bar, 0);
  } else {
    for (int i = threadIdx.x; i < cnt; i += kGatherThreads) {
      if (m.is64) reinterpret_cast<int64_t*>(s_idx)[i] = __ldg(reinterpret_cast<const int64_t*>(src) + i);
      else reinterpret_cast<int32_t*>(s_idx)[i] = __ldg(reinterpret_cast<const int32_t*>(src) + i);
    }
    __syncthreads();
  }

  constexpr int kGroups = kGatherThreads / GS;
  const int group = threadIdx.x / GS;
  const int lane = threadIdx.x % GS;
  const bool active = lane * VEC < D;
  bool oob = false;

  for (int i0 = group * kGatherUnroll; i0 < cnt; i0 += kGroups * kGatherUnroll) {
    int64_t row[kGatherUnroll];
    Row<VEC> val[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const int i = i0 + u;
      row[u] = -2;
      if (i < cnt) {
        const int64_t raw = m.is64 ? reinterpret_cast<const int64_t*>(s_idx)[i]
                                   : static_cast<int64_t>(reinterpret_cast<const int32_t*>(s_idx)[i]);
        row[u] = map_raw(m, raw, base + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      val[u] = zero_row<VEC>();
      if (row[u] >= 0 && active) val[u] = ld_row<VEC>(table + row[u] * D + lane * VEC);
      oob |= (row[u] == -1);
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      if (row[u] != -2 && active) st_row_stream<VEC>(out + (base + i0 + u) * out_stride + lane * VEC, val[u]);
    }
  }
  if (oob && oob_flag != nullptr) *oob_flag = 1;
}

// ---- pooled lookup (sum / mean / masked mean) --------------------------------------------------
constexpr int kPoolThreads = 256;
constexpr int kPoolUnroll = 4;

template <int VEC, int GS>
__global__ void __launch_bounds__(kPoolThreads)
bag_pool_kernel(const float* __restrict__ table, IndexMap m, int64_t B, int L, int D, int pool_mode,
                const void* __restrict__ mask_idx, float* __restrict__ out, int64_t out_stride,
                float* __restrict__ count_out, int* __restrict__ oob_flag) {
  const int64_t bag = static_cast<int64_t>(blockIdx.x) * (kPoolThreads / GS) + threadIdx.x / GS;
  const int lane = threadIdx.x % GS;
  const bool bag_ok = bag < B;
  const bool active = bag_ok && lane * VEC < D;
  const int64_t p0 = (bag_ok ? bag : 0) * L;
  const bool masked = (pool_mode == RB_POOL_MASKED_MEAN);
  const void* mk = (mask_idx != nullptr) ? mask_idx : m.idx;

  Row<VEC> acc = zero_row<VEC>();
  int valid = 0;
  bool oob = false;
  // The group's lanes fetch GS consecutive indices at once (coalesced), then every lane walks
  // them via shuffles; kPoolUnroll independent row loads are in flight per group.
  for (int l0 = 0; l0 < L; l0 += GS) {
    const int l = l0 + lane;
    int64_t my_row = -2;  // -2: skip (past L or masked)
    if (bag_ok && l < L) {
      bool keep = true;
      if (masked) keep = load_raw_index(mk, m.is64, p0 + l) != 0;
      if (keep) my_row = map_index(m, p0 + l);
    }
    const int steps = min(GS, L - l0);
    for (int j0 = 0; j0 < steps; j0 += kPoolUnroll) {
      int64_t row[kPoolUnroll];
      Row<VEC> val[kPoolUnroll];
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        // all 32 lanes take part in the shuffle; (j0+u) may run past `steps` -> masked below
        row[u] = __shfl_sync(0xffffffffu, my_row, (j0 + u) % GS, GS);
        if (j0 + u >= steps) row[u] = -2;
      }
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        val[u] = zero_row<VEC>();
        if (row[u] >= 0 && active) val[u] = ld_row<VEC>(table + row[u] * D + lane * VEC);
      }
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        if (row[u] != -2) {
          ++valid;
          oob |= (row[u] == -1);
#pragma unroll
          for (int k = 0; k < VEC; ++k) acc.v[k] += val[u].v[k];  // in position order
        }
      }
    }
  }
  if (active) {
    float denom = 1.f;
    if (pool_mode == RB_POOL_MEAN) denom = static_cast<float>(L);
    if (masked) denom = static_cast<float>(valid);
    if (pool_mode != RB_POOL_SUM) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc.v[k] = __fdiv_rn(acc.v[k], denom);  // 0/0 -> NaN like the reference
    }
    st_row<VEC>(out + bag * out_stride + lane * VEC, acc);
  }
  if (bag_ok && lane == 0 && count_out != nullptr) count_out[bag] = static_cast<float>(valid);
  if (oob && oob_flag != nullptr) *oob_flag = 1;
}

// ---- DeepFM front end: gather + FM second order in one pass -----------------------------------------
template <int VEC, int GS>
__global__ void __launch_bounds__(kPoolThreads)
gather_fm_kernel(const float* __restrict__ table, IndexMap m, int64_t B, int F, int D, float* __restrict__ E,
                 float* __restrict__ s_out, float* __restrict__ fm_out, int* __restrict__ oob_flag, DeepRow deep) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * (kPoolThreads / GS) + threadIdx.x / GS;
  const int lane = threadIdx.x % GS;
  const bool b_ok = b < B;
  const bool active = b_ok && lane * VEC < D;
  const int64_t p0 = (b_ok ? b : 0) * F;

  Row<VEC> s = zero_row<VEC>();
  Row<VEC> q = zero_row<VEC>();
  bool oob = false;
  for (int f0 = 0; f0 < F; f0 += GS) {
    const int f = f0 + lane;
    int64_t my_row = -2;
    if (b_ok && f < F) my_row = map_index(m, p0 + f);
    const int steps = min(GS, F - f0);
    for (int j0 = 0; j0 < steps; j0 += kPoolUnroll) {
      int64_t row[kPoolUnroll];
      Row<VEC> val[kPoolUnroll];
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        row[u] = __shfl_sync(0xffffffffu, my_row, (j0 + u) % GS, GS);
        if (j0 + u >= steps) row[u] = -2;
      }
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        val[u] = zero_row<VEC>();
        if (row[u] >= 0 && active) val[u] = ld_row<VEC>(table + row[u] * D + lane * VEC);
      }
#pragma unroll
      for (int u = 0; u < kPoolUnroll; ++u) {
        if (row[u] != -2) {
          oob |= (row[u] == -1);
          if (active && E != nullptr) st_row_stream<VEC>(E + (p0 + f0 + j0 + u) * D + lane * VEC, val[u]);
          if (active && deep.out != nullptr) {
            __nv_bfloat16* dst = deep.out + b * deep.ld + (f0 + j0 + u) * D + lane * VEC;
#pragma unroll
            for (int k = 0; k < VEC; ++k) dst[k] = __float2bfloat16_rn(val[u].v[k]);
          }
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            s.v[k] += val[u].v[k];
            q.v[k] = fmaf(val[u].v[k], val[u].v[k], q.v[k]);
          }
        }
      }
    }
  }
  // fm = 0.5 * sum_d (s^2 - q)   (ctr/model.py:21-23)
  float part = 0.f;
  if (active) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) part += s.v[k] * s.v[k] - q.v[k];
    if (s_out != nullptr) st_row<VEC>(s_out + b * D + lane * VEC, s);
  }
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o, GS);
  if (b_ok && lane == 0 && fm_out != nullptr) fm_out[b] = 0.5f * part;
  if (b_ok && deep.out != nullptr) {      // the tail of the deep row: dense features, the ones column, zero pads
    const int c0 = F * D;
    for (int c = c0 + lane; c < deep.ld; c += GS) {
      const int j = c - c0;
      const float v = j < deep.num_dense ? __ldg(deep.dense + b * deep.dense_ld + j) : (j == deep.num_dense ? 1.0f : 0.f);
      deep.out[b * deep.ld + c] = __float2bfloat16_rn(v);
    }
  }
  if (oob && oob_flag != nullptr) *oob_flag = 1;
}

// ---- id -> row map -----------------------------------------------------------------------------
__global__ void hash_ids_kernel(const void* __restrict__ ids, int is64, int64_t n, int64_t vocab, int world,
                                int64_t* __restrict__ rows_out, int32_t* __restrict__ owner_out,
                                int64_t* __restrict__ local_out) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint64_t id = static_cast<uint64_t>(load_raw_index(ids, is64, p));
  const int64_t row = static_cast<int64_t>(id % static_cast<uint64_t>(vocab));
  if (rows_out != nullptr) rows_out[p] = row;
  if (world > 1) {
    if (owner_out != nullptr) owner_out[p] = static_cast<int32_t>(row % world);
    if (local_out != nullptr) local_out[p] = row / world;
  } else {
    if (owner_out != nullptr) owner_out[p] = 0;
    if (local_out != nullptr) local_out[p] = row;
  }
}

// ---- dispatch -------------------------------------------------------------------------------------
#define RB_DISPATCH_GEOM(geom, CALL)                                  \
  do {                                                                \
    if (geom.vec == 4) {                                              \
      if (geom.gs == 4) { CALL(4, 4); }                               \
      else if (geom.gs == 8) { CALL(4, 8); }                          \
      else if (geom.gs == 16) { CALL(4, 16); }                        \
      else { CALL(4, 32); }                                           \
    } else if (geom.vec == 2) {                                       \
      if (geom.gs == 4) { CALL(2, 4); }                               \
      else if (geom.gs == 8) { CALL(2, 8); }                          \
      else if (geom.gs == 16) { CALL(2, 16); }                        \
      else { CALL(2, 32); }                                           \
    } else {                                                          \
      if (geom.gs == 4) { CALL(1, 4); }                               \
      else if (geom.gs == 8) { CALL(1, 8); }                          \
      else if (geom.gs == 16) { CALL(1, 16); }                        \
      else { CALL(1, 32); }                                           \
    }                                                                 \
  } while (0)

static int check_table(const float* table, int64_t rows, int D, RowGeom* g) {
  RB_CHECK_ARG(table != nullptr && rows > 0, RB_ERR_ARG, "table is null or has no rows");
  RB_CHECK_ARG(row_geom(D, g), RB_ERR_SHAPE, "unsupported embedding dim D=%d (need D/vec <= 32)", D);
  RB_CHECK_ARG(aligned_for(table, g->vec), RB_ERR_ALIGN, "table pointer not aligned to %d bytes", g->vec * 4);
  return RB_OK;
}

}  // namespace rb

using namespace rb;

// Per-table id validation (TF's CPU Gather rejects an id outside [0, input_dim) of ITS table; the lookups above range-check
// the final row against the total row count only, so with T tables stored back to back an id >= rows(f) would land in
// table f+1).  One coalesced pass over the ids; any violation raises the flag.
__global__ void check_ids_kernel(const void* __restrict__ idx, int is64, int64_t n, int L, const int64_t* __restrict__ field_rows,
                                 int64_t hash_mod, int32_t* __restrict__ oob_flag) {
  bool bad = false;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < n;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t id = load_raw_index(idx, is64, p);
    if (hash_mod > 0) id = static_cast<int64_t>(static_cast<uint64_t>(id) % static_cast<uint64_t>(hash_mod));
    bad |= id < 0 || id >= __ldg(field_rows + (p % L));
  }
  if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) *oob_flag = 1;
}

extern "C" int rb_gather_fwd(const float* table, int64_t rows, int32_t D, const void* idx, int32_t idx_type,
                             int64_t n, int32_t L, const int64_t* field_row_offset, int64_t hash_mod, float* out,
                             int64_t out_stride, int32_t* oob_flag, void* stream) {
  RowGeom g;
  int rc = check_table(table, rows, D, &g);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(n >= 0 && (idx_type == RB_I32 || idx_type == RB_I64), RB_ERR_ARG, "bad n or index type");
  if (n == 0) return RB_OK;
  RB_CHECK_ARG(idx != nullptr && out != nullptr, RB_ERR_ARG, "idx/out is null");
  RB_CHECK_ARG(out_stride >= D && (out_stride % g.vec) == 0 && aligned_for(out, g.vec), RB_ERR_ALIGN,
               "out stride/pointer not aligned for vec=%d", g.vec);
  IndexMap m = make_index_map(idx, idx_type, field_row_offset, hash_mod, rows, L);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int grid = grid_for(n, kGatherChunk);
#define CALL(V, G) gather_rows_kernel<V, G><<<grid, kGatherThreads, 0, st>>>(table, m, n, D, out, out_stride, oob_flag)
  RB_DISPATCH_GEOM(g, CALL);
#undef CALL
  RB_LAUNCH_CHECK("gather_rows_kernel");
  return RB_OK;
}

extern "C" int rb_bag_pool_fwd(const float* table, int64_t rows, int32_t D, const void* idx, int32_t idx_type,
                               int64_t B, int32_t L, const int64_t* field_row_offset, int64_t hash_mod,
                               int32_t pool_mode, const void* mask_idx, float* out, int64_t out_stride,
                               float* count_out, int32_t* oob_flag, void* stream) {
  RowGeom g;
  int rc = check_table(table, rows, D, &g);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(B >= 0 && L > 0 && (idx_type == RB_I32 || idx_type == RB_I64), RB_ERR_ARG, "bad B/L or index type");
  RB_CHECK_ARG(pool_mode >= RB_POOL_SUM && pool_mode <= RB_POOL_MASKED_MEAN, RB_ERR_ARG, "bad pool mode %d", pool_mode);
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(idx != nullptr && out != nullptr, RB_ERR_ARG, "idx/out is null");
  RB_CHECK_ARG(out_stride >= D && (out_stride % g.vec) == 0 && aligned_for(out, g.vec), RB_ERR_ALIGN,
               "out stride/pointer not aligned for vec=%d", g.vec);
  IndexMap m = make_index_map(idx, idx_type, field_row_offset, hash_mod, rows, L);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CALL(V, G)                                                                                              \
  bag_pool_kernel<V, G><<<grid_for(B, kPoolThreads / G), kPoolThreads, 0, st>>>(table, m, B, L, D, pool_mode,  \
                                                                                mask_idx, out, out_stride,    \
                                                                                count_out, oob_flag)
  RB_DISPATCH_GEOM(g, CALL);
#undef CALL
  RB_LAUNCH_CHECK("bag_pool_kernel");
  return RB_OK;
}

extern "C" int rb_gather_fm_fwd(const float* table, int64_t rows, int32_t D, const void* idx, int32_t idx_type,
                                int64_t B, int32_t F, const int64_t* field_row_offset, int64_t hash_mod, float* E,
                                float* s, float* fm, int32_t* oob_flag, void* stream) {
  RowGeom g;
  int rc = check_table(table, rows, D, &g);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(B >= 0 && F > 0 && (idx_type == RB_I32 || idx_type == RB_I64), RB_ERR_ARG, "bad B/F or index type");
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(idx != nullptr, RB_ERR_ARG, "idx is null");
  RB_CHECK_ARG((E == nullptr || aligned_for(E, g.vec)) && (s == nullptr || aligned_for(s, g.vec)), RB_ERR_ALIGN,
               "E/s pointer not aligned for vec=%d", g.vec);
  IndexMap m = make_index_map(idx, idx_type, field_row_offset, hash_mod, rows, F);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeepRow deep{nullptr, nullptr, 0, 0, 0};
#define CALL(V, G) \
  gather_fm_kernel<V, G><<<grid_for(B, kPoolThreads / G), kPoolThreads, 0, st>>>(table, m, B, F, D, E, s, fm, oob_flag, deep)
  RB_DISPATCH_GEOM(g, CALL);
#undef CALL
  RB_LAUNCH_CHECK("gather_fm_kernel");
  return RB_OK;
}

extern "C" int rb_gather_fm_deep_fwd(const float* table, int64_t rows, int32_t D, const void* idx, int32_t idx_type, int64_t B, int32_t F,
                                     const int64_t* field_row_offset, int64_t hash_mod, const float* dense, int32_t num_dense,
                                     int64_t dense_ld, void* deep_bf16, int32_t ld_deep, float* E, float* s, float* fm, int32_t* oob_flag,
                                     void* stream) {
  RowGeom g;
  int rc = check_table(table, rows, D, &g);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(B >= 0 && F > 0 && (idx_type == RB_I32 || idx_type == RB_I64), RB_ERR_ARG, "bad B/F or index type");
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(idx != nullptr && deep_bf16 != nullptr, RB_ERR_ARG, "idx / deep row is null");
  RB_CHECK_ARG(num_dense >= 0 && (num_dense == 0 || (dense != nullptr && dense_ld >= num_dense)), RB_ERR_ARG, "bad dense features");
  RB_CHECK_ARG(ld_deep >= F * D + num_dense, RB_ERR_SHAPE, "deep row of %d columns cannot hold %d x %d + %d values", ld_deep, F, D, num_dense);
  RB_CHECK_ARG((E == nullptr || aligned_for(E, g.vec)) && (s == nullptr || aligned_for(s, g.vec)), RB_ERR_ALIGN,
               "E/s pointer not aligned for vec=%d", g.vec);
  IndexMap m = make_index_map(idx, idx_type, field_row_offset, hash_mod, rows, F);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeepRow deep{static_cast<__nv_bfloat16*>(deep_bf16), dense, num_dense, static_cast<int>(dense_ld), ld_deep};
#define CALL(V, G) \
  gather_fm_kernel<V, G><<<grid_for(B, kPoolThreads / G), kPoolThreads, 0, st>>>(table, m, B, F, D, E, s, fm, oob_flag, deep)
  RB_DISPATCH_GEOM(g, CALL);
#undef CALL
  RB_LAUNCH_CHECK("gather_fm_kernel");
  return RB_OK;
}

extern "C" int rb_hash_ids(const void* ids, int32_t idx_type, int64_t n, int64_t vocab, int32_t world,
                           int64_t* rows_out, int32_t* owner_out, int64_t* local_out, void* stream) {
  RB_CHECK_ARG(n >= 0 && vocab > 0 && (idx_type == RB_I32 || idx_type == RB_I64), RB_ERR_ARG, "bad n/vocab/index type");
  if (n == 0) return RB_OK;
  RB_CHECK_ARG(ids != nullptr, RB_ERR_ARG, "ids is null");
  hash_ids_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, idx_type == RB_I64, n, vocab,
                                                                                  world, rows_out, owner_out, local_out);
  RB_LAUNCH_CHECK("hash_ids_kernel");
  return RB_OK;
}

extern "C" int rb_check_indices(const void* idx, int32_t idx_type, int64_t n, int32_t L, const int64_t* field_rows,
                                int64_t hash_mod, int32_t* oob_flag, void* stream) {
  RB_CHECK_ARG(n >= 0 && L > 0 && (idx_type == RB_I32 || idx_type == RB_I64), RB_ERR_ARG, "bad n / L or index type");
  if (n == 0) return RB_OK;
  RB_CHECK_ARG(idx != nullptr && field_rows != nullptr && oob_flag != nullptr, RB_ERR_ARG, "idx / field_rows / oob_flag is null");
  RB_CHECK_ARG(n % L == 0, RB_ERR_SHAPE, "n is not a multiple of L");
  const unsigned int grid = static_cast<unsigned int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  check_ids_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, idx_type == RB_I64, n, L, field_rows, hash_mod, oob_flag);
  RB_LAUNCH_CHECK("check_ids_kernel");
  return RB_OK;
}
