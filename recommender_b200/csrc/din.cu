// DIN's LocalActivationUnit (dien/layers.py:34-59, called at dien/model.py:42-53) — attention-weighted pooling
// of a behaviour history whose per-position weights come from a small MLP on [t, h, t - h, t * h].
//
// The B200 form never materialises history[B, L, E] nor the 4E-wide feature tensor for the padded positions.  The VALID
// positions (mask != 0, dien/model.py:43) are numbered p = offsets[b] + (rank of l among the sample's valid positions) and
// only they get a feature row:
//
//   rb_din_offsets            offsets[B + 1] = exclusive scan of the per-sample valid counts (device); offsets[B] = P
//   rb_din_build_features     X[p, :] = bf16([t_b | h_p | t_b - h_p | t_b * h_p]), h_p = [W_item[item[b,l]] | W_cat[cat[b,l]]] read
//                             straight from the tables (dien/model.py:14-19 compute_flat_embedding + dien/layers.py:48-49)
//   (three Dense layers 4E -> 80 -> 40 -> 1 on the tcgen05 kernels of mlp.cu, sigmoid / sigmoid / none: dien/layers.py:37-39,50-52)
//   rb_din_pool_fwd           rep[b, :] = sum_p w_p * h_p   (dien/layers.py:53-57: weights *= mask; weights^T . history)
//   rb_din_pool_bwd_weights   dw_p = <d_rep[b], h_p>
//   rb_din_feature_bwd        dh[b, l, :] = dX[p, E:2E] - dX[p, 2E:3E] + dX[p, 3E:4E] * t_b + w_p * d_rep[b]   (the IndexedSlices rows)
//                             dt[b, :]    = sum_p dX[p, 0:E] + dX[p, 2E:3E] + dX[p, 3E:4E] * h_p
//
// One warp per sample; the ids and mask of 32 positions are fetched coalesced and handed round by shuffle; lanes own the
// columns c = lane, lane + 32, ... of the E-wide rows.  All kernels are HBM-bound (rows of E * 4 bytes gathered per valid
// position; X rows of 8 E bytes written once).
#include <cub/cub.cuh>

#include "common.cuh"

namespace rb {
namespace din {

constexpr int kMaxColsPerLane = 4;       // E <= 128
constexpr int kWarpsPerBlock = 8;

struct Tables {
  const float* tab[2];
  int64_t rows[2];
  int D[2];
  const void* idx[2];     // [B, L] ids into tab[k]; idx[1] / tab[1] may be null (D[1] == 0)
  int is64;
  const void* mask;       // [B, L]; nonzero = valid.  null: idx[0] != 0 (keras mask_zero, dien/model.py:11-12,43)
  int mask_is64;
  int64_t B;
  int L, E;
};

__device__ __forceinline__ bool valid_at(const Tables& t, int64_t p) {
  return t.mask != nullptr ? load_raw_index(t.mask, t.mask_is64, p) != 0 : load_raw_index(t.idx[0], t.is64, p) != 0;
}

// column c of the history row at ids (i0, i1); out-of-range ids read as zeros (TF's GPU gather)
__device__ __forceinline__ float hist_col(const Tables& t, int64_t i0, int64_t i1, int c) {
  if (c < t.D[0]) return (i0 >= 0 && i0 < t.rows[0]) ? __ldg(t.tab[0] + i0 * t.D[0] + c) : 0.f;
  return (i1 >= 0 && i1 < t.rows[1]) ? __ldg(t.tab[1] + i1 * t.D[1] + (c - t.D[0])) : 0.f;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) count_valid_kernel(Tables t, int32_t* __restrict__ counts) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  int n = 0;
  for (int l0 = 0; l0 < t.L; l0 += 32) {
    const int l = l0 + lane;
    const bool v = l < t.L && valid_at(t, b * t.L + l);
    n += __popc(__ballot_sync(0xffffffffu, v));
  }
  if (lane == 0) counts[b] = n;
}

// Walks the valid positions of sample b in order of l and calls f(l, p, i0, i1) with all 32 lanes converged.
template <class F>
__device__ __forceinline__ void for_valid_positions(const Tables& t, int64_t b, int lane, int32_t p0, F&& f) {
  int32_t p = p0;
  for (int l0 = 0; l0 < t.L; l0 += 32) {
    const int l = l0 + lane;
    const bool in = l < t.L;
    const int64_t q = b * t.L + l;
    const bool v = in && valid_at(t, q);
    const int64_t my0 = in ? load_raw_index(t.idx[0], t.is64, q) : 0;
    const int64_t my1 = (in && t.idx[1] != nullptr) ? load_raw_index(t.idx[1], t.is64, q) : 0;
    unsigned m = __ballot_sync(0xffffffffu, v);
    while (m != 0) {
      const int k = __ffs(m) - 1;
      m &= m - 1;
      const int64_t i0 = __shfl_sync(0xffffffffu, my0, k);
      const int64_t i1 = __shfl_sync(0xffffffffu, my1, k);
      f(l0 + k, p, i0, i1);
      ++p;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
build_features_kernel(Tables t, const float* __restrict__ target, const int32_t* __restrict__ offsets, __nv_bfloat16* __restrict__ X,
                      int64_t ldx) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float tv[kMaxColsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) {
    const int c = lane + 32 * j;
    tv[j] = c < E ? __ldg(target + b * E + c) : 0.f;
  }
  const int pad = static_cast<int>(ldx) - 4 * E;
  for_valid_positions(t, b, lane, offsets[b], [&](int, int32_t p, int64_t i0, int64_t i1) {
    __nv_bfloat16* row = X + static_cast<int64_t>(p) * ldx;
    float hv[kMaxColsPerLane];
#pragma unroll
    for (int j = 0; j < kMaxColsPerLane; ++j) {
      const int c = lane + 32 * j;
      hv[j] = c < E ? hist_col(t, i0, i1, c) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kMaxColsPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < E) {
        row[c] = __float2bfloat16_rn(tv[j]);
        row[E + c] = __float2bfloat16_rn(hv[j]);
        row[2 * E + c] = __float2bfloat16_rn(__fsub_rn(tv[j], hv[j]));
        row[3 * E + c] = __float2bfloat16_rn(__fmul_rn(tv[j], hv[j]));
      }
    }
    if (lane < pad) row[4 * E + lane] = __float2bfloat16_rn(0.f);
  });
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pool_fwd_kernel(Tables t, const int32_t* __restrict__ offsets, const float* __restrict__ w, float* __restrict__ rep) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float acc[kMaxColsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) acc[j] = 0.f;
  for_valid_positions(t, b, lane, offsets[b], [&](int, int32_t p, int64_t i0, int64_t i1) {
    const float wp = __ldg(w + p);
#pragma unroll
    for (int j = 0; j < kMaxColsPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < E) acc[j] = __fadd_rn(acc[j], __fmul_rn(wp, hist_col(t, i0, i1, c)));      // position order, explicit rounding
    }
  });
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < E) rep[b * E + c] = acc[j];
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pool_bwd_weights_kernel(Tables t, const int32_t* __restrict__ offsets, const float* __restrict__ d_rep, float* __restrict__ dw) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float g[kMaxColsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) {
    const int c = lane + 32 * j;
    g[j] = c < E ? __ldg(d_rep + b * E + c) : 0.f;
  }
  for_valid_positions(t, b, lane, offsets[b], [&](int, int32_t p, int64_t i0, int64_t i1) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxColsPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < E) s = fmaf(g[j], hist_col(t, i0, i1, c), s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);      // fixed tree: deterministic
    if (lane == 0) dw[p] = s;
  });
}

__device__ __forceinline__ float bf16_at(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
feature_bwd_kernel(Tables t, const float* __restrict__ target, const int32_t* __restrict__ offsets, const __nv_bfloat16* __restrict__ dX,
                   int64_t ldx, const float* __restrict__ w, const float* __restrict__ d_rep, float* __restrict__ dh,
                   float* __restrict__ d_target, int zero_masked) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float tv[kMaxColsPerLane], gv[kMaxColsPerLane], dt[kMaxColsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) {
    const int c = lane + 32 * j;
    tv[j] = c < E ? __ldg(target + b * E + c) : 0.f;
    gv[j] = c < E ? __ldg(d_rep + b * E + c) : 0.f;
    dt[j] = 0.f;
  }
  if (zero_masked) {      // a dense history gradient (materialised-history form): masked positions get explicit zeros
    for (int l = 0; l < t.L; ++l) {
      if (!valid_at(t, b * t.L + l)) {
#pragma unroll
        for (int j = 0; j < kMaxColsPerLane; ++j) {
          const int c = lane + 32 * j;
          if (c < E) dh[(b * t.L + l) * E + c] = 0.f;
        }
      }
    }
  }
  for_valid_positions(t, b, lane, offsets[b], [&](int l, int32_t p, int64_t i0, int64_t i1) {
    const __nv_bfloat16* row = dX + static_cast<int64_t>(p) * ldx;
    const float wp = __ldg(w + p);
    float* out = dh + (b * t.L + l) * E;
#pragma unroll
    for (int j = 0; j < kMaxColsPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < E) {
        const float h = hist_col(t, i0, i1, c);
        const float d0 = bf16_at(row + c), d1 = bf16_at(row + E + c), d2 = bf16_at(row + 2 * E + c), d3 = bf16_at(row + 3 * E + c);
        // every op rounded explicitly, in this order: the oracle (oracle/ctr_oracle.py local_activation_unit_backward) does the same
        out[c] = __fadd_rn(__fadd_rn(__fsub_rn(d1, d2), __fmul_rn(d3, tv[j])), __fmul_rn(wp, gv[j]));
        dt[j] = __fadd_rn(dt[j], __fadd_rn(__fadd_rn(d0, d2), __fmul_rn(d3, h)));
      }
    }
  });
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < E) d_target[b * E + c] = dt[j];
  }
}

static int fill(Tables* t, const rb_din_history* h) {
  RB_CHECK_ARG(h != nullptr, RB_ERR_ARG, "history description is null");
  RB_CHECK_ARG(h->B >= 0 && h->L > 0, RB_ERR_ARG, "bad B / L");
  RB_CHECK_ARG(h->table0 != nullptr && h->idx0 != nullptr && h->D0 > 0 && h->rows0 > 0, RB_ERR_ARG, "the first table, its ids and D0 > 0 are required");
  RB_CHECK_ARG((h->D1 == 0) == (h->table1 == nullptr) && (h->D1 == 0 || (h->idx1 != nullptr && h->rows1 > 0)), RB_ERR_ARG,
               "second table: give table1, idx1, rows1 and D1 > 0, or none of them");
  RB_CHECK_ARG(h->idx_type == RB_I32 || h->idx_type == RB_I64, RB_ERR_ARG, "bad index type");
  RB_CHECK_ARG(h->mask == nullptr || h->mask_type == RB_I32 || h->mask_type == RB_I64, RB_ERR_ARG, "bad mask type");
  RB_CHECK_ARG(h->D0 + h->D1 <= 32 * kMaxColsPerLane, RB_ERR_SHAPE, "history rows of at most %d columns, got %d", 32 * kMaxColsPerLane,
               h->D0 + h->D1);
  RB_CHECK_ARG(h->B * h->L < 0x7FFFFFFFll, RB_ERR_ARG, "B * L must stay below 2^31");
  t->tab[0] = h->table0;
  t->tab[1] = h->table1;
  t->rows[0] = h->rows0;
  t->rows[1] = h->rows1;
  t->D[0] = h->D0;
  t->D[1] = h->D1;
  t->idx[0] = h->idx0;
  t->idx[1] = h->idx1;
  t->is64 = h->idx_type == RB_I64;
  t->mask = h->mask;
  t->mask_is64 = h->mask_type == RB_I64;
  t->B = h->B;
  t->L = h->L;
  t->E = h->D0 + h->D1;
  return RB_OK;
}

}  // namespace din
}  // namespace rb

using namespace rb;
using namespace rb::din;

extern "C" size_t rb_din_workspace_bytes(int64_t B) {
  if (B <= 0 || B >= 0x7FFFFFFFll) return 0;
  size_t scan = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), static_cast<int>(B));
  return ((static_cast<size_t>(B) * 4 + 255) & ~static_cast<size_t>(255)) + scan + 256;
}

extern "C" int rb_din_offsets(const rb_din_history* h, int32_t* offsets, void* ws, size_t ws_bytes, void* stream) {
  Tables t;
  int rc = fill(&t, h);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(offsets != nullptr, RB_ERR_ARG, "offsets is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RB_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int32_t), st));
  if (t.B == 0) return RB_OK;
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= rb_din_workspace_bytes(t.B) && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, RB_ERR_WORKSPACE,
               "workspace too small or not 256 B aligned: need %zu bytes", rb_din_workspace_bytes(t.B));
  int32_t* counts = static_cast<int32_t*>(ws);
  const size_t counts_bytes = (static_cast<size_t>(t.B) * 4 + 255) & ~static_cast<size_t>(255);
  count_valid_kernel<<<grid_for(t.B, kWarpsPerBlock), kWarpsPerBlock * 32, 0, st>>>(t, counts);
  RB_LAUNCH_CHECK("din count_valid_kernel");
  size_t scan = ws_bytes - counts_bytes;
  RB_CUDA(cub::DeviceScan::InclusiveSum(static_cast<unsigned char*>(ws) + counts_bytes, scan, counts, offsets + 1, static_cast<int>(t.B), st));
  return RB_OK;
}

extern "C" int rb_din_build_features(const rb_din_history* h, const float* target, const int32_t* offsets, void* x_bf16, int64_t ldx,
                                     void* stream) {
  Tables t;
  int rc = fill(&t, h);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(target != nullptr && offsets != nullptr && x_bf16 != nullptr, RB_ERR_ARG, "target / offsets / x is null");
  RB_CHECK_ARG(ldx >= 4 * t.E && ldx - 4 * t.E < 32, RB_ERR_ARG, "ldx must be 4E plus fewer than 32 pad columns");
  if (t.B == 0) return RB_OK;
  build_features_kernel<<<grid_for(t.B, kWarpsPerBlock), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      t, target, offsets, static_cast<__nv_bfloat16*>(x_bf16), ldx);
  RB_LAUNCH_CHECK("din build_features_kernel");
  return RB_OK;
}

extern "C" int rb_din_pool_fwd(const rb_din_history* h, const int32_t* offsets, const float* w, float* rep, void* stream) {
  Tables t;
  int rc = fill(&t, h);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(offsets != nullptr && w != nullptr && rep != nullptr, RB_ERR_ARG, "offsets / w / rep is null");
  if (t.B == 0) return RB_OK;
  pool_fwd_kernel<<<grid_for(t.B, kWarpsPerBlock), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(t, offsets, w, rep);
  RB_LAUNCH_CHECK("din pool_fwd_kernel");
  return RB_OK;
}

extern "C" int rb_din_pool_bwd_weights(const rb_din_history* h, const int32_t* offsets, const float* d_rep, float* dw, void* stream) {
  Tables t;
  int rc = fill(&t, h);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(offsets != nullptr && d_rep != nullptr && dw != nullptr, RB_ERR_ARG, "offsets / d_rep / dw is null");
  if (t.B == 0) return RB_OK;
  pool_bwd_weights_kernel<<<grid_for(t.B, kWarpsPerBlock), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(t, offsets, d_rep, dw);
  RB_LAUNCH_CHECK("din pool_bwd_weights_kernel");
  return RB_OK;
}

extern "C" int rb_din_feature_bwd(const rb_din_history* h, const float* target, const int32_t* offsets, const void* dx_bf16, int64_t ldx,
                                  const float* w, const float* d_rep, float* dh, float* d_target, int32_t zero_masked, void* stream) {
  Tables t;
  int rc = fill(&t, h);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(target != nullptr && offsets != nullptr && dx_bf16 != nullptr && w != nullptr && d_rep != nullptr && dh != nullptr &&
                   d_target != nullptr,
               RB_ERR_ARG, "a required pointer is null");
  RB_CHECK_ARG(ldx >= 4 * t.E, RB_ERR_ARG, "ldx smaller than 4E");
  if (t.B == 0) return RB_OK;
  feature_bwd_kernel<<<grid_for(t.B, kWarpsPerBlock), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      t, target, offsets, static_cast<const __nv_bfloat16*>(dx_bf16), ldx, w, d_rep, dh, d_target, zero_masked);
  RB_LAUNCH_CHECK("din feature_bwd_kernel");
  return RB_OK;
}
