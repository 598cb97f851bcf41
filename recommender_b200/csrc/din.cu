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

// VEC consecutive columns starting at c of the history row at ids (i0, i1); out-of-range ids read as zeros (TF's GPU gather,
// SURVEY A.6).  VEC == 2 needs even D[0], D[1] (8-byte aligned pairs that never straddle the two tables).
template <int VEC, int COLS>
struct HRow {
  float v[COLS][VEC];
};
constexpr int kUnroll = 4;      // history rows per batch and warp, two batches in flight (8 per batch: fewer resident warps, slower — r2_51)

template <int VEC>
__device__ __forceinline__ void hist_cols(const Tables& t, int64_t i0, int64_t i1, int c, float (&out)[VEC]) {
  const float* src = nullptr;
  if (c < t.D[0]) {
    if (i0 >= 0 && i0 < t.rows[0]) src = t.tab[0] + i0 * t.D[0] + c;
  } else if (i1 >= 0 && i1 < t.rows[1]) {
    src = t.tab[1] + i1 * t.D[1] + (c - t.D[0]);
  }
  if (src == nullptr) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) out[k] = 0.f;
  } else if constexpr (VEC == 2) {
    const float2 x = __ldg(reinterpret_cast<const float2*>(src));
    out[0] = x.x;
    out[1] = x.y;
  } else {
    out[0] = __ldg(src);
  }
}

template <int VEC, int COLS>
__device__ __forceinline__ HRow<VEC, COLS> load_hist(const Tables& t, int64_t i0, int64_t i1, int lane) {
  HRow<VEC, COLS> h;
#pragma unroll
  for (int j = 0; j < COLS; ++j) {
    const int c = (lane + 32 * j) * VEC;
    if (c < t.E) hist_cols<VEC>(t, i0, i1, c, h.v[j]);
    else
#pragma unroll
      for (int k = 0; k < VEC; ++k) h.v[j][k] = 0.f;
  }
  return h;
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

// Walks the valid positions of sample b in order of l and calls f(l, p, h) with all 32 lanes converged, h = this lane's
// columns of the position's history row.  The rows of kUnroll positions are requested together before the first is used.
template <int VEC, int COLS, class F>
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
    // batches of kUnroll positions, double-buffered: the rows of the NEXT batch are requested before this batch is used
    int ks[kUnroll], kn[kUnroll];
    HRow<VEC, COLS> rows[kUnroll], rows_n[kUnroll];
    auto fetch = [&](int (&kk)[kUnroll], HRow<VEC, COLS> (&rr)[kUnroll]) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        kk[u] = m != 0 ? __ffs(m) - 1 : -1;
        if (kk[u] >= 0) {
          m &= m - 1;
          rr[u] = load_hist<VEC, COLS>(t, __shfl_sync(0xffffffffu, my0, kk[u]), __shfl_sync(0xffffffffu, my1, kk[u]), lane);
        }
      }
    };
    fetch(ks, rows);
    while (ks[0] >= 0) {
      fetch(kn, rows_n);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (ks[u] >= 0) {
          f(l0 + ks[u], p, rows[u]);
          ++p;
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        ks[u] = kn[u];
        rows[u] = rows_n[u];
      }
    }
  }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int VEC, int COLS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
build_features_kernel(Tables t, const float* __restrict__ target, const int32_t* __restrict__ offsets, __nv_bfloat16* __restrict__ X,
                      int64_t ldx) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float tv[COLS][VEC];
#pragma unroll
  for (int j = 0; j < COLS; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = (lane + 32 * j) * VEC + k;
      tv[j][k] = c < E ? __ldg(target + b * E + c) : 0.f;
    }
  const int pad = static_cast<int>(ldx) - 4 * E;
  for_valid_positions<VEC, COLS>(t, b, lane, offsets[b], [&](int, int32_t p, const HRow<VEC, COLS>& h) {
    __nv_bfloat16* row = X + static_cast<int64_t>(p) * ldx;
#pragma unroll
    for (int j = 0; j < COLS; ++j) {
      const int c = (lane + 32 * j) * VEC;
      if (c < E) {
        if constexpr (VEC == 2) {
          *reinterpret_cast<uint32_t*>(row + c) = pack2(tv[j][0], tv[j][1]);
          *reinterpret_cast<uint32_t*>(row + E + c) = pack2(h.v[j][0], h.v[j][1]);
          *reinterpret_cast<uint32_t*>(row + 2 * E + c) = pack2(__fsub_rn(tv[j][0], h.v[j][0]), __fsub_rn(tv[j][1], h.v[j][1]));
          *reinterpret_cast<uint32_t*>(row + 3 * E + c) = pack2(__fmul_rn(tv[j][0], h.v[j][0]), __fmul_rn(tv[j][1], h.v[j][1]));
        } else {
          row[c] = __float2bfloat16_rn(tv[j][0]);
          row[E + c] = __float2bfloat16_rn(h.v[j][0]);
          row[2 * E + c] = __float2bfloat16_rn(__fsub_rn(tv[j][0], h.v[j][0]));
          row[3 * E + c] = __float2bfloat16_rn(__fmul_rn(tv[j][0], h.v[j][0]));
        }
      }
    }
    if (lane < pad) row[4 * E + lane] = __float2bfloat16_rn(0.f);
  });
}

template <int VEC, int COLS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pool_fwd_kernel(Tables t, const int32_t* __restrict__ offsets, const float* __restrict__ w, float* __restrict__ rep) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float acc[COLS][VEC];
#pragma unroll
  for (int j = 0; j < COLS; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[j][k] = 0.f;
  for_valid_positions<VEC, COLS>(t, b, lane, offsets[b], [&](int, int32_t p, const HRow<VEC, COLS>& h) {
    const float wp = __ldg(w + p);
#pragma unroll
    for (int j = 0; j < COLS; ++j)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[j][k] = __fadd_rn(acc[j][k], __fmul_rn(wp, h.v[j][k]));      // position order, explicit rounding
  });
#pragma unroll
  for (int j = 0; j < COLS; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = (lane + 32 * j) * VEC + k;
      if (c < E) rep[b * E + c] = acc[j][k];
    }
}

template <int VEC, int COLS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pool_bwd_weights_kernel(Tables t, const int32_t* __restrict__ offsets, const float* __restrict__ d_rep, float* __restrict__ dw) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float g[COLS][VEC];
#pragma unroll
  for (int j = 0; j < COLS; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = (lane + 32 * j) * VEC + k;
      g[j][k] = c < E ? __ldg(d_rep + b * E + c) : 0.f;
    }
  for_valid_positions<VEC, COLS>(t, b, lane, offsets[b], [&](int, int32_t p, const HRow<VEC, COLS>& h) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < COLS; ++j)
#pragma unroll
      for (int k = 0; k < VEC; ++k) s = fmaf(g[j][k], h.v[j][k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);      // fixed tree: deterministic
    if (lane == 0) dw[p] = s;
  });
}

template <int VEC, int COLS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
feature_bwd_kernel(Tables t, const float* __restrict__ target, const int32_t* __restrict__ offsets, const __nv_bfloat16* __restrict__ dX,
                   int64_t ldx, const float* __restrict__ w, const float* __restrict__ d_rep, float* __restrict__ dh,
                   float* __restrict__ d_target, int zero_masked) {
  const int lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + threadIdx.x / 32;
  if (b >= t.B) return;
  const int E = t.E;
  float tv[COLS][VEC], gv[COLS][VEC], dt[COLS][VEC];
#pragma unroll
  for (int j = 0; j < COLS; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = (lane + 32 * j) * VEC + k;
      tv[j][k] = c < E ? __ldg(target + b * E + c) : 0.f;
      gv[j][k] = c < E ? __ldg(d_rep + b * E + c) : 0.f;
      dt[j][k] = 0.f;
    }
  if (zero_masked) {      // a dense history gradient (materialised-history form): masked positions get explicit zeros
    for (int l = 0; l < t.L; ++l) {
      if (!valid_at(t, b * t.L + l)) {
        for (int c = lane; c < E; c += 32) dh[(b * t.L + l) * E + c] = 0.f;
      }
    }
  }
  for_valid_positions<VEC, COLS>(t, b, lane, offsets[b], [&](int l, int32_t p, const HRow<VEC, COLS>& h) {
    const __nv_bfloat16* row = dX + static_cast<int64_t>(p) * ldx;
    const float wp = __ldg(w + p);
    float* out = dh + (b * t.L + l) * E;
#pragma unroll
    for (int j = 0; j < COLS; ++j) {
      const int c = (lane + 32 * j) * VEC;
      if (c < E) {
        float d[4][VEC];
        if constexpr (VEC == 2) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const uint32_t u = __ldcs(reinterpret_cast<const uint32_t*>(row + s4 * E + c));
            d[s4][0] = __uint_as_float(u << 16);
            d[s4][1] = __uint_as_float(u & 0xFFFF0000u);
          }
        } else {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) d[s4][0] = __bfloat162float(row[s4 * E + c]);
        }
        float o[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          // every op rounded explicitly, in this order: the oracle (oracle/ctr_oracle.py local_activation_unit_backward) does the same
          o[k] = __fadd_rn(__fadd_rn(__fsub_rn(d[1][k], d[2][k]), __fmul_rn(d[3][k], tv[j][k])), __fmul_rn(wp, gv[j][k]));
          dt[j][k] = __fadd_rn(dt[j][k], __fadd_rn(__fadd_rn(d[0][k], d[2][k]), __fmul_rn(d[3][k], h.v[j][k])));
        }
        if constexpr (VEC == 2) __stcs(reinterpret_cast<float2*>(out + c), make_float2(o[0], o[1]));
        else out[c] = o[0];
      }
    }
  });
#pragma unroll
  for (int j = 0; j < COLS; ++j)
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = (lane + 32 * j) * VEC + k;
      if (c < E) d_target[b * E + c] = dt[j][k];
    }
}

// pairs of columns when every pair is 8-byte aligned and stays inside one table
static bool pairs_ok(const Tables& t, int64_t ldx = 0) { return t.D[0] % 2 == 0 && t.D[1] % 2 == 0 && ldx % 2 == 0; }

// KERNEL<VEC, COLS> for the row width at hand: COLS = column groups per lane = ceil(E / (32 * VEC))
#define DIN_DISPATCH(KERNEL, PAIRS, ...)                                                                     \
  {                                                                                                          \
    const dim3 grid(grid_for(t.B, kWarpsPerBlock));                                                          \
    const int vec = (PAIRS) ? 2 : 1;                                                                         \
    const int cols = (t.E + 32 * vec - 1) / (32 * vec);                                                      \
    if (vec == 2) {                                                                                          \
      if (cols <= 1) KERNEL<2, 1><<<grid, kWarpsPerBlock * 32, 0, st>>>(__VA_ARGS__);                        \
      else KERNEL<2, 2><<<grid, kWarpsPerBlock * 32, 0, st>>>(__VA_ARGS__);                                  \
    } else if (cols <= 1) KERNEL<1, 1><<<grid, kWarpsPerBlock * 32, 0, st>>>(__VA_ARGS__);                   \
    else if (cols <= 2) KERNEL<1, 2><<<grid, kWarpsPerBlock * 32, 0, st>>>(__VA_ARGS__);                     \
    else KERNEL<1, 4><<<grid, kWarpsPerBlock * 32, 0, st>>>(__VA_ARGS__);                                    \
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DIN_DISPATCH(build_features_kernel, pairs_ok(t, ldx), t, target, offsets, static_cast<__nv_bfloat16*>(x_bf16), ldx)
  RB_LAUNCH_CHECK("din build_features_kernel");
  return RB_OK;
}

extern "C" int rb_din_pool_fwd(const rb_din_history* h, const int32_t* offsets, const float* w, float* rep, void* stream) {
  Tables t;
  int rc = fill(&t, h);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(offsets != nullptr && w != nullptr && rep != nullptr, RB_ERR_ARG, "offsets / w / rep is null");
  if (t.B == 0) return RB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DIN_DISPATCH(pool_fwd_kernel, pairs_ok(t), t, offsets, w, rep)
  RB_LAUNCH_CHECK("din pool_fwd_kernel");
  return RB_OK;
}

extern "C" int rb_din_pool_bwd_weights(const rb_din_history* h, const int32_t* offsets, const float* d_rep, float* dw, void* stream) {
  Tables t;
  int rc = fill(&t, h);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(offsets != nullptr && d_rep != nullptr && dw != nullptr, RB_ERR_ARG, "offsets / d_rep / dw is null");
  if (t.B == 0) return RB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DIN_DISPATCH(pool_bwd_weights_kernel, pairs_ok(t), t, offsets, d_rep, dw)
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DIN_DISPATCH(feature_bwd_kernel, pairs_ok(t, ldx), t, target, offsets, static_cast<const __nv_bfloat16*>(dx_bf16), ldx, w, d_rep, dh, d_target,
               zero_masked)
  RB_LAUNCH_CHECK("din feature_bwd_kernel");
  return RB_OK;
}
