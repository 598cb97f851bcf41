// K3..K6 / K10: DotInteraction forward and backward on bf16 tensor cores (sm_100a).
//
// One warp owns one sample.  The F' <= 32 feature rows (F embedding rows — read from E or, in
// the fused-gather form, straight from the table through idx — plus the optional dense vector)
// are converted to bf16 into a padded shared tile; Z = X X^T (forward) and dX = (G+G^T) X
// (backward) run as m16n8k16 bf16 MMAs with fp32 accumulators.  The triangular mask, the
// zero-fill / compaction, the DLRM concat of the dense vector and the "|| bmlp" tail are all
// applied in the epilogue, so none of the reference's temporaries exist.
// The per-sample problem (27x27x64) is far below a tcgen05 128-row tile and the kernel is
// HBM-bound (9 FLOP/B vs a ridge of ~214); the tcgen05 form lives in interaction_umma.cu.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rb {

constexpr int kIxWarps = 4;  // samples per CTA
constexpr int kPad = 8;      // bf16 elements of row padding (16 B) -> conflict-free ldmatrix

struct IxArgs {
  const float* E;          // [B,F,D] or null (fused gather)
  const float* table;      // used when E == null
  IndexMap map;            // idx[B,F]
  const float* dense_vec;  // [B,D] or null
  int64_t B;
  int F;                   // embedding features
  int Fp;                  // F + (dense_vec != null)
  int self_interaction;
  int skip_gather;
  int tail;
  int ncols;               // interaction columns (without tail)
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// kept(i,j) and its position in the compact (skip_gather == 0) layout   (ctr/layers.py:27-42)
__device__ __forceinline__ bool kept(int i, int j, int self_interaction) { return self_interaction ? (j <= i) : (j > i); }
__device__ __forceinline__ int compact_pos(int i, int j, int Fp, int self_interaction) {
  return self_interaction ? (i * (i + 1) / 2 + j) : (i * Fp - i * (i + 1) / 2 + (j - i - 1));
}
__device__ __forceinline__ int out_pos(int i, int j, const IxArgs& a) {
  return a.skip_gather ? (i * a.Fp + j) : compact_pos(i, j, a.Fp, a.self_interaction);
}

// Load the sample's F' rows (fp32) and park them as bf16 in xs[32][D+kPad]; rows >= F' are zero.
template <int D>
__device__ __forceinline__ void load_rows_bf16(const IxArgs& a, int64_t b, __nv_bfloat16* xs, int lane) {
  constexpr int kLanesPerRow = D / 4;
  constexpr int kRowsPerIter = 32 / kLanesPerRow;
  constexpr int kIters = 32 / kRowsPerIter;
  const int sub = lane / kLanesPerRow;
  const int c = (lane % kLanesPerRow) * 4;
  float4 v[kIters];
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int r = it * kRowsPerIter + sub;
    const float* src = nullptr;
    if (r < a.F) {
      if (a.E != nullptr) {
        src = a.E + (b * a.F + r) * D;
      } else {
        const int64_t row = map_index(a.map, b * a.F + r);
        if (row >= 0) src = a.table + row * D;
      }
    } else if (r == a.F && a.dense_vec != nullptr) {
      src = a.dense_vec + b * D;
    }
    v[it] = (src != nullptr) ? __ldg(reinterpret_cast<const float4*>(src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int r = it * kRowsPerIter + sub;
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[it].x, v[it].y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v[it].z, v[it].w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(xs + r * (D + kPad) + c) = pk;
  }
}

template <int D>
__global__ void __launch_bounds__(kIxWarps * 32)
dot_interaction_fwd_kernel(IxArgs a, float* __restrict__ out, int64_t out_stride, int out_smem_floats) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kIxWarps + warp;
  if (b >= a.B) return;
  constexpr int kXsBytes = 32 * (D + kPad) * 2;
  unsigned char* my = smem + static_cast<size_t>(warp) * (kXsBytes + out_smem_floats * 4);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(my);
  float* os = reinterpret_cast<float*>(my + kXsBytes);

  load_rows_bf16<D>(a, b, xs, lane);
  const int total = a.ncols + (a.tail ? D : 0);
  if (!a.skip_gather) {
    // compact layout: every column is written by exactly one kept (i,j); nothing to clear
  } else {
    for (int i = lane; i < a.ncols; i += 32) os[i] = 0.f;
  }
  __syncwarp();

  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[mt][nt][k] = 0.f;

#pragma unroll
  for (int k0 = 0; k0 < D; k0 += 16) {
    uint32_t afrag[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int row = mt * 16 + (lane % 8) + ((lane / 8) % 2) * 8;
      const int col = k0 + (lane / 16) * 8;
      ldmatrix_x4(afrag[mt], smem_addr(xs + row * (D + kPad) + col));
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t bfrag[2];
      const int row = nt * 8 + (lane % 8);
      const int col = k0 + ((lane / 8) % 2) * 8;
      ldmatrix_x2(bfrag, smem_addr(xs + row * (D + kPad) + col));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(acc[mt][nt], afrag[mt], bfrag);
    }
  }

  // epilogue: mask + placement into the staged output row
  const int g = lane / 4, t2 = (lane % 4) * 2;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = mt * 16 + g + (k / 2) * 8;
        const int j = nt * 8 + t2 + (k % 2);
        if (i < a.Fp && j < a.Fp && kept(i, j, a.self_interaction)) os[out_pos(i, j, a)] = acc[mt][nt][k];
      }
  if (a.tail) {
    for (int d = lane; d < D; d += 32) os[a.ncols + d] = __ldg(a.dense_vec + b * D + d);
  }
  __syncwarp();
  float* dst = out + b * out_stride;
  for (int i = lane; i < total; i += 32) __stcs(dst + i, os[i]);
}

template <int D>
__global__ void __launch_bounds__(kIxWarps * 32)
dot_interaction_bwd_kernel(IxArgs a, const float* __restrict__ dOut, int64_t dout_stride, float* __restrict__ dE,
                           float* __restrict__ d_dense, int out_smem_floats) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kIxWarps + warp;
  if (b >= a.B) return;
  constexpr int kXsBytes = 32 * (D + kPad) * 2;
  constexpr int kSsBytes = 32 * (32 + kPad) * 2;
  unsigned char* my = smem + static_cast<size_t>(warp) * (kXsBytes + kSsBytes + out_smem_floats * 4);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(my);
  __nv_bfloat16* ss = reinterpret_cast<__nv_bfloat16*>(my + kXsBytes);
  float* gs = reinterpret_cast<float*>(my + kXsBytes + kSsBytes);

  // stage the dOut row (interaction columns + tail) with coalesced loads
  const int total = a.ncols + (a.tail ? D : 0);
  const float* src = dOut + b * dout_stride;
  for (int i = lane; i < total; i += 32) gs[i] = __ldcs(src + i);
  load_rows_bf16<D>(a, b, xs, lane);
  __syncwarp();

  // S = G + G^T with G = mask (.) dOut, as bf16 [32][32+kPad]; lane owns column j = lane
  for (int i = 0; i < 32; ++i) {
    const int j = lane;
    float sv = 0.f;
    if (i < a.Fp && j < a.Fp) {
      if (kept(i, j, a.self_interaction)) sv += gs[out_pos(i, j, a)];
      if (kept(j, i, a.self_interaction)) sv += gs[out_pos(j, i, a)];
    }
    ss[i * (32 + kPad) + j] = __float2bfloat16_rn(sv);
  }
  __syncwarp();

  // A = S (two k-steps of 16), fragments kept for all n-chunks
  uint32_t afrag[2][2][4];  // [mt][kstep]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int row = mt * 16 + (lane % 8) + ((lane / 8) % 2) * 8;
      const int col = ks * 16 + (lane / 16) * 8;
      ldmatrix_x4(afrag[mt][ks], smem_addr(ss + row * (32 + kPad) + col));
    }

  const int g = lane / 4, t2 = (lane % 4) * 2;
#pragma unroll
  for (int n0 = 0; n0 < D; n0 += 32) {  // chunks of 4 n-tiles bound the accumulator registers
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[mt][nt][k] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (n0 + nt * 8 < D) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t bfrag[2];
          // B[k][n] = X[k][n]: transposed 8x8 loads of X rows k, columns n
          const int row = ks * 16 + (lane % 8) + ((lane / 8) % 2) * 8;
          ldmatrix_x2_trans(bfrag, smem_addr(xs + row * (D + kPad) + n0 + nt * 8));
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(acc[mt][nt], afrag[mt][ks], bfrag);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = n0 + nt * 8 + t2;
        if (col < D) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = mt * 16 + g + h * 8;
            float2 v = make_float2(acc[mt][nt][h * 2], acc[mt][nt][h * 2 + 1]);
            if (i < a.F) {
              if (dE != nullptr) __stcs(reinterpret_cast<float2*>(dE + (b * a.F + i) * D + col), v);
            } else if (i == a.F && a.dense_vec != nullptr && d_dense != nullptr) {
              if (a.tail) {
                v.x += gs[a.ncols + col];
                v.y += gs[a.ncols + col + 1];
              }
              *reinterpret_cast<float2*>(d_dense + b * D + col) = v;
            }
          }
        }
      }
  }
}

static int fill_args(IxArgs* a, const float* E, const float* table, int64_t rows, const void* idx, int idx_type,
                     const int64_t* off, const float* dense_vec, int64_t B, int F, int D, int self_interaction,
                     int skip_gather, int tail) {
  RB_CHECK_ARG(B >= 0 && F > 0, RB_ERR_ARG, "bad B/F");
  RB_CHECK_ARG(D == 16 || D == 32 || D == 64 || D == 128, RB_ERR_SHAPE, "dot interaction needs D in {16,32,64,128}, got %d", D);
  const int Fp = F + (dense_vec != nullptr ? 1 : 0);
  RB_CHECK_ARG(Fp <= 32, RB_ERR_SHAPE, "dot interaction supports at most 32 features, got %d", Fp);
  RB_CHECK_ARG(!tail || dense_vec != nullptr, RB_ERR_ARG, "tail requires dense_vec");
  if (E == nullptr) {
    RB_CHECK_ARG(table != nullptr && idx != nullptr && rows > 0, RB_ERR_ARG, "fused gather needs table and idx");
    RB_CHECK_ARG(idx_type == RB_I32 || idx_type == RB_I64, RB_ERR_ARG, "bad index type");
    RB_CHECK_ARG(aligned_for(table, 4), RB_ERR_ALIGN, "table not 16 B aligned");
  } else {
    RB_CHECK_ARG(aligned_for(E, 4), RB_ERR_ALIGN, "E not 16 B aligned");
  }
  RB_CHECK_ARG(dense_vec == nullptr || aligned_for(dense_vec, 4), RB_ERR_ALIGN, "dense_vec not 16 B aligned");
  a->E = E;
  a->table = table;
  a->map = make_index_map(idx, idx_type, off, 0, rows, F);
  a->dense_vec = dense_vec;
  a->B = B;
  a->F = F;
  a->Fp = Fp;
  a->self_interaction = self_interaction ? 1 : 0;
  a->skip_gather = skip_gather ? 1 : 0;
  a->tail = tail ? 1 : 0;
  a->ncols = skip_gather ? Fp * Fp : (self_interaction ? Fp * (Fp + 1) / 2 : Fp * (Fp - 1) / 2);
  return RB_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return RB_OK;
}

}  // namespace rb

using namespace rb;

extern "C" int rb_dot_interaction_fwd(const float* E, const float* table, int64_t rows, const void* idx,
                                      int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                      int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                      int32_t tail, float* out, int64_t out_stride, void* stream) {
  IxArgs a;
  int rc = fill_args(&a, E, table, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D, self_interaction,
                     skip_gather, tail);
  if (rc != RB_OK) return rc;
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(out != nullptr && out_stride >= a.ncols + (a.tail ? D : 0), RB_ERR_ARG, "out is null or out_stride too small");
  const int out_floats = (a.ncols + D + 3) / 4 * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned int grid = grid_for(B, kIxWarps);
#define LAUNCH(DD)                                                                                   \
  {                                                                                                  \
    size_t smem = static_cast<size_t>(kIxWarps) * (32 * (DD + kPad) * 2 + out_floats * 4);          \
    rc = set_smem(dot_interaction_fwd_kernel<DD>, smem);                                             \
    if (rc != RB_OK) return rc;                                                                      \
    dot_interaction_fwd_kernel<DD><<<grid, kIxWarps * 32, smem, st>>>(a, out, out_stride, out_floats); \
  }
  if (D == 16) LAUNCH(16) else if (D == 32) LAUNCH(32) else if (D == 64) LAUNCH(64) else LAUNCH(128)
#undef LAUNCH
  RB_LAUNCH_CHECK("dot_interaction_fwd_kernel");
  return RB_OK;
}

extern "C" int rb_dot_interaction_bwd(const float* E, const float* table, int64_t rows, const void* idx,
                                      int32_t idx_type, const int64_t* field_row_offset, const float* dense_vec,
                                      int64_t B, int32_t F, int32_t D, int32_t self_interaction, int32_t skip_gather,
                                      int32_t tail, const float* dOut, int64_t dout_stride, float* dE, float* d_dense,
                                      void* stream) {
  IxArgs a;
  int rc = fill_args(&a, E, table, rows, idx, idx_type, field_row_offset, dense_vec, B, F, D, self_interaction,
                     skip_gather, tail);
  if (rc != RB_OK) return rc;
  if (B == 0) return RB_OK;
  RB_CHECK_ARG(dOut != nullptr && dout_stride >= a.ncols + (a.tail ? D : 0), RB_ERR_ARG, "dOut is null or stride too small");
  RB_CHECK_ARG((dE == nullptr || aligned_for(dE, 4)) && (d_dense == nullptr || aligned_for(d_dense, 4)), RB_ERR_ALIGN,
               "dE/d_dense not 16 B aligned");
  const int out_floats = (a.ncols + D + 3) / 4 * 4;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned int grid = grid_for(B, kIxWarps);
#define LAUNCH(DD)                                                                                               \
  {                                                                                                              \
    size_t smem = static_cast<size_t>(kIxWarps) * (32 * (DD + kPad) * 2 + 32 * (32 + kPad) * 2 + out_floats * 4); \
    rc = set_smem(dot_interaction_bwd_kernel<DD>, smem);                                                         \
    if (rc != RB_OK) return rc;                                                                                  \
    dot_interaction_bwd_kernel<DD><<<grid, kIxWarps * 32, smem, st>>>(a, dOut, dout_stride, dE, d_dense, out_floats); \
  }
  if (D == 16) LAUNCH(16) else if (D == 32) LAUNCH(32) else if (D == 64) LAUNCH(64) else LAUNCH(128)
#undef LAUNCH
  RB_LAUNCH_CHECK("dot_interaction_bwd_kernel");
  return RB_OK;
}
