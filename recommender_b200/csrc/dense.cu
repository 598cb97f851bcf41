// Dense-side helpers around the sparse path: the optimizer step of ALL dense
// parameters in one launch, and the bias gradient (column sum of a [B, N] activation gradient).
// Both replace chains of tiny framework kernels that cost more in launch latency than in bytes.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rb {

// ---- multi-tensor dense optimizer step --------------------------------------------------------------
struct DenseSlots {
  int num;
  int opt;
  float lr, b1, b2, omb1, omb2, eps, alpha;
  const float* alpha_dev;
  float* p[RB_MAX_DENSE_TENSORS];
  float* s0[RB_MAX_DENSE_TENSORS];
  float* s1[RB_MAX_DENSE_TENSORS];
  const float* g[RB_MAX_DENSE_TENSORS];
  __nv_bfloat16* shadow[RB_MAX_DENSE_TENSORS];   // optional bf16 copy of the parameter (what the bf16 GEMMs read)
  int chunk_start[RB_MAX_DENSE_TENSORS + 1];  // prefix sum of ceil(n / kDenseChunk)
  int64_t n[RB_MAX_DENSE_TENSORS];
};

constexpr int kDenseChunk = 2048;  // elements per CTA
constexpr int kDenseThreads = 256;

// Keras `_resource_apply_dense` formulas, every op explicitly rounded so
// the result matches the numpy oracle bit for bit.
__global__ void __launch_bounds__(kDenseThreads) dense_opt_kernel(const __grid_constant__ DenseSlots s) {
  griddep_wait();
  griddep_release();
  int t = 0;
  while (t + 1 < s.num && static_cast<int>(blockIdx.x) >= s.chunk_start[t + 1]) ++t;
  const int64_t base = static_cast<int64_t>(blockIdx.x - s.chunk_start[t]) * kDenseChunk;
  const int64_t end = min(base + kDenseChunk, s.n[t]);
  float* __restrict__ p = s.p[t];
  float* __restrict__ m = s.s0[t];
  float* __restrict__ v = s.s1[t];
  const float* __restrict__ g = s.g[t];
  const float alpha = s.alpha_dev != nullptr ? __ldg(s.alpha_dev) : s.alpha;
  __nv_bfloat16* __restrict__ sh = s.shadow[t];
  for (int64_t i = base + threadIdx.x; i < end; i += kDenseThreads) {
    const float gi = g[i];
    if (s.opt == RB_OPT_ADAM_LAZY || s.opt == RB_OPT_ADAM_TF_DENSE) {
      const float mi = __fadd_rn(__fmul_rn(m[i], s.b1), __fmul_rn(gi, s.omb1));
      const float vi = __fadd_rn(__fmul_rn(v[i], s.b2), __fmul_rn(__fmul_rn(gi, gi), s.omb2));
      m[i] = mi;
      v[i] = vi;
      p[i] = __fsub_rn(p[i], __fdiv_rn(__fmul_rn(alpha, mi), __fadd_rn(__fsqrt_rn(vi), s.eps)));
    } else if (s.opt == RB_OPT_ADAGRAD) {
      const float ai = __fadd_rn(m[i], __fmul_rn(gi, gi));
      m[i] = ai;
      p[i] = __fsub_rn(p[i], __fdiv_rn(__fmul_rn(s.lr, gi), __fadd_rn(__fsqrt_rn(ai), s.eps)));
    } else {
      p[i] = __fsub_rn(p[i], __fmul_rn(s.lr, gi));
    }
    if (sh != nullptr) sh[i] = __float2bfloat16_rn(p[i]);
  }
}

// ---- column sum of a [rows, cols] matrix (bf16 or fp32) -> fp32 [cols] ---------------------------------
// Stage 1: each CTA owns a slab of rows; a thread owns 8 consecutive columns (one 16 B / 32 B load
// per row) and kRowLanes rows in parallel; partial[cta][cols].  Stage 2: fixed-order sum of the
// partials (deterministic, no float atomics).
constexpr int kColThreads = 256;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = __uint_as_float(w[k] << 16);
    v[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldcs(reinterpret_cast<const float4*>(p));
  const float4 b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <typename T>
__global__ void __launch_bounds__(kColThreads)
colsum_partial_kernel(const T* __restrict__ x, int64_t rows, int cols, int64_t row_stride, float* __restrict__ partial) {
  griddep_wait();
  griddep_release();
  __shared__ float s_red[kColThreads * 8];
  const int vcols = cols / 8;                      // vector columns
  const int row_lanes = kColThreads / vcols;       // rows handled in parallel (vcols <= 256)
  const int vc = threadIdx.x % vcols;
  const int rl = threadIdx.x / vcols;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (rl < row_lanes) {
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * per;
    const int64_t r1 = min(r0 + per, rows);
    int64_t r = r0 + rl;
    for (; r + 3 * row_lanes < r1; r += 4 * row_lanes) {   // 4 independent loads in flight
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8<T>(x + (r + u * row_lanes) * row_stride + vc * 8, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[u][k];
    }
    for (; r < r1; r += row_lanes) {
      float v[8];
      load8<T>(x + r * row_stride + vc * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  // thread c < cols sums its column over the row lanes in fixed order
  for (int c = threadIdx.x; c < cols; c += kColThreads) {
    float t = 0.f;
    for (int l = 0; l < row_lanes; ++l) t += s_red[(l * vcols + c / 8) * 8 + (c % 8)];
    partial[static_cast<int64_t>(blockIdx.x) * cols + c] = t;
  }
}

// Stage 2: a CTA owns 32 columns; warp w adds partials w, w+8, ... (coalesced 128 B reads, 8 loads in
// flight), then the 8 warp sums are combined in fixed order.
constexpr int kFinalWarps = 32;
__global__ void __launch_bounds__(kFinalWarps * 32) colsum_final_kernel(const float* __restrict__ partial, int nparts, int cols,
                                                                         float* __restrict__ out) {
  griddep_wait();
  griddep_release();
  __shared__ float s_w[kFinalWarps][32];
  const int lane = threadIdx.x % 32, w = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + lane;
  float t = 0.f;
  if (c < cols) {
    int p = w;
    for (; p + 7 * kFinalWarps < nparts; p += 8 * kFinalWarps) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = partial[static_cast<int64_t>(p + kFinalWarps * u) * cols + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) t += v[u];
    }
    for (; p < nparts; p += kFinalWarps) t += partial[static_cast<int64_t>(p) * cols + c];
  }
  s_w[w][lane] = t;
  __syncthreads();
  if (w == 0 && c < cols) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < kFinalWarps; ++k) r += s_w[k][lane];
    out[c] = r;
  }
}

// ---- loss head: Keras binary_crossentropy on probabilities + its gradient, one pass ------------------------
// ctr/train.py:85 compiles BinaryCrossentropy(from_logits=False); for DLRM (last op Squeeze) Keras evaluates the
// clipped-probability form:  p <- clip(p, 1e-7, 1-1e-7);  l = -(y log(p+1e-7) + (1-y) log(1-p+1e-7)),
// batch mean.  dprob = dl/dp / n with the clip's pass-through mask.  Partials per CTA, fixed-order final sum.
constexpr int kBceThreads = 256;

__global__ void __launch_bounds__(kBceThreads)
bce_partial_kernel(const float* __restrict__ prob, const void* __restrict__ label, int label_is_i64, int64_t n, float inv_n,
                   float* __restrict__ dprob, float* __restrict__ partial) {
  griddep_wait();
  griddep_release();
  __shared__ float s_red[kBceThreads / 32];
  const float eps = 1e-7f;
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kBceThreads + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * kBceThreads) {
    const float p = prob[i];
    const float y = label_is_i64 ? static_cast<float>(static_cast<const int64_t*>(label)[i]) : static_cast<const float*>(label)[i];
    const float pc = fminf(fmaxf(p, eps), 1.0f - eps);
    const float a = pc + eps, b = (1.0f - pc) + eps;
    acc += -(y * logf(a) + (1.0f - y) * logf(b));
    if (dprob != nullptr) {
      const bool pass = (p >= eps) && (p <= 1.0f - eps);
      dprob[i] = pass ? (-(y / a) + (1.0f - y) / b) * inv_n : 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x % 32 == 0) s_red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kBceThreads / 32; ++w) t += s_red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void bce_final_kernel(const float* __restrict__ partial, int nparts, float inv_n, float* __restrict__ loss) {
  griddep_wait();
  griddep_release();
  float t = 0.f;
  for (int i = 0; i < nparts; ++i) t += partial[i];
  loss[0] = t * inv_n;
}

static int bce_parts(int64_t n) {
  int64_t parts = (n + kBceThreads * 4 - 1) / (kBceThreads * 4);
  if (parts < 1) parts = 1;
  if (parts > 512) parts = 512;
  return static_cast<int>(parts);
}

static int colsum_parts(int64_t rows) {
  int64_t parts = rows / 64;
  if (parts < 1) parts = 1;
  if (parts > 4 * kNumSMs) parts = 4 * kNumSMs;
  return static_cast<int>(parts);
}

}  // namespace rb

using namespace rb;

extern "C" int rb_dense_opt_step(const rb_dense_slot* slots, int32_t num, const rb_opt_params* opt, void* stream) {
  RB_CHECK_ARG(num >= 0 && num <= RB_MAX_DENSE_TENSORS, RB_ERR_ARG, "0..%d dense tensors per call, got %d", RB_MAX_DENSE_TENSORS, num);
  if (num == 0) return RB_OK;
  RB_CHECK_ARG(slots != nullptr && opt != nullptr, RB_ERR_ARG, "slots/opt is null");
  const int o = opt->optimizer;
  RB_CHECK_ARG(o >= RB_OPT_SGD && o <= RB_OPT_ADAM_TF_DENSE, RB_ERR_ARG, "bad optimizer %d", o);
  const bool adam = (o == RB_OPT_ADAM_LAZY || o == RB_OPT_ADAM_TF_DENSE);
  RB_CHECK_ARG(!adam || opt->step >= 1, RB_ERR_ARG, "Adam needs step >= 1");
  DenseSlots s;
  s.num = num;
  s.opt = o;
  s.lr = opt->lr;
  s.b1 = opt->beta_1;
  s.b2 = opt->beta_2;
  s.omb1 = 1.0f - opt->beta_1;
  s.omb2 = 1.0f - opt->beta_2;
  s.eps = opt->epsilon;
  s.alpha = adam ? rb_adam_alpha_t(opt->lr, opt->beta_1, opt->beta_2, opt->step) : 0.f;
  s.alpha_dev = adam ? opt->alpha_t_dev : nullptr;
  int chunks = 0;
  for (int t = 0; t < num; ++t) {
    const rb_dense_slot& d = slots[t];
    RB_CHECK_ARG(d.param != nullptr && d.grad != nullptr && d.n >= 0, RB_ERR_ARG, "dense tensor %d: null param/grad or n < 0", t);
    RB_CHECK_ARG(!adam || (d.state0 != nullptr && d.state1 != nullptr), RB_ERR_ARG, "dense tensor %d: Adam needs m and v", t);
    RB_CHECK_ARG(o != RB_OPT_ADAGRAD || d.state0 != nullptr, RB_ERR_ARG, "dense tensor %d: Adagrad needs its accumulator", t);
    s.p[t] = d.param;
    s.s0[t] = d.state0;
    s.s1[t] = d.state1;
    s.g[t] = d.grad;
    s.shadow[t] = static_cast<__nv_bfloat16*>(d.shadow_bf16);
    s.n[t] = d.n;
    s.chunk_start[t] = chunks;
    chunks += static_cast<int>((d.n + kDenseChunk - 1) / kDenseChunk);
  }
  for (int t = num; t <= RB_MAX_DENSE_TENSORS; ++t) s.chunk_start[t] = chunks;
  if (chunks == 0) return RB_OK;
  RB_CUDA(launch_dependent(dense_opt_kernel, chunks, kDenseThreads, 0, static_cast<cudaStream_t>(stream), true, s));
  RB_LAUNCH_CHECK("dense_opt_kernel");
  return RB_OK;
}

extern "C" size_t rb_colsum_workspace_bytes(int64_t rows, int32_t cols) {
  if (rows < 0 || cols <= 0) return 0;
  return static_cast<size_t>(colsum_parts(rows)) * cols * sizeof(float) + 256;
}

extern "C" int rb_colsum(const void* x, int32_t dtype, int64_t rows, int32_t cols, int64_t row_stride, float* out, void* ws,
                         size_t ws_bytes, void* stream) {
  RB_CHECK_ARG(dtype == RB_F32 || dtype == RB_BF16, RB_ERR_ARG, "bad dtype %d", dtype);
  RB_CHECK_ARG(rows >= 0 && cols > 0 && row_stride >= cols, RB_ERR_ARG, "bad rows/cols/stride");
  RB_CHECK_ARG(cols % 8 == 0 && cols / 8 <= kColThreads, RB_ERR_SHAPE, "colsum needs cols %% 8 == 0 and cols <= %d, got %d", 8 * kColThreads, cols);
  RB_CHECK_ARG(out != nullptr, RB_ERR_ARG, "out is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    RB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    return RB_OK;
  }
  const size_t esz = dtype == RB_BF16 ? 2 : 4;
  RB_CHECK_ARG(x != nullptr && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (row_stride * esz) % 16 == 0, RB_ERR_ALIGN,
               "x / row stride not 16 B aligned");
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= rb_colsum_workspace_bytes(rows, cols), RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
               rb_colsum_workspace_bytes(rows, cols), ws_bytes);
  const int parts = colsum_parts(rows);
  float* partial = static_cast<float*>(ws);
  if (dtype == RB_BF16)
    RB_CUDA(launch_dependent(colsum_partial_kernel<__nv_bfloat16>, parts, kColThreads, 0, st, true, static_cast<const __nv_bfloat16*>(x), rows, cols, row_stride, partial));
  else
    RB_CUDA(launch_dependent(colsum_partial_kernel<float>, parts, kColThreads, 0, st, true, static_cast<const float*>(x), rows, cols, row_stride, partial));
  RB_LAUNCH_CHECK("colsum_partial_kernel");
  RB_CUDA(launch_dependent(colsum_final_kernel, (cols + 31) / 32, kFinalWarps * 32, 0, st, true, partial, parts, cols, out));
  RB_LAUNCH_CHECK("colsum_final_kernel");
  return RB_OK;
}

extern "C" size_t rb_bce_workspace_bytes(int64_t n) { return n < 0 ? 0 : static_cast<size_t>(bce_parts(n)) * sizeof(float) + 256; }

extern "C" int rb_bce_clipped(const float* prob, const void* label, int32_t label_type, int64_t n, float* loss_out, float* dprob_out,
                              void* ws, size_t ws_bytes, void* stream) {
  RB_CHECK_ARG(n > 0 && prob != nullptr && label != nullptr && loss_out != nullptr, RB_ERR_ARG, "null pointer or n <= 0");
  RB_CHECK_ARG(label_type == 0 || label_type == 1, RB_ERR_ARG, "label_type: 0 = f32, 1 = i64");
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= rb_bce_workspace_bytes(n), RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
               rb_bce_workspace_bytes(n), ws_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int parts = bce_parts(n);
  const float inv_n = 1.0f / static_cast<float>(n);
  RB_CUDA(launch_dependent(bce_partial_kernel, parts, kBceThreads, 0, st, true, prob, label, label_type, n, inv_n, dprob_out, static_cast<float*>(ws)));
  RB_LAUNCH_CHECK("bce_partial_kernel");
  RB_CUDA(launch_dependent(bce_final_kernel, 1, 1, 0, st, true, static_cast<const float*>(ws), parts, inv_n, loss_out));
  RB_LAUNCH_CHECK("bce_final_kernel");
  return RB_OK;
}
