// Shared device/host helpers for the recsys_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/recsys_b200.h"

namespace rb {

// ---- error plumbing (thread-local message, C ABI returns codes) ---------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launches(int n);  // kernels of THIS library launched so far (rb_kernel_launches)

#define RB_CHECK_ARG(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::rb::set_error(__VA_ARGS__);        \
      return (code);                       \
    }                                      \
  } while (0)

#define RB_CUDA(call)                                            \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return ::rb::cuda_fail(e__, #call);  \
  } while (0)

#define RB_LAUNCH_CHECK(name)                                          \
  do {                                                                 \
    cudaError_t e__ = cudaPeekAtLastError();                           \
    if (e__ != cudaSuccess) return ::rb::cuda_fail(e__, name);         \
    ::rb::count_launches(1);                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- layout of one embedding row over a thread group --------------------------------------
// A row of D floats is handled by a power-of-two group of GS lanes, each moving VEC floats
// (VEC = 4 -> 128-bit, 2 -> 64-bit, 1 -> 32-bit), chosen from D's alignment.  D / VEC <= 32.
struct RowGeom {
  int vec;  // 4, 2 or 1
  int gs;   // 4, 8, 16 or 32 lanes per row
};

inline bool row_geom(int D, RowGeom* g) {
  if (D <= 0) return false;
  int vec = (D % 4 == 0) ? 4 : (D % 2 == 0) ? 2 : 1;
  int q = D / vec;  // vector elements per row
  if (q > 32) return false;
  int gs = 4;
  while (gs < q) gs <<= 1;
  g->vec = vec;
  g->gs = gs;
  return true;
}

inline bool aligned_for(const void* p, int vec) { return (reinterpret_cast<uintptr_t>(p) % (vec * 4)) == 0; }

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
};
template <>
struct Vec<2> {
  using T = float2;
};
template <>
struct Vec<1> {
  using T = float;
};

template <int VEC>
struct Row {  // VEC floats in registers
  float v[VEC];
};

// streaming (read-once / write-once) and default loads/stores of VEC floats
template <int VEC>
__device__ __forceinline__ Row<VEC> ld_row(const float* p) {
  Row<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

// plain (coherent) load: for table/state rows this kernel itself rewrites
template <int VEC>
__device__ __forceinline__ Row<VEC> ld_row_rw(const float* p) {
  Row<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = *p;
  }
  return r;
}

// evict-first streaming load (gradient rows are read exactly once)
template <int VEC>
__device__ __forceinline__ Row<VEC> ld_row_stream(const float* p) {
  Row<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else if constexpr (VEC == 2) {
    float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  } else {
    r.v[0] = __ldcs(p);
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void st_row(float* p, const Row<VEC>& r) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
  } else {
    *p = r.v[0];
  }
}

template <int VEC>
__device__ __forceinline__ void st_row_stream(float* p, const Row<VEC>& r) {
  if constexpr (VEC == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  } else if constexpr (VEC == 2) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(r.v[0], r.v[1]));
  } else {
    __stcs(p, r.v[0]);
  }
}

template <int VEC>
__device__ __forceinline__ Row<VEC> zero_row() {
  Row<VEC> r;
#pragma unroll
  for (int i = 0; i < VEC; ++i) r.v[i] = 0.f;
  return r;
}

// ---- the optimizers' row arithmetic, every op explicitly rounded (no FMA contraction) so that
// every kernel that updates a row — the segmented reduction's sink and the interaction backward's fused update of rows a
// step touches once — produces the same bits as the numpy oracle
struct OptMath {
  int opt;  // rb_optimizer
  float lr, b1, b2, omb1, omb2, eps, alpha;
};

template <int VEC>
__device__ __forceinline__ void opt_row_math(const OptMath& p, const Row<VEC>& g, Row<VEC>& w, Row<VEC>& m, Row<VEC>& v) {
  if (p.opt == RB_OPT_ADAM_LAZY) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      m.v[i] = __fadd_rn(__fmul_rn(m.v[i], p.b1), __fmul_rn(g.v[i], p.omb1));
      v.v[i] = __fadd_rn(__fmul_rn(v.v[i], p.b2), __fmul_rn(__fmul_rn(g.v[i], g.v[i]), p.omb2));
      w.v[i] = __fsub_rn(w.v[i], __fdiv_rn(__fmul_rn(p.alpha, m.v[i]), __fadd_rn(__fsqrt_rn(v.v[i]), p.eps)));
    }
  } else if (p.opt == RB_OPT_ADAM_TF_DENSE) {   // scatter-add phase only: the decay / apply passes cover every row
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      m.v[i] = __fadd_rn(m.v[i], __fmul_rn(g.v[i], p.omb1));
      v.v[i] = __fadd_rn(v.v[i], __fmul_rn(__fmul_rn(g.v[i], g.v[i]), p.omb2));
    }
  } else if (p.opt == RB_OPT_ADAGRAD) {         // accumulator in m
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      m.v[i] = __fadd_rn(m.v[i], __fmul_rn(g.v[i], g.v[i]));
      w.v[i] = __fsub_rn(w.v[i], __fdiv_rn(__fmul_rn(p.lr, g.v[i]), __fadd_rn(__fsqrt_rn(m.v[i]), p.eps)));
    }
  } else {                                      // SGD
#pragma unroll
    for (int i = 0; i < VEC; ++i) w.v[i] = __fsub_rn(w.v[i], __fmul_rn(p.lr, g.v[i]));
  }
}

// ---- cp.async (LDGSTS): global -> shared copies that hold no registers while in flight ----------------
__device__ __forceinline__ uint32_t smem_u32addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 16-byte copy, L2 only; src_bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32addr(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// 16-byte copy through L1 (.ca): for PEER memory, which bypasses the local L2 — L1 line fills ask NVLink for
// whole 128-byte lines instead of 32-byte sectors
__device__ __forceinline__ void cp_async16_ca(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32addr(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32addr(dst)), "l"(src) : "memory");
}
// VEC floats per thread (16 / 8 / 4 bytes)
// through_l1: the source may be PEER memory (see cp_async16_ca)
template <int VEC>
__device__ __forceinline__ void cp_async_vec(float* dst, const float* src, bool through_l1 = false) {
  if constexpr (VEC == 4) {
    if (through_l1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32addr(dst)), "l"(src) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32addr(dst)), "l"(src) : "memory");
  } else if constexpr (VEC == 2) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32addr(dst)), "l"(src) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32addr(dst)), "l"(src) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- index decoding --------------------------------------------------------------------------
// Folds (optional) the id with uint64 mod, adds the per-field row offset, range-checks.
struct IndexMap {
  const void* idx;
  const int64_t* field_row_offset;  // may be null
  int64_t hash_mod;                 // 0 = off
  int64_t rows;
  int32_t L;
  int32_t is64;
};

__device__ __forceinline__ int64_t load_raw_index(const void* idx, int is64, int64_t p) {
  return is64 ? __ldg(reinterpret_cast<const int64_t*>(idx) + p)
              : static_cast<int64_t>(__ldg(reinterpret_cast<const int32_t*>(idx) + p));
}

// returns the table row of position p, or -1 when out of range
__device__ __forceinline__ int64_t map_index(const IndexMap& m, int64_t p) {
  int64_t id = load_raw_index(m.idx, m.is64, p);
  if (m.hash_mod > 0) id = static_cast<int64_t>(static_cast<uint64_t>(id) % static_cast<uint64_t>(m.hash_mod));
  if (m.field_row_offset != nullptr) id += __ldg(m.field_row_offset + (p % m.L));
  return (id >= 0 && id < m.rows) ? id : -1;
}

inline IndexMap make_index_map(const void* idx, int idx_type, const int64_t* off, int64_t hash_mod, int64_t rows, int L) {
  IndexMap m;
  m.idx = idx;
  m.field_row_offset = off;
  m.hash_mod = hash_mod;
  m.rows = rows;
  m.L = L > 0 ? L : 1;
  m.is64 = (idx_type == RB_I64);
  return m;
}

// keys / positions input buffers inside a sparse-backward workspace (sparse_update.cu), filled by p2p.cu
int sparse_ws_key_buffers(int64_t n, int D, int64_t rows, void* ws, uint32_t** keys, uint32_t** vals);

// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// A kernel launched with the attribute below may be placed on the SMs while the kernel before it in the stream still drains;
// it orders itself behind that kernel's completion (and memory flush) with griddep_wait(), which EVERY kernel launched this way
// executes before it touches global memory, so that a chain of such kernels stays transitively ordered.  Without the
// attribute (or when the preceding stream operation is not a kernel) the wait returns at once.  Works inside stream capture
// (a programmatic edge of the graph).  RB_PDL=0 / rb_set_pdl(0) turn the attribute off everywhere: measured on a B200, the
// single-GPU step gains 1.7 % (r2_67: 1.328 -> 1.306 ms) and the peer-memory sharded step LOSES 1.5 % (r2_69, N = 2: 1.380 ->
// 1.400 ms), so p2p.py switches it off.
#if defined(__CUDACC__)
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// RB_PDL_RELEASE=1 (build time): small kernels (no shared memory to speak of) release their dependents right after their own
// wait, so that the next kernel's CTAs — a Dense GEMM's, say — become resident beside them and run their prologue early.
// Measured SLOWER than leaving the release to the kernel's end (r2_72, A/B on one box, twice: 1.331 / 1.342 against 1.318 /
// 1.316 ms per step): the early CTAs hold shared memory and tensor memory the other streams' kernels could have used.  Off.
#ifndef RB_PDL_RELEASE
#define RB_PDL_RELEASE 0
#endif
__device__ __forceinline__ void griddep_release() {
#if RB_PDL_RELEASE
  griddep_launch_dependents();
#endif
}

bool pdl_enabled();      // api.cu: rb_set_pdl, else the RB_PDL environment variable (default on)

template <class... KArgs, class... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), unsigned int grid, unsigned int block, size_t smem, cudaStream_t st,
                                    bool programmatic, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (programmatic && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

inline unsigned int grid_for(int64_t work_items, int items_per_block) {
  int64_t b = (work_items + items_per_block - 1) / items_per_block;
  if (b < 1) b = 1;
  return static_cast<unsigned int>(b);
}

}  // namespace rb
