// the Dense layers around the embedding path (ctr/layers.py:5-14, called at ctr/model.py:27,50,56;
// shapes ctr/train.py:74-75,82) as hand-written sm_100a GEMMs: TMA (cp.async.bulk.tensor, 128-byte swizzle) stages the
// bf16 operands in shared memory, one thread issues tcgen05.mma (M = 128, N <= 256, K = 16 per instruction) with fp32
// accumulators in tensor memory, four epilogue warps read them back with tcgen05.ld, fuse bias / activation / the bf16
// cast and hand 128-byte-wide slabs to TMA stores.  Persistent, warp-specialised, one CTA per SM:
//
//   warp 0   TMA producer     ring of kStages {A tile 128x64, B tile Nx64} stages, full/empty mbarriers
//   warp 1   MMA issuer       4 x tcgen05.mma per stage, tcgen05.commit frees the stage / publishes the accumulator
//   warp 2   TMEM allocator   512 columns = two accumulators (the epilogue of tile i overlaps the MMAs of tile i+1)
//   warp 4-11 epilogue        TMEM lane quarter q = warp % 4: thread = one row of the 128-row tile; two groups of four
//                             warps drain alternate 128-byte-wide column slabs
//
// One kernel serves the three products of a Dense layer; they differ only in which operand is "K-major" (reduction
// index contiguous in memory) and which is "MN-major" (row / column index contiguous), which the shared-memory
// descriptors and the instruction descriptor express directly — nothing is transposed in memory:
//
//   forward       y  = x . W         A = x  [rows, in]    K-major    B = W  [in, units]   MN-major   reduce over in
//   input grad    dx = dy . W^T      A = dy [rows, units] K-major    B = W  [in, units]   K-major    reduce over units
//   weight grad   dW = x^T . dy      A = x  [rows, in]    MN-major   B = dy [rows, units] MN-major   reduce over rows
//
// The weight gradient reduces over the batch: its few output tiles are split along the batch over all SMs, every split
// writes an fp32 partial tile, and a second kernel adds the partials in split order (deterministic, no float atomics).
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <cuda_bf16.h>

#include "common.cuh"

namespace rb {
namespace mlp {

constexpr int kBlockM = 128;                       // rows of C per tile = TMEM lanes
constexpr int kBlockK = 64;                        // reduction elements per stage: 64 bf16 = one 128-byte swizzle span
constexpr int kMaxN = 256;                         // columns of C per tile (tcgen05.mma N <= 256)
constexpr int kStages = 4;
constexpr int kEpilogueWarp0 = 4;
constexpr int kPrefetchAhead = 8;                    // stages the L2 prefetch warp runs ahead of the TMA producer
constexpr int kEpilogueGroups = 2;                   // groups of four warps (one per TMEM lane quarter)
constexpr int kThreads = (kEpilogueWarp0 + 4 * kEpilogueGroups) * 32;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;    // 16 KiB
constexpr int kBStageBytes = kMaxN * kBlockK * 2;      // 32 KiB
constexpr int kBResidentBytes = kStages * kBStageBytes;   // 128 KiB: the B ring's shared memory, used as one resident panel
constexpr int kSlabBytes = kBlockM * 128;              // per epilogue group: one fp32 slab (128 rows x 128 bytes) or two bf16 slabs (x 64 bytes)
constexpr int kSlabs = kEpilogueGroups;
constexpr int kAtomBytes = 64 * kBlockK * 2;           // MN-major operands: one 64-wide box of 64 reduction rows = 8 KiB
constexpr int kTmemCols = 512;
constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + kSlabs * kSlabBytes + 256 /* barriers */ + kEpilogueGroups * kMaxN * 4 /* bias */;
static_assert(kSmemBytes <= 227 * 1024, "dynamic shared memory of the Dense kernel exceeds the 227 KiB a CTA may own");

struct GemmArgs {
  int32_t M, N, K;             // C[M, N] = sum_k A[m, k] * B[n, k]
  int32_t block_n;             // columns per tile: multiple of 64, <= 256
  int32_t a_mn, b_mn;          // 1: the operand is MN-major in memory
  int32_t a_atoms, b_atoms;    // MN-major operand whose M / N extent is a multiple of 64: its map is 3-D {64, K, extent / 64} and ONE
                               // TMA instruction stages all 64-column boxes of the tile (a single thread issues ~1 TMA per 130 cycles:
                               // five per stage starved the MMAs, stats r2_09); 0: one 2-D box per instruction
  int32_t m_blocks, n_blocks, splits, k_blocks, k_blocks_per_split;
  int32_t out_f32;             // 0: bf16 C;  1: fp32 C (split s writes rows [s * m_blocks * 128, ...) of the output map)
  int32_t activation;          // rb_activation, applied after the bias
  int32_t b_resident;          // the whole B panel of the CTA's column block (K x block_n, <= kBResidentBytes) is loaded ONCE and stays in
                               // shared memory; only A streams through the ring (single-CTA form, no split): small-K products
  int32_t pf_a, pf_b;          // opt-in experiment (RB_DENSE_PREFETCH=1): a spare warp prefetches the operand's boxes into L2
                               // kPrefetchAhead stages ahead of the TMA loads; see prefetch_enabled() for why it is off
  const float* bias;           // f32[N] or null
  unsigned long long* stats;   // diagnostic (rb_dense_debug_stats): per CTA, cycles each role spent waiting; null = off
};

// ---- PTX: mbarrier, TMA, tcgen05 ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t saddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
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
      : "r"(saddr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// A wait that cannot hang the GPU: a protocol error (a phase that never completes) traps after ~2 s instead of
// spinning until the watchdog resets the device for everybody.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) {
        printf("rb::mlp: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}

// One lane of a CONVERGED warp.  The producer and the MMA issuer run their loops with all 32 lanes and predicate only the
// TMA / tcgen05 instructions on the elected lane: operands that are warp-uniform then live in uniform registers.  (Under
// `if (lane == 0)` the compiler wraps every UTCHMMA / UTMALDG in an ELECT retry loop fed by R2UR moves — ~130 cycles per
// issue, SASS of r2_18 — which made a 128-cycle MMA issue-bound.)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(saddr(dst)),
               "l"(map), "r"(saddr(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(saddr(dst)),
               "l"(map), "r"(saddr(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(saddr(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 operands, fp32 accumulate; one thread issues for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (asynchronous: tmem_ld_wait() before the registers are read)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (tcgen05), 128-byte swizzle.  Offsets are in bytes, encoded without their 4 LSBs.
//   K-major operand:  rows of 128 bytes (64 bf16 of the reduction axis), 8-row groups 1024 bytes apart (SBO); one
//                     instruction consumes 32 bytes of every row -> the start address advances by 32 per k-step.
//   MN-major operand: boxes of 64 reduction rows x 128 bytes (64 bf16 of the M/N axis); 8-row groups 1024 bytes apart
//                     (SBO), the next 64 columns of M/N one box (8 KiB) further (LBO); one instruction consumes 16 rows
//                     -> the start address advances by 2048 per k-step.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  d |= 2ull << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float apply_activation(float x, int act) {
  if (act == RB_ACT_RELU) return fmaxf(x, 0.f);
  if (act == RB_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(-x));
  return x;
}

// 32 accumulator columns of one row -> (+ bias, activation) -> the slab row in shared memory: 128 bytes of fp32 (128-byte
// swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)) or 64 bytes of bf16 (64-byte swizzle: chunk c ^ ((r >> 1) & 3)).
// `sw` is that row term; s_bias points at the 32 biases in shared memory.
template <bool F32, int ACT>
__device__ __forceinline__ void store_columns(const uint32_t (&v)[32], const float* s_bias, uint8_t* srow, int sw) {
  constexpr int chunk0 = 0;
  uint32_t pk[2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b = *reinterpret_cast<const float4*>(s_bias + 4 * j);      // same address in every lane: a broadcast
    const float f0 = apply_activation(__uint_as_float(v[4 * j]) + b.x, ACT), f1 = apply_activation(__uint_as_float(v[4 * j + 1]) + b.y, ACT);
    const float f2 = apply_activation(__uint_as_float(v[4 * j + 2]) + b.z, ACT), f3 = apply_activation(__uint_as_float(v[4 * j + 3]) + b.w, ACT);
    if constexpr (F32) {
      *reinterpret_cast<float4*>(srow + (((chunk0 + j) ^ sw) << 4)) = make_float4(f0, f1, f2, f3);
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(f0, f1), hi = __floats2bfloat162_rn(f2, f3);
      if ((j & 1) == 0) {       // two column quads make one 16-byte chunk
        pk[0] = *reinterpret_cast<uint32_t*>(&lo);
        pk[1] = *reinterpret_cast<uint32_t*>(&hi);
      } else {
        *reinterpret_cast<uint4*>(srow + (((chunk0 + (j >> 1)) ^ sw) << 4)) =
            make_uint4(pk[0], pk[1], *reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
  }
}

// ---- the 2-CTA form (cta_group::2) ---------------------------------------------------------------------------------------
// A CTA pair (the two SMs of a TPC, a cluster of 2) computes a 256 x N tile: CTA r owns rows [128 r, 128 r + 128) of A and of
// the accumulator, and stages HALF of the B tile (N / 2 of its rows / columns); the leader's tcgen05.mma.cta_group::2 reads
// both halves, so every byte of B crosses L2 -> SM once per 256 rows of C instead of once per 128.  At 128 x 256 tiles one
// CTA alone needs 96 B / clk of operand fill against ~3000 cycles of L2 / DRAM latency — more bytes in flight than its
// shared memory holds (ncu r2_04: tensor pipe 55 % busy, HBM 32 %, L2 35 %: waiting on data); the pair needs 64 B / clk.
//   * both producers issue cp.async.bulk.tensor ... .cta_group::2 whose completion bytes land on the LEADER's full barrier;
//     the leader's producer arms it with the bytes of both CTAs;
//   * the leader's tcgen05.commit ... .multicast::cluster frees the stage / publishes the accumulator in BOTH CTAs;
//   * both CTAs' epilogue warps arrive on the leader's acc_empty barrier (remote arrive through mapa).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;       // shared::cluster address of the same offset in the pair's even (leader) CTA

__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(saddr(dst)),
      "l"(map), "r"(saddr(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          saddr(dst)),
      "l"(map), "r"(saddr(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(saddr(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(saddr(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

// CG = CTAs per tile (1, or 2 = a cta_group::2 pair launched as a cluster of 2)
template <int CG, bool F32, int ACT>
__device__ __forceinline__ void dense_gemm_body(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_c,
                                                const GemmArgs& g) {
  constexpr int kNumStages = CG == 2 ? 6 : kStages;
  constexpr int kBBytes = kBStageBytes / CG;              // this CTA's share of a B stage
  extern __shared__ __align__(1024) uint8_t smem[];      // 128-byte-swizzled tiles need 1024-byte aligned bases
  if ((saddr(smem) & 1023u) != 0) __trap();
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + kNumStages * kAStageBytes;
  uint8_t* smem_out = smem_b + kNumStages * kBBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_out + kSlabs * kSlabBytes);
  uint64_t* empty = full + kNumStages;
  uint64_t* acc_full = empty + kNumStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_full = acc_empty + 2;                     // B-resident form: the panel has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  volatile int* produced = reinterpret_cast<volatile int*>(tmem_slot + 1);      // stages the TMA producer has issued (paces the prefetch warp)
  float* s_bias = reinterpret_cast<float*>(smem_out + kSlabs * kSlabBytes + 256);   // the tile's kMaxN biases, one copy per group

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile_id0 = blockIdx.x / CG, tile_stride = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kNumStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    *produced = 0;
    mbar_init(b_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4 * kEpilogueGroups * CG);     // one arrival per epilogue warp of the tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();      // the peer's barriers exist before anything is signalled on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (descriptor prefetch, barriers, tensor-memory allocation) may run while the kernel before this one in the
  // stream drains (programmatic dependent launch, common.cuh); nothing below may
  griddep_wait();

  const int total = g.m_blocks * g.n_blocks * g.splits;

  if (warp == 0) {
    // ===== TMA producer (converged warp, one elected lane issues) ================================================
    {
      const bool leader = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      long long t_empty = 0;
      int issued = 0;
      const long long t_begin = clock64();
      if (CG == 1 && g.b_resident) {
        // every tile of this CTA has the same column block (the grid is a multiple of n_blocks): its B panel is loaded once,
        // k-block after k-block, in the layout a stage would have
        const int n0 = (tile_id0 % g.n_blocks) * g.block_n;
        const uint32_t tile_bytes = static_cast<uint32_t>(g.block_n) * 128u;
        if (leader) {
          mbar_expect_tx(b_full, tile_bytes * static_cast<uint32_t>(g.k_blocks));
          for (int kb = 0; kb < g.k_blocks; ++kb) {
            uint8_t* dst = smem_b + static_cast<size_t>(kb) * tile_bytes;
            if (!g.b_mn) tma_load_2d(dst, &map_b, b_full, kb * kBlockK, n0);
            else if (g.b_atoms) tma_load_3d(dst, &map_b, b_full, 0, kb * kBlockK, n0 / 64);
            else
              for (int j = 0; j < g.block_n / 64; ++j) tma_load_2d(dst + j * kAtomBytes, &map_b, b_full, n0 + 64 * j, kb * kBlockK);
          }
        }
        __syncwarp();
      }
      for (int w = tile_id0; w < total; w += tile_stride) {
        const int n_blk = w % g.n_blocks, m_blk = (w / g.n_blocks) % g.m_blocks, split = w / (g.n_blocks * g.m_blocks);
        const int m0 = m_blk * (kBlockM * CG) + rank * kBlockM, n0 = n_blk * g.block_n;
        const int kb0 = split * g.k_blocks_per_split, kb1 = min(kb0 + g.k_blocks_per_split, g.k_blocks);
        // a tile that hangs over the right edge of C multiplies only the columns that exist (rounded up to the MMA's N step);
        // this CTA stages share `rank` of them
        const int n_round = (g.b_atoms ? 64 : 16) * CG;
        const int n_eff = min(g.block_n, (g.N - n0 + n_round - 1) / n_round * n_round);
        const int n_mine = n_eff / CG, nb0 = n0 + rank * n_mine;
        const int b_boxes = (n_mine + 63) / 64;
        const uint32_t my_bytes = kAStageBytes + ((g.b_mn && !g.b_atoms) ? static_cast<uint32_t>(b_boxes) * kAtomBytes
                                                                         : static_cast<uint32_t>(g.block_n / CG) * 128u);
        for (int kb = kb0; kb < kb1; ++kb) {
          const long long t0 = clock64();
          mbar_wait(&empty[stage], phase ^ 1);
          t_empty += clock64() - t0;
          const int k0 = kb * kBlockK;
          uint8_t* a_dst = smem_a + stage * kAStageBytes;
          uint8_t* b_dst = smem_b + stage * kBBytes;
          if (leader) {
          if constexpr (CG == 2) {
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * my_bytes);      // both CTAs stage the same number of bytes
            if (!g.a_mn) {
              tma_load_2d_pair(a_dst, &map_a, &full[stage], k0, m0);
            } else if (g.a_atoms) {
              tma_load_3d_pair(a_dst, &map_a, &full[stage], 0, k0, m0 / 64);
            } else {
              tma_load_2d_pair(a_dst, &map_a, &full[stage], m0, k0);
              tma_load_2d_pair(a_dst + kAtomBytes, &map_a, &full[stage], m0 + 64, k0);
            }
            if (!g.b_mn) {
              tma_load_2d_pair(b_dst, &map_b, &full[stage], k0, nb0);
            } else if (g.b_atoms) {
              tma_load_3d_pair(b_dst, &map_b, &full[stage], 0, k0, nb0 / 64);
            } else {
              for (int j = 0; j < b_boxes; ++j) tma_load_2d_pair(b_dst + j * kAtomBytes, &map_b, &full[stage], nb0 + 64 * j, k0);
            }
          } else {
            mbar_expect_tx(&full[stage], g.b_resident ? static_cast<uint32_t>(kAStageBytes) : my_bytes);
            if (!g.a_mn) {
              tma_load_2d(a_dst, &map_a, &full[stage], k0, m0);
            } else if (g.a_atoms) {
              tma_load_3d(a_dst, &map_a, &full[stage], 0, k0, m0 / 64);
            } else {
              tma_load_2d(a_dst, &map_a, &full[stage], m0, k0);
              tma_load_2d(a_dst + kAtomBytes, &map_a, &full[stage], m0 + 64, k0);
            }
            if (g.b_resident) {
              // B is already there
            } else if (!g.b_mn) {
              tma_load_2d(b_dst, &map_b, &full[stage], k0, nb0);
            } else if (g.b_atoms) {
              tma_load_3d(b_dst, &map_b, &full[stage], 0, k0, nb0 / 64);
            } else {
              for (int j = 0; j < b_boxes; ++j) tma_load_2d(b_dst + j * kAtomBytes, &map_b, &full[stage], nb0 + 64 * j, k0);
            }
          }
          }
          __syncwarp();
          if (++stage == kNumStages) {
            stage = 0;
            phase ^= 1;
          }
          ++issued;
          if (leader) *produced = issued;
        }
      }
      if (g.stats != nullptr && leader) {
        g.stats[blockIdx.x * 8 + 3] = t_empty;
        g.stats[blockIdx.x * 8 + 4] = clock64() - t_begin;
      }
    }
  } else if (warp == 3) {
    // ===== L2 prefetch of the streamed operands, kPrefetchAhead stages ahead of the TMA loads ===================================
    if (lane == 0 && (g.pf_a || g.pf_b)) {
      int ahead = 0;
      for (int w = tile_id0; w < total; w += tile_stride) {
        const int n_blk = w % g.n_blocks, m_blk = (w / g.n_blocks) % g.m_blocks, split = w / (g.n_blocks * g.m_blocks);
        const int m0 = m_blk * (kBlockM * CG) + rank * kBlockM, n0 = n_blk * g.block_n;
        const int kb0 = split * g.k_blocks_per_split, kb1 = min(kb0 + g.k_blocks_per_split, g.k_blocks);
        const int n_round = (g.b_atoms ? 64 : 16) * CG;
        const int n_eff = min(g.block_n, (g.N - n0 + n_round - 1) / n_round * n_round);
        const int n_mine = n_eff / CG, nb0 = n0 + rank * n_mine;
        const int b_boxes = (n_mine + 63) / 64;
        for (int kb = kb0; kb < kb1; ++kb, ++ahead) {
          while (ahead >= *produced + kPrefetchAhead) __nanosleep(64);
          const int k0 = kb * kBlockK;
          if (g.pf_a) {
            if (!g.a_mn) tma_prefetch_2d(&map_a, k0, m0);
            else if (g.a_atoms) tma_prefetch_3d(&map_a, 0, k0, m0 / 64);
            else {
              tma_prefetch_2d(&map_a, m0, k0);
              tma_prefetch_2d(&map_a, m0 + 64, k0);
            }
          }
          if (g.pf_b) {
            if (!g.b_mn) tma_prefetch_2d(&map_b, k0, nb0);
            else if (g.b_atoms) tma_prefetch_3d(&map_b, 0, k0, nb0 / 64);
            else
              for (int j = 0; j < b_boxes; ++j) tma_prefetch_2d(&map_b, nb0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (the pair's leader CTA only; converged warp, one elected lane issues) ==========================
    if (rank == 0) {
      const bool leader = elect_one();
      // instruction descriptor: D fp32, A/B bf16, majors, N >> 3, M >> 4
      const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(g.a_mn) << 15) |
                              (static_cast<uint32_t>(g.b_mn) << 16) | (static_cast<uint32_t>((kBlockM * CG) >> 4) << 24);
      const uint32_t a_lbo = g.a_mn ? kAtomBytes : 16, b_lbo = g.b_mn ? kAtomBytes : 16;
      // descriptors of k-step 0 of stage 0; a k-step advances the 14-bit start-address field (16-byte units) by 2 (K-major:
      // 32 bytes along the swizzled row) or 128 (MN-major: 16 rows of 128 bytes), a stage by its size
      const uint64_t a_desc0 = smem_desc(saddr(smem_a), a_lbo, 1024), b_desc0 = smem_desc(saddr(smem_b), b_lbo, 1024);
      const uint32_t a_kstep = g.a_mn ? 128u : 2u, b_kstep = g.b_mn ? 128u : 2u;
      long long t_full = 0, t_acc = 0;
      const long long t_begin = clock64();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t b_tile16 = static_cast<uint32_t>(g.block_n) * 128u >> 4;      // B-resident: k-block stride of the panel
      if (CG == 1 && g.b_resident) {
        mbar_wait(b_full, 0);
        tc_fence_after();
      }
      for (int w = tile_id0; w < total; w += tile_stride, ++it) {
        const int split = w / (g.n_blocks * g.m_blocks);
        const int kb0 = split * g.k_blocks_per_split, kb1 = min(kb0 + g.k_blocks_per_split, g.k_blocks);
        const int n_round = (g.b_atoms ? 64 : 16) * CG;
        const int n_eff = min(g.block_n, (g.N - (w % g.n_blocks) * g.block_n + n_round - 1) / n_round * n_round);
        const uint32_t idesc = idesc0 | (static_cast<uint32_t>(n_eff >> 3) << 17);   // N of this tile's MMAs
        const int acc = it & 1;
        long long t0 = clock64();
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator
        t_acc += clock64() - t0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kMaxN);
        for (int kb = kb0; kb < kb1; ++kb) {
          t0 = clock64();
          mbar_wait(&full[stage], phase);
          t_full += clock64() - t0;
          tc_fence_after();
          const uint64_t ad = a_desc0 + static_cast<uint64_t>(stage * (kAStageBytes >> 4));
          const uint64_t bd = b_desc0 + static_cast<uint64_t>(g.b_resident ? (kb - kb0) * b_tile16 : stage * (kBBytes >> 4));
          const int k_valid = g.K - kb * kBlockK;
          if (leader) {
            if (k_valid >= kBlockK) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if constexpr (CG == 2) umma_bf16_pair(tmem_d, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                else umma_bf16(tmem_d, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              }
            } else {                                              // the ragged end of the reduction axis
              const int k_steps = (k_valid + 15) / 16;
              for (int k = 0; k < k_steps; ++k) {
                if constexpr (CG == 2) umma_bf16_pair(tmem_d, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                else umma_bf16(tmem_d, ad + k * a_kstep, bd + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              }
            }
            // the stage is free (in both CTAs) once these MMAs have read it
            if constexpr (CG == 2) umma_commit_pair(&empty[stage]);
            else umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == kNumStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // the accumulator (each CTA's 128 rows of it) is complete
        if (leader) {
          if constexpr (CG == 2) umma_commit_pair(&acc_full[acc]);
          else umma_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
      if (g.stats != nullptr && leader) {
        g.stats[blockIdx.x * 8 + 0] = t_full;
        g.stats[blockIdx.x * 8 + 1] = t_acc;
        g.stats[blockIdx.x * 8 + 2] = clock64() - t_begin;
      }
    }
  } else if (warp >= kEpilogueWarp0) {
    // ===== epilogue: TMEM -> registers -> (bias, activation, cast) -> swizzled smem slab -> TMA store ===========
    // Two groups of four warps; group g drains the slabs of parity g of every tile through its own slab buffer, bias copy
    // and named barrier, so two slabs are in flight per tile.
    const int ew = warp - kEpilogueWarp0;
    const int q = ew & 3;                                       // == warp % 4: the TMEM lane quarter this warp may read
    const int grp = ew >> 2;
    const int row = q * 32 + lane;
    const int etid = q * 32 + lane;                             // thread index inside the group
    const int bar_id = 1 + grp;
    float* bias_g = s_bias + grp * kMaxN;
    // a slab = 32 output columns of the tile's 128 rows: fp32 -> 128-byte rows (16 KiB, one buffer per group), bf16 -> 64-byte rows
    // (8 KiB, TWO buffers per group: the TMA store of one slab drains while the next is written)
    constexpr int cols_per_slab = 32;
    constexpr int kRowBytes = F32 ? 128 : 64;
    constexpr int kBufs = F32 ? 1 : 2;
    uint8_t* slab0 = smem_out + grp * kSlabBytes;
    const int sw = F32 ? (row & 7) : ((row >> 1) & 3);
    uint32_t slab_seq = 0;
    const int n_slabs = g.block_n / cols_per_slab;
    long long t_wait = 0;
    const long long t_begin = clock64();
    int it = 0;
    for (int w = tile_id0; w < total; w += tile_stride, ++it) {
      const int n_blk = w % g.n_blocks, m_blk = (w / g.n_blocks) % g.m_blocks, split = w / (g.n_blocks * g.m_blocks);
      const int n0 = n_blk * g.block_n;
      const int out_row0 = (F32 ? split * g.m_blocks * (kBlockM * CG) : 0) + m_blk * (kBlockM * CG) + rank * kBlockM;
      const int acc = it & 1;
      // the tile's biases -> this group's shared-memory copy (zeros without a bias / past the edge of C); ordered before their
      // first use by the bar.sync that opens the group's first slab, and after the previous tile's last use by the bar.sync
      // that closed its last slab
      for (int c = etid; c < g.block_n; c += 128) bias_g[c] = (g.bias != nullptr && n0 + c < g.N) ? __ldg(g.bias + n0 + c) : 0.f;
      const long long t0 = clock64();
      mbar_wait(&acc_full[acc], (it >> 1) & 1);
      t_wait += clock64() - t0;
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kMaxN);
      for (int s = grp; s < n_slabs; s += kEpilogueGroups) {
        const int c0 = n0 + s * cols_per_slab;                  // first output column of the slab
        if (c0 >= g.N) break;                                   // uniform over the group: the tile hangs over the edge of C
        uint32_t v0[32];
        tmem_ld32(t_row + static_cast<uint32_t>(s * cols_per_slab), v0);          // in flight across the waits below
        uint8_t* slab = slab0 + (slab_seq % kBufs) * (kBlockM * kRowBytes);
        ++slab_seq;
        if (etid == 0) tma_store_wait_read<kBufs - 1>();        // the store that last read this slab buffer is done with it
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        tmem_ld_wait();
        store_columns<F32, ACT>(v0, bias_g + s * cols_per_slab, slab + row * kRowBytes, sw);
        fence_proxy_async();                                    // generic-proxy writes -> visible to the TMA engine
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (etid == 0) {
          tma_store_2d(&map_c, slab, c0, out_row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                          // this warp's share of the accumulator is drained
        if constexpr (CG == 2) mbar_arrive_leader(&acc_empty[acc]);
        else mbar_arrive(&acc_empty[acc]);
      }
    }
    if (etid == 0) tma_store_wait_all();
    if (g.stats != nullptr && etid == 0 && grp == 0) {
      g.stats[blockIdx.x * 8 + 5] = t_wait;
      g.stats[blockIdx.x * 8 + 6] = clock64() - t_begin;
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();      // no CTA leaves while its peer may still read its shared memory or signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <bool F32, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_c, const __grid_constant__ GemmArgs g) {
  dense_gemm_body<1, F32, ACT>(map_a, map_b, map_c, g);
}

template <bool F32, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
dense_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_c, const __grid_constant__ GemmArgs g) {
  dense_gemm_body<2, F32, ACT>(map_a, map_b, map_c, g);
}

// dW[m, n] = sum over splits (in split order) of the fp32 partial tiles; eight loads in flight per thread
__global__ void __launch_bounds__(256)
reduce_splits_kernel(const float* __restrict__ part, int splits, int64_t split_stride, int M, int N, float* __restrict__ out, int64_t ldo) {
  griddep_wait();
  griddep_release();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int nv = N / 4;
  if (i >= static_cast<int64_t>(M) * nv) return;
  const int m = static_cast<int>(i / nv), n = static_cast<int>(i % nv) * 4;
  const float* p = part + static_cast<int64_t>(m) * N + n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = 0; s0 < splits; s0 += 8) {
    float4 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      t[u] = (s0 + u < splits) ? __ldcs(reinterpret_cast<const float4*>(p + (s0 + u) * split_stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x += t[u].x;
      acc.y += t[u].y;
      acc.z += t[u].z;
      acc.w += t[u].w;
    }
  }
  *reinterpret_cast<float4*>(out + static_cast<int64_t>(m) * ldo + n) = acc;
}

// ---- the Dense(1) head (ctr/train.py:75,82: the last unit of [512, 256, 1]) on CUDA cores: a row dot product ----------
// z[r] = sum_k x[r, k] * w[k] + b;  out[r] = act(z[r]).  One warp per row, 16-byte loads, fp32 accumulation.
__global__ void __launch_bounds__(256)
head_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int in_dim, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                const float* __restrict__ bias, int act, float* __restrict__ out) {
  griddep_wait();
  griddep_release();
  const int lane = threadIdx.x % 32;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + threadIdx.x / 32;
  if (r >= rows) return;
  float acc = 0.f;
  for (int k = lane * 8; k < in_dim; k += 256) {
    const uint4 xv = __ldcs(reinterpret_cast<const uint4*>(x + r * ldx + k));
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + k));
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc = fmaf(__uint_as_float(xs[j] << 16), __uint_as_float(ws[j] << 16), acc);
      acc = fmaf(__uint_as_float(xs[j] & 0xFFFF0000u), __uint_as_float(ws[j] & 0xFFFF0000u), acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[r] = apply_activation(acc + (bias != nullptr ? __ldg(bias) : 0.f), act);
}

// Narrow heads (in_dim <= 64, e.g. the 40 -> 1 logit of DIN's attention unit, dien/layers.py:39): a warp per row would leave most
// lanes idle, so one THREAD takes a row — its in_dim / 8 16-byte loads are contiguous and neighbouring threads read neighbouring rows.
__global__ void __launch_bounds__(256)
head_fwd_narrow_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int in_dim, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                       const float* __restrict__ bias, int act, float* __restrict__ out) {
  griddep_wait();
  griddep_release();
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (r >= rows) return;
  float acc = 0.f;
  for (int k = 0; k < in_dim; k += 8) {
    const uint4 xv = __ldcs(reinterpret_cast<const uint4*>(x + r * ldx + k));
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + k));
    const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc = fmaf(__uint_as_float(xs[j] << 16), __uint_as_float(ws[j] << 16), acc);
      acc = fmaf(__uint_as_float(xs[j] & 0xFFFF0000u), __uint_as_float(ws[j] & 0xFFFF0000u), acc);
    }
  }
  out[r] = apply_activation(acc + (bias != nullptr ? __ldg(bias) : 0.f), act);
}

// Backward of the head.  dz[r] = dout[r] * act'(out[r]);  dx[r, :] = bf16(dz[r] * w[:]);  partial sums of
// dW[k] = sum_r x[r, k] * dz[r] and db = sum_r dz[r] per CTA (slab of rows), reduced in CTA order by head_bwd_final_kernel.
constexpr int kHeadBwdThreads = 256;
__global__ void __launch_bounds__(kHeadBwdThreads)
head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, int act, const __nv_bfloat16* __restrict__ x, int64_t rows,
                int in_dim, int64_t ldx, const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ dx, int64_t lddx,
                float* __restrict__ partial /* [grid, 2 * in_dim + 1]: dW | db | column sums of dx */) {
  griddep_wait();
  griddep_release();
  // thread = (row lane rl, vector column vc): vc covers 8 consecutive columns
  const int vcols = in_dim / 8;
  const int row_lanes = kHeadBwdThreads / vcols;
  const int vc = threadIdx.x % vcols, rl = threadIdx.x / vcols;
  __shared__ float s_red[kHeadBwdThreads * 8];
  __shared__ float s_db[kHeadBwdThreads];
  const int pstride = 2 * in_dim + 1;
  float acc[8], cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = cs[k] = 0.f;
  float db = 0.f;
  if (rl < row_lanes) {
    float wf[8];
    {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + vc * 8));
      const uint32_t ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        wf[2 * j] = __uint_as_float(ws[j] << 16);
        wf[2 * j + 1] = __uint_as_float(ws[j] & 0xFFFF0000u);
      }
    }
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * per, r1 = min(r0 + per, rows);
    for (int64_t r = r0 + rl; r < r1; r += row_lanes) {
      float dz = dout[r];
      const float o = act != RB_ACT_NONE ? out[r] : 0.f;      // a head without activation has no forward output to read
      if (act == RB_ACT_SIGMOID) dz = dz * o * (1.0f - o);
      else if (act == RB_ACT_RELU) dz = o > 0.f ? dz : 0.f;
      if (vc == 0) db += dz;
      const uint4 xv = __ldcs(reinterpret_cast<const uint4*>(x + r * ldx + vc * 8));
      const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] = fmaf(__uint_as_float(xs[j] << 16), dz, acc[2 * j]);
        acc[2 * j + 1] = fmaf(__uint_as_float(xs[j] & 0xFFFF0000u), dz, acc[2 * j + 1]);
        __nv_bfloat162 h = __floats2bfloat162_rn(dz * wf[2 * j], dz * wf[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
        cs[2 * j] += __uint_as_float(pk[j] << 16);            // what the next layer's bias gradient sums: the ROUNDED dx
        cs[2 * j + 1] += __uint_as_float(pk[j] & 0xFFFF0000u);
      }
      if (dx != nullptr) __stcs(reinterpret_cast<uint4*>(dx + r * lddx + vc * 8), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = acc[k];
  s_db[threadIdx.x] = db;
  __syncthreads();
  for (int c = threadIdx.x; c < in_dim; c += kHeadBwdThreads) {
    float t = 0.f;
    for (int l = 0; l < row_lanes; ++l) t += s_red[(l * vcols + c / 8) * 8 + (c % 8)];
    partial[static_cast<int64_t>(blockIdx.x) * pstride + c] = t;
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int l = 0; l < row_lanes; ++l) t += s_db[l * vcols];
    partial[static_cast<int64_t>(blockIdx.x) * pstride + in_dim] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = cs[k];
  __syncthreads();
  for (int c = threadIdx.x; c < in_dim; c += kHeadBwdThreads) {
    float t = 0.f;
    for (int l = 0; l < row_lanes; ++l) t += s_red[(l * vcols + c / 8) * 8 + (c % 8)];
    partial[static_cast<int64_t>(blockIdx.x) * pstride + in_dim + 1 + c] = t;
  }
}

// one warp per output column: lane l adds partials l, l + 32, ... (independent loads), then a fixed shuffle tree
// Dense(1) + sigmoid, the clipped binary cross-entropy of its output and the head's backward in ONE pass over x (ctr/model.py:56-57,
// ctr/train.py:85-87,97 for the model whose last layer is Dense(1, sigmoid)): the three-kernel sequence head_fwd -> bce -> head_bwd
// reads x twice and sits on the step's critical path with ~8 small launches (50 us of a 1.33 ms step, profiles/r2_20_timeline.txt).
// Same thread layout, row partition and accumulation order as head_bwd_kernel, same dot-product order as head_fwd_kernel, same
// loss / dprob formulas as bce_partial_kernel (dense.cu): prob, dx, dW, db and the dx column sums come out bit-identical to the
// unfused sequence; the loss differs only in the order its per-row terms are added.  The backward is seeded with d loss = 1.
// vcols = in_dim / 8 must be a power of two <= 32 (a row's threads are lanes of one warp).
__global__ void __launch_bounds__(kHeadBwdThreads)
head_bce_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int in_dim, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                const float* __restrict__ bias, const void* __restrict__ label, int label_is_i64, float inv_n, float* __restrict__ prob,
                __nv_bfloat16* __restrict__ dx, int64_t lddx, float* __restrict__ partial /* [grid, 2 * in_dim + 2]: dW | db | colsums | loss */) {
  griddep_wait();
  griddep_release();
  const int vcols = in_dim / 8;
  const int row_lanes = kHeadBwdThreads / vcols;
  const int vc = threadIdx.x % vcols, rl = threadIdx.x / vcols;
  __shared__ float s_red[kHeadBwdThreads * 8];
  __shared__ float s_db[kHeadBwdThreads];
  __shared__ float s_loss[kHeadBwdThreads];
  const int pstride = 2 * in_dim + 2;
  const float eps = 1e-7f;
  float acc[8], cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = cs[k] = 0.f;
  float db = 0.f, loss = 0.f;
  float wf[8];
  {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + vc * 8));
    const uint32_t ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wf[2 * j] = __uint_as_float(ws[j] << 16);
      wf[2 * j + 1] = __uint_as_float(ws[j] & 0xFFFF0000u);
    }
  }
  const float b0 = bias != nullptr ? __ldg(bias) : 0.f;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * per, r1 = min(r0 + per, rows);
  // every lane of a warp runs the same number of iterations (rows of one warp are r, r + 1, ... for its 32 / vcols groups):
  // the shuffles below need the full warp
  const int64_t n_iter = (r1 - r0 + row_lanes - 1) / row_lanes;
  constexpr int kU = 4;                                   // rows whose x (and label) loads are in flight together per thread
  for (int64_t it0 = 0; it0 < n_iter; it0 += kU) {
    uint32_t xs[kU][4];
    float yv[kU];
    bool live[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t r = r0 + rl + (it0 + u) * row_lanes;
      live[u] = (it0 + u < n_iter) && r < r1;
      xs[u][0] = xs[u][1] = xs[u][2] = xs[u][3] = 0u;
      yv[u] = 0.f;
      if (live[u]) {
        const uint4 xv = __ldcs(reinterpret_cast<const uint4*>(x + r * ldx + vc * 8));
        xs[u][0] = xv.x; xs[u][1] = xv.y; xs[u][2] = xv.z; xs[u][3] = xv.w;
        yv[u] = label_is_i64 ? static_cast<float>(static_cast<const int64_t*>(label)[r]) : static_cast<const float*>(label)[r];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (it0 + u >= n_iter) break;                       // uniform over the CTA
      const int64_t r = r0 + rl + (it0 + u) * row_lanes;
      float z = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        z = fmaf(__uint_as_float(xs[u][j] << 16), wf[2 * j], z);
        z = fmaf(__uint_as_float(xs[u][j] & 0xFFFF0000u), wf[2 * j + 1], z);
      }
      for (int o = 16; o > 0; o >>= 1) {                 // offsets >= vcols add lanes of other rows: skipped
        const float t = __shfl_xor_sync(0xffffffffu, z, o);
        if (o < vcols) z += t;
      }
      if (!live[u]) continue;
      const float p = apply_activation(z + b0, RB_ACT_SIGMOID);
      const float y = yv[u];
      const float pc = fminf(fmaxf(p, eps), 1.0f - eps);
      const float a = pc + eps, bq = (1.0f - pc) + eps;
      const bool pass = (p >= eps) && (p <= 1.0f - eps);
      const float dprob = pass ? (-(y / a) + (1.0f - y) / bq) * inv_n : 0.f;
      const float dz = dprob * p * (1.0f - p);
      if (vc == 0) {
        prob[r] = p;
        loss += -(y * logf(a) + (1.0f - y) * logf(bq));
        db += dz;
      }
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] = fmaf(__uint_as_float(xs[u][j] << 16), dz, acc[2 * j]);
        acc[2 * j + 1] = fmaf(__uint_as_float(xs[u][j] & 0xFFFF0000u), dz, acc[2 * j + 1]);
        __nv_bfloat162 h = __floats2bfloat162_rn(dz * wf[2 * j], dz * wf[2 * j + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&h);
        cs[2 * j] += __uint_as_float(pk[j] << 16);
        cs[2 * j + 1] += __uint_as_float(pk[j] & 0xFFFF0000u);
      }
      if (dx != nullptr) __stcs(reinterpret_cast<uint4*>(dx + r * lddx + vc * 8), make_uint4(pk[0], pk[1], pk[2], pk[3]));
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = acc[k];
  s_db[threadIdx.x] = db;
  s_loss[threadIdx.x] = loss;
  __syncthreads();
  for (int c = threadIdx.x; c < in_dim; c += kHeadBwdThreads) {
    float t = 0.f;
    for (int l = 0; l < row_lanes; ++l) t += s_red[(l * vcols + c / 8) * 8 + (c % 8)];
    partial[static_cast<int64_t>(blockIdx.x) * pstride + c] = t;
  }
  if (threadIdx.x == 0) {
    float t = 0.f, tl = 0.f;
    for (int l = 0; l < row_lanes; ++l) {
      t += s_db[l * vcols];
      tl += s_loss[l * vcols];
    }
    partial[static_cast<int64_t>(blockIdx.x) * pstride + in_dim] = t;
    partial[static_cast<int64_t>(blockIdx.x) * pstride + 2 * in_dim + 1] = tl;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[threadIdx.x * 8 + k] = cs[k];
  __syncthreads();
  for (int c = threadIdx.x; c < in_dim; c += kHeadBwdThreads) {
    float t = 0.f;
    for (int l = 0; l < row_lanes; ++l) t += s_red[(l * vcols + c / 8) * 8 + (c % 8)];
    partial[static_cast<int64_t>(blockIdx.x) * pstride + in_dim + 1 + c] = t;
  }
}

// partial[parts][2 * in_dim + 2] -> dW | db | dx column sums | loss (mean), CTA order
__global__ void __launch_bounds__(256)
head_bce_final_kernel(const float* __restrict__ partial, int parts, int in_dim, float inv_n, float* __restrict__ dW, float* __restrict__ db,
                      float* __restrict__ dx_colsum, float* __restrict__ loss) {
  griddep_wait();
  griddep_release();
  const int c = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  const int pstride = 2 * in_dim + 2;
  if (c >= pstride) return;
  float t = 0.f;
  for (int p = lane; p < parts; p += 32) t += partial[static_cast<int64_t>(p) * pstride + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane != 0) return;
  if (c < in_dim) dW[c] = t;
  else if (c == in_dim) db[0] = t;
  else if (c == 2 * in_dim + 1) loss[0] = t * inv_n;
  else if (dx_colsum != nullptr) dx_colsum[c - in_dim - 1] = t;
}

__global__ void __launch_bounds__(256)
head_bwd_final_kernel(const float* __restrict__ partial, int parts, int in_dim, float* __restrict__ dW, float* __restrict__ db,
                      float* __restrict__ dx_colsum) {
  griddep_wait();
  griddep_release();
  const int c = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  const int pstride = 2 * in_dim + 1;
  if (c >= pstride) return;
  float t = 0.f;
  for (int p = lane; p < parts; p += 32) t += partial[static_cast<int64_t>(p) * pstride + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane != 0) return;
  if (c < in_dim) dW[c] = t;
  else if (c == in_dim) db[0] = t;
  else if (dx_colsum != nullptr) dx_colsum[c - in_dim - 1] = t;
}

// dy_pre[r, c] = bf16(dy[r, c] * act'(y[r, c])): the activation's backward in front of the last layer's GEMMs
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act, int64_t n4, __nv_bfloat16* __restrict__ out) {
  griddep_wait();
  griddep_release();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 d = __ldcs(reinterpret_cast<const float4*>(dy) + i);
  float v[4] = {d.x, d.y, d.z, d.w};
  if (act != RB_ACT_NONE) {
    const float4 o = __ldcs(reinterpret_cast<const float4*>(y) + i);
    const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = act == RB_ACT_RELU ? (ov[k] > 0.f ? v[k] : 0.f) : v[k] * ov[k] * (1.0f - ov[k]);
  }
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// the same for a hidden layer whose output and incoming gradient are bf16 (dien/layers.py:37-38): 8 elements per thread
__global__ void __launch_bounds__(256)
act_bwd_bf16_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, int act, int64_t n8, uint4* __restrict__ out) {
  griddep_wait();
  griddep_release();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const uint4 d = __ldcs(dy + i);
  const uint4 o = act != RB_ACT_NONE ? __ldcs(y + i) : make_uint4(0, 0, 0, 0);
  const uint32_t dv[4] = {d.x, d.y, d.z, d.w}, ov[4] = {o.x, o.y, o.z, o.w};
  uint32_t r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float g0 = __uint_as_float(dv[k] << 16), g1 = __uint_as_float(dv[k] & 0xFFFF0000u);
    if (act != RB_ACT_NONE) {
      const float y0 = __uint_as_float(ov[k] << 16), y1 = __uint_as_float(ov[k] & 0xFFFF0000u);
      g0 = act == RB_ACT_RELU ? (y0 > 0.f ? g0 : 0.f) : __fmul_rn(__fmul_rn(g0, y0), __fsub_rn(1.0f, y0));
      g1 = act == RB_ACT_RELU ? (y1 > 0.f ? g1 : 0.f) : __fmul_rn(__fmul_rn(g1, y1), __fsub_rn(1.0f, y1));
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(g0, g1);
    r[k] = *reinterpret_cast<uint32_t*>(&h);
  }
  out[i] = make_uint4(r[0], r[1], r[2], r[3]);
}

// x f32[rows, in_dim] -> bf16[rows, ld] = [x | 1 | 0 ...]: the K operand of the first Dense layer; the ones column makes
// row in_dim of that layer's weight-gradient GEMM its bias gradient
__global__ void __launch_bounds__(256)
pack_input_kernel(const float* __restrict__ x, int64_t rows, int in_dim, int64_t ldx, __nv_bfloat16* __restrict__ out, int ld, int ones_col) {
  griddep_wait();
  griddep_release();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= rows * ld) return;
  const int64_t r = i / ld;
  const int c = static_cast<int>(i % ld);
  float v = 0.f;
  if (c < in_dim) v = x[r * ldx + c];
  else if (c == in_dim && ones_col) v = 1.0f;
  out[i] = __float2bfloat16_rn(v);
}

// ---- host side: tensor maps ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D row-major matrix [outer, inner] with row stride `ld` elements; box = [box_outer, box_inner]; 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, bool f32, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_tiled();
  RB_CHECK_ARG(fn != nullptr, RB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const size_t esz = f32 ? 4 : 2;
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * esz) % 16 == 0, RB_ERR_ALIGN,
               "Dense operands need 16-byte aligned base pointers and row strides (ld * %zu bytes)", esz);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * esz};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RB_CHECK_ARG(r == CUDA_SUCCESS, RB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (inner %lld outer %lld ld %lld)", static_cast<int>(r),
               static_cast<long long>(inner), static_cast<long long>(outer), static_cast<long long>(ld));
  return RB_OK;
}

// MN-major operand [K, MN] (row stride ld) with MN a multiple of 64, seen as {64, K, MN / 64}: a box of `atoms` 64-column slabs
// of 64 reduction rows lands in shared memory slab after slab, exactly the layout the per-box loads produce
static int make_map_atoms(CUtensorMap* map, const void* base, int64_t mn, int64_t k, int64_t ld, int atoms) {
  EncodeTiledFn fn = encode_tiled();
  RB_CHECK_ARG(fn != nullptr, RB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, RB_ERR_ALIGN,
               "Dense operands need 16-byte aligned base pointers and row strides");
  cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(mn / 64)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2, 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(atoms)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RB_CHECK_ARG(r == CUDA_SUCCESS, RB_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with %d (mn %lld k %lld ld %lld)", static_cast<int>(r),
               static_cast<long long>(mn), static_cast<long long>(k), static_cast<long long>(ld));
  return RB_OK;
}

struct Operand {
  const void* p;
  int64_t rows, cols, ld;   // row-major [rows, cols]
  bool mn_major;            // true: `cols` is the M / N axis of the product and `rows` the reduction axis
};

static int block_n_for(int N) {
  int bn = (std::min(N, kMaxN) + 63) / 64 * 64;
  return bn;
}

static unsigned long long* g_debug_stats = nullptr;    // rb_dense_debug_stats

// RB_DENSE_PAIR=1 in the environment runs the products whose tiles allow it on cta_group::2 pairs.  Measured (r2_11, 65536 x 800
// x 512): forward 51.8 us against 52.9 us single-CTA, input gradient 62.0 against 59.5, weight gradient 62.5 against 59.5; the
// whole step 1.374 ms against 1.343 ms — the single-CTA form is the default.
static bool pair_enabled() {
  const char* e = getenv("RB_DENSE_PAIR");
  return e != nullptr && e[0] == '1';
}

// RB_DENSE_ATOMS=0: MN-major operands are staged one 64-column box per TMA instruction (A/B measurements)
static bool atoms_enabled() {
  const char* e = getenv("RB_DENSE_ATOMS");
  return e == nullptr || e[0] != '0';
}

// RB_DENSE_PREFETCH=1: a spare warp prefetches the streamed operand's boxes into L2 ahead of the TMA loads.  Off by default:
// measured (r2_12) it makes every product SLOWER (forward 53.2 -> 55.8 us, weight gradient 59.5 -> 67.6 us) and the MMA
// issuer waits on operands MORE, not less — the wait is not DRAM latency but the SM's L2 -> shared-memory ingest rate (a
// 128 x 256 tile needs 96 B / clk of operand fill), which extra TMA requests only load further.  What relieves it is the
// 2-CTA form (64 B / clk per SM): there the wait halves (stats r2_11), but its other costs cancel the gain.
static bool prefetch_enabled() {
  const char* e = getenv("RB_DENSE_PREFETCH");
  return e != nullptr && e[0] == '1';
}

static int max_pair_clusters() {
  static const int n = [] {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kNumSMs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int c = 0;
    if (cudaFuncSetAttribute(dense_gemm_pair_kernel<false, RB_ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&c, dense_gemm_pair_kernel<false, RB_ACT_NONE>, &cfg) != cudaSuccess)
      c = 0;
    (void)cudaGetLastError();
    return std::min(c, kNumSMs / 2);
  }();
  return n;
}

// How a product is cut into work items: tiles of (128 * cg) x block_n, the reduction split over `splits` workers
struct Plan {
  int cg;          // CTAs per tile: 2 = cta_group::2 pairs
  int block_n, m_blocks, n_blocks, k_blocks, splits, kbps, workers;
  int resident;    // B panel resident in shared memory (small K): grid is a multiple of n_blocks
};

// RB_DENSE_RESIDENT: 0 (default) = never, 1 = when every CTA gets at least two tiles, 2 = whenever the panel fits (tests).
// Off by default — measured (r2_18, B = 65536): no product gets faster and the ones whose column block has to shrink to fit
// the panel get much slower (800 -> 512 forward with 64-column blocks 53.9 -> 139 us, 512 -> 800 input gradient with 128-column
// blocks 59.8 -> 85.7 us, 512 -> 256 forward 25.3 -> 30.2 us): one thread issues one tcgen05.mma per ~130 cycles, so an MMA must
// be N = 256 wide (128 cycles of tensor work) to keep the pipe busy; halving N halves the work per issue slot.
static int resident_mode() {
  const char* e = getenv("RB_DENSE_RESIDENT");
  return e == nullptr ? 0 : atoi(e);
}

static Plan make_plan(int M, int N, int K, bool split_k, bool use_pairs) {
  Plan p;
  p.block_n = block_n_for(N);
  p.workers = kNumSMs;
  p.cg = 1;
  if (use_pairs && p.block_n % 128 == 0 && M >= 2 * kBlockM) {
    p.cg = 2;
    p.workers = kNumSMs / 2;
  }
  p.m_blocks = (M + kBlockM * p.cg - 1) / (kBlockM * p.cg);
  p.n_blocks = (N + p.block_n - 1) / p.block_n;
  p.k_blocks = (K + kBlockK - 1) / kBlockK;
  const int tiles = p.m_blocks * p.n_blocks;
  int s = 1;
  if (split_k && tiles < p.workers) s = std::max(1, std::min(p.workers / tiles, p.k_blocks));
  p.kbps = (p.k_blocks + s - 1) / s;
  p.splits = (p.k_blocks + p.kbps - 1) / p.kbps;
  // Small reduction axis (K <= 512: every product of the towers but the 800-wide ones): the operand B (the layer's kernel) of a
  // column block is at most 128 KiB and the same for every row tile — keep it in shared memory and stream A alone.  A 128 x 256
  // tile then needs 32 B / clk of operand fill instead of 96 (the SM ingests ~64): at K = 512 the column block shrinks to 128 so
  // that the panel fits.
  p.resident = 0;
  const int mode = resident_mode();
  if (mode > 0 && p.cg == 1 && !split_k && p.splits == 1) {
    const int k_pad = p.k_blocks * kBlockK;
    int bn = 0;
    for (int cand : {256, 192, 128, 64})
      if (bn == 0 && cand <= p.block_n && static_cast<int64_t>(k_pad) * cand * 2 <= kBResidentBytes) bn = cand;
    if (bn != 0) {
      const int nb = (N + bn - 1) / bn;
      if (nb <= kNumSMs && (mode >= 2 || static_cast<int64_t>(p.m_blocks) * nb >= 2 * kNumSMs)) {
        p.resident = 1;
        p.block_n = bn;
        p.n_blocks = nb;
      }
    }
  }
  return p;
}

// C[M, N] = A . B^T over K.  `c` is bf16 [M, N] / f32 [M, N] (splits == 1) or the f32 partial buffer.
static int launch_gemm(const Operand& a, const Operand& b, int M, int N, int K, void* c, bool c_f32, int64_t ldc, bool split_k, const float* bias,
                       int activation, void* ws, size_t ws_bytes, cudaStream_t st, int* splits_out, int64_t* split_stride_out) {
  Plan pl = make_plan(M, N, K, split_k, pair_enabled());
  if (pl.cg == 2 && max_pair_clusters() < 1) pl = make_plan(M, N, K, split_k, false);
  GemmArgs g;
  g.M = M;
  g.N = N;
  g.K = K;
  g.block_n = pl.block_n;
  g.a_mn = a.mn_major;
  g.b_mn = b.mn_major;
  g.m_blocks = pl.m_blocks;
  g.n_blocks = pl.n_blocks;
  g.k_blocks = pl.k_blocks;
  g.splits = pl.splits;
  g.k_blocks_per_split = pl.kbps;
  g.out_f32 = c_f32;
  g.activation = activation;
  g.b_resident = pl.resident;
  g.bias = bias;
  g.stats = g_debug_stats;
  const bool pf = prefetch_enabled();
  g.pf_a = pf && static_cast<int64_t>(M) * K * 2 > (8ll << 20);
  g.pf_b = pf && static_cast<int64_t>(N) * K * 2 > (8ll << 20);
  const int tile_m = kBlockM * pl.cg;
  CUtensorMap ma, mb, mc;
  int rc;
  // K-major operand [MN, K]: box = 64 reduction elements x (128 | this CTA's share of block_n) rows.
  // MN-major operand [K, MN]: box = 64 x 64.
  g.a_atoms = (a.mn_major && M % 64 == 0 && atoms_enabled()) ? 2 : 0;
  g.b_atoms = (b.mn_major && N % 64 == 0 && atoms_enabled()) ? g.block_n / 64 / pl.cg : 0;
  if (!a.mn_major) rc = make_map(&ma, a.p, false, K, M, a.ld, kBlockK, kBlockM);
  else if (g.a_atoms) rc = make_map_atoms(&ma, a.p, M, K, a.ld, g.a_atoms);
  else rc = make_map(&ma, a.p, false, M, K, a.ld, 64, kBlockK);
  if (rc != RB_OK) return rc;
  if (!b.mn_major) rc = make_map(&mb, b.p, false, K, N, b.ld, kBlockK, g.block_n / pl.cg);
  else if (g.b_atoms) rc = make_map_atoms(&mb, b.p, N, K, b.ld, g.b_atoms);
  else rc = make_map(&mb, b.p, false, N, K, b.ld, 64, kBlockK);
  if (rc != RB_OK) return rc;
  void* out = c;
  int64_t out_rows = M, out_ld = ldc;
  if (split_k) {
    const int64_t split_stride = static_cast<int64_t>(g.m_blocks) * tile_m * N;
    const size_t need = static_cast<size_t>(g.splits) * split_stride * sizeof(float);
    RB_CHECK_ARG(ws != nullptr && ws_bytes >= need, RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    out = ws;
    out_rows = static_cast<int64_t>(g.splits) * g.m_blocks * tile_m;
    out_ld = N;
    *splits_out = g.splits;
    *split_stride_out = split_stride;
  }
  rc = make_map(&mc, out, c_f32, N, out_rows, out_ld, 32, kBlockM, c_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != RB_OK) return rc;
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmArgs);
  static const KernelFn kernels[2][2][3] = {
      {{dense_gemm_kernel<false, RB_ACT_NONE>, dense_gemm_kernel<false, RB_ACT_RELU>, dense_gemm_kernel<false, RB_ACT_SIGMOID>},
       {dense_gemm_kernel<true, RB_ACT_NONE>, dense_gemm_kernel<true, RB_ACT_RELU>, dense_gemm_kernel<true, RB_ACT_SIGMOID>}},
      {{dense_gemm_pair_kernel<false, RB_ACT_NONE>, dense_gemm_pair_kernel<false, RB_ACT_RELU>, dense_gemm_pair_kernel<false, RB_ACT_SIGMOID>},
       {dense_gemm_pair_kernel<true, RB_ACT_NONE>, dense_gemm_pair_kernel<true, RB_ACT_RELU>, dense_gemm_pair_kernel<true, RB_ACT_SIGMOID>}}};
  static bool attr_set = false;
  if (!attr_set) {
    for (int q = 0; q < 2; ++q)
      for (int f = 0; f < 2; ++f)
        for (int t = 0; t < 3; ++t) RB_CUDA(cudaFuncSetAttribute(kernels[q][f][t], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int total = g.m_blocks * g.n_blocks * g.splits;
  const int workers = pl.cg == 2 ? std::min(pl.workers, max_pair_clusters()) : pl.workers;
  int grid = std::min(total, workers) * pl.cg;
  if (pl.resident) grid = std::min(workers / g.n_blocks, g.m_blocks) * g.n_blocks;      // a CTA keeps its column block
  RB_CUDA(launch_dependent(kernels[pl.cg - 1][c_f32 ? 1 : 0][activation], static_cast<unsigned int>(grid), kThreads, kSmemBytes, st, true, ma, mb,
                           mc, g));
  RB_LAUNCH_CHECK("dense_gemm_kernel");
  return RB_OK;
}

// the larger of the two plans' partial buffers: the workspace query does not know which form will run
static size_t weight_ws_bytes(int64_t rows, int in_dim, int units) {
  size_t need = 0;
  for (int pairs = 0; pairs < 2; ++pairs) {
    const Plan pl = make_plan(in_dim, units, static_cast<int>(std::min<int64_t>(rows, 0x7FFFFFFF)), true, pairs != 0);
    need = std::max(need, static_cast<size_t>(pl.splits) * pl.m_blocks * kBlockM * pl.cg * units * sizeof(float));
  }
  return need + 256;
}

static bool dims_ok(int64_t rows, int a, int b) { return rows > 0 && rows < (1ll << 31) && a > 0 && b > 0 && a % 8 == 0 && b % 8 == 0; }

}  // namespace mlp
}  // namespace rb

using namespace rb;
using namespace rb::mlp;

extern "C" int rb_dense_fwd(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* w, int32_t units, int64_t ldw,
                            const float* bias, int32_t activation, void* y, int32_t y_type, int64_t ldy, void* stream) {
  RB_CHECK_ARG(x != nullptr && w != nullptr && y != nullptr, RB_ERR_ARG, "x / w / y is null");
  RB_CHECK_ARG(activation >= RB_ACT_NONE && activation <= RB_ACT_SIGMOID, RB_ERR_ARG, "bad activation %d", activation);
  RB_CHECK_ARG(y_type == RB_F32 || y_type == RB_BF16, RB_ERR_ARG, "y_type must be RB_F32 or RB_BF16");
  RB_CHECK_ARG(dims_ok(rows, in_dim, units), RB_ERR_SHAPE, "Dense needs rows > 0 and in_dim, units multiples of 8 (got %lld, %d, %d)",
               static_cast<long long>(rows), in_dim, units);
  RB_CHECK_ARG(ldx >= in_dim && ldw >= units && ldy >= units, RB_ERR_ARG, "a leading dimension is smaller than its row");
  Operand a{x, rows, in_dim, ldx, false}, b{w, in_dim, units, ldw, true};
  return launch_gemm(a, b, static_cast<int>(rows), units, in_dim, y, y_type == RB_F32, ldy, false, bias, activation, nullptr, 0,
                     static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

extern "C" int rb_dense_bwd_input(const void* dy, int64_t rows, int32_t units, int64_t lddy, const void* w, int32_t in_dim, int64_t ldw,
                                  void* dx, int64_t lddx, void* stream) {
  RB_CHECK_ARG(dy != nullptr && w != nullptr && dx != nullptr, RB_ERR_ARG, "dy / w / dx is null");
  RB_CHECK_ARG(dims_ok(rows, in_dim, units), RB_ERR_SHAPE, "Dense needs rows > 0 and in_dim, units multiples of 8 (got %lld, %d, %d)",
               static_cast<long long>(rows), in_dim, units);
  RB_CHECK_ARG(lddy >= units && ldw >= units && lddx >= in_dim, RB_ERR_ARG, "a leading dimension is smaller than its row");
  // dx[r, i] = sum_u dy[r, u] * W[i, u]: both operands have the reduction index u contiguous
  Operand a{dy, rows, units, lddy, false}, b{w, in_dim, units, ldw, false};
  return launch_gemm(a, b, static_cast<int>(rows), in_dim, units, dx, false, lddx, false, nullptr, RB_ACT_NONE, nullptr, 0,
                     static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

extern "C" size_t rb_dense_bwd_weight_workspace_bytes(int64_t rows, int32_t in_dim, int32_t units) {
  if (!dims_ok(rows, in_dim, units)) return 0;
  return weight_ws_bytes(rows, in_dim, units);
}

extern "C" int rb_dense_bwd_weight(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* dy, int32_t units, int64_t lddy,
                                   float* dw, int64_t lddw, void* ws, size_t ws_bytes, void* stream) {
  RB_CHECK_ARG(x != nullptr && dy != nullptr && dw != nullptr, RB_ERR_ARG, "x / dy / dw is null");
  RB_CHECK_ARG(dims_ok(rows, in_dim, units), RB_ERR_SHAPE, "Dense needs rows > 0 and in_dim, units multiples of 8 (got %lld, %d, %d)",
               static_cast<long long>(rows), in_dim, units);
  RB_CHECK_ARG(ldx >= in_dim && lddy >= units && lddw >= units && lddw % 4 == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0, RB_ERR_ARG,
               "bad leading dimension / dw alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // dW[i, u] = sum_r x[r, i] * dy[r, u]: the reduction index r is the slow axis of both operands (MN-major)
  Operand a{x, rows, in_dim, ldx, true}, b{dy, rows, units, lddy, true};
  int splits = 1;
  int64_t split_stride = 0;
  int rc = launch_gemm(a, b, in_dim, units, static_cast<int>(rows), nullptr, true, 0, true, nullptr, RB_ACT_NONE, ws, ws_bytes, st, &splits,
                       &split_stride);
  if (rc != RB_OK) return rc;
  const int64_t work = static_cast<int64_t>(in_dim) * (units / 4);
  RB_CUDA(launch_dependent(reduce_splits_kernel, grid_for(work, 256), 256, 0, st, true, static_cast<const float*>(ws), splits, split_stride, in_dim, units, dw, lddw));
  RB_LAUNCH_CHECK("reduce_splits_kernel");
  return RB_OK;
}

extern "C" int rb_dense_head_fwd(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* w, const float* bias, int32_t activation,
                                 float* out, void* stream) {
  RB_CHECK_ARG(x != nullptr && w != nullptr && out != nullptr, RB_ERR_ARG, "x / w / out is null");
  RB_CHECK_ARG(rows > 0 && in_dim > 0 && in_dim % 8 == 0 && ldx % 8 == 0 && ldx >= in_dim, RB_ERR_SHAPE, "head: in_dim and ldx must be multiples of 8");
  RB_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0, RB_ERR_ALIGN, "head: x / w not 16-byte aligned");
  RB_CHECK_ARG(activation >= RB_ACT_NONE && activation <= RB_ACT_SIGMOID, RB_ERR_ARG, "bad activation %d", activation);
  if (in_dim <= 64)
    RB_CUDA(launch_dependent(head_fwd_narrow_kernel, grid_for(rows, 256), 256, 0, static_cast<cudaStream_t>(stream), true, static_cast<const __nv_bfloat16*>(x), rows, in_dim, ldx, static_cast<const __nv_bfloat16*>(w), bias, activation, out));
  else
    RB_CUDA(launch_dependent(head_fwd_kernel, grid_for(rows, 8), 256, 0, static_cast<cudaStream_t>(stream), true, static_cast<const __nv_bfloat16*>(x), rows, in_dim, ldx, static_cast<const __nv_bfloat16*>(w), bias, activation, out));
  RB_LAUNCH_CHECK("head_fwd_kernel");
  return RB_OK;
}

static int head_parts(int64_t rows) { return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(4 * kNumSMs, rows / 64))); }

extern "C" size_t rb_dense_head_bwd_workspace_bytes(int64_t rows, int32_t in_dim) {
  if (rows <= 0 || in_dim <= 0) return 0;
  return static_cast<size_t>(head_parts(rows)) * (2 * in_dim + 1) * sizeof(float) + 256;
}

extern "C" int rb_dense_head_bwd(const float* dout, const float* out, int32_t activation, const void* x, int64_t rows, int32_t in_dim, int64_t ldx,
                                 const void* w, void* dx, int64_t lddx, float* dw, float* db, float* dx_colsum, void* ws, size_t ws_bytes,
                                 void* stream) {
  RB_CHECK_ARG(dout != nullptr && x != nullptr && w != nullptr && dw != nullptr && db != nullptr, RB_ERR_ARG, "a required pointer is null");
  RB_CHECK_ARG(activation == RB_ACT_NONE || out != nullptr, RB_ERR_ARG, "the activation's backward needs the forward output");
  RB_CHECK_ARG(rows > 0 && in_dim > 0 && in_dim % 8 == 0 && in_dim <= 8 * kHeadBwdThreads && ldx % 8 == 0 && ldx >= in_dim &&
                   (dx == nullptr || (lddx % 8 == 0 && lddx >= in_dim)),
               RB_ERR_SHAPE, "head: in_dim (<= %d) and the leading dimensions must be multiples of 8", 8 * kHeadBwdThreads);
  RB_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, RB_ERR_ALIGN,
               "head: x / w / dx not 16-byte aligned");
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= rb_dense_head_bwd_workspace_bytes(rows, in_dim), RB_ERR_WORKSPACE,
               "workspace too small: need %zu bytes, got %zu", rb_dense_head_bwd_workspace_bytes(rows, in_dim), ws_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int parts = head_parts(rows);
  RB_CUDA(launch_dependent(head_bwd_kernel, parts, kHeadBwdThreads, 0, st, true, dout, out, activation, static_cast<const __nv_bfloat16*>(x), rows, in_dim, ldx, static_cast<const __nv_bfloat16*>(w), static_cast<__nv_bfloat16*>(dx), lddx, static_cast<float*>(ws)));
  RB_LAUNCH_CHECK("head_bwd_kernel");
  RB_CUDA(launch_dependent(head_bwd_final_kernel, (2 * in_dim + 1 + 7) / 8, 256, 0, st, true, static_cast<const float*>(ws), parts, in_dim, dw, db, dx_colsum));
  RB_LAUNCH_CHECK("head_bwd_final_kernel");
  return RB_OK;
}

extern "C" size_t rb_dense_head_bce_workspace_bytes(int64_t rows, int32_t in_dim) {
  if (rows <= 0 || in_dim <= 0) return 0;
  return static_cast<size_t>(head_parts(rows)) * (2 * in_dim + 2) * sizeof(float) + 256;
}

extern "C" int rb_dense_head_bce(const void* x, int64_t rows, int32_t in_dim, int64_t ldx, const void* w, const float* bias, const void* label,
                                 int32_t label_type, float* prob, float* loss, void* dx, int64_t lddx, float* dw, float* db, float* dx_colsum,
                                 void* ws, size_t ws_bytes, void* stream) {
  RB_CHECK_ARG(x != nullptr && w != nullptr && label != nullptr && prob != nullptr && loss != nullptr && dw != nullptr && db != nullptr, RB_ERR_ARG,
               "a required pointer is null");
  RB_CHECK_ARG(label_type == 0 || label_type == 1, RB_ERR_ARG, "label_type: 0 = f32, 1 = i64");
  const int vcols = in_dim / 8;
  RB_CHECK_ARG(rows > 0 && in_dim >= 8 && in_dim % 8 == 0 && vcols <= 32 && (vcols & (vcols - 1)) == 0 && ldx % 8 == 0 && ldx >= in_dim &&
                   (dx == nullptr || (lddx % 8 == 0 && lddx >= in_dim)),
               RB_ERR_SHAPE, "fused head + loss: in_dim in {8, 16, 32, 64, 128, 256}, leading dimensions multiples of 8");
  RB_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, RB_ERR_ALIGN,
               "head: x / w / dx not 16-byte aligned");
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= rb_dense_head_bce_workspace_bytes(rows, in_dim), RB_ERR_WORKSPACE,
               "workspace too small: need %zu bytes, got %zu", rb_dense_head_bce_workspace_bytes(rows, in_dim), ws_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int parts = head_parts(rows);
  const float inv_n = 1.0f / static_cast<float>(rows);
  RB_CUDA(launch_dependent(head_bce_kernel, parts, kHeadBwdThreads, 0, st, true, static_cast<const __nv_bfloat16*>(x), rows, in_dim, ldx, static_cast<const __nv_bfloat16*>(w), bias, label, label_type, inv_n, prob, static_cast<__nv_bfloat16*>(dx), lddx, static_cast<float*>(ws)));
  RB_LAUNCH_CHECK("head_bce_kernel");
  RB_CUDA(launch_dependent(head_bce_final_kernel, (2 * in_dim + 2 + 7) / 8, 256, 0, st, true, static_cast<const float*>(ws), parts, in_dim, inv_n, dw, db, dx_colsum, loss));
  RB_LAUNCH_CHECK("head_bce_final_kernel");
  return RB_OK;
}

extern "C" int rb_dense_act_bwd(const float* dy, const float* y, int32_t activation, int64_t n, void* out_bf16, void* stream) {
  RB_CHECK_ARG(dy != nullptr && out_bf16 != nullptr && (activation == RB_ACT_NONE || y != nullptr), RB_ERR_ARG, "a required pointer is null");
  RB_CHECK_ARG(n > 0 && n % 4 == 0, RB_ERR_SHAPE, "act_bwd: element count must be a positive multiple of 4");
  RB_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0,
               RB_ERR_ALIGN, "act_bwd: pointers not aligned");
  RB_CHECK_ARG(activation >= RB_ACT_NONE && activation <= RB_ACT_SIGMOID, RB_ERR_ARG, "bad activation %d", activation);
  RB_CUDA(launch_dependent(act_bwd_kernel, grid_for(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream), true, dy, y, activation, n / 4, static_cast<__nv_bfloat16*>(out_bf16)));
  RB_LAUNCH_CHECK("act_bwd_kernel");
  return RB_OK;
}

extern "C" int rb_dense_act_bwd_bf16(const void* dy_bf16, const void* y_bf16, int32_t activation, int64_t n, void* out_bf16, void* stream) {
  RB_CHECK_ARG(dy_bf16 != nullptr && out_bf16 != nullptr && (activation == RB_ACT_NONE || y_bf16 != nullptr), RB_ERR_ARG, "a required pointer is null");
  RB_CHECK_ARG(n > 0 && n % 8 == 0, RB_ERR_SHAPE, "act_bwd_bf16: element count must be a positive multiple of 8");
  RB_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy_bf16) | reinterpret_cast<uintptr_t>(y_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0,
               RB_ERR_ALIGN, "act_bwd_bf16: pointers not 16-byte aligned");
  RB_CHECK_ARG(activation >= RB_ACT_NONE && activation <= RB_ACT_SIGMOID, RB_ERR_ARG, "bad activation %d", activation);
  RB_CUDA(launch_dependent(act_bwd_bf16_kernel, grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream), true, static_cast<const uint4*>(dy_bf16), static_cast<const uint4*>(y_bf16), activation, n / 8, static_cast<uint4*>(out_bf16)));
  RB_LAUNCH_CHECK("act_bwd_bf16_kernel");
  return RB_OK;
}

extern "C" int rb_dense_pack_input(const float* x, int64_t rows, int32_t in_dim, int64_t ldx, void* out_bf16, int32_t ld_out, int32_t ones_col,
                                   void* stream) {
  RB_CHECK_ARG(x != nullptr && out_bf16 != nullptr, RB_ERR_ARG, "x / out is null");
  RB_CHECK_ARG(rows > 0 && in_dim > 0 && ld_out >= in_dim + (ones_col ? 1 : 0) && ldx >= in_dim, RB_ERR_SHAPE, "pack_input: bad sizes");
  RB_CUDA(launch_dependent(pack_input_kernel, grid_for(rows * ld_out, 256), 256, 0, static_cast<cudaStream_t>(stream), true, x, rows, in_dim, ldx, static_cast<__nv_bfloat16*>(out_bf16), ld_out, ones_col));
  RB_LAUNCH_CHECK("pack_input_kernel");
  return RB_OK;
}

extern "C" int rb_dense_debug_stats(void* device_buf) {
  g_debug_stats = static_cast<unsigned long long*>(device_buf);
  return RB_OK;
}
