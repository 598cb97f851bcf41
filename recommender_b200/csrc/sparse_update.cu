// K7..K9: backward scatter of the lookup + fused sparse optimizer row update (sm_100a).
//
// Deterministic and free of floating-point atomics:
//   1. keys[p] = table row of lookup position p, vals[p] = p;  stable LSD radix sort by row
//      (only ceil(log2(rows)) key bits are sorted).  Within a row, positions stay ascending,
//      i.e. the order TF's CPU UnsortedSegmentSum adds them.
//   2. seg_reduce_tiles_kernel: every GS-lane group walks a tile of 32 consecutive sorted
//      entries, builds each entry's gradient row on the fly (multi-consumer sum, mean /
//      masked-mean scaling of a bag-level gradient, FM term) and sums runs of equal rows in
//      order.  A run that lies inside the tile is handed straight to the sink (optimizer row
//      update, or the compact IndexedSlices writer).  Runs that cross tile borders leave
//      per-tile partial sums.
//   3. seg_chain_kernel: the tile where a crossing run starts adds the partials of the
//      following tiles in order and calls the sink; chains longer than kLongChain tiles (hot
//      rows: OOV id 0, 3-row ESMM tables) are queued and reduced by a whole CTA each in
//      seg_long_chain_kernel with a fixed split, so the result is run-to-run identical.
//      Both are launched programmatically (griddepcontrol) while the tile kernel drains: which tiles own a crossing run is
//      decided from the sorted keys alone, and only the owners wait for the tile kernel's partial sums (r2_65: 511 -> 507 us).
// HBM traffic is the algorithmic minimum: each gradient row once, each touched table/state
// row read once and written once; sort traffic is 16 B per lookup per pass.
#include <cuda_bf16.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace rb {

constexpr int kTile = 32;         // sorted entries per group
constexpr int kSegThreads = 256;  // CTA size of the reduction kernels
constexpr int kChunk = 4;         // gradient rows in flight per group
constexpr int kLongChain = 64;    // tiles; longer chains go to the CTA-wide kernel
#ifndef RB_SEG_RING
#define RB_SEG_RING 3
#endif
constexpr int kRing = RB_SEG_RING;  // sorted entries in flight per lane (cp.async ring in shared memory)

struct GradSrcDev {
  int num_src, scale_mode, L, is64;
  uint32_t L_recip;   // floor(2^32 / L): p / L = umulhi(p, L_recip) (+1 after one correction)
  const float* src[RB_MAX_GRAD_SOURCES];
  int64_t bag_stride[RB_MAX_GRAD_SOURCES];
  int64_t pos_stride[RB_MAX_GRAD_SOURCES];
  const void* mask_idx;
  const float* count;
  const float* fm_g;
  const float* fm_s;
};

// All uses of one table inside a step (their IndexedSlices are concatenated in use
// order).  Global lookup position p belongs to group k when start[k] <= p < start[k+1].
constexpr int kMaxGroups = RB_MAX_LOOKUP_GROUPS > RB_MAX_RANKS ? RB_MAX_LOOKUP_GROUPS : RB_MAX_RANKS;

struct GradGroupsDev {
  int num;
  int D;
  int peer;                                  // gradient rows may live in peer memory (sharded path): copy them through L1
  uint32_t start[kMaxGroups + 1];            // unused entries = 0xFFFFFFFF
  const float* table;                        // read-only view of the table for the FM term
  GradSrcDev g[kMaxGroups];
};

struct LongChain {
  uint32_t row, first_tile, last_tile, seg_first;
};

// ---- gradient row of one lookup position ----------------------------------------------------------
// Where position p's gradient lives: which use of the table (group), which bag, which slot of the bag.
struct GradPos {
  int gi;
  uint32_t p, b, l;   // p: position inside the group
};

__device__ __forceinline__ GradPos decode_pos(const GradGroupsDev& gg, uint32_t p) {
  GradPos q;
  q.gi = 0;
#pragma unroll
  for (int k = 1; k < kMaxGroups; ++k) q.gi += (p >= gg.start[k]) ? 1 : 0;
  const uint32_t L = static_cast<uint32_t>(gg.g[q.gi].L);
  q.p = p - gg.start[q.gi];
  // q.p / L without the ~20-instruction runtime division: the reciprocal estimate is low by at most one
  q.b = __umulhi(q.p, gg.g[q.gi].L_recip);
  q.l = q.p - q.b * L;
  if (q.l >= L) {
    ++q.b;
    q.l -= L;
  }
  return q;
}

// SIMPLE gradients (one use of the table, one source tensor, no scaling, no FM term — the DLRM / sharded case):
// the second half of the work on a gradient row vanishes and the group search is skipped
__device__ __forceinline__ GradPos decode_pos_simple(const GradGroupsDev& gg, uint32_t p) {
  GradPos q;
  q.gi = 0;
  const uint32_t L = static_cast<uint32_t>(gg.g[0].L);
  q.p = p;
  q.b = __umulhi(p, gg.g[0].L_recip);
  q.l = p - q.b * L;
  if (q.l >= L) {
    ++q.b;
    q.l -= L;
  }
  return q;
}

// address of the first source's row for this lane's columns (the part fetched by cp.async)
__device__ __forceinline__ const float* grad_src0(const GradGroupsDev& gg, const GradPos& q, int c) {
  const GradSrcDev& g = gg.g[q.gi];
  return g.src[0] + q.b * g.bag_stride[0] + q.l * g.pos_stride[0] + c;
}

// r = the first source's row (already loaded); adds the other consumers, the mean / masked-mean scaling
// and the FM term (w = this lane's slice of W[row,:], read before any update of that row)
template <int VEC>
__device__ __forceinline__ void finish_grad(const GradGroupsDev& gg, const GradPos& q, int c, Row<VEC>& r, const float* w) {
  const GradSrcDev& g = gg.g[q.gi];
  for (int k = 1; k < g.num_src; ++k) {  // consumers are added left to right (esmm/esmm.py:23-24)
    Row<VEC> t = ld_row_stream<VEC>(g.src[k] + q.b * g.bag_stride[k] + q.l * g.pos_stride[k] + c);
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __fadd_rn(r.v[i], t.v[i]);
  }
  if (g.scale_mode == RB_SCALE_MEAN) {
    const float denom = static_cast<float>(g.L);
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __fdiv_rn(r.v[i], denom);
  } else if (g.scale_mode == RB_SCALE_MASKED_MEAN) {
    const bool keep = load_raw_index(g.mask_idx, g.is64, q.p) != 0;
    const float denom = __ldg(g.count + q.b);
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = keep ? __fdiv_rn(r.v[i], denom) : 0.f;
  } else if (g.scale_mode == RB_SCALE_MASKED) {   // rows of masked positions were never written: whatever was read is dropped
    const bool keep = load_raw_index(g.mask_idx, g.is64, q.p) != 0;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = keep ? r.v[i] : 0.f;
  }
  if (g.fm_g != nullptr) {  // dE += g_fm[b] * (s[b,:] - W[row,:])      (ctr/model.py:21-23 backward)
    const float gb = __ldg(g.fm_g + q.b);
    Row<VEC> sv = ld_row<VEC>(g.fm_s + static_cast<int64_t>(q.b) * gg.D + c);
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = __fadd_rn(r.v[i], __fmul_rn(gb, __fsub_rn(sv.v[i], w[i])));
  }
}

// ---- sinks ----------------------------------------------------------------------------------------
struct OptSink {  // fused optimizer row update; every op explicitly rounded (no FMA) to match numpy
  float* table;
  float* s0;
  float* s1;
  int D;
  int opt;  // rb_optimizer, RB_OPT_ADAM_TF_DENSE = scatter-add phase only
  float lr, b1, b2, omb1, omb2, eps, alpha;
  const float* alpha_dev;   // optional: alpha_t read from device memory (CUDA-graph replays change it per step)
  __nv_bfloat16* shadow;    // optional bf16 copy of the table kept in step with it (what sharded forwards read over NVLink)

  template <int VEC>
  __device__ __forceinline__ void store_shadow(int64_t o, const Row<VEC>& w) const {
    if (shadow == nullptr) return;
    if constexpr (VEC == 4) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(w.v[0], w.v[1]);
      __nv_bfloat162 hi = __floats2bfloat162_rn(w.v[2], w.v[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(shadow + o) = pk;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) shadow[o + i] = __float2bfloat16_rn(w.v[i]);
    }
  }

  __device__ __forceinline__ void prepare() {
    if (alpha_dev != nullptr) alpha = __ldg(alpha_dev);
  }

  static constexpr int kStateRows = 3;  // ring slots per entry after the gradient: W, state0, state1

  // Start the async copies of this lane's slices of W / state rows into its ring slots
  // (slot k at st + k*kstride).  update: the entry closes a run this tile owns; need_w: FM term.
  template <int VEC>
  __device__ __forceinline__ void issue_state(uint32_t row, int c, bool update, bool need_w, float* st, int kstride) const {
    const int64_t o = static_cast<int64_t>(row) * D + c;
    if ((update && opt != RB_OPT_ADAM_TF_DENSE) || need_w) cp_async_vec<VEC>(st, table + o);
    if (update && opt != RB_OPT_SGD) cp_async_vec<VEC>(st + kstride, s0 + o);
    if (update && (opt == RB_OPT_ADAM_LAZY || opt == RB_OPT_ADAM_TF_DENSE)) cp_async_vec<VEC>(st + 2 * kstride, s1 + o);
  }

  // the row update with W / state slices already in shared memory (same arithmetic as apply())
  template <int VEC>
  __device__ __forceinline__ void apply_staged(uint32_t row, const Row<VEC>& g, int lane, uint32_t /*seg_first*/, const float* st,
                                               int kstride) const {
    const int64_t o = static_cast<int64_t>(row) * D + lane * VEC;
    Row<VEC> w, m, v;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      w.v[i] = st[i];
      m.v[i] = st[kstride + i];
      v.v[i] = st[2 * kstride + i];
    }
    update_row<VEC>(o, g, w, m, v);
  }

  __device__ __forceinline__ OptMath math() const { return OptMath{opt, lr, b1, b2, omb1, omb2, eps, alpha}; }

  template <int VEC>
  __device__ __forceinline__ void update_row(int64_t o, const Row<VEC>& g, Row<VEC>& w, Row<VEC>& m, Row<VEC>& v) const {
    opt_row_math<VEC>(math(), g, w, m, v);
    if (opt == RB_OPT_ADAM_LAZY) {
      st_row<VEC>(s0 + o, m);
      st_row<VEC>(s1 + o, v);
      st_row<VEC>(table + o, w);
      store_shadow<VEC>(o, w);
    } else if (opt == RB_OPT_ADAM_TF_DENSE) {
      st_row<VEC>(s0 + o, m);
      st_row<VEC>(s1 + o, v);
    } else if (opt == RB_OPT_ADAGRAD) {   // accumulator in slot 1 (m)
      st_row<VEC>(s0 + o, m);
      st_row<VEC>(table + o, w);
      store_shadow<VEC>(o, w);
    } else {  // SGD
      st_row<VEC>(table + o, w);
      store_shadow<VEC>(o, w);
    }
  }

  // the row update straight from global memory (chain kernels)
  template <int VEC>
  __device__ __forceinline__ void apply(uint32_t row, const Row<VEC>& g, int lane, uint32_t /*seg_first*/) const {
    const int64_t o = static_cast<int64_t>(row) * D + lane * VEC;
    Row<VEC> w = zero_row<VEC>(), m = zero_row<VEC>(), v = zero_row<VEC>();
    if (opt != RB_OPT_ADAM_TF_DENSE) w = ld_row_rw<VEC>(table + o);
    if (opt != RB_OPT_SGD) m = ld_row_rw<VEC>(s0 + o);
    if (opt == RB_OPT_ADAM_LAZY || opt == RB_OPT_ADAM_TF_DENSE) v = ld_row_rw<VEC>(s1 + o);
    update_row<VEC>(o, g, w, m, v);
  }
};

struct DedupSink {  // writes the deduplicated IndexedSlices (rows ascending)
  const int32_t* seg_incl;  // inclusive scan of the run-head flags over the sorted entries
  int64_t* uniq_rows;
  float* uniq_grad;
  int D;

  static constexpr int kStateRows = 0;
  __device__ __forceinline__ void prepare() {}

  template <int VEC>
  __device__ __forceinline__ void issue_state(uint32_t, int, bool, bool, float*, int) const {}

  template <int VEC>
  __device__ __forceinline__ void apply_staged(uint32_t row, const Row<VEC>& g, int lane, uint32_t seg_first, const float*, int) const {
    apply<VEC>(row, g, lane, seg_first);
  }

  template <int VEC>
  __device__ __forceinline__ void apply(uint32_t row, const Row<VEC>& g, int lane, uint32_t seg_first) const {
    const int64_t u = static_cast<int64_t>(seg_incl[seg_first]) - 1;
    if (lane == 0) uniq_rows[u] = static_cast<int64_t>(row);
    st_row<VEC>(uniq_grad + u * D + lane * VEC, g);
  }
};

// ---- step 1: keys ------------------------------------------------------------------------------------
// drop_mask != null (masked-mean pooling, dien/layers.py:13): a masked position carries an exactly-zero gradient row.
// TF still lists it in the IndexedSlices, so its row counts as touched — but one zero pair per row says that as well as
// fifty.  A masked position whose predecessor in the same bag is masked too and maps to the same row gets `invalid_key`,
// sorts behind every real pair and is cut off by the device-side pair count; the first pad of each run stays.  The sums
// are unchanged bit for bit (x + 0 == x), every optimizer sees the same touched rows, and the trailing pads of a
// behaviour history (dien/data_loader.py:44) no longer form one run of B*L/2 pairs on row 0.
__global__ void make_keys_kernel(IndexMap m, int64_t n, int64_t start, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals, int* __restrict__ oob_flag, const void* __restrict__ drop_mask,
                                 uint32_t invalid_key) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int64_t row = map_index(m, p);
  if (row < 0) {  // the forward already produced zeros for it; keep the pair harmless and flag it
    if (oob_flag != nullptr) *oob_flag = 1;
    row = 0;
  }
  uint32_t key = static_cast<uint32_t>(row);
  if (drop_mask != nullptr && (p % m.L) != 0 && load_raw_index(drop_mask, m.is64, p) == 0 &&
      load_raw_index(drop_mask, m.is64, p - 1) == 0) {
    int64_t prev = map_index(m, p - 1);
    if (prev < 0) prev = 0;
    if (prev == row) key = invalid_key;
  }
  keys[start + p] = key;
  vals[start + p] = static_cast<uint32_t>(start + p);  // global position over the concatenated groups
}

// number of real pairs = first sorted position holding invalid_key
__global__ void count_valid_kernel(const uint32_t* __restrict__ keys, int n, uint32_t invalid_key, int* __restrict__ n_valid) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (keys[mid] < invalid_key) lo = mid + 1;
    else hi = mid;
  }
  *n_valid = lo;
}

__global__ void head_flags_kernel(const uint32_t* __restrict__ keys, int n, int32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void write_num_unique_kernel(const int32_t* __restrict__ seg_incl, int n, int64_t* __restrict__ out) {
  out[0] = (n > 0) ? static_cast<int64_t>(seg_incl[n - 1]) : 0;
}

// ---- programmatic dependent launch (griddepcontrol): the border kernels of step 3 are launched while the tile kernel drains ----
// The tile kernel releases its dependents as soon as every one of its CTAs has started; seg_chain_kernel's CTAs then take the SM
// slots the tile kernel's last wave frees, decide from the sorted keys alone (final before the tile kernel started) whether
// their tile owns a border-crossing run — all but a few hundred of 53 k tiles at config 2 do not, and exit — and only the owners
// wait for the tile kernel's completion before they read its partial sums.  Thread 0 of the grid always waits, so that the
// completion of each kernel of the chain still implies the completion of the one before it.

// ---- step 2: tiles -------------------------------------------------------------------------------------
// Every lane runs a private kRing-deep pipeline over ITS columns of the tile's entries: the gradient
// row slice (and, for an entry that closes a run owned by this tile, the W / state row slices) are
// fetched with cp.async into the lane's ring slots, so up to kRing entries x 4 rows are in flight per
// lane without holding registers; a lane only ever reads back the bytes it copied itself, so no
// cross-lane synchronisation is needed.  dynamic smem: ring[kRing][1 + Sink::kStateRows][kSegThreads][VEC].
template <int VEC, int GS, class Sink, bool SIMPLE>
__global__ void __launch_bounds__(kSegThreads)
seg_reduce_tiles_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n, const int* __restrict__ n_dev,
                        const __grid_constant__ GradGroupsDev gsrc, Sink sink, float* __restrict__ head_part,
                        float* __restrict__ tail_part, int* __restrict__ long_count) {
  griddep_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) *long_count = 0;   // seg_chain_kernel appends to the list only after this grid is complete
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));   // padded capacity (sharded path): only the first *n_dev pairs are real
  sink.prepare();
  constexpr int kGroups = kSegThreads / GS;
  constexpr int kEntries = kGroups * kTile;
  constexpr int kKinds = 1 + Sink::kStateRows;
  constexpr int kKStride = kSegThreads * VEC;        // floats between the kinds of one ring slot
  constexpr int kSlotStride = kKinds * kKStride;     // floats between ring slots
  __shared__ uint32_t s_key[kEntries + 2];  // [0] = key before the CTA's range, [kEntries+1] = key after
  __shared__ uint32_t s_pos[kEntries];
  extern __shared__ __align__(16) float ring[];

  const int cta_base = blockIdx.x * kEntries;
  const int cta_cnt = min(kEntries, n - cta_base);
  if (cta_cnt <= 0) return;
  for (int i = threadIdx.x; i < cta_cnt; i += kSegThreads) {
    s_key[i + 1] = keys[cta_base + i];
    s_pos[i] = vals[cta_base + i];
  }
  if (threadIdx.x == 0) {
    // sentinels differ from every real key on the respective side only when no neighbour exists;
    // has_prev / has_next below guard their use
    s_key[0] = (cta_base > 0) ? keys[cta_base - 1] : 0u;
    s_key[cta_cnt + 1] = (cta_base + cta_cnt < n) ? keys[cta_base + cta_cnt] : 0u;
  }
  __syncthreads();

  const int group = threadIdx.x / GS;
  const int lane = threadIdx.x % GS;
  const int tbase = group * kTile;                 // offset inside the CTA range
  const int tcnt = min(kTile, cta_cnt - tbase);    // entries of this tile (<= 0: idle group)
  if (tcnt <= 0) return;
  const bool active = lane * VEC < gsrc.D;
  const int c = lane * VEC;
  const int gbase = cta_base + tbase;              // global sorted index of the tile's first entry
  const int tile_id = gbase / kTile;
  const bool has_prev = gbase > 0;
  const bool has_next = gbase + tcnt < n;
  const uint32_t* tk = s_key + 1 + tbase;          // tk[-1] and tk[tcnt] are the neighbours
  const uint32_t* tp = s_pos + tbase;
  const uint32_t first_key = tk[0];
  const bool cont_first = has_prev && (tk[-1] == first_key);   // the leading run started in an earlier tile
  float* my_ring = ring + threadIdx.x * VEC;

  auto run_ends_at = [&](int j, uint32_t key) {
    return (j == tcnt - 1) ? !(has_next && tk[tcnt] == key) : (tk[j + 1] != key);
  };
  auto issue = [&](int j, int slot) {
    if (j < tcnt && active) {
      const uint32_t key = tk[j];
      const GradPos q = SIMPLE ? decode_pos_simple(gsrc, tp[j]) : decode_pos(gsrc, tp[j]);
      float* dst = my_ring + slot * kSlotStride;
      cp_async_vec<VEC>(dst, grad_src0(gsrc, q, c), gsrc.peer != 0);
      if (kKinds > 1) {
        const bool update = run_ends_at(j, key) && !(cont_first && key == first_key);
        sink.template issue_state<VEC>(key, c, update, !SIMPLE && gsrc.g[q.gi].fm_g != nullptr, dst + kKStride, kKStride);
      }
    }
    cp_async_commit();   // one group per entry, also when nothing was copied: keeps wait_group uniform
  };

  Row<VEC> acc = zero_row<VEC>();
  uint32_t cur = first_key;
  int seg_start = 0;
  bool continues_prev = cont_first;
#pragma unroll
  for (int j = 0; j < kRing - 1; ++j) issue(j, j);

  int slot = 0;
  for (int j = 0; j < tcnt; ++j) {
    {
      int s_issue = slot + (kRing - 1);
      if (s_issue >= kRing) s_issue -= kRing;
      issue(j + kRing - 1, s_issue);
    }
    cp_async_wait<kRing - 1>();
    const uint32_t key = tk[j];
    const float* sl = my_ring + slot * kSlotStride;
    Row<VEC> g = zero_row<VEC>();
    if (active) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) g.v[i] = sl[i];
      if constexpr (!SIMPLE) finish_grad<VEC>(gsrc, decode_pos(gsrc, tp[j]), c, g, sl + kKStride);
    }
    if (key != cur) {  // previous run was closed below; start a new one
      cur = key;
      acc = zero_row<VEC>();
      seg_start = j;
      continues_prev = false;
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc.v[i] = __fadd_rn(acc.v[i], g.v[i]);
    const bool last_in_tile = (j == tcnt - 1);
    if (run_ends_at(j, cur)) {
      if (continues_prev) {
        if (active) st_row<VEC>(head_part + static_cast<int64_t>(tile_id) * gsrc.D + c, acc);
      } else if (active) {
        sink.template apply_staged<VEC>(cur, acc, lane, static_cast<uint32_t>(gbase + seg_start), sl + kKStride, kKStride);
      }
    } else if (last_in_tile && active) {  // run goes on in the next tile
      float* dst = continues_prev ? head_part : tail_part;
      st_row<VEC>(dst + static_cast<int64_t>(tile_id) * gsrc.D + c, acc);
    }
    if (++slot == kRing) slot = 0;
  }
}

// ---- step 2, row-granular form: one bulk async copy (TMA 1-D, SASS UBLKCP) per row ---------------------------------------
// Same tiles, same order of additions, same sink arithmetic as seg_reduce_tiles_kernel — what changes is who moves the bytes.
// There, every lane issues one 16-byte cp.async per row kind and entry (4 LDGSTS per entry and lane): at config 2 the
// LSU / MIO queues saturate before HBM does (ncu r1_15: mio_throttle + short_scoreboard = 41 % of the stalls at 73 % DRAM
// utilisation).  Here lane 0 of a group issues ONE cp.async.bulk per row — the gradient row, and for an entry that closes
// a run this tile owns the W / m / v rows — into the group's ring slot, completion counted in bytes by the slot's mbarrier;
// the lanes then read their 16-byte slices.  Applies to plain gradients (one source tensor per use of the table, no
// scaling, no FM term: DLRM, and the per-rank dE buffers of the sharded path) with rows that are multiples of 16 bytes.
#ifndef RB_SEG_EVICT_FIRST
#define RB_SEG_EVICT_FIRST 0     // 1: gradient rows are bulk-copied with an L2 evict-first hint.  Measured slower (r2_43: 519 vs 515 us uniform, 408 vs 394 us Zipf)
#endif
#ifndef RB_BULK_RING
#define RB_BULK_RING 4
#endif
constexpr int kBulkRing = RB_BULK_RING;       // entries in flight per group

__device__ __forceinline__ uint32_t su_saddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void su_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su_saddr(bar)), "r"(count));
}
__device__ __forceinline__ void su_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void su_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su_saddr(dst)), "l"(src),
               "r"(bytes), "r"(su_saddr(bar))
               : "memory");
}
// the same with an L2 eviction-priority hint: gradient rows are read exactly once
__device__ __forceinline__ void su_bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(su_saddr(dst)),
               "l"(src), "r"(bytes), "r"(su_saddr(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t su_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void su_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(su_saddr(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();     // a byte count that never completes must not hang the device
  } while (!ok);
}

// FLAT: one use of the table whose gradient rows lie one after the other in position order (dE[B, L, D] of the un-pooled lookup:
// bag_stride == L * pos_stride) — row p sits at src + p * pos_stride, no group search and no division by L (SASS r2_33: the
// general addressing was ~40 of the 241 warp instructions per entry of a kernel that issues 71 % of its cycles)
template <int VEC, int GS, bool FLAT>
__global__ void __launch_bounds__(kSegThreads)
seg_reduce_tiles_bulk_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n, const int* __restrict__ n_dev,
                             const __grid_constant__ GradGroupsDev gsrc, OptSink sink, float* __restrict__ head_part,
                             float* __restrict__ tail_part, int* __restrict__ long_count) {
  griddep_launch_dependents();
  if (blockIdx.x == 0 && threadIdx.x == 0) *long_count = 0;   // seg_chain_kernel appends to the list only after this grid is complete
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  sink.prepare();
  constexpr int kGroups = kSegThreads / GS;
  constexpr int kEntries = kGroups * kTile;
  constexpr int kKinds = 4;                            // gradient, W, state0, state1
  __shared__ uint32_t s_key[kEntries + 2];
  __shared__ uint32_t s_pos[kEntries];
  __shared__ __align__(8) uint64_t s_bar[kGroups * kBulkRing];
  extern __shared__ __align__(16) float ring[];        // [kGroups][kBulkRing][kKinds][D]

  const int D = gsrc.D;
  const uint32_t row_bytes = static_cast<uint32_t>(D) * 4u;
  const int cta_base = blockIdx.x * kEntries;
  const int cta_cnt = min(kEntries, n - cta_base);
  if (cta_cnt <= 0) return;
  for (int i = threadIdx.x; i < cta_cnt; i += kSegThreads) {
    s_key[i + 1] = keys[cta_base + i];
    s_pos[i] = vals[cta_base + i];
  }
  if (threadIdx.x == 0) {
    s_key[0] = (cta_base > 0) ? keys[cta_base - 1] : 0u;
    s_key[cta_cnt + 1] = (cta_base + cta_cnt < n) ? keys[cta_base + cta_cnt] : 0u;
  }
  if (threadIdx.x < kGroups * kBulkRing) su_mbar_init(&s_bar[threadIdx.x], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int group = threadIdx.x / GS;
  const int lane = threadIdx.x % GS;
  const int tbase = group * kTile;
  const int tcnt = min(kTile, cta_cnt - tbase);
  if (tcnt <= 0) return;
  const unsigned gmask = (GS == 32) ? 0xFFFFFFFFu : (((1u << GS) - 1u) << ((threadIdx.x % 32) / GS * GS));
  const bool active = lane * VEC < D;
  const int c = lane * VEC;
  const int gbase = cta_base + tbase;
  const int tile_id = gbase / kTile;
  const bool has_prev = gbase > 0;
  const bool has_next = gbase + tcnt < n;
  const uint32_t* tk = s_key + 1 + tbase;
  const uint32_t* tp = s_pos + tbase;
  const uint32_t first_key = tk[0];
  const bool cont_first = has_prev && (tk[-1] == first_key);
  float* my_ring = ring + static_cast<size_t>(group) * kBulkRing * kKinds * D;
  uint64_t* my_bar = s_bar + group * kBulkRing;
  const int opt = sink.opt;
  const bool ld_w = opt != RB_OPT_ADAM_TF_DENSE, ld_s0 = opt != RB_OPT_SGD, ld_s1 = (opt == RB_OPT_ADAM_LAZY || opt == RB_OPT_ADAM_TF_DENSE);
  const uint32_t state_bytes = row_bytes * (static_cast<uint32_t>(ld_w) + static_cast<uint32_t>(ld_s0) + static_cast<uint32_t>(ld_s1));

  auto run_ends_at = [&](int j, uint32_t key) {
    return (j == tcnt - 1) ? !(has_next && tk[tcnt] == key) : (tk[j + 1] != key);
  };
  const float* const flat_src = gsrc.g[0].src[0];
  const int64_t flat_stride = gsrc.g[0].pos_stride[0];
  const uint64_t evict_first = su_policy_evict_first();
  const bool stream_grads = RB_SEG_EVICT_FIRST != 0 && gsrc.peer == 0;
  auto issue = [&](int j, int slot) {            // lane 0 of the group only
    if (j >= tcnt) return;
    const uint32_t key = tk[j];
    uint64_t* bar = my_bar + slot;
    float* dst = my_ring + slot * kKinds * D;
    const bool update = run_ends_at(j, key) && !(cont_first && key == first_key);
    su_mbar_expect_tx(bar, row_bytes + (update ? state_bytes : 0u));
    const float* grow;
    if constexpr (FLAT) grow = flat_src + static_cast<int64_t>(tp[j]) * flat_stride;
    else grow = grad_src0(gsrc, decode_pos(gsrc, tp[j]), 0);
    if (stream_grads) su_bulk_g2s_hint(dst, grow, row_bytes, bar, evict_first);
    else su_bulk_g2s(dst, grow, row_bytes, bar);
    if (update) {
      const int64_t o = static_cast<int64_t>(key) * D;
      if (ld_w) su_bulk_g2s(dst + D, sink.table + o, row_bytes, bar);
      if (ld_s0) su_bulk_g2s(dst + 2 * D, sink.s0 + o, row_bytes, bar);
      if (ld_s1) su_bulk_g2s(dst + 3 * D, sink.s1 + o, row_bytes, bar);
    }
  };

  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kBulkRing; ++j) issue(j, j);
  }
  Row<VEC> acc = zero_row<VEC>();
  uint32_t cur = first_key;
  int seg_start = 0;
  bool continues_prev = cont_first;
  int slot = 0;
  uint32_t parity = 0;
  for (int j = 0; j < tcnt; ++j) {
    su_mbar_wait(my_bar + slot, parity);
    const uint32_t key = tk[j];
    const float* sl = my_ring + slot * kKinds * D;
    Row<VEC> g = zero_row<VEC>(), w = zero_row<VEC>(), m = zero_row<VEC>(), v = zero_row<VEC>();
    if (key != cur) {
      cur = key;
      acc = zero_row<VEC>();
      seg_start = j;
      continues_prev = false;
    }
    const bool ends = run_ends_at(j, cur);
    const bool update = ends && !continues_prev;
    if (active) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) g.v[i] = sl[c + i];
      if (update) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          w.v[i] = sl[D + c + i];
          m.v[i] = sl[2 * D + c + i];
          v.v[i] = sl[3 * D + c + i];
        }
      }
    }
    __syncwarp(gmask);                           // every lane of the group has read the slot: it may be refilled
    if (lane == 0) issue(j + kBulkRing, slot);
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc.v[i] = __fadd_rn(acc.v[i], g.v[i]);
    if (ends) {
      if (continues_prev) {
        if (active) st_row<VEC>(head_part + static_cast<int64_t>(tile_id) * D + c, acc);
      } else if (active) {
        sink.template update_row<VEC>(static_cast<int64_t>(cur) * D + c, acc, w, m, v);
      }
    } else if (j == tcnt - 1 && active) {
      float* dst = continues_prev ? head_part : tail_part;
      st_row<VEC>(dst + static_cast<int64_t>(tile_id) * D + c, acc);
    }
    if (++slot == kBulkRing) {
      slot = 0;
      parity ^= 1;
    }
  }
}

// ---- step 3: runs that cross tile borders -------------------------------------------------------------------
__device__ __forceinline__ int run_end(const uint32_t* __restrict__ keys, int lo, int n, uint32_t key) {
  // first index in [lo, n) whose key differs from `key` (keys are sorted, keys[lo-1] == key)
  int hi = n;
  while (lo < hi) {
    const int mid = lo + (hi - lo) / 2;
    if (keys[mid] == key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <int VEC, int GS, class Sink>
__global__ void __launch_bounds__(kSegThreads)
seg_chain_kernel(const uint32_t* __restrict__ keys, int n, const int* __restrict__ n_dev, int D, Sink sink,
                 const float* __restrict__ head_part, const float* __restrict__ tail_part, LongChain* __restrict__ long_list,
                 int* __restrict__ long_count, int long_cap) {
  griddep_launch_dependents();
  if (n_dev != nullptr) n = min(n, __ldg(n_dev));
  sink.prepare();
  const int tile_id = blockIdx.x * (kSegThreads / GS) + threadIdx.x / GS;
  const int lane = threadIdx.x % GS;
  const int gbase = tile_id * kTile;
  // the sorted keys alone say whether this tile owns a run that crosses its border: decided while the tile kernel still runs
  bool owns = gbase + kTile < n;                        // else no following tile: nothing can continue
  uint32_t key = 0;
  if (owns) {
    key = keys[gbase + kTile - 1];
    owns = keys[gbase + kTile] == key;                  // else the trailing run ends here
    if (owns && keys[gbase] == key && gbase > 0 && keys[gbase - 1] == key) owns = false;  // interior tile of someone else's chain
  }
  const bool anchor = blockIdx.x == 0 && threadIdx.x == 0;   // this grid must not complete before the tile kernel has
  if (!owns && !anchor) return;
  griddep_wait();                                       // the tile kernel's partial sums (and its rows) are complete and visible
  if (!owns) return;
  // this tile owns the run: find where it starts (inside the tile) and where it ends
  int start = kTile - 1;
  while (start > 0 && keys[gbase + start - 1] == key) --start;
  const int end = run_end(keys, gbase + kTile, n, key);  // exclusive
  const int last_tile = (end - 1) / kTile;
  const int chain = last_tile - tile_id;                 // following tiles that hold a head partial
  if (chain > kLongChain) {
    if (lane == 0) {
      const int slot = atomicAdd(long_count, 1);         // integer append; order does not affect values
      if (slot < long_cap) long_list[slot] = LongChain{key, static_cast<uint32_t>(tile_id), static_cast<uint32_t>(last_tile),
                                                            static_cast<uint32_t>(gbase + start)};
    }
    return;
  }
  if (lane * VEC >= D) return;
  Row<VEC> acc = ld_row_rw<VEC>(tail_part + static_cast<int64_t>(tile_id) * D + lane * VEC);
  for (int t0 = tile_id + 1; t0 <= last_tile; t0 += kChunk) {
    Row<VEC> h[kChunk];
#pragma unroll
    for (int u = 0; u < kChunk; ++u)
      h[u] = (t0 + u <= last_tile) ? ld_row_rw<VEC>(head_part + static_cast<int64_t>(t0 + u) * D + lane * VEC) : zero_row<VEC>();
#pragma unroll
    for (int u = 0; u < kChunk; ++u)
      if (t0 + u <= last_tile) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc.v[i] = __fadd_rn(acc.v[i], h[u].v[i]);
      }
  }
  sink.template apply<VEC>(key, acc, lane, static_cast<uint32_t>(gbase + start));
}

template <int VEC, int GS, class Sink>
__global__ void __launch_bounds__(kSegThreads)
seg_long_chain_kernel(int D, Sink sink, const float* __restrict__ head_part, const float* __restrict__ tail_part,
                      const LongChain* __restrict__ long_list, const int* __restrict__ long_count, int long_cap) {
  griddep_wait();                                       // the list seg_chain_kernel appended to is complete
  sink.prepare();
  constexpr int kGroups = kSegThreads / GS;
  __shared__ float s_part[kGroups * 128];  // D <= 128
  const int group = threadIdx.x / GS, lane = threadIdx.x % GS;
  const bool active = lane * VEC < D;
  const int count = min(*long_count, long_cap);
  for (int c = blockIdx.x; c < count; c += gridDim.x) {
    const LongChain lc = long_list[c];
    const int ntiles = static_cast<int>(lc.last_tile - lc.first_tile);  // head partials to add
    const int per = (ntiles + kGroups - 1) / kGroups;
    const int t_begin = static_cast<int>(lc.first_tile) + 1 + group * per;
    const int t_end = min(t_begin + per, static_cast<int>(lc.last_tile) + 1);
    Row<VEC> acc = zero_row<VEC>();
    if (active) {
      for (int t0 = t_begin; t0 < t_end; t0 += kChunk) {
        Row<VEC> h[kChunk];
#pragma unroll
        for (int u = 0; u < kChunk; ++u)
          h[u] = (t0 + u < t_end) ? ld_row_rw<VEC>(head_part + static_cast<int64_t>(t0 + u) * D + lane * VEC) : zero_row<VEC>();
#pragma unroll
        for (int u = 0; u < kChunk; ++u)
          if (t0 + u < t_end) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc.v[i] = __fadd_rn(acc.v[i], h[u].v[i]);
          }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) s_part[group * 128 + lane * VEC + i] = acc.v[i];
    }
    __syncthreads();
    if (group == 0 && active) {
      Row<VEC> tot = ld_row_rw<VEC>(tail_part + static_cast<int64_t>(lc.first_tile) * D + lane * VEC);
      for (int gidx = 0; gidx < kGroups; ++gidx) {  // fixed order -> deterministic
#pragma unroll
        for (int i = 0; i < VEC; ++i) tot.v[i] = __fadd_rn(tot.v[i], s_part[gidx * 128 + lane * VEC + i]);
      }
      sink.template apply<VEC>(lc.row, tot, lane, lc.seg_first);
    }
    __syncthreads();
  }
}

// ---- Keras-exact Adam: the dense passes over every row ----------------------------------------
__global__ void adam_decay_all_kernel(float* __restrict__ m, float* __restrict__ v, int64_t count, float b1, float b2) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    m[i] = __fmul_rn(m[i], b1);
    v[i] = __fmul_rn(v[i], b2);
  }
}
__global__ void adam_apply_all_kernel(float* __restrict__ w, const float* __restrict__ m, const float* __restrict__ v,
                                      int64_t count, float alpha, const float* __restrict__ alpha_dev, float eps) {
  if (alpha_dev != nullptr) alpha = __ldg(alpha_dev);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    w[i] = __fsub_rn(w[i], __fdiv_rn(__fmul_rn(alpha, m[i]), __fadd_rn(__fsqrt_rn(v[i]), eps)));
  }
}

// ---- workspace -----------------------------------------------------------------------------------------------
struct WsLayout {
  size_t keys_a, keys_b, vals_a, vals_b, seg_incl, head_part, tail_part, long_list, long_count, cub_temp, cub_bytes, total;
};

// floor(2^32 / L) clamped to 32 bits (L = 1): umulhi(p, r) is p / L or one less
static uint32_t recip32(int L) {
  const uint64_t r = (1ull << 32) / static_cast<uint64_t>(L);
  return r > 0xFFFFFFFFull ? 0xFFFFFFFFu : static_cast<uint32_t>(r);
}

static size_t align_up(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// a chain is "long" when it spans more than kLongChain tiles, so at most this many can exist
static int long_list_cap(int64_t n) { return static_cast<int>(n / (static_cast<int64_t>(kLongChain) * kTile)) + 2; }

static int key_bits(int64_t rows) {
  int bits = 1;
  while (bits < 32 && (int64_t(1) << bits) < rows) ++bits;
  return bits;
}

static WsLayout ws_layout(int64_t n, int D, int64_t rows) {
  WsLayout w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    return o;
  };
  const size_t n4 = static_cast<size_t>(n) * 4;
  const size_t tiles = static_cast<size_t>((n + kTile - 1) / kTile) + 1;
  w.keys_a = take(n4);
  w.keys_b = take(n4);
  w.vals_a = take(n4);
  w.vals_b = take(n4);
  w.seg_incl = take(n4);
  w.head_part = take(tiles * D * 4);
  w.tail_part = take(tiles * D * 4);
  w.long_list = take(sizeof(LongChain) * static_cast<size_t>(long_list_cap(n)));
  w.long_count = take(256);
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, dk, dv, static_cast<int>(n), 0, key_bits(rows + 1));
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                static_cast<int>(n));
  w.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  w.cub_temp = take(w.cub_bytes + 256);
  w.total = off;
  return w;
}

static int fill_grad_src(GradSrcDev* d, const rb_grad_source* g, int L, int idx_type, const void* idx, int vec) {
  RB_CHECK_ARG(g != nullptr && g->num_src >= 1 && g->num_src <= RB_MAX_GRAD_SOURCES, RB_ERR_ARG, "grad source count out of range");
  RB_CHECK_ARG(g->scale_mode >= RB_SCALE_NONE && g->scale_mode <= RB_SCALE_MASKED, RB_ERR_ARG, "bad scale mode");
  d->num_src = g->num_src;
  d->scale_mode = g->scale_mode;
  d->L = L;
  d->L_recip = recip32(L);
  d->is64 = (idx_type == RB_I64);
  for (int k = 0; k < RB_MAX_GRAD_SOURCES; ++k) {
    d->src[k] = nullptr;
    d->bag_stride[k] = d->pos_stride[k] = 0;
  }
  for (int k = 0; k < g->num_src; ++k) {
    RB_CHECK_ARG(g->src[k] != nullptr, RB_ERR_ARG, "grad source %d is null", k);
    RB_CHECK_ARG(aligned_for(g->src[k], vec) && g->bag_stride[k] % vec == 0 && g->pos_stride[k] % vec == 0, RB_ERR_ALIGN,
                 "grad source %d not aligned for vec=%d", k, vec);
    d->src[k] = g->src[k];
    d->bag_stride[k] = g->bag_stride[k];
    d->pos_stride[k] = g->pos_stride[k];
  }
  d->mask_idx = g->mask_idx != nullptr ? g->mask_idx : idx;
  d->count = g->count;
  RB_CHECK_ARG(g->scale_mode != RB_SCALE_MASKED_MEAN || g->count != nullptr, RB_ERR_ARG, "masked mean needs count");
  d->fm_g = g->fm_g;
  d->fm_s = g->fm_s;
  RB_CHECK_ARG(g->fm_g == nullptr || (g->fm_s != nullptr && aligned_for(g->fm_s, vec)), RB_ERR_ARG, "fm_g needs an aligned fm_s");
  return RB_OK;
}

// RB_SEG_BULK=0 in the environment keeps the per-lane cp.async kernel (A/B measurements)
static bool bulk_enabled() {
  static const bool on = [] {
    const char* e = getenv("RB_SEG_BULK");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

// RB_SEG_PAD_KB=n adds n KiB to the bulk kernel's dynamic shared memory (occupancy experiments: fewer resident CTAs per SM)
static size_t bulk_pad_bytes() {
  static const size_t pad = [] {
    const char* e = getenv("RB_SEG_PAD_KB");
    return e == nullptr ? size_t(0) : static_cast<size_t>(atoi(e)) * 1024;
  }();
  return pad;
}

// gradient rows in PEER memory (sharded path): opt-in until measured over NVLink
static bool bulk_peer_enabled() {
  static const bool on = [] {
    const char* e = getenv("RB_SEG_BULK_PEER");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

template <class Sink>
static int run_segments(const RowGeom& geo, const uint32_t* keys, const uint32_t* vals, int n, const GradGroupsDev& gsrc,
                        const Sink& sink, unsigned char* ws, const WsLayout& lay, cudaStream_t st, const int* n_dev = nullptr) {
  float* head = reinterpret_cast<float*>(ws + lay.head_part);
  float* tail = reinterpret_cast<float*>(ws + lay.tail_part);
  LongChain* ll = reinterpret_cast<LongChain*>(ws + lay.long_list);
  int* lc = reinterpret_cast<int*>(ws + lay.long_count);
  static const bool pdl = [] { const char* e = getenv("RB_SEG_PDL"); return e == nullptr || e[0] != '0'; }();
  const int tiles = (n + kTile - 1) / kTile;
  const int cap = long_list_cap(n);
  const GradSrcDev& g0 = gsrc.g[0];
  const bool simple = gsrc.num == 1 && g0.num_src == 1 && g0.scale_mode == RB_SCALE_NONE && g0.fm_g == nullptr;
  // plain gradient rows (every use of the table: one source tensor, no scaling, no FM term) of 16-byte multiples, all
  // 16-byte aligned: the row-granular bulk-copy kernel
  bool bulk = std::is_same<Sink, OptSink>::value && geo.vec == 4 && bulk_enabled() && (gsrc.peer == 0 || bulk_peer_enabled());
  for (int k = 0; k < gsrc.num && bulk; ++k) {
    const GradSrcDev& gk = gsrc.g[k];
    bulk = gk.num_src == 1 && gk.scale_mode == RB_SCALE_NONE && gk.fm_g == nullptr && gk.bag_stride[0] % 4 == 0 && gk.pos_stride[0] % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(gk.src[0]) & 15) == 0;
  }
  const bool flat = bulk && gsrc.num == 1 && g0.bag_stride[0] == static_cast<int64_t>(g0.L) * g0.pos_stride[0];
#define CALL(V, G)                                                                                                      \
  {                                                                                                                     \
    constexpr int kGroups = kSegThreads / G;                                                                            \
    const size_t ring_bytes = static_cast<size_t>(kRing) * (1 + Sink::kStateRows) * kSegThreads * V * sizeof(float);   \
    if (bulk) {                                                                                                         \
      const size_t bulk_bytes = static_cast<size_t>(kGroups) * kBulkRing * 4 * gsrc.D * sizeof(float) + bulk_pad_bytes(); \
      if constexpr (std::is_same<Sink, OptSink>::value) {                                                               \
        if (flat) {                                                                                                     \
          RB_CUDA(cudaFuncSetAttribute(seg_reduce_tiles_bulk_kernel<V, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       static_cast<int>(bulk_bytes)));                                                  \
          seg_reduce_tiles_bulk_kernel<V, G, true><<<grid_for(tiles, kGroups), kSegThreads, bulk_bytes, st>>>(keys, vals, n, n_dev, \
                                                                                                              gsrc, sink, head, tail, lc); \
        } else {                                                                                                        \
          RB_CUDA(cudaFuncSetAttribute(seg_reduce_tiles_bulk_kernel<V, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       static_cast<int>(bulk_bytes)));                                                  \
          seg_reduce_tiles_bulk_kernel<V, G, false><<<grid_for(tiles, kGroups), kSegThreads, bulk_bytes, st>>>(keys, vals, n, n_dev, \
                                                                                                               gsrc, sink, head, tail, lc); \
        }                                                                                                               \
      }                                                                                                                 \
    } else if (simple) {                                                                                                \
      if (ring_bytes > 40 * 1024)                                                                                       \
        RB_CUDA(cudaFuncSetAttribute(seg_reduce_tiles_kernel<V, G, Sink, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     static_cast<int>(ring_bytes)));                                                    \
      seg_reduce_tiles_kernel<V, G, Sink, true><<<grid_for(tiles, kGroups), kSegThreads, ring_bytes, st>>>(keys, vals, n, n_dev, \
                                                                                                            gsrc, sink, head, tail, lc); \
    } else {                                                                                                            \
      if (ring_bytes > 40 * 1024)                                                                                       \
        RB_CUDA(cudaFuncSetAttribute(seg_reduce_tiles_kernel<V, G, Sink, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     static_cast<int>(ring_bytes)));                                                    \
      seg_reduce_tiles_kernel<V, G, Sink, false><<<grid_for(tiles, kGroups), kSegThreads, ring_bytes, st>>>(keys, vals, n, n_dev, \
                                                                                                             gsrc, sink, head, tail, lc); \
    }                                                                                                                   \
    RB_CUDA(launch_dependent(seg_chain_kernel<V, G, Sink>, grid_for(tiles, kGroups), kSegThreads, 0, st, pdl, keys, n, n_dev, gsrc.D, sink, \
                             head, tail, ll, lc, cap));                                                                  \
    RB_CUDA(launch_dependent(seg_long_chain_kernel<V, G, Sink>, 2 * kNumSMs, kSegThreads, 0, st, pdl, gsrc.D, sink, head, tail, ll, lc, cap)); \
  }
  if (geo.vec == 4) {
    if (geo.gs == 4) CALL(4, 4) else if (geo.gs == 8) CALL(4, 8) else if (geo.gs == 16) CALL(4, 16) else CALL(4, 32)
  } else if (geo.vec == 2) {
    if (geo.gs == 4) CALL(2, 4) else if (geo.gs == 8) CALL(2, 8) else if (geo.gs == 16) CALL(2, 16) else CALL(2, 32)
  } else {
    if (geo.gs == 4) CALL(1, 4) else if (geo.gs == 8) CALL(1, 8) else if (geo.gs == 16) CALL(1, 16) else CALL(1, 32)
  }
#undef CALL
  RB_LAUNCH_CHECK("segmented reduction kernels");
  count_launches(2);  // three kernels above, one counted by the check
  return RB_OK;
}

static int check_common(int64_t rows, int D, int64_t n, RowGeom* geo) {
  RB_CHECK_ARG(rows > 0 && rows <= 0xFFFFFFFFll, RB_ERR_ARG, "rows must be in (0, 2^32)");
  RB_CHECK_ARG(row_geom(D, geo), RB_ERR_SHAPE, "unsupported embedding dim D=%d", D);
  RB_CHECK_ARG(D <= 128, RB_ERR_SHAPE, "D must be <= 128");
  RB_CHECK_ARG(n >= 0 && n < 0x7FFFFFFFll, RB_ERR_ARG, "the number of lookups must be in [0, 2^31)");
  return RB_OK;
}

// Phase 1 (depends on the ids only): writes and stably sorts the (row, position) pairs of all groups.
// *sel_out = which half of the double buffers holds the sorted pairs.
static int sort_groups(const rb_lookup_group* groups, int num_groups, int64_t rows, int64_t n, unsigned char* ws,
                       const WsLayout& lay, int* oob_flag, cudaStream_t st, int* sel_out, bool drop_masked = false,
                       bool* dropped_out = nullptr) {
  bool dropped = false;
  uint32_t* ka = reinterpret_cast<uint32_t*>(ws + lay.keys_a);
  uint32_t* kb = reinterpret_cast<uint32_t*>(ws + lay.keys_b);
  uint32_t* va = reinterpret_cast<uint32_t*>(ws + lay.vals_a);
  uint32_t* vb = reinterpret_cast<uint32_t*>(ws + lay.vals_b);
  int64_t start = 0;
  for (int k = 0; k < num_groups; ++k) {
    const rb_lookup_group& g = groups[k];
    RB_CHECK_ARG(g.idx_type == RB_I32 || g.idx_type == RB_I64, RB_ERR_ARG, "group %d: bad index type", k);
    RB_CHECK_ARG(g.n >= 0 && g.L > 0 && (g.n == 0 || g.idx != nullptr), RB_ERR_ARG, "group %d: bad n/L/idx", k);
    if (g.n > 0) {
      IndexMap m = make_index_map(g.idx, g.idx_type, g.field_row_offset, g.hash_mod, rows, g.L);
      const void* mask = nullptr;
      if (drop_masked && (g.grad.scale_mode == RB_SCALE_MASKED_MEAN || g.grad.scale_mode == RB_SCALE_MASKED) && rows < 0xFFFFFFFFll) {
        mask = g.grad.mask_idx != nullptr ? g.grad.mask_idx : g.idx;
        dropped = true;
      }
      make_keys_kernel<<<grid_for(g.n, 256), 256, 0, st>>>(m, g.n, start, ka, va, oob_flag, mask, static_cast<uint32_t>(rows));
      RB_LAUNCH_CHECK("make_keys_kernel");
    }
    start += g.n;
  }
  cub::DoubleBuffer<uint32_t> dk(ka, kb), dv(va, vb);
  size_t temp = lay.cub_bytes;
  // the padding key is `rows` itself: one more key bit when pairs were dropped
  const int bits = dropped ? key_bits(rows + 1) : key_bits(rows);
  RB_CUDA(cub::DeviceRadixSort::SortPairs(ws + lay.cub_temp, temp, dk, dv, static_cast<int>(n), 0, bits, st));
  *sel_out = dk.selector;
  if (dropped) {
    int* n_valid = reinterpret_cast<int*>(ws + lay.long_count) + 1;
    count_valid_kernel<<<1, 1, 0, st>>>(dk.Current(), static_cast<int>(n), static_cast<uint32_t>(rows), n_valid);
    RB_LAUNCH_CHECK("count_valid_kernel");
  }
  if (dropped_out != nullptr) *dropped_out = dropped;
  return RB_OK;
}

static void sorted_pairs(unsigned char* ws, const WsLayout& lay, int sel, const uint32_t** keys, const uint32_t** vals) {
  *keys = reinterpret_cast<const uint32_t*>(ws + (sel ? lay.keys_b : lay.keys_a));
  *vals = reinterpret_cast<const uint32_t*>(ws + (sel ? lay.vals_b : lay.vals_a));
}

// Phase 2 set-up: validates the groups' gradient sources and fills the device-side description.
static int describe_groups(const rb_lookup_group* groups, int num_groups, int D, const float* table, const RowGeom& geo,
                           GradGroupsDev* gg) {
  gg->num = num_groups;
  gg->D = D;
  gg->peer = 0;
  gg->table = table;
  for (int k = 0; k <= kMaxGroups; ++k) gg->start[k] = 0xFFFFFFFFu;
  int64_t start = 0;
  for (int k = 0; k < num_groups; ++k) {
    const rb_lookup_group& g = groups[k];
    RB_CHECK_ARG(g.idx_type == RB_I32 || g.idx_type == RB_I64, RB_ERR_ARG, "group %d: bad index type", k);
    RB_CHECK_ARG(g.n >= 0 && g.L > 0 && (g.n == 0 || g.idx != nullptr), RB_ERR_ARG, "group %d: bad n/L/idx", k);
    int rc = fill_grad_src(&gg->g[k], &g.grad, g.L, g.idx_type, g.idx, geo.vec);
    if (rc != RB_OK) return rc;
    RB_CHECK_ARG(g.grad.fm_g == nullptr || table != nullptr, RB_ERR_ARG, "the FM term needs the table");
    gg->start[k] = static_cast<uint32_t>(start);
    start += g.n;
  }
  for (int k = num_groups; k < kMaxGroups; ++k) gg->g[k] = gg->g[0];
  return RB_OK;
}

static OptSink make_sink(float* table, float* state0, float* state1, int D, const rb_opt_params* opt) {
  const int o = opt->optimizer;
  const bool adam = (o == RB_OPT_ADAM_LAZY || o == RB_OPT_ADAM_TF_DENSE);
  OptSink sink;
  sink.table = table;
  sink.s0 = state0;
  sink.s1 = state1;
  sink.D = D;
  sink.opt = o;
  sink.lr = opt->lr;
  sink.b1 = opt->beta_1;
  sink.b2 = opt->beta_2;
  sink.omb1 = 1.0f - opt->beta_1;
  sink.omb2 = 1.0f - opt->beta_2;
  sink.eps = opt->epsilon;
  sink.alpha = adam ? rb_adam_alpha_t(opt->lr, opt->beta_1, opt->beta_2, opt->step) : 0.f;
  sink.alpha_dev = adam ? opt->alpha_t_dev : nullptr;
  sink.shadow = nullptr;
  return sink;
}

static int check_opt(const rb_opt_params* opt, const float* state0, const float* state1, const RowGeom& geo) {
  RB_CHECK_ARG(opt != nullptr, RB_ERR_ARG, "opt is null");
  const int o = opt->optimizer;
  RB_CHECK_ARG(o >= RB_OPT_SGD && o <= RB_OPT_ADAM_TF_DENSE, RB_ERR_ARG, "bad optimizer %d", o);
  const bool adam = (o == RB_OPT_ADAM_LAZY || o == RB_OPT_ADAM_TF_DENSE);
  RB_CHECK_ARG(!adam || (state0 != nullptr && state1 != nullptr && opt->step >= 1), RB_ERR_ARG, "Adam needs m, v and step >= 1");
  RB_CHECK_ARG(o != RB_OPT_ADAGRAD || state0 != nullptr, RB_ERR_ARG, "Adagrad needs its accumulator");
  RB_CHECK_ARG((state0 == nullptr || aligned_for(state0, geo.vec)) && (state1 == nullptr || aligned_for(state1, geo.vec)),
               RB_ERR_ALIGN, "optimizer state not aligned for vec=%d", geo.vec);
  return RB_OK;
}

static int64_t total_lookups(const rb_lookup_group* groups, int num_groups) {
  int64_t n = 0;
  for (int k = 0; k < num_groups; ++k) n += groups[k].n > 0 ? groups[k].n : 0;
  return n;
}

static int check_ws(const void* ws, size_t ws_bytes, const WsLayout& lay) {
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= lay.total, RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", lay.total,
               ws_bytes);
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0, RB_ERR_ALIGN, "workspace must be 256 B aligned");
  return RB_OK;
}

}  // namespace rb

using namespace rb;

extern "C" size_t rb_sparse_bwd_update_workspace_bytes(int64_t n, int32_t D, int64_t rows) {
  if (n < 0 || n >= 0x7FFFFFFFll || D <= 0 || rows <= 0) return 0;
  return ws_layout(n > 0 ? n : 1, D, rows).total;
}

extern "C" float rb_adam_alpha_t(float lr, float beta_1, float beta_2, int32_t step) {
  const float t = static_cast<float>(step);
  const float b1p = powf(beta_1, t);
  const float b2p = powf(beta_2, t);
  return lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
}

// Hot-row census over the sorted keys: position i starts a window of kHotRun equal keys <=> its row takes at least kHotRun of
// the step's lookups.  The count is (run length - kHotRun + 1) summed over such rows; when it exceeds 1/16 of the lookups
// the flag says "copy table rows through L1" to the fused lookups that follow (RB_ROW_CACHE_AUTO).  Calibration (B = 65536,
// F = 26, kbench r2_10 / r2_11, bench r2_12): the bench's Zipf ids (P(id >= x) = x^-0.05, + 2 % id 0) on ONE shared table put
// 11 % of the lookups into such windows (18 rows with >= 4096 lookups each) — forward 347 -> 105 us, backward 447 -> 193 us
// through L1; the same ids over 26 tables (the hottest row of a table takes ~2.2k lookups: census 0) run 6 % SLOWER through
// L1 (111 -> 118 us) and uniform ids 20 % slower: both stay on L2.
constexpr int kHotRun = 4096;
__global__ void __launch_bounds__(256) hot_rows_count_kernel(const uint32_t* __restrict__ keys, int n, int* __restrict__ count) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const bool hot = (i + kHotRun - 1 < n) && keys[i] == keys[i + kHotRun - 1];
  const unsigned m = __ballot_sync(0xffffffffu, hot);
  if (threadIdx.x % 32 == 0 && m != 0) atomicAdd(count, __popc(m));        // integer count: order-independent
}
__global__ void hot_rows_flag_kernel(int* __restrict__ count, int n, int32_t* __restrict__ flag) {
  *flag = (static_cast<int64_t>(*count) * 16 > n) ? 1 : 0;
}

extern "C" int rb_sparse_bwd_prepare(int64_t rows, int32_t D, const rb_lookup_group* groups, int32_t num_groups, void* ws,
                                     size_t ws_bytes, int32_t* oob_flag, int32_t* sorted_sel, int32_t* hot_rows_flag, void* stream) {
  RB_CHECK_ARG(groups != nullptr && num_groups >= 1 && num_groups <= RB_MAX_LOOKUP_GROUPS, RB_ERR_ARG,
               "1..%d lookup groups, got %d", RB_MAX_LOOKUP_GROUPS, num_groups);
  RB_CHECK_ARG(sorted_sel != nullptr, RB_ERR_ARG, "sorted_sel is null");
  const int64_t n = total_lookups(groups, num_groups);
  RowGeom geo;
  int rc = check_common(rows, D, n, &geo);
  if (rc != RB_OK) return rc;
  *sorted_sel = 0;
  if (n == 0) return RB_OK;
  const WsLayout lay = ws_layout(n, D, rows);
  rc = check_ws(ws, ws_bytes, lay);
  if (rc != RB_OK) return rc;
  int sel = 0;
  rc = sort_groups(groups, num_groups, rows, n, static_cast<unsigned char*>(ws), lay, oob_flag, static_cast<cudaStream_t>(stream), &sel);
  *sorted_sel = sel;
  if (rc == RB_OK && hot_rows_flag != nullptr) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t *keys, *vals;
    sorted_pairs(static_cast<unsigned char*>(ws), lay, sel, &keys, &vals);
    int* count = reinterpret_cast<int*>(static_cast<unsigned char*>(ws) + lay.long_count) + 2;     // [0] long chains, [1] valid pairs
    RB_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
    hot_rows_count_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, static_cast<int>(n), count);
    RB_LAUNCH_CHECK("hot_rows_count_kernel");
    hot_rows_flag_kernel<<<1, 1, 0, st>>>(count, static_cast<int>(n), hot_rows_flag);
    RB_LAUNCH_CHECK("hot_rows_flag_kernel");
  }
  return rc;
}

// n_dev != null: only the first *n_dev sorted pairs are real (masked pads collapsed by sort_groups)
static int apply_sorted(float* table, float* state0, float* state1, int64_t rows, int32_t D, const rb_lookup_group* groups,
                        int32_t num_groups, const rb_opt_params* opt, void* ws, size_t ws_bytes, int32_t sorted_sel,
                        void* stream, const int* n_dev, int32_t flags = 0) {
  RB_CHECK_ARG(groups != nullptr && num_groups >= 1 && num_groups <= RB_MAX_LOOKUP_GROUPS, RB_ERR_ARG,
               "1..%d lookup groups, got %d", RB_MAX_LOOKUP_GROUPS, num_groups);
  RB_CHECK_ARG(sorted_sel == 0 || sorted_sel == 1, RB_ERR_ARG, "sorted_sel must come from rb_sparse_bwd_prepare");
  const int64_t n = total_lookups(groups, num_groups);
  RowGeom geo;
  int rc = check_common(rows, D, n, &geo);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(table != nullptr && opt != nullptr, RB_ERR_ARG, "table/opt is null");
  RB_CHECK_ARG(aligned_for(table, geo.vec), RB_ERR_ALIGN, "table not aligned for vec=%d", geo.vec);
  const int o = opt->optimizer;
  RB_CHECK_ARG(o >= RB_OPT_SGD && o <= RB_OPT_ADAM_TF_DENSE, RB_ERR_ARG, "bad optimizer %d", o);
  const bool adam = (o == RB_OPT_ADAM_LAZY || o == RB_OPT_ADAM_TF_DENSE);
  RB_CHECK_ARG(!adam || (state0 != nullptr && state1 != nullptr && opt->step >= 1), RB_ERR_ARG, "Adam needs m, v and step >= 1");
  RB_CHECK_ARG(o != RB_OPT_ADAGRAD || state0 != nullptr, RB_ERR_ARG, "Adagrad needs its accumulator");
  RB_CHECK_ARG((state0 == nullptr || aligned_for(state0, geo.vec)) && (state1 == nullptr || aligned_for(state1, geo.vec)),
               RB_ERR_ALIGN, "optimizer state not aligned for vec=%d", geo.vec);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const OptSink sink = make_sink(table, state0, state1, D, opt);

  const int64_t count = rows * D;
  if (o == RB_OPT_ADAM_TF_DENSE) {
    adam_decay_all_kernel<<<8 * kNumSMs, 256, 0, st>>>(state0, state1, count, sink.b1, sink.b2);
    RB_LAUNCH_CHECK("adam_decay_all_kernel");
  }
  if (n > 0) {
    const WsLayout lay = ws_layout(n, D, rows);
    rc = check_ws(ws, ws_bytes, lay);
    if (rc != RB_OK) return rc;
    GradGroupsDev gg;
    rc = describe_groups(groups, num_groups, D, table, geo, &gg);
    if (rc != RB_OK) return rc;
    const uint32_t *keys, *vals;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    sorted_pairs(wsb, lay, sorted_sel, &keys, &vals);
    RB_CHECK_ARG((flags & RB_APPLY_SKIP_SINGLETONS) == 0 || (o != RB_OPT_ADAM_TF_DENSE && num_groups == 1), RB_ERR_ARG,
                 "RB_APPLY_SKIP_SINGLETONS: one lookup group, row-sparse optimizers only");
    // the pairs of rows touched once were compacted away by rb_sparse_bwd_mark_singletons: the survivors' count is on the device
    if (flags & RB_APPLY_SKIP_SINGLETONS) n_dev = reinterpret_cast<const int*>(wsb + lay.long_count) + 1;
    rc = run_segments(geo, keys, vals, static_cast<int>(n), gg, sink, wsb, lay, st, n_dev);
    if (rc != RB_OK) return rc;
  }
  if (o == RB_OPT_ADAM_TF_DENSE) {
    adam_apply_all_kernel<<<8 * kNumSMs, 256, 0, st>>>(table, state0, state1, count, sink.alpha, sink.alpha_dev, sink.eps);
    RB_LAUNCH_CHECK("adam_apply_all_kernel");
  }
  return RB_OK;
}

extern "C" int rb_sparse_bwd_apply(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                                   const rb_lookup_group* groups, int32_t num_groups, const rb_opt_params* opt, void* ws,
                                   size_t ws_bytes, int32_t sorted_sel, void* stream) {
  return apply_sorted(table, state0, state1, rows, D, groups, num_groups, opt, ws, ws_bytes, sorted_sel, stream, nullptr);
}

extern "C" int rb_sparse_bwd_apply_ex(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                                      const rb_lookup_group* groups, int32_t num_groups, const rb_opt_params* opt, void* ws,
                                      size_t ws_bytes, int32_t sorted_sel, int32_t flags, void* stream) {
  RB_CHECK_ARG((flags & ~RB_APPLY_SKIP_SINGLETONS) == 0, RB_ERR_ARG, "unknown flags 0x%x", flags);
  return apply_sorted(table, state0, state1, rows, D, groups, num_groups, opt, ws, ws_bytes, sorted_sel, stream, nullptr, flags);
}

// single[p] = 1 when the row of lookup position p occurs exactly once among the step's sorted pairs; keep[i] = 1 for the sorted
// pairs of every other row
__global__ void __launch_bounds__(256)
mark_singletons_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n, uint8_t* __restrict__ single,
                       int32_t* __restrict__ keep) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = keys[i];
  const bool alone = (i == 0 || keys[i - 1] != k) && (i == n - 1 || keys[i + 1] != k);
  single[vals[i]] = alone ? 1 : 0;
  keep[i] = alone ? 0 : 1;
}

// stable compaction of the kept pairs (incl = inclusive scan of keep) into the other half of the double buffers
__global__ void __launch_bounds__(256)
compact_pairs_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n, const int32_t* __restrict__ incl,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int* __restrict__ n_kept) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int32_t c = incl[i];
  if (c != (i > 0 ? incl[i - 1] : 0)) {
    keys_out[c - 1] = keys[i];
    vals_out[c - 1] = vals[i];
  }
  if (i == n - 1) *n_kept = c;
}

extern "C" int rb_sparse_bwd_mark_singletons(int64_t rows, int32_t D, int64_t n, void* ws, size_t ws_bytes, int32_t* sorted_sel,
                                             uint8_t* single, void* stream) {
  RB_CHECK_ARG(sorted_sel != nullptr && (*sorted_sel == 0 || *sorted_sel == 1), RB_ERR_ARG, "sorted_sel must come from rb_sparse_bwd_prepare");
  RB_CHECK_ARG(single != nullptr, RB_ERR_ARG, "single is null");
  RowGeom geo;
  int rc = check_common(rows, D, n, &geo);
  if (rc != RB_OK) return rc;
  if (n == 0) return RB_OK;
  const WsLayout lay = ws_layout(n, D, rows);
  rc = check_ws(ws, ws_bytes, lay);
  if (rc != RB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  const uint32_t *keys, *vals;
  sorted_pairs(wsb, lay, *sorted_sel, &keys, &vals);
  int32_t* keep = reinterpret_cast<int32_t*>(wsb + lay.seg_incl);
  mark_singletons_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, vals, static_cast<int>(n), single, keep);
  RB_LAUNCH_CHECK("mark_singletons_kernel");
  size_t temp = lay.cub_bytes;
  RB_CUDA(cub::DeviceScan::InclusiveSum(wsb + lay.cub_temp, temp, keep, keep, static_cast<int>(n), st));
  const int other = 1 - *sorted_sel;
  uint32_t* keys_out = reinterpret_cast<uint32_t*>(wsb + (other ? lay.keys_b : lay.keys_a));
  uint32_t* vals_out = reinterpret_cast<uint32_t*>(wsb + (other ? lay.vals_b : lay.vals_a));
  int* n_kept = reinterpret_cast<int*>(wsb + lay.long_count) + 1;
  compact_pairs_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, vals, static_cast<int>(n), keep, keys_out, vals_out, n_kept);
  RB_LAUNCH_CHECK("compact_pairs_kernel");
  count_launches(2);
  *sorted_sel = other;
  return RB_OK;
}

// ---- tables replicated on every rank (p2p.py): one update from G short sorted (row, summed gradient) lists ------------------
// list k (rank k's contribution) is `cap` records of D + 1 floats: D gradient values and the row id (int32 bits) in the last
// one, rows ascending, padded behind the valid records with row ids >= rows.  One group of GS lanes per table row looks the row
// up in every list (binary search), adds the hits in rank order and applies the optimizer once if any list held the row —
// the lazy-update rule: untouched rows do not move.  Deterministic; every rank computes the same bits from the same lists.
template <int VEC, int GS>
__global__ void __launch_bounds__(256)
replicated_rows_update_kernel(OptSink sink, int rows, const float* __restrict__ lists, int G, int cap) {
  sink.prepare();
  const int D = sink.D;
  const int rec = D + 1;
  const int r = (blockIdx.x * 256 + threadIdx.x) / GS;
  const int lane = threadIdx.x % GS;
  if (r >= rows) return;
  const int c = lane * VEC;
  const bool active = c < D;
  Row<VEC> acc = zero_row<VEC>();
  bool touched = false;
  for (int k = 0; k < G; ++k) {
    const float* L = lists + static_cast<int64_t>(k) * cap * rec;
    int lo = 0, hi = cap;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__float_as_int(__ldg(L + static_cast<int64_t>(mid) * rec + D)) < r) lo = mid + 1;
      else hi = mid;
    }
    if (lo < cap && __float_as_int(__ldg(L + static_cast<int64_t>(lo) * rec + D)) == r) {
      touched = true;
      if (active) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc.v[i] = __fadd_rn(acc.v[i], __ldg(L + static_cast<int64_t>(lo) * rec + c + i));
      }
    }
  }
  if (touched && active) {
    const int64_t o = static_cast<int64_t>(r) * D + c;
    Row<VEC> w = zero_row<VEC>(), m = zero_row<VEC>(), v = zero_row<VEC>();
    w = ld_row_rw<VEC>(sink.table + o);
    if (sink.opt != RB_OPT_SGD) m = ld_row_rw<VEC>(sink.s0 + o);
    if (sink.opt == RB_OPT_ADAM_LAZY) v = ld_row_rw<VEC>(sink.s1 + o);
    sink.template update_row<VEC>(o, acc, w, m, v);
  }
}

extern "C" int rb_replicated_rows_update(float* table, float* state0, float* state1, int64_t rows, int32_t D, const float* lists,
                                         int32_t num_lists, int64_t cap, const rb_opt_params* opt, void* shadow_bf16, void* stream) {
  RB_CHECK_ARG(table != nullptr && lists != nullptr && opt != nullptr, RB_ERR_ARG, "table / lists / opt is null");
  RB_CHECK_ARG(rows > 0 && rows < 0x7FFFFFFFll && num_lists >= 1 && num_lists <= RB_MAX_RANKS && cap >= 1 && cap < 0x7FFFFFFFll, RB_ERR_ARG,
               "bad rows / num_lists / cap");
  RowGeom geo;
  RB_CHECK_ARG(row_geom(D, &geo) && D <= 128, RB_ERR_SHAPE, "unsupported embedding dim D=%d", D);
  const int o = opt->optimizer;
  RB_CHECK_ARG(o == RB_OPT_SGD || o == RB_OPT_ADAGRAD || o == RB_OPT_ADAM_LAZY, RB_ERR_ARG, "row-sparse optimizers only, got %d", o);
  int rc = check_opt(opt, state0, state1, geo);
  if (rc != RB_OK) return rc;
  OptSink sink = make_sink(table, state0, state1, D, opt);
  sink.shadow = static_cast<__nv_bfloat16*>(shadow_bf16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t threads = rows * geo.gs;
#define RRU(V, GSZ) replicated_rows_update_kernel<V, GSZ><<<grid_for(threads, 256), 256, 0, st>>>(sink, static_cast<int>(rows), lists, num_lists, static_cast<int>(cap))
  if (geo.vec == 4) {
    if (geo.gs == 4) RRU(4, 4); else if (geo.gs == 8) RRU(4, 8); else if (geo.gs == 16) RRU(4, 16); else RRU(4, 32);
  } else if (geo.vec == 2) {
    if (geo.gs == 4) RRU(2, 4); else if (geo.gs == 8) RRU(2, 8); else if (geo.gs == 16) RRU(2, 16); else RRU(2, 32);
  } else {
    if (geo.gs == 4) RRU(1, 4); else if (geo.gs == 8) RRU(1, 8); else if (geo.gs == 16) RRU(1, 16); else RRU(1, 32);
  }
#undef RRU
  RB_LAUNCH_CHECK("replicated_rows_update_kernel");
  return RB_OK;
}

extern "C" int rb_sparse_bwd_update_groups(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                                           const rb_lookup_group* groups, int32_t num_groups, const rb_opt_params* opt,
                                           void* ws, size_t ws_bytes, int32_t* oob_flag, void* stream) {
  // argument checks that must precede any launch (the two phases repeat the rest)
  RB_CHECK_ARG(table != nullptr && opt != nullptr, RB_ERR_ARG, "table/opt is null");
  RB_CHECK_ARG(groups != nullptr && num_groups >= 1 && num_groups <= RB_MAX_LOOKUP_GROUPS, RB_ERR_ARG,
               "1..%d lookup groups, got %d", RB_MAX_LOOKUP_GROUPS, num_groups);
  const int o = opt->optimizer;
  RB_CHECK_ARG(o >= RB_OPT_SGD && o <= RB_OPT_ADAM_TF_DENSE, RB_ERR_ARG, "bad optimizer %d", o);
  const bool adam = (o == RB_OPT_ADAM_LAZY || o == RB_OPT_ADAM_TF_DENSE);
  RB_CHECK_ARG(!adam || (state0 != nullptr && state1 != nullptr && opt->step >= 1), RB_ERR_ARG, "Adam needs m, v and step >= 1");
  RB_CHECK_ARG(o != RB_OPT_ADAGRAD || state0 != nullptr, RB_ERR_ARG, "Adagrad needs its accumulator");
  {
    RowGeom geo;
    int rc0 = check_common(rows, D, total_lookups(groups, num_groups), &geo);
    if (rc0 != RB_OK) return rc0;
    GradGroupsDev gg;
    rc0 = describe_groups(groups, num_groups, D, table, geo, &gg);
    if (rc0 != RB_OK) return rc0;
  }
  // The one-call form knows the gradient sources while it sorts, so masked-mean pads are collapsed (make_keys_kernel);
  // the two-phase form (rb_sparse_bwd_prepare reads ids only) keeps every pair.
  const int64_t n = total_lookups(groups, num_groups);
  int sel = 0;
  const int* n_dev = nullptr;
  if (n > 0) {
    const WsLayout lay = ws_layout(n, D, rows);
    int rc = check_ws(ws, ws_bytes, lay);
    if (rc != RB_OK) return rc;
    bool dropped = false;
    rc = sort_groups(groups, num_groups, rows, n, static_cast<unsigned char*>(ws), lay, oob_flag, static_cast<cudaStream_t>(stream),
                     &sel, true, &dropped);
    if (rc != RB_OK) return rc;
    if (dropped) n_dev = reinterpret_cast<const int*>(static_cast<unsigned char*>(ws) + lay.long_count) + 1;
  }
  return apply_sorted(table, state0, state1, rows, D, groups, num_groups, opt, ws, ws_bytes, sel, stream, n_dev);
}

extern "C" int rb_sparse_bwd_update(float* table, float* state0, float* state1, int64_t rows, int32_t D,
                                    const void* idx, int32_t idx_type, int64_t n, int32_t L,
                                    const int64_t* field_row_offset, int64_t hash_mod, const rb_grad_source* grad,
                                    const rb_opt_params* opt, void* ws, size_t ws_bytes, int32_t* oob_flag, void* stream) {
  RB_CHECK_ARG(grad != nullptr, RB_ERR_ARG, "grad is null");
  rb_lookup_group g;
  g.idx = idx;
  g.idx_type = idx_type;
  g.L = L;
  g.n = n;
  g.field_row_offset = field_row_offset;
  g.hash_mod = hash_mod;
  g.grad = *grad;
  return rb_sparse_bwd_update_groups(table, state0, state1, rows, D, &g, 1, opt, ws, ws_bytes, oob_flag, stream);
}

extern "C" int rb_sparse_bwd_dedup(int64_t rows, int32_t D, const void* idx, int32_t idx_type, int64_t n, int32_t L,
                                   const int64_t* field_row_offset, int64_t hash_mod, const rb_grad_source* grad,
                                   int64_t* uniq_rows, float* uniq_grad, int64_t* num_unique, void* ws, size_t ws_bytes,
                                   int32_t* oob_flag, void* stream) {
  RowGeom geo;
  int rc = check_common(rows, D, n, &geo);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(num_unique != nullptr, RB_ERR_ARG, "num_unique is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    RB_CUDA(cudaMemsetAsync(num_unique, 0, sizeof(int64_t), st));
    return RB_OK;
  }
  RB_CHECK_ARG(uniq_rows != nullptr && uniq_grad != nullptr && aligned_for(uniq_grad, geo.vec), RB_ERR_ARG,
               "uniq_rows/uniq_grad null or misaligned");
  RB_CHECK_ARG(grad != nullptr && grad->fm_g == nullptr, RB_ERR_ARG, "the dedup entry point takes no FM term (it has no table)");
  const WsLayout lay = ws_layout(n, D, rows);
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= lay.total, RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
               lay.total, ws_bytes);
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0, RB_ERR_ALIGN, "workspace must be 256 B aligned");
  rb_lookup_group g;
  g.idx = idx;
  g.idx_type = idx_type;
  g.L = L;
  g.n = n;
  g.field_row_offset = field_row_offset;
  g.hash_mod = hash_mod;
  g.grad = *grad;
  GradGroupsDev gg;
  const uint32_t *keys, *vals;
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  rc = describe_groups(&g, 1, D, nullptr, geo, &gg);
  if (rc != RB_OK) return rc;
  int sel = 0;
  rc = sort_groups(&g, 1, rows, n, wsb, lay, oob_flag, st, &sel);
  if (rc != RB_OK) return rc;
  sorted_pairs(wsb, lay, sel, &keys, &vals);
  int32_t* seg = reinterpret_cast<int32_t*>(wsb + lay.seg_incl);
  head_flags_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, static_cast<int>(n), seg);
  RB_LAUNCH_CHECK("head_flags_kernel");
  size_t temp = lay.cub_bytes;
  RB_CUDA(cub::DeviceScan::InclusiveSum(wsb + lay.cub_temp, temp, seg, seg, static_cast<int>(n), st));
  write_num_unique_kernel<<<1, 1, 0, st>>>(seg, static_cast<int>(n), num_unique);
  RB_LAUNCH_CHECK("write_num_unique_kernel");
  DedupSink sink{seg, uniq_rows, uniq_grad, D};
  return run_segments(geo, keys, vals, static_cast<int>(n), gg, sink, wsb, lay, st);
}

// ---- sharded path: pairs collected from the peers' bucket arrays (p2p.cu) ---------------------------------------

namespace rb {
int sparse_ws_key_buffers(int64_t n, int D, int64_t rows, void* ws, uint32_t** keys, uint32_t** vals) {
  const WsLayout lay = ws_layout(n, D, rows);
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  *keys = reinterpret_cast<uint32_t*>(wsb + lay.keys_a);
  *vals = reinterpret_cast<uint32_t*>(wsb + lay.vals_a);
  return RB_OK;
}
}  // namespace rb

extern "C" int rb_sparse_bwd_prepare_collected(int64_t local_rows, int32_t D, int64_t capacity, void* ws, size_t ws_bytes,
                                               int32_t* sorted_sel, void* stream) {
  RB_CHECK_ARG(sorted_sel != nullptr, RB_ERR_ARG, "sorted_sel is null");
  RowGeom geo;
  int rc = check_common(local_rows + 1, D, capacity, &geo);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(capacity > 0, RB_ERR_ARG, "capacity must be positive");
  const WsLayout lay = ws_layout(capacity, D, local_rows + 1);
  rc = check_ws(ws, ws_bytes, lay);
  if (rc != RB_OK) return rc;
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  cub::DoubleBuffer<uint32_t> dk(reinterpret_cast<uint32_t*>(wsb + lay.keys_a), reinterpret_cast<uint32_t*>(wsb + lay.keys_b));
  cub::DoubleBuffer<uint32_t> dv(reinterpret_cast<uint32_t*>(wsb + lay.vals_a), reinterpret_cast<uint32_t*>(wsb + lay.vals_b));
  size_t temp = lay.cub_bytes;
  // the padding key is `local_rows` itself: key_bits(local_rows + 1) covers it and it sorts behind every real pair
  RB_CUDA(cub::DeviceRadixSort::SortPairs(wsb + lay.cub_temp, temp, dk, dv, static_cast<int>(capacity), 0, key_bits(local_rows + 1),
                                          static_cast<cudaStream_t>(stream)));
  *sorted_sel = dk.selector;
  return RB_OK;
}

extern "C" int rb_sparse_bwd_apply_p2p(float* table, float* state0, float* state1, int64_t local_rows, int32_t D, int32_t world,
                                       int64_t n_local, int32_t L, const void* const* dE_ptrs, int64_t capacity,
                                       const int32_t* n_valid_dev, const rb_opt_params* opt, void* ws, size_t ws_bytes,
                                       int32_t sorted_sel, void* shadow_bf16, void* stream) {
  RB_CHECK_ARG(world >= 1 && world <= RB_MAX_RANKS && dE_ptrs != nullptr && n_valid_dev != nullptr, RB_ERR_ARG,
               "world must be in [1, %d]; dE_ptrs / n_valid_dev must not be null", RB_MAX_RANKS);
  RB_CHECK_ARG(n_local > 0 && L > 0 && n_local % L == 0 && n_local * world < 0xFFFFFFFFll, RB_ERR_ARG, "bad n_local / L");
  RB_CHECK_ARG(sorted_sel == 0 || sorted_sel == 1, RB_ERR_ARG, "sorted_sel must come from rb_sparse_bwd_prepare_collected");
  RowGeom geo;
  int rc = check_common(local_rows + 1, D, capacity, &geo);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(table != nullptr && aligned_for(table, geo.vec), RB_ERR_ALIGN, "table null or not aligned for vec=%d", geo.vec);
  rc = check_opt(opt, state0, state1, geo);
  if (rc != RB_OK) return rc;
  RB_CHECK_ARG(opt->optimizer != RB_OPT_ADAM_TF_DENSE, RB_ERR_ARG, "the sharded apply supports adam_lazy / adagrad / sgd");
  const WsLayout lay = ws_layout(capacity, D, local_rows + 1);
  rc = check_ws(ws, ws_bytes, lay);
  if (rc != RB_OK) return rc;
  // group k = source rank k: positions [k*n_local, (k+1)*n_local) address rank k's dE[B_local, L, D] (peer memory)
  GradGroupsDev gg;
  gg.num = world;
  gg.D = D;
  gg.peer = world > 1 ? 1 : 0;
  gg.table = table;
  for (int k = 0; k <= kMaxGroups; ++k) gg.start[k] = 0xFFFFFFFFu;
  for (int k = 0; k < kMaxGroups; ++k) {
    GradSrcDev& g = gg.g[k];
    const int src_rank = k < world ? k : 0;
    RB_CHECK_ARG(dE_ptrs[src_rank] != nullptr && aligned_for(dE_ptrs[src_rank], geo.vec), RB_ERR_ALIGN, "dE of rank %d null or misaligned", src_rank);
    g.num_src = 1;
    g.scale_mode = RB_SCALE_NONE;
    g.L = L;
    g.L_recip = recip32(L);
    g.is64 = 0;
    for (int j = 0; j < RB_MAX_GRAD_SOURCES; ++j) {
      g.src[j] = nullptr;
      g.bag_stride[j] = g.pos_stride[j] = 0;
    }
    g.src[0] = static_cast<const float*>(dE_ptrs[src_rank]);
    g.bag_stride[0] = static_cast<int64_t>(L) * D;
    g.pos_stride[0] = D;
    g.mask_idx = nullptr;
    g.count = nullptr;
    g.fm_g = nullptr;
    g.fm_s = nullptr;
    if (k < world) gg.start[k] = static_cast<uint32_t>(k * n_local);
  }
  OptSink sink = make_sink(table, state0, state1, D, opt);
  RB_CHECK_ARG(shadow_bf16 == nullptr || (reinterpret_cast<uintptr_t>(shadow_bf16) & 7) == 0, RB_ERR_ALIGN, "shadow not 8 B aligned");
  sink.shadow = static_cast<__nv_bfloat16*>(shadow_bf16);
  const uint32_t *keys, *vals;
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  sorted_pairs(wsb, lay, sorted_sel, &keys, &vals);
  return run_segments(geo, keys, vals, static_cast<int>(capacity), gg, sink, wsb, lay, static_cast<cudaStream_t>(stream), n_valid_dev);
}
