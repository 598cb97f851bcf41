// Row-wise sharding support: route every lookup to the rank that owns its row.
//
// owner = row mod world, local = row div world.  rb_bucket_by_owner is the send side of the
// index all-to-all: a STABLE partition of the lookups by owner (one radix pass over log2(world)
// bits), the local row ids in bucket order, the permutation and its inverse.  The inverse
// permutation is what lets the interaction / gather kernels consume the rows that come back
// from the owners in place: the receive buffer is addressed like a table whose "ids" are
// inv_perm, so no un-permute pass over [n, D] ever runs.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace rb {

// ---- stable partition by owner in three small kernels (world <= kMaxOwners) ------------------------------------------------
// A chunk of kChunk consecutive lookups per CTA, kItems consecutive lookups per thread, so "earlier thread, earlier item"
// is position order and the partition is stable by construction:
//   1. owner_count_kernel    per-chunk lookup counts per owner
//   2. owner_offsets_kernel  one CTA: per owner, the exclusive prefix over the chunks (+ the start of the owner's bucket)
//   3. owner_scatter_kernel  rank of every lookup inside its chunk (block scans of the per-thread counts) -> slot
// 20 B per lookup of traffic and no sort: the one-pass radix sort this replaces (cub, + a key-building and a finishing
// kernel) took ~100 us per step at 1.7M lookups (bench r2_13: p2p.route).
constexpr int kMaxOwners = 8;
constexpr int kOwnerThreads = 256;
constexpr int kItems = 8;
constexpr int kChunk = kOwnerThreads * kItems;

__device__ __forceinline__ int64_t row_or_zero(const IndexMap& m, int64_t p) {
  const int64_t row = map_index(m, p);
  return row < 0 ? 0 : row;  // out-of-range ids are clamped to row 0 (documented in the header)
}

// skip_from: lookups whose row is >= skip_from belong to no owner (tables replicated on every rank, p2p.py): they are neither
// counted nor given a slot
__global__ void __launch_bounds__(kOwnerThreads) owner_count_kernel(IndexMap m, int64_t n, int world, int64_t skip_from,
                                                                     int32_t* __restrict__ chunk_counts) {
  __shared__ int s_cnt[kMaxOwners];
  if (threadIdx.x < kMaxOwners) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kChunk + threadIdx.x * kItems;
  int cnt[kMaxOwners];
#pragma unroll
  for (int q = 0; q < kMaxOwners; ++q) cnt[q] = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    if (base + i < n) {
      const int64_t r = row_or_zero(m, base + i);
      const int o = r >= skip_from ? -1 : static_cast<int>(r % world);
#pragma unroll
      for (int q = 0; q < kMaxOwners; ++q) cnt[q] += (o == q) ? 1 : 0;
    }
  }
#pragma unroll
  for (int q = 0; q < kMaxOwners; ++q) {
    int c = cnt[q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (threadIdx.x % 32 == 0 && c != 0) atomicAdd(&s_cnt[q], c);          // integer counts: order-independent
  }
  __syncthreads();
  if (threadIdx.x < kMaxOwners) chunk_counts[static_cast<int64_t>(blockIdx.x) * kMaxOwners + threadIdx.x] = s_cnt[threadIdx.x];
}

// chunk_counts[c][q] -> chunk_offsets[c][q] = first slot of chunk c's lookups owned by q; counts[q] = lookups owned by q
__global__ void __launch_bounds__(1024) owner_offsets_kernel(int32_t* __restrict__ chunk_counts, int chunks, int world,
                                                              int64_t* __restrict__ counts) {
  __shared__ int s_part[32][kMaxOwners];
  __shared__ int s_base[kMaxOwners];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  // warp w scans the chunks [w * per, (w + 1) * per) of every owner; the warp totals are then scanned by thread 0
  const int per = (chunks + 31) / 32;
  const int c0 = warp * per, c1 = min(c0 + per, chunks);
  for (int q = 0; q < world; ++q) {
    int run = 0;
    for (int cb = c0; cb < c1; cb += 32) {
      const int c = cb + lane;
      const int v = c < c1 ? chunk_counts[static_cast<int64_t>(c) * kMaxOwners + q] : 0;
      int incl = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
      }
      if (c < c1) chunk_counts[static_cast<int64_t>(c) * kMaxOwners + q] = run + incl - v;     // exclusive within the warp's range
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_part[warp][q] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int start = 0;
    for (int q = 0; q < world; ++q) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) {
        const int t = s_part[w][q];
        s_part[w][q] = tot;
        tot += t;
      }
      s_base[q] = start;
      counts[q] = tot;
      start += tot;
    }
  }
  __syncthreads();
  for (int q = 0; q < world; ++q)
    for (int c = c0 + lane; c < c1; c += 32) chunk_counts[static_cast<int64_t>(c) * kMaxOwners + q] += s_base[q] + s_part[warp][q];
}

__global__ void __launch_bounds__(kOwnerThreads)
owner_scatter_kernel(IndexMap m, int64_t n, int world, int64_t skip_from, const int32_t* __restrict__ chunk_offsets,
                     int64_t* __restrict__ local_rows, int32_t* __restrict__ perm, int32_t* __restrict__ inv_perm) {
  __shared__ int s_warp[kOwnerThreads / 32][kMaxOwners];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kChunk + threadIdx.x * kItems;
  int64_t row[kItems];
  int owner[kItems];
  int cnt[kMaxOwners];
#pragma unroll
  for (int q = 0; q < kMaxOwners; ++q) cnt[q] = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    owner[i] = -1;
    if (base + i < n) {
      row[i] = row_or_zero(m, base + i);
      owner[i] = row[i] >= skip_from ? -1 : static_cast<int>(row[i] % world);
      if (owner[i] < 0) inv_perm[base + i] = -1;
#pragma unroll
      for (int q = 0; q < kMaxOwners; ++q) cnt[q] += (owner[i] == q) ? 1 : 0;
    }
  }
  // exclusive prefix of the per-thread counts over the CTA, per owner: warp scan, then the warps' totals
  int excl[kMaxOwners];
#pragma unroll
  for (int q = 0; q < kMaxOwners; ++q) {
    int incl = cnt[q];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    excl[q] = incl - cnt[q];
    if (lane == 31) s_warp[warp][q] = incl;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < kMaxOwners; ++q) {
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w][q];
    excl[q] += before + (q < world ? __ldg(chunk_offsets + static_cast<int64_t>(blockIdx.x) * kMaxOwners + q) : 0);
  }
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    if (owner[i] >= 0) {
      int slot = 0;
#pragma unroll
      for (int q = 0; q < kMaxOwners; ++q)
        if (owner[i] == q) slot = excl[q]++;
      local_rows[slot] = row[i] / world;
      perm[slot] = static_cast<int32_t>(base + i);
      inv_perm[base + i] = slot;
    }
  }
}

__global__ void owner_keys_kernel(IndexMap m, int64_t n, int world, uint32_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int64_t row = map_index(m, p);
  if (row < 0) row = 0;  // out-of-range ids are clamped to row 0 (documented in the header)
  keys[p] = static_cast<uint32_t>(row % world);
  vals[p] = static_cast<uint32_t>(p);
}

__global__ void bucket_finish_kernel(IndexMap m, int64_t n, int world, const uint32_t* __restrict__ sorted_owner,
                                     const int32_t* __restrict__ perm, int64_t* __restrict__ local_rows,
                                     int32_t* __restrict__ inv_perm, int64_t* __restrict__ counts) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const int32_t p = perm[i];
    int64_t row = map_index(m, p);
    if (row < 0) row = 0;
    local_rows[i] = row / world;
    inv_perm[p] = static_cast<int32_t>(i);
  }
  // the first `world` threads of the grid count the buckets by binary search in the sorted owners
  if (i < world) {
    const uint32_t g = static_cast<uint32_t>(i);
    int64_t lo = 0, hi = n;
    while (lo < hi) {  // first slot with owner >= g
      const int64_t mid = lo + (hi - lo) / 2;
      if (sorted_owner[mid] < g) lo = mid + 1;
      else hi = mid;
    }
    const int64_t first = lo;
    hi = n;
    while (lo < hi) {  // first slot with owner > g
      const int64_t mid = lo + (hi - lo) / 2;
      if (sorted_owner[mid] <= g) lo = mid + 1;
      else hi = mid;
    }
    counts[g] = lo - first;
  }
}

__global__ void zero_counts_kernel(int64_t* counts, int world) {
  if (threadIdx.x < world) counts[threadIdx.x] = 0;
}

struct BucketWs {
  size_t chunk_counts, keys_in, keys_out, vals_in, cub_temp, cub_bytes, total;
};

static int owner_bits(int world) {
  int bits = 1;
  while ((1 << bits) < world) ++bits;
  return bits;
}

static BucketWs bucket_ws(int64_t n, int world) {
  BucketWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~static_cast<size_t>(255);
    return o;
  };
  const size_t n4 = static_cast<size_t>(n > 0 ? n : 1) * 4;
  w.chunk_counts = take(static_cast<size_t>((n > 0 ? n : 1) / kChunk + 1) * kMaxOwners * sizeof(int32_t));
  w.keys_in = take(n4);
  w.keys_out = take(n4);
  w.vals_in = take(n4);
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<int>(n > 0 ? n : 1), 0, owner_bits(world));
  w.cub_bytes = sort_bytes;
  w.cub_temp = take(sort_bytes + 256);
  w.total = off;
  return w;
}

}  // namespace rb

using namespace rb;

extern "C" size_t rb_bucket_by_owner_workspace_bytes(int64_t n, int32_t world) {
  if (n < 0 || n >= 0x7FFFFFFFll || world < 1 || world > 1024) return 0;
  return bucket_ws(n, world).total;
}

extern "C" int rb_bucket_by_owner_skip(const void* idx, int32_t idx_type, int64_t n, int32_t L, const int64_t* field_row_offset,
                                       int64_t hash_mod, int32_t world, int64_t skip_from_row, int64_t* local_rows_out, int32_t* perm_out,
                                       int32_t* inv_perm_out, int64_t* counts_out, void* ws, size_t ws_bytes, void* stream);

extern "C" int rb_bucket_by_owner(const void* idx, int32_t idx_type, int64_t n, int32_t L, const int64_t* field_row_offset,
                                  int64_t hash_mod, int32_t world, int64_t* local_rows_out, int32_t* perm_out,
                                  int32_t* inv_perm_out, int64_t* counts_out, void* ws, size_t ws_bytes, void* stream) {
  return rb_bucket_by_owner_skip(idx, idx_type, n, L, field_row_offset, hash_mod, world, INT64_MAX, local_rows_out, perm_out, inv_perm_out,
                                 counts_out, ws, ws_bytes, stream);
}

extern "C" int rb_bucket_by_owner_skip(const void* idx, int32_t idx_type, int64_t n, int32_t L, const int64_t* field_row_offset,
                                       int64_t hash_mod, int32_t world, int64_t skip_from_row, int64_t* local_rows_out, int32_t* perm_out,
                                       int32_t* inv_perm_out, int64_t* counts_out, void* ws, size_t ws_bytes, void* stream) {
  RB_CHECK_ARG(skip_from_row == INT64_MAX || world <= kMaxOwners, RB_ERR_ARG, "skip_from_row needs world <= %d (the counting partition)", kMaxOwners);
  RB_CHECK_ARG(n >= 0 && n < 0x7FFFFFFFll && L > 0, RB_ERR_ARG, "n must be in [0, 2^31) and L > 0");
  RB_CHECK_ARG(world >= 1 && world <= 1024, RB_ERR_ARG, "world must be in [1, 1024]");
  RB_CHECK_ARG(idx_type == RB_I32 || idx_type == RB_I64, RB_ERR_ARG, "bad index type");
  RB_CHECK_ARG(counts_out != nullptr, RB_ERR_ARG, "counts_out is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    zero_counts_kernel<<<1, 1024, 0, st>>>(counts_out, world);
    RB_LAUNCH_CHECK("zero_counts_kernel");
    return RB_OK;
  }
  RB_CHECK_ARG(idx != nullptr && local_rows_out != nullptr && perm_out != nullptr && inv_perm_out != nullptr, RB_ERR_ARG,
               "idx or an output is null");
  const BucketWs lay = bucket_ws(n, world);
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= lay.total, RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu",
               lay.total, ws_bytes);
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0, RB_ERR_ALIGN, "workspace must be 256 B aligned");
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  uint32_t* keys_in = reinterpret_cast<uint32_t*>(wsb + lay.keys_in);
  uint32_t* keys_out = reinterpret_cast<uint32_t*>(wsb + lay.keys_out);
  uint32_t* vals_in = reinterpret_cast<uint32_t*>(wsb + lay.vals_in);
  // the row count only bounds the range check inside map_index; the caller guarantees ids in range
  IndexMap m = make_index_map(idx, idx_type, field_row_offset, hash_mod, INT64_MAX, L);
  if (world <= kMaxOwners) {      // the box's 8 GPUs: counting partition, no sort
    int32_t* chunk_counts = reinterpret_cast<int32_t*>(wsb + lay.chunk_counts);
    const int chunks = static_cast<int>((n + kChunk - 1) / kChunk);
    owner_count_kernel<<<chunks, kOwnerThreads, 0, st>>>(m, n, world, skip_from_row, chunk_counts);
    RB_LAUNCH_CHECK("owner_count_kernel");
    owner_offsets_kernel<<<1, 1024, 0, st>>>(chunk_counts, chunks, world, counts_out);
    RB_LAUNCH_CHECK("owner_offsets_kernel");
    owner_scatter_kernel<<<chunks, kOwnerThreads, 0, st>>>(m, n, world, skip_from_row, chunk_counts, local_rows_out, perm_out, inv_perm_out);
    RB_LAUNCH_CHECK("owner_scatter_kernel");
    return RB_OK;
  }
  owner_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(m, n, world, keys_in, vals_in);
  RB_LAUNCH_CHECK("owner_keys_kernel");
  size_t temp = lay.cub_bytes;
  RB_CUDA(cub::DeviceRadixSort::SortPairs(wsb + lay.cub_temp, temp, keys_in, keys_out, vals_in,
                                          reinterpret_cast<uint32_t*>(perm_out), static_cast<int>(n), 0, owner_bits(world), st));
  const int64_t work = n > world ? n : world;
  bucket_finish_kernel<<<grid_for(work, 256), 256, 0, st>>>(m, n, world, keys_out, perm_out, local_rows_out, inv_perm_out,
                                                            counts_out);
  RB_LAUNCH_CHECK("bucket_finish_kernel");
  return RB_OK;
}
