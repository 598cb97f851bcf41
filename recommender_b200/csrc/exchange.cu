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
  size_t keys_in, keys_out, vals_in, cub_temp, cub_bytes, total;
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

extern "C" int rb_bucket_by_owner(const void* idx, int32_t idx_type, int64_t n, int32_t L, const int64_t* field_row_offset,
                                  int64_t hash_mod, int32_t world, int64_t* local_rows_out, int32_t* perm_out,
                                  int32_t* inv_perm_out, int64_t* counts_out, void* ws, size_t ws_bytes, void* stream) {
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
