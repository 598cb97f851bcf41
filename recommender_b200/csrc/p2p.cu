// Peer-memory plumbing for the sharded path: buffers that the other GPUs of the box
// read directly over NVLink/NVSwitch — the table shards (rows are gathered by the interaction
// kernel of ANY rank straight from the owner's HBM), the per-rank gradient tensor dE (pulled by
// the owner's segmented reduction) and the small bucket arrays that route lookups to owners.
// One process per GPU: buffers are cudaMalloc'ed here, exported with cudaIpcGetMemHandle and
// opened in the peer processes; torch.distributed only carries the 64-byte handles.
#include <string.h>

#include "common.cuh"

namespace rb {

// Owner q collects, from every source rank k, the slice of k's bucket arrays addressed to q
// (lookups whose row mod world == q), as (key = local row, val = k * n_local + position) pairs
// padded with `invalid_key` up to `capacity`.  All inputs are read over peer pointers.
struct CollectArgs {
  int world, me;
  int capacity;
  uint32_t invalid_key;
  uint32_t n_local;                       // positions per source rank (B_local * F)
  const int64_t* rows[RB_MAX_RANKS];      // peer k: local row ids, bucket order       int64[n_local]
  const int32_t* perm[RB_MAX_RANKS];      // peer k: bucket slot -> lookup position     int32[n_local]
  const int64_t* counts[RB_MAX_RANKS];    // peer k: lookups per owner                  int64[world]
};

__global__ void __launch_bounds__(256)
collect_keys_kernel(const __grid_constant__ CollectArgs a, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                    int32_t* __restrict__ n_valid, int32_t* __restrict__ overflow_flag) {
  __shared__ int64_t s_src_off[RB_MAX_RANKS];   // where my slice starts inside peer k's bucket arrays
  __shared__ int64_t s_dst_off[RB_MAX_RANKS + 1];
  if (threadIdx.x == 0) {
    int64_t dst = 0;
    for (int k = 0; k < a.world; ++k) {
      int64_t off = 0;
      for (int j = 0; j < a.me; ++j) off += a.counts[k][j];
      s_src_off[k] = off;
      s_dst_off[k] = dst;
      dst += a.counts[k][a.me];
    }
    s_dst_off[a.world] = dst;
    if (blockIdx.x == 0) {
      *n_valid = static_cast<int32_t>(dst < a.capacity ? dst : a.capacity);
      if (dst > a.capacity && overflow_flag != nullptr) *overflow_flag = 1;
    }
  }
  __syncthreads();
  const int64_t total = s_dst_off[a.world];
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.capacity;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint32_t key = a.invalid_key, val = 0;
    if (i < total) {
      int k = 0;
      while (k + 1 < a.world && i >= s_dst_off[k + 1]) ++k;
      const int64_t s = s_src_off[k] + (i - s_dst_off[k]);
      key = static_cast<uint32_t>(a.rows[k][s]);
      val = static_cast<uint32_t>(k) * a.n_local + static_cast<uint32_t>(a.perm[k][s]);
    }
    keys[i] = key;
    vals[i] = val;
  }
}

}  // namespace rb

using namespace rb;

extern "C" int rb_shared_alloc(size_t bytes, void** ptr_out, void* handle_out) {
  RB_CHECK_ARG(ptr_out != nullptr && handle_out != nullptr && bytes > 0, RB_ERR_ARG, "bad arguments");
  void* p = nullptr;
  RB_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaIpcGetMemHandle");
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == RB_IPC_HANDLE_BYTES, "handle size");
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return RB_OK;
}

extern "C" int rb_shared_free(void* ptr) {
  if (ptr != nullptr) RB_CUDA(cudaFree(ptr));
  return RB_OK;
}

extern "C" int rb_ipc_open(const void* handle, void** ptr_out) {
  RB_CHECK_ARG(handle != nullptr && ptr_out != nullptr, RB_ERR_ARG, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  RB_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return RB_OK;
}

extern "C" int rb_enable_peer_access(int32_t peer_device) {
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return RB_OK;
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
  return RB_OK;
}

extern "C" int rb_ipc_close(void* ptr) {
  if (ptr != nullptr) RB_CUDA(cudaIpcCloseMemHandle(ptr));
  return RB_OK;
}

extern "C" int rb_p2p_collect_keys(int32_t world, int32_t me, int64_t n_local, const void* const* rows_ptrs,
                                   const void* const* perm_ptrs, const void* const* counts_ptrs, int64_t local_rows,
                                   int64_t capacity, void* ws, size_t ws_bytes, int32_t D, int32_t* n_valid_dev,
                                   int32_t* overflow_flag, void* stream) {
  RB_CHECK_ARG(world >= 1 && world <= RB_MAX_RANKS && me >= 0 && me < world, RB_ERR_ARG, "world must be in [1, %d]", RB_MAX_RANKS);
  RB_CHECK_ARG(n_local > 0 && n_local * world < 0xFFFFFFFFll && capacity > 0 && capacity < 0x7FFFFFFFll, RB_ERR_ARG,
               "n_local * world must be below 2^32 and capacity in (0, 2^31)");
  RB_CHECK_ARG(rows_ptrs != nullptr && perm_ptrs != nullptr && counts_ptrs != nullptr && n_valid_dev != nullptr, RB_ERR_ARG,
               "null pointer");
  RB_CHECK_ARG(local_rows > 0 && local_rows < 0x7FFFFFFFll, RB_ERR_ARG, "local_rows must be in (0, 2^31)");
  const size_t need = rb_sparse_bwd_update_workspace_bytes(capacity, D, local_rows + 1);
  RB_CHECK_ARG(ws != nullptr && need > 0 && ws_bytes >= need, RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need,
               ws_bytes);
  CollectArgs a;
  a.world = world;
  a.me = me;
  a.capacity = static_cast<int>(capacity);
  a.invalid_key = static_cast<uint32_t>(local_rows);   // one past the largest real key: sorts behind every real pair
  a.n_local = static_cast<uint32_t>(n_local);
  for (int k = 0; k < world; ++k) {
    RB_CHECK_ARG(rows_ptrs[k] != nullptr && perm_ptrs[k] != nullptr && counts_ptrs[k] != nullptr, RB_ERR_ARG, "rank %d: null pointer", k);
    a.rows[k] = static_cast<const int64_t*>(rows_ptrs[k]);
    a.perm[k] = static_cast<const int32_t*>(perm_ptrs[k]);
    a.counts[k] = static_cast<const int64_t*>(counts_ptrs[k]);
  }
  uint32_t *keys, *vals;
  int rc = sparse_ws_key_buffers(capacity, D, local_rows + 1, ws, &keys, &vals);
  if (rc != RB_OK) return rc;
  const unsigned int grid = static_cast<unsigned int>((capacity + 255) / 256 < 4 * kNumSMs ? (capacity + 255) / 256 : 4 * kNumSMs);
  collect_keys_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, keys, vals, n_valid_dev, overflow_flag);
  RB_LAUNCH_CHECK("collect_keys_kernel");
  return RB_OK;
}
