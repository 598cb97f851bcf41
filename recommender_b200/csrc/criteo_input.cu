// Input side of the CTR path: Criteo TSV bytes resident in HBM -> the batch the models consume.
// Replaces the per-line Python of ctr/tfrecord_io.py (build_vocab :15-35, write_tfrecord :38-75, read_tfrecord
// :78-96) — the reference is input-bound (≈ 3 k examples/s from its Python generator).
//
//   rb_criteo_index_lines   newline positions -> line starts              (2 passes over the bytes, 16 B per thread)
//   rb_criteo_parse         one warp per line: stage the line in shared memory, find the 39 tabs with ballots, then
//                           lane f parses column f (and f + 32): label, log(max(x,0)+1), token key -> vocabulary id
//   rb_vocab_build          the global dictionary: stable sort of (token key, position), run lengths, count > 10,
//                           ids in first-seen order (a second sort by first position) — no host round trip
//   rb_vocab_table_build / rb_vocab_lookup   open-addressing table key -> id in HBM; OOV -> 0
//
// Byte / integer work, HBM- and latency-bound: no tensor cores here.
#include <cub/cub.cuh>

#include "common.cuh"
#include "criteo_fields.h"

namespace rb {
namespace {

using namespace criteo;

constexpr int kTileThreads = 256;
constexpr int kTileBytes = kTileThreads * 16;  // one 128-bit load per thread
constexpr int kMaxLine = 1024;                 // bytes of one line without its newline (Criteo lines are < 450)
constexpr int kLineBuf = kMaxLine + 32;        // the copy starts at the 16-byte boundary below the line; 16 B of slack
                                               // behind it for load8 (criteo_fields.h)
constexpr int kParseWarps = 8;

__device__ __forceinline__ unsigned newline_mask(uint32_t w) { return __vcmpeq4(w, 0x0A0A0A0Au) & 0x01010101u; }

// 16-bit mask of the tab bytes of v (bit j = byte j).  Per word: 0xFF where the byte matches, one bit per byte kept,
// and a multiply gathers bits 0, 8, 16, 24 into one nibble (no two partial products share a bit, so no carries).
__device__ __forceinline__ uint32_t tab_mask16(const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) m |= (((__vcmpeq4(w[k], 0x09090909u) & 0x01010101u) * 0x01020408u) >> 24) << (4 * k);
  return m;
}

// newline bytes among the first `valid` (1..16) bytes of v
__device__ __forceinline__ int count_newlines(const uint4& v, int valid) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  int c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned m = newline_mask(w[k]);
    const int vb = valid - 4 * k;  // valid bytes of this word
    if (vb <= 0) m = 0;
    else if (vb < 4) m &= (1u << (8 * vb)) - 1u;
    c += __popc(m);
  }
  return c;
}

// A newline at position p separates two lines iff p + 1 < nbytes: only bytes [0, nbytes - 1) are examined, so a final
// newline does not open an empty last line (`for line in f`, ctr/tfrecord_io.py:18).
__global__ void __launch_bounds__(kTileThreads) count_newlines_kernel(const uint8_t* __restrict__ text, int64_t nbytes,
                                                                      int32_t* __restrict__ tile_counts) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kTileBytes + threadIdx.x * 16;
  const int64_t limit = nbytes - 1;
  int c = 0;
  if (base < limit) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + base));
    c = count_newlines(v, static_cast<int>(min(static_cast<int64_t>(16), limit - base)));
  }
  using Reduce = cub::BlockReduce<int, kTileThreads>;
  __shared__ typename Reduce::TempStorage tmp;
  const int total = Reduce(tmp).Sum(c);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kTileThreads) write_line_starts_kernel(const uint8_t* __restrict__ text, int64_t nbytes,
                                                                         const int32_t* __restrict__ tile_offsets,
                                                                         int64_t max_lines, int64_t* __restrict__ line_start,
                                                                         int64_t* __restrict__ num_lines) {
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kTileBytes + threadIdx.x * 16;
  const int64_t limit = nbytes - 1;
  uint4 v = make_uint4(0, 0, 0, 0);
  int valid = 0, c = 0;
  if (base < limit) {
    v = __ldg(reinterpret_cast<const uint4*>(text + base));
    valid = static_cast<int>(min(static_cast<int64_t>(16), limit - base));
    c = count_newlines(v, valid);
  }
  using Scan = cub::BlockScan<int, kTileThreads>;
  __shared__ typename Scan::TempStorage tmp;
  int before = 0, total = 0;
  Scan(tmp).ExclusiveSum(c, before, total);
  const int64_t tile_off = tile_offsets[blockIdx.x];
  if (c > 0) {
    int64_t out = 1 + tile_off + before;  // line 0 starts at byte 0
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < valid && ((w[k >> 2] >> (8 * (k & 3))) & 0xFFu) == 0x0Au) {
        if (out < max_lines) line_start[out] = base + k + 1;
        ++out;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && max_lines > 0) line_start[0] = 0;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *num_lines = 1 + tile_off + total;
}

struct ParseArgs {
  const uint8_t* text;
  int64_t nbytes;
  const int64_t* line_start;
  int64_t num_lines;
  int64_t* label;
  float* int_features;
  uint64_t* cat_tokens;   // may be null
  int64_t* cat_features;  // may be null
  const uint64_t* vocab_keys;
  const int32_t* vocab_vals;
  uint64_t vocab_mask;
  int32_t* error_flag;
};

__global__ void __launch_bounds__(kParseWarps * 32) parse_lines_kernel(ParseArgs a) {
  __shared__ __align__(16) uint8_t s_line[kParseWarps][kLineBuf];
  __shared__ int16_t s_tab[kParseWarps][kCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* buf = s_line[warp];
  int16_t* tab = s_tab[warp];
  int err = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kParseWarps + warp; i < a.num_lines;
       i += static_cast<int64_t>(gridDim.x) * kParseWarps) {
    const int64_t beg = __ldg(a.line_start + i);
    int64_t end;
    bool has_nl = true;
    if (i + 1 < a.num_lines) {
      end = __ldg(a.line_start + i + 1) - 1;        // the separating newline
    } else if (a.text[a.nbytes - 1] == '\n') {
      end = a.nbytes - 1;
    } else {
      end = a.nbytes;                               // a last line without a newline: nothing stays attached to C26
      has_nl = false;
    }
    const int64_t len64 = end - beg;
    const bool too_long = len64 > kMaxLine;
    const int len = too_long ? 0 : static_cast<int>(len64);
    if (too_long) err |= kErrLongLine;
    // ---- stage the line: 128-bit loads from the 16-byte boundary below `beg`
    const int64_t abeg = beg & ~static_cast<int64_t>(15);
    const int shift = static_cast<int>(beg - abeg);
    const int n16 = (shift + len + 15) >> 4;
    __syncwarp();                                   // the previous line's readers are done with buf / tab
    // Each lane moves 16 bytes to shared memory and, from the same registers, finds the tabs among them; a warp
    // prefix sum numbers them in line order.  Tab k closes column k.
    const uint8_t* line = buf + shift;
    int ntabs = 0;
    for (int k0 = 0; k0 < n16; k0 += 32) {
      const int k = k0 + lane;
      uint32_t tm = 0;                              // bit j: byte j of this lane's 16 is a tab inside the line
      if (k < n16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.text + abeg) + k);
        reinterpret_cast<uint4*>(buf)[k] = v;
        tm = tab_mask16(v);
        const int lo = shift - 16 * k;              // bytes before the line (first 16 only) ...
        const int hi = shift + len - 16 * k;        // ... and behind it are not part of it
        if (lo > 0) tm &= ~((1u << lo) - 1u);
        if (hi < 16) tm &= (1u << (hi > 0 ? hi : 0)) - 1u;
      }
      const int c = __popc(tm);
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
      }
      int idx = ntabs + incl - c;
      while (tm != 0) {
        const int j = __ffs(tm) - 1;
        tm &= tm - 1;
        if (idx < kCols) tab[idx] = static_cast<int16_t>(16 * k + j - shift);
        ++idx;
      }
      ntabs += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    __syncwarp();
    const bool short_line = ntabs < kCols - 1;      // line.split('\t')[39] raises IndexError in the reference
    if (short_line && !too_long) err |= kErrShortLine;
    const bool ok = !short_line && !too_long;
    for (int f = lane; f < kCols; f += 32) {
      int start = 0, stop = 0;
      if (ok) column_span(tab, ntabs, len, f, &start, &stop);
      const uint8_t* s = line + start;
      const int flen = stop - start;
      if (f == 0) {
        int64_t v = 0;
        if (ok && !parse_int(s, flen, &v)) err |= kErrBadInt;
        a.label[i] = v;
      } else if (f <= kNumInt) {
        int64_t v = 0;
        if (ok && !int_column(s, flen, &v)) err |= kErrBadInt;
        // :51-53  int64 -> float32 (round to nearest even), + 1 and log in float32
        a.int_features[i * kNumInt + (f - 1)] = logf(__ll2float_rn(v) + 1.0f);
      } else {
        const int field = f - kNumInt - 1;
        uint64_t key = missing_key(field);
        if (ok) key = token_key(s, flen, f == kCols - 1 && last_column_keeps_newline(ntabs, has_nl), field, &err);
        if (a.cat_tokens != nullptr) a.cat_tokens[i * kNumCat + field] = key;
        if (a.cat_features != nullptr)
          a.cat_features[i * kNumCat + field] = ok ? vocab_find(a.vocab_keys, a.vocab_vals, a.vocab_mask, key) : 0;
      }
    }
  }
  err = __reduce_or_sync(0xFFFFFFFFu, err);
  if (lane == 0 && err != 0 && a.error_flag != nullptr) atomicOr(a.error_flag, err);
}

// ---- vocabulary -------------------------------------------------------------------------------------------------

__global__ void iota_kernel(uint32_t* __restrict__ v, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) v[i] = static_cast<uint32_t>(i);
}

// head index of position i's run, as the running maximum of (i is a run head ? i : 0)
__global__ void head_index_kernel(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ h) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) h[i] = (i == 0 || keys[i] != keys[i - 1]) ? static_cast<uint32_t>(i) : 0u;
}

// first[i] = first position of the token if i heads a run longer than min_count (ctr/tfrecord_io.py:31), else
// 0xFFFFFFFF.  The pair sort was stable, so the run head carries the smallest position = the first occurrence.
__global__ void first_seen_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pos,
                                  const uint32_t* __restrict__ head_of, int64_t n, int32_t min_count,
                                  uint32_t* __restrict__ first, unsigned long long* __restrict__ num_vocab) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool head = (i == 0 || keys[i] != keys[i - 1]);
  const bool tail = (i == n - 1 || keys[i] != keys[i + 1]);
  if (!head) first[i] = 0xFFFFFFFFu;
  if (tail) {
    const uint32_t h = head_of[i];
    const int64_t count = i - h + 1;
    const bool keep = count > min_count;
    first[h] = keep ? pos[h] : 0xFFFFFFFFu;
    if (keep) atomicAdd(num_vocab, 1ull);
  }
}

__global__ void emit_vocab_kernel(const uint64_t* __restrict__ by_first, const unsigned long long* __restrict__ num_vocab,
                                  int64_t max_vocab, uint64_t* __restrict__ out, int64_t* __restrict__ num_out) {
  const int64_t nv = static_cast<int64_t>(*num_vocab);
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0) *num_out = nv;
  if (i < nv && i < max_vocab) out[i] = by_first[i];
}

__global__ void table_insert_kernel(const uint64_t* __restrict__ vocab_keys, int64_t n, unsigned long long* __restrict__ tkeys,
                                    int32_t* __restrict__ tvals, uint64_t mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = vocab_keys[i];
  uint64_t slot = mix64(key) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(tkeys + slot, static_cast<unsigned long long>(kEmptySlot), key);
    if (prev == kEmptySlot || prev == key) {
      tvals[slot] = static_cast<int32_t>(i);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__global__ void table_lookup_kernel(const uint64_t* __restrict__ tokens, int64_t n, const uint64_t* __restrict__ tkeys,
                                    const int32_t* __restrict__ tvals, uint64_t mask, int64_t* __restrict__ ids) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = vocab_find(tkeys, tvals, mask, tokens[i]);
}

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct IndexWs {
  size_t counts, offsets, cub_temp, cub_bytes, total;
};
IndexWs index_ws(int64_t nbytes) {
  IndexWs w{};
  const size_t tiles = static_cast<size_t>((nbytes + kTileBytes - 1) / kTileBytes) + 1;
  size_t off = 0;
  w.counts = off;
  off += align256(tiles * 4);
  w.offsets = off;
  off += align256(tiles * 4);
  cub::DeviceScan::ExclusiveSum(nullptr, w.cub_bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                static_cast<int>(tiles));
  w.cub_temp = off;
  off += align256(w.cub_bytes + 256);
  w.total = off;
  return w;
}

struct VocabWs {
  size_t keys_a, keys_b, vals_a, vals_b, head, first_b, count, cub_temp, cub_bytes, total;
};
VocabWs vocab_ws(int64_t n) {
  VocabWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += align256(bytes);
    return o;
  };
  const size_t un = static_cast<size_t>(n);
  w.keys_a = take(un * 8);
  w.keys_b = take(un * 8);
  w.vals_a = take(un * 4);   // positions, later the first-seen keys of the second sort
  w.vals_b = take(un * 4);
  w.head = take(un * 4);
  w.first_b = take(un * 4);
  w.count = take(256);
  size_t s1 = 0, s2 = 0, s3 = 0;
  cub::DoubleBuffer<uint64_t> k64(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> v32(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, s1, k64, v32, static_cast<int>(n), 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, s2, v32, k64, static_cast<int>(n), 0, 32);
  cub::DeviceScan::InclusiveScan(nullptr, s3, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), cub::Max(),
                                 static_cast<int>(n));
  w.cub_bytes = s1 > s2 ? s1 : s2;
  if (s3 > w.cub_bytes) w.cub_bytes = s3;
  w.cub_temp = take(w.cub_bytes + 256);
  w.total = off;
  return w;
}

int grid1d(int64_t n, int threads) { return static_cast<int>((n + threads - 1) / threads); }

bool pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }

}  // namespace
}  // namespace rb

using namespace rb;

extern "C" size_t rb_criteo_index_workspace_bytes(int64_t nbytes) {
  if (nbytes < 0 || nbytes >= 0x7FFFFFFFll) return 0;
  return index_ws(nbytes > 0 ? nbytes : 1).total;
}

extern "C" int rb_criteo_index_lines(const uint8_t* text, int64_t nbytes, int64_t max_lines, int64_t* line_start,
                                     int64_t* num_lines_dev, void* ws, size_t ws_bytes, void* stream) {
  RB_CHECK_ARG(nbytes >= 0 && nbytes < 0x7FFFFFFFll, RB_ERR_ARG, "nbytes must be in [0, 2^31): feed the file in chunks");
  RB_CHECK_ARG(num_lines_dev != nullptr && max_lines >= 0 && (max_lines == 0 || line_start != nullptr), RB_ERR_ARG,
               "num_lines_dev / line_start is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nbytes == 0) {
    RB_CUDA(cudaMemsetAsync(num_lines_dev, 0, sizeof(int64_t), st));
    return RB_OK;
  }
  RB_CHECK_ARG(text != nullptr && (reinterpret_cast<uintptr_t>(text) & 15) == 0, RB_ERR_ALIGN, "text must be 16 B aligned");
  const IndexWs w = index_ws(nbytes);
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= w.total, RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
               ws_bytes);
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0, RB_ERR_ALIGN, "workspace must be 256 B aligned");
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  int32_t* counts = reinterpret_cast<int32_t*>(wsb + w.counts);
  int32_t* offsets = reinterpret_cast<int32_t*>(wsb + w.offsets);
  const int tiles = static_cast<int>((nbytes + kTileBytes - 1) / kTileBytes);
  count_newlines_kernel<<<tiles, kTileThreads, 0, st>>>(text, nbytes, counts);
  RB_LAUNCH_CHECK("count_newlines_kernel");
  size_t temp = w.cub_bytes;
  RB_CUDA(cub::DeviceScan::ExclusiveSum(wsb + w.cub_temp, temp, counts, offsets, tiles, st));
  write_line_starts_kernel<<<tiles, kTileThreads, 0, st>>>(text, nbytes, offsets, max_lines, line_start, num_lines_dev);
  RB_LAUNCH_CHECK("write_line_starts_kernel");
  return RB_OK;
}

extern "C" int rb_criteo_parse(const uint8_t* text, int64_t nbytes, const int64_t* line_start, int64_t num_lines,
                               int64_t* label, float* int_features, uint64_t* cat_tokens, int64_t* cat_features,
                               const uint64_t* vocab_table_keys, const int32_t* vocab_table_vals, int64_t vocab_capacity,
                               int32_t* error_flag, void* stream) {
  RB_CHECK_ARG(num_lines >= 0 && nbytes >= 0 && nbytes < 0x7FFFFFFFll, RB_ERR_ARG, "bad num_lines / nbytes");
  if (num_lines == 0) return RB_OK;
  RB_CHECK_ARG(nbytes > 0 && text != nullptr && line_start != nullptr && label != nullptr && int_features != nullptr, RB_ERR_ARG,
               "text / line_start / label / int_features is null");
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(text) & 15) == 0, RB_ERR_ALIGN, "text must be 16 B aligned");
  RB_CHECK_ARG(cat_tokens != nullptr || cat_features != nullptr, RB_ERR_ARG, "ask for cat_tokens, cat_features or both");
  RB_CHECK_ARG(cat_features == nullptr || (vocab_table_keys != nullptr && vocab_table_vals != nullptr && pow2(vocab_capacity)),
               RB_ERR_ARG, "cat_features needs a vocabulary table whose capacity is a power of two");
  ParseArgs a;
  a.text = text;
  a.nbytes = nbytes;
  a.line_start = line_start;
  a.num_lines = num_lines;
  a.label = label;
  a.int_features = int_features;
  a.cat_tokens = cat_tokens;
  a.cat_features = cat_features;
  a.vocab_keys = vocab_table_keys;
  a.vocab_vals = vocab_table_vals;
  a.vocab_mask = cat_features != nullptr ? static_cast<uint64_t>(vocab_capacity - 1) : 0;
  a.error_flag = error_flag;
  const int64_t blocks = (num_lines + kParseWarps - 1) / kParseWarps;
  const int grid = static_cast<int>(blocks < 16 * kNumSMs ? blocks : 16 * kNumSMs);
  parse_lines_kernel<<<grid, kParseWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  RB_LAUNCH_CHECK("parse_lines_kernel");
  return RB_OK;
}

extern "C" size_t rb_vocab_build_workspace_bytes(int64_t n) {
  if (n < 0 || n >= 0x7FFFFFFFll) return 0;
  return vocab_ws(n > 0 ? n : 1).total;
}

extern "C" int rb_vocab_build(const uint64_t* tokens, int64_t n, int32_t min_count, uint64_t* vocab_keys_out, int64_t max_vocab,
                              int64_t* num_vocab_dev, void* ws, size_t ws_bytes, void* stream) {
  RB_CHECK_ARG(n >= 0 && n < 0x7FFFFFFFll, RB_ERR_ARG, "the number of tokens must be in [0, 2^31)");
  RB_CHECK_ARG(num_vocab_dev != nullptr && max_vocab >= 0 && (max_vocab == 0 || vocab_keys_out != nullptr) && min_count >= 0,
               RB_ERR_ARG, "num_vocab_dev / vocab_keys_out null or negative size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    RB_CUDA(cudaMemsetAsync(num_vocab_dev, 0, sizeof(int64_t), st));
    return RB_OK;
  }
  RB_CHECK_ARG(tokens != nullptr, RB_ERR_ARG, "tokens is null");
  const VocabWs w = vocab_ws(n);
  RB_CHECK_ARG(ws != nullptr && ws_bytes >= w.total, RB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total,
               ws_bytes);
  RB_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0, RB_ERR_ALIGN, "workspace must be 256 B aligned");
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  uint64_t* ka = reinterpret_cast<uint64_t*>(wsb + w.keys_a);
  uint64_t* kb = reinterpret_cast<uint64_t*>(wsb + w.keys_b);
  uint32_t* va = reinterpret_cast<uint32_t*>(wsb + w.vals_a);
  uint32_t* vb = reinterpret_cast<uint32_t*>(wsb + w.vals_b);
  uint32_t* head = reinterpret_cast<uint32_t*>(wsb + w.head);
  uint32_t* first_b = reinterpret_cast<uint32_t*>(wsb + w.first_b);
  unsigned long long* count = reinterpret_cast<unsigned long long*>(wsb + w.count);
  const int ni = static_cast<int>(n);
  const int grid = grid1d(n, 256);
  RB_CUDA(cudaMemcpyAsync(ka, tokens, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToDevice, st));
  RB_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
  iota_kernel<<<grid, 256, 0, st>>>(va, n);
  RB_LAUNCH_CHECK("iota_kernel");
  // 1. stable sort of (token key, scan position): equal tokens become one run, first occurrence at its head
  cub::DoubleBuffer<uint64_t> dk(ka, kb);
  cub::DoubleBuffer<uint32_t> dv(va, vb);
  size_t temp = w.cub_bytes;
  RB_CUDA(cub::DeviceRadixSort::SortPairs(wsb + w.cub_temp, temp, dk, dv, ni, 0, 64, st));
  const uint64_t* keys = dk.Current();
  const uint32_t* pos = dv.Current();
  uint32_t* first_a = dv.Alternate();   // the positions' other half is free now
  // 2. run lengths: every element learns its run's head index (max-scan), the tail knows the length
  head_index_kernel<<<grid, 256, 0, st>>>(keys, n, head);
  RB_LAUNCH_CHECK("head_index_kernel");
  temp = w.cub_bytes;
  RB_CUDA(cub::DeviceScan::InclusiveScan(wsb + w.cub_temp, temp, head, head, cub::Max(), ni, st));
  first_seen_kernel<<<grid, 256, 0, st>>>(keys, pos, head, n, min_count, first_a, count);
  RB_LAUNCH_CHECK("first_seen_kernel");
  // 3. ids in first-seen order: sort the kept tokens by their first position (dropped ones carry 0xFFFFFFFF)
  cub::DoubleBuffer<uint32_t> df(first_a, first_b);
  cub::DoubleBuffer<uint64_t> dt(const_cast<uint64_t*>(keys), dk.Alternate());
  temp = w.cub_bytes;
  RB_CUDA(cub::DeviceRadixSort::SortPairs(wsb + w.cub_temp, temp, df, dt, ni, 0, 32, st));
  const int64_t m = max_vocab < n ? max_vocab : n;
  emit_vocab_kernel<<<grid1d(m > 0 ? m : 1, 256), 256, 0, st>>>(dt.Current(), count, max_vocab, vocab_keys_out, num_vocab_dev);
  RB_LAUNCH_CHECK("emit_vocab_kernel");
  return RB_OK;
}

extern "C" int rb_vocab_table_build(const uint64_t* vocab_keys, int64_t num_vocab, uint64_t* table_keys, int32_t* table_vals,
                                    int64_t capacity, void* stream) {
  RB_CHECK_ARG(num_vocab >= 0 && num_vocab < 0x7FFFFFFFll, RB_ERR_ARG, "num_vocab must be in [0, 2^31)");
  RB_CHECK_ARG(pow2(capacity) && capacity > num_vocab, RB_ERR_ARG, "capacity must be a power of two greater than num_vocab");
  RB_CHECK_ARG(table_keys != nullptr && table_vals != nullptr && (num_vocab == 0 || vocab_keys != nullptr), RB_ERR_ARG,
               "table_keys / table_vals / vocab_keys is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RB_CUDA(cudaMemsetAsync(table_keys, 0xFF, static_cast<size_t>(capacity) * 8, st));
  RB_CUDA(cudaMemsetAsync(table_vals, 0, static_cast<size_t>(capacity) * 4, st));
  if (num_vocab > 0) {
    table_insert_kernel<<<grid1d(num_vocab, 256), 256, 0, st>>>(vocab_keys, num_vocab,
                                                                reinterpret_cast<unsigned long long*>(table_keys), table_vals,
                                                                static_cast<uint64_t>(capacity - 1));
    RB_LAUNCH_CHECK("table_insert_kernel");
  }
  return RB_OK;
}

extern "C" int rb_vocab_lookup(const uint64_t* tokens, int64_t n, const uint64_t* table_keys, const int32_t* table_vals,
                               int64_t capacity, int64_t* ids_out, void* stream) {
  RB_CHECK_ARG(n >= 0, RB_ERR_ARG, "negative n");
  if (n == 0) return RB_OK;
  RB_CHECK_ARG(tokens != nullptr && ids_out != nullptr && table_keys != nullptr && table_vals != nullptr && pow2(capacity), RB_ERR_ARG,
               "null pointer or capacity not a power of two");
  table_lookup_kernel<<<grid1d(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(tokens, n, table_keys, table_vals,
                                                                                     static_cast<uint64_t>(capacity - 1), ids_out);
  RB_LAUNCH_CHECK("table_lookup_kernel");
  return RB_OK;
}
