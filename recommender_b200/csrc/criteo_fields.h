// Field-level pieces of the Criteo TSV parser (ctr/tfrecord_io.py:15-66), written so that the SAME code compiles
// for the device (criteo_input.cu) and for the host (tests/test_criteo_fields_cpu.py builds it with g++ and checks it
// against the oracle without a GPU).  No CUDA types in here.
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

namespace rb {
namespace criteo {

constexpr int kNumInt = 13;   // ctr/tfrecord_io.py:8
constexpr int kNumCat = 26;   // :9
constexpr int kCols = 40;     // :10  label + 13 + 26
constexpr uint64_t kNewlineBit = 1ull << 63;
constexpr uint64_t kEmptySlot = ~0ull;   // not a token key: a packed ASCII token never has bits 7, 15, ... set

// bits of *error_flag (documented in include/recsys_b200.h)
constexpr int kErrShortLine = 1;   // fewer than 40 columns (the reference raises IndexError, :21/:45)
constexpr int kErrBadInt = 2;      // a label / integer column that int() would reject, or beyond 18 digits
constexpr int kErrLongToken = 4;   // a categorical token longer than 8 bytes has no 64-bit key
constexpr int kErrNonAscii = 8;    // a byte >= 0x80 inside a categorical token
constexpr int kErrLongLine = 16;   // a line longer than the parser's staging buffer

// The 8 bytes at s (any alignment), little-endian: byte i of the result is s[i].  Bytes behind the column are garbage
// the caller masks off; the caller guarantees that 11 bytes behind s are readable (the parser's line buffer has slack).
RB_HD uint64_t load8(const uint8_t* s) {
#if defined(__CUDA_ARCH__)
  const uintptr_t a = reinterpret_cast<uintptr_t>(s);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const unsigned sh = static_cast<unsigned>(a & 3) * 8;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  return (static_cast<uint64_t>(__funnelshift_r(w1, w2, sh)) << 32) | __funnelshift_r(w0, w1, sh);
#else
  uint64_t v;
  memcpy(&v, s, 8);  // little-endian hosts (x86-64, aarch64)
  return v;
#endif
}

// 1..8 decimal digits held in the low `len` bytes of `chunk` (first character in byte 0), all at once: pad with
// leading '0's, check every byte is a digit, then three multiply-shift steps combine pairs, quads and the two halves.
RB_HD bool parse_digits8(uint64_t chunk, int len, int64_t* out) {
  const int pad = 8 * (8 - len);
  uint64_t v = pad ? ((chunk << pad) | (0x3030303030303030ull >> (64 - pad))) : chunk;
  if (((v & 0xF0F0F0F0F0F0F0F0ull) | (((v + 0x0606060606060606ull) & 0xF0F0F0F0F0F0F0F0ull) >> 4)) != 0x3333333333333333ull)
    return false;
  v = ((v & 0x0F0F0F0F0F0F0F0Full) * 2561) >> 8;
  v = ((v & 0x00FF00FF00FF00FFull) * 6553601) >> 16;
  *out = static_cast<int64_t>(((v & 0x0000FFFF0000FFFFull) * 42949672960001ull) >> 32);
  return true;
}

// `int(s)` for the label and the 13 integer columns: optional sign, decimal digits.  '' is the caller's business
// (:46-47 turns it into '0' for the integer columns; the label is never empty).  Returns false when int() would raise
// (or the value does not fit: more than 18 digits).  Up to 8 digits go through parse_digits8; longer ones and
// anything odd through the byte loop.
RB_HD bool parse_int(const uint8_t* s, int len, int64_t* out) {
  if (len >= 1 && len <= 8) {
    uint64_t c = load8(s);
    int l = len;
    const unsigned first = static_cast<unsigned>(c & 0xFF);
    const bool sign = first == '-' || first == '+';
    if (sign) {
      c >>= 8;
      l -= 1;
    }
    int64_t v;
    if (l >= 1 && parse_digits8(c, l, &v)) {
      *out = first == '-' ? -v : v;
      return true;
    }
  }
  int i = 0;
  bool neg = false;
  if (len > 0 && (s[0] == '-' || s[0] == '+')) {
    neg = s[0] == '-';
    i = 1;
  }
  if (i >= len || len - i > 18) return false;
  int64_t v = 0;
  for (; i < len; ++i) {
    const unsigned d = static_cast<unsigned>(s[i]) - '0';
    if (d > 9u) return false;
    v = v * 10 + static_cast<int64_t>(d);
  }
  *out = neg ? -v : v;
  return true;
}

// :46-53 for one integer column: '' -> 0, negative -> 0; the caller converts to float32 and takes log(x + 1).
RB_HD bool int_column(const uint8_t* s, int len, int64_t* out) {
  if (len == 0) {
    *out = 0;
    return true;
  }
  int64_t v;
  if (!parse_int(s, len, &v)) return false;
  *out = v < 0 ? 0 : v;
  return true;
}

// Key of the null-imputation token of categorical field `field` (0-based): low byte zero, which no non-empty token
// has; one key per field like the reference's 26 random strings (:11-12).
RB_HD uint64_t missing_key(int field) { return static_cast<uint64_t>(field + 1) << 8; }

// 64-bit key of one categorical token (:21-23 / :54-57).  s[0..len) are the column's bytes WITHOUT the line's
// newline; trailing_newline says whether str.split left a '\n' attached (last column of a newline-terminated line),
// which makes it a different dictionary key from the same bytes elsewhere.  Empty -> the field's imputation key.
RB_HD uint64_t token_key(const uint8_t* s, int len, bool trailing_newline, int field, int* err) {
  if (len == 0) return missing_key(field);      // '' and '\n' are both imputed
  if (len > 8) {
    *err |= kErrLongToken;
    len = 8;
  }
  uint64_t k = load8(s);
  if (len < 8) k &= (1ull << (8 * len)) - 1;
  if (k & 0x8080808080808080ull) {
    *err |= kErrNonAscii;
    k &= 0x7F7F7F7F7F7F7F7Full;
  }
  return trailing_newline ? (k | kNewlineBit) : k;
}

// Column f of a line whose tab k sits at byte tab[k] (k < min(ntabs, 40)): [*start, *stop).  line.split('\t')[f].
// Needs ntabs >= 39 (a shorter line raises IndexError in the reference).
template <class TabT>
RB_HD void column_span(const TabT* tab, int ntabs, int len, int f, int* start, int* stop) {
  *start = f == 0 ? 0 : static_cast<int>(tab[f - 1]) + 1;
  *stop = f < ntabs ? static_cast<int>(tab[f]) : len;
}

// str.split leaves the line's '\n' attached to column 39 only when it is the line's last column.
RB_HD bool last_column_keeps_newline(int ntabs, bool line_has_newline) { return ntabs < kCols && line_has_newline; }

// splitmix64 finaliser: slot hash of the vocabulary table
RB_HD uint64_t mix64(uint64_t k) {
  k ^= k >> 30;
  k *= 0xBF58476D1CE4E5B9ull;
  k ^= k >> 27;
  k *= 0x94D049BB133111EBull;
  k ^= k >> 31;
  return k;
}

// Open addressing, linear probing, capacity = mask + 1 a power of two with at least one empty slot.
// A missing key maps to id 0 — the reference's OOV rule (:61-64), which collides with the first vocabulary entry.
RB_HD int64_t vocab_find(const uint64_t* keys, const int32_t* vals, uint64_t mask, uint64_t key) {
  uint64_t slot = mix64(key) & mask;
  for (;;) {
    const uint64_t k = keys[slot];
    if (k == key) return vals[slot];
    if (k == kEmptySlot) return 0;
    slot = (slot + 1) & mask;
  }
}

}  // namespace criteo
}  // namespace rb
