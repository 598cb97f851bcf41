// Host build of recommender_b200/csrc/criteo_fields.h for tests/test_criteo_fields_cpu.py (g++, no CUDA): the
// field-level functions the parse kernel calls, driven by a scalar walk over one line.  TEST INFRASTRUCTURE.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../recommender_b200/csrc/criteo_fields.h"

using namespace rb::criteo;

// line[0..len): one line WITHOUT its newline; has_nl: the newline was there.  Returns the error bits.
extern "C" int t_parse_line(const uint8_t* text, int len, int has_nl, int64_t* label, float* ints, uint64_t* keys) {
  std::vector<uint8_t> padded(static_cast<size_t>(len) + 16, 0xAB);   // load8 reads up to 11 bytes behind a column
  memcpy(padded.data(), text, static_cast<size_t>(len));
  const uint8_t* line = padded.data();
  int16_t tab[kCols];
  int ntabs = 0;
  for (int p = 0; p < len; ++p)
    if (line[p] == '\t') {
      if (ntabs < kCols) tab[ntabs] = static_cast<int16_t>(p);
      ++ntabs;
    }
  if (ntabs < kCols - 1) return kErrShortLine;
  int err = 0;
  for (int f = 0; f < kCols; ++f) {
    int start, stop;
    column_span(tab, ntabs, len, f, &start, &stop);
    const uint8_t* s = line + start;
    const int flen = stop - start;
    if (f == 0) {
      int64_t v = 0;
      if (!parse_int(s, flen, &v)) err |= kErrBadInt;
      *label = v;
    } else if (f <= kNumInt) {
      int64_t v = 0;
      if (!int_column(s, flen, &v)) err |= kErrBadInt;
      ints[f - 1] = logf(static_cast<float>(v) + 1.0f);
    } else {
      const int field = f - kNumInt - 1;
      keys[field] = token_key(s, flen, f == kCols - 1 && last_column_keeps_newline(ntabs, has_nl != 0), field, &err);
    }
  }
  return err;
}

extern "C" uint64_t t_mix64(uint64_t k) { return mix64(k); }

extern "C" int64_t t_vocab_find(const uint64_t* keys, const int32_t* vals, uint64_t mask, uint64_t key) {
  return vocab_find(keys, vals, mask, key);
}

// sequential counterpart of table_insert_kernel
extern "C" void t_table_build(const uint64_t* vocab_keys, int64_t n, uint64_t* tkeys, int32_t* tvals, uint64_t mask) {
  for (uint64_t s = 0; s <= mask; ++s) {
    tkeys[s] = kEmptySlot;
    tvals[s] = 0;
  }
  for (int64_t i = 0; i < n; ++i) {
    uint64_t slot = mix64(vocab_keys[i]) & mask;
    while (tkeys[slot] != kEmptySlot && tkeys[slot] != vocab_keys[i]) slot = (slot + 1) & mask;
    tkeys[slot] = vocab_keys[i];
    tvals[slot] = static_cast<int32_t>(i);
  }
}

extern "C" int t_parse_int(const uint8_t* text, int len, int64_t* out) {
  uint8_t padded[64];
  memset(padded, 0xAB, sizeof(padded));
  memcpy(padded, text, static_cast<size_t>(len < 40 ? len : 40));
  return parse_int(padded, len, out) ? 1 : 0;
}
