// K2 / K3 device-side arithmetic coder: 32-bit binary AC with pending-bit (E3) counter.
// Bit-exact restatement of src/main.rs:261-404 (encoder) and 406-449, 500-548 (decoder, integer path).
// One lane (thread) per independent stream; integer-only; bits are MSB-first (src/main.rs:299-309).
#pragma once
#include <stdint.h>

namespace czk {

struct AcEncoder {
  uint64_t low, high, carry_run;
  uint8_t *out;      // lane's output region (global)
  uint64_t n_bytes;  // bytes_out, including bytes still in the staging buffer
  uint32_t bit_buffer;
  uint32_t bit_count;
  uint8_t *stage;    // optional shared-memory staging buffer (nullptr: write straight to `out`)
  uint32_t stage_cap, stage_n;

  __device__ __forceinline__ void init(uint8_t *o) {
    low = 0;
    high = 0xFFFFFFFFull;
    carry_run = 0;
    out = o;
    n_bytes = 0;
    bit_buffer = 0;
    bit_count = 0;
    stage = nullptr;
    stage_cap = stage_n = 0;
  }
  // single-thread flush of the staging buffer (slow path: only when a carry burst overfills it)
  __device__ __forceinline__ void flush_stage_serial() {
    const uint64_t b0 = n_bytes - stage_n;
    for (uint32_t i = 0; i < stage_n; i++) out[b0 + i] = stage[i];
    stage_n = 0;
  }
  __device__ __forceinline__ void put_bit_internal(uint32_t bit) {
    bit_buffer = (bit_buffer << 1) | (bit & 1u);
    if (++bit_count == 8) {
      if (stage) {
        if (stage_n == stage_cap) flush_stage_serial();
        stage[stage_n++] = (uint8_t)bit_buffer;
        n_bytes++;
      } else {
        out[n_bytes++] = (uint8_t)bit_buffer;
      }
      bit_buffer = 0;
      bit_count = 0;
    }
  }
  __device__ __forceinline__ void put_bit(uint32_t bit) {
    put_bit_internal(bit);
    while (carry_run > 0) {
      put_bit_internal((~bit) & 1u);
      carry_run--;
    }
  }
  // total is fixed at 2^30 (AC_CDF_TOTAL), so floor(range*c/total) is a shift. Caller guarantees c_lo < c_hi <= 2^30.
  __device__ __forceinline__ void encode_counts(uint32_t c_lo, uint32_t c_hi) {
    const uint64_t range = high - low + 1;
    const uint64_t new_low = low + ((range * (uint64_t)c_lo) >> 30);
    const uint64_t new_high = low + ((range * (uint64_t)c_hi) >> 30) - 1;
    low = new_low & 0xFFFFFFFFull;
    high = new_high & 0xFFFFFFFFull;
    for (;;) {
      if (high < 0x80000000ull) {
        put_bit(0);
      } else if (low >= 0x80000000ull) {
        put_bit(1);
        low -= 0x80000000ull;
        high -= 0x80000000ull;
      } else if (low >= 0x40000000ull && high < 0xC0000000ull) {
        carry_run++;
        low -= 0x40000000ull;
        high -= 0x40000000ull;
      } else {
        break;
      }
      low = (low << 1) & 0xFFFFFFFFull;
      high = ((high << 1) & 0xFFFFFFFFull) | 1ull;
    }
  }
  __device__ __forceinline__ uint64_t finish() {  // src/main.rs:387-399
    carry_run++;
    put_bit(low < 0x40000000ull ? 0u : 1u);
    if (bit_count > 0) {
      uint32_t remaining = 8 - bit_count;
      for (uint32_t i = 0; i < remaining; i++) put_bit_internal(0);
    }
    return n_bytes;
  }
};

// Persisted between decode steps (one per stream) -- 40 bytes.
struct AcDecoderState {
  uint64_t low, high, code;
  uint64_t byte_pos;
  uint32_t bit_pos;
  uint32_t pad;
};

struct AcDecoder {
  AcDecoderState s;
  const uint8_t *in;
  uint64_t len;

  __device__ __forceinline__ uint32_t get_bit() {  // past EOF reads as 1 (src/main.rs:434, 494, 544)
    if (s.byte_pos >= len) return 1u;
    uint32_t bit = (in[s.byte_pos] >> (7 - s.bit_pos)) & 1u;
    if (++s.bit_pos >= 8) {
      s.bit_pos = 0;
      s.byte_pos++;
    }
    return bit;
  }
  __device__ __forceinline__ void init(const uint8_t *payload, uint64_t n) {
    in = payload;
    len = n;
    s.low = 0;
    s.high = 0xFFFFFFFFull;
    s.code = 0;
    s.byte_pos = 0;
    s.bit_pos = 0;
    s.pad = 0;
    for (int i = 0; i < 32; i++) s.code = (s.code << 1) | get_bit();
  }
  __device__ __forceinline__ void resume(const AcDecoderState &st, const uint8_t *payload, uint64_t n) {
    s = st;
    in = payload;
    len = n;
  }
  // value = ((code - low + 1) * total - 1) / range   (src/main.rs:503-505), total = 2^30
  __device__ __forceinline__ uint32_t peek_value() const {
    const uint64_t range = s.high - s.low + 1;
    return (uint32_t)((((s.code - s.low + 1) << 30) - 1) / range);
  }
  __device__ __forceinline__ void consume(uint32_t c_lo, uint32_t c_hi) {  // src/main.rs:518-545
    const uint64_t range = s.high - s.low + 1;
    const uint64_t new_low = s.low + ((range * (uint64_t)c_lo) >> 30);
    const uint64_t new_high = s.low + ((range * (uint64_t)c_hi) >> 30) - 1;
    s.low = new_low;
    s.high = new_high;
    for (;;) {
      if (s.high < 0x80000000ull) {
      } else if (s.low >= 0x80000000ull) {
        s.low -= 0x80000000ull;
        s.high -= 0x80000000ull;
        s.code -= 0x80000000ull;
      } else if (s.low >= 0x40000000ull && s.high < 0xC0000000ull) {
        s.low -= 0x40000000ull;
        s.high -= 0x40000000ull;
        s.code -= 0x40000000ull;
      } else {
        break;
      }
      s.low = (s.low << 1) & 0xFFFFFFFFull;
      s.high = ((s.high << 1) & 0xFFFFFFFFull) | 1ull;
      s.code = ((s.code << 1) & 0xFFFFFFFFull) | get_bit();
    }
  }
};

}  // namespace czk
