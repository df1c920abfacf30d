// K2 / K3 device-side arithmetic coder: 32-bit binary AC with pending-bit (E3) counter.
// Bit-exact restatement of src/main.rs:261-404 (encoder) and 406-449, 500-548 (decoder, integer path).
// One lane (thread) per independent stream; integer-only; bits are MSB-first (src/main.rs:299-309).
#pragma once
#include <stdint.h>

namespace czk {

struct AcEncoder {
  uint64_t low, high, carry_run;
  uint8_t *out;      // lane's output region (global)
  uint64_t n_bytes;  // bytes emitted so far (whole 32-bit words until finish()), including bytes still in the staging buffer
  // Bit packer (src/main.rs:299-318 emits one bit at a time into a byte): here the bits queue up in a 64-bit accumulator (fewer
  // than 32 pending between calls) and leave as big-endian 32-bit words -- MSB-first within the stream, so the byte sequence is the
  // same -- one store per 32 bits instead of a shift / compare / store per byte on the single thread that owns the stream.
  uint64_t bit_acc;
  uint32_t n_bits;
  uint8_t *stage;    // optional shared-memory staging buffer, 4-byte aligned (nullptr: write straight to `out`)
  uint32_t stage_cap, stage_n;

  __device__ __forceinline__ void init(uint8_t *o) {
    low = 0;
    high = 0xFFFFFFFFull;
    carry_run = 0;
    out = o;
    n_bytes = 0;
    bit_acc = 0;
    n_bits = 0;
    stage = nullptr;
    stage_cap = stage_n = 0;
  }
  // single-thread flush of the staging buffer (slow path: only when a carry burst overfills it)
  __device__ __forceinline__ void flush_stage_serial() {
    const uint64_t b0 = n_bytes - stage_n;
    for (uint32_t i = 0; i < stage_n; i++) out[b0 + i] = stage[i];
    stage_n = 0;
  }
  __device__ __forceinline__ void emit_byte(uint32_t v) {
    if (stage) {
      if (stage_n == stage_cap) flush_stage_serial();
      stage[stage_n++] = (uint8_t)v;
    } else {
      out[n_bytes] = (uint8_t)v;
    }
    n_bytes++;
  }
  __device__ __forceinline__ void emit_word(uint32_t w) {  // the next four bytes of the stream, first byte = top byte of w
    if (stage) {
      if (stage_n + 4 > stage_cap) flush_stage_serial();
      *reinterpret_cast<uint32_t *>(stage + stage_n) = __byte_perm(w, 0u, 0x0123);  // (stage_n stays a multiple of 4 until finish())
      stage_n += 4;
    } else {
      out[n_bytes] = (uint8_t)(w >> 24);
      out[n_bytes + 1] = (uint8_t)(w >> 16);
      out[n_bytes + 2] = (uint8_t)(w >> 8);
      out[n_bytes + 3] = (uint8_t)w;
    }
    n_bytes += 4;
  }
  // append the low `n` (1..32) bits of `bits`, MSB first, to the byte stream
  __device__ __forceinline__ void put_bits_plain(uint32_t bits, uint32_t n) {
    const uint64_t v = n >= 32 ? (uint64_t)bits : ((uint64_t)bits & ((1ull << n) - 1ull));
    bit_acc = (bit_acc << n) | v;
    n_bits += n;
    if (n_bits >= 32) {
      n_bits -= 32;
      emit_word((uint32_t)(bit_acc >> n_bits));
    }
  }
  __device__ __forceinline__ void put_bit_internal(uint32_t bit) { put_bits_plain(bit & 1u, 1); }
  __device__ __forceinline__ void put_bit(uint32_t bit) {
    put_bit_internal(bit);
    while (carry_run > 0) {
      put_bit_internal((~bit) & 1u);
      carry_run--;
    }
  }
  // total is fixed at 2^30 (AC_CDF_TOTAL), so floor(range*c/total) is a shift. Caller guarantees c_lo < c_hi <= 2^30.
  //
  // The reference renormalises one bit per loop iteration (src/main.rs:367-383).  The iteration sequence is always
  // (A|B)^n C^k: cases A/B (emit the common MSB) repeat while the top bits of low and high agree; a case-C step
  // (low = 01.., high = 10..: carry_run++) leaves low = 0.., high = 1.., after which only C can follow.  So the loop is
  // done in bulk: n = clz(low ^ high) common bits are emitted at once (the first one followed by the pending carry_run
  // inverted bits), then k = number of straddle steps is read off the bit patterns.  Bit-identical output, ~10x fewer
  // dependent instructions per symbol for the single thread that owns the stream.
  __device__ __forceinline__ void encode_counts(uint32_t c_lo, uint32_t c_hi) {
    const uint64_t range = high - low + 1;
    uint32_t lo32 = (uint32_t)(low + ((range * (uint64_t)c_lo) >> 30));
    uint32_t hi32 = (uint32_t)(low + ((range * (uint64_t)c_hi) >> 30) - 1);
    const uint32_t n = (uint32_t)__clz((int)(lo32 ^ hi32));  // common leading bits (32 if equal)
    if (n > 0) {
      const uint32_t top = n == 32 ? lo32 : (lo32 >> (32 - n));
      const uint32_t first = (top >> (n - 1)) & 1u;
      put_bits_plain(first, 1);
      if (carry_run > 0) {  // pending straddle bits: the inverse of the first resolved bit
        const uint32_t inv = first ? 0u : 0xFFFFFFFFu;
        while (carry_run >= 32) {
          put_bits_plain(inv, 32);
          carry_run -= 32;
        }
        if (carry_run > 0) put_bits_plain(inv, (uint32_t)carry_run);
        carry_run = 0;
      }
      if (n > 1) put_bits_plain(top & ((n - 1 == 32) ? 0xFFFFFFFFu : ((1u << (n - 1)) - 1u)), n - 1);
      if (n == 32) {
        lo32 = 0u;
        hi32 = 0xFFFFFFFFu;
      } else {
        lo32 <<= n;
        hi32 = (hi32 << n) | ((1u << n) - 1u);
      }
    }
    // now lo32 = 0..., hi32 = 1...; straddle steps while lo32 = 01.. and hi32 = 10..
    uint32_t k = (uint32_t)__clz((int)~(lo32 << 1));          // ones in lo32 starting at bit 30
    const uint32_t kz = (uint32_t)__clz((int)((hi32 << 1) | 1u));  // zeros in hi32 starting at bit 30 (<= 31)
    k = k < kz ? k : kz;
    if (k > 31) k = 31;
    if (k > 0) {
      carry_run += k;
      lo32 = (lo32 << k) & 0x7FFFFFFFu;
      hi32 = 0x80000000u | ((hi32 << k) & 0x7FFFFFFFu) | ((1u << k) - 1u);
    }
    low = lo32;
    high = hi32;
  }
  // one-bit-per-iteration form, literally src/main.rs:353-385 (kept for finish() and as the definition)
  __device__ __forceinline__ void encode_counts_ref(uint32_t c_lo, uint32_t c_hi) {
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
    if (n_bits & 7u) put_bits_plain(0u, 8u - (n_bits & 7u));  // zero-pad to a byte
    while (n_bits > 0) {                                      // the pending whole bytes (at most three)
      n_bits -= 8;
      emit_byte((uint32_t)(bit_acc >> n_bits) & 0xFFu);
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
