/*
 * cz_coder.c -- ORACLE (test infrastructure only; never linked by the product).
 * Restates src/main.rs:230-238 (ac_p_min), 261-404 (ArithmeticEncoder), 406-549
 * (ArithmeticDecoder, integer path), 551-670 (container v2), 758-824 (pdf / CDF).
 * Compile with -ffp-contract=off: every fused op below is written as an explicit fma().
 */
#include "cz_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * expf.  Rust's f32::exp lowers to the platform libm expf.  On x86_64 glibc >= 2.28 that is the
 * ARM optimized-routines algorithm (sysdeps/ieee754/flt-32/e_expf.c) built with FMA contraction
 * (the __expf_fma ifunc variant).  Restated here so the oracle is independent of the host libm;
 * oracle/expf_exhaustive.c proves bit-equality with the host expf over the whole finite domain.
 * ------------------------------------------------------------------------------------------ */
static const uint64_t EXP2F_TAB[32] = {
    0x3ff0000000000000, 0x3fefd9b0d3158574, 0x3fefb5586cf9890f, 0x3fef9301d0125b51,
    0x3fef72b83c7d517b, 0x3fef54873168b9aa, 0x3fef387a6e756238, 0x3fef1e9df51fdee1,
    0x3fef06fe0a31b715, 0x3feef1a7373aa9cb, 0x3feedea64c123422, 0x3feece086061892d,
    0x3feebfdad5362a27, 0x3feeb42b569d4f82, 0x3feeab07dd485429, 0x3feea47eb03a5585,
    0x3feea09e667f3bcd, 0x3fee9f75e8ec5f74, 0x3feea11473eb0187, 0x3feea589994cce13,
    0x3feeace5422aa0db, 0x3feeb737b0cdc5e5, 0x3feec49182a3f090, 0x3feed503b23e255d,
    0x3feee89f995ad3ad, 0x3feeff76f2fb5e47, 0x3fef199bdd85529c, 0x3fef3720dcef9069,
    0x3fef5818dcfba487, 0x3fef7c97337b9b5f, 0x3fefa4afa2a490da, 0x3fefd0765b6e4540,
};

float czo_expf(float x) {
  const double inv_ln2_n = 0x1.71547652b82fep+0 * 32;
  const double shift = 0x1.8p+52;
  const double c0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32;
  const double c1 = 0x1.ebfce50fac4f3p-3 / 32 / 32;
  const double c2 = 0x1.62e42ff0c52d6p-1 / 32;
  if (x != x) return x;
  if (x < -0x1.9fe368p6f) return 0.0f;           /* underflow (also -inf) */
  if (x > 0x1.62e42ep6f) return INFINITY;          /* overflow */
  double xd = (double)x;
  double z = inv_ln2_n * xd;
  double kd = z + shift;
  uint64_t ki;
  memcpy(&ki, &kd, 8);
  kd -= shift;
  double r = fma(inv_ln2_n, xd, -kd);              /* contracted z - kd */
  uint64_t t = EXP2F_TAB[ki % 32];
  t += ki << (52 - 5);
  double s;
  memcpy(&s, &t, 8);
  double p = fma(c0, r, c1);
  double r2 = r * r;
  double y = fma(c2, r, 1.0);
  y = fma(p, r2, y);
  y = y * s;
  return (float)y;
}

/* Checksum of czo_expf over the non-negative f32 bit patterns b in [b_lo, b_hi): sum of bits(expf(-float(b))) * (2 b + 1) mod 2^64,
 * NaN results counted as 0x7fc00000.  The GPU computes the same sum with its own expf (cz_test_expf_exhaustive): equal sums over
 * the whole domain [0, 2^31) tie the device expf to this one -- which expf_exhaustive.c ties to the host glibc expf that Rust's
 * f32::exp calls (src/main.rs:791). */
uint64_t czo_expf_checksum(uint64_t b_lo, uint64_t b_hi) {
  uint64_t sum = 0;
#pragma omp parallel for reduction(+ : sum) schedule(static)
  for (uint64_t b = b_lo; b < b_hi; b++) {
    uint32_t u = (uint32_t)b, r;
    float a, e;
    memcpy(&a, &u, 4);
    e = czo_expf(-a);
    memcpy(&r, &e, 4);
    if (e != e) r = 0x7fc00000u;
    sum += (uint64_t)r * (2ull * b + 1ull);
  }
  return sum;
}

double czo_ac_p_min(void) { return 2.0 * pow(2.0, -30.0); } /* src/main.rs:235-238: 2*2^-(32-2) = 2^-29 */

/* src/main.rs:784-801 */
void czo_softmax_pdf(const float *logits, size_t v, double *pdf) {
  float max = -INFINITY;
  for (size_t i = 0; i < v; i++)
    if (logits[i] > max) max = logits[i];
  double sum = 0.0;
  for (size_t i = 0; i < v; i++) {
    double e = (double)czo_expf(logits[i] - max);
    pdf[i] = e;
    sum += e;
  }
  if (sum <= 0.0) {
    double u = 1.0 / (double)v;
    for (size_t i = 0; i < v; i++) pdf[i] = u;
    return;
  }
  for (size_t i = 0; i < v; i++) pdf[i] = pdf[i] / sum;
}

/* src/main.rs:758-767 */
void czo_softmax_pdf_floor(const float *logits, size_t v, double p_floor, double *pdf) {
  float max = -INFINITY;
  for (size_t i = 0; i < v; i++)
    if (logits[i] > max) max = logits[i];
  double sum = 0.0;
  for (size_t i = 0; i < v; i++) {
    double e = (double)czo_expf(logits[i] - max);
    pdf[i] = e;
    sum += e;
  }
  for (size_t i = 0; i < v; i++) {
    double p = pdf[i] / sum;
    pdf[i] = p > p_floor ? p : p_floor; /* f64::max: NaN-ignoring; p is never NaN for finite logits */
  }
  double norm = 0.0;
  for (size_t i = 0; i < v; i++) norm += pdf[i];
  for (size_t i = 0; i < v; i++) pdf[i] /= norm;
}

/* src/main.rs:769-782 */
void czo_combined_pdf_with_literals(const float *logits, size_t v, double *pdf) {
  czo_softmax_pdf_floor(logits, v, czo_ac_p_min(), pdf);
  double p_escape_total = 256.0 * czo_ac_p_min();
  double scale = 1.0 - p_escape_total;
  if (scale < 0.0) scale = 0.0;
  for (size_t i = 0; i < v; i++) pdf[i] *= scale;
  double p_literal_each = p_escape_total > 0.0 ? p_escape_total / 256.0 : 0.0;
  for (size_t i = 0; i < 256; i++) pdf[v + i] = p_literal_each;
  double sum = 0.0;
  for (size_t i = 0; i < v + 256; i++) sum += pdf[i];
  if (sum > 0.0)
    for (size_t i = 0; i < v + 256; i++) pdf[i] /= sum;
}

/* src/main.rs:805-824 (and the identical inline RWKV quantiser 2305-2321) */
void czo_quantize_pdf_to_cdf(const double *pdf, size_t n, uint32_t *cdf) {
  double acc = 0.0;
  cdf[0] = 0;
  for (size_t i = 0; i < n; i++) {
    acc += pdf[i];
    double f = floor(acc * (double)CZO_AC_CDF_TOTAL);
    int64_t v;
    if (f != f) v = 0;                       /* Rust `as i64` maps NaN to 0 */
    else if (f >= 9.2e18) v = INT64_MAX;     /* saturating cast */
    else if (f <= -9.2e18) v = INT64_MIN;
    else v = (int64_t)f;
    if (v < 0) v = 0;
    if (v > (int64_t)CZO_AC_CDF_TOTAL) v = (int64_t)CZO_AC_CDF_TOTAL;
    int64_t prev = (int64_t)cdf[i];
    if (v < prev) v = prev;
    cdf[i + 1] = (uint32_t)v;
  }
  cdf[n] = CZO_AC_CDF_TOTAL;
}

void czo_logits_to_cdf(const float *logits, size_t v, int mode, uint32_t *cdf) {
  size_t n = mode == 1 ? v + 256 : v;
  double *pdf = (double *)malloc(n * sizeof(double));
  if (mode == 1) czo_combined_pdf_with_literals(logits, v, pdf);
  else czo_softmax_pdf(logits, v, pdf);
  czo_quantize_pdf_to_cdf(pdf, n, cdf);
  free(pdf);
}

/* ------------------------------------------------------------------------------------------
 * Arithmetic encoder: src/main.rs:261-404
 * ------------------------------------------------------------------------------------------ */
struct czo_encoder {
  uint64_t b_to_pm1, b_to_pm2, mask, low, high, carry_run;
  uint8_t *out;
  size_t len, cap;
  uint8_t bit_buffer, bit_count;
};

czo_encoder *czo_encoder_new(void) {
  czo_encoder *e = (czo_encoder *)calloc(1, sizeof(*e));
  e->b_to_pm1 = 1ull << 31;
  e->b_to_pm2 = 1ull << 30;
  e->mask = (1ull << 32) - 1;
  e->low = 0;
  e->high = e->mask;
  e->cap = 1024;
  e->out = (uint8_t *)malloc(e->cap);
  return e;
}
void czo_encoder_free(czo_encoder *e) {
  if (!e) return;
  free(e->out);
  free(e);
}
static void enc_write_byte(czo_encoder *e, uint8_t b) {
  if (e->len == e->cap) {
    e->cap *= 2;
    e->out = (uint8_t *)realloc(e->out, e->cap);
  }
  e->out[e->len++] = b;
}
static void enc_put_bit_internal(czo_encoder *e, uint8_t bit) { /* :299-309 */
  e->bit_buffer = (uint8_t)((e->bit_buffer << 1) | (bit & 1));
  e->bit_count++;
  if (e->bit_count == 8) {
    enc_write_byte(e, e->bit_buffer);
    e->bit_buffer = 0;
    e->bit_count = 0;
  }
}
static void enc_put_bit(czo_encoder *e, uint8_t bit) { /* :311-318 */
  enc_put_bit_internal(e, bit);
  while (e->carry_run > 0) {
    enc_put_bit_internal(e, (uint8_t)((~bit) & 1));
    e->carry_run--;
  }
}
int czo_encoder_encode_counts(czo_encoder *e, uint64_t c_lo, uint64_t c_hi, uint64_t total) { /* :353-385 */
  if (c_hi <= c_lo || c_hi > total) return -1;
  unsigned __int128 range = (unsigned __int128)(e->high - e->low + 1);
  unsigned __int128 new_low = (unsigned __int128)e->low + (range * c_lo) / total;
  unsigned __int128 new_high = (unsigned __int128)e->low + (range * c_hi) / total - 1;
  e->low = (uint64_t)(new_low & e->mask);
  e->high = (uint64_t)(new_high & e->mask);
  for (;;) {
    if (e->high < e->b_to_pm1) {
      enc_put_bit(e, 0);
    } else if (e->low >= e->b_to_pm1) {
      enc_put_bit(e, 1);
      e->low -= e->b_to_pm1;
      e->high -= e->b_to_pm1;
    } else if (e->low >= e->b_to_pm2 && e->high < e->b_to_pm2 * 3) {
      e->carry_run++;
      e->low -= e->b_to_pm2;
      e->high -= e->b_to_pm2;
    } else {
      break;
    }
    e->low = (e->low << 1) & e->mask;
    e->high = ((e->high << 1) & e->mask) | 1;
  }
  return 0;
}
uint64_t czo_encoder_bytes_written(const czo_encoder *e) { return (uint64_t)e->len; }
const uint8_t *czo_encoder_finish(czo_encoder *e, size_t *len) { /* :387-399 */
  e->carry_run++;
  if (e->low < e->b_to_pm2) enc_put_bit(e, 0);
  else enc_put_bit(e, 1);
  if (e->bit_count > 0) {
    int remaining = 8 - e->bit_count;
    for (int i = 0; i < remaining; i++) enc_put_bit_internal(e, 0);
  }
  *len = e->len;
  return e->out;
}

/* ------------------------------------------------------------------------------------------
 * Arithmetic decoder: src/main.rs:406-449, 500-548
 * ------------------------------------------------------------------------------------------ */
struct czo_decoder {
  uint64_t b_to_pm1, b_to_pm2, mask, low, high, code;
  const uint8_t *input;
  size_t len, byte_pos;
  uint8_t bit_pos;
};
static uint8_t dec_get_bit(czo_decoder *d) { /* :439-449, EOF reads as 1 (:434, :544) */
  if (d->byte_pos >= d->len) return 1;
  uint8_t byte = d->input[d->byte_pos];
  uint8_t bit = (uint8_t)((byte >> (7 - d->bit_pos)) & 1);
  d->bit_pos++;
  if (d->bit_pos >= 8) {
    d->bit_pos = 0;
    d->byte_pos++;
  }
  return bit;
}
czo_decoder *czo_decoder_new(const uint8_t *payload, size_t len) { /* :419-437 */
  czo_decoder *d = (czo_decoder *)calloc(1, sizeof(*d));
  d->b_to_pm1 = 1ull << 31;
  d->b_to_pm2 = 1ull << 30;
  d->mask = (1ull << 32) - 1;
  d->low = 0;
  d->high = d->mask;
  d->input = payload;
  d->len = len;
  for (int i = 0; i < 32; i++) d->code = (d->code << 1) | dec_get_bit(d);
  return d;
}
void czo_decoder_free(czo_decoder *d) { free(d); }
uint32_t czo_decoder_peek_value(const czo_decoder *d, uint32_t total) { /* :502-505 */
  uint64_t range = d->high - d->low + 1;
  unsigned __int128 value = ((unsigned __int128)(d->code - d->low + 1) * total - 1) / range;
  return (uint32_t)value;
}
size_t czo_decoder_decode_symbol_counts(czo_decoder *d, const uint32_t *cdf, size_t cdf_len, uint32_t total) {
  uint32_t value_u = czo_decoder_peek_value(d, total);
  size_t lo = 0, hi = cdf_len - 1; /* :508-513 */
  while (lo + 1 < hi) {
    size_t mid = (lo + hi) / 2;
    if (cdf[mid] <= value_u) lo = mid;
    else hi = mid;
  }
  size_t s = lo;
  uint64_t c_lo = cdf[s], c_hi = cdf[s + 1];
  unsigned __int128 range = (unsigned __int128)(d->high - d->low + 1);
  unsigned __int128 new_low = (unsigned __int128)d->low + (range * c_lo) / total;
  unsigned __int128 new_high = (unsigned __int128)d->low + (range * c_hi) / total - 1;
  d->low = (uint64_t)new_low;
  d->high = (uint64_t)new_high;
  for (;;) { /* :528-545 */
    if (d->high < d->b_to_pm1) {
    } else if (d->low >= d->b_to_pm1) {
      d->low -= d->b_to_pm1;
      d->high -= d->b_to_pm1;
      d->code -= d->b_to_pm1;
    } else if (d->low >= d->b_to_pm2 && d->high < d->b_to_pm2 * 3) {
      d->low -= d->b_to_pm2;
      d->high -= d->b_to_pm2;
      d->code -= d->b_to_pm2;
    } else {
      break;
    }
    d->low = (d->low << 1) & d->mask;
    d->high = ((d->high << 1) & d->mask) | 1;
    d->code = ((d->code << 1) & d->mask) | dec_get_bit(d);
  }
  return s;
}

/* ------------------------------------------------------------------------------------------
 * Container v2: src/main.rs:227-259, 551-646
 * ------------------------------------------------------------------------------------------ */
size_t czo_write_var_u64(uint8_t *buf, uint64_t v) { /* :566-574 */
  size_t n = 0;
  while (v >= 0x80) {
    buf[n++] = (uint8_t)((v & 0x7F) | 0x80);
    v >>= 7;
  }
  buf[n++] = (uint8_t)v;
  return n;
}
static size_t read_var_u64(const uint8_t *buf, size_t len, uint64_t *out) { /* :576-590 */
  uint32_t shift = 0;
  uint64_t v = 0;
  size_t n = 0;
  for (;;) {
    if (n >= len) return 0;
    uint8_t byte = buf[n++];
    v |= (uint64_t)(byte & 0x7F) << shift;
    if ((byte & 0x80) == 0) break;
    shift += 7;
    if (shift > 63) return 0;
  }
  *out = v;
  return n;
}
static void put_u32(uint8_t *b, uint32_t v) { b[0] = (uint8_t)v; b[1] = (uint8_t)(v >> 8); b[2] = (uint8_t)(v >> 16); b[3] = (uint8_t)(v >> 24); }
static uint32_t get_u32(const uint8_t *b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); }

uint32_t czo_flags_pack(int agent_used, int agent_mock, int gates_present, uint32_t agent_chunk) { /* :247-255 */
  uint32_t f = 0;
  if (agent_used) f |= 1u << 0;
  if (agent_mock) f |= 1u << 1;
  if (gates_present) f |= 1u << 2;
  f |= (agent_chunk & 0xFFFF) << 16;
  return f;
}

size_t czo_write_header_v2(uint8_t *buf, size_t cap, const czo_header_v2 *h, const uint8_t *repr) { /* :592-609 */
  if (cap < 96 + (size_t)h->model_file_repr_len) return 0;
  size_t n = 0;
  put_u32(buf + n, 0x5a505447u); n += 4;
  buf[n++] = 2; buf[n++] = 0;
  put_u32(buf + n, h->bos_token_id); n += 4;
  n += czo_write_var_u64(buf + n, h->token_count);
  n += czo_write_var_u64(buf + n, h->orig_len_bytes);
  memcpy(buf + n, h->model_hash16, 16); n += 16;
  memcpy(buf + n, h->tokenizer_hash16, 16); n += 16;
  memcpy(buf + n, h->orig_hash16, 16); n += 16;
  put_u32(buf + n, h->reserved_flags); n += 4;
  put_u32(buf + n, h->context_window); n += 4;
  put_u32(buf + n, h->vocab_size); n += 4;
  put_u32(buf + n, h->model_file_repr_len); n += 4;
  put_u32(buf + n, h->reprime_interval); n += 4;
  memcpy(buf + n, repr, h->model_file_repr_len); n += h->model_file_repr_len;
  return n;
}

size_t czo_read_header_v2(const uint8_t *buf, size_t len, czo_header_v2 *h, size_t *repr_off) { /* :611-646 */
  size_t n = 0, k;
  if (len < 10) return 0;
  if (get_u32(buf) != 0x5a505447u) return 0;
  n = 4;
  if (buf[n] != 2 || buf[n + 1] != 0) return 0;
  n += 2;
  h->bos_token_id = get_u32(buf + n); n += 4;
  if (!(k = read_var_u64(buf + n, len - n, &h->token_count))) return 0;
  n += k;
  if (!(k = read_var_u64(buf + n, len - n, &h->orig_len_bytes))) return 0;
  n += k;
  if (len < n + 48 + 20) return 0;
  memcpy(h->model_hash16, buf + n, 16); n += 16;
  memcpy(h->tokenizer_hash16, buf + n, 16); n += 16;
  memcpy(h->orig_hash16, buf + n, 16); n += 16;
  h->reserved_flags = get_u32(buf + n); n += 4;
  h->context_window = get_u32(buf + n); n += 4;
  h->vocab_size = get_u32(buf + n); n += 4;
  h->model_file_repr_len = get_u32(buf + n); n += 4;
  h->reprime_interval = get_u32(buf + n); n += 4;
  if (len < n + h->model_file_repr_len) return 0;
  if (repr_off) *repr_off = n;
  n += h->model_file_repr_len;
  return n;
}

void czo_free(void *p) { free(p); }
