/*
 * cz_loop.c -- ORACLE (test infrastructure only; never linked by the product).
 * Restates the reference's coding loops and reprime schedule:
 *   encode: src/main.rs:1913-1916, 1935, 1979, 2275-2300 (smollm) / 2301-2326 (rwkv), 2344-2350, 2358
 *   decode: src/main.rs:2485, 2506, 2528-2541, 2621-2627 (smollm); 2706-2864 (rwkv7)
 *   hint priming in the main stream: 2123-2149 (enc), 2586-2614 (dec) -- supplied as czo_prime_event
 *   gate cross-entropy: 1725-1751, 1753-1787
 */
#include "cz_oracle.h"
#include "cz_session_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }

int czo_encode_tokens(czo_session *s, const uint32_t *ids, size_t n_ids, czo_loop_opts *o, uint8_t **out,
                      size_t *out_len) {
  const size_t v = s->vocab;
  const int rwkv = o->backend == 1;
  const size_t n_sym = rwkv ? v + 256 : v;
  uint32_t *cdf = (uint32_t *)malloc((n_sym + 1) * sizeof(uint32_t));
  czo_encoder *ace = czo_encoder_new();
  const float *logits = czo_session_step_logits(s, ids[0]); /* :1916 */
  const size_t effective_context = rwkv ? (size_t)-1 : min_sz(o->context, 511); /* :1935 */
  size_t hold_until = 0, ev = 0;
  int rc = 0;
  o->n_reprimes = 0;
  for (size_t i = 0; i + 1 < n_ids; i++) { /* :1979 */
    uint32_t sym = ids[i + 1];
    /* gated hint prime at an agent boundary (:1981, 2123-2149) */
    if (ev < o->n_events && o->events[ev].i == i) {
      logits = czo_session_reprime(s, o->events[ev].prime, o->events[ev].prime_len);
      hold_until = o->events[ev].hold_until;
      ev++;
    }
    if (!rwkv) { /* :2275-2290 */
      if (i < hold_until) {
      } else if (s->index_pos >= effective_context && (i % o->reprime_interval) == 0 && i > 0) {
        size_t end = 1 + i;
        size_t start = end > effective_context ? end - effective_context : 0;
        logits = czo_session_reprime(s, ids + start, end - start);
        if (o->reprime_log && o->n_reprimes < o->reprime_log_cap) o->reprime_log[o->n_reprimes] = i;
        o->n_reprimes++;
      }
    }
    if ((size_t)sym >= n_sym) { rc = -2; break; }
    czo_logits_to_cdf(logits, v, rwkv, cdf); /* :2294-2296 / 2303-2321 */
    if (czo_encoder_encode_counts(ace, cdf[sym], cdf[sym + 1], CZO_AC_CDF_TOTAL) != 0) { rc = -1; break; }
    if (!rwkv || (size_t)sym < v) logits = czo_session_step_logits(s, sym); /* :2344-2350 */
  }
  if (rc == 0) {
    size_t len;
    const uint8_t *p = czo_encoder_finish(ace, &len); /* :2358 */
    *out = (uint8_t *)malloc(len ? len : 1);
    memcpy(*out, p, len);
    *out_len = len;
  }
  czo_encoder_free(ace);
  free(cdf);
  return rc;
}

int czo_decode_tokens(czo_session *s, const uint8_t *payload, size_t len, uint32_t bos, size_t token_count,
                      czo_loop_opts *o, uint32_t *ids_out) {
  const size_t v = s->vocab;
  const int rwkv = o->backend == 1;
  const size_t n_sym = rwkv ? v + 256 : v;
  uint32_t *cdf = (uint32_t *)malloc((n_sym + 1) * sizeof(uint32_t));
  czo_decoder *acd = czo_decoder_new(payload, len); /* :2485 */
  const float *logits = czo_session_step_logits(s, bos);
  size_t n_out = 0;
  ids_out[n_out++] = bos;
  /* :2506 -- note (context-1).min(511), vs context.min(511) on the encode side */
  const size_t effective_context = rwkv ? (size_t)-1 : min_sz(o->context ? o->context - 1 : 0, 511);
  size_t hold_until = 0, ev = 0;
  o->n_reprimes = 0;
  for (size_t i = 0; i < token_count; i++) { /* :2528 */
    if (!rwkv) {
      if (i < hold_until) {
      } else if (s->index_pos >= effective_context && (i % o->reprime_interval) == 0 && i > 0) { /* :2532 */
        size_t end = n_out;
        size_t start = end > effective_context ? end - effective_context : 0;
        logits = czo_session_reprime(s, ids_out + start, end - start);
        if (o->reprime_log && o->n_reprimes < o->reprime_log_cap) o->reprime_log[o->n_reprimes] = i;
        o->n_reprimes++;
      }
    }
    if (ev < o->n_events && o->events[ev].i == i) { /* :2543, 2586-2614 */
      logits = czo_session_reprime(s, o->events[ev].prime, o->events[ev].prime_len);
      hold_until = o->events[ev].hold_until;
      ev++;
    }
    czo_logits_to_cdf(logits, v, rwkv, cdf); /* :2622-2624 */
    uint32_t sym = (uint32_t)czo_decoder_decode_symbol_counts(acd, cdf, n_sym + 1, CZO_AC_CDF_TOTAL);
    ids_out[n_out++] = sym;
    if (!rwkv || (size_t)sym < v) logits = czo_session_step_logits(s, sym); /* :2626-2627 / 2832-2834 */
  }
  czo_decoder_free(acd);
  free(cdf);
  return 0;
}

double czo_xe_bits_over_span(czo_session *s, int backend, const uint32_t *history, size_t nh,
                             const uint32_t *targets, size_t nt, const uint32_t *hint, size_t nhint) {
  if (nt == 0) return 0.0;
  const size_t v = s->vocab;
  size_t max_ctx = s->max_context_length ? s->max_context_length - 1 : 0; /* saturating_sub(1) */
  uint32_t *hist_f = NULL;
  if (backend == 1) { /* :1763-1765 filter literals out of the history */
    hist_f = (uint32_t *)malloc((nh ? nh : 1) * sizeof(uint32_t));
    size_t k = 0;
    for (size_t i = 0; i < nh; i++)
      if ((size_t)history[i] < v) hist_f[k++] = history[i];
    history = hist_f;
    nh = k;
  }
  size_t hint_budget = min_sz(nhint, max_ctx);
  size_t remaining = max_ctx > hint_budget ? max_ctx - hint_budget : 0;
  size_t hist_take = min_sz(remaining, nh);
  size_t np = hist_take + hint_budget;
  uint32_t *prime = (uint32_t *)calloc(np ? np : 1, sizeof(uint32_t));
  memcpy(prime, history + (nh - hist_take), hist_take * sizeof(uint32_t));
  if (hint_budget) memcpy(prime + hist_take, hint, hint_budget * sizeof(uint32_t));
  const float *logits = czo_session_reprime(s, prime, np);
  size_t n_sym = backend == 1 ? v + 256 : v;
  double *pdf = (double *)malloc(n_sym * sizeof(double));
  double bits = 0.0;
  if (logits) {
    for (size_t t = 0; t < nt; t++) {
      uint32_t sym = targets[t];
      if (backend == 1) czo_combined_pdf_with_literals(logits, v, pdf); /* :1778 */
      else czo_softmax_pdf_floor(logits, v, czo_ac_p_min(), pdf);       /* :1745 */
      double p = (size_t)sym < n_sym ? pdf[sym] : czo_ac_p_min();
      if (p < 1e-300) p = 1e-300;
      bits += -log2(p); /* :1747 */
      if (backend != 1 || (size_t)sym < v) logits = czo_session_step_logits(s, sym);
    }
  } else {
    bits = NAN; /* the reference bails on an empty prime */
  }
  free(pdf);
  free(prime);
  free(hist_f);
  return bits;
}
