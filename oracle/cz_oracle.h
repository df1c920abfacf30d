/*
 * cz_oracle.h -- CPU ORACLE for the candlezip hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's algorithm (turtle261/candlezip,
 * file:line citations are relative to /root/reference).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call it; the product library (candlezip_b200/csrc) never does.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - coder / CDF / schedule / container: restated from src/main.rs, pinned against the
 *     reference's shipped artefacts (container headers, reprime positions in
 *     results_300s_nomem watchdog traces) -- the Rust reference itself cannot be
 *     compiled in this environment (no cargo/rustc), so coder bitstreams are
 *     "parity unpinned" against a live reference run.
 *   - expf: restated glibc 2.39 x86_64 (__expf_fma) and checked exhaustively against the
 *     host libm (oracle/expf_exhaustive.c): 0 mismatches on [-104, 88].
 *   - SmolLM forward: candle-transformers 0.9.1 models::llama is NOT under
 *     /root/reference; restated from the published LLaMA/HF semantics and cross-checked
 *     against transformers.LlamaForCausalLM (tests/golden) -- "parity unpinned" at bit level.
 *   - RWKV-7: restated from candle_rwkv7/src/models/rwkv7.rs (t==1 path).
 */
#ifndef CZ_ORACLE_H
#define CZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CZO_AC_CDF_TOTAL (1u << 30) /* src/main.rs:803 */

/* ---- numeric kernels: src/main.rs:230-238, 758-824 ---- */
float czo_expf(float x);
uint64_t czo_expf_checksum(uint64_t b_lo, uint64_t b_hi);
double czo_ac_p_min(void);
void czo_softmax_pdf(const float *logits, size_t v, double *pdf);
void czo_softmax_pdf_floor(const float *logits, size_t v, double p_floor, double *pdf);
void czo_combined_pdf_with_literals(const float *logits, size_t v, double *pdf /* v+256 */);
void czo_quantize_pdf_to_cdf(const double *pdf, size_t n, uint32_t *cdf /* n+1 */);
/* logits -> integer CDF exactly as the encode/decode loops do it.
 * mode 0: SmolLM (src/main.rs:2294-2296); mode 1: RWKV with literals (2303-2321), cdf has v+257 entries */
void czo_logits_to_cdf(const float *logits, size_t v, int mode, uint32_t *cdf);

/* ---- arithmetic coder: src/main.rs:261-404, 406-549 ---- */
typedef struct czo_encoder czo_encoder;
czo_encoder *czo_encoder_new(void);
void czo_encoder_free(czo_encoder *e);
/* returns 0, or -1 on a zero-width / inverted interval (the reference would corrupt the stream) */
int czo_encoder_encode_counts(czo_encoder *e, uint64_t c_lo, uint64_t c_hi, uint64_t total);
uint64_t czo_encoder_bytes_written(const czo_encoder *e);
/* finish(): flushes; returns pointer to the internal buffer and its length */
const uint8_t *czo_encoder_finish(czo_encoder *e, size_t *len);

typedef struct czo_decoder czo_decoder;
czo_decoder *czo_decoder_new(const uint8_t *payload, size_t len);
void czo_decoder_free(czo_decoder *d);
size_t czo_decoder_decode_symbol_counts(czo_decoder *d, const uint32_t *cdf, size_t cdf_len, uint32_t total);
/* the target value the decoder derives before the search (src/main.rs:504) */
uint32_t czo_decoder_peek_value(const czo_decoder *d, uint32_t total);

/* ---- container v2: src/main.rs:227-259, 551-670 ---- */
typedef struct {
  uint32_t bos_token_id;
  uint64_t token_count;
  uint64_t orig_len_bytes;
  uint8_t model_hash16[16];
  uint8_t tokenizer_hash16[16];
  uint8_t orig_hash16[16];
  uint32_t reserved_flags;
  uint32_t context_window;
  uint32_t vocab_size;
  uint32_t model_file_repr_len;
  uint32_t reprime_interval;
} czo_header_v2;
/* returns bytes consumed (header + repr), or 0 on error; repr_off receives offset of the repr string */
size_t czo_read_header_v2(const uint8_t *buf, size_t len, czo_header_v2 *h, size_t *repr_off);
/* returns bytes written (needs cap >= 96 + repr_len) */
size_t czo_write_header_v2(uint8_t *buf, size_t cap, const czo_header_v2 *h, const uint8_t *repr);
size_t czo_write_var_u64(uint8_t *buf, uint64_t v);
uint32_t czo_flags_pack(int agent_used, int agent_mock, int gates_present, uint32_t agent_chunk);

/* ---- model sessions: src/models.rs:28-33 ---- */
typedef struct czo_session czo_session;

typedef struct {
  int vocab, d_model, n_layers, n_heads, n_kv_heads, head_dim, d_ffn;
  float rms_eps, rope_theta;
  int max_pos;        /* KV capacity */
  int round_bf16;     /* 0: pure f32 (reference CPU semantics). 1: round GEMM inputs / KV to bf16 like the B200 path */
} czo_llama_config;

czo_session *czo_llama_new(const czo_llama_config *cfg);
/* tensors by HF name: "model.embed_tokens.weight", "model.layers.%d.self_attn.q_proj.weight", ...
 * data is f32, row-major [out,in]; copied. returns 0 / -1 unknown name or size mismatch */
int czo_session_set_tensor(czo_session *s, const char *name, const float *data, size_t n);

typedef struct {
  int vocab, d_model, n_layers, head_dim, d_ffn;
  int lora_w, lora_a, lora_v, lora_g;
  float norm_eps;
  int round_bf16;     /* as in czo_llama_config */
} czo_rwkv7_config;
czo_session *czo_rwkv7_new(const czo_rwkv7_config *cfg);

void czo_session_free(czo_session *s);
size_t czo_session_vocab_size(const czo_session *s);
size_t czo_session_max_context_length(const czo_session *s);
size_t czo_session_index_pos(const czo_session *s);
/* both return a pointer to an internal [vocab] f32 buffer valid until the next call */
const float *czo_session_step_logits(czo_session *s, uint32_t token);
const float *czo_session_reprime(czo_session *s, const uint32_t *history, size_t n);
/* test hook: make the session return logits from a table instead of a model.
 * logits(t) = table[(hash of (pos, token)) % n_rows]. Used to test the loops without weights. */
czo_session *czo_table_session_new(size_t vocab, const float *table, size_t n_rows);

/* ---- loops: src/main.rs:1913-1941, 1979, 2275-2358 (encode); 2485, 2506, 2528-2541, 2621-2627 (decode) ---- */
typedef struct {
  size_t i;            /* iteration index at which the prime is applied (i+1 == agent boundary) */
  const uint32_t *prime;
  size_t prime_len;
  size_t hold_until;   /* reprime_hold_until after the prime (src/main.rs:2149, 2614) */
} czo_prime_event;

typedef struct {
  int backend;               /* 0 smollm, 1 rwkv7 */
  size_t context;            /* --context */
  size_t reprime_interval;   /* --reprime-interval */
  const czo_prime_event *events; /* gated hint primes, sorted by i; may be NULL */
  size_t n_events;
  size_t *reprime_log;       /* optional: receives i of each context_reprime */
  size_t reprime_log_cap;
  size_t n_reprimes;         /* out */
} czo_loop_opts;

/* ids[0] is BOS; codes ids[1..n_ids). out gets malloc'd payload (caller frees with czo_free).
 * returns 0, or negative on error (-1 zero-width interval, -2 symbol out of range) */
int czo_encode_tokens(czo_session *s, const uint32_t *ids, size_t n_ids, czo_loop_opts *o,
                      uint8_t **out, size_t *out_len);
/* decodes token_count symbols into ids_out[1..], ids_out[0] = bos */
int czo_decode_tokens(czo_session *s, const uint8_t *payload, size_t len, uint32_t bos, size_t token_count,
                      czo_loop_opts *o, uint32_t *ids_out);
/* src/main.rs:1725-1751 (backend 0) / 1753-1787 (backend 1) */
double czo_xe_bits_over_span(czo_session *s, int backend, const uint32_t *history, size_t nh,
                             const uint32_t *targets, size_t nt, const uint32_t *hint, size_t nhint);
void czo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
