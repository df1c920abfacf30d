/*
 * cz_models.c -- ORACLE (test infrastructure only; never linked by the product).
 *
 * Model sessions behind the reference's LanguageModelSession trait (src/models.rs:28-33):
 *   - SmolLM / LLaMA (src/models.rs:37-120).  The forward itself lives in candle-transformers 0.9.1
 *     `models::llama` (Cargo.toml:33-35), which is NOT vendored under /root/reference; it is restated
 *     here from the published LLaMA architecture as implemented by HF `LlamaForCausalLM`
 *     (RMSNorm eps, rotate-half RoPE, GQA repeat, f32 softmax, SwiGLU, tied head, last-position logits)
 *     and cross-checked against transformers in tests/golden/make_llama_golden.py.
 *   - RWKV-7 (src/models.rs:124-180 -> candle_rwkv7/src/models/rwkv7.rs t==1 path): see cz_rwkv7.c.
 *   - a table-driven fake session for exercising the loops without weights.
 */
#include "cz_oracle.h"
#include "cz_session_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------------- shared helpers ---------------- */
float czo_bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x; /* inf / nan unchanged */
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  u &= 0xffff0000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

void czo_session_free(czo_session *s) {
  if (!s) return;
  if (s->destroy) s->destroy(s);
  free(s->logits);
  free(s);
}
size_t czo_session_vocab_size(const czo_session *s) { return s->vocab; }
size_t czo_session_max_context_length(const czo_session *s) { return s->max_context_length; }
size_t czo_session_index_pos(const czo_session *s) { return s->index_pos; }
const float *czo_session_step_logits(czo_session *s, uint32_t token) { return s->step(s, token); }
const float *czo_session_reprime(czo_session *s, const uint32_t *history, size_t n) {
  if (n == 0) return NULL; /* src/models.rs:108 bails */
  return s->reprime(s, history, n);
}
int czo_session_set_tensor(czo_session *s, const char *name, const float *data, size_t n) {
  if (!s->set_tensor) return -1;
  return s->set_tensor(s, name, data, n);
}

/* ---------------- table session (test hook) ---------------- */
typedef struct {
  const float *table;
  size_t n_rows;
} table_impl;

static const float *table_row(czo_session *s, uint32_t token) {
  table_impl *t = (table_impl *)s->impl;
  uint64_t h = (uint64_t)s->index_pos * 0x9E3779B97F4A7C15ull + (uint64_t)token * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  const float *row = t->table + (size_t)(h % t->n_rows) * s->vocab;
  memcpy(s->logits, row, s->vocab * sizeof(float));
  return s->logits;
}
static const float *table_step(czo_session *s, uint32_t token) {
  s->index_pos += 1;
  return table_row(s, token);
}
static const float *table_reprime(czo_session *s, const uint32_t *h, size_t n) {
  s->index_pos = n;
  return table_row(s, h[n - 1]);
}
static void table_destroy(czo_session *s) { free(s->impl); }

czo_session *czo_table_session_new(size_t vocab, const float *table, size_t n_rows) {
  czo_session *s = (czo_session *)calloc(1, sizeof(*s));
  table_impl *t = (table_impl *)calloc(1, sizeof(*t));
  t->table = table;
  t->n_rows = n_rows;
  s->impl = t;
  s->vocab = vocab;
  s->max_context_length = 512;
  s->logits = (float *)malloc(vocab * sizeof(float));
  s->step = table_step;
  s->reprime = table_reprime;
  s->destroy = table_destroy;
  return s;
}

/* ---------------- LLaMA / SmolLM ---------------- */
typedef struct {
  float *attn_norm, *wq, *wk, *wv, *wo, *ffn_norm, *wg, *wu, *wd;
  float *kcache, *vcache; /* [max_pos][n_kv*hd] */
} llama_layer;

typedef struct {
  czo_llama_config c;
  float *embed;      /* [V][D] (tied head) */
  float *final_norm; /* [D] */
  llama_layer *layers;
  float *inv_freq;   /* [hd/2] */
  /* scratch for up to T tokens */
  size_t cap_t;
  float *x, *h, *q, *k, *v, *att, *g, *u, *scores;
} llama_impl;

static void llama_scratch(llama_impl *m, size_t t) {
  if (t <= m->cap_t) return;
  const czo_llama_config *c = &m->c;
  size_t d = (size_t)c->d_model, kv = (size_t)c->n_kv_heads * c->head_dim, f = (size_t)c->d_ffn;
  free(m->x); free(m->h); free(m->q); free(m->k); free(m->v); free(m->att); free(m->g); free(m->u);
  m->x = (float *)malloc(t * d * 4);
  m->h = (float *)malloc(t * d * 4);
  m->q = (float *)malloc(t * d * 4);
  m->k = (float *)malloc(t * kv * 4);
  m->v = (float *)malloc(t * kv * 4);
  m->att = (float *)malloc(t * d * 4);
  m->g = (float *)malloc(t * f * 4);
  m->u = (float *)malloc(t * f * 4);
  m->cap_t = t;
}

/* Y[t][n] = sum_k X[t][k] * W[n][k]   (Linear without bias, weight [out,in]) */
static void linear(const float *x, size_t t, size_t kdim, const float *w, size_t n, float *y) {
#pragma omp parallel for schedule(static)
  for (size_t j = 0; j < n; j++) {
    const float *wr = w + j * kdim;
    for (size_t i = 0; i < t; i++) {
      const float *xr = x + i * kdim;
      float acc = 0.f;
#pragma omp simd reduction(+ : acc)
      for (size_t k = 0; k < kdim; k++) acc += xr[k] * wr[k];
      y[i * n + j] = acc;
    }
  }
}

static void rmsnorm(const float *x, const float *w, size_t t, size_t d, float eps, int rb, float *y) {
  for (size_t i = 0; i < t; i++) {
    const float *xr = x + i * d;
    float ss = 0.f;
    for (size_t k = 0; k < d; k++) ss += xr[k] * xr[k];
    float inv = 1.0f / sqrtf(ss / (float)d + eps);
    for (size_t k = 0; k < d; k++) {
      float o = xr[k] * inv * w[k];
      y[i * d + k] = rb ? czo_bf16_round(o) : o;
    }
  }
}

/* forward of t tokens starting at position pos0; appends to the KV cache; logits of the LAST token only */
static const float *llama_forward(czo_session *s, const uint32_t *tok, size_t t, size_t pos0) {
  llama_impl *m = (llama_impl *)s->impl;
  const czo_llama_config *c = &m->c;
  const size_t d = (size_t)c->d_model, hd = (size_t)c->head_dim, nh = (size_t)c->n_heads, nkv = (size_t)c->n_kv_heads;
  const size_t kvd = nkv * hd, f = (size_t)c->d_ffn, half = hd / 2, grp = nh / nkv;
  const int rb = c->round_bf16;
  if (pos0 + t > (size_t)c->max_pos) {
    fprintf(stderr, "czo llama: KV capacity exceeded (%zu + %zu > %d)\n", pos0, t, c->max_pos);
    abort();
  }
  llama_scratch(m, t);
  for (size_t i = 0; i < t; i++) memcpy(m->x + i * d, m->embed + (size_t)tok[i] * d, d * 4);
  const float scale = 1.0f / sqrtf((float)hd);
  for (int l = 0; l < c->n_layers; l++) {
    llama_layer *L = &m->layers[l];
    rmsnorm(m->x, L->attn_norm, t, d, c->rms_eps, rb, m->h);
    linear(m->h, t, d, L->wq, d, m->q);
    linear(m->h, t, d, L->wk, kvd, m->k);
    linear(m->h, t, d, L->wv, kvd, m->v);
    /* rotate-half RoPE (non-interleaved), f32 */
    for (size_t i = 0; i < t; i++) {
      float pos = (float)(pos0 + i);
      for (size_t hh = 0; hh < nh + nkv; hh++) {
        float *vec = hh < nh ? m->q + i * d + hh * hd : m->k + i * kvd + (hh - nh) * hd;
        for (size_t j = 0; j < half; j++) {
          float ang = pos * m->inv_freq[j];
          float cs = cosf(ang), sn = sinf(ang);
          float a = vec[j], b = vec[j + half];
          vec[j] = a * cs - b * sn;
          vec[j + half] = b * cs + a * sn;
        }
      }
    }
    if (rb) {
      for (size_t i = 0; i < t * d; i++) m->q[i] = czo_bf16_round(m->q[i]);
      for (size_t i = 0; i < t * kvd; i++) m->k[i] = czo_bf16_round(m->k[i]);
      for (size_t i = 0; i < t * kvd; i++) m->v[i] = czo_bf16_round(m->v[i]);
    }
    memcpy(L->kcache + pos0 * kvd, m->k, t * kvd * 4);
    memcpy(L->vcache + pos0 * kvd, m->v, t * kvd * 4);
    /* causal attention with f32 softmax */
#pragma omp parallel for collapse(2) schedule(static)
    for (size_t i = 0; i < t; i++) {
      for (size_t hh = 0; hh < nh; hh++) {
        size_t kvh = hh / grp, n_keys = pos0 + i + 1;
        float *sc = m->scores + ((size_t)omp_get_thread_num_safe()) * (size_t)c->max_pos;
        const float *qv = m->q + i * d + hh * hd;
        float mx = -INFINITY;
        for (size_t p = 0; p < n_keys; p++) {
          const float *kv = L->kcache + p * kvd + kvh * hd;
          float acc = 0.f;
#pragma omp simd reduction(+ : acc)
          for (size_t j = 0; j < hd; j++) acc += qv[j] * kv[j];
          acc *= scale;
          sc[p] = acc;
          if (acc > mx) mx = acc;
        }
        float sum = 0.f;
        for (size_t p = 0; p < n_keys; p++) {
          sc[p] = expf(sc[p] - mx);
          sum += sc[p];
        }
        float *o = m->att + i * d + hh * hd;
        for (size_t j = 0; j < hd; j++) o[j] = 0.f;
        for (size_t p = 0; p < n_keys; p++) {
          const float *vv = L->vcache + p * kvd + kvh * hd;
          float pw = sc[p];
          for (size_t j = 0; j < hd; j++) o[j] += pw * vv[j];
        }
        float inv = 1.0f / sum;
        for (size_t j = 0; j < hd; j++) {
          float val = o[j] * inv;
          o[j] = rb ? czo_bf16_round(val) : val;
        }
      }
    }
    linear(m->att, t, d, L->wo, d, m->h);
    for (size_t i = 0; i < t * d; i++) m->x[i] += m->h[i];
    rmsnorm(m->x, L->ffn_norm, t, d, c->rms_eps, rb, m->h);
    linear(m->h, t, d, L->wg, f, m->g);
    linear(m->h, t, d, L->wu, f, m->u);
    for (size_t i = 0; i < t * f; i++) {
      float gv = m->g[i];
      float a = gv / (1.0f + expf(-gv)) * m->u[i];
      m->g[i] = rb ? czo_bf16_round(a) : a;
    }
    linear(m->g, t, f, L->wd, d, m->h);
    for (size_t i = 0; i < t * d; i++) m->x[i] += m->h[i];
  }
  rmsnorm(m->x + (t - 1) * d, m->final_norm, 1, d, c->rms_eps, rb, m->h);
  linear(m->h, 1, d, m->embed, (size_t)c->vocab, s->logits);
  return s->logits;
}

static const float *llama_step(czo_session *s, uint32_t token) { /* src/models.rs:92-103 */
  const float *l = llama_forward(s, &token, 1, s->index_pos);
  s->index_pos += 1;
  return l;
}
static const float *llama_reprime(czo_session *s, const uint32_t *h, size_t n) { /* src/models.rs:104-119 */
  s->index_pos = 0; /* fresh cache */
  const float *l = llama_forward(s, h, n, 0);
  s->index_pos += n;
  return l;
}

static int llama_set_tensor(czo_session *s, const char *name, const float *data, size_t n) {
  llama_impl *m = (llama_impl *)s->impl;
  const czo_llama_config *c = &m->c;
  size_t d = (size_t)c->d_model, kvd = (size_t)c->n_kv_heads * c->head_dim, f = (size_t)c->d_ffn;
  float *dst = NULL;
  size_t want = 0;
  int l = -1;
  char rest[128];
  if (!strcmp(name, "model.embed_tokens.weight") || !strcmp(name, "lm_head.weight")) { dst = m->embed; want = (size_t)c->vocab * d; }
  else if (!strcmp(name, "model.norm.weight")) { dst = m->final_norm; want = d; }
  else if (sscanf(name, "model.layers.%d.%127s", &l, rest) == 2 && l >= 0 && l < c->n_layers) {
    llama_layer *L = &m->layers[l];
    if (!strcmp(rest, "input_layernorm.weight")) { dst = L->attn_norm; want = d; }
    else if (!strcmp(rest, "post_attention_layernorm.weight")) { dst = L->ffn_norm; want = d; }
    else if (!strcmp(rest, "self_attn.q_proj.weight")) { dst = L->wq; want = d * d; }
    else if (!strcmp(rest, "self_attn.k_proj.weight")) { dst = L->wk; want = kvd * d; }
    else if (!strcmp(rest, "self_attn.v_proj.weight")) { dst = L->wv; want = kvd * d; }
    else if (!strcmp(rest, "self_attn.o_proj.weight")) { dst = L->wo; want = d * d; }
    else if (!strcmp(rest, "mlp.gate_proj.weight")) { dst = L->wg; want = f * d; }
    else if (!strcmp(rest, "mlp.up_proj.weight")) { dst = L->wu; want = f * d; }
    else if (!strcmp(rest, "mlp.down_proj.weight")) { dst = L->wd; want = d * f; }
  }
  if (!dst || want != n) return -1;
  memcpy(dst, data, n * 4);
  return 0;
}

static void llama_destroy(czo_session *s) {
  llama_impl *m = (llama_impl *)s->impl;
  for (int l = 0; l < m->c.n_layers; l++) {
    llama_layer *L = &m->layers[l];
    free(L->attn_norm); free(L->wq); free(L->wk); free(L->wv); free(L->wo);
    free(L->ffn_norm); free(L->wg); free(L->wu); free(L->wd); free(L->kcache); free(L->vcache);
  }
  free(m->layers); free(m->embed); free(m->final_norm); free(m->inv_freq);
  free(m->x); free(m->h); free(m->q); free(m->k); free(m->v); free(m->att); free(m->g); free(m->u); free(m->scores);
  free(m);
}

czo_session *czo_llama_new(const czo_llama_config *cfg) {
  if (cfg->n_heads * cfg->head_dim != cfg->d_model || cfg->n_heads % cfg->n_kv_heads) return NULL;
  czo_session *s = (czo_session *)calloc(1, sizeof(*s));
  llama_impl *m = (llama_impl *)calloc(1, sizeof(*m));
  m->c = *cfg;
  size_t d = (size_t)cfg->d_model, kvd = (size_t)cfg->n_kv_heads * cfg->head_dim, f = (size_t)cfg->d_ffn;
  m->embed = (float *)calloc((size_t)cfg->vocab * d, 4);
  m->final_norm = (float *)calloc(d, 4);
  m->layers = (llama_layer *)calloc((size_t)cfg->n_layers, sizeof(llama_layer));
  for (int l = 0; l < cfg->n_layers; l++) {
    llama_layer *L = &m->layers[l];
    L->attn_norm = (float *)calloc(d, 4);
    L->ffn_norm = (float *)calloc(d, 4);
    L->wq = (float *)calloc(d * d, 4);
    L->wk = (float *)calloc(kvd * d, 4);
    L->wv = (float *)calloc(kvd * d, 4);
    L->wo = (float *)calloc(d * d, 4);
    L->wg = (float *)calloc(f * d, 4);
    L->wu = (float *)calloc(f * d, 4);
    L->wd = (float *)calloc(d * f, 4);
    L->kcache = (float *)calloc((size_t)cfg->max_pos * kvd, 4);
    L->vcache = (float *)calloc((size_t)cfg->max_pos * kvd, 4);
  }
  size_t half = (size_t)cfg->head_dim / 2;
  m->inv_freq = (float *)malloc(half * 4);
  for (size_t j = 0; j < half; j++)
    m->inv_freq[j] = 1.0f / powf(cfg->rope_theta, (float)(2 * j) / (float)cfg->head_dim);
  m->scores = (float *)malloc((size_t)omp_max_threads_safe() * (size_t)cfg->max_pos * 4);
  s->impl = m;
  s->vocab = (size_t)cfg->vocab;
  s->max_context_length = 512; /* src/models.rs:91 */
  s->logits = (float *)malloc((size_t)cfg->vocab * 4);
  s->step = llama_step;
  s->reprime = llama_reprime;
  s->set_tensor = llama_set_tensor;
  s->destroy = llama_destroy;
  return s;
}
