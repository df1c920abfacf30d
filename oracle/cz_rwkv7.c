/*
 * cz_rwkv7.c -- ORACLE (test infrastructure only; never linked by the product).
 *
 * RWKV-7 session behind the reference's LanguageModelSession trait (src/models.rs:124-180), restating the t==1 path of
 * candle_rwkv7/src/models/rwkv7.rs (the only path CandleZip reaches: models.rs:153 feeds [[token]] and reprime replays
 * the history token by token, models.rs:167-170):
 *   Model::forward          rwkv7.rs:509-524     embed -> blocks -> ln_out -> head
 *   Block::forward          rwkv7.rs:468-478     [pre_ln on layer 0], x += att(ln1(x)), x += ffn(ln2(x))
 *   SelfAttention::forward  rwkv7.rs:179-328     token shift, r/k/v, LoRA w/a/g/v, k-hat, WKV state update, group-norm, bonus, gate
 *   FeedForward::forward    rwkv7.rs:412-430     token shift, relu(key x)^2, value
 *   group_norm              rwkv7.rs:529-542     population variance, eps 64e-5
 * Everything is f32 (models.rs:139).  candle's LayerNorm (candle-nn 0.9.1, not in tree) is restated from its documented
 * semantics: mean, centred, population variance, /sqrt(var+eps), *w + b.
 * round_bf16 = 1 rounds the operands of every dense contraction to bf16 like the B200 path (tighter comparison).
 */
#include "cz_oracle.h"
#include "cz_session_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float *pre_w, *pre_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  float *wr, *wk, *wv, *wo; /* [C][C] */
  float *gn_w, *gn_b;
  float *x_r, *x_w, *x_k, *x_v, *x_a, *x_g, *k_k, *k_a, *r_k;
  float *w1, *w2, *w0; /* [Dw][C], [C][Dw], [C] */
  float *a1, *a2, *a0;
  float *v1, *v2, *v0; /* layers > 0 */
  float *g1, *g2;
  float *ffn_xk, *ffn_key, *ffn_val; /* [C], [F][C], [C][F] */
  /* state (rwkv7.rs:35-61) */
  float *att_x_prev, *ffn_x_prev; /* [C] */
  float *att_state;               /* [H][N][N] */
} rwkv_layer;

typedef struct {
  czo_rwkv7_config c;
  float *embed, *head; /* [V][C] */
  float *lnout_w, *lnout_b;
  rwkv_layer *layers;
  /* scratch */
  float *x, *xn, *xx, *mix[6], *r, *k, *v, *vfirst, *w, *a, *g, *kk, *y, *t1, *t2, *f1, *tmp;
} rwkv_impl;

static float *fzalloc(size_t n) { return (float *)calloc(n ? n : 1, sizeof(float)); }

/* y[n] = sum_k x[k] * W[n][k]; rb: round x and W operands to bf16 first (W is expected to be bf16-exact already) */
static void matvec(const float *x, size_t kdim, const float *w, size_t n, float *y, int rb, float *tmp) {
  const float *xs = x;
  if (rb) {
    for (size_t k = 0; k < kdim; k++) tmp[k] = czo_bf16_round(x[k]);
    xs = tmp;
  }
#pragma omp parallel for schedule(static) if (n * kdim > 65536)
  for (size_t j = 0; j < n; j++) {
    const float *wr = w + j * kdim;
    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
    for (size_t k = 0; k < kdim; k++) acc += xs[k] * wr[k];
    y[j] = acc;
  }
}

static void layer_norm(const float *x, const float *w, const float *b, size_t d, float eps, float *y) {
  float mean = 0.f;
  for (size_t i = 0; i < d; i++) mean += x[i];
  mean /= (float)d;
  float var = 0.f;
  for (size_t i = 0; i < d; i++) {
    float c = x[i] - mean;
    var += c * c;
  }
  var /= (float)d;
  float inv = 1.0f / sqrtf(var + eps);
  for (size_t i = 0; i < d; i++) y[i] = (x[i] - mean) * inv * w[i] + b[i];
}

static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

static const float *rwkv_step(czo_session *s, uint32_t token) {
  rwkv_impl *m = (rwkv_impl *)s->impl;
  const czo_rwkv7_config *c = &m->c;
  const size_t C = (size_t)c->d_model, N = (size_t)c->head_dim, H = C / N, F = (size_t)c->d_ffn;
  const int rb = c->round_bf16;
  memcpy(m->x, m->embed + (size_t)token * C, C * 4); /* rwkv7.rs:511 */
  for (int l = 0; l < c->n_layers; l++) {
    rwkv_layer *L = &m->layers[l];
    if (l == 0) { /* rwkv7.rs:470-472 */
      layer_norm(m->x, L->pre_w, L->pre_b, C, c->norm_eps, m->xn);
      memcpy(m->x, m->xn, C * 4);
    }
    /* ---- attention (rwkv7.rs:179-328) ---- */
    layer_norm(m->x, L->ln1_w, L->ln1_b, C, c->norm_eps, m->xn);
    const float *mixw[6] = {L->x_r, L->x_w, L->x_k, L->x_v, L->x_a, L->x_g};
    for (size_t i = 0; i < C; i++) m->xx[i] = L->att_x_prev[i] - m->xn[i]; /* :191-192 */
    for (int q = 0; q < 6; q++)
      for (size_t i = 0; i < C; i++) m->mix[q][i] = m->xn[i] + m->xx[i] * mixw[q][i]; /* :196-201 */
    float *xr = m->mix[0], *xw = m->mix[1], *xk = m->mix[2], *xv = m->mix[3], *xa = m->mix[4], *xg = m->mix[5];
    matvec(xr, C, L->wr, C, m->r, rb, m->tmp); /* :204 */
    /* w = exp(-exp(-softplus(-(w0 + tanh(xw W1^T) W2^T)) - 0.5))   :209-221 */
    matvec(xw, C, L->w1, (size_t)c->lora_w, m->t1, rb, m->tmp);
    for (int i = 0; i < c->lora_w; i++) m->t1[i] = tanhf(m->t1[i]);
    matvec(m->t1, (size_t)c->lora_w, L->w2, C, m->w, rb, m->tmp);
    for (size_t i = 0; i < C; i++) {
      float z = m->w[i] + L->w0[i];
      float sp = logf(expf(-z) + 1.0f);
      float wt = -sp + -0.5f;
      m->w[i] = expf(-expf(wt));
    }
    matvec(xk, C, L->wk, C, m->k, rb, m->tmp); /* :223 */
    matvec(xv, C, L->wv, C, m->v, rb, m->tmp); /* :224 */
    /* a = sigmoid(a0 + (xa A1^T) A2^T)   :227-230 */
    matvec(xa, C, L->a1, (size_t)c->lora_a, m->t1, rb, m->tmp);
    matvec(m->t1, (size_t)c->lora_a, L->a2, C, m->a, rb, m->tmp);
    for (size_t i = 0; i < C; i++) m->a[i] = sigmoidf_(m->a[i] + L->a0[i]);
    /* g = sigmoid(xg G1^T) G2^T   :233-235 */
    matvec(xg, C, L->g1, (size_t)c->lora_g, m->t2, rb, m->tmp);
    for (int i = 0; i < c->lora_g; i++) m->t2[i] = sigmoidf_(m->t2[i]);
    matvec(m->t2, (size_t)c->lora_g, L->g2, C, m->g, rb, m->tmp);
    /* value residual   :238-248 */
    if (l == 0) {
      memcpy(m->vfirst, m->v, C * 4);
    } else {
      matvec(xv, C, L->v1, (size_t)c->lora_v, m->t1, rb, m->tmp);
      matvec(m->t1, (size_t)c->lora_v, L->v2, C, m->y, rb, m->tmp); /* m->y is free at this point */
      for (size_t i = 0; i < C; i++) {
        float nu = sigmoidf_(m->y[i] + L->v0[i]);
        m->v[i] = m->v[i] + (m->vfirst[i] - m->v[i]) * nu;
      }
    }
    /* kk = normalize_head(k * k_k)   :251-259 */
    for (size_t h = 0; h < H; h++) {
      float ss = 0.f;
      for (size_t j = 0; j < N; j++) {
        float q = m->k[h * N + j] * L->k_k[h * N + j];
        m->kk[h * N + j] = q;
        ss += q * q;
      }
      float nrm = sqrtf(ss);
      if (!(nrm > 1e-12f)) nrm = 1e-12f;
      for (size_t j = 0; j < N; j++) m->kk[h * N + j] /= nrm;
    }
    /* k = k * (1 + (a - 1) * k_a)   :263-265 */
    for (size_t i = 0; i < C; i++) m->k[i] = m->k[i] * (1.0f + (m->a[i] - 1.0f) * L->k_a[i]);
    /* state update and readout per head   :285-302 */
#pragma omp parallel for schedule(static)
    for (size_t h = 0; h < H; h++) {
      float *S = L->att_state + h * N * N;
      const float *wv_ = m->w + h * N, *kkv = m->kk + h * N, *av = m->a + h * N, *kv = m->k + h * N, *vv = m->v + h * N,
                  *rv = m->r + h * N;
      for (size_t i = 0; i < N; i++) {
        float *Si = S + i * N;
        float sa = 0.f; /* (S @ kk)[i] with the OLD state */
        for (size_t j = 0; j < N; j++) sa += Si[j] * kkv[j];
        float y = 0.f;
        for (size_t j = 0; j < N; j++) {
          float nv = (Si[j] * wv_[j] - sa * (kkv[j] * av[j])) + vv[i] * kv[j];
          Si[j] = nv;
          y += nv * rv[j];
        }
        m->y[h * N + i] = y;
      }
    }
    memcpy(L->att_x_prev, m->xn, C * 4); /* :306 */
    /* group norm (H groups, eps 64e-5)   :312-314, 529-542 */
    for (size_t h = 0; h < H; h++) {
      float *yh = m->y + h * N;
      float mean = 0.f;
      for (size_t j = 0; j < N; j++) mean += yh[j];
      mean /= (float)N;
      float var = 0.f;
      for (size_t j = 0; j < N; j++) {
        float cc = yh[j] - mean;
        var += cc * cc;
      }
      var /= (float)N;
      float den = sqrtf(var + (float)64e-5);
      /* bonus: (sum_j r*k*r_k) * v   :317-321 */
      float alpha = 0.f;
      for (size_t j = 0; j < N; j++) alpha += m->r[h * N + j] * m->k[h * N + j] * L->r_k[h * N + j];
      for (size_t j = 0; j < N; j++) {
        float o = ((yh[j] - mean) / den) * L->gn_w[h * N + j] + L->gn_b[h * N + j];
        o += alpha * m->v[h * N + j];
        yh[j] = o * m->g[h * N + j]; /* gate :324 */
      }
    }
    matvec(m->y, C, L->wo, C, m->xx, rb, m->tmp); /* :325 */
    for (size_t i = 0; i < C; i++) m->x[i] += m->xx[i];
    /* ---- feed forward (rwkv7.rs:412-430) ---- */
    layer_norm(m->x, L->ln2_w, L->ln2_b, C, c->norm_eps, m->xn);
    for (size_t i = 0; i < C; i++) m->xx[i] = m->xn[i] + (L->ffn_x_prev[i] - m->xn[i]) * L->ffn_xk[i];
    matvec(m->xx, C, L->ffn_key, F, m->f1, rb, m->tmp);
    for (size_t i = 0; i < F; i++) {
      float q = m->f1[i] > 0.f ? m->f1[i] : 0.f;
      m->f1[i] = q * q;
    }
    matvec(m->f1, F, L->ffn_val, C, m->xx, rb, m->tmp);
    memcpy(L->ffn_x_prev, m->xn, C * 4);
    for (size_t i = 0; i < C; i++) m->x[i] += m->xx[i];
  }
  layer_norm(m->x, m->lnout_w, m->lnout_b, C, c->norm_eps, m->xn);
  matvec(m->xn, C, m->head, (size_t)c->vocab, s->logits, rb, m->tmp);
  s->index_pos += 1;
  return s->logits;
}

static void rwkv_reset(rwkv_impl *m) { /* State::new, rwkv7.rs:47-60 */
  const size_t C = (size_t)m->c.d_model, N = (size_t)m->c.head_dim, H = C / N;
  for (int l = 0; l < m->c.n_layers; l++) {
    memset(m->layers[l].att_x_prev, 0, C * 4);
    memset(m->layers[l].ffn_x_prev, 0, C * 4);
    memset(m->layers[l].att_state, 0, H * N * N * 4);
  }
}

static const float *rwkv_reprime(czo_session *s, const uint32_t *h, size_t n) { /* src/models.rs:162-170 */
  rwkv_reset((rwkv_impl *)s->impl);
  s->index_pos = 0;
  const float *l = NULL;
  for (size_t i = 0; i < n; i++) l = rwkv_step(s, h[i]);
  return l;
}

static int rwkv_set_tensor(czo_session *s, const char *name, const float *data, size_t n) {
  rwkv_impl *m = (rwkv_impl *)s->impl;
  const czo_rwkv7_config *c = &m->c;
  const size_t C = (size_t)c->d_model, F = (size_t)c->d_ffn, V = (size_t)c->vocab;
  float *dst = NULL;
  size_t want = 0;
  int l = -1;
  char rest[128];
  if (!strcmp(name, "model.embeddings.weight")) { dst = m->embed; want = V * C; }
  else if (!strcmp(name, "lm_head.weight")) { dst = m->head; want = V * C; }
  else if (!strcmp(name, "model.norm.weight")) { dst = m->lnout_w; want = C; }
  else if (!strcmp(name, "model.norm.bias")) { dst = m->lnout_b; want = C; }
  else if (sscanf(name, "model.layers.%d.%127s", &l, rest) == 2 && l >= 0 && l < c->n_layers) {
    rwkv_layer *L = &m->layers[l];
#define T_(nm, ptr, sz) else if (!strcmp(rest, nm)) { dst = ptr; want = sz; }
    if (0) {}
    T_("pre_norm.weight", L->pre_w, C) T_("pre_norm.bias", L->pre_b, C)
    T_("attn_norm.weight", L->ln1_w, C) T_("attn_norm.bias", L->ln1_b, C)
    T_("ffn_norm.weight", L->ln2_w, C) T_("ffn_norm.bias", L->ln2_b, C)
    T_("attn.r_proj.weight", L->wr, C * C) T_("attn.k_proj.weight", L->wk, C * C)
    T_("attn.v_proj.weight", L->wv, C * C) T_("attn.o_proj.weight", L->wo, C * C)
    T_("attn.g_norm.weight", L->gn_w, C) T_("attn.g_norm.bias", L->gn_b, C)
    T_("attn.x_r", L->x_r, C) T_("attn.x_w", L->x_w, C) T_("attn.x_k", L->x_k, C)
    T_("attn.x_v", L->x_v, C) T_("attn.x_a", L->x_a, C) T_("attn.x_g", L->x_g, C)
    T_("attn.k_k", L->k_k, C) T_("attn.k_a", L->k_a, C) T_("attn.r_k", L->r_k, C)
    T_("attn.w_lora.lora.0.weight", L->w1, (size_t)c->lora_w * C) T_("attn.w_lora.lora.2.weight", L->w2, C * c->lora_w)
    T_("attn.w_lora.lora.2.bias", L->w0, C)
    T_("attn.a_lora.lora.0.weight", L->a1, (size_t)c->lora_a * C) T_("attn.a_lora.lora.2.weight", L->a2, C * c->lora_a)
    T_("attn.a_lora.lora.2.bias", L->a0, C)
    T_("attn.v_lora.lora.0.weight", L->v1, (size_t)c->lora_v * C) T_("attn.v_lora.lora.2.weight", L->v2, C * c->lora_v)
    T_("attn.v_lora.lora.2.bias", L->v0, C)
    T_("attn.g_lora.lora.0.weight", L->g1, (size_t)c->lora_g * C) T_("attn.g_lora.lora.2.weight", L->g2, C * c->lora_g)
    T_("ffn.x_k", L->ffn_xk, C) T_("ffn.key.weight", L->ffn_key, F * C) T_("ffn.value.weight", L->ffn_val, C * F)
#undef T_
  }
  if (!dst || want != n) return -1;
  memcpy(dst, data, n * 4);
  return 0;
}

static void rwkv_destroy(czo_session *s) {
  rwkv_impl *m = (rwkv_impl *)s->impl;
  for (int l = 0; l < m->c.n_layers; l++) {
    rwkv_layer *L = &m->layers[l];
    float *ptrs[] = {L->pre_w, L->pre_b, L->ln1_w, L->ln1_b, L->ln2_w, L->ln2_b, L->wr, L->wk, L->wv, L->wo, L->gn_w, L->gn_b,
                     L->x_r, L->x_w, L->x_k, L->x_v, L->x_a, L->x_g, L->k_k, L->k_a, L->r_k, L->w1, L->w2, L->w0, L->a1, L->a2,
                     L->a0, L->v1, L->v2, L->v0, L->g1, L->g2, L->ffn_xk, L->ffn_key, L->ffn_val, L->att_x_prev, L->ffn_x_prev,
                     L->att_state};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
  }
  free(m->layers);
  float *ptrs[] = {m->embed, m->head, m->lnout_w, m->lnout_b, m->x, m->xn, m->xx, m->mix[0], m->mix[1], m->mix[2], m->mix[3],
                   m->mix[4], m->mix[5], m->r, m->k, m->v, m->vfirst, m->w, m->a, m->g, m->kk, m->y, m->t1, m->t2, m->f1, m->tmp};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
  free(m);
}

czo_session *czo_rwkv7_new(const czo_rwkv7_config *cfg) {
  if (cfg->head_dim <= 0 || cfg->d_model % cfg->head_dim) return NULL;
  czo_session *s = (czo_session *)calloc(1, sizeof(*s));
  rwkv_impl *m = (rwkv_impl *)calloc(1, sizeof(*m));
  m->c = *cfg;
  const size_t C = (size_t)cfg->d_model, N = (size_t)cfg->head_dim, H = C / N, F = (size_t)cfg->d_ffn, V = (size_t)cfg->vocab;
  m->embed = fzalloc(V * C);
  m->head = fzalloc(V * C);
  m->lnout_w = fzalloc(C);
  m->lnout_b = fzalloc(C);
  m->layers = (rwkv_layer *)calloc((size_t)cfg->n_layers, sizeof(rwkv_layer));
  for (int l = 0; l < cfg->n_layers; l++) {
    rwkv_layer *L = &m->layers[l];
    L->pre_w = fzalloc(C); L->pre_b = fzalloc(C); L->ln1_w = fzalloc(C); L->ln1_b = fzalloc(C); L->ln2_w = fzalloc(C); L->ln2_b = fzalloc(C);
    L->wr = fzalloc(C * C); L->wk = fzalloc(C * C); L->wv = fzalloc(C * C); L->wo = fzalloc(C * C);
    L->gn_w = fzalloc(C); L->gn_b = fzalloc(C);
    L->x_r = fzalloc(C); L->x_w = fzalloc(C); L->x_k = fzalloc(C); L->x_v = fzalloc(C); L->x_a = fzalloc(C); L->x_g = fzalloc(C);
    L->k_k = fzalloc(C); L->k_a = fzalloc(C); L->r_k = fzalloc(C);
    L->w1 = fzalloc((size_t)cfg->lora_w * C); L->w2 = fzalloc(C * cfg->lora_w); L->w0 = fzalloc(C);
    L->a1 = fzalloc((size_t)cfg->lora_a * C); L->a2 = fzalloc(C * cfg->lora_a); L->a0 = fzalloc(C);
    L->v1 = fzalloc((size_t)cfg->lora_v * C); L->v2 = fzalloc(C * cfg->lora_v); L->v0 = fzalloc(C);
    L->g1 = fzalloc((size_t)cfg->lora_g * C); L->g2 = fzalloc(C * cfg->lora_g);
    L->ffn_xk = fzalloc(C); L->ffn_key = fzalloc(F * C); L->ffn_val = fzalloc(C * F);
    L->att_x_prev = fzalloc(C); L->ffn_x_prev = fzalloc(C); L->att_state = fzalloc(H * N * N);
  }
  size_t big = F > C ? F : C;
  m->x = fzalloc(C); m->xn = fzalloc(C); m->xx = fzalloc(C);
  for (int q = 0; q < 6; q++) m->mix[q] = fzalloc(C);
  m->r = fzalloc(C); m->k = fzalloc(C); m->v = fzalloc(C); m->vfirst = fzalloc(C); m->w = fzalloc(C); m->a = fzalloc(C);
  m->g = fzalloc(C); m->kk = fzalloc(C); m->y = fzalloc(C);
  m->t1 = fzalloc(C); m->t2 = fzalloc(C); m->f1 = fzalloc(F); m->tmp = fzalloc(big);
  s->impl = m;
  s->vocab = V;
  s->max_context_length = (size_t)-1; /* usize::MAX, src/models.rs:150 */
  s->logits = (float *)malloc(V * 4);
  s->step = rwkv_step;
  s->reprime = rwkv_reprime;
  s->set_tensor = rwkv_set_tensor;
  s->destroy = rwkv_destroy;
  return s;
}
