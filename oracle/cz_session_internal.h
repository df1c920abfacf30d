/* ORACLE internal (test infrastructure only). Session vtable shared by cz_models.c / cz_rwkv7.c / cz_loop.c */
#ifndef CZ_SESSION_INTERNAL_H
#define CZ_SESSION_INTERNAL_H
#include "cz_oracle.h"
#ifdef _OPENMP
#include <omp.h>
static inline int omp_get_thread_num_safe(void) { return omp_get_thread_num(); }
static inline int omp_max_threads_safe(void) { return omp_get_max_threads(); }
#else
static inline int omp_get_thread_num_safe(void) { return 0; }
static inline int omp_max_threads_safe(void) { return 1; }
#endif

struct czo_session {
  void *impl;
  size_t vocab;
  size_t max_context_length;
  size_t index_pos;
  float *logits;
  const float *(*step)(czo_session *, uint32_t);
  const float *(*reprime)(czo_session *, const uint32_t *, size_t);
  int (*set_tensor)(czo_session *, const char *, const float *, size_t);
  void (*destroy)(czo_session *);
};
float czo_bf16_round(float x);
#endif
