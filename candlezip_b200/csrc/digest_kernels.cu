// f-4: watchdog-compatible logits digests on the GPU.  The reference's self-test can record, per coded token, the BLAKE3-128 of
// the logits vector it coded from (blake3_f32_bin16: the f32 little-endian bytes of the V logits, src/main.rs:955-961; enabled by
// CANDLEZIP_WATCHDOG_DIGEST, :1085-1090; written per step at :2328-2342 and compared between encode and decode).  Here the logits
// never leave HBM, so the digest is computed next to the CDF pass from the same vocab-major logits batch [V][ld]:
//   pass 1: one thread per (column, 1 KiB chunk) compresses the chunk's 16 blocks -> chaining value (BLAKE3 chunk state)
//   pass 2: one thread per column merges its chunk CVs with the spec's stack algorithm (left subtrees are complete binary
//           trees; the last parent carries the ROOT flag) and writes the first 16 bytes of the root output.
// Integer-only, order fixed by the hash: identical digests for every wave size, batch size and GPU count -- a 16-byte
// cross-run determinism audit per token (tests: digest == blake3 package on the same logits; encode digests == decode digests).
#include "cz_common.cuh"
#include "model.h"

namespace czk {

__constant__ uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 };

__device__ __forceinline__ uint32_t b3_rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
#define B3_G(a, b, c, d, mx, my) \
  do {                           \
    a = a + b + (mx);            \
    d = b3_rotr(d ^ a, 16);      \
    c = c + d;                   \
    b = b3_rotr(b ^ c, 12);      \
    a = a + b + (my);            \
    d = b3_rotr(d ^ a, 8);       \
    c = c + d;                   \
    b = b3_rotr(b ^ c, 7);       \
  } while (0)

// message word order of round r = the permutation applied r times (fully unrolled: the words never move between registers)
__device__ constexpr int B3_SCHED[7][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},  {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8},
    {3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1},  {10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6},
    {12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4},  {9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7},
    {11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13}};

// out[0..8) = the new chaining value (first half of the compression output)
__device__ __forceinline__ void b3_compress(const uint32_t (&cv)[8], const uint32_t (&m)[16], uint32_t counter_lo, uint32_t counter_hi,
                                            uint32_t block_len, uint32_t flags, uint32_t (&out)[8]) {
  uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
  uint32_t s8 = B3_IV[0], s9 = B3_IV[1], s10 = B3_IV[2], s11 = B3_IV[3], s12 = counter_lo, s13 = counter_hi, s14 = block_len, s15 = flags;
#pragma unroll
  for (int r = 0; r < 7; r++) {
    B3_G(s0, s4, s8, s12, m[B3_SCHED[r][0]], m[B3_SCHED[r][1]]);
    B3_G(s1, s5, s9, s13, m[B3_SCHED[r][2]], m[B3_SCHED[r][3]]);
    B3_G(s2, s6, s10, s14, m[B3_SCHED[r][4]], m[B3_SCHED[r][5]]);
    B3_G(s3, s7, s11, s15, m[B3_SCHED[r][6]], m[B3_SCHED[r][7]]);
    B3_G(s0, s5, s10, s15, m[B3_SCHED[r][8]], m[B3_SCHED[r][9]]);
    B3_G(s1, s6, s11, s12, m[B3_SCHED[r][10]], m[B3_SCHED[r][11]]);
    B3_G(s2, s7, s8, s13, m[B3_SCHED[r][12]], m[B3_SCHED[r][13]]);
    B3_G(s3, s4, s9, s14, m[B3_SCHED[r][14]], m[B3_SCHED[r][15]]);
  }
  out[0] = s0 ^ s8;
  out[1] = s1 ^ s9;
  out[2] = s2 ^ s10;
  out[3] = s3 ^ s11;
  out[4] = s4 ^ s12;
  out[5] = s5 ^ s13;
  out[6] = s6 ^ s14;
  out[7] = s7 ^ s15;
}

// pass 1.  logits [V][ld] vocab-major: word i of column `col` is the bit pattern of logits[i * ld + col] (f32 little-endian
// bytes == the u32 in message-word order).  cv [(chunk * 8 + k) * cols_pad + col].  A single-chunk vector (V <= 256) is the root.
__global__ void __launch_bounds__(128) b3_chunk_cv_kernel(const float *__restrict__ logits, size_t ld, int V, int col0, int n_cols, int n_chunks,
                                                          uint32_t *__restrict__ cv, size_t cols_pad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  if (c >= n_cols) return;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(logits) + (size_t)(col0 + c);
  const int w0 = chunk * 256, n_words = min(256, V - w0);  // words of this chunk (>= 1)
  const int n_blocks = (n_words + 15) >> 4;
  uint32_t h[8];
#pragma unroll
  for (int k = 0; k < 8; k++) h[k] = B3_IV[k];
  for (int b = 0; b < n_blocks; b++) {
    uint32_t m[16];
    const int bw = min(16, n_words - b * 16);
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = i < bw ? src[(size_t)(w0 + b * 16 + i) * ld] : 0u;
    uint32_t flags = (b == 0 ? B3_CHUNK_START : 0) | (b + 1 == n_blocks ? B3_CHUNK_END : 0);
    const bool root = n_chunks == 1 && b + 1 == n_blocks;
    if (root) flags |= B3_ROOT;
    uint32_t o[8];
    b3_compress(h, m, root ? 0u : (uint32_t)chunk, 0u, (uint32_t)(bw * 4), flags, o);
#pragma unroll
    for (int k = 0; k < 8; k++) h[k] = o[k];
  }
#pragma unroll
  for (int k = 0; k < 8; k++) cv[((size_t)chunk * 8 + k) * cols_pad + c] = h[k];
}

// pass 2.  Destination of column c: digest index out_index[col0 + c] (RWKV-7: columns are scattered, literal escapes do not step the
// model), else seg_start[col0 + c] + ctr[0], else out_first + c; stepwise decoding passes
// the segment table and the device-resident step counter so that one captured graph serves every step, and idles finished lanes.
__global__ void __launch_bounds__(128) b3_merge_kernel(const uint32_t *__restrict__ cv, size_t cols_pad, int col0, int n_cols, int n_chunks,
                                                       uint8_t *__restrict__ out, unsigned long long out_first,
                                                       const unsigned long long *__restrict__ out_index,
                                                       const uint64_t *__restrict__ seg_start, const unsigned long long *__restrict__ ctr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  unsigned long long dst = out_index ? out_index[col0 + c] : out_first + (unsigned long long)c;
  if (seg_start) {
    const unsigned long long i = ctr[0], a = seg_start[col0 + c], b = seg_start[col0 + c + 1];
    if (i >= b - a) return;
    dst = a + i;
  }
  uint32_t cur[8];
  auto load = [&](int chunk, uint32_t(&v)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = cv[((size_t)chunk * 8 + k) * cols_pad + c];
  };
  if (n_chunks > 1) {
    uint32_t stack[24][8];  // one entry per set bit of the chunk count: 2^24 chunks = 16 GiB of logits per column
    int sp = 0;
    uint32_t iv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) iv[k] = B3_IV[k];
    for (int ch = 0; ch + 1 < n_chunks; ch++) {
      load(ch, cur);
      unsigned total = (unsigned)ch + 1u;
      while ((total & 1u) == 0u) {  // completes a subtree: parent(left = top of the stack, right = cur)
        uint32_t m[16], o[8];
        sp--;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          m[k] = stack[sp][k];
          m[8 + k] = cur[k];
        }
        b3_compress(iv, m, 0u, 0u, 64u, B3_PARENT, o);
#pragma unroll
        for (int k = 0; k < 8; k++) cur[k] = o[k];
        total >>= 1;
      }
#pragma unroll
      for (int k = 0; k < 8; k++) stack[sp][k] = cur[k];
      sp++;
    }
    load(n_chunks - 1, cur);
    while (sp > 0) {
      uint32_t m[16], o[8];
      sp--;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        m[k] = stack[sp][k];
        m[8 + k] = cur[k];
      }
      b3_compress(iv, m, 0u, 0u, 64u, B3_PARENT | (sp == 0 ? B3_ROOT : 0), o);
#pragma unroll
      for (int k = 0; k < 8; k++) cur[k] = o[k];
    }
  } else {
    load(0, cur);
  }
  uint4 *d = reinterpret_cast<uint4 *>(out + dst * 16);
  *d = make_uint4(cur[0], cur[1], cur[2], cur[3]);
}

}  // namespace czk

namespace cz {

size_t digest_scratch_bytes(size_t V, size_t n_cols) {
  const size_t n_chunks = (V + 255) / 256, cols = n_cols < 16384 ? n_cols : 16384;
  return n_chunks * 8 * ((cols + 31) & ~(size_t)31) * 4;
}

// digests of columns [0, n_cols) of logits [V][ld] -> out[(index) * 16]; scratch >= digest_scratch_bytes(V, n_cols)
int launch_logits_digest(cz_ctx *ctx, const float *logits, size_t V, size_t n_cols, size_t ld, void *scratch, uint8_t *out,
                         unsigned long long out_first, const unsigned long long *out_index, const uint64_t *seg_start,
                         const unsigned long long *ctr, cudaStream_t st) {
  if (n_cols == 0) return CZ_OK;
  const int n_chunks = (int)((V + 255) / 256);
  if (n_chunks > 65535) {
    set_error("logits digest: vocabulary too large");
    return CZ_ERR_UNSUPPORTED;
  }
  for (size_t c0 = 0; c0 < n_cols; c0 += 16384) {
    const size_t nc = n_cols - c0 < 16384 ? n_cols - c0 : 16384, pad = (nc + 31) & ~(size_t)31;
    dim3 grid((unsigned)ceil_div(nc, 128), (unsigned)n_chunks);
    CZ_LAUNCH(ctx, CZ_K_OTHER,
              (czk::b3_chunk_cv_kernel<<<grid, 128, 0, st>>>(logits, ld, (int)V, (int)c0, (int)nc, n_chunks, (uint32_t *)scratch, pad)));
    CZ_CHECK_LAUNCH();
    CZ_LAUNCH(ctx, CZ_K_OTHER,
              (czk::b3_merge_kernel<<<(unsigned)ceil_div(nc, 128), 128, 0, st>>>((const uint32_t *)scratch, pad, (int)c0, (int)nc, n_chunks, out,
                                                                                out_first + c0, out_index, seg_start, ctr)));
    CZ_CHECK_LAUNCH();
  }
  return CZ_OK;
}

int digest_begin(cz_model *m, size_t n_tokens, size_t max_cols, cudaStream_t st) {
  if (!m->digest_host) return CZ_OK;
  if (n_tokens > m->digest_cap) {
    set_error("digest sink too small: " + std::to_string(m->digest_cap) + " < " + std::to_string(n_tokens) + " tokens");
    return CZ_ERR_NOMEM;
  }
  CZ_TRY(m->sb[SB_DIGEST].reserve(n_tokens * 16 + 16, st));
  CZ_TRY(m->sb[SB_CV].reserve(digest_scratch_bytes((size_t)m->cfg.vocab, max_cols) + 256, st));
  return CZ_OK;
}
int digest_end(cz_model *m, size_t n_tokens, cudaStream_t st) {
  if (!m->digest_host || n_tokens == 0) return CZ_OK;
  CZ_CUDA_TRY(cudaMemcpyAsync(m->digest_host, m->sb[SB_DIGEST].p, n_tokens * 16, cudaMemcpyDeviceToHost, st));
  return CZ_OK;
}

}  // namespace cz

extern "C" int cz_model_set_digest_out(cz_model *m, uint8_t *digests_out, size_t cap_tokens) {
  if (!m || (digests_out && cap_tokens == 0)) return CZ_ERR_INVALID;
  m->digest_host = digests_out;
  m->digest_cap = digests_out ? cap_tokens : 0;
  return CZ_OK;
}
