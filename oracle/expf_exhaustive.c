/* ORACLE tool: prove czo_expf == host libm expf bit-for-bit over every finite float in [-104, 88]
 * (all inputs the softmax can produce are <= 0).  Exit code 0 iff zero mismatches. ~25 s on 8 cores.
 * Usage: expf_exhaustive [--neg-only]   (neg-only: x <= 0, ~3 s) */
#include "cz_oracle.h"
#include <math.h>
#include <stdio.h>
#include <string.h>
int main(int argc, char **argv) {
  int neg_only = argc > 1 && !strcmp(argv[1], "--neg-only");
  long bad = 0, n = 0;
  uint64_t lo = neg_only ? 0x80000000ull : 0, hi = 0xffffffffull;
#pragma omp parallel for reduction(+ : bad, n) schedule(static)
  for (uint64_t b = lo; b <= hi; b++) {
    uint32_t u = (uint32_t)b;
    float x;
    memcpy(&x, &u, 4);
    if (!(x <= 88.0f && x >= -104.0f)) continue;
    volatile float xv = x;
    float h = expf(xv), a = czo_expf(x);
    uint32_t hb, ab;
    memcpy(&hb, &h, 4);
    memcpy(&ab, &a, 4);
    n++;
    if (hb != ab) {
      bad++;
      if (bad < 10) printf("mismatch x=%a host=%a oracle=%a\n", x, h, a);
    }
  }
  printf("checked=%ld mismatches=%ld\n", n, bad);
  return bad != 0;
}
