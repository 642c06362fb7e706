/* knn2_abi_check — orbm_knn2 through the C ABI with the tensor-core path on (default) and off (ORBM_KNN2_TC=0): the two
 * kernels must agree on every output word. No Python, so a run costs seconds of GPU box time:
 *   gcc -O2 -o knn2_abi_check knn2_abi_check.c -I../../include -L../../orb_slam3_fast_b200 -lorbx \
 *       -Wl,-rpath,'$ORIGIN/../../orb_slam3_fast_b200'  &&  ./knn2_abi_check                                       */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "orbm.h"

static uint64_t s_ = 0x9e3779b97f4a7c15ull;
static uint64_t rnd(void) { s_ ^= s_ << 13; s_ ^= s_ >> 7; s_ ^= s_ << 17; return s_; }
static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* proto > 0: rows are one of `proto` prototypes with a few flipped bits (many exact distance ties) */
static uint8_t* descriptors(int n, int proto) {
  uint8_t* d = malloc((size_t)n * 32 + 32);
  for (size_t i = 0; i < (size_t)n * 32; i++) d[i] = (uint8_t)(rnd() >> 24);
  if (proto > 0)
    for (int r = proto; r < n; r++) {
      memcpy(d + (size_t)r * 32, d + (size_t)(rnd() % proto) * 32, 32);
      for (int k = (int)(rnd() % 4); k > 0; k--) d[(size_t)r * 32 + rnd() % 32] ^= (uint8_t)(1u << (rnd() % 8));
    }
  return d;
}

int main(void) {
  static const int cases[][3] = {{10000, 10000, 0}, {6000, 6000, 64},   {8192, 4097, 16},     {300, 120000, 0},
                                 {257, 140001, 8},  {100000, 383, 0},   {100000, 100000, 0},  {50000, 70000, 256}};
  orbm_matcher* m = NULL;
  if (orbm_create(&m, 0) != 0) return printf("orbm_create: %s\n", orbm_last_error(NULL)), 1;
  int failed = 0;
  for (unsigned c = 0; c < sizeof(cases) / sizeof(cases[0]); c++) {
    const int nq = cases[c][0], nt = cases[c][1], proto = cases[c][2];
    uint8_t *q = descriptors(nq, proto), *t = descriptors(nt, proto);
    if (proto) memcpy(q, t, (size_t)(nq < nt ? nq : nt) / 2 * 32); /* exact hits, d = 0, and duplicates among them */
    int32_t* out[2];
    double ms[2];
    for (int pass = 0; pass < 2; pass++) {
      setenv("ORBM_KNN2_TC", pass == 0 ? "1" : "0", 1);
      out[pass] = malloc((size_t)nq * 16);
      int32_t* o = out[pass];
      for (int rep = 0; rep < 2; rep++) { /* second call is warm (buffers allocated) */
        const double t0 = now_ms();
        const int rc = orbm_knn2(m, q, nq, t, nt, o, o + nq, o + 2 * (size_t)nq, o + 3 * (size_t)nq);
        ms[pass] = now_ms() - t0;
        if (rc != 0) return printf("case %u pass %d: rc %d: %s\n", c, pass, rc, orbm_last_error(m)), 1;
      }
    }
    long bad = 0;
    for (size_t i = 0; i < (size_t)nq * 4; i++) bad += out[0][i] != out[1][i];
    printf("%6d x %6d proto %3d: %ld words differ; host call %.2f ms tensor-core, %.2f ms POPC\n", nq, nt, proto, bad,
           ms[0], ms[1]);
    failed += bad != 0;
    free(q); free(t); free(out[0]); free(out[1]);
  }
  orbm_destroy(m);
  printf(failed ? "FAILED: %d cases\n" : "all cases identical\n", failed);
  return failed ? 2 : 0;
}
