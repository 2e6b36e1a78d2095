/*
 * exhaustive_trig.c -- pins oracle_sinf/oracle_cosf (the restatement of glibc
 * 2.39 sinf/cosf) against the libm this process is linked with, bit for bit,
 * over EVERY binary32 value with |x| <= limit (default 128: covers the < pi/4,
 * reduce_fast and the start of the reduce_large ranges), plus a strided sweep
 * of the remaining finite range.  TEST INFRASTRUCTURE ONLY.
 *
 * usage: exhaustive_trig [limit] [threads]     exit 0 iff zero mismatches
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mgym_oracle.h"

typedef struct {
  uint32_t lo, hi, stride;
  uint64_t checked, bad_sin, bad_cos;
  uint32_t first_bad;
} job_t;

static inline float from_bits(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint32_t to_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

static void *run(void *arg) {
  job_t *j = (job_t *)arg;
  for (uint64_t b = j->lo; b < j->hi; b += j->stride) {
    for (uint32_t sign = 0; sign < 2; ++sign) {
      float x = from_bits((uint32_t)b | (sign << 31));
      uint32_t s0 = to_bits(sinf(x)), s1 = to_bits(oracle_sinf(x));
      uint32_t c0 = to_bits(cosf(x)), c1 = to_bits(oracle_cosf(x));
      if (s0 != s1) { if (!j->bad_sin && !j->bad_cos) j->first_bad = (uint32_t)b | (sign << 31); j->bad_sin++; }
      if (c0 != c1) { if (!j->bad_sin && !j->bad_cos) j->first_bad = (uint32_t)b | (sign << 31); j->bad_cos++; }
      j->checked++;
    }
  }
  return NULL;
}

static int sweep(uint32_t lo, uint32_t hi, uint32_t stride, int nt, const char *what) {
  pthread_t th[64];
  job_t jobs[64];
  if (nt > 64) nt = 64;
  uint64_t span = ((uint64_t)hi - lo + nt - 1) / nt;
  span = (span + stride - 1) / stride * stride;
  for (int i = 0; i < nt; ++i) {
    memset(&jobs[i], 0, sizeof(job_t));
    uint64_t a = lo + span * i, b = a + span;
    if (a > hi) a = hi;
    if (b > hi) b = hi;
    jobs[i].lo = (uint32_t)a;
    jobs[i].hi = (uint32_t)b;
    jobs[i].stride = stride;
    pthread_create(&th[i], NULL, run, &jobs[i]);
  }
  uint64_t checked = 0, bs = 0, bc = 0;
  uint32_t first = 0;
  for (int i = 0; i < nt; ++i) {
    pthread_join(th[i], NULL);
    checked += jobs[i].checked;
    if (!first && (jobs[i].bad_sin || jobs[i].bad_cos)) first = jobs[i].first_bad;
    bs += jobs[i].bad_sin;
    bc += jobs[i].bad_cos;
  }
  printf("%s: checked=%llu sin_mismatch=%llu cos_mismatch=%llu", what, (unsigned long long)checked,
         (unsigned long long)bs, (unsigned long long)bc);
  if (bs || bc) printf(" first_bad_bits=0x%08x", first);
  printf("\n");
  return (bs || bc) ? 1 : 0;
}

int main(int argc, char **argv) {
  float limit = argc > 1 ? (float)atof(argv[1]) : 128.0f;
  int nt = argc > 2 ? atoi(argv[2]) : 8;
  int rc = 0;
  rc |= sweep(0u, to_bits(limit) + 1u, 1u, nt, "exhaustive |x|<=limit");
  rc |= sweep(to_bits(limit), 0x7f800000u, 257u, nt, "strided  limit<|x|<inf");
  /* inf / nan produce nan in both */
  float specials[3] = {INFINITY, -INFINITY, NAN};
  for (int i = 0; i < 3; ++i)
    if (!isnan(oracle_sinf(specials[i])) || !isnan(oracle_cosf(specials[i]))) rc = 1;
  printf(rc ? "FAIL\n" : "OK\n");
  return rc;
}
