// Exhaustive check: for every finite float x, fma-based division by 127.5f equals IEEE x / 127.5f.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
int main() {
  const float b = 127.5f;
  const float y = 1.0f / b;  // correctly rounded reciprocal
  long long bad = 0, badnorm = 0;
#pragma omp parallel for reduction(+ : bad, badnorm) schedule(static)
  for (long long i = 0; i < (1ll << 32); ++i) {
    uint32_t u = (uint32_t)i;
    float x;
    memcpy(&x, &u, 4);
    if (!isfinite(x)) continue;
    float q = x * y;
    float r = fmaf(-q, b, x);
    float q2 = fmaf(r, y, q);
    float ref = x / b;
    uint32_t a, c;
    memcpy(&a, &q2, 4);
    memcpy(&c, &ref, 4);
    if (a != c) {
      ++bad;
      if (fabsf(x) >= 1e-30f && fabsf(x) < 1e38f) ++badnorm;
    }
  }
  printf("mismatches: %lld (of which |x| in [1e-30, 1e38): %lld)\n", bad, badnorm);
  return 0;
}
