// Exhaustive proof (all 2^32 fp32 bit patterns) that  x / C  ==  q1  for
//     r  = (float)(1.0 / C)        (nearest fp32 to the reciprocal)
//     q0 = x * r                   (one rounding)
//     e  = fmaf(-q0, C, x)         (exact residual of the estimate, one rounding)
//     q1 = fmaf(e, r, q0)          (corrected quotient, one rounding)
// i.e. that an IEEE-754 correctly rounded division by the CONSTANT C can be replaced by FMUL + 2 FFMA
// (T / 250 in maxWater, V / 127 in calcEvaporation).  Counts every input whose results differ
// (NaN == NaN; signed zeros compared by bits).  gcc -O2 -fopenmp -ffp-contract=off const_div_proof.c -lm
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(int argc, char** argv) {
  for (int a = 1; a < argc; a++) {
    const float C = (float)atof(argv[a]);
    const float r = (float)(1.0 / (double)C);
    unsigned long long bad = 0, bad_normal = 0;
    uint32_t first = 0; int have = 0;
#pragma omp parallel for reduction(+ : bad, bad_normal) schedule(static)
    for (long long i = 0; i < (1LL << 32); i++) {
      const float x = u2f((uint32_t)i);
      volatile float want = x / C;
      const float q0 = x * r;
      const float e = fmaf(-q0, C, x);
      const float q1 = fmaf(e, r, q0);
      const float w = want;
      const int same = (w != w && q1 != q1) || f2u(w) == f2u(q1);
      if (!same) {
        bad++;
        const float ax = fabsf(x);
        if (ax >= 1e-30f && ax <= 1e30f) bad_normal++;
#pragma omp critical
        if (!have) { have = 1; first = (uint32_t)i; }
      }
    }
    printf("C = %g (r = %.9g): %llu of 2^32 inputs differ, %llu of them with 1e-30 <= |x| <= 1e30", C, r, bad, bad_normal);
    if (have) printf("; e.g. x = %.9g (0x%08x): x / C = %.9g, sequence = %.9g", u2f(first), first, u2f(first) / C, fmaf(fmaf(-(u2f(first) * r), C, u2f(first)), r, u2f(first) * r));
    printf("\n");
  }
  return 0;
}
