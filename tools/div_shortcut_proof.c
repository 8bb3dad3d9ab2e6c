/* CPU check of the "divide a product by one of its factors" shortcut proposed
 * in DESIGN.md section 9 (not used by the engine yet).
 *
 *   y = fl(t * rho);  wanted: q = fl(y / rho)  (IEEE round-to-nearest-even)
 *
 * The exact quotient is t + r0/rho with r0 = y - t*rho (exactly representable,
 * one FMA) and |r0/rho| < ulp(t), so q is t or one of its two neighbours,
 * decided by comparing |r0| with (ulp(t)/2)*|rho| -- a power-of-two scaling of
 * rho, exact.  t exactly a power of two (spacing changes below it) is left to
 * the ordinary division.
 *
 *   gcc -O2 -fopenmp -ffp-contract=off tools/div_shortcut_proof.c -lm -o /tmp/dsp
 *   /tmp/dsp [pairs per thread] [seed]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* returns 1 and sets *q when the shortcut applies */
static inline int shortcut(double t, double rho, double y, double* q) {
  const uint64_t tb = bits(t);
  const int et = (int)((tb >> 52) & 0x7ff);
  if ((tb & 0xfffffffffffffULL) == 0 || et == 0 || et == 0x7ff) return 0; /* 2^k, 0, subnormal, inf/nan */
  const double r0 = fma(-t, rho, y);                 /* y - t*rho, exact */
  const double thr = ldexp(fabs(rho), et - 1023 - 53); /* (ulp(t)/2) * |rho| */
  const double a = fabs(r0);
  uint64_t qb = tb;
  if (a > thr || (a == thr && (tb & 1))) {
    /* move one ulp towards the exact quotient: away from zero when r0/rho has t's sign */
    const int away = ((bits(r0) ^ bits(rho) ^ tb) >> 63) == 0;
    qb = away ? tb + 1 : tb - 1;
  }
  *q = from_bits(qb);
  return 1;
}

static inline uint64_t rng(uint64_t* s) {
  uint64_t x = *s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return *s = x;
}

#include <stdlib.h>
int main(int argc, char** argv) {
  long long bad = 0, done = 0, skipped = 0;
  const long long per_thread = argc > 1 ? atoll(argv[1]) : 40000000LL;
  const uint64_t seed = argc > 2 ? (uint64_t)atoll(argv[2]) : 1;
#pragma omp parallel reduction(+ : bad, done, skipped)
  {
    uint64_t s = 0x9E3779B97F4A7C15ULL * seed;
#ifdef _OPENMP
    extern int omp_get_thread_num(void);
    s ^= (uint64_t)(omp_get_thread_num() + 1) * 0xD1B54A32D192ED03ULL;
#endif
    for (long long i = 0; i < per_thread; ++i) {
      uint64_t m1 = rng(&s), m2 = rng(&s), c = rng(&s);
      /* mantissa patterns: random, few low bits, all ones minus few bits */
      uint64_t mt = m1 & 0xfffffffffffffULL, mr = m2 & 0xfffffffffffffULL;
      switch (c & 7) {
        case 0: mt &= 0xff; break;
        case 1: mt |= 0xfffffffffff00ULL; break;
        case 2: mr &= 0xff; break;
        case 3: mr |= 0xfffffffffff00ULL; break;
        case 4: mt = (mt & 0xffff) | 0x8000000000000ULL; break;
        default: break;
      }
      const int e_t = 1023 + (int)((c >> 8) % 61) - 30;     /* |t| in 2^-30 .. 2^30 */
      const int e_r = 1023 + (int)((c >> 16) % 41) - 30;    /* rho in 2^-30 .. 2^10 */
      const uint64_t sign = (c >> 40) & 1;
      const double t = from_bits((sign << 63) | ((uint64_t)e_t << 52) | mt);
      const uint64_t sign_r = (c >> 41) & 1;   /* the model has rho > 0; the rule holds either way */
      const double rho = from_bits((sign_r << 63) | ((uint64_t)e_r << 52) | mr);
      const double y = t * rho;
      double q;
      if (!shortcut(t, rho, y, &q)) { ++skipped; continue; }
      ++done;
      if (bits(q) != bits(y / rho)) {
        if (++bad <= 5) printf("MISMATCH t=%a rho=%a y=%a got %a want %a\n", t, rho, y, q, y / rho);
      }
    }
  }
  printf("%lld pairs checked, %lld left to the ordinary division, %lld mismatches\n", done, skipped, bad);
  return bad != 0;
}
