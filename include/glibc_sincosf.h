/* glibc_sincosf.h -- restatement of glibc >= 2.28 sinf/cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c,
 * s_sincosf.h; the ARM "optimized routines" algorithm): reduce in double, degree-7/8 polynomial in double,
 * one rounding to float.  The reference calls cos()/sin() on a float angle (src/ORBextractor.cc:97), which binds to
 * these libm functions; they are NOT correctly rounded (about 1.3 % of the angles in [0, 2 pi] differ by 1 ulp from
 * the correctly rounded value), so bit-exact descriptors need this very algorithm, not a better one.
 * Only the branches reachable for |x| < 120 are restated (the descriptor angle lies in [0, 2 pi]).
 * tests/test_oracle_cpu.py checks this header against the host libm on millions of inputs.
 * Usable from C, C++ and CUDA (define B200_HD before including for __host__ __device__). */
#ifndef B200_GLIBC_SINCOSF_H_
#define B200_GLIBC_SINCOSF_H_
#include <stdint.h>
#ifndef B200_HD
#define B200_HD static inline
#endif

B200_HD uint32_t b200_abstop12(float x) {
  union { float f; uint32_t u; } v;
  v.f = x;
  return (v.u >> 20) & 0x7ff;
}

/* n even: sine polynomial, n odd: cosine polynomial; neg selects the second (negated cosine) table */
B200_HD float b200_sinf_poly(double x, double x2, int n, int neg) {
  const double c0 = neg ? -0x1p0 : 0x1p0, c1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2,
               c2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5,
               c3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10,
               c4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  if ((n & 1) == 0) {
    const double x3 = x * x2;
    const double t1 = s2 + x2 * s3;
    const double x7 = x3 * x2;
    const double s = x + x3 * s1;
    return (float)(s + x7 * t1);
  } else {
    const double x4 = x2 * x2;
    const double t2 = c3 + x2 * c4;
    const double t1 = c0 + x2 * c1;
    const double x6 = x4 * x2;
    const double c = t1 + x4 * c2;
    return (float)(c + x6 * t2);
  }
}

B200_HD double b200_reduce_fast(double x, int* np) {
  const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
  const double r = x * hpi_inv;
  const int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return x - n * hpi;
}

B200_HD float b200_cosf(float y) {
  double x = y;
  if (b200_abstop12(y) < b200_abstop12(0x1.921FB6p-1f)) {
    if (b200_abstop12(y) < b200_abstop12(0x1p-12f)) return 1.0f;
    return b200_sinf_poly(x, x * x, 1, 0);
  }
  int n;
  x = b200_reduce_fast(x, &n);
  const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   /* sign[n & 3] = {1,-1,-1,1} */
  return b200_sinf_poly(x * s, x * x, n ^ 1, (n & 2) != 0);
}

B200_HD float b200_sinf(float y) {
  double x = y;
  if (b200_abstop12(y) < b200_abstop12(0x1.921FB6p-1f)) {
    if (b200_abstop12(y) < b200_abstop12(0x1p-12f)) return y;
    return b200_sinf_poly(x, x * x, 0, 0);
  }
  int n;
  x = b200_reduce_fast(x, &n);
  const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return b200_sinf_poly(x * s, x * x, n, (n & 2) != 0);
}
#endif
