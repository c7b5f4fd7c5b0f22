/*
 * tests/csrc/fastdiv_model.c -- host model of the division fast paths of pic1dp_b200/csrc/particle_kernels.cuh
 * (div_const / div_pos + div_suspect), compiled by tests/test_fastdiv_model_cpu.py.
 *
 * The kernels claim: q1 = fma(r, y, q0) with q0 = a*y, r = fma(-q0, b, a) equals RN(a/b) whenever the exact residual test
 * |fma(-q1, b, a)| <= b*ulp(q1)/2 passes; operands that fail it (or are out of range) are flagged and redone with the
 * IEEE division.  This program checks exactly that on the host, where fma() is exact and a/b is the IEEE quotient:
 * every NON-flagged case must equal a/b bit for bit, and the flag rate must stay small.  It models y either as
 * RN(1/b) (div_const) or as a ~20-bit approximation refined by the kernel's cubic Newton step (div_pos, MUFU.RCP64H).
 *
 * usage: fastdiv_model <ncases> <seed>   -> prints "checked N flagged F mismatches M random_flagged R of K"
 * (R of K: flags among the plainly random operands, i.e. how often the hot loop would fall back on ordinary data)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2];
static uint64_t next(void) { /* xorshift128+ */
  uint64_t x = s[0], y = s[1];
  s[0] = y;
  x ^= x << 23;
  s[1] = x ^ y ^ (x >> 17) ^ (y >> 26);
  return s[1] + y;
}
static double u01(void) { return (double)(next() >> 11) * 0x1p-53; }
static uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

/* div_suspect: true when q1 is NOT provably RN(a/b) */
static int div_suspect(double a, double b, double q1) {
  const double r1 = fma(-q1, b, a);
  const uint64_t e = bits(q1) & 0x7ff0000000000000ULL;
  const double h = from_bits(bits(b) + e - (1076ULL << 52)); /* b * ulp(q1) / 2 */
  return fabs(r1) > h;
}

/* div_const: y = RN(1/b) */
static int div_const(double a, double b, double y, double *q) {
  const double q0 = a * y;
  const double r = fma(-q0, b, a);
  *q = fma(r, y, q0);
  return !(a >= 0x1p-800) || div_suspect(a, b, *q);
}

/* div_pos: y from a 20-bit reciprocal + the kernel's one cubic Newton step */
static int div_pos(double a, double b, double *q) {
  const uint32_t eb = (uint32_t)((bits(b) >> 32) & 0x7ff00000u) - (523u << 20);
  const uint32_t ea = (uint32_t)((bits(a) >> 32) & 0x7ff00000u) - (523u << 20);
  double y = from_bits(bits(1.0 / b) & 0xffffffff00000000ULL); /* ~20 good bits like MUFU.RCP64H */
  double e = fma(-b, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  const double q0 = a * y;
  const double r = fma(-q0, b, a);
  *q = fma(r, y, q0);
  return ((eb | ea) >= (1000u << 20)) || !(b > 0.0) || div_suspect(a, b, *q);
}

int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  s[0] = 0x9E3779B97F4A7C15ULL ^ (uint64_t)(argc > 2 ? atol(argv[2]) : 1);
  s[1] = 0xD1B54A32D192ED03ULL;
  long checked = 0, flagged = 0, bad = 0, rnd = 0, rnd_flagged = 0;
  /* 1. x / lx for box lengths of the configurations (lx = 2 pi / k) and x in [0, lx], plus adversarial x = q*lx rounded */
  const double lxs[5] = {2.0 * 3.1415926535897932384626 / 0.36, 4.0 * 3.14159265358979323846, 1.0, 17.0, 0.1};
  for (int c = 0; c < 5; c++) {
    const double lx = lxs[c], y = 1.0 / lx;
    for (long i = 0; i < n; i++) {
      double x = u01() * lx;
      const int plain = (i & 7) != 0 && (i & 1023) != 0;
      if ((i & 7) == 0) x = ((double)(next() % 4096) + 0.5 * (double)(next() & 1)) / 4096.0 * lx; /* near cell edges */
      if ((i & 1023) == 0) x = nextafter(lx, (i & 1024) ? 0.0 : 2.0 * lx);
      double q;
      const int flag = div_const(x, lx, y, &q);
      checked++;
      rnd += plain;
      rnd_flagged += plain && flag;
      if (flag) { flagged++; continue; }
      if (bits(q) != bits(x / lx)) bad++;
    }
  }
  /* 2. general positive quotients over 50 binades (2^-40 .. 2^10: the operands of -d ln f0/dv are O(1) or smaller; the
   * kernel's one-instruction range test (eb | ea) also flags some harmless operand pairs above 2^12, which only costs
   * the slow path), and quotients constructed to sit next to rounding midpoints */
  for (long i = 0; i < 4 * n; i++) {
    double a = ldexp(1.0 + u01(), (int)(next() % 50) - 40), b = ldexp(1.0 + u01(), (int)(next() % 50) - 40);
    if ((i & 3) == 0) { /* a = b * (m + 1/2 ulp) rounded: a/b lands near a midpoint */
      const double m = 1.0 + u01();
      a = b * (m + 0x1p-53);
    }
    double q;
    const int flag = div_pos(a, b, &q);
    checked++;
    rnd += (i & 3) != 0;
    rnd_flagged += (i & 3) != 0 && flag;
    if (flag) { flagged++; continue; }
    if (bits(q) != bits(a / b)) bad++;
  }
  /* 3. wrap_x / wrap_n: for -lx < x < 2 lx the select form (x >= lx ? x - lx : x; x < 0 ? x + lx : .) must equal the
   * reference's  x = mod(x, lx); if (x < 0) x = x + lx  bit for bit (src/pic1dp_interaction.F90:102-104) */
  for (int c = 0; c < 5; c++) {
    const double lx = lxs[c];
    for (long i = 0; i < n; i++) {
      double x = (3.0 * u01() - 1.0) * lx;
      if ((i & 15) == 0) x = nextafter((double)(next() % 3) * lx - ((i & 16) ? lx : 0.0), (i & 32) ? 1e300 : -1e300);
      if ((i & 255) == 0) x = (i & 256) ? -0.0 : 0.0;
      if (!(x > -lx && x < lx + lx)) continue; /* the kernels send these to fmod */
      double ref = fmod(x, lx);
      if (ref < 0.0) ref = ref + lx;
      double xw = (x >= lx) ? x - lx : x;
      xw = (x < 0.0) ? x + lx : xw;
      checked++;
      if (bits(xw) != bits(ref)) bad++;
    }
  }
  printf("checked %ld flagged %ld mismatches %ld random_flagged %ld of %ld\n", checked, flagged, bad, rnd_flagged, rnd);
  return bad != 0;
}
