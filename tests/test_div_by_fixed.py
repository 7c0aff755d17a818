"""csrc/samplers.cuh: div_by_fixed -- a / b for a divisor fixed per parameter set without the
division sequence (reciprocal + two fma residual corrections; Markstein).  The engine's claim
is "the bits of a / b"; on the GPU the bit-exact filter tests cover it through the compare's
exponential noise, here the same operation sequence is run in C (gcc, fma, no contraction)
against `/` over random numerators and a set of divisors, the SIR default rate 1e6 among them."""
import os
import subprocess
import textwrap

SRC = textwrap.dedent(r"""
    #include <math.h>
    #include <stdint.h>
    #include <stdio.h>
    #include <string.h>
    static uint64_t s[4] = {1, 2, 3, 4};
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    static uint64_t next(void) {
      uint64_t r = s[0] + s[3], t = s[1] << 17;
      s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
      return r;
    }
    static double div_by_fixed(double a, double b, double y) {
      double q = a * y;
      double r = fma(-q, b, a);
      q = fma(r, y, q);
      r = fma(-q, b, a);
      return fma(r, y, q);
    }
    int main(void) {
      const double bs[] = {1e6, 3.0, 0.1, 7.25e-3, 123456.789, 1000001.0, 0.9999999, 1.0000001,
                           2.5e5, 1e-6, 4096.0, 1e12};
      long bad = 0, tot = 0;
      for (unsigned k = 0; k < sizeof bs / sizeof bs[0]; ++k) {
        const double b = bs[k], y = 1.0 / b;
        for (long i = 0; i < 1500000; ++i) {
          const double u = (double)(next() >> 11) * 0x1p-53;
          if (u == 0.0) continue;
          double a = -log(u);                       /* what the compare divides */
          if (i & 1) {                               /* and numerators over 120 binades */
            uint64_t bits = (next() >> 12) | ((uint64_t)(1023 - 60 + (next() % 120)) << 52);
            memcpy(&a, &bits, 8);
          }
          ++tot;
          if (div_by_fixed(a, b, y) != a / b) ++bad;
        }
      }
      printf("%ld %ld\n", tot, bad);
      return 0;
    }
""")


def test_div_by_fixed_gives_the_bits_of_division(tmp_path):
    src = tmp_path / "div.c"
    src.write_text(SRC)
    exe = tmp_path / "div"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", str(exe), str(src), "-lm"],
                   check=True)
    tot, bad = (int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True,
                                               check=True).stdout.split())
    assert tot > 10_000_000 and bad == 0
