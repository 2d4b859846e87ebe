/* TEST INFRASTRUCTURE (oracle/): CPU check of the quotient sequence the axisymmetric kernel uses for
 * the division by r^2 (sv_quotient in pyfds_b200/csrc/fds_streamv.cuh; the reference divides with
 * numpy's IEEE division, pyfds/acoustics.py:210-212):
 *     y = RN(1 / b);  q0 = RN(a y);  t = b q0 - a (one fma, exact);  q = RN(q0 - t y)
 * compared bit for bit with a / b over the ranges the kernel admits (2^-256 <= b <= 2^256, the
 * significand of b not all ones, 2^-693 <= |a| <= 2^677 or a = +-0).
 *
 *   quotient_check N  ->  prints "cases C mismatches M"; exit status 1 if M > 0.
 *
 * Cases: (1) b = ((i + 1/2) dx)^2 as the model builds it, a with random significand and exponent over
 * the whole admitted range; (2) random b, a whose significand is next to b's; (3) quotients placed next
 * to a rounding boundary (a = b (q + ulp/2) and a = b q, each +-2 ulp); (4) signed zeros. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static uint64_t rnd(void) {
    uint64_t a = s[0], b = s[1];
    s[0] = b;
    a ^= a << 23;
    s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}
static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t to_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

static double quotient(double a, double b, double y) {
    double q0 = a * y;
    double t = fma(b, q0, -a);
    return fma(-t, y, q0);
}

static long cases = 0, mismatches = 0;
static void check(double a, double b) {
    double y = 1.0 / b;
    double q = quotient(a, b, y), ref = a / b;
    ++cases;
    if (to_bits(q) != to_bits(ref)) {
        if (++mismatches < 10) printf("MISMATCH a=%a b=%a got %a want %a\n", a, b, q, ref);
    }
}
static double random_numerator(void) {
    uint64_t m = rnd() & 0xfffffffffffffull;
    uint64_t e = (1023 - 693) + rnd() % (693 + 677 + 1);
    return from_bits(((rnd() & 1) << 63) | (e << 52) | m);
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 1000000;
    const double dxs[] = {1e-3, 1e-4, 2.5e-4, 1e-5, 3.3e-3, 1e-2, 7e-6};
    for (int d = 0; d < 7; ++d)
        for (long i = 0; i < n; ++i) {
            double r = (double)(rnd() % 40000) * dxs[d] + dxs[d] / 2;
            double b = r * r;
            if ((to_bits(b) & 0xfffffffffffffull) == 0xfffffffffffffull) continue;
            check(random_numerator(), b);
        }
    for (long i = 0; i < 4 * n; ++i) {
        uint64_t mb = rnd() & 0xfffffffffffffull;
        if (i % 7 == 0) mb |= 0xffffffffff000ull;
        if (i % 11 == 0) mb &= 0xfffull;
        if (mb == 0xfffffffffffffull) continue;
        uint64_t eb = 1023 - 256 + rnd() % 513;
        double b = from_bits((eb << 52) | mb);
        double a = random_numerator();
        if (i % 5 == 0) {
            uint64_t m = (mb + rnd() % 5 - 2) & 0xfffffffffffffull;
            a = from_bits((to_bits(a) & 0xfff0000000000000ull) | m);
        }
        check(a, b);
    }
    for (long i = 0; i < n; ++i) {
        uint64_t mb = rnd() & 0xfffffffffffffull;
        if (i % 3 == 0) mb = ((rnd() & 0xfffff) << (rnd() % 32)) & 0xfffffffffffffull;
        if (mb == 0xfffffffffffffull) continue;
        double b = from_bits((1023ull << 52) | mb);
        double q = from_bits((1023ull << 52) | (rnd() & 0xfffffffffffffull));
        double near_half = fma(b, q, b * ldexp(1.0, -53)), near_q = b * q;
        for (int k = -2; k <= 2; ++k) {
            check(from_bits(to_bits(near_half) + k), b);
            check(from_bits(to_bits(near_q) + k), b);
        }
    }
    {
        double b = 2.5e-7, y = 1.0 / b;
        double zp = quotient(0.0, b, y), zn = quotient(-0.0, b, y);
        cases += 2;
        if (to_bits(zp) != 0 || to_bits(zn) != 0x8000000000000000ull) {
            ++mismatches;
            printf("MISMATCH zeros %a %a\n", zp, zn);
        }
    }
    printf("cases %ld mismatches %ld\n", cases, mismatches);
    return mismatches != 0;
}
