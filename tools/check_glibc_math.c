// check_glibc_math.c — host check that csrc/glibc_math.cuh reproduces this image's libm exp()/log() bit for bit.
//   g++ -O2 -ffp-contract=off -mfma -x c++ tools/check_glibc_math.c -o /tmp/check_glibc_math -lm && /tmp/check_glibc_math 100000000
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../ellalgo_rs_b200/csrc/glibc_math.cuh"

static uint64_t s = 0x9E3779B97F4A7C15ull;
static inline uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline double u01() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char** argv) {
    long n = argc > 1 ? atol(argv[1]) : 10000000;
    long bad_e = 0, bad_l = 0;
    for (long i = 0; i < n; ++i) {
        double x;
        switch (i & 7) {
            case 0: x = -10.0 + 20.0 * u01(); break;           // profit's range
            case 1: x = -745.0 + 1455.0 * u01(); break;        // whole finite range incl. subnormal results
            case 2: x = (u01() - 0.5) * 1e-3; break;
            case 3: x = ell_as_f64_(next()); break;            // arbitrary bit patterns (nan, inf, huge, tiny)
            default: x = -2.0 + 8.0 * u01(); break;
        }
        double a = exp(x), b = ellcuda::glibc::exp(x);
        if (ell_as_u64_(a) != ell_as_u64_(b) && !(a != a && b != b)) { if (bad_e++ < 5) printf("exp(%a): libm %a port %a\n", x, a, b); }
        double y;
        switch (i & 7) {
            case 0: y = 0.9 + 0.2 * u01(); break;               // the near-1 branch
            case 1: y = ell_as_f64_(next() & 0x7fffffffffffffffull); break;  // any non-negative pattern
            case 2: y = u01() * 1e-310; break;                  // subnormals
            case 3: y = ell_as_f64_(next()); break;
            default: y = exp(-5.0 + 20.0 * u01()); break;       // profit's range
        }
        a = log(y); b = ellcuda::glibc::log(y);
        if (ell_as_u64_(a) != ell_as_u64_(b) && !(a != a && b != b)) { if (bad_l++ < 5) printf("log(%a): libm %a port %a\n", y, a, b); }
    }
    printf("checked %ld arguments each: exp mismatches %ld, log mismatches %ld\n", n, bad_e, bad_l);
    return (bad_e || bad_l) ? 1 : 0;
}
