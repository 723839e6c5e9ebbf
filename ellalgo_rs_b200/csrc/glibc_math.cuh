// glibc_math.cuh — exp() and log() that reproduce glibc 2.39's x86-64 FMA build bit for bit, for device code.
//
// Why: the reference's oracles call f64::exp / f64::ln (src/oracles/profit_oracle.rs:54-113, src/quasicvx.rs:25-49),
// which on Linux are glibc's.  CUDA's libdevice exp/log differ from glibc in the last bit for a fraction of a percent
// of arguments — enough to move iteration counts of the cutting-plane loop on a few problems in ten thousand.  The
// algorithms are the published ones of glibc's sysdeps/ieee754/dbl-64/e_exp.c and e_log.c (ARM optimized-routines:
// table-driven, 2^(k/128) resp. 128 sub-intervals of [0x1.6p-1, 0x1.6p0), low-order polynomial), written out again
// here with every fused multiply-add that GCC's -ffp-contract=fast forms in the libm binary spelled explicitly
// (the library is compiled with -fmad=false, so nothing else fuses).  tools/check_glibc_math.c checks the same
// source, compiled for the host, against libm on ~10^8 arguments; tables: glibc_tables.h.
#pragma once
#include <cstdint>

#include "glibc_tables.h"

#if defined(__CUDA_ARCH__)
#define ELL_MATH_FN __device__ __forceinline__
#define ELL_FMA(a, b, c) __fma_rn((a), (b), (c))
#define ELL_AS_U64(x) ((uint64_t)__double_as_longlong(x))
#define ELL_AS_F64(u) __longlong_as_double((long long)(u))
#define ELL_TAB_ATTR __device__
#else
#include <cstring>
#define ELL_MATH_FN static inline
#define ELL_FMA(a, b, c) __builtin_fma((a), (b), (c))
static inline uint64_t ell_as_u64_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double ell_as_f64_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define ELL_AS_U64(x) ell_as_u64_(x)
#define ELL_AS_F64(u) ell_as_f64_(u)
#define ELL_TAB_ATTR static
#endif

namespace ellcuda {
namespace glibc {

ELL_TAB_ATTR const double kExpHdr[8] = ELL_GLIBC_EXP_HDR;   // invln2N, shift, negln2hiN, negln2loN, C2..C5
ELL_TAB_ATTR const uint64_t kExpTab[256] = ELL_GLIBC_EXP_TAB;
ELL_TAB_ATTR const double kLogHdr[18] = ELL_GLIBC_LOG_HDR;  // ln2hi, ln2lo, A[5], B[11]
ELL_TAB_ATTR const double kLogTab[256] = ELL_GLIBC_LOG_TAB; // {invc, logc} x 128

// e_exp.c: exp(x) = 2^(k/N) * exp(r),  x = k ln2/N + r,  N = 128
ELL_MATH_FN double exp_special(double tmp, uint64_t sbits, uint64_t ki) {  // specialcase(): |x| > 512 region
    if ((ki & 0x80000000u) == 0) {  // k > 0: the exponent of scale might have overflowed by <= 460
        sbits -= 1009ull << 52;
        const double scale = ELL_AS_F64(sbits);
        return 0x1p1009 * ELL_FMA(scale, tmp, scale);
    }
    sbits += 1022ull << 52;  // k < 0: careful in the subnormal range
    const double scale = ELL_AS_F64(sbits);
    const double st = scale * tmp;  // used twice below, so the libm binary keeps it as a plain product
    double y = scale + st;
    if (y < 1.0) {
        double lo = scale - y + st;
        double hi = 1.0 + y;
        lo = 1.0 - hi + y + lo;
        y = (hi + lo) - 1.0;
        if (y == 0.0) y = 0.0;
    }
    return 0x1p-1022 * y;
}

ELL_MATH_FN double exp(double x) {
    const double InvLn2N = kExpHdr[0], Shift = kExpHdr[1], NegLn2hiN = kExpHdr[2], NegLn2loN = kExpHdr[3];
    const double C2 = kExpHdr[4], C3 = kExpHdr[5], C4 = kExpHdr[6], C5 = kExpHdr[7];
    uint32_t abstop = (uint32_t)(ELL_AS_U64(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) {      // |x| < 2^-54 or |x| >= 512 or nan
        if (abstop - 0x3c9u >= 0x80000000u) return 1.0 + x;
        if (abstop >= 0x409u) {                    // |x| >= 1024
            if (ELL_AS_U64(x) == 0xfff0000000000000ull) return 0.0;
            if (abstop >= 0x7ffu) return 1.0 + x;
            return (ELL_AS_U64(x) >> 63) ? 0.0 : ELL_AS_F64(0x7ff0000000000000ull);
        }
        abstop = 0;
    }
    double kd = ELL_FMA(InvLn2N, x, Shift);        // z = InvLn2N*x; kd = z + Shift   (fused in the libm binary)
    const uint64_t ki = ELL_AS_U64(kd);
    kd -= Shift;
    const double r = ELL_FMA(kd, NegLn2loN, ELL_FMA(kd, NegLn2hiN, x));
    const uint64_t idx = 2 * (ki % 128);
    const uint64_t top = ki << (52 - 7);
    const double tail = ELL_AS_F64(kExpTab[idx]);
    const uint64_t sbits = kExpTab[idx + 1] + top;
    const double r2 = r * r;
    // tmp = tail + r + r2*(C2 + r*C3) + r2*r2*(C4 + r*C5)
    double tmp = ELL_FMA(r2, ELL_FMA(r, C3, C2), tail + r);
    tmp = ELL_FMA(r2 * r2, ELL_FMA(r, C5, C4), tmp);
    if (abstop == 0) return exp_special(tmp, sbits, ki);
    const double scale = ELL_AS_F64(sbits);
    return ELL_FMA(scale, tmp, scale);
}

// e_log.c
ELL_MATH_FN double log(double x) {
    const double Ln2hi = kLogHdr[0], Ln2lo = kLogHdr[1];
    const double* A = kLogHdr + 2;
    const double* B = kLogHdr + 7;
    uint64_t ix = ELL_AS_U64(x);
    uint32_t top = (uint32_t)(ix >> 48);
    const uint64_t LO = 0x3fee000000000000ull;  // asuint64(1.0 - 0x1p-4)
    const uint64_t HI = 0x3ff1090000000000ull;  // asuint64(1.0 + 0x1.09p-4)
    if (ix - LO < HI - LO) {
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = x - 1.0;
        const double r2 = r * r;
        const double r3 = r * r2;
        double p3 = ELL_FMA(r3, B[10], ELL_FMA(r2, B[9], ELL_FMA(r, B[8], B[7])));
        double p2 = ELL_FMA(r3, p3, ELL_FMA(r2, B[6], ELL_FMA(r, B[5], B[4])));
        double p1 = ELL_FMA(r3, p2, ELL_FMA(r2, B[3], ELL_FMA(r, B[2], B[1])));
        double w = r * 0x1p27;
        const double rhi = r + w - w;
        const double rlo = r - rhi;
        w = rhi * rhi * B[0];
        const double hi = r + w;
        double lo = r - hi + w;
        lo = ELL_FMA(B[0] * rlo, rhi + r, lo);
        double y = ELL_FMA(r3, p1, lo);
        y += hi;
        return y;
    }
    if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
        if (ix * 2 == 0) return -ELL_AS_F64(0x7ff0000000000000ull);         // log(0) = -inf
        if (ix == 0x7ff0000000000000ull) return x;                          // log(inf) = inf
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return (x - x) / (x - x);  // x < 0 or nan -> nan
        ix = ELL_AS_U64(x * 0x1p52);                                        // subnormal: normalise
        ix -= 52ull << 52;
    }
    const uint64_t OFF = 0x3fe6000000000000ull;
    const uint64_t tmp = ix - OFF;
    const int i = (int)((tmp >> (52 - 7)) % 128);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffull << 52));
    const double invc = kLogTab[2 * i], logc = kLogTab[2 * i + 1];
    const double z = ELL_AS_F64(iz);
    const double r = ELL_FMA(z, invc, -1.0);
    const double kd = (double)k;
    const double w = ELL_FMA(kd, Ln2hi, logc);
    const double hi = w + r;
    const double lo = ELL_FMA(kd, Ln2lo, w - hi + r);
    const double r2 = r * r;
    // y = lo + r2*A[0] + r*r2*(A[1] + r*A[2] + r2*(A[3] + r*A[4])) + hi
    const double p = ELL_FMA(r2, ELL_FMA(r, A[4], A[3]), ELL_FMA(r, A[2], A[1]));
    double y = ELL_FMA(r * r2, p, ELL_FMA(r2, A[0], lo));
    y += hi;
    return y;
}

}  // namespace glibc
}  // namespace ellcuda
