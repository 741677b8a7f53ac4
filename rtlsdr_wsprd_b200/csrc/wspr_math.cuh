// Bit-faithful replicas of the three glibc float functions that sit on the data-dependent part of the decode
// path, usable from host and device code:
//
//   glibc_sinf / glibc_cosf   sync_and_demodulate phasor seeds (wsprd/wsprd.c:159-172) and the reference
//                             signal of subtract_signal2 (wsprd/wsprd.c:347-348)
//   glibc_log10f              candidate SNR (wsprd/wsprd.c:616)
//
// glibc (>= 2.28, x86-64) evaluates sinf/cosf/logf in binary64 with short minimax polynomials and rounds once to
// binary32; it is not correctly rounded, so a generic device libm differs from it in ~1 % of calls, which would
// perturb the decoder's float recurrences.  The functions below evaluate the same published polynomials in the
// same order.  glibc selects an FMA build of these routines at run time on CPUs with FMA3 (every current x86-64
// server part); the fma() calls below are placed where that build contracts.  tools/check_sincosf_replica.c and
// tools/check_log10f_replica.c compare the replicas against the host libm exhaustively (every float with
// |x| < 2^17 for sin/cos, every float in [1e-6, 1e9] for log10f): zero mismatches on the build host.
// tests/test_codec_host.py repeats a sampled version of that check through the shared library.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define WMATH __host__ __device__ __forceinline__
#else
#define WMATH inline
#endif

namespace wspr {

WMATH uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
WMATH float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
WMATH uint32_t abstop12(float f) { return (f2u(f) >> 20) & 0x7ffu; }

// polynomial for sin (n even) or cos (n odd) on [-pi/4, pi/4]; negcos flips the cosine polynomial's sign
WMATH float sincos_poly(double x, double x2, int n, bool negcos) {
    if ((n & 1) == 0) {
        const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
        double x3 = x * x2;
        double s1 = fma(x2, S3, S2);
        double x7 = x3 * x2;
        double s = fma(x3, S1, x);
        return (float)fma(x7, s1, s);
    }
    const double sg = negcos ? -1.0 : 1.0;
    const double C0 = sg * 0x1p0, C1 = sg * -0x1.ffffffd0c621cp-2, C2 = sg * 0x1.55553e1068f19p-5,
                 C3 = sg * -0x1.6c087e89a359dp-10, C4 = sg * 0x1.99343027bf8c3p-16;
    double x4 = x2 * x2;
    double c2 = fma(x2, C4, C3);
    double c1 = fma(x2, C1, C0);
    double x6 = x4 * x2;
    double c = fma(x4, C2, c1);
    return (float)fma(x6, c2, c);
}

// |x| < 120: quadrant by scaled conversion (hpi_inv carries a 2^24 factor)
WMATH double reduce_small(double x, int *np) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    double r = x * HPI_INV;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma(-(double)n, HPI, x);
}

// larger arguments: 96 bits of 4/pi selected by the exponent.  (On the device the table sits in global memory: as a local
// array every thread rebuilt it on its stack -- 24 stores and three local loads per call, which is what bound k_sub_ref.)
#define WSPR_INV_PIO4 {0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, \
                       0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, \
                       0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041}
#ifdef __CUDACC__
static __device__ const uint32_t g_inv_pio4[24] = WSPR_INV_PIO4;
#endif
WMATH double reduce_big(uint32_t xi, int *np) {
#ifdef __CUDA_ARCH__
    const uint32_t *inv_pio4 = g_inv_pio4;
#else
    static const uint32_t inv_pio4[24] = WSPR_INV_PIO4;
#endif
    const double PI63 = 0x1.921FB54442D18p-62;
    const uint32_t *arr = &inv_pio4[(xi >> 26) & 15];
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    res0 = (uint64_t)(uint32_t)(xi * arr[0]);
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * PI63;
}

WMATH float sign_of_quadrant(int n) { return ((n + 1) & 2) ? -1.0f : 1.0f; }   // {1,-1,-1,1}[n & 3]

WMATH float glibc_sinf(float y) {
    double x = y;
    int n;
    if (abstop12(y) < abstop12(0x1.921fb6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sincos_poly(x, x * x, 0, false);
    }
    if (abstop12(y) < abstop12(120.0f)) {
        x = reduce_small(x, &n);
        double s = sign_of_quadrant(n & 3);
        return sincos_poly(x * s, x * x, n, (n & 2) != 0);
    }
    if (abstop12(y) < 0x7f8u) {
        uint32_t xi = f2u(y);
        int sign = (int)(xi >> 31);
        x = reduce_big(xi, &n);
        double s = sign_of_quadrant((n + sign) & 3);
        return sincos_poly(x * s, x * x, n, ((n + sign) & 2) != 0);
    }
    return y - y;
}

WMATH float glibc_cosf(float y) {
    double x = y;
    int n;
    if (abstop12(y) < abstop12(0x1.921fb6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sincos_poly(x, x * x, 1, false);
    }
    if (abstop12(y) < abstop12(120.0f)) {
        x = reduce_small(x, &n);
        double s = sign_of_quadrant(n & 3);
        return sincos_poly(x * s, x * x, n ^ 1, (n & 2) != 0);
    }
    if (abstop12(y) < 0x7f8u) {
        uint32_t xi = f2u(y);
        int sign = (int)(xi >> 31);
        x = reduce_big(xi, &n);
        double s = sign_of_quadrant((n + sign) & 3);
        return sincos_poly(x * s, x * x, n ^ 1, ((n + sign) & 2) != 0);
    }
    return y - y;
}

// both at once: sinf(y) and cosf(y) share the argument reduction (bit-identical to calling the two above)
WMATH void glibc_sincosf(float y, float *sp, float *cp) {
    double x = y;
    int n;
    if (abstop12(y) < abstop12(0x1.921fb6p-1f)) {
        if (abstop12(y) < abstop12(0x1p-12f)) {
            *sp = y;
            *cp = 1.0f;
            return;
        }
        const double x2 = x * x;
        *sp = sincos_poly(x, x2, 0, false);
        *cp = sincos_poly(x, x2, 1, false);
        return;
    }
    if (abstop12(y) < abstop12(120.0f)) {
        x = reduce_small(x, &n);
        const double s = sign_of_quadrant(n & 3), xs = x * s, x2 = x * x;
        *sp = sincos_poly(xs, x2, n, (n & 2) != 0);
        *cp = sincos_poly(xs, x2, n ^ 1, (n & 2) != 0);
        return;
    }
    if (abstop12(y) < 0x7f8u) {
        const uint32_t xi = f2u(y);
        const int sign = (int)(xi >> 31);
        x = reduce_big(xi, &n);
        const double s = sign_of_quadrant((n + sign) & 3), xs = x * s, x2 = x * x;
        *sp = sincos_poly(xs, x2, n, ((n + sign) & 2) != 0);
        *cp = sincos_poly(xs, x2, n ^ 1, ((n + sign) & 2) != 0);
        return;
    }
    *sp = *cp = y - y;
}

// natural log, positive normal arguments only (the decode path passes values in [1, 2))
WMATH float glibc_logf_normal(float x) {
    const double T[16][2] = {{0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
                             {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
                             {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
                             {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
                             {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
                             {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
                             {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
                             {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
    const double LN2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2,
                 A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) % 16);
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & (0x1ffu << 23));
    double invc = T[i][0], logc = T[i][1], z = (double)u2f(iz);
    double r = fma(z, invc, -1.0);
    double y0 = fma((double)k, LN2, logc);
    double r2 = r * r;
    double y = fma(A1, r, A2);
    y = fma(A0, r2, y);
    y = fma(y, r2, y0 + r);
    return (float)y;
}

// log10f for positive finite x (smoothed-spectrum peaks are > 0); float arithmetic around logf as in glibc 2.39
WMATH float glibc_log10f(float x) {
    const float two25 = 3.3554432000e+07f, ivln10 = 4.3429449201e-01f, log10_2hi = 3.0102920532e-01f,
                log10_2lo = 7.9034151668e-07f;
    int32_t hx = (int32_t)f2u(x), k = 0;
    if (hx < 0x00800000) {
        if ((hx & 0x7fffffff) == 0) return -two25 / fabsf(x);
        if (hx < 0) return (x - x) / (x - x);
        k -= 25;
        x *= two25;
        hx = (int32_t)f2u(x);
    }
    if (hx >= 0x7f800000) return x + x;
    k += (hx >> 23) - 127;
    int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
    float y = (float)(k + i);
    float z = y * log10_2lo + ivln10 * glibc_logf_normal(u2f((uint32_t)hx));
    return z + y * log10_2hi;
}

}  // namespace wspr
