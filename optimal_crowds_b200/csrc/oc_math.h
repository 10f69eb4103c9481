/* oc_math.h -- fully specified FP64 elementary functions (exp, atan2, sin, cos, hypot2).
 *
 * Why: the GCFM update is chaotic at round-off level: a 1-ulp difference
 * in one pair force reaches 1e-8 within ~70 steps.  libdevice, glibc and numpy's SIMD loops each
 * round exp/atan2/sin/cos differently, so "the same" formula gives different trajectories on every
 * platform -- including between two CPUs running the numpy reference.  To make whole runs
 * reproducible bit for bit between the CUDA kernels and a host restatement, every transcendental
 * on the GCFM path goes through the functions below, which use only IEEE-754 correctly rounded
 * operations (+ - * / sqrt fma) in a fixed order.  Compiled with contraction OFF on both sides
 * (nvcc -fmad=false, gcc -ffp-contract=off) they return identical bits on sm_100a and on any host.
 *
 * Algorithms: classic argument reduction + minimax polynomials (Sun fdlibm family of
 * approximations: exp via r = x - k ln2 and a degree-5 rational-form kernel, atan via 4-interval
 * reduction and a degree-11 odd polynomial, sin/cos via 3-term Cody-Waite reduction by pi/2 and
 * degree-13/14 kernels).  Max observed error vs numpy < 1 ulp (tests/test_cpu_oracle.py::test_oc_math_*).
 *
 * The polynomial coefficients and the reduction constants (P1..P5, aT[], atanhi/atanlo, S1..S6, C1..C6, the
 * pio2 splits) are those of FreeBSD/Sun fdlibm (e_exp.c, s_atan.c, k_sin.c, k_cos.c, e_rem_pio2.c), whose notice
 * is reproduced here as its licence asks:
 *
 *   ====================================================
 *   Copyright (C) 1993 by Sun Microsystems, Inc. All rights reserved.
 *
 *   Developed at SunSoft, a Sun Microsystems, Inc. business.
 *   Permission to use, copy, modify, and distribute this
 *   software is freely granted, provided that this notice
 *   is preserved.
 *   ====================================================
 *
 * This header is product code.  The CPU oracle (oracle/) includes it for these elementary
 * functions ONLY, and tests pin them against numpy; all GCFM formulae are restated independently
 * on both sides.
 */
#ifndef OC_MATH_H
#define OC_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OCM_FN __host__ __device__ __forceinline__
/* the polynomial cores are called from several places of the GCFM kernels: keeping one copy of each keeps the
 * sweep kernel inside the instruction cache (ncu: no_instruction stalls with everything inlined) */
#define OCM_CORE static __host__ __device__ __noinline__
#else
#define OCM_FN static inline
#define OCM_CORE static inline
#endif

OCM_FN uint64_t ocm_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
OCM_FN double ocm_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}
OCM_FN uint32_t ocm_hi(double x) { return (uint32_t)(ocm_bits(x) >> 32); }
OCM_FN double ocm_abs(double x) { return ocm_from_bits(ocm_bits(x) & 0x7fffffffffffffffULL); }
OCM_FN int ocm_isnan(double x) { return x != x; }

/* sqrt(fma(y,y,x*x)): what np.linalg.norm of a 2-vector evaluates to through OpenBLAS ddot on
 * FMA-capable x86.  fma and sqrt are correctly rounded everywhere. */
OCM_FN double ocm_norm2(double x, double y) { return sqrt(fma(y, y, x * x)); }

/* np.minimum / np.maximum: a NaN operand gives NaN (C's and CUDA's fmin / fmax return the other operand instead).
 * pedestrians.py:243-245,262,278,332 use them on quantities that are NaN for coincident agents or an agent exactly
 * on a wall node (R = 0 -> e = 0/0), and the reference's forces are NaN there. */
OCM_FN double ocm_npmin(double a, double b) { return (a <= b || a != a) ? a : b; }  /* numpy's scalar loop, verbatim */
OCM_FN double ocm_npmax(double a, double b) { return (a >= b || a != a) ? a : b; }

/* 2^k for k in [-1022, 1023] */
OCM_FN double ocm_pow2i(int k) { return ocm_from_bits((uint64_t)(k + 1023) << 52); }

/* ---------------------------------------------------------------- exp */
OCM_CORE double ocm_exp(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 inv_ln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    if (ocm_isnan(x)) return x;
    if (x > 7.09782712893383973096e+02) return INFINITY;
    if (x < -7.45133219101941108420e+02) return 0.0;
    double ax = ocm_abs(x);
    if (ax < 0x1p-28) return 1.0 + x;
    double hi, lo, r;
    int k;
    if (ax > 0.34657359027997264 /* ln2/2 */) {
        double kf = inv_ln2 * x + (x < 0 ? -0.5 : 0.5);
        k = (int)kf; /* truncation toward zero == round-half-away of inv_ln2*x */
        double t = (double)k;
        hi = x - t * ln2_hi; /* t*ln2_hi exact: ln2_hi has 32 significant bits, |k| < 2^11 */
        lo = t * ln2_lo;
        r = hi - lo;
    } else {
        k = 0;
        hi = x;
        lo = 0.0;
        r = x;
    }
    double z = r * r;
    double c = r - z * (P1 + z * (P2 + z * (P3 + z * (P4 + z * P5))));
    double y;
    if (k == 0) return 1.0 - ((r * c) / (c - 2.0) - r);
    y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
    if (k >= -1021 && k <= 1023) return y * ocm_pow2i(k);
    if (k > 1023) return y * 2.0 * ocm_pow2i(k - 1); /* k == 1024 */
    return y * ocm_pow2i(k + 1000) * 0x1p-1000;      /* gradual underflow */
}

/* ---------------------------------------------------------------- atan / atan2 */
OCM_CORE double ocm_atan(double x) {
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01,
                 aT2 = 1.42857142725034663711e-01, aT3 = -1.11111104054623557880e-01,
                 aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02,
                 aT8 = 4.97687799461593236017e-02, aT9 = -3.65315727442169155270e-02,
                 aT10 = 1.62858201153657823623e-02;
    if (ocm_isnan(x)) return x;
    int neg = (ocm_bits(x) >> 63) != 0;
    double ax = ocm_abs(x);
    double hi_c, lo_c;
    int id;
    if (ax >= 0x1p66) { /* |x| huge: pi/2 */
        double r = 1.57079632679489655800e+00 + 6.12323399573676603587e-17;
        return neg ? -r : r;
    }
    if (ax < 0.4375) {
        if (ax < 0x1p-27) return x;
        id = -1;
        hi_c = 0;
        lo_c = 0;
    } else if (ax < 1.1875) {
        if (ax < 0.6875) { /* atan(0.5) + atan((2x-1)/(2+x)) */
            id = 0;
            hi_c = 4.63647609000806093515e-01;
            lo_c = 2.26987774529616870924e-17;
            ax = (2.0 * ax - 1.0) / (2.0 + ax);
        } else { /* atan(1) + atan((x-1)/(x+1)) */
            id = 1;
            hi_c = 7.85398163397448278999e-01;
            lo_c = 3.06161699786838301793e-17;
            ax = (ax - 1.0) / (ax + 1.0);
        }
    } else {
        if (ax < 2.4375) { /* atan(1.5) + atan((x-1.5)/(1+1.5x)) */
            id = 2;
            hi_c = 9.82793723247329054082e-01;
            lo_c = 1.39033110312309984516e-17;
            ax = (ax - 1.5) / (1.0 + 1.5 * ax);
        } else { /* pi/2 - atan(1/x) */
            id = 3;
            hi_c = 1.57079632679489655800e+00;
            lo_c = 6.12323399573676603587e-17;
            ax = -1.0 / ax;
        }
    }
    double z = ax * ax, w = z * z;
    double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) {
        double r = ax - ax * (s1 + s2);
        return neg ? -r : r;
    }
    double r = hi_c - ((ax * (s1 + s2) - lo_c) - ax);
    return neg ? -r : r;
}

OCM_FN double ocm_atan2(double y, double x) {
    const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16,
                 pi_o_2 = 1.5707963267948965580e+00, pi_o_4 = 7.8539816339744827900e-01;
    if (ocm_isnan(x) || ocm_isnan(y)) return x + y;
    int sy = (int)(ocm_bits(y) >> 63), sx = (int)(ocm_bits(x) >> 63);
    double ay = ocm_abs(y), ax = ocm_abs(x);
    if (x == 1.0) return ocm_atan(y);
    if (ay == 0.0) { /* y = +-0 */
        if (!sx) return y;            /* atan2(+-0, +x) = +-0 (also x = +0) */
        return sy ? -pi : pi;         /* atan2(+-0, -x) = +-pi (also x = -0) */
    }
    if (ax == 0.0) return sy ? -pi_o_2 : pi_o_2;
    int x_inf = ax == INFINITY, y_inf = ay == INFINITY;
    if (x_inf) {
        if (y_inf) {
            double r = sx ? 3.0 * pi_o_4 : pi_o_4;
            return sy ? -r : r;
        }
        double r = sx ? pi : 0.0;
        return sy ? -r : r;
    }
    if (y_inf) return sy ? -pi_o_2 : pi_o_2;
    /* exponent gap k = ilogb(y) - ilogb(x) (from the raw exponent fields, as in the classic scheme) */
    int ky = (int)((ocm_bits(y) >> 52) & 0x7ff), kx = (int)((ocm_bits(x) >> 52) & 0x7ff);
    int k = ky - kx;
    double z;
    if (k > 60) z = pi_o_2 + 0.5 * pi_lo;      /* |y/x| > 2^60 */
    else if (sx && k < -60) z = 0.0;           /* 0 > |y|/x > -2^-60 */
    else z = ocm_atan(ocm_abs(y / x));
    if (!sx) return sy ? -z : z;
    if (!sy) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

/* ---------------------------------------------------------------- sin / cos */
/* kernels on |x| <= pi/4 with tail y */
OCM_FN double ocm_ksin(double x, double y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x, v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
OCM_FN double ocm_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}
/* reduce x (|x| < ~1e5) to r + t = x - n*pi/2, |r| <= pi/4(+eps); returns n & 3 */
OCM_CORE int ocm_rem_pio2(double x, double *r, double *t) {
    const double inv_pio2 = 6.36619772367581382433e-01, p1 = 1.57079632673412561417e+00,
                 p2 = 6.07710050630396597660e-11, p3 = 2.02226624871116645580e-21,
                 p3t = 8.47842766036889956997e-32;
    double fn = x * inv_pio2;
    fn = (fn < 0) ? (double)(long long)(fn - 0.5) : (double)(long long)(fn + 0.5);
    /* p1, p2, p3 carry 33 significant bits each: fn*p1, fn*p2, fn*p3 are exact for |fn| < 2^20 */
    double a = x - fn * p1;
    double b = a - fn * p2;
    double w = fn * p3;
    double hi = b - w;
    double lo = (b - hi) - w; /* rounding error of b - w, exact when |b| >= |w| or by Sterbenz */
    lo = lo - fn * p3t;
    *r = hi + lo;
    *t = (hi - *r) + lo;
    return (int)((long long)fn) & 3;
}
OCM_FN double ocm_sin(double x) {
    if (ocm_isnan(x) || ocm_abs(x) == INFINITY) return x - x;
    double ax = ocm_abs(x);
    if (ax <= 7.85398163397448278999e-01) {
        if (ax < 0x1p-26) return x;
        return ocm_ksin(x, 0.0);
    }
    double r, t;
    int n = ocm_rem_pio2(x, &r, &t);
    switch (n) {
        case 0: return ocm_ksin(r, t);
        case 1: return ocm_kcos(r, t);
        case 2: return -ocm_ksin(r, t);
        default: return -ocm_kcos(r, t);
    }
}
OCM_FN double ocm_cos(double x) {
    if (ocm_isnan(x) || ocm_abs(x) == INFINITY) return x - x;
    double ax = ocm_abs(x);
    if (ax <= 7.85398163397448278999e-01) {
        if (ax < 0x1p-27) return 1.0;
        return ocm_kcos(x, 0.0);
    }
    double r, t;
    int n = ocm_rem_pio2(x, &r, &t);
    switch (n) {
        case 0: return ocm_kcos(r, t);
        case 1: return -ocm_ksin(r, t);
        case 2: return -ocm_kcos(r, t);
        default: return ocm_ksin(r, t);
    }
}

/* ---------------------------------------------------------------- shared-work variants
 * Same bits as the separate calls (tests/test_cpu_oracle.py), less work: the GCFM pair force needs
 * atan2(y,x) AND atan2(-y,-x) (pedestrians.py:266,268) and sin AND cos of the same angle (:270-271). */
OCM_CORE void ocm_sincos(double x, double *s, double *c) {
    double ax = ocm_abs(x);
    if (ocm_isnan(x) || ax == INFINITY || ax <= 7.85398163397448278999e-01) {
        *s = ocm_sin(x);
        *c = ocm_cos(x);
        return;
    }
    double r, t;
    int n = ocm_rem_pio2(x, &r, &t);
    double ks = ocm_ksin(r, t), kc = ocm_kcos(r, t);
    switch (n) {
        case 0: *s = ks; *c = kc; break;
        case 1: *s = kc; *c = -ks; break;
        case 2: *s = -ks; *c = -kc; break;
        default: *s = -kc; *c = ks; break;
    }
}

OCM_FN void ocm_atan2_both(double y, double x, double *a_pos, double *a_neg) {
    const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
    double ay = ocm_abs(y), ax = ocm_abs(x);
    int ky = (int)((ocm_bits(y) >> 52) & 0x7ff), kx = (int)((ocm_bits(x) >> 52) & 0x7ff);
    int k = ky - kx;
    /* anything that takes a special path in ocm_atan2 is evaluated by the two plain calls */
    if (ocm_isnan(x) || ocm_isnan(y) || ay == 0.0 || ax == 0.0 || ax == INFINITY || ay == INFINITY || ax == 1.0 ||
        k > 60 || k < -60) {
        *a_pos = ocm_atan2(y, x);
        *a_neg = ocm_atan2(-y, -x);
        return;
    }
    int sy = (int)(ocm_bits(y) >> 63), sx = (int)(ocm_bits(x) >> 63);
    double z = ocm_atan(ocm_abs(y / x)); /* |(-y)/(-x)| == |y/x| exactly */
    double zq = z - pi_lo;
    /* (y,x): signs (sy,sx); (-y,-x): signs (!sy,!sx) -- same quadrant rules as ocm_atan2 */
    double p, n;
    if (!sx) p = sy ? -z : z; else p = (!sy) ? pi - zq : zq - pi;
    if (sx) n = (!sy) ? -z : z; else n = sy ? pi - zq : zq - pi;
    *a_pos = p;
    *a_neg = n;
}

#endif /* OC_MATH_H */
