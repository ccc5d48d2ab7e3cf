/* fd_oracle_math.h -- TEST INFRASTRUCTURE (oracle). Not part of the product.
 *
 * f32 sin/cos/acos/atan2 used by the oracle's restatement of the reference's
 * geometric hash (src/geometry/pdb_tr.rs:44-58 calls f32::sin/cos,
 * src/structure/coordinate.rs:128 f32::acos, :214 f32::atan2).
 *
 * The reference gets these from the platform libm through Rust std, i.e. the
 * last ulp is not pinned by the reference.  The oracle
 * therefore has two modes:
 *   - FD_MATH_EXACT (default): evaluate in IEEE binary64 with a fixed sequence
 *     of + - * / sqrt fma (all correctly rounded on every IEEE machine, CPU or
 *     GPU) and round ONCE to binary32.  The binary64 value is good to ~1e-16,
 *     so the result is the correctly rounded f32 value except with
 *     probability ~1e-8 per call, and it is reproducible bit-for-bit anywhere.
 *   - FD_MATH_LIBM: call glibc sinf/cosf/acosf/atan2f like the Rust binary
 *     would on this machine.  tests/ count the hash disagreements between the
 *     two modes on every shipped structure.
 *
 * The double kernels are the classic fdlibm minimax approximations
 * (k_sin.c, k_cos.c, e_acos.c, s_atan.c; Sun Microsystems, freely usable).
 */
#ifndef FD_ORACLE_MATH_H
#define FD_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- binary64 kernels -------------------------------------------------- */

static inline double fdo_ksin(double r) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = r * r;
    double p = S6;
    p = p * z + S5;
    p = p * z + S4;
    p = p * z + S3;
    p = p * z + S2;
    p = p * z + S1;
    return r + (r * z) * p;
}

static inline double fdo_kcos(double r) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = r * r;
    double p = C6;
    p = p * z + C5;
    p = p * z + C4;
    p = p * z + C3;
    p = p * z + C2;
    p = p * z + C1;
    return (1.0 - 0.5 * z) + (z * z) * p;
}

/* sin and cos of a binary32 argument, one shared range reduction. */
static inline void fdo_sincos_d(float xf, double *s, double *c) {
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_HI = 1.57079632679489655800e+00;
    const double PIO2_LO = 6.12323399573676603587e-17;
    double x = (double)xf;
    if (!(fabs(x) < 1.0e9)) { /* inf, nan, or far outside any angle we hash */
        *s = x - x;           /* nan for inf/nan; 0 for huge finite (never hit) */
        *c = x - x;
        if (x == x && fabs(x) != INFINITY) { *s = 0.0; *c = 1.0; }
        return;
    }
    double k = rint(x * TWO_OVER_PI);
    double r = fma(-k, PIO2_HI, x);
    r = fma(-k, PIO2_LO, r);
    int q = (int)((long long)k & 3LL);
    double ks = fdo_ksin(r), kc = fdo_kcos(r);
    switch (q) {
        case 0: *s = ks;  *c = kc;  break;
        case 1: *s = kc;  *c = -ks; break;
        case 2: *s = -ks; *c = -kc; break;
        default: *s = -kc; *c = ks; break;
    }
}

static inline double fdo_acos_d(double x) {
    const double PIO2_HI = 1.57079632679489655800e+00, PIO2_LO = 6.12323399573676603587e-17;
    const double PI = 3.14159265358979311600e+00;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    double ax = fabs(x);
    if (!(ax <= 1.0)) return NAN; /* |x|>1 or nan */
    if (ax == 1.0) return x > 0.0 ? 0.0 : PI;
    if (ax < 0.5) {
        double z = x * x;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double r = p / q;
        return PIO2_HI - (x - (PIO2_LO - x * r));
    }
    double z = (1.0 - ax) * 0.5;
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    double s = sqrt(z);
    double r = p / q;
    if (x < 0.0) {
        double w = r * s - PIO2_LO;
        return PI - 2.0 * (s + w);
    }
    /* x > 0.5: acos(x) = 2*asin(sqrt((1-x)/2)); binary64 throughout is plenty
     * for a binary32 result, so no hi/lo split of s is needed. */
    return 2.0 * (s + r * s);
}

static inline double fdo_atan_d(double x) { /* x >= 0, finite or +inf */
    const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01,
                              9.82793723247329054082e-01, 1.57079632679489655800e+00};
    const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17,
                              1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01,
                           1.42857142725034663711e-01,  -1.11111104054623557880e-01,
                           9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                           6.66107313738753120669e-02,  -5.83357013379057348645e-02,
                           4.97687799461593236017e-02,  -3.65315727442169155270e-02,
                           1.62858201153657823623e-02};
    int id;
    if (x >= 1.8446744073709552e19) return atanhi[3] + atanlo[3]; /* >= 2^64 */
    if (x < 0.4375) {
        id = -1;
    } else if (x < 1.1875) {
        if (x < 0.6875) { id = 0; x = (2.0 * x - 1.0) / (2.0 + x); }
        else            { id = 1; x = (x - 1.0) / (x + 1.0); }
    } else {
        if (x < 2.4375) { id = 2; x = (x - 1.5) / (1.0 + 1.5 * x); }
        else            { id = 3; x = -1.0 / x; }
    }
    double z = x * x;
    double w = z * z;
    double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    return atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
}

static inline double fdo_atan2_d(double y, double x) {
    const double PI = 3.14159265358979311600e+00, PI_LO = 1.2246467991473531772e-16;
    const double PIO2 = 1.57079632679489655800e+00;
    if (x != x || y != y) return NAN;
    if (y == 0.0) {
        if (x > 0.0 || (x == 0.0 && !signbit(x))) return y; /* +-0 */
        return signbit(y) ? -PI : PI;
    }
    if (x == 0.0) return y < 0.0 ? -PIO2 : PIO2;
    double ay = fabs(y), ax = fabs(x);
    double z;
    if (isinf(ax) || isinf(ay)) {
        if (isinf(ax) && isinf(ay)) z = x > 0.0 ? PIO2 * 0.5 : 3.0 * (PIO2 * 0.5);
        else if (isinf(ay)) z = PIO2;
        else z = x > 0.0 ? 0.0 : PI;
        return y < 0.0 ? -z : z;
    }
    z = fdo_atan_d(ay / ax);
    if (x < 0.0) z = PI - (z - PI_LO);
    return y < 0.0 ? -z : z;
}

/* ---- binary32 front ends ---------------------------------------------- */

#ifndef FD_MATH_LIBM
#define FD_MATH_LIBM 0
#endif

extern int fdo_math_use_libm; /* runtime switch, defined in fd_oracle.cpp */

static inline float fdo_sinf(float x) {
    if (fdo_math_use_libm) return sinf(x);
    double s, c; fdo_sincos_d(x, &s, &c); return (float)s;
}
static inline float fdo_cosf(float x) {
    if (fdo_math_use_libm) return cosf(x);
    double s, c; fdo_sincos_d(x, &s, &c); return (float)c;
}
static inline float fdo_acosf(float x) {
    if (fdo_math_use_libm) return acosf(x);
    return (float)fdo_acos_d((double)x);
}
static inline float fdo_atan2f(float y, float x) {
    if (fdo_math_use_libm) return atan2f(y, x);
    return (float)fdo_atan2_d((double)y, (double)x);
}

#ifdef __cplusplus
}
#endif
#endif /* FD_ORACLE_MATH_H */
