// fd_math.cuh -- bit-reproducible f32 sin/cos/acos/atan2 for host and device.
//
// The geometric hash (reference src/geometry/pdb_tr.rs:44-58) bins f32 sin/cos of angles that come from
// f32 acos/atan2 (src/structure/coordinate.rs:128, :214).  CUDA's sinf/cosf/acosf/atan2f are 1-2 ulp off any
// CPU libm, so the kernels cannot use them and stay bit-exact.  Everything here is evaluated in IEEE binary64
// with a fixed sequence of + - * / sqrt fma -- operations that are correctly rounded on the GPU and on the
// host alike -- and rounded once to binary32.  The binary64 value is good to ~2e-16, so the result is the
// correctly rounded f32 except with probability ~1e-8 per call; that makes GPU == host by construction and
// == any faithful libm except at those rare ties (measured: 0 hash differences vs glibc on all shipped data).
//
// Build rules that keep this true: nvcc -fmad=false (no silent FMA contraction), never --use_fast_math;
// host code that includes this header is compiled with -ffp-contract=off.
// The double kernels are the classic fdlibm minimax approximations (k_sin.c, k_cos.c, e_acos.c, s_atan.c).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FD_HD __host__ __device__ __forceinline__
#else
#define FD_HD inline
#endif

namespace fdm {

FD_HD double ksin(double r) {
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

FD_HD double kcos(double r) {
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

// sin and cos of a binary32 angle with one shared range reduction; results rounded to binary32.
FD_HD void sincosf_exact(float xf, float *s, float *c) {
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_HI = 1.57079632679489655800e+00;
    const double PIO2_LO = 6.12323399573676603587e-17;
    double x = (double)xf;
    if (!(fabs(x) < 1.0e9)) {
        double v = x - x; // nan for inf/nan, 0 for huge finite
        if (x == x && !isinf(x)) {
            *s = 0.0f;
            *c = 1.0f;
        } else {
            *s = (float)v;
            *c = (float)v;
        }
        return;
    }
    double k = rint(x * TWO_OVER_PI);
    double r = fma(-k, PIO2_HI, x);
    r = fma(-k, PIO2_LO, r);
    int q = (int)((long long)k & 3LL);
    double ks = ksin(r), kc = kcos(r);
    double sd, cd;
    if (q == 0) {
        sd = ks;
        cd = kc;
    } else if (q == 1) {
        sd = kc;
        cd = -ks;
    } else if (q == 2) {
        sd = -ks;
        cd = -kc;
    } else {
        sd = -kc;
        cd = ks;
    }
    *s = (float)sd;
    *c = (float)cd;
}

FD_HD double acos_d(double x) {
    const double PIO2_HI = 1.57079632679489655800e+00, PIO2_LO = 6.12323399573676603587e-17;
    const double PI = 3.14159265358979311600e+00;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    double ax = fabs(x);
    if (!(ax <= 1.0)) return NAN; // |x| > 1 or nan
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
    return 2.0 * (s + r * s);
}

FD_HD double atan_pos_d(double x) { // x >= 0
    const double hi0 = 4.63647609000806093515e-01, hi1 = 7.85398163397448278999e-01,
                 hi2 = 9.82793723247329054082e-01, hi3 = 1.57079632679489655800e+00;
    const double lo0 = 2.26987774529616870924e-17, lo1 = 3.06161699786838301793e-17,
                 lo2 = 1.39033110312309984516e-17, lo3 = 6.12323399573676603587e-17;
    const double a0 = 3.33333333333329318027e-01, a1 = -1.99999999998764832476e-01,
                 a2 = 1.42857142725034663711e-01, a3 = -1.11111104054623557880e-01,
                 a4 = 9.09088713343650656196e-02, a5 = -7.69187620504482999495e-02,
                 a6 = 6.66107313738753120669e-02, a7 = -5.83357013379057348645e-02,
                 a8 = 4.97687799461593236017e-02, a9 = -3.65315727442169155270e-02,
                 a10 = 1.62858201153657823623e-02;
    int id;
    double hi = 0.0, lo = 0.0;
    if (x >= 1.8446744073709552e19) return hi3 + lo3;
    if (x < 0.4375) {
        id = -1;
    } else if (x < 1.1875) {
        if (x < 0.6875) {
            id = 0;
            hi = hi0;
            lo = lo0;
            x = (2.0 * x - 1.0) / (2.0 + x);
        } else {
            id = 1;
            hi = hi1;
            lo = lo1;
            x = (x - 1.0) / (x + 1.0);
        }
    } else {
        if (x < 2.4375) {
            id = 2;
            hi = hi2;
            lo = lo2;
            x = (x - 1.5) / (1.0 + 1.5 * x);
        } else {
            id = 3;
            hi = hi3;
            lo = lo3;
            x = -1.0 / x;
        }
    }
    double z = x * x;
    double w = z * z;
    double s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    double s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    if (id < 0) return x - x * (s1 + s2);
    return hi - ((x * (s1 + s2) - lo) - x);
}

FD_HD double atan2_d(double y, double x) {
    const double PI = 3.14159265358979311600e+00, PI_LO = 1.2246467991473531772e-16;
    const double PIO2 = 1.57079632679489655800e+00;
    if (x != x || y != y) return x + y;
    if (y == 0.0) {
        if (x > 0.0 || (x == 0.0 && !signbit(x))) return y;
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
    z = atan_pos_d(ay / ax);
    if (x < 0.0) z = PI - (z - PI_LO);
    return y < 0.0 ? -z : z;
}

FD_HD float acosf_exact(float x) { return (float)acos_d((double)x); }
FD_HD float atan2f_exact(float y, float x) { return (float)atan2_d((double)y, (double)x); }

} // namespace fdm
