// fd_kabsch.cuh -- device Kabsch (kabsch.rs:157-554, mode 2) shared by fd_kabsch.cu and fd_verify.cu.
#pragma once
#include <stdint.h>

namespace fdk {

struct P3 {
    double x, y, z;
};
struct FlatPoints { // points stored contiguously as xyz triples
    const float *p;
    __device__ P3 operator()(uint32_t i) const { return {(double)p[3 * i], (double)p[3 * i + 1], (double)p[3 * i + 2]}; }
};
struct GatherPoints { // CA, CB interleaved, gathered by residue index: point 2k = CA(res[k]), 2k+1 = CB(res[k])
    const float *ca, *cb;
    const uint32_t *res;
    uint64_t base;
    __device__ P3 operator()(uint32_t i) const {
        const uint64_t r = base + res[i >> 1];
        const float *s = (i & 1) ? cb : ca;
        return {(double)s[3 * r], (double)s[3 * r + 1], (double)s[3 * r + 2]};
    }
};

template <class PX, class PY>
__device__ void kabsch_one(PX px, PY py, uint32_t n, float *U, float *T, float *rmsd_out) {
    const double EPSILON = 1.0e-8, TOLERANCE = 0.01, SQRT3 = 1.7320508075688772;
    const int IP[9] = {0, 1, 3, 1, 2, 4, 3, 4, 5};
    const int IP2312[4] = {1, 2, 0, 1};
    double u[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double t[3] = {0, 0, 0};
    if (n == 0) {
        for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1.0f : 0.0f;
        T[0] = T[1] = T[2] = 0.0f;
        *rmsd_out = 3.40282347e+38f;
        return;
    }
    double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0}, sx[3] = {0, 0, 0}, sy[3] = {0, 0, 0}, sz[3] = {0, 0, 0};
    for (uint32_t i = 0; i < n; i++) {
        const P3 a1 = px(i), a2 = py(i);
        const double c1[3] = {a1.x, a1.y, a1.z};
        const double c2[3] = {a2.x, a2.y, a2.z};
        for (int j = 0; j < 3; j++) {
            s1[j] += c1[j];
            s2[j] += c2[j];
        }
        sx[0] += c1[0] * c2[0];
        sx[1] += c1[0] * c2[1];
        sx[2] += c1[0] * c2[2];
        sy[0] += c1[1] * c2[0];
        sy[1] += c1[1] * c2[1];
        sy[2] += c1[1] * c2[2];
        sz[0] += c1[2] * c2[0];
        sz[1] += c1[2] * c2[1];
        sz[2] += c1[2] * c2[2];
    }
    double xc[3], yc[3];
    for (int j = 0; j < 3; j++) {
        xc[j] = s1[j] / (double)n;
        yc[j] = s2[j] / (double)n;
    }
    double r[3][3];
    for (int j = 0; j < 3; j++) {
        r[j][0] = sx[j] - s1[0] * s2[j] / (double)n;
        r[j][1] = sy[j] - s1[1] * s2[j] / (double)n;
        r[j][2] = sz[j] - s1[2] * s2[j] / (double)n;
    }
    const double det_r = r[0][0] * (r[1][1] * r[2][2] - r[1][2] * r[2][1]) -
                         r[0][1] * (r[1][0] * r[2][2] - r[1][2] * r[2][0]) +
                         r[0][2] * (r[1][0] * r[2][1] - r[1][1] * r[2][0]);
    double rr[6];
    {
        int m = 0;
        for (int j = 0; j < 3; j++)
            for (int i = 0; i <= j; i++) rr[m++] = r[0][i] * r[0][j] + r[1][i] * r[1][j] + r[2][i] * r[2][j];
    }
    const double spur = (rr[0] + rr[2] + rr[5]) / 3.0;
    const double cof =
        (((rr[2] * rr[5] - rr[4] * rr[4]) + rr[0] * rr[5] - rr[3] * rr[3]) + rr[0] * rr[2] - rr[1] * rr[1]) / 3.0;
    const double det = det_r * det_r;
    if (spur > 0.0) {
        const double d = spur * spur;
        const double h = d - cof;
        const double g = (spur * cof - det) / 2.0 - spur * h;
        if (h > 0.0) {
            const double sqrth = sqrt(h);
            double disc = h * h * h - g * g;
            if (disc < 0.0) disc = 0.0;
            const double sqrt_disc = sqrt(disc);
            double d_ang;
            if (fabs(g) > 1e18) d_ang = g > 0.0 ? 3.14159265358979323846 / 3.0 : 0.0;
            else d_ang = atan2(sqrt_disc, -g) / 3.0;
            const double cth = sqrth * cos(d_ang);
            const double sth = sqrth * SQRT3 * sin(d_ang);
            double e[3];
            e[0] = spur + 2.0 * cth;
            e[1] = spur - cth + sth;
            e[2] = spur - cth - sth;
            double a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            bool a_failed = false, b_failed = false;
            for (int li = 0; li < 2; li++) {
                const int l = li == 0 ? 0 : 2;
                const double dl = e[l];
                double ss[6];
                ss[0] = (dl - rr[2]) * (dl - rr[5]) - rr[4] * rr[4];
                ss[1] = (dl - rr[5]) * rr[1] + rr[3] * rr[4];
                ss[2] = (dl - rr[0]) * (dl - rr[5]) - rr[3] * rr[3];
                ss[3] = (dl - rr[2]) * rr[3] + rr[1] * rr[4];
                ss[4] = (dl - rr[0]) * rr[4] + rr[1] * rr[3];
                ss[5] = (dl - rr[0]) * (dl - rr[2]) - rr[1] * rr[1];
                for (int k = 0; k < 6; k++)
                    if (fabs(ss[k]) <= EPSILON) ss[k] = 0.0;
                const double A = fabs(ss[0]), B = fabs(ss[2]), C = fabs(ss[5]);
                const int j = (A >= B && A >= C) ? 0 : (B >= C ? 1 : 2);
                double dn = 0.0;
                for (int i = 0; i < 3; i++) {
                    const int k = IP[3 * j + i];
                    a[i][l] = ss[k];
                    dn += ss[k] * ss[k];
                }
                dn = dn > EPSILON ? 1.0 / sqrt(dn) : 0.0;
                for (int i = 0; i < 3; i++) a[i][l] *= dn;
            }
            const double dot = a[0][0] * a[0][2] + a[1][0] * a[1][2] + a[2][0] * a[2][2];
            int m1, mm;
            if (e[0] - e[1] > e[1] - e[2]) {
                m1 = 2;
                mm = 0;
            } else {
                m1 = 0;
                mm = 2;
            }
            double p = 0.0;
            for (int i = 0; i < 3; i++) {
                a[i][m1] = a[i][m1] - dot * a[i][mm];
                p += a[i][m1] * a[i][m1];
            }
            if (p <= TOLERANCE) {
                int j = 0;
                p = 1.0;
                for (int i = 0; i < 3; i++)
                    if (p < fabs(a[i][mm])) {
                        p = fabs(a[i][mm]);
                        j = i;
                    }
                const int k = IP2312[j], l = IP2312[j + 1];
                p = sqrt(a[k][mm] * a[k][mm] + a[l][mm] * a[l][mm]);
                if (p > TOLERANCE) {
                    a[j][m1] = 0.0;
                    a[k][m1] = -a[l][mm] / p;
                    a[l][m1] = a[k][mm] / p;
                } else {
                    a_failed = true;
                }
            } else {
                p = 1.0 / sqrt(p);
                for (int i = 0; i < 3; i++) a[i][m1] *= p;
            }
            if (!a_failed) {
                a[0][1] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
                a[1][1] = a[2][2] * a[0][0] - a[2][0] * a[0][2];
                a[2][1] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
                for (int l = 0; l < 2; l++) {
                    double db = 0.0;
                    for (int i = 0; i < 3; i++) {
                        b[i][l] = r[i][0] * a[0][l] + r[i][1] * a[1][l] + r[i][2] * a[2][l];
                        db += b[i][l] * b[i][l];
                    }
                    db = db > EPSILON ? 1.0 / sqrt(db) : 0.0;
                    for (int i = 0; i < 3; i++) b[i][l] *= db;
                }
                double dot_b = 0.0;
                for (int i = 0; i < 3; i++) dot_b += b[i][0] * b[i][1];
                double pb = 0.0;
                for (int i = 0; i < 3; i++) {
                    b[i][1] -= dot_b * b[i][0];
                    pb += b[i][1] * b[i][1];
                }
                if (pb <= TOLERANCE) {
                    pb = 1.0;
                    int j = 0;
                    for (int i = 0; i < 3; i++)
                        if (pb < fabs(b[i][0])) {
                            pb = fabs(b[i][0]);
                            j = i;
                        }
                    const int k = IP2312[j], l = IP2312[j + 1];
                    pb = sqrt(b[k][0] * b[k][0] + b[l][0] * b[l][0]);
                    if (pb > TOLERANCE) {
                        b[j][1] = 0.0;
                        b[k][1] = -b[l][0] / pb;
                        b[l][1] = b[k][0] / pb;
                    } else {
                        b_failed = true;
                    }
                } else {
                    pb = 1.0 / sqrt(pb);
                    for (int i = 0; i < 3; i++) b[i][1] *= pb;
                }
                if (!b_failed) {
                    b[0][2] = b[1][0] * b[2][1] - b[1][1] * b[2][0];
                    b[1][2] = b[2][0] * b[0][1] - b[2][1] * b[0][0];
                    b[2][2] = b[0][0] * b[1][1] - b[0][1] * b[1][0];
                    for (int i = 0; i < 3; i++)
                        for (int j = 0; j < 3; j++)
                            u[i][j] = b[i][0] * a[j][0] + b[i][1] * a[j][1] + b[i][2] * a[j][2];
                    for (int i = 0; i < 3; i++)
                        t[i] = yc[i] - (u[i][0] * xc[0] + u[i][1] * xc[1] + u[i][2] * xc[2]);
                }
            }
        }
    } else {
        for (int i = 0; i < 3; i++) t[i] = yc[i] - (u[i][0] * xc[0] + u[i][1] * xc[1] + u[i][2] * xc[2]);
    }
    double sum_sq = 0.0;
    for (uint32_t i = 0; i < n; i++) {
        const P3 xp = px(i), yp = py(i);
        const double x0 = xp.x, x1 = xp.y, x2 = xp.z;
        const double yv[3] = {yp.x, yp.y, yp.z};
        const double tr[3] = {u[0][0] * x0 + u[0][1] * x1 + u[0][2] * x2 + t[0],
                              u[1][0] * x0 + u[1][1] * x1 + u[1][2] * x2 + t[1],
                              u[2][0] * x0 + u[2][1] * x1 + u[2][2] * x2 + t[2]};
        for (int j = 0; j < 3; j++) {
            const double diff = tr[j] - yv[j];
            sum_sq += diff * diff;
        }
    }
    const double rms = sqrt(sum_sq / (double)n);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) U[3 * i + j] = (float)u[i][j];
    for (int i = 0; i < 3; i++) T[i] = (float)t[i];
    float rf = (float)rms;
    if (rf != rf) rf = 3.40282347e+38f;
    *rmsd_out = rf;
}


} // namespace fdk
