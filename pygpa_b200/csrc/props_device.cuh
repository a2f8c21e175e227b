// Per-pixel 2x2 SVD with LAPACK's conventions and the property formulas of props_from_Jac
// (pyGPA/property_extract.py:137-178).  __host__ __device__ so that tests/test_props_host.py can
// run exactly this code on the CPU against numpy.linalg.svd and the oracle.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define GPA_HD __host__ __device__
#else
#define GPA_HD
#endif

namespace gpa {

// ---------------------------------------------------------------------------------------------
// LAPACK's 2x2 SVD
// ---------------------------------------------------------------------------------------------
GPA_HD inline double fsign(double a, double b) {   // Fortran SIGN(a, b)
    return copysign(a, b);
}

// dlasv2: [[csl, snl], [-snl, csl]] [[f, g], [0, h]] [[csr, -snr], [snr, csr]] = diag(ssmax, ssmin)
GPA_HD inline void dlasv2(double f, double g, double h, double& ssmin, double& ssmax, double& snr, double& csr,
                       double& snl, double& csl) {
    const double eps = 1.1102230246251565e-16;
    double ft = f, fa = fabs(f), ht = h, ha = fabs(h);
    int pmax = 1;
    const bool swap = ha > fa;
    if (swap) {
        pmax = 3;
        double t = ft; ft = ht; ht = t;
        t = fa; fa = ha; ha = t;
    }
    const double gt = g, ga = fabs(g);
    double clt, crt, slt, srt;
    if (ga == 0.0) {
        ssmin = ha; ssmax = fa; clt = 1.0; crt = 1.0; slt = 0.0; srt = 0.0;
    } else {
        bool gasmal = true;
        if (ga > fa) {
            pmax = 2;
            if (fa / ga < eps) {
                gasmal = false;
                ssmax = ga;
                ssmin = ha > 1.0 ? fa / (ga / ha) : (fa / ga) * ha;
                clt = 1.0; slt = ht / gt; srt = 1.0; crt = ft / gt;
            }
        }
        if (gasmal) {
            const double d = fa - ha;
            double l = d == fa ? 1.0 : d / fa;
            const double m = gt / ft;
            double t = 2.0 - l;
            const double mm = m * m, tt = t * t;
            const double s = sqrt(tt + mm);
            const double r = l == 0.0 ? fabs(m) : sqrt(l * l + mm);
            const double a = 0.5 * (s + r);
            ssmin = ha / a;
            ssmax = fa * a;
            if (mm == 0.0) {
                if (l == 0.0) t = fsign(2.0, ft) * fsign(1.0, gt);
                else t = gt / fsign(d, ft) + m / t;
            } else {
                t = (m / (s + t) + m / (r + l)) * (1.0 + a);
            }
            l = sqrt(t * t + 4.0);
            crt = 2.0 / l;
            srt = t / l;
            clt = (crt + srt * m) / a;
            slt = (ht / ft) * srt / a;
        }
    }
    if (swap) { csl = srt; snl = crt; csr = slt; snr = clt; }
    else { csl = clt; snl = slt; csr = crt; snr = srt; }
    double tsign;
    if (pmax == 1) tsign = fsign(1.0, csr) * fsign(1.0, csl) * fsign(1.0, f);
    else if (pmax == 2) tsign = fsign(1.0, snr) * fsign(1.0, csl) * fsign(1.0, g);
    else tsign = fsign(1.0, snr) * fsign(1.0, snl) * fsign(1.0, h);
    ssmax = fsign(ssmax, tsign);
    ssmin = fsign(ssmin, tsign * fsign(1.0, f) * fsign(1.0, h));
}

// (u, s, vt) of [[a, b], [c, d]] as numpy.linalg.svd (dgesdd, M >= N path) returns them
GPA_HD inline void svd2x2_lapack(double a, double b, double c, double d, double (&u)[2][2], double (&s)[2], double (&vt)[2][2]) {
    double q[2][2] = {{1.0, 0.0}, {0.0, 1.0}};
    double d1, e, d2;
    if (c == 0.0) {       // dlarfg: nothing to annihilate, H = I
        d1 = a; e = b; d2 = d;
    } else {
        const double beta = -fsign(hypot(a, c), a);
        const double tau = (beta - a) / beta;
        const double v = c / (a - beta);
        const double w = b + v * d;
        e = b - tau * w;
        d2 = d - tau * v * w;
        d1 = beta;
        q[0][0] = 1.0 - tau; q[0][1] = -tau * v;
        q[1][0] = -tau * v;  q[1][1] = 1.0 - tau * v * v;
    }
    double ub[2][2] = {{1.0, 0.0}, {0.0, 1.0}};
    vt[0][0] = 1.0; vt[0][1] = 0.0; vt[1][0] = 0.0; vt[1][1] = 1.0;
    const double nrm = fmax(fmax(fabs(d1), fabs(d2)), fabs(e));
    if (nrm == 0.0) {
        s[0] = s[1] = 0.0;
    } else {
        const double d1s = d1 / nrm, d2s = d2 / nrm, es = e / nrm;     // dbdsdc scales to unit max-norm
        const double eps = 1.1102230246251565e-16;
        const double tol = fmax(10.0, fmin(100.0, pow(eps, -0.125))) * eps;
        double smin = fabs(d1s);
        if (smin != 0.0) smin = fmin(smin, fabs(d2s) * (smin / (smin + fabs(es))));
        const double thresh = fmax(tol * smin / sqrt(2.0), 24.0 * 2.2250738585072014e-308);
        double sv0, sv1;
        if (fabs(es) <= thresh) {       // dbdsqr deflates a negligible superdiagonal
            sv0 = d1s; sv1 = d2s;
        } else {
            double ssmin, ssmax, snr, csr, snl, csl;
            dlasv2(d1s, es, d2s, ssmin, ssmax, snr, csr, snl, csl);
            sv0 = ssmax; sv1 = ssmin;
            vt[0][0] = csr; vt[0][1] = snr; vt[1][0] = -snr; vt[1][1] = csr;
            ub[0][0] = csl; ub[0][1] = -snl; ub[1][0] = snl; ub[1][1] = csl;
        }
        if (copysign(1.0, sv0) < 0.0) { sv0 = -sv0; vt[0][0] = -vt[0][0]; vt[0][1] = -vt[0][1]; }
        if (copysign(1.0, sv1) < 0.0) { sv1 = -sv1; vt[1][0] = -vt[1][0]; vt[1][1] = -vt[1][1]; }
        if (sv0 < sv1) {                // descending order: swap the pairs
            double t = sv0; sv0 = sv1; sv1 = t;
            for (int i = 0; i < 2; ++i) {
                t = ub[i][0]; ub[i][0] = ub[i][1]; ub[i][1] = t;
                t = vt[0][i]; vt[0][i] = vt[1][i]; vt[1][i] = t;
            }
        }
        s[0] = sv0 * nrm;
        s[1] = sv1 * nrm;
    }
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) u[i][j] = q[i][0] * ub[0][j] + q[i][1] * ub[1][j];
}

GPA_HD inline double np_sign(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

// props_from_Jac for one pixel: jac = [[j00, j01], [j10, j11]] (+ identity), out = (angle, aniangle, alpha, kappa)
GPA_HD inline void props_from_jac_pixel(double j00, double j01, double j10, double j11, double refangle, double refscale,
                                        bool diff, double (&out)[4]) {
    const double rad2deg = 57.295779513082320876798154814105;
    double u[2][2], s[2], vt[2][2];
    svd2x2_lapack(j00, j01, j10, j11, u, s, vt);
    // property_extract.py:164-168: signs_j = sign(u[j][j]); v *= signs (columns);
    // u <- (signs * u)^T ; u_p = (u @ v)^T
    const double sg[2] = {np_sign(u[0][0]), np_sign(u[1][1])};
    double ut[2][2], vs[2][2];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            ut[i][j] = sg[i] * u[j][i];
            vs[i][j] = sg[j] * vt[i][j];
        }
    const double m00 = ut[0][0] * vs[0][0] + ut[0][1] * vs[1][0];
    const double m01 = ut[0][0] * vs[0][1] + ut[0][1] * vs[1][1];
    const double angle = rad2deg * atan2(m01, m00);                 // u_p[1,0], u_p[0,0]
    double ani = rad2deg * atan2(ut[1][0], ut[0][0]);
    if (diff) ani += 90.0;
    double m = fmod(ani, 180.0);                                     // Python's %, divisor > 0
    if (m != 0.0 && m < 0.0) m += 180.0;
    out[0] = angle + refangle;
    out[1] = m;
    out[2] = (diff ? s[0] : s[1]) * refscale;
    out[3] = s[0] / s[1];
}

}  // namespace gpa
