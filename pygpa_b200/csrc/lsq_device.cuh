// Per-pixel weighted least squares shared by K3 (lstsq.cu) and the property extraction (props.cu).
//
// Minimise |a0 x0 + a1 x1 - y| over x for NRHS right-hand sides that share the d x 2 matrix
// [a0 a1] (rows already multiplied by the pixel's weights): column-pivoted Gram-Schmidt QR in
// registers (error ~ cond * eps like LAPACK's SVD path) with gelsd's rank rule — singular values
// <= eps * s_max are dropped and the minimum-norm solution is returned (all-zero rows -> 0),
// which is what np.linalg.lstsq gives myweighed_lstsq (geometric_phase_analysis.py:97-113).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define GPA_HD __host__ __device__
#else
#define GPA_HD
#endif

namespace gpa {

constexpr int kMaxD = 8;

// DM = compile-time capacity of the row arrays (d <= DM): the kernels instantiate DM = 3 — the
// three primary k-vectors of every pipeline in the reference — next to the general DM = kMaxD, which
// halves their register footprint.
template <int NRHS, int DM>
GPA_HD inline void lsq_solve2(const double (&a0)[DM], const double (&a1)[DM],
                                           const double (&y)[NRHS][DM], int d, double (&x)[NRHS][2]) {
#pragma unroll
    for (int k = 0; k < NRHS; ++k) x[k][0] = x[k][1] = 0.0;
    double n0 = 0.0, n1 = 0.0;
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        if (i < d) {
            n0 = fma(a0[i], a0[i], n0);
            n1 = fma(a1[i], a1[i], n1);
        }
    }
    const bool swap = n1 > n0;   // column pivoting: the larger column first
    const double f2 = swap ? n1 : n0;
    if (!(f2 > 0.0)) return;
    const double f = sqrt(f2);
    const double inv_f = 1.0 / f;
    double q[DM];
    double g = 0.0, z1[NRHS];
#pragma unroll
    for (int k = 0; k < NRHS; ++k) z1[k] = 0.0;
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        if (i < d) {
            q[i] = (swap ? a1[i] : a0[i]) * inv_f;
            g = fma(q[i], swap ? a0[i] : a1[i], g);
#pragma unroll
            for (int k = 0; k < NRHS; ++k) z1[k] = fma(q[i], y[k][i], z1[k]);
        }
    }
    double h2 = 0.0, z2h[NRHS];   // second column orthogonalised against the first; z2h = h * z2
#pragma unroll
    for (int k = 0; k < NRHS; ++k) z2h[k] = 0.0;
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        if (i < d) {
            const double e = (swap ? a0[i] : a1[i]) - g * q[i];
            h2 = fma(e, e, h2);
#pragma unroll
            for (int k = 0; k < NRHS; ++k) z2h[k] = fma(e, y[k][i], z2h[k]);
        }
    }
    const double h = sqrt(h2);
    // singular values of [[f, g], [0, h]]
    const double t = f * f + g * g + h * h;
    const double det = f * h;
    const double disc = sqrt(fmax(t * t - 4.0 * det * det, 0.0));
    const double s1 = sqrt(0.5 * (t + disc));
    const double s2 = det / s1;
    const bool full_rank = s2 > 2.220446049250313e-16 * s1;
    const double nn = f * f + g * g;
#pragma unroll
    for (int k = 0; k < NRHS; ++k) {
        double u0, u1;
        if (full_rank) {
            u1 = z2h[k] / h2;                 // z2 / h
            u0 = (z1[k] - g * u1) * inv_f;
        } else {                               // rank 1: minimum-norm solution of [f g] u = z1
            u0 = f * z1[k] / nn;
            u1 = g * z1[k] / nn;
        }
        x[k][0] = swap ? u1 : u0;
        x[k][1] = swap ? u0 : u1;
    }
}

GPA_HD inline double wrap_pi(double v) {
    // (v + pi) mod 2 pi - pi with a non-negative modulo: mathtools.py:72-75
    const double two_pi = 6.283185307179586476925286766559;
    const double pi = 3.141592653589793238462643383279;
    double t = (v + pi) / two_pi;
    t -= floor(t);
    return t * two_pi - pi;
}

}  // namespace gpa
