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
    // pivot column c and the other column o, selected ONCE (the selects inside the three loops below used to cost more
    // issue slots than the arithmetic)
    double c[DM], o[DM];
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        c[i] = swap ? a1[i] : a0[i];
        o[i] = swap ? a0[i] : a1[i];
    }
    // Gram-Schmidt with the UNNORMALISED pivot column c (R = [[f, g], [0, h]], f^2 = |c|^2, g = c.o / f, h^2 = |o - (c.o / f^2) c|^2):
    // the generic pixel costs two divisions and no square root, which keeps this streaming kernel on the HBM roofline
    // instead of the fp64 divide / square-root pipe (r1: 0.33 of the HBM peak with the normalised form).
    const double inv_f2 = 1.0 / f2;
    double co = 0.0, cy[NRHS];
#pragma unroll
    for (int k = 0; k < NRHS; ++k) cy[k] = 0.0;
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        if (i < d) {
            co = fma(c[i], o[i], co);
#pragma unroll
            for (int k = 0; k < NRHS; ++k) cy[k] = fma(c[i], y[k][i], cy[k]);
        }
    }
    const double gs = co * inv_f2;            // g / f
    double h2 = 0.0, z2h[NRHS];   // second column orthogonalised against the first; z2h = h * z2
#pragma unroll
    for (int k = 0; k < NRHS; ++k) z2h[k] = 0.0;
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        if (i < d) {
            const double e = o[i] - gs * c[i];
            h2 = fma(e, e, h2);
#pragma unroll
            for (int k = 0; k < NRHS; ++k) z2h[k] = fma(e, y[k][i], z2h[k]);
        }
    }
    // gelsd's rank rule s2 > eps s1 for the singular values of R: s1^2 + s2^2 = t = |a0|^2 + |a1|^2, s1 s2 = det = f h.
    // s1^2 <= t, so det > eps t proves full rank without a square root; only the (rare) pixels that fail this test take
    // the exact evaluation.
    const double eps = 2.220446049250313e-16;
    const double t = n0 + n1;
    const double det2 = f2 * h2;
    bool full_rank = det2 > (eps * t) * (eps * t);
    if (!full_rank && h2 > 0.0) {
        const double det = sqrt(det2);
        const double disc = sqrt(fmax(t * t - 4.0 * det2, 0.0));
        const double s1 = sqrt(0.5 * (t + disc));
        full_rank = det / s1 > eps * s1;
    }
    if (full_rank) {
        const double inv_h2 = 1.0 / h2;
#pragma unroll
        for (int k = 0; k < NRHS; ++k) {
            const double u1 = z2h[k] * inv_h2;                    // z2 / h
            const double u0 = fma(-co, u1, cy[k]) * inv_f2;      // (z1 - g u1) / f
            x[k][0] = swap ? u1 : u0;
            x[k][1] = swap ? u0 : u1;
        }
    } else {                                   // rank 1: minimum-norm solution of [f g] u = z1, i.e. u = (c.y) (f^2, c.o) / (f^4 + (c.o)^2)
        const double inv_nn = 1.0 / fma(f2, f2, co * co);
#pragma unroll
        for (int k = 0; k < NRHS; ++k) {
            const double u0 = f2 * cy[k] * inv_nn;
            const double u1 = co * cy[k] * inv_nn;
            x[k][0] = swap ? u1 : u0;
            x[k][1] = swap ? u0 : u1;
        }
    }
}

GPA_HD inline double wrap_pi(double v) {
    // (v + pi) mod 2 pi - pi with a non-negative modulo (mathtools.py:72-75): quotient by multiplication, remainder by fma
    // (closer to numpy's exact fmod remainder than t * 2 pi, and no fp64 division)
    const double two_pi = 6.283185307179586476925286766559;
    const double inv_two_pi = 0.15915494309189533576888376337251;
    const double pi = 3.141592653589793238462643383279;
    const double s = v + pi;
    double r = fma(-floor(s * inv_two_pi), two_pi, s);
    if (r < 0.0) r += two_pi;
    else if (r >= two_pi) r -= two_pi;
    return r - pi;
}

}  // namespace gpa
