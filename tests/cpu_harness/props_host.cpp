// Host build of the per-pixel device code of K3 / K5 (pygpa_b200/csrc/lsq_device.cuh,
// props_device.cuh) for tests/test_props_host.py: the same source the CUDA kernels compile, run on
// the CPU against numpy.linalg.svd and the oracle.  Test infrastructure, not a fallback: nothing in
// pygpa_b200 loads this.
#include "../../pygpa_b200/csrc/lsq_device.cuh"
#include "../../pygpa_b200/csrc/props_device.cuh"

extern "C" {

// n matrices (n, 2, 2) -> u (n, 2, 2), s (n, 2), vt (n, 2, 2)
void host_svd2x2(const double* a, long n, double* u, double* s, double* vt) {
    for (long i = 0; i < n; ++i) {
        double uu[2][2], ss[2], vv[2][2];
        gpa::svd2x2_lapack(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3], uu, ss, vv);
        for (int k = 0; k < 4; ++k) {
            u[4 * i + k] = uu[k / 2][k % 2];
            vt[4 * i + k] = vv[k / 2][k % 2];
        }
        s[2 * i] = ss[0];
        s[2 * i + 1] = ss[1];
    }
}

// jac (n, 2, 2) -> props (4, n)
void host_props(const double* jac, long n, double refangle, double refscale, int diff, int add_identity, double* props) {
    const double id = add_identity ? 1.0 : 0.0;
    for (long i = 0; i < n; ++i) {
        double out[4];
        gpa::props_from_jac_pixel(jac[4 * i] + id, jac[4 * i + 1], jac[4 * i + 2], jac[4 * i + 3] + id, refangle, refscale,
                                  diff != 0, out);
        for (int k = 0; k < 4; ++k) props[k * n + i] = out[k];
    }
}

// per pixel: rows (w_i K_i), right-hand sides w_i b_i[rhs]; a (n, d, 2), y (n, 2, d) -> x (n, 2, 2) [rhs][component]
void host_lsq2(const double* a, const double* y, long n, int d, double* x) {
    for (long p = 0; p < n; ++p) {
        double a0[gpa::kMaxD], a1[gpa::kMaxD], yy[2][gpa::kMaxD], xx[2][2];
        for (int i = 0; i < d; ++i) {
            a0[i] = a[(p * d + i) * 2];
            a1[i] = a[(p * d + i) * 2 + 1];
            yy[0][i] = y[(p * 2 + 0) * d + i];
            yy[1][i] = y[(p * 2 + 1) * d + i];
        }
        gpa::lsq_solve2<2, gpa::kMaxD>(a0, a1, yy, d, xx);
        for (int k = 0; k < 4; ++k) x[4 * p + k] = xx[k / 2][k % 2];
    }
}

}  // extern "C"
