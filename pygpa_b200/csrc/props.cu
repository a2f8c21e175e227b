// K5 — local lattice properties from the sweep's phase-gradient maps (float64), sm_100a.
//
// Reference semantics: phasegradient2J (pyGPA/property_extract.py:69-101): isotropic
// re-referencing of the per-peak phase gradients, two per-pixel weighted least-squares solves
// (myweighed_lstsq, geometric_phase_analysis.py:97-113) -> J (N, M, 2, 2); props_from_Jac
// (property_extract.py:137-178): per-pixel 2x2 SVD -> lattice angle, anisotropy angle, scale,
// anisotropy magnitude.  Both are HBM-bound streaming kernels, one thread per pixel.
//
// props_from_Jac is NOT invariant under the sign freedom of the SVD: its sign normalisation
// multiplies the COLUMNS of V^T, so the result depends on the relative sign LAPACK gives the two
// singular-vector pairs.  The kernel therefore restates LAPACK's dgesdd for a 2x2 matrix
// (dgebrd Householder reflector -> dbdsqr deflation test -> dlasv2 -> sign fix -> sort) rather
// than using a textbook closed form; oracle/props_numpy.py:svd2x2_lapack is the same restatement
// in NumPy and is checked against numpy.linalg.svd.
#include "common.cuh"
#include "lsq_device.cuh"
#include "props_device.cuh"

namespace gpa {

struct Grad2JParams {
    const double* grads;    // (d, N, M, 2)
    const double* w;        // (d, wn, wm)
    double* J;              // (N, M, 2, 2)
    int d, N, M, wn, wm, do_wrap, add_identity;
    int order[kMaxD];       // source plane of solve row i
    double K[kMaxD][2];     // solve matrix rows (2 pi (k + dk))
    double sub[kMaxD][2];   // subtracted from the gradient before wrapping (2 pi dk)
    double nmperpixel;
};

// FULL (d == DM) and WRAP as template parameters: no launch-uniform branch is left around the loads, so the DM weights and DM
// gradient pairs of a pixel are all in flight before the first is used (as in k_lstsq: the runtime switches serialised them).
template <int DM, bool FULL, bool WRAP>
__global__ void __launch_bounds__(256) k_grad2J(const Grad2JParams p) {
    const int c = blockIdx.x * 64 + (threadIdx.x & 63);
    const int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= p.N || c >= p.M) return;
    const int d = FULL ? DM : p.d;
    const size_t pix = (size_t)r * p.M + c;
    const size_t npix = (size_t)p.N * p.M;
    double wv[DM];
    double2 gv[DM];
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        wv[i] = 0.0;
        gv[i] = make_double2(0.0, 0.0);
        if (i < d) {
            wv[i] = p.w[(size_t)i * p.wn * p.wm + (size_t)r * p.wm + c];
            gv[i] = *reinterpret_cast<const double2*>(p.grads + ((size_t)p.order[i] * npix + pix) * 2);
        }
    }
    double a0[DM], a1[DM], y[2][DM], x[2][2];
#pragma unroll
    for (int i = 0; i < DM; ++i) {
        double b0 = gv[i].x - p.sub[i][0], b1 = gv[i].y - p.sub[i][1];
        if (WRAP) {
            b0 = wrap_pi(b0);
            b1 = wrap_pi(b1);
        }
        a0[i] = wv[i] * p.K[i][0];
        a1[i] = wv[i] * p.K[i][1];
        y[0][i] = wv[i] * b0;
        y[1][i] = wv[i] * b1;
    }
    lsq_solve2<2, DM>(a0, a1, y, d, x);
    // J[i][j] = d u_i / d x_j: right-hand side j (gradient along axis j) gives column j
    const double id = p.add_identity ? 1.0 : 0.0;
    double2* out = reinterpret_cast<double2*>(p.J + pix * 4);
    out[0] = make_double2(x[0][0] / p.nmperpixel + id, x[1][0] / p.nmperpixel);
    out[1] = make_double2(x[0][1] / p.nmperpixel, x[1][1] / p.nmperpixel + id);
}

struct PropsParams {
    const double* jac;     // (N, M, 2, 2)
    double* props;         // (4, npix): angle, aniangle, alpha, kappa
    size_t npix;
    double refangle, refscale;
    int diff, add_identity;
};

__global__ void __launch_bounds__(256) k_props_from_jac(const PropsParams p) {
    for (size_t pix = (size_t)blockIdx.x * 256 + threadIdx.x; pix < p.npix; pix += (size_t)gridDim.x * 256) {
        const double2* in = reinterpret_cast<const double2*>(p.jac + pix * 4);
        const double2 r0 = in[0], r1 = in[1];
        const double id = p.add_identity ? 1.0 : 0.0;
        double out[4];
        props_from_jac_pixel(r0.x + id, r0.y, r1.x, r1.y + id, p.refangle, p.refscale, p.diff != 0, out);
#pragma unroll
        for (int k = 0; k < 4; ++k) p.props[k * p.npix + pix] = out[k];
    }
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_phasegradient_to_j(const double* grads, const double* weights, int wn, int wm,
                                      const double* K /*host (d,2)*/, const double* sub /*host (d,2) or null*/,
                                      const int* order /*host d or null*/, int do_wrap, int d, int N, int M,
                                      double nmperpixel, int add_identity, double* J, void* stream) {
    GPA_REQUIRE(grads && weights && K && J, "null pointer argument");
    GPA_REQUIRE(d >= 1 && d <= kMaxD, "d must be in [1, %d] (got %d)", kMaxD, d);
    GPA_REQUIRE(N >= 1 && M >= 1, "bad shape");
    GPA_REQUIRE(wn >= N && wm >= M, "weights (%d x %d) smaller than the frame (%d x %d)", wn, wm, N, M);
    GPA_REQUIRE(nmperpixel != 0.0, "nmperpixel must be non-zero");
    Grad2JParams p;
    std::memset(&p, 0, sizeof(p));
    p.grads = grads; p.w = weights; p.J = J; p.d = d; p.N = N; p.M = M; p.wn = wn; p.wm = wm;
    p.do_wrap = do_wrap; p.add_identity = add_identity; p.nmperpixel = nmperpixel;
    for (int i = 0; i < d; ++i) {
        p.order[i] = order ? order[i] : i;
        GPA_REQUIRE(p.order[i] >= 0 && p.order[i] < d, "order[%d] = %d out of range", i, p.order[i]);
        p.K[i][0] = K[2 * i]; p.K[i][1] = K[2 * i + 1];
        if (sub) { p.sub[i][0] = sub[2 * i]; p.sub[i][1] = sub[2 * i + 1]; }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KernelTimer t("k_grad2J", st);
    dim3 grid(ceil_div(M, 64), ceil_div(N, 4));
    if (d == 3) {
        if (do_wrap) k_grad2J<3, true, true><<<grid, 256, 0, st>>>(p);
        else k_grad2J<3, true, false><<<grid, 256, 0, st>>>(p);
    } else if (d < 3) {
        if (do_wrap) k_grad2J<3, false, true><<<grid, 256, 0, st>>>(p);
        else k_grad2J<3, false, false><<<grid, 256, 0, st>>>(p);
    } else {
        if (do_wrap) k_grad2J<kMaxD, false, true><<<grid, 256, 0, st>>>(p);
        else k_grad2J<kMaxD, false, false><<<grid, 256, 0, st>>>(p);
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_props_from_jac(const double* jac, size_t npix, double refangle, double refscale, int diff,
                                  int add_identity, double* props, void* stream) {
    GPA_REQUIRE(jac && props, "null pointer argument");
    if (npix == 0) return GPA_OK;
    PropsParams p;
    p.jac = jac; p.props = props; p.npix = npix; p.refangle = refangle; p.refscale = refscale;
    p.diff = diff != 0; p.add_identity = add_identity != 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    size_t blocks = (npix + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    KernelTimer t("k_props_from_jac", st);
    k_props_from_jac<<<(unsigned)blocks, 256, 0, st>>>(p);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}
