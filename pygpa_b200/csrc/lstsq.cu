// K3 — per-pixel phase -> displacement least squares (float64), sm_100a.
//
// Reference semantics: myweighed_lstsq (pyGPA/geometric_phase_analysis.py:97-113), the three
// branches of reconstruct_u_inv (157-193) and the gradient assembly of
// reconstruct_u_inv_from_phases (196-237).  One thread per pixel; the d x 2 system
// (w_i K_i) x = w_i b_i is solved by a column-pivoted Gram-Schmidt QR in registers (error ~
// cond * eps, like LAPACK's SVD path, not cond^2 * eps like the normal equations) with
// gelsd's rank rule: singular values <= eps * s_max are dropped and the minimum-norm solution
// returned (all-zero weights -> 0).  HBM-bound: (2d + 2) doubles per pixel.
#include "common.cuh"
#include "lsq_device.cuh"

namespace gpa {

struct LsqParams {
    const double* src;    // phases / gradients / unwrapped phases
    const double* w;      // weights (d, wn, wm) or null
    const double* means;  // per-plane mean to subtract (d) or null
    double* out;          // (2, n, m)
    long long base, ps, rs, cs, doff;   // b_i(r, c) = src[base + i*ps + r*rs + c*cs (+ doff)] ...
    int do_wrap, d, n, m, wn, wm;
    double K[kMaxD][2];   // 2 pi k
    double P[2][kMaxD];   // unweighted / two-k branch: x = P b
    int use_matrix;
};

// DIFF / WRAP / MEANS / MATRIX are the launch-uniform switches of LsqParams as template parameters, FULL = (d == DM): the
// kernel body has no uniform branches left, so ALL of a pixel's loads (DM sources, DM forward neighbours, DM weights) are
// issued back to back before the first one is consumed.  With the runtime switches the compiler kept every load inside its
// branch — four serialised DRAM round trips per thread, 59 % of the stall samples on long_scoreboard at 0.45 of the HBM rate
// (profiles/r02f_k_lstsq_ncu_summary.json).
// A thread solves kLsqPPT pixels, 64 columns apart (a CTA covers 4 rows x 64 kLsqPPT columns): with DM = 3 two pixels per
// thread double the loads in flight per warp at 72 registers; the general DM = 8 form keeps one.
template <int DM>
struct LsqShape { static constexpr int PPT = DM <= 3 ? 2 : 1; };

template <int DM, bool FULL, bool DIFF, bool WRAP, bool MEANS, bool MATRIX>
__global__ void __launch_bounds__(256) k_lstsq(const LsqParams p) {
    constexpr int kLsqPPT = LsqShape<DM>::PPT;
    const int c0 = blockIdx.x * (64 * kLsqPPT) + (threadIdx.x & 63);
    const int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= p.n) return;
    const int d = FULL ? DM : p.d;
    double s0[kLsqPPT][DM], s1[kLsqPPT][DM], w[kLsqPPT][DM];
#pragma unroll
    for (int q = 0; q < kLsqPPT; ++q) {
        const int c = c0 + 64 * q;
        const long long o = p.base + (long long)r * p.rs + (long long)c * p.cs;
#pragma unroll
        for (int i = 0; i < DM; ++i) {
            s0[q][i] = s1[q][i] = w[q][i] = 0.0;
            if (i < d && c < p.m) {
                s0[q][i] = p.src[o + i * p.ps];
                if (DIFF) s1[q][i] = p.src[o + i * p.ps + p.doff];
                if (!MATRIX) w[q][i] = p.w[(long long)i * p.wn * p.wm + (long long)r * p.wm + c];
            }
        }
    }
    const size_t np = (size_t)p.n * p.m;
#pragma unroll
    for (int q = 0; q < kLsqPPT; ++q) {
        const int c = c0 + 64 * q;
        if (c >= p.m) continue;
        double b[DM];
#pragma unroll
        for (int i = 0; i < DM; ++i) {
            double v = DIFF ? s1[q][i] - s0[q][i] : s0[q][i];
            if (WRAP) v = wrap_pi(v);
            if (MEANS && i < d) v -= p.means[i];
            b[i] = v;
        }
        double x0 = 0.0, x1 = 0.0;
        if (MATRIX) {
#pragma unroll
            for (int i = 0; i < DM; ++i) {
                if (i < d) {
                    x0 = fma(p.P[0][i], b[i], x0);
                    x1 = fma(p.P[1][i], b[i], x1);
                }
            }
        } else {
            double a0[DM], a1[DM], y[1][DM], x[1][2];
#pragma unroll
            for (int i = 0; i < DM; ++i) {
                a0[i] = w[q][i] * p.K[i][0];
                a1[i] = w[q][i] * p.K[i][1];
                y[0][i] = w[q][i] * b[i];
            }
            lsq_solve2<1, DM>(a0, a1, y, d, x);
            x0 = x[0][0];
            x1 = x[0][1];
        }
        p.out[(size_t)r * p.m + c] = x0;
        p.out[np + (size_t)r * p.m + c] = x1;
    }
}

template <int DM, bool FULL>
static void launch_lstsq(const LsqParams& p, cudaStream_t st) {
    const dim3 grid(ceil_div(p.m, 64 * LsqShape<DM>::PPT), ceil_div(p.n, 4));
    const bool diff = p.doff != 0, wrap = p.do_wrap != 0, means = p.means != nullptr, matrix = p.use_matrix != 0;
    // the combinations the entry point can produce: plain (+ means) with either solver, wrapped differences / pre-differences
    if (diff) {            // GPA_LSQ_SRC_DIFF0 / DIFF1: always wrapped, never mean-subtracted
        if (matrix) k_lstsq<DM, FULL, true, true, false, true><<<grid, 256, 0, st>>>(p);
        else k_lstsq<DM, FULL, true, true, false, false><<<grid, 256, 0, st>>>(p);
    } else if (wrap) {     // GPA_LSQ_SRC_PREDIFF0 / PREDIFF1
        if (matrix) k_lstsq<DM, FULL, false, true, false, true><<<grid, 256, 0, st>>>(p);
        else k_lstsq<DM, FULL, false, true, false, false><<<grid, 256, 0, st>>>(p);
    } else if (means) {    // GPA_LSQ_SRC_PLAIN with the per-plane mean removed
        if (matrix) k_lstsq<DM, FULL, false, false, true, true><<<grid, 256, 0, st>>>(p);
        else k_lstsq<DM, FULL, false, false, true, false><<<grid, 256, 0, st>>>(p);
    } else {
        if (matrix) k_lstsq<DM, FULL, false, false, false, true><<<grid, 256, 0, st>>>(p);
        else k_lstsq<DM, FULL, false, false, false, false><<<grid, 256, 0, st>>>(p);
    }
}

// per-plane mean of a (d, n) array, deterministic two-stage reduction
__global__ void __launch_bounds__(256) k_plane_partial(const double* __restrict__ src, double* __restrict__ part, size_t n) {
    __shared__ double sh[256];
    const int plane = blockIdx.y;
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s += src[(size_t)plane * n + i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)plane * gridDim.x + blockIdx.x] = sh[0];
}

__global__ void k_plane_mean_final(const double* __restrict__ part, double* __restrict__ means, int nblk, double inv_n) {
    __shared__ double sh[256];
    const int plane = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 256) s += part[(size_t)plane * nblk + i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) means[plane] = sh[0] * inv_n;
}

// out[r, c] = sqrt(sum_i w[i, r, c]^2)   (np.linalg.norm(weights, axis=0), geometric_phase_analysis.py:240)
__global__ void __launch_bounds__(256) k_norm_axis0(const double* __restrict__ w, double* __restrict__ out, int d, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) {
            const double v = w[(size_t)k * n + i];
            s = fma(v, v, s);
        }
        out[i] = sqrt(s);
    }
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_lstsq_workspace_bytes(int d, size_t* bytes) {
    GPA_REQUIRE(bytes && d >= 1 && d <= kMaxD, "d must be in [1, %d]", kMaxD);
    *bytes = (size_t)(d * 1024 + kMaxD) * sizeof(double) + 512;
    return GPA_OK;
}

extern "C" int gpa_lstsq_u(const double* src, int src_kind, const double* w, int wn, int wm,
                           const double* kvecs /*host (d,2), cycles/pixel*/, int d, int N, int M,
                           int solver, const double* matrix /*host (2,d) or null*/, int subtract_mean,
                           double* out, void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(src && kvecs && out, "null pointer argument");
    GPA_REQUIRE(d >= 1 && d <= kMaxD, "d must be in [1, %d] (got %d)", kMaxD, d);
    GPA_REQUIRE(N >= 1 && M >= 1, "bad shape");
    GPA_REQUIRE(src_kind >= GPA_LSQ_SRC_PLAIN && src_kind <= GPA_LSQ_SRC_PREDIFF1, "bad src_kind %d", src_kind);
    GPA_REQUIRE(solver == GPA_LSQ_WEIGHTED || solver == GPA_LSQ_MATRIX, "bad solver %d", solver);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LsqParams p;
    std::memset(&p, 0, sizeof(p));
    p.src = src; p.w = w; p.out = out; p.d = d;
    const long long NM = (long long)N * M;
    switch (src_kind) {
        case GPA_LSQ_SRC_PLAIN: p.ps = NM; p.rs = M; p.cs = 1; p.n = N; p.m = M; break;
        case GPA_LSQ_SRC_DIFF1: p.ps = NM; p.rs = M; p.cs = 1; p.doff = 1; p.do_wrap = 1; p.n = N; p.m = M - 1; break;
        case GPA_LSQ_SRC_DIFF0: p.ps = NM; p.rs = M; p.cs = 1; p.doff = M; p.do_wrap = 1; p.n = N - 1; p.m = M; break;
        case GPA_LSQ_SRC_PREDIFF0: p.ps = 2 * NM; p.rs = 2LL * M; p.cs = 2; p.base = 0; p.do_wrap = 1; p.n = N; p.m = M - 1; break;
        case GPA_LSQ_SRC_PREDIFF1: p.ps = 2 * NM; p.rs = 2LL * M; p.cs = 2; p.base = 1; p.do_wrap = 1; p.n = N - 1; p.m = M; break;
    }
    GPA_REQUIRE(p.n >= 1 && p.m >= 1, "frame too small for a difference");
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < d; ++i) {
        p.K[i][0] = two_pi * kvecs[2 * i];
        p.K[i][1] = two_pi * kvecs[2 * i + 1];
    }
    if (solver == GPA_LSQ_MATRIX) {
        GPA_REQUIRE(matrix != nullptr, "matrix is null");
        p.use_matrix = 1;
        for (int i = 0; i < d; ++i) {
            p.P[0][i] = matrix[i];
            p.P[1][i] = matrix[d + i];
        }
    } else {
        GPA_REQUIRE(w != nullptr, "weights are null");
        GPA_REQUIRE(wn >= p.n && wm >= p.m, "weights (%d x %d) smaller than the solve grid (%d x %d)", wn, wm, p.n, p.m);
        p.wn = wn; p.wm = wm;
    }
    if (subtract_mean) {
        GPA_REQUIRE(src_kind == GPA_LSQ_SRC_PLAIN, "mean subtraction applies to plain sources only");
        size_t need = 0;
        gpa_lstsq_workspace_bytes(d, &need);
        if (!ws || ws_bytes < need) {
            set_error("workspace too small (%zu < %zu)", ws_bytes, need);
            return GPA_ERR_WORKSPACE;
        }
        Arena a(ws, ws_bytes);
        double* part = a.take<double>((size_t)d * 1024);
        double* means = a.take<double>(kMaxD);
        const int nblk = (int)((NM + 256 * 8 - 1) / (256 * 8) < 1024 ? (NM + 256 * 8 - 1) / (256 * 8) : 1024);
        {
            KernelTimer t("k_plane_mean", st);
            k_plane_partial<<<dim3(nblk, d), 256, 0, st>>>(src, part, (size_t)NM);
            k_plane_mean_final<<<d, 256, 0, st>>>(part, means, nblk, 1.0 / (double)NM);
        }
        p.means = means;
    }
    {
        KernelTimer t("k_lstsq", st);
        if (d == 3) launch_lstsq<3, true>(p, st);
        else if (d < 3) launch_lstsq<3, false>(p, st);
        else launch_lstsq<kMaxD, false>(p, st);
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_norm_axis0(const double* w, int d, size_t n, double* out, void* stream) {
    GPA_REQUIRE(w && out && d >= 1, "bad argument");
    if (n == 0) return GPA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    KernelTimer t("k_norm_axis0", st);
    k_norm_axis0<<<(unsigned)blocks, 256, 0, st>>>(w, out, d, n);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}
