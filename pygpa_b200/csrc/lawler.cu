// K4 — Lawler-Fujita: fixed-point inversion of the displacement field and cubic-spline
// resampling (float64), sm_100a.
//
// Reference semantics: invert_u_overlap (pyGPA/geometric_phase_analysis.py:262-300) and
// undistort_image (:935-974), i.e. scipy.ndimage.map_coordinates(order=3) in mode='nearest'
// (inversion) and mode='constant', cval=0 (final resample).  SciPy's rules are restated in
// oracle/spline_restatement.py; in short: cubic B-spline, pole sqrt(3)-2, gain 6 per axis;
// 'nearest' = 12-sample edge padding + reflect initial conditions + tap indices clamped;
// 'constant' = mirror initial conditions + mirrored taps + 0 outside [0, n-1].
//
// What is different from the reference: it re-runs the spline prefilter of the unchanged u in
// each of its 72 map_coordinates calls; here the prefilter runs once, both components are
// interleaved so one 16-byte load fetches both coefficients of a tap, and the whole fixed-point
// loop of a pixel runs in registers inside ONE kernel (no intermediate u_it ever touches HBM).
#include "common.cuh"

namespace gpa {

constexpr int kPad = 12;                 // scipy's _prepad_for_spline_filter
constexpr int kInitTerms = 96;           // |pole|^96 ~ 1e-55: the initial-condition sums are exact in double
__device__ __constant__ const double kPole = -0.26794919243112270647255365849413;   // sqrt(3) - 2

enum { kNearest = 0, kConstant = 1 };

// Prefilter along axis 0 of a (n x cols) array: one thread per column, coalesced across the warp.
// Source mapping (first pass only): s[i][c] = scale * src[clamp(i - pad)][clamp(c - pad)];
// later passes run in place (src == dst, pad == 0, src dims == dst dims).
struct PrefilterArgs {
    const double* src;
    double* dst;
    int n, cols;            // destination extent along the filter axis / across it
    int src_n, src_cols;    // source extent
    int pad;
    double scale;
    int boundary;           // kNearest -> reflect init, kConstant -> mirror init
    double gain;            // (1 - z)(1 - 1/z) = 6 per axis pass
};

constexpr int kSegLen = 32;     // outputs per thread along the filter axis (A/B at 2048^2, two fields: 64 -> 0.298 ms, 32 -> 0.271, 16 -> 0.271)
constexpr int kWarm = 40;       // warm-up samples of a segment's recursion: |pole|^40 = 1.3e-23

// The recursions c+[i] = s[i] + z c+[i-1] and c[i] = z (c[i+1] - c+[i]) forget their start after
// ~40 samples (|z| = 0.268), so each column is cut into segments of kSegLen outputs that run in
// parallel, each warming its recursion up on the kWarm samples before (after) the segment; only the
// first (last) segment uses scipy's exact boundary initialisation.  Threads of a warp sit on
// adjacent columns (coalesced); results are exact to double rounding.
__global__ void __launch_bounds__(128) k_prefilter_causal(const PrefilterArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.cols) return;
    const double z = kPole;
    const int n = a.n;
    int cs = c - a.pad;
    cs = cs < 0 ? 0 : (cs >= a.src_cols ? a.src_cols - 1 : cs);
    auto s = [&](int i) -> double {
        int r = i - a.pad;
        r = r < 0 ? 0 : (r >= a.src_n ? a.src_n - 1 : r);
        return a.gain * a.scale * a.src[(size_t)r * a.src_cols + cs];
    };
    double* out = a.dst + c;
    const size_t st = (size_t)a.cols;
    if (n == 1) {
        out[0] = a.scale * a.src[cs];
        return;
    }
    const int i0 = blockIdx.y * kSegLen;
    const int i1 = min(i0 + kSegLen, n);
    double prev;
    int i;
    if (i0 <= kWarm) {
        // exact causal initial condition (scipy _init_causal_reflect / _init_causal_mirror); segments that begin within
        // kWarm samples of the start run the recursion up from sample 0 without storing
        const int terms = n < kInitTerms ? n : kInitTerms;
        if (a.boundary == kNearest) {
            const double zn = pow(z, (double)n);
            double acc = 0.0, zi = 1.0;
            for (int t = 0; t < terms; ++t) {
                acc += zi * (s(t) + zn * s(n - 1 - t));
                zi *= z;
            }
            prev = acc * z / (1.0 - zn * zn) + s(0);
        } else {
            const double zn1 = pow(z, (double)(n - 1));
            double acc = s(0) + zn1 * s(n - 1), zi = z;
            const int last = (n - 1) < kInitTerms ? (n - 1) : kInitTerms;
            for (int t = 1; t < last; ++t) {
                acc += zi * (s(t) + zn1 * s(n - 1 - t));
                zi *= z;
            }
            prev = acc / (1.0 - zn1 * zn1);
        }
        if (i0 == 0) {
            out[0] = prev;
            i = 1;
        } else {
            for (int t = 1; t < i0; ++t) prev = fma(z, prev, s(t));
            i = i0;
        }
    } else {
        const int w0 = i0 - kWarm;           // > 0
        prev = s(w0) / (1.0 - z);
        for (int t = w0 + 1; t < i0; ++t) prev = fma(z, prev, s(t));
        i = i0;
    }
    for (; i < i1; ++i) {
        prev = fma(z, prev, s(i));
        out[i * st] = prev;
    }
}

// src = c+ (n x cols), dst = c (n x cols); never in place (a segment reads its successor's c+)
__global__ void __launch_bounds__(128) k_prefilter_anticausal(const double* __restrict__ cp, double* __restrict__ dst,
                                                                int n, int cols, int boundary) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const double z = kPole;
    const double* __restrict__ in = cp + c;
    double* out = dst + c;
    const size_t st = (size_t)cols;
    if (n == 1) {
        out[0] = in[0];
        return;
    }
    const int i0 = blockIdx.y * kSegLen;
    const int i1 = min(i0 + kSegLen, n);
    double nxt;
    int i;
    if (i1 + kWarm >= n) {
        // the true end is within reach: exact anti-causal initial condition, then run down to i1
        if (boundary == kNearest) nxt = in[(size_t)(n - 1) * st] * (z / (z - 1.0));
        else nxt = (z * in[(size_t)(n - 2) * st] + in[(size_t)(n - 1) * st]) * z / (z * z - 1.0);
        if (i1 == n) out[(size_t)(n - 1) * st] = nxt;
        for (int t = n - 2; t >= i1; --t) nxt = z * (nxt - in[t * st]);
        i = (i1 == n ? n - 2 : i1 - 1);
    } else {
        const int w1 = i1 + kWarm - 1;
        nxt = in[w1 * st] * (z / (z - 1.0));
        for (int t = w1 - 1; t >= i1; --t) nxt = z * (nxt - in[t * st]);
        i = i1 - 1;
    }
    for (; i >= i0; --i) {
        nxt = z * (nxt - in[i * st]);
        out[i * st] = nxt;
    }
}

__global__ void k_transpose_plain(const double* __restrict__ in, double* __restrict__ out, int rows, int cols,
                                  int out_elem, int out_off) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[((size_t)c * rows + r) * out_elem + out_off] = tile[threadIdx.x][i];
    }
}

// cubic B-spline weights for the fractional offset t in [0, 1)
__device__ __forceinline__ void bspline_weights(double t, double (&w)[4]) {
    const double t2 = t * t, t3 = t2 * t, u = 1.0 - t;
    w[0] = u * u * u * (1.0 / 6.0);
    w[1] = (3.0 * t3 - 6.0 * t2 + 4.0) * (1.0 / 6.0);
    w[2] = (-3.0 * t3 + 3.0 * t2 + 3.0 * t + 1.0) * (1.0 / 6.0);
    w[3] = t3 * (1.0 / 6.0);
}

template <int MODE>
__device__ __forceinline__ int tap_index(int i, int n) {
    if (MODE == kNearest) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    if (n == 1) return 0;
    const int s2 = 2 * n - 2;
    i = (i < 0 ? -i : i) % s2;
    return i >= n ? s2 - i : i;
}

// coefficient arrays: (Np, Mp) of T (double or double2), evaluated at (cx, cy) in array coordinates
template <int MODE, typename T>
__device__ __forceinline__ T spline_eval(const T* __restrict__ coef, int Np, int Mp, double cx, double cy) {
    // beyond one sample outside every tap is clamped anyway; this also keeps the int conversion safe
    if (MODE == kNearest) {
        cx = fmin(fmax(cx, -4.0), (double)Np + 3.0);
        cy = fmin(fmax(cy, -4.0), (double)Mp + 3.0);
    }
    const double fx = floor(cx), fy = floor(cy);
    double wx[4], wy[4];
    bspline_weights(cx - fx, wx);
    bspline_weights(cy - fy, wy);
    const int ix = (int)fx - 1, iy = (int)fy - 1;
    int ty[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ty[j] = tap_index<MODE>(iy + j, Mp);
    T acc;
    if constexpr (sizeof(T) == sizeof(double2)) acc = make_double2(0.0, 0.0);
    else acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const T* __restrict__ row = coef + (size_t)tap_index<MODE>(ix + i, Np) * Mp;
        T r;
        if constexpr (sizeof(T) == sizeof(double2)) {
            r = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 v = __ldg(row + ty[j]);
                r.x = fma(wy[j], v.x, r.x);
                r.y = fma(wy[j], v.y, r.y);
            }
            acc.x = fma(wx[i], r.x, acc.x);
            acc.y = fma(wx[i], r.y, acc.y);
        } else {
            r = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) r = fma(wy[j], __ldg(row + ty[j]), r);
            acc = fma(wx[i], r, acc);
        }
    }
    return acc;
}

// whole fixed-point inversion of one output pixel in registers:
//   u_it <- u(r - e0);  repeat iters times: u_it <- u(r - e1 + u_it),  r on an (on, om) grid
// invert_u_overlap (geometric_phase_analysis.py:291-299): e0 = e1 = edge on the grown grid (N + 2 edge, M + 2 edge);
// invert_u (:255-258): e0 = 0, e1 = edge on the (N, M) grid (the reference's `- edge` only enters the iterations).
__global__ void __launch_bounds__(256) k_invert_u(const double2* __restrict__ coef, int Np, int Mp, int on, int om,
                                                  int e0, int e1, int iters, double* __restrict__ out) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (r >= on || c >= om) return;
    const double x = (double)(r - e1) + kPad, y = (double)(c - e1) + kPad;   // padded-array coordinates
    double2 u = spline_eval<kNearest>(coef, Np, Mp, (double)(r - e0) + kPad, (double)(c - e0) + kPad);
    for (int it = 0; it < iters; ++it) u = spline_eval<kNearest>(coef, Np, Mp, x + u.x, y + u.y);
    out[(size_t)r * om + c] = u.x;
    out[(size_t)on * om + (size_t)r * om + c] = u.y;
}

// out(r, c) = spline(img)(r + u0(r,c), c + u1(r,c)), 0 outside [0, N-1] x [0, M-1]     (:973)
__global__ void __launch_bounds__(256) k_resample(const double* __restrict__ coef, int N, int M,
                                                  const double* __restrict__ u, double* __restrict__ out) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (r >= N || c >= M) return;
    const size_t i = (size_t)r * M + c;
    const double cx = (double)r + u[i], cy = (double)c + u[(size_t)N * M + i];
    double v = 0.0;
    // comparisons written so that NaN coordinates fall outside
    if (cx >= 0.0 && cx <= (double)(N - 1) && cy >= 0.0 && cy <= (double)(M - 1))
        v = spline_eval<kConstant>(coef, N, M, cx, cy);
    out[i] = v;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Prefilter `src` (N, M) [times scale] into dst with element stride dst_elem / offset dst_off;
// dst extent is (N + 2 pad, M + 2 pad).  tmp: two scratch arrays of that extent.
static int prefilter_2d(const double* src, int N, int M, double scale, int boundary, double* dst, int dst_elem,
                        int dst_off, double* tmp0, double* tmp1, cudaStream_t st) {
    const int pad = boundary == kNearest ? kPad : 0;
    const int Np = N + 2 * pad, Mp = M + 2 * pad;
    PrefilterArgs a;
    a.src = src; a.dst = tmp0; a.n = Np; a.cols = Mp; a.src_n = N; a.src_cols = M; a.pad = pad; a.scale = scale;
    a.boundary = boundary; a.gain = 6.0;
    dim3 gA(ceil_div(Mp, 128), ceil_div(Np, kSegLen));
    k_prefilter_causal<<<gA, 128, 0, st>>>(a);                                      // along axis 0 (+ padding): tmp0 = c+
    k_prefilter_anticausal<<<gA, 128, 0, st>>>(tmp0, tmp1, Np, Mp, boundary);       // tmp1 = c
    dim3 g1(ceil_div(Mp, 32), ceil_div(Np, 32));
    k_transpose_plain<<<g1, dim3(32, 8), 0, st>>>(tmp1, tmp0, Np, Mp, 1, 0);        // tmp0: (Mp, Np)
    a.src = tmp0; a.dst = tmp1; a.n = Mp; a.cols = Np; a.src_n = Mp; a.src_cols = Np; a.pad = 0; a.scale = 1.0;
    dim3 gB(ceil_div(Np, 128), ceil_div(Mp, kSegLen));
    k_prefilter_causal<<<gB, 128, 0, st>>>(a);                                      // along axis 1: tmp1 = c+
    k_prefilter_anticausal<<<gB, 128, 0, st>>>(tmp1, tmp0, Mp, Np, boundary);       // tmp0 = c
    dim3 g2(ceil_div(Np, 32), ceil_div(Mp, 32));
    k_transpose_plain<<<g2, dim3(32, 8), 0, st>>>(tmp0, dst, Mp, Np, dst_elem, dst_off);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_lawler_workspace_bytes(int N, int M, int edge, size_t* bytes) {
    GPA_REQUIRE(bytes && N >= 1 && M >= 1 && edge >= 0, "bad argument");
    const size_t np = (size_t)(N + 2 * kPad) * (M + 2 * kPad);
    const size_t out = (size_t)(N + 2 * edge) * (M + 2 * edge);
    // interleaved coefficients (2 np) + 2 scratch (np each) + image coefficients (N M) + u_inv (2 out)
    *bytes = (4 * np + (size_t)N * M + 2 * out) * sizeof(double) + 8 * 256;
    return GPA_OK;
}

static int invert_u_impl(const double* u, int N, int M, double scale, int iters, int edge, bool overlap, double* out,
                         void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(u && out && ws, "null pointer argument");
    GPA_REQUIRE(N >= 1 && M >= 1 && iters >= 0 && edge >= 0, "bad argument");
    const int Np = N + 2 * kPad, Mp = M + 2 * kPad;
    const size_t np = (size_t)Np * Mp;
    Arena a(ws, ws_bytes);
    double* coef = a.take<double>(2 * np);
    double* t0 = a.take<double>(np);
    double* t1 = a.take<double>(np);
    if (!a.ok()) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, a.off);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    {
        KernelTimer t("lf_prefilter", st);
        for (int comp = 0; comp < 2; ++comp)
            if ((rc = prefilter_2d(u + (size_t)comp * N * M, N, M, scale, kNearest, coef, 2, comp, t0, t1, st))) return rc;
    }
    {
        KernelTimer t("k_invert_u", st);
        const int on = overlap ? N + 2 * edge : N, om = overlap ? M + 2 * edge : M;
        dim3 grid(ceil_div(om, 32), ceil_div(on, 8));
        k_invert_u<<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(coef), Np, Mp, on, om, overlap ? edge : 0, edge, iters, out);
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_invert_u(const double* u, int N, int M, double scale, int iters, int edge, double* out,
                            void* ws, size_t ws_bytes, void* stream) {
    return invert_u_impl(u, N, M, scale, iters, edge, true, out, ws, ws_bytes, stream);
}

extern "C" int gpa_invert_u_plain(const double* u, int N, int M, double scale, int iters, int edge, double* out,
                                  void* ws, size_t ws_bytes, void* stream) {
    return invert_u_impl(u, N, M, scale, iters, edge, false, out, ws, ws_bytes, stream);
}

extern "C" int gpa_resample_image(const double* img, int N, int M, const double* u_inv, double* out,
                                  void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(img && u_inv && out && ws, "null pointer argument");
    GPA_REQUIRE(N >= 1 && M >= 1, "bad shape");
    const size_t nm = (size_t)N * M;
    Arena a(ws, ws_bytes);
    double* coef = a.take<double>(nm);
    double* t0 = a.take<double>(nm);
    double* t1 = a.take<double>(nm);
    if (!a.ok()) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, a.off);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    {
        KernelTimer t("lf_prefilter", st);
        if ((rc = prefilter_2d(img, N, M, 1.0, kConstant, coef, 1, 0, t0, t1, st))) return rc;
    }
    {
        KernelTimer t("k_resample", st);
        dim3 grid(ceil_div(M, 32), ceil_div(N, 8));
        k_resample<<<grid, 256, 0, st>>>(coef, N, M, u_inv, out);
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_undistort_image(const double* img, const double* u, int N, int M, int iters, double* out,
                                   void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(ws != nullptr && N >= 1 && M >= 1, "bad argument");
    Arena a(ws, ws_bytes);
    double* u_inv = a.take<double>(2 * (size_t)N * M);
    const size_t used = align_up(a.off, 256);
    if (used >= ws_bytes) {
        set_error("workspace too small (%zu)", ws_bytes);
        return GPA_ERR_WORKSPACE;
    }
    char* rest = static_cast<char*>(ws) + used;
    int rc = gpa_invert_u(u, N, M, -1.0, iters, 0, u_inv, rest, ws_bytes - used, stream);    // invert_u_overlap(-u), :971
    if (rc) return rc;
    return gpa_resample_image(img, N, M, u_inv, out, rest, ws_bytes - used, stream);
}

// ==============================================================================================
// K7 — unit-cell averaging: pyGPA/unit_cell_averaging.py:132-249
// ==============================================================================================
// unit_cell_average: every pixel r (+ u(r)) is folded into the unit cell spanned by the two
// k-vectors ks — fractional coordinates f = (r + u) ks^T mod 1, cartesian R = z (f ks^-T - rmin) — and
// dropped "drizzle like" onto the 2 x 2 cluster of cells around R (:208-217).  A scatter-add: fp64
// atomics into the (small, L2-resident) cell array; the order of the additions is not fixed, so
// results agree with the reference to rounding (1e-13 relative), not bit for bit.
// expand_unitcell (:234-249) is the gather back: the K4 cubic-spline evaluation (mode='constant')
// of the NaN-cleared cell at the folded coordinate of every output pixel.
namespace gpa {

struct UcGeom {
    double ks[2][2], kinv[2][2], rmin[2], z;
};

__device__ __forceinline__ void fold_into_cell(const UcGeom& g, double x, double y, double& R0, double& R1) {
    double f0 = x * g.ks[0][0] + y * g.ks[0][1];         // forward_transform: vecs @ ks.T
    double f1 = x * g.ks[1][0] + y * g.ks[1][1];
    f0 -= floor(f0);                                     // % 1.
    f1 -= floor(f1);
    R0 = (f0 * g.kinv[0][0] + f1 * g.kinv[0][1] - g.rmin[0]) * g.z;   // backward_transform, - rmin, * z
    R1 = (f0 * g.kinv[1][0] + f1 * g.kinv[1][1] - g.rmin[1]) * g.z;
}

__global__ void __launch_bounds__(256) k_uc_scatter(const double* __restrict__ img, const double* __restrict__ u, int N, int M,
                                                    const UcGeom g, int rs0, int rs1, double* __restrict__ res,
                                                    double* __restrict__ weights) {
    const size_t total = (size_t)N * M;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const double v = img[i];
        if (isnan(v)) continue;                          // NaN masks a pixel (:194)
        const int r = (int)(i / M), c = (int)(i % M);
        double x = (double)r, y = (double)c;
        if (u) {
            x += u[i];
            y += u[total + i];
        }
        double R0, R1;
        fold_into_cell(g, x, y, R0, R1);
        const double fl0 = floor(R0), fl1 = floor(R1);
        const double f0 = R0 - fl0, f1 = R1 - fl1;
        const int i0 = (int)fl0, i1 = (int)fl1;
        // float_overlap (:37-42) as written: overlap[li][lj] = (lj ? f0 : 1 - f0) * (li ? f1 : 1 - f1)
#pragma unroll
        for (int li = 0; li < 2; ++li)
#pragma unroll
            for (int lj = 0; lj < 2; ++lj) {
                const int a = i0 + li, b = i1 + lj;
                if (a < 0 || a >= rs0 || b < 0 || b >= rs1) continue;    // the reference would write out of bounds
                const double w = (lj ? f0 : 1.0 - f0) * (li ? f1 : 1.0 - f1);
                atomicAdd(res + (size_t)a * rs1 + b, v * w);
                atomicAdd(weights + (size_t)a * rs1 + b, w);
            }
    }
}

// res / weights (NaN where nothing landed), or with clear_nan the NaN-cleared copy expand_unitcell filters
__global__ void k_uc_divide(const double* __restrict__ res, const double* __restrict__ weights, size_t n, double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = res[i] / weights[i];
}

__global__ void k_nan_to_num(const double* __restrict__ in, size_t n, double* __restrict__ out) {
    const double big = 1.7976931348623157e308;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = in[i];
        out[i] = isnan(v) ? 0.0 : (isinf(v) ? copysign(big, v) : v);
    }
}

__global__ void __launch_bounds__(256) k_uc_expand(const double* __restrict__ coef, int n, int m, int H, int W,
                                                   const double* __restrict__ u, double u_const, double z2, const UcGeom g,
                                                   double* __restrict__ out) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int r = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (r >= H || c >= W) return;
    const size_t i = (size_t)r * W + c;
    double x = (double)r / z2, y = (double)c / z2;              // np.mgrid / z2
    if (u) {
        x += u[i];
        y += u[(size_t)H * W + i];
    } else {
        x += u_const;
        y += u_const;
    }
    double R0, R1;
    fold_into_cell(g, x, y, R0, R1);
    double v = 0.0;
    if (R0 >= 0.0 && R0 <= (double)(n - 1) && R1 >= 0.0 && R1 <= (double)(m - 1)) v = spline_eval<kConstant>(coef, n, m, R0, R1);
    out[i] = v;
}

static void fill_geom(UcGeom& g, const double* ks, const double* kinv, const double* rmin, double z) {
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            g.ks[i][j] = ks[2 * i + j];
            g.kinv[i][j] = kinv[2 * i + j];
        }
    g.rmin[0] = rmin[0]; g.rmin[1] = rmin[1]; g.z = z;
}

}  // namespace gpa

extern "C" int gpa_uc_workspace_bytes(int rs0, int rs1, size_t* bytes) {
    GPA_REQUIRE(bytes && rs0 >= 1 && rs1 >= 1, "bad argument");
    *bytes = 4 * ((size_t)rs0 * rs1 * sizeof(double) + 256) + 1024;
    return GPA_OK;
}

extern "C" int gpa_uc_average(const double* img, const double* u /*(2,N,M) or null*/, int N, int M,
                              const double* ks /*host 4*/, const double* kinv /*host 4*/, const double* rmin /*host 2*/,
                              double z, int rs0, int rs1, double* out /*(rs0, rs1)*/, void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(img && ks && kinv && rmin && out && ws, "null pointer argument");
    GPA_REQUIRE(N >= 1 && M >= 1 && rs0 >= 1 && rs1 >= 1, "bad shape");
    size_t need = 0;
    gpa_uc_workspace_bytes(rs0, rs1, &need);
    if (ws_bytes < need) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)rs0 * rs1;
    Arena a(ws, ws_bytes);
    double* res = a.take<double>(n);
    double* wts = a.take<double>(n);
    GPA_CHECK_CUDA(cudaMemsetAsync(res, 0, n * sizeof(double), st));
    GPA_CHECK_CUDA(cudaMemsetAsync(wts, 0, n * sizeof(double), st));
    UcGeom g;
    fill_geom(g, ks, kinv, rmin, z);
    size_t blocks = ((size_t)N * M + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    {
        KernelTimer t("k_uc_scatter", st);
        k_uc_scatter<<<(unsigned)blocks, 256, 0, st>>>(img, u, N, M, g, rs0, rs1, res, wts);
    }
    k_uc_divide<<<(unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(res, wts, n, out);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_uc_expand(const double* ucell /*(n, m)*/, int n, int m, int H, int W, const double* u /*(2,H,W) or null*/,
                             double u_const, double z2, const double* ks, const double* kinv, const double* rmin, double z,
                             double* out /*(H, W)*/, void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(ucell && ks && kinv && rmin && out && ws, "null pointer argument");
    GPA_REQUIRE(n >= 2 && m >= 2 && H >= 1 && W >= 1 && z2 != 0.0, "bad argument");
    size_t need = 0;
    gpa_uc_workspace_bytes(n, m, &need);
    if (ws_bytes < need) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t cnt = (size_t)n * m;
    Arena a(ws, ws_bytes);
    double* clean = a.take<double>(cnt);
    double* coef = a.take<double>(cnt);
    double* t0 = a.take<double>(cnt);
    double* t1 = a.take<double>(cnt);
    k_nan_to_num<<<(unsigned)((cnt + 255) / 256 < 1184 ? (cnt + 255) / 256 : 1184), 256, 0, st>>>(ucell, cnt, clean);   // np.nan_to_num (:246)
    int rc = prefilter_2d(clean, n, m, 1.0, kConstant, coef, 1, 0, t0, t1, st);
    if (rc) return rc;
    UcGeom g;
    fill_geom(g, ks, kinv, rmin, z);
    KernelTimer t("k_uc_expand", st);
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8));
    k_uc_expand<<<grid, 256, 0, st>>>(coef, n, m, H, W, u, u_const, z2, g, out);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}
