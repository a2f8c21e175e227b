// K6 — device pieces of the k-vector refinement loop iterate_GPA (float64), sm_100a.
//
// Reference semantics: iterate_GPA (pyGPA/geometric_phase_analysis.py:116-154): fixed-reference
// lock-in per k -> angle / abs cropped by `edge` (:134-139) -> weighted unwrap with
// sqrt(w / max w) (:141-143) -> robust plane fit of the unwrapped phase (fit_delta_k :92-94 ->
// mathtools.fit_plane :30-47: scipy.optimize.least_squares(loss='huber'), f_scale = 1) -> the slope
// corrects the k-vector.  The lock-in (K1) and the unwrap (K2) exist; this file adds
//   k_phase_amp_crop   angle, abs on the cropped window + running max of abs
//   k_sqrt_norm        sqrt(w / max w)
//   k_plane_irls       one iteratively-reweighted-least-squares step of the Huber plane fit: the
//                      minimiser of sum rho(r_i^2), rho = Huber, satisfies the weighted normal
//                      equations with w_i = min(1, 1/|r_i|); every step is one streaming reduction
//                      of nine sums, the CTA that finishes last solves the 3x3 system on the device.
// All HBM-bound streaming kernels.
#include "common.cuh"

namespace gpa {

__device__ __forceinline__ double block_sum32(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = lane < nw ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;       // every thread of every warp holds the total
}

template <typename T2>
__global__ void __launch_bounds__(256) k_phase_amp_crop(const T2* __restrict__ lockin, int N, int M, int edge,
                                                        double* __restrict__ phases, double* __restrict__ amp,
                                                        unsigned long long* __restrict__ amp_max_bits) {
    __shared__ double sh[32];
    const int n = N - 2 * edge, m = M - 2 * edge;
    const size_t total = (size_t)n * m;
    double mx = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int r = (int)(i / m), c = (int)(i % m);
        const T2 v = lockin[(size_t)(r + edge) * M + c + edge];
        const double re = v.x, im = v.y;
        const double a = hypot(re, im);
        phases[i] = atan2(im, re);
        amp[i] = a;
        mx = fmax(mx, a);
    }
    // CTA max through the sum helper's layout: reuse shuffles directly
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmax(mx, sh[w]);
        // non-negative doubles order like their bit patterns
        atomicMax(amp_max_bits, (unsigned long long)__double_as_longlong(mx));
    }
}

__global__ void __launch_bounds__(256) k_sqrt_norm(const double* __restrict__ amp, const double* __restrict__ amp_max,
                                                   size_t n, double* __restrict__ out) {
    const double mx = *amp_max;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        out[i] = sqrt(amp[i] / mx);          // np.sqrt(we / we.max()), geometric_phase_analysis.py:141
}

struct PlaneFitState {
    double theta[3];      // point the next launch evaluates, in scaled, centred coordinates: z ~ t0 xs + t1 ys + t2
    double accepted[3];   // last accepted point and its objective
    double f_accepted;
    double irls[3];       // IRLS step from the last accepted point (the safe fallback)
    double delta;         // largest change of the plane over the frame in the last step
    int done, iter, max_iter, was_newton;
    unsigned ticket;
};

constexpr int kPlaneSums = 19;   // IRLS normal equations 9, inlier Hessian 6, gradient 3, objective 1

// 3x3 solve by Gaussian elimination with partial pivoting; false when singular
__device__ bool solve3(double A[3][4], double (&x)[3]) {
    for (int col = 0; col < 3; ++col) {
        int piv = col;
        for (int r = col + 1; r < 3; ++r)
            if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
        if (!(fabs(A[piv][col]) > 0.0)) return false;
        for (int k = 0; k < 4; ++k) { const double t = A[col][k]; A[col][k] = A[piv][k]; A[piv][k] = t; }
        for (int r = col + 1; r < 3; ++r) {
            const double f = A[r][col] / A[col][col];
            for (int k = col; k < 4; ++k) A[r][k] -= f * A[col][k];
        }
    }
    x[2] = A[2][3] / A[2][2];
    x[1] = (A[1][3] - A[1][2] * x[2]) / A[1][1];
    x[0] = (A[0][3] - A[0][1] * x[1] - A[0][2] * x[2]) / A[0][0];
    return isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]);
}

// One step of the Huber plane fit.  The objective sum rho(r_i) is convex and piecewise quadratic, so
// Newton's method on the current inlier set (|r| <= f_scale) lands on the minimiser in a handful of
// steps once the inlier set settles; the IRLS step (weights min(1, f_scale/|r|), a majorise-minimise
// step that can never increase the objective) is computed in the same sweep and taken instead whenever
// the Newton step is unavailable (singular inlier Hessian) or made the objective worse.  Launch k
// evaluates everything at state->theta; the CTA that finishes last decides the next point.
__global__ void __launch_bounds__(256) k_plane_irls(const double* __restrict__ img, int n, int m, PlaneFitState* st,
                                                    double* __restrict__ partial, double f_scale, double tol) {
    if (st->done) return;
    __shared__ double sh[32];
    __shared__ int s_last;
    const bool first = st->iter == 0;
    const double t0 = st->theta[0], t1 = st->theta[1], t2 = st->theta[2];
    const double xc = 0.5 * (n - 1), yc = 0.5 * (m - 1), sx = 1.0 / n, sy = 1.0 / m;
    double acc[kPlaneSums];
#pragma unroll
    for (int k = 0; k < kPlaneSums; ++k) acc[k] = 0.0;
    const size_t total = (size_t)n * m;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int r = (int)(i / m), c = (int)(i % m);
        const double x = ((double)r - xc) * sx, y = ((double)c - yc) * sy, z = img[i];
        const double res = z - (t0 * x + t1 * y + t2), ares = fabs(res);
        const bool in = first || ares <= f_scale;           // the first step is ordinary least squares
        const double w = in ? 1.0 : f_scale / ares;          // Huber: psi(r) / r
        const double wx = w * x, wy = w * y;
        acc[0] += wx * x; acc[1] += wx * y; acc[2] += wy * y;
        acc[3] += wx;     acc[4] += wy;     acc[5] += w;
        acc[6] += wx * z; acc[7] += wy * z; acc[8] += w * z;
        if (in) {
            acc[9] += x * x; acc[10] += x * y; acc[11] += y * y;
            acc[12] += x;    acc[13] += y;     acc[14] += 1.0;
        }
        const double psi = w * res;                           // clip(res, -f_scale, f_scale)
        acc[15] += psi * x; acc[16] += psi * y; acc[17] += psi;
        acc[18] += in ? 0.5 * res * res : f_scale * ares - 0.5 * f_scale * f_scale;
    }
    const int nblk = gridDim.x;
#pragma unroll
    for (int k = 0; k < kPlaneSums; ++k) {
        const double s = block_sum32(acc[k], sh);
        if (threadIdx.x == 0) partial[(size_t)k * nblk + blockIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double S[kPlaneSums];
    for (int k = 0; k < kPlaneSums; ++k) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblk; i += 256) s += __ldcg(partial + (size_t)k * nblk + i);
        S[k] = block_sum32(s, sh);
    }
    if (threadIdx.x != 0) return;
    st->ticket = 0u;
    st->iter += 1;
    const double F = S[18];
    if (!first && st->was_newton && !(F <= st->f_accepted)) {
        // the Newton step overshot: go back and take the IRLS step from the last accepted point
        st->theta[0] = st->irls[0]; st->theta[1] = st->irls[1]; st->theta[2] = st->irls[2];
        st->was_newton = 0;
        if (st->iter >= st->max_iter) {
            st->theta[0] = st->accepted[0]; st->theta[1] = st->accepted[1]; st->theta[2] = st->accepted[2];
            st->done = 1;
        }
        return;
    }
    st->accepted[0] = t0; st->accepted[1] = t1; st->accepted[2] = t2;
    st->f_accepted = F;
    double next[3] = {t0, t1, t2};
    bool newton = false, have = false;
    {
        double A[3][4] = {{S[0], S[1], S[3], S[6]}, {S[1], S[2], S[4], S[7]}, {S[3], S[4], S[5], S[8]}};
        double th[3];
        if (solve3(A, th)) {
            st->irls[0] = th[0]; st->irls[1] = th[1]; st->irls[2] = th[2];
            next[0] = th[0]; next[1] = th[1]; next[2] = th[2];
            have = true;
        }
    }
    if (!first && S[14] >= 3.0) {
        double H[3][4] = {{S[9], S[10], S[12], S[15]}, {S[10], S[11], S[13], S[16]}, {S[12], S[13], S[14], S[17]}};
        double d[3];
        if (solve3(H, d)) {
            next[0] = t0 + d[0]; next[1] = t1 + d[1]; next[2] = t2 + d[2];
            newton = have = true;
        }
    }
    if (!have) {                 // degenerate frame (all weights zero): keep the current plane
        st->done = 1;
        return;
    }
    // change of the fitted plane anywhere on the frame (scaled coordinates span [-1/2, 1/2])
    const double d = 0.5 * fabs(next[0] - t0) + 0.5 * fabs(next[1] - t1) + fabs(next[2] - t2);
    st->delta = d;
    st->was_newton = newton ? 1 : 0;
    st->theta[0] = next[0]; st->theta[1] = next[1]; st->theta[2] = next[2];
    if ((!first && d < tol) || st->iter >= st->max_iter) st->done = 1;
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_lockin_phase_amp(const void* lockin, int is_f64, int N, int M, int edge, double* phases,
                                    double* amp, double* amp_max /*device, 1*/, void* stream) {
    GPA_REQUIRE(lockin && phases && amp && amp_max, "null pointer argument");
    GPA_REQUIRE(edge >= 0 && N - 2 * edge >= 1 && M - 2 * edge >= 1, "edge %d leaves nothing of a %d x %d frame", edge, N, M);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GPA_CHECK_CUDA(cudaMemsetAsync(amp_max, 0, sizeof(double), st));
    const size_t total = (size_t)(N - 2 * edge) * (M - 2 * edge);
    size_t blocks = (total + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    KernelTimer t("k_phase_amp_crop", st);
    if (is_f64)
        k_phase_amp_crop<double2><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const double2*>(lockin), N, M, edge, phases, amp,
                                                                    reinterpret_cast<unsigned long long*>(amp_max));
    else
        k_phase_amp_crop<float2><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float2*>(lockin), N, M, edge, phases, amp,
                                                                   reinterpret_cast<unsigned long long*>(amp_max));
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_weight_sqrt_norm(const double* amp, const double* amp_max /*device*/, size_t n, double* out, void* stream) {
    GPA_REQUIRE(amp && amp_max && out, "null pointer argument");
    if (n == 0) return GPA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    size_t blocks = (n + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    KernelTimer t("k_sqrt_norm", st);
    k_sqrt_norm<<<(unsigned)blocks, 256, 0, st>>>(amp, amp_max, n, out);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_fit_plane_workspace_bytes(size_t* bytes) {
    GPA_REQUIRE(bytes != nullptr, "bytes is null");
    *bytes = (size_t)kPlaneSums * 148 * 8 * sizeof(double) + sizeof(PlaneFitState) + 1024;
    return GPA_OK;
}

// Huber plane fit z ~ a0 x + a1 y + a2 (x, y = array indices), the minimiser mathtools.fit_plane
// (mathtools.py:30-47) approaches with scipy's trust-region solver.  theta: HOST, 3 doubles.
extern "C" int gpa_fit_plane_huber(const double* img, int n, int m, double f_scale, int max_iter, double tol,
                                   double* theta /*host*/, int* iterations /*host, may be null*/, void* ws,
                                   size_t ws_bytes, void* stream) {
    GPA_REQUIRE(img && theta && ws, "null pointer argument");
    GPA_REQUIRE(n >= 2 && m >= 2 && f_scale > 0.0 && max_iter >= 1, "bad argument");
    size_t need = 0;
    gpa_fit_plane_workspace_bytes(&need);
    if (ws_bytes < need) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena a(ws, ws_bytes);
    double* partial = a.take<double>((size_t)kPlaneSums * 148 * 8);
    PlaneFitState* state = a.take<PlaneFitState>(1);
    PlaneFitState h;
    std::memset(&h, 0, sizeof(h));
    h.max_iter = max_iter;
    GPA_CHECK_CUDA(cudaMemcpyAsync(state, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    const size_t total = (size_t)n * m;
    size_t blocks = (total + 2047) / 2048;
    if (blocks > 148 * 8) blocks = 148 * 8;
    {
        KernelTimer t("k_plane_irls", st);
        for (int it = 0; it < max_iter; ++it) {
            k_plane_irls<<<(unsigned)blocks, 256, 0, st>>>(img, n, m, state, partial, f_scale, tol);
            if ((it & 7) == 7) {          // kernels return at once after convergence; stop enqueueing them too
                GPA_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
                GPA_CHECK_CUDA(cudaStreamSynchronize(st));
                if (h.done) break;
            }
        }
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    GPA_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
    GPA_CHECK_CUDA(cudaStreamSynchronize(st));
    // back to index coordinates: xs = (x - xc) / n, ys = (y - yc) / m
    const double xc = 0.5 * (n - 1), yc = 0.5 * (m - 1);
    theta[0] = h.theta[0] / n;
    theta[1] = h.theta[1] / m;
    theta[2] = h.theta[2] - theta[0] * xc - theta[1] * yc;
    if (iterations) *iterations = h.iter;
    return GPA_OK;
}
