// Small device utilities of the C ABI (dtype staging for the host binding).
#include "common.cuh"

namespace gpa {
__global__ void k_cast_f64_f32(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}
}  // namespace gpa

extern "C" int gpa_cast_f64_to_f32(const double* in, float* out, size_t n, void* stream) {
    GPA_REQUIRE(in && out, "null pointer argument");
    if (n == 0) return GPA_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gpa::k_cast_f64_f32<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

namespace gpa {
// phases = angle(lockin), weights = |lockin| * (mask + eps), mask = 1 on [dr, N-dr) x [dr, M-dr)
// (extract_displacement_field, geometric_phase_analysis.py:922-926)
template <typename T2>
__global__ void k_phase_weight(const T2* __restrict__ lockin, double* __restrict__ phases, double* __restrict__ weights,
                               int N, int M, int dr, double eps) {
    const size_t n = (size_t)N * M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / M), c = (int)(i % M);
        const T2 v = lockin[i];
        const double re = v.x, im = v.y;
        // numpy's mask[dr:-dr, dr:-dr] is empty for dr == 0
        const bool inside = dr > 0 && r >= dr && r < N - dr && c >= dr && c < M - dr;
        phases[i] = atan2(im, re);
        weights[i] = hypot(re, im) * ((inside ? 1.0 : 0.0) + eps);
    }
}
}  // namespace gpa

extern "C" int gpa_phase_weight(const void* lockin, int is_f64, int N, int M, int border, double eps,
                                double* phases, double* weights, void* stream) {
    GPA_REQUIRE(lockin && phases && weights && N > 0 && M > 0 && border >= 0, "bad argument");
    size_t blocks = ((size_t)N * M + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (is_f64)
        gpa::k_phase_weight<double2><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const double2*>(lockin), phases, weights, N, M, border, eps);
    else
        gpa::k_phase_weight<float2><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float2*>(lockin), phases, weights, N, M, border, eps);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

namespace gpa {
// kidx = flat candidate index packed in the key, -1 where no candidate won (high word 0)
__global__ void k_key_to_kidx(const unsigned long long* __restrict__ key, int* __restrict__ kidx, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long k = key[i];
        kidx[i] = (k >> 32) == 0ull ? -1 : (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
    }
}
}  // namespace gpa

extern "C" int gpa_key_to_kidx(const unsigned long long* key, int* kidx, size_t n, void* stream) {
    GPA_REQUIRE(key && kidx, "null pointer argument");
    if (n == 0) return GPA_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gpa::k_key_to_kidx<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(key, kidx, n);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

// ---------------------------------------------------------------------------------------------
// FP32 pipe peak, measured: the denominator of the K1 roofline (MEASURED_PEAKS.json has no fp32 figure).
// Pure register FFMA2 (fma.rn.f32x2, the instruction the sweep kernels issue), 16 independent accumulators
// per thread, 4 CTAs of 256 threads per SM, no memory traffic in the loop.
// ---------------------------------------------------------------------------------------------
namespace gpa {
__global__ void __launch_bounds__(256) k_ffma_peak(float2* __restrict__ out, const float2* __restrict__ in, int iters) {
    float2 acc[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        acc[i] = in[threadIdx.x + i * 256];
        b[i] = in[threadIdx.x + (i + 16) * 256];
    }
    for (int it = 0; it < iters; ++it) {
        const float2 g = make_float2(b[0].x + (float)it, b[0].y + (float)it);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(g, b[(i + r) & 15], acc[i]);
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        s.x += acc[i].x;
        s.y += acc[i].y;
    }
    out[(size_t)blockIdx.x * 256 + threadIdx.x] = s;
}
}  // namespace gpa

extern "C" int gpa_fp32_peak_tflops(void* ws, size_t ws_bytes, double* tflops, void* stream) {
    GPA_REQUIRE(ws && tflops, "null pointer argument");
    int dev = 0, sms = 0;
    GPA_CHECK_CUDA(cudaGetDevice(&dev));
    GPA_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 4, iters = 8192;
    const size_t need = (size_t)(32 * 256 + grid * 256) * sizeof(float2);
    if (ws_bytes < need) {
        gpa::set_error("workspace too small (%zu bytes, need %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float2* in = static_cast<float2*>(ws);
    float2* out = in + 32 * 256;
    GPA_CHECK_CUDA(cudaMemsetAsync(in, 0, 32 * 256 * sizeof(float2), st));
    cudaEvent_t e0, e1;
    GPA_CHECK_CUDA(cudaEventCreate(&e0));
    GPA_CHECK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {       // the first repetition warms up
        cudaEventRecord(e0, st);
        gpa::k_ffma_peak<<<grid, 256, 0, st>>>(out, in, iters);
        cudaEventRecord(e1, st);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            gpa::set_error("gpa_fp32_peak_tflops: %s", cudaGetErrorString(e));
            return GPA_ERR_CUDA;
        }
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // per thread and iteration: 4 x 16 FFMA2 = 64 x 2 lanes x 2 flop
    *tflops = (double)grid * 256.0 * iters * 64.0 * 4.0 / ((double)best * 1e-3) / 1e12;
    return GPA_OK;
}
