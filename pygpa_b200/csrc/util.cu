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
