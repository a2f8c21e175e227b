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
