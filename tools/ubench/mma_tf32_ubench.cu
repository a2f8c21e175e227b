// Throughput of the legacy warp-level tensor path on sm_100a: mma.sync.m16n8k8 tf32 (SASS HMMA.1688.F32.TF32) and
// m16n8k16 bf16, register operands only.  Evidence for the tensor-core decision of K1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32_ubench mma_tf32_ubench.cu && ./mma_tf32_ubench
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) k_tf32(float* out, const float* in, int iters) {
    unsigned a[4], b[2];
    float c[NACC][4];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(in[threadIdx.x + 32 * i]);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(in[threadIdx.x + 32 * (i + 4)]);
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_bf16(float* out, const float* in, int iters) {
    unsigned a[4], b[2];
    float c[NACC][4];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(in[threadIdx.x + 32 * i]);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(in[threadIdx.x + 32 * (i + 4)]);
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <typename F>
static double run(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *in, *out;
    cudaMalloc(&in, 1 << 16);
    cudaMemset(in, 0, 1 << 16);
    cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
    const int iters = 4096;
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
        double ms = run([&] { k_tf32<8><<<sms * ctas, 256>>>(out, in, iters); });
        double macs = (double)sms * ctas * 8 /*warps*/ * iters * 8 /*NACC*/ * (16.0 * 8 * 8);
        printf("tf32 m16n8k8  %d CTA/SM: %.3f ms  %.1f dense TFLOP/s  %.1f MAC/clk/SM @1.965GHz\n", ctas, ms, 2 * macs / ms / 1e9,
               macs / (ms * 1e-3) / sms / 1.965e9);
        ms = run([&] { k_bf16<8><<<sms * ctas, 256>>>(out, in, iters); });
        macs = (double)sms * ctas * 8 * iters * 8 * (16.0 * 8 * 16);
        printf("bf16 m16n8k16 %d CTA/SM: %.3f ms  %.1f dense TFLOP/s  %.1f MAC/clk/SM @1.965GHz\n", ctas, ms, 2 * macs / ms / 1e9,
               macs / (ms * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
