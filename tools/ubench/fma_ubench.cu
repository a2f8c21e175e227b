// Microbenchmark: FP32 FMA issue-rate variants on sm_100a (throw-away, informs K1 design).
#include <cstdio>
#include <cuda_runtime.h>
#define NACC 16
#define ITERS 4096

__constant__ float2 c_taps[256];

template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, const float2* in, int iters) {
    float2 acc[NACC];
    float2 b[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i] = in[threadIdx.x + i * 256]; b[i] = in[threadIdx.x + (i + NACC) * 256]; }
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = in[i];
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // scalar FFMA, 3 regs, per-thread g
            float g = b[0].x + it;
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < NACC; i++) { acc[i].x = fmaf(g, b[(i + r) & 15].x, acc[i].x); acc[i].y = fmaf(g, b[(i + r) & 15].y, acc[i].y); }
        } else if (MODE == 1) {     // FFMA2, regs
            float2 g = make_float2(b[0].x + it, b[0].y + it);
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = __ffma2_rn(g, b[(i + r) & 15], acc[i]);
        } else if (MODE == 2) {     // scalar FFMA with constant-bank tap
#pragma unroll
            for (int r = 0; r < 4; r++) {
            float g = c_taps[(it * 4 + r) & 255].x;
#pragma unroll
            for (int i = 0; i < NACC; i++) { acc[i].x = fmaf(g, b[(i + r) & 15].x, acc[i].x); acc[i].y = fmaf(g, b[(i + r) & 15].y, acc[i].y); }
            }
        } else if (MODE == 3) {     // FFMA2 with constant-bank tap pair
#pragma unroll
            for (int r = 0; r < 4; r++) {
            float2 g = c_taps[(it * 4 + r) & 255];
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = __ffma2_rn(g, b[(i + r) & 15], acc[i]);
            }
        } else if (MODE == 4) {     // FFMA2 + 1 LDS.64 per 16 (sliding window realistic)
#pragma unroll
            for (int r = 0; r < 4; r++) {
            float2 g = c_taps[(it * 4 + r) & 255];
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = __ffma2_rn(g, b[(i + r) & 15], acc[i]);
            b[r] = sm[(threadIdx.x + it * 4 + r) & 2047];
            }
        } else if (MODE == 5) {     // scalar FFMA + LDS, const taps
#pragma unroll
            for (int r = 0; r < 4; r++) {
            float g = c_taps[(it * 4 + r) & 255].x;
#pragma unroll
            for (int i = 0; i < NACC; i++) { acc[i].x = fmaf(g, b[(i + r) & 15].x, acc[i].x); acc[i].y = fmaf(g, b[(i + r) & 15].y, acc[i].y); }
            b[r] = sm[(threadIdx.x + it * 4 + r) & 2047];
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < NACC; i++) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, float2* out, float2* in) {
    int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, in, ITERS);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(out, in, ITERS);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fma = (double)grid * 256 * ITERS * 4 * NACC * 2;  // scalar fma count
    printf("%-44s ctas/sm=%d  %.3f ms  %.1f TFLOP/s (fp32 fma*2)  err=%s\n", name, ctas_per_sm, best, fma * 2 / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float2 *in, *out; cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 148 * 8 * 256 * 8);
    cudaMemset(in, 0, 1 << 20);
    float2 h[256]; for (int i = 0; i < 256; i++) h[i] = make_float2(1e-3f * i, 1e-3f * i);
    cudaMemcpyToSymbol(c_taps, h, sizeof(h));
    for (int c = 1; c <= 2; c++) {
        run<0>("scalar FFMA 3-reg", c, out, in);
        run<1>("FFMA2 3-reg", c, out, in);
        run<2>("scalar FFMA const tap", c, out, in);
        run<3>("FFMA2 const tap", c, out, in);
        run<4>("FFMA2 const tap + LDS.64/16", c, out, in);
        run<5>("scalar FFMA const tap + LDS.64/32", c, out, in);
    }
    return 0;
}
