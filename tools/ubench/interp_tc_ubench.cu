// Tensor-core decision for K1 (BASELINE north_star: "tensor cores are used only if ncu shows a banded-Toeplitz GEMM form of
// the Gaussian beating the CUDA-core version").  The y-interpolation stage of k_mr_interp — per candidate, a 64 x 128 pixel
// tile out[x][y] = sum_w T[y % 4][w] * p3t[y / 4 + w][x] (11 real taps x complex sample), then |sf|^2 and the running
// arg-max — written two ways on identical inputs:
//
//   A  CUDA cores: the production scheme (lockin_mr.cuh): 2 regions x 16 outputs per thread, packed FFMA2
//   B  tensor cores, warp-level path (mma.sync.m16n8k8 tf32 = SASS HMMA.1688.F32.TF32), block-banded GEMM: every 16-column
//      block is [64 x 16] . [16 x 16] (the 11-tap band padded to K = 16 -> 69 % useful), 3xTF32 split (hi*hi + lo*hi + hi*lo)
//      to keep |sf|^2 at fp32 accuracy, because the arg-max must match the reference except at 1e-5 near-ties
//
// Both stream the same x-interpolated tiles from global memory (cp.async, double buffered, one barrier per candidate) and
// keep (best |sf|^2, candidate index) per pixel in registers.  Reports ms per launch, the largest relative |sf|^2 difference
// and the number of pixels whose winner differs.  Run it under ncu for the pipe utilisations (profiles/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o interp_tc_ubench interp_tc_ubench.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int S = 4, W = 12, TX = 64, TY = 128, CY = TY / S + W - 2;   // 42 coarse columns feed 128 fine ones
constexpr int PA = TX + 1;      // tile pitch of variant A (as in production)
constexpr int PB = TX + 4;      // tile pitch of variant B: fragment loads of a half-warp hit 16 distinct 8-byte slots
constexpr int CYB = CY + 2;    // variant B reads K = 16 coarse columns per block: two zero rows of padding (taps there are 0)
constexpr int NCAND = 96;

struct Taps {
    float2 g[S * W];            // (tap, tap) for FFMA2:  g[phase * W + w]
};

__device__ __forceinline__ void cp_async8(void* s, const void* g) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(g));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// A: CUDA cores
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_interp_cuda(const float2* __restrict__ tiles, int ncand, const __grid_constant__ Taps taps,
                                                       float* __restrict__ best_out, int* __restrict__ idx_out) {
    extern __shared__ float2 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto fetch = [&](int c) {
        const float2* src = tiles + (size_t)c * CY * PA;
        float2* dst = smem + (c & 1) * CY * PA;
        for (int i = threadIdx.x; i < CY * PA; i += 256) cp_async8(dst + i, src + i);
        cp_commit();
    };
    float best[2][16];
    int bidx[2][16];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int p = 0; p < 16; ++p) { best[h][p] = 0.f; bidx[h][p] = 0; }
    fetch(0);
    for (int c = 0; c < ncand; ++c) {
        cp_wait_all();
        __syncthreads();
        if (c + 1 < ncand) fetch(c + 1);
        const float2* t = smem + (c & 1) * CY * PA;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 smp[16 / S + W - 2], acc[16];
#pragma unroll
            for (int i = 0; i < 16 / S + W - 2; ++i) smp[i] = t[(warp * (16 / S) + i) * PA + lane + 32 * h];
#pragma unroll
            for (int p = 0; p < 16; ++p) acc[p] = make_float2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < 16; p += S) acc[p] = __ffma2_rn(taps.g[0], smp[p / S], acc[p]);
#pragma unroll
            for (int w = 1; w < W - 1; ++w)
#pragma unroll
                for (int p = 0; p < 16; ++p) acc[p] = __ffma2_rn(taps.g[(p % S) * W + w], smp[p / S + w], acc[p]);
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const float a2 = fmaf(acc[p].x, acc[p].x, acc[p].y * acc[p].y);
                if (a2 > best[h][p]) { best[h][p] = a2; bidx[h][p] = c; }
            }
        }
    }
    if (blockIdx.x == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int x = lane + 32 * h, y = warp * 16 + p;
                best_out[x * TY + y] = best[h][p];
                idx_out[x * TY + y] = bidx[h][p];
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// B: mma.sync tf32, 3xTF32
// ---------------------------------------------------------------------------------------------------------------------
struct BFrag {
    unsigned hi[2][2][2], lo[2][2][2];     // [k-step][n-tile][reg], per lane: filled on the host per lane id
};

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned to_tf32(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__global__ void __launch_bounds__(256, 2) k_interp_mma(const float2* __restrict__ tiles, int ncand, const BFrag* __restrict__ bfrag,
                                                      float* __restrict__ best_out, int* __restrict__ idx_out) {
    extern __shared__ float2 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    auto fetch = [&](int c) {
        const float2* src = tiles + (size_t)c * CYB * PB;
        float2* dst = smem + (c & 1) * CYB * PB;
        for (int i = threadIdx.x; i < CYB * PB; i += 256) cp_async8(dst + i, src + i);
        cp_commit();
    };
    const BFrag bf = bfrag[lane];          // the band matrix is the same for every 16-column block: registers, loaded once
    float best[4][2][4];
    int bidx[4][2][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) { best[m][n][i] = 0.f; bidx[m][n][i] = 0; }
    fetch(0);
    for (int c = 0; c < ncand; ++c) {
        cp_wait_all();
        __syncthreads();
        if (c + 1 < ncand) fetch(c + 1);
        const float2* tl = smem + (c & 1) * CYB * PB + (warp * (16 / S)) * PB;     // first coarse column of this warp's block
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            float cre[2][4], cim[2][4];
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) cre[n][i] = cim[n][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                // A[m][k] = tile[k][x = 16 m + row]: a0 (g, t), a1 (g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
                float2 v[4];
                v[0] = tl[(8 * ks + t) * PB + 16 * m + g];
                v[1] = tl[(8 * ks + t) * PB + 16 * m + g + 8];
                v[2] = tl[(8 * ks + t + 4) * PB + 16 * m + g];
                v[3] = tl[(8 * ks + t + 4) * PB + 16 * m + g + 8];
                unsigned rh[4], rl[4], ih[4], il[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    rh[i] = to_tf32(v[i].x);
                    rl[i] = to_tf32(v[i].x - __uint_as_float(rh[i]));
                    ih[i] = to_tf32(v[i].y);
                    il[i] = to_tf32(v[i].y - __uint_as_float(ih[i]));
                }
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    mma_tf32(cre[n], rl, bf.hi[ks][n][0], bf.hi[ks][n][1]);      // small terms first
                    mma_tf32(cre[n], rh, bf.lo[ks][n][0], bf.lo[ks][n][1]);
                    mma_tf32(cre[n], rh, bf.hi[ks][n][0], bf.hi[ks][n][1]);
                    mma_tf32(cim[n], il, bf.hi[ks][n][0], bf.hi[ks][n][1]);
                    mma_tf32(cim[n], ih, bf.lo[ks][n][0], bf.lo[ks][n][1]);
                    mma_tf32(cim[n], ih, bf.hi[ks][n][0], bf.hi[ks][n][1]);
                }
            }
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a2 = fmaf(cre[n][i], cre[n][i], cim[n][i] * cim[n][i]);
                    if (a2 > best[m][n][i]) { best[m][n][i] = a2; bidx[m][n][i] = c; }
                }
        }
    }
    if (blockIdx.x == 0) {
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int x = 16 * m + g + 8 * (i >> 1), y = warp * 16 + 8 * n + 2 * t + (i & 1);
                    best_out[x * TY + y] = best[m][n][i];
                    idx_out[x * TY + y] = bidx[m][n][i];
                }
    }
}

static float tf32_round(float v) {      // round to nearest, ties away (cvt.rna.tf32.f32)
    unsigned u;
    memcpy(&u, &v, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // interpolation taps of the production plan at sigma = 10: S = 4, sigma_b = 4.4, S * G_b(phase + S (5 - w))
    const double sb = 4.4;
    Taps taps;
    double T[S][W];
    for (int ph = 0; ph < S; ++ph)
        for (int w = 0; w < W; ++w) {
            const int d = ph + S * (5 - w);
            const double v = (abs(d) <= 20) ? S * exp(-0.5 * d * d / (sb * sb)) / (sb * sqrt(2 * M_PI)) : 0.0;
            T[ph][w] = (w == W - 1 || (w == 0 && ph != 0)) ? 0.0 : v;
            taps.g[ph * W + w] = make_float2((float)T[ph][w], (float)T[ph][w]);
        }
    // band matrix B[k][n] = T[n % 4][k - n / 4]  (k < 16 coarse columns, n < 16 fine columns), split into tf32 hi + lo
    std::vector<BFrag> bfr(32);
    for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        for (int ks = 0; ks < 2; ++ks)
            for (int n = 0; n < 2; ++n)
                for (int r = 0; r < 2; ++r) {
                    const int k = 8 * ks + t + 4 * r, nn = 8 * n + g;
                    const int w = k - nn / 4;
                    const float v = (w >= 0 && w < W) ? (float)T[nn % 4][w] : 0.f;
                    const float hi = tf32_round(v), lo = tf32_round(v - hi);
                    memcpy(&bfr[lane].hi[ks][n][r], &hi, 4);
                    memcpy(&bfr[lane].lo[ks][n][r], &lo, 4);
                }
    }
    // x-interpolated tiles: smooth complex fields with candidate-dependent amplitude (near-ties included)
    std::vector<float2> ha((size_t)NCAND * CY * PA), hb((size_t)NCAND * CYB * PB, make_float2(0.f, 0.f));
    srand(1);
    for (int c = 0; c < NCAND; ++c) {
        const double amp = 1.0 + 0.02 * cos(0.37 * c), kx = 0.011 * (c % 10), ky = 0.013 * (c / 10);
        for (int cy = 0; cy < CY; ++cy)
            for (int x = 0; x < TX; ++x) {
                const double ph = 2 * M_PI * (kx * x + ky * S * cy), nz = 0.05 * (rand() / (double)RAND_MAX - 0.5);
                const float2 v = make_float2((float)(amp * cos(ph) + nz), (float)(amp * sin(ph) - nz));
                ha[((size_t)c * CY + cy) * PA + x] = v;
                hb[((size_t)c * CYB + cy) * PB + x] = v;
            }
    }
    float2 *da, *db;
    BFrag* dbf;
    float *best_a, *best_b;
    int *idx_a, *idx_b;
    cudaMalloc(&da, ha.size() * sizeof(float2));
    cudaMalloc(&db, hb.size() * sizeof(float2));
    cudaMalloc(&dbf, 32 * sizeof(BFrag));
    cudaMalloc(&best_a, TX * TY * 4); cudaMalloc(&best_b, TX * TY * 4);
    cudaMalloc(&idx_a, TX * TY * 4); cudaMalloc(&idx_b, TX * TY * 4);
    cudaMemcpy(da, ha.data(), ha.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(dbf, bfr.data(), 32 * sizeof(BFrag), cudaMemcpyHostToDevice);
    const int grid = sms * 2 * 4;          // four waves of 2 CTAs per SM
    const size_t sa = 2 * CY * PA * sizeof(float2), sbm = 2 * CYB * PB * sizeof(float2);
    cudaFuncSetAttribute(k_interp_cuda, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_interp_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms_a = 1e30f, ms_b = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        float ms;
        cudaEventRecord(e0);
        k_interp_cuda<<<grid, 256, sa>>>(da, NCAND, taps, best_a, idx_a);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < ms_a) ms_a = ms;
        cudaEventRecord(e0);
        k_interp_mma<<<grid, 256, sbm>>>(db, NCAND, dbf, best_b, idx_b);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < ms_b) ms_b = ms;
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    std::vector<float> ba(TX * TY), bb(TX * TY);
    std::vector<int> ia(TX * TY), ib(TX * TY);
    cudaMemcpy(ba.data(), best_a, TX * TY * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(bb.data(), best_b, TX * TY * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ia.data(), idx_a, TX * TY * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ib.data(), idx_b, TX * TY * 4, cudaMemcpyDeviceToHost);
    double maxrel = 0;
    int differ = 0;
    for (int i = 0; i < TX * TY; ++i) {
        maxrel = fmax(maxrel, fabs((double)ba[i] - bb[i]) / ba[i]);
        differ += ia[i] != ib[i];
    }
    const double units = (double)grid * NCAND * TX * TY;      // pixel * candidates per launch
    printf("tiles: %d CTAs x %d candidates x %dx%d pixels\n", grid, NCAND, TX, TY);
    printf("A cuda-core FFMA2 : %.3f ms  %.1f Gpixel*cand/s\n", ms_a, units / ms_a / 1e6);
    printf("B mma.sync 3xTF32 : %.3f ms  %.1f Gpixel*cand/s   (B / A time = %.2f)\n", ms_b, units / ms_b / 1e6, ms_b / ms_a);
    printf("max relative |sf|^2 difference %.3g, winners that differ: %d of %d\n", maxrel, differ, TX * TY);
    return 0;
}
