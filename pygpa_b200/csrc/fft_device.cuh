// Shared-memory FFT used by K2 (DCT rows, unwrap.cu) and the Wiener deconvolution (wiener.cu).
#pragma once
#include <cuda_runtime.h>

namespace gpa {

// ------------------------------------------------------------------------------------------
// in-place Stockham FFT of buf[0..n) (forward, e^{-2 pi i jk/n}); all threads of the CTA.
// Radix-8 stages (a third of the shared-memory round trips of radix 2), preceded by one radix-2 or
// radix-4 stage when log2(n) is not a multiple of 3.  Every stage reads all its inputs into
// registers, synchronises, and writes in autosort order: no second buffer.  MAXB = radix-8
// butterflies per thread (blockDim.x * MAXB >= n / 8); tw[t] = e^{-2 pi i t/n}, t < n.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 zmul_mi(double2 a) { return make_double2(a.y, -a.x); }     // -i a

// x <- DFT_4(x), natural order
__device__ __forceinline__ void dft4(double2& x0, double2& x1, double2& x2, double2& x3) {
    const double2 t0 = zadd(x0, x2), t1 = zsub(x0, x2), t2 = zadd(x1, x3), t3 = zmul_mi(zsub(x1, x3));
    x0 = zadd(t0, t2);
    x1 = zadd(t1, t3);
    x2 = zsub(t0, t2);
    x3 = zsub(t1, t3);
}

// u <- DFT_8(u), natural order
__device__ __forceinline__ void dft8(double2 (&u)[8]) {
    const double r = 0.70710678118654752440084436210485;
    double2 a0 = zadd(u[0], u[4]), a1 = zadd(u[1], u[5]), a2 = zadd(u[2], u[6]), a3 = zadd(u[3], u[7]);
    double2 b0 = zsub(u[0], u[4]), b1 = zsub(u[1], u[5]), b2 = zsub(u[2], u[6]), b3 = zsub(u[3], u[7]);
    b1 = make_double2(r * (b1.x + b1.y), r * (b1.y - b1.x));       // * e^{-i pi/4}
    b2 = zmul_mi(b2);                                              // * e^{-i pi/2}
    b3 = make_double2(r * (b3.y - b3.x), -r * (b3.x + b3.y));      // * e^{-3 i pi/4}
    dft4(a0, a1, a2, a3);
    dft4(b0, b1, b2, b3);
    u[0] = a0; u[1] = b0; u[2] = a1; u[3] = b1; u[4] = a2; u[5] = b2; u[6] = a3; u[7] = b3;
}

// fft_pow2_g: the same transform run by a GROUP of nthr threads (tid = index inside the group) on the group's own
// buffer; every group of the CTA must call it with the same n (the barriers are CTA-wide).
template <int MAXB>
__device__ __forceinline__ void fft_pow2_g(double2* buf, int n, const double2* __restrict__ tw, int tid, int nthr);

template <int MAXB>
__device__ __forceinline__ void fft_pow2(double2* buf, int n, const double2* __restrict__ tw) {
    fft_pow2_g<MAXB>(buf, n, tw, (int)threadIdx.x, (int)blockDim.x);
}

template <int MAXB>
__device__ __forceinline__ void fft_pow2_g(double2* buf, int n, const double2* __restrict__ tw, const int tid, const int nthr) {
    const int lg = 31 - __clz(n);
    int ns = 1;
    if (lg % 3 == 1) {                               // one radix-2 stage (ns = 1: twiddles are 1)
        const int half = n >> 1;
        double2 a[4 * MAXB], b[4 * MAXB];
#pragma unroll
        for (int q = 0; q < 4 * MAXB; ++q) {
            const int j = tid + q * nthr;
            if (j < half) {
                a[q] = buf[j];
                b[q] = buf[j + half];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4 * MAXB; ++q) {
            const int j = tid + q * nthr;
            if (j < half) {
                buf[2 * j] = zadd(a[q], b[q]);
                buf[2 * j + 1] = zsub(a[q], b[q]);
            }
        }
        __syncthreads();
        ns = 2;
    } else if (lg % 3 == 2) {                        // one radix-4 stage
        const int quarter = n >> 2;
        double2 v[2 * MAXB][4];
#pragma unroll
        for (int q = 0; q < 2 * MAXB; ++q) {
            const int j = tid + q * nthr;
            if (j < quarter) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[q][i] = buf[j + i * quarter];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2 * MAXB; ++q) {
            const int j = tid + q * nthr;
            if (j < quarter) {
                dft4(v[q][0], v[q][1], v[q][2], v[q][3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) buf[4 * j + i] = v[q][i];
            }
        }
        __syncthreads();
        ns = 4;
    }
    const int e = n >> 3;
    for (; ns < n; ns <<= 3) {
        double2 u[MAXB][8];
        const int tstep = e / ns;                    // e^{-2 pi i q k/(8 ns)} = tw[q k n/(8 ns)]
#pragma unroll
        for (int q = 0; q < MAXB; ++q) {
            const int j = tid + q * nthr;
            if (j < e) {
#pragma unroll
                for (int i = 0; i < 8; ++i) u[q][i] = buf[j + i * e];
                if (ns > 1) {
                    const int k = j & (ns - 1);
                    const double2 w1 = __ldg(tw + k * tstep), w2 = __ldg(tw + 2 * k * tstep), w4 = __ldg(tw + 4 * k * tstep);
                    const double2 w3 = zmul(w1, w2), w5 = zmul(w1, w4), w6 = zmul(w2, w4);
                    const double2 w7 = zmul(w3, w4);
                    u[q][1] = zmul(u[q][1], w1);
                    u[q][2] = zmul(u[q][2], w2);
                    u[q][3] = zmul(u[q][3], w3);
                    u[q][4] = zmul(u[q][4], w4);
                    u[q][5] = zmul(u[q][5], w5);
                    u[q][6] = zmul(u[q][6], w6);
                    u[q][7] = zmul(u[q][7], w7);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < MAXB; ++q) {
            const int j = tid + q * nthr;
            if (j < e) {
                const int k = j & (ns - 1);
                const int j0 = ((j - k) << 3) + k;
                dft8(u[q]);
#pragma unroll
                for (int i = 0; i < 8; ++i) buf[j0 + i * ns] = u[q][i];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Bluestein (chirp-z) DFT of ANY length P inside shared memory, built on fft_pow2:
//   X[k] = c[k] sum_n (x[n] c[n]) conj(c)[k - n],  c[n] = exp(-i pi n^2 / P)
// i.e. a circular convolution of length L >= 2P - 1 (a power of two) = two FFT_L.
// ------------------------------------------------------------------------------------------
struct AxisPlan {
    int P, L;
    double2* chirp;     // [P]  exp(-i pi n^2 / P)
    double2* bhat;      // [L]  FFT_L of the wrapped conjugate chirp
    double2* tw;        // [L]  exp(-2 pi i t / L)
};

static __global__ void k_bs_tables(double2* chirp, double2* tw, int P, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        const long long q = ((long long)i * i) % (2LL * P);        // n^2 mod 2P: exact range reduction
        double s, c;
        sincospi(-(double)q / (double)P, &s, &c);
        chirp[i] = make_double2(c, s);
    }
    if (i < L) {
        double s, c;
        sincospi(-2.0 * (double)i / (double)L, &s, &c);
        tw[i] = make_double2(c, s);
    }
}

// bhat = FFT_L(b),  b[m mod L] = conj(chirp[|m|]), |m| < P        (one CTA)
template <int MAXB>
static __global__ void __launch_bounds__(512, 1) k_bs_prep(const AxisPlan ax) {
    extern __shared__ double2 bs_buf[];
    for (int i = threadIdx.x; i < ax.L; i += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (i < ax.P) v = ax.chirp[i];
        else if (ax.L - i < ax.P) v = ax.chirp[ax.L - i];
        bs_buf[i] = make_double2(v.x, -v.y);
    }
    __syncthreads();
    fft_pow2<MAXB>(bs_buf, ax.L, ax.tw);
    for (int i = threadIdx.x; i < ax.L; i += blockDim.x) ax.bhat[i] = bs_buf[i];
}

// buf[0..P) <- DFT_P(buf[0..P)); buf holds L entries (the rest is scratch); all threads of the CTA,
// which must have synchronised after writing buf.  On return buf[0..P) is visible to every thread.
template <int MAXB>
__device__ __forceinline__ void bluestein_dft(double2* buf, const AxisPlan& ax) {
    const int P = ax.P, L = ax.L;
    for (int n = threadIdx.x; n < L; n += blockDim.x)
        buf[n] = n < P ? zmul(buf[n], __ldg(ax.chirp + n)) : make_double2(0.0, 0.0);
    __syncthreads();
    fft_pow2<MAXB>(buf, L, ax.tw);
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        const double2 y = zmul(buf[k], __ldg(ax.bhat + k));
        buf[k] = make_double2(y.x, -y.y);                 // conj: the second forward FFT then inverts
    }
    __syncthreads();
    fft_pow2<MAXB>(buf, L, ax.tw);
    const double inv_l = 1.0 / (double)L;
    for (int k = threadIdx.x; k < P; k += blockDim.x) {
        const double2 v = zmul(make_double2(buf[k].x, -buf[k].y), __ldg(ax.chirp + k));
        buf[k] = make_double2(v.x * inv_l, v.y * inv_l);
    }
    __syncthreads();
}

static inline int bs_pow2_at_least(int v) {
    int l = 1;
    while (l < v) l <<= 1;
    return l;
}

// launch geometry of the shared-memory FFT of length L: threads, and whether a thread carries two butterflies
static inline void fft_launch_shape(int L, int& threads, int& per) {
    const int eighth = L / 8 > 0 ? L / 8 : 1;
    threads = eighth < 32 ? 32 : (eighth > 512 ? 512 : eighth);
    per = (eighth + threads - 1) / threads;
}

}  // namespace gpa
