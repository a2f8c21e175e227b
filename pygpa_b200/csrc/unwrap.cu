// K2 — weighted least-squares phase unwrapping: PCG with a DCT Poisson preconditioner (float64).
//
// Reference semantics: phase_unwrap / phase_unwrap_prediff (pyGPA/phase_unwrap.py:141-208,
// 282-350), solvePoisson_precomped (:95-103), precomp_Poissonscaling (:106-115, including the
// swapped N/M in the cosine arguments), applyQ (:118-132).  Ghiglia & Romero, JOSA A 11 (1994).
//
// Everything stays on the device: the PCG scalars (alpha, beta, the norms and the stop flag) live
// in device memory, every kernel returns immediately once the stop flag is set, and the host only
// enqueues (it polls the flag every few iterations to stop enqueueing early).
//
// DCT-II / DCT-III (scipy.fft.dctn / idctn, unnormalised): one row per CTA in shared memory.
// Makhoul's permutation turns each DCT into ONE n-point complex DFT per PAIR of rows: a radix-8
// Stockham FFT in shared memory for power-of-two n, a Bluestein chirp-z transform (two FFTs of
// length L >= 2n - 1) for any other n up to 4096 — iterate_GPA crops the frame by `edge`, so odd
// sizes are the norm there — and the O(n^2) cosine-table sum beyond that.  The 2-D
// transform is row pass -> transpose -> row pass; the 1/scale of the Poisson solve is fused into
// the second forward pass and <r, z> into the last inverse pass.
#include "common.cuh"
#include "fft_device.cuh"
#include "dct_pipe.cuh"
#include "lsq_device.cuh"

#include <utility>

namespace gpa {

constexpr double kPi = 3.141592653589793238462643383279;
constexpr int kMaxFftLen = 8192;   // 16 B * 8192 = 128 KB of shared memory per row

struct UwScalars {
    double rz, rz_prev, pqp, r0sq, rsq, alpha, beta;
    int done, k, kmax, pad;
    unsigned ticket[4];      // "last CTA finishes the reduction" counters, one per reducing kernel
};

enum { kTicketInit = 0, kTicketBeta = 1, kTicketAlpha = 2, kTicketStop = 3 };

// (v + pi) mod 2 pi - pi (phase_unwrap.py:135-138): the division-free form shared with K3 (lsq_device.cuh)
__device__ __forceinline__ double wrap_pi_d(double v) { return wrap_pi(v); }

// sum over the CTA (any multiple of 32 threads up to 1024); sh needs 32 doubles
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = lane < nw ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) sh[0] = r;
    }
    __syncthreads();
    r = sh[0];
    __syncthreads();
    return r;
}

// Tail of every reducing kernel: each CTA has written its partial sums; the CTA that takes the last
// ticket adds up all n partials (fixed order: the result does not depend on which CTA is last) and
// applies the scalar update `fin(total)` — the PCG scalars never need a launch of their own.
template <typename Fin>
__device__ __forceinline__ void finish_reduction(unsigned* ticket, const double* part, int n, double* sh, Fin fin) {
    __shared__ int s_last;
    __threadfence();                       // this CTA's partials are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;   // ... before its ticket
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(part + i);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
        fin(s);
        *ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------
// tables
// ------------------------------------------------------------------------------------------
// tw[t] = exp(-2 pi i t / n), t < n ;  mk[k] = exp(-i pi k / (2n)), k < n ; ct[j] = cos(pi j / (2n)), j < 4n
// cs[i] = cos(pi i / denom), i < n
__global__ void k_uw_cos_table(double* cs, int n, int denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cs[i] = cospi((double)i / (double)denom);
}

// mode 2: power-of-two FFT (tw, mk), 1: Bluestein FFT (mk only; the chirp tables come from k_bs_tables), 0: direct (ct)
__global__ void k_uw_tables(double2* tw, double2* mk, double* ct, int n, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode) {
        if (i < n) {
            double s, c;
            if (mode == 2) {
                sincospi(-2.0 * (double)i / (double)n, &s, &c);
                tw[i] = make_double2(c, s);
            }
            sincospi(-(double)i / (2.0 * (double)n), &s, &c);
            mk[i] = make_double2(c, s);
        }
    } else if (i < 4 * n) {
        ct[i] = cospi((double)i / (2.0 * (double)n));
    }
}

struct DctArgs {
    const double* in;
    double* out;
    int rows, n;               // rows of length n, contiguous
    const double2 *tw, *mk;    // FFT tables
    AxisPlan bs;               // Bluestein plan of this axis (bs.L == 0: n is a power of two)
    const double* ct;          // direct table
    // fused Poisson scaling (second forward pass; data is in the transposed layout: row = axis-1
    // frequency J, column = axis-0 frequency I): out /= 2 (cos(pi I / dimM) + cos(pi J / dimN) - 2)
    int fuse_scale, dimN, dimM;
    const double *cos_col, *cos_row;   // cos(pi k / dimM), k < n ; cos(pi row / dimN), row < rows
    // fused <r, z> (last inverse pass): partial[row] = sum_c out[row][c] * dot_with[row][c]
    const double* dot_with;
    double* partial;
    UwScalars* sc;
    // fused direction update of the PCG (pipelined inverse pass only): instead of storing the result z to `out`,
    // p <- z + beta p (p <- z in the first iteration), beta from sc (complete before this kernel starts)   (phase_unwrap.py:188-196)
    double* p_update;
};

// <r, z> is complete: k += 1, beta = rz / rz_prev (phase_unwrap.py:188-193)
__device__ __forceinline__ void sc_beta(UwScalars* sc, double s) {
    sc->k += 1;
    sc->rz = s;
    sc->beta = (sc->k == 1) ? 0.0 : s / sc->rz_prev;
    sc->rz_prev = s;
}

__device__ __forceinline__ double poisson_scale(int I, int J, int dimN, int dimM) {
    // phase_unwrap.py:109: 2 (cos(pi I / M) + cos(pi J / N) - 2), [0,0] := 1
    if (I == 0 && J == 0) return 1.0;
    return 2.0 * (cospi((double)I / (double)dimM) + cospi((double)J / (double)dimN) - 2.0);
}

// Forward DCT-II of every row, y[k] = 2 sum_m x[m] cos(pi k (2m+1) / (2n)), by Makhoul's
// permutation + one n-point complex FFT — of TWO rows at a time: rows 2b and 2b+1 ride in the real
// and imaginary parts, Z = FFT(v_a + i v_b), V_a[k] = (Z[k] + conj Z[n-k]) / 2,
// V_b[k] = (Z[k] - conj Z[n-k]) / 2i, y[k] = 2 Re(e^{-i pi k / 2n} V[k]).
template <int MAXB>
__global__ void __launch_bounds__(512, MAXB >= 2 ? 1 : 2) k_dct2_rows_pow2(const DctArgs a) {
    if (a.sc->done) return;
    extern __shared__ double2 cbuf[];
    const int n = a.n, row = 2 * blockIdx.x;
    const bool two = row + 1 < a.rows;
    const double* __restrict__ xa = a.in + (size_t)row * n;
    const double* __restrict__ xb = xa + n;
    // fixed trip count (blockDim.x * 8 MAXB >= n): every global load of the thread is issued before the
    // first shared-memory store waits on one (a runtime-bound loop serialised them: 24 % of the kernel)
    {
        double va[8 * MAXB], vb[8 * MAXB];
#pragma unroll
        for (int q = 0; q < 8 * MAXB; ++q) {
            const int j = threadIdx.x + q * blockDim.x;
            va[q] = j < n ? xa[j] : 0.0;
            vb[q] = (j < n && two) ? xb[j] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8 * MAXB; ++q) {
            const int j = threadIdx.x + q * blockDim.x;
            if (j < n) cbuf[(j & 1) ? n - 1 - (j >> 1) : (j >> 1)] = make_double2(va[q], vb[q]);
        }
    }
    __syncthreads();
    if (a.bs.L) bluestein_dft<MAXB>(cbuf, a.bs);
    else fft_pow2<MAXB>(cbuf, n, a.tw);
    double* __restrict__ ya = a.out + (size_t)row * n;
    double* __restrict__ yb = ya + n;
    const double cra = a.fuse_scale ? a.cos_row[row] : 0.0;
    const double crb = a.fuse_scale && two ? a.cos_row[row + 1] : 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double2 w = __ldg(a.mk + k), zk = cbuf[k], zn = cbuf[k ? n - k : 0];
        const double sx = zk.x + zn.x, sy = zk.y - zn.y;      // Z[k] + conj Z[n-k]
        const double dx = zk.x - zn.x, dy = zk.y + zn.y;      // Z[k] - conj Z[n-k]
        double ra = w.x * sx - w.y * sy;                      // Re(w (Z + conj Z'))
        double rb = w.x * dy + w.y * dx;                      // Im(w (Z - conj Z'))
        if (a.fuse_scale) {
            // phase_unwrap.py:109: 2 (cos(pi I / M) + cos(pi J / N) - 2), [0,0] := 1  (I = k, J = row)
            const double ck = a.cos_col[k];
            ra /= (k == 0 && row == 0) ? 1.0 : 2.0 * (ck + cra - 2.0);
            rb /= 2.0 * (ck + crb - 2.0);
        }
        ya[k] = ra;
        if (two) yb[k] = rb;
    }
}

// inverse (scipy idct type 2, norm=None), two rows per FFT: with V[k] = e^{+i pi k/2n} (y[k] - i y[n-k]) / 2
// the forward FFT of conj(V_a + i V_b) is n (v_a - i v_b)
template <int MAXB>
__global__ void __launch_bounds__(512, MAXB >= 2 ? 1 : 2) k_idct2_rows_pow2(const DctArgs a) {
    if (a.sc->done) return;
    extern __shared__ double2 cbuf[];
    __shared__ double red[32];
    const int n = a.n, row = 2 * blockIdx.x;
    const bool two = row + 1 < a.rows;
    const double* __restrict__ ya = a.in + (size_t)row * n;
    const double* __restrict__ yb = ya + n;
#pragma unroll
    for (int half = 0; half < 2; ++half) {             // two batches of 4 MAXB elements: 16 MAXB loads in flight
        double ak[4 * MAXB], ank[4 * MAXB], bk[4 * MAXB], bnk[4 * MAXB];
#pragma unroll
        for (int q = 0; q < 4 * MAXB; ++q) {           // fixed trip count: the loads are issued back to back
            const int k = threadIdx.x + (half * 4 * MAXB + q) * blockDim.x;
            const bool in = k < n;
            ak[q] = in ? ya[k] : 0.0;
            ank[q] = (in && k) ? ya[n - k] : 0.0;
            bk[q] = (in && two) ? yb[k] : 0.0;
            bnk[q] = (in && two && k) ? yb[n - k] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4 * MAXB; ++q) {
            const int k = threadIdx.x + (half * 4 * MAXB + q) * blockDim.x;
            if (k >= n) continue;
            const double2 w = __ldg(a.mk + k);             // e^{-i pi k/2n} = (c, -s)
            // e^{+i t}(yk - i ynk) = (c yk + s ynk) + i (s yk - c ynk), with c = w.x, s = -w.y
            const double re_a = 0.5 * (w.x * ak[q] - w.y * ank[q]), im_a = 0.5 * (-w.y * ak[q] - w.x * ank[q]);
            const double re_b = 0.5 * (w.x * bk[q] - w.y * bnk[q]), im_b = 0.5 * (-w.y * bk[q] - w.x * bnk[q]);
            cbuf[k] = make_double2(re_a - im_b, -im_a - re_b);   // conj(V_a) - i conj(V_b)
        }
    }
    __syncthreads();
    if (a.bs.L) bluestein_dft<MAXB>(cbuf, a.bs);
    else fft_pow2<MAXB>(cbuf, n, a.tw);
    double* __restrict__ xa = a.out + (size_t)row * n;
    double* __restrict__ xb = xa + n;
    const double inv = 1.0 / (double)n;
    double dot_a = 0.0, dot_b = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int src = (j & 1) ? n - 1 - (j >> 1) : (j >> 1);
        const double2 f = cbuf[src];
        const double va = f.x * inv, vb = -f.y * inv;
        xa[j] = va;
        if (two) xb[j] = vb;
        if (a.dot_with) {
            dot_a = fma(va, a.dot_with[(size_t)row * n + j], dot_a);
            if (two) dot_b = fma(vb, a.dot_with[(size_t)(row + 1) * n + j], dot_b);
        }
    }
    if (a.dot_with) {
        const double sa = block_sum(dot_a, red);
        const double sb = block_sum(dot_b, red);
        if (threadIdx.x == 0) {
            a.partial[row] = sa;
            if (two) a.partial[row + 1] = sb;
        }
        UwScalars* sc = a.sc;
        finish_reduction(&sc->ticket[kTicketBeta], a.partial, a.rows, red, [sc](double s) { sc_beta(sc, s); });
    }
}

// ---- pipelined forms (dct_pipe.cuh): persistent CTAs of n / 8 threads, pair p = blockIdx.x, blockIdx.x + gridDim.x, ...
template <int LOGN>
__global__ void __launch_bounds__(1 << (LOGN - 3), LOGN == 12 ? 1 : 768 >> (LOGN - 3)) k_dct2_rows_pipe(const DctArgs a) {
    if (a.sc->done) return;
    constexpr int n = 1 << LOGN, T = n >> 3;
    extern __shared__ __align__(128) unsigned char dp_smem[];
    __shared__ unsigned long long mbar;
    double* const stage = reinterpret_cast<double*>(dp_smem);                  // rows 2b | 2b + 1 as they lie in memory
    double2* const buf = reinterpret_cast<double2*>(dp_smem + 2 * n * sizeof(double));
    double2* const t8 = buf + n + n / 8;
    const int tid = threadIdx.x, pairs = (a.rows + 1) >> 1, G = gridDim.x;
    auto row_bytes = [&](int pair) -> unsigned { return (2 * pair + 1 < a.rows ? 2u : 1u) * n * (unsigned)sizeof(double); };
    if (tid == 0) {
        dp_mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    dp_tw_fill<LOGN>(t8, a.tw, tid, T);
    const double2 mk0 = a.mk[tid];                       // e^{-i pi tid/2n}; position tid + m n/8 needs mk0 e^{-i pi m/16}
    __syncthreads();
    int pair = blockIdx.x;
    if (tid == 0 && pair < pairs) dp_bulk_load(stage, a.in + (size_t)2 * pair * n, row_bytes(pair), &mbar);
    unsigned parity = 0;
    for (; pair < pairs; pair += G) {
        const int row = 2 * pair;
        const bool two = row + 1 < a.rows;
        dp_mbar_wait(&mbar, parity);
        parity ^= 1u;
        // Makhoul order: v[p] = x[2p] (p < n/2), x[2(n-1-p)+1] (p >= n/2); rows a, b in the real / imaginary part
        auto get = [&](int m) -> double2 {
            const int p = tid + m * T;
            const int src = m < 4 ? 2 * p : 2 * (n - 1 - p) + 1;
            return make_double2(stage[src], two ? stage[n + src] : 0.0);
        };
        const int next = pair + G;
        dp_fft<LOGN>(buf, t8, tid, get, [] { __syncthreads(); }, [&] {
            if (tid == 0 && next < pairs) dp_bulk_load(stage, a.in + (size_t)2 * next * n, row_bytes(next), &mbar);
        }, [](int p) { return dp_pad(p); });
        double* __restrict__ ya = a.out + (size_t)row * n;
        double* __restrict__ yb = ya + n;
        const double cra = a.fuse_scale ? a.cos_row[row] : 0.0;
        const double crb = a.fuse_scale && two ? a.cos_row[row + 1] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = tid + q * T;
            const double2 w = zmul(mk0, dp_rot16(q)), zk = buf[dp_pad(k)], zn = buf[dp_pad(k ? n - k : 0)];
            const double sx = zk.x + zn.x, sy = zk.y - zn.y;      // Z[k] + conj Z[n-k]
            const double dx = zk.x - zn.x, dy = zk.y + zn.y;      // Z[k] - conj Z[n-k]
            double ra = w.x * sx - w.y * sy;                      // Re(w (Z + conj Z'))
            double rb = w.x * dy + w.y * dx;                      // Im(w (Z - conj Z'))
            if (a.fuse_scale) {
                const double ck = a.cos_col[k];
                ra /= (k == 0 && row == 0) ? 1.0 : 2.0 * (ck + cra - 2.0);
                rb /= 2.0 * (ck + crb - 2.0);
            }
            ya[k] = ra;
            if (two) yb[k] = rb;
        }
    }
}

// MODE 0: out = idct rows; 1: + partial sums of <out, dot_with>; 2: p_update <- out + beta p_update instead of storing out
template <int LOGN, int MODE>
__global__ void __launch_bounds__(1 << (LOGN - 3), LOGN == 12 ? 1 : 768 >> (LOGN - 3)) k_idct2_rows_pipe(const DctArgs a) {
    if (a.sc->done) return;
    constexpr int n = 1 << LOGN, T = n >> 3;
    extern __shared__ __align__(128) unsigned char dp_smem[];
    __shared__ unsigned long long mbar;
    __shared__ double red[32];
    double* const stage = reinterpret_cast<double*>(dp_smem);
    double2* const buf = reinterpret_cast<double2*>(dp_smem + 2 * n * sizeof(double));
    double2* const t8 = buf + n + n / 8;
    const int tid = threadIdx.x, pairs = (a.rows + 1) >> 1, G = gridDim.x;
    auto row_bytes = [&](int pair) -> unsigned { return (2 * pair + 1 < a.rows ? 2u : 1u) * n * (unsigned)sizeof(double); };
    if (tid == 0) {
        dp_mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    dp_tw_fill<LOGN>(t8, a.tw, tid, T);
    const double2 mk0 = a.mk[tid];                       // e^{-i pi tid/2n}; position tid + m n/8 needs mk0 e^{-i pi m/16}
    __syncthreads();
    int pair = blockIdx.x;
    if (tid == 0 && pair < pairs) dp_bulk_load(stage, a.in + (size_t)2 * pair * n, row_bytes(pair), &mbar);
    unsigned parity = 0;
    const double inv = 1.0 / (double)n;
    constexpr bool DOT = MODE == 1, PUPD = MODE == 2;
    const double beta = PUPD ? a.sc->beta : 0.0;
    const bool first = PUPD ? a.sc->k == 1 : true;
    for (; pair < pairs; pair += G) {
        const int row = 2 * pair;
        const bool two = row + 1 < a.rows;
        dp_mbar_wait(&mbar, parity);
        parity ^= 1u;
        // FFT input at k: conj(V_a) - i conj(V_b), V[k] = e^{+i pi k/2n} (y[k] - i y[n-k]) / 2     (as k_idct2_rows_pow2)
        auto get = [&](int m) -> double2 {
            const int k = tid + m * T;
            const double2 w = zmul(mk0, dp_rot16(m));      // e^{-i pi k/2n} = (c, -s)
            const double ak = stage[k], ank = k ? stage[n - k] : 0.0;
            const double bk = two ? stage[n + k] : 0.0, bnk = (two && k) ? stage[2 * n - k] : 0.0;
            const double re_a = 0.5 * (w.x * ak - w.y * ank), im_a = 0.5 * (-w.y * ak - w.x * ank);
            const double re_b = 0.5 * (w.x * bk - w.y * bnk), im_b = 0.5 * (-w.y * bk - w.x * bnk);
            return make_double2(re_a - im_b, -im_a - re_b);
        };
        const int next = pair + G;
        dp_fft<LOGN>(buf, t8, tid, get, [] { __syncthreads(); }, [&] {
            if (tid == 0 && next < pairs) dp_bulk_load(stage, a.in + (size_t)2 * next * n, row_bytes(next), &mbar);
        }, [](int p) { return dp_pad(p); });
        double* __restrict__ xa = (PUPD ? a.p_update : a.out) + (size_t)row * n;
        double* __restrict__ xb = xa + n;
        double dot_a = 0.0, dot_b = 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                      // two batches of 8 loads in flight (16 at once spill at 80 registers)
            double da[4], db[4];
            if (DOT) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = tid + (4 * h + q) * T;
                    da[q] = a.dot_with[(size_t)row * n + j];
                    db[q] = two ? a.dot_with[(size_t)(row + 1) * n + j] : 0.0;
                }
            }
            if (PUPD && !first) {                          // the old direction, 8 loads in flight
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = tid + (4 * h + q) * T;
                    da[q] = xa[j];
                    db[q] = two ? xb[j] : 0.0;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = tid + (4 * h + q) * T;
                const int src = (j & 1) ? n - 1 - (j >> 1) : (j >> 1);
                const double2 f = buf[dp_pad(src)];
                double va = f.x * inv, vb = -f.y * inv;
                if (PUPD && !first) {
                    va = fma(beta, da[q], va);
                    vb = fma(beta, db[q], vb);
                }
                xa[j] = va;
                if (two) xb[j] = vb;
                if (DOT) {
                    dot_a = fma(va, da[q], dot_a);
                    if (two) dot_b = fma(vb, db[q], dot_b);
                }
            }
        }
        if (DOT) {
            const double sa = block_sum(dot_a, red);
            const double sb = block_sum(dot_b, red);
            if (tid == 0) {
                a.partial[row] = sa;
                if (two) a.partial[row + 1] = sb;
            }
        }
    }
    if (DOT) {
        UwScalars* sc = a.sc;
        finish_reduction(&sc->ticket[kTicketBeta], a.partial, a.rows, red, [sc](double s) { sc_beta(sc, s); });
    }
}

// any length: direct cosine sums from the table ct[j] = cos(pi j / 2n), j < 4n
template <int INVERSE>
__global__ void k_dct2_rows_direct(const DctArgs a) {
    if (a.sc->done) return;
    extern __shared__ double rbuf[];
    __shared__ double red[32];
    const int n = a.n, row = blockIdx.x, n4 = 4 * n;
    const double* __restrict__ in = a.in + (size_t)row * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) rbuf[j] = in[j];
    __syncthreads();
    double* __restrict__ out = a.out + (size_t)row * n;
    double dot = 0.0;
    for (int o = threadIdx.x; o < n; o += blockDim.x) {
        double acc = 0.0;
        if (!INVERSE) {
            int idx = o % n4;                      // k (2m+1) mod 4n, m = 0
            const int step = (2 * o) % n4;
            for (int m = 0; m < n; ++m) {
                acc = fma(rbuf[m], __ldg(a.ct + idx), acc);
                idx += step;
                if (idx >= n4) idx -= n4;
            }
            acc *= 2.0;
            if (a.fuse_scale) acc /= poisson_scale(o, row, a.dimN, a.dimM);
        } else {
            const int step = (2 * o + 1) % n4;     // k (2m+1) mod 4n, k = 1, 2, ...
            int idx = step;
            for (int k = 1; k < n; ++k) {
                acc = fma(rbuf[k], __ldg(a.ct + idx), acc);
                idx += step;
                if (idx >= n4) idx -= n4;
            }
            acc = (rbuf[0] + 2.0 * acc) / (2.0 * (double)n);
            if (a.dot_with) dot = fma(acc, a.dot_with[(size_t)row * n + o], dot);
        }
        out[o] = acc;
    }
    if (INVERSE && a.dot_with) {
        const double s = block_sum(dot, red);
        if (threadIdx.x == 0) a.partial[row] = s;
        UwScalars* sc = a.sc;
        finish_reduction(&sc->ticket[kTicketBeta], a.partial, a.rows, red, [sc](double t) { sc_beta(sc, t); });
    }
}

// Fused column stage of the Poisson solve for power-of-two N: the CTA owns a strip of CW columns of the (N, M) array
// (row-DCT coefficients), transforms every column (DCT-II along axis 0), divides by the Poisson scale and transforms
// back (DCT-III), all inside shared memory: z <- idct_0(dct_0(z) / scale).  One read and one write of the array instead
// of transpose + row pass + row pass + transpose (6 passes -> 2).  Columns are paired like the rows of
// k_dct2_rows_pow2 (columns c, c+1 ride in the real / imaginary part of one complex FFT); a group of n / (8 MAXB)
// threads runs each FFT, CW / 2 groups per CTA.  Global accesses are CW x 8 B contiguous per row (whole sectors for
// CW >= 4).  The arithmetic per element is that of k_dct2_rows_pow2 (epilogue) followed by k_idct2_rows_pow2 (prologue).
struct ColArgs {
    double* z;                 // (N, M) row-major, in place
    int N, M, CW, tpf;         // strip width, threads per FFT
    const double2 *tw, *mk;    // FFT tables of length N
    const double *cos_k, *cos_c;   // cos(pi k / M), k < N ; cos(pi c / N), c < M      (phase_unwrap.py:109, swapped on purpose)
    const UwScalars* sc;
};

template <int MAXB>
__global__ void __launch_bounds__(MAXB == 1 ? 1024 : 512, 1) k_poisson_cols(const ColArgs a) {
    if (a.sc->done) return;
    extern __shared__ double2 cbuf[];
    const int n = a.N, M = a.M, CW = a.CW, tpf = a.tpf;
    const int gstride = n + 1;                      // one complex of padding: the groups' buffers start 4 banks apart
    const int c0 = blockIdx.x * CW;
    const int nthr = blockDim.x;
    double* const flat = reinterpret_cast<double*>(cbuf);
    // ---- load the strip, Makhoul order, columns (2g, 2g+1) -> (re, im) of group g
    for (int base = 0; base < n * CW; base += nthr * 16) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int idx = base + threadIdx.x + q * nthr;
            const int r = idx / CW, cc = idx - r * CW;
            v[q] = (idx < n * CW && c0 + cc < M) ? a.z[(size_t)r * M + c0 + cc] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int idx = base + threadIdx.x + q * nthr;
            if (idx < n * CW) {
                const int r = idx / CW, cc = idx - r * CW;
                const int pos = (r & 1) ? n - 1 - (r >> 1) : (r >> 1);
                flat[((size_t)(cc >> 1) * gstride + pos) * 2 + (cc & 1)] = v[q];
            }
        }
    }
    __syncthreads();
    const int g = threadIdx.x / tpf, gt = threadIdx.x - g * tpf;
    double2* const buf = cbuf + (size_t)g * gstride;
    fft_pow2_g<MAXB>(buf, n, a.tw, gt, tpf);
    // ---- DCT-II epilogue, Poisson scale, DCT-III prologue: pairs (k, n - k) are independent of each other
    {
        const int ca = c0 + 2 * g, cb = ca + 1;                  // the two columns of this group (axis-1 frequencies J)
        const double cra = ca < M ? a.cos_c[ca] : 0.0, crb = cb < M ? a.cos_c[cb] : 0.0;
        const int half = n >> 1;
        for (int k = gt; k <= half; k += tpf) {
            const int kn = k ? n - k : 0;
            const double2 zk = buf[k], zn = buf[kn];
            const double2 wk = __ldg(a.mk + k);
            // forward coefficients at k (as k_dct2_rows_pow2)
            double ra_k = wk.x * (zk.x + zn.x) - wk.y * (zk.y - zn.y);
            double rb_k = wk.x * (zk.y + zn.y) + wk.y * (zk.x - zn.x);
            const double ck = a.cos_k[k];
            ra_k /= (k == 0 && ca == 0) ? 1.0 : 2.0 * (ck + cra - 2.0);
            rb_k /= 2.0 * (ck + crb - 2.0);
            double ra_n = 0.0, rb_n = 0.0;
            double2 wn = make_double2(0.0, 0.0);
            if (k != 0) {
                if (k == half) {
                    ra_n = ra_k; rb_n = rb_k; wn = wk;
                } else {      // forward coefficients at n - k: Z and conj-partner swap roles
                    wn = __ldg(a.mk + kn);
                    ra_n = wn.x * (zn.x + zk.x) - wn.y * (zn.y - zk.y);
                    rb_n = wn.x * (zn.y + zk.y) + wn.y * (zn.x - zk.x);
                    const double cn = a.cos_k[kn];
                    ra_n /= 2.0 * (cn + cra - 2.0);
                    rb_n /= 2.0 * (cn + crb - 2.0);
                }
            }
            // inverse prologue at k (as k_idct2_rows_pow2): partner value is the coefficient at n - k (0 for k = 0)
            {
                const double re_a = 0.5 * (wk.x * ra_k - wk.y * ra_n), im_a = 0.5 * (-wk.y * ra_k - wk.x * ra_n);
                const double re_b = 0.5 * (wk.x * rb_k - wk.y * rb_n), im_b = 0.5 * (-wk.y * rb_k - wk.x * rb_n);
                buf[k] = make_double2(re_a - im_b, -im_a - re_b);
            }
            if (k != 0 && k != half) {
                const double re_a = 0.5 * (wn.x * ra_n - wn.y * ra_k), im_a = 0.5 * (-wn.y * ra_n - wn.x * ra_k);
                const double re_b = 0.5 * (wn.x * rb_n - wn.y * rb_k), im_b = 0.5 * (-wn.y * rb_n - wn.x * rb_k);
                buf[kn] = make_double2(re_a - im_b, -im_a - re_b);
            }
        }
    }
    __syncthreads();
    fft_pow2_g<MAXB>(buf, n, a.tw, gt, tpf);
    // ---- store: x[j] = Re / -Im of F[perm(j)] / n
    const double inv = 1.0 / (double)n;
    for (int idx = threadIdx.x; idx < n * CW; idx += nthr) {
        const int r = idx / CW, cc = idx - r * CW;
        if (c0 + cc < M) {
            const int pos = (r & 1) ? n - 1 - (r >> 1) : (r >> 1);
            const double f = flat[((size_t)(cc >> 1) * gstride + pos) * 2 + (cc & 1)];
            a.z[(size_t)r * M + c0 + cc] = (cc & 1) ? -f * inv : f * inv;
        }
    }
}

// Pipelined column stage of the Poisson solve: z <- idct_0(dct_0(z) / scale), in place (same arithmetic as k_poisson_cols).
// A CTA of n / 4 threads owns a strip of FOUR columns = one 32-byte sector per row: the strip arrives as TMA boxes
// (cp.async.bulk.tensor.2d, 256 rows x 4 columns each) in the row-major layout [row][4], which IS two interleaved complex
// sequences — (column 0, column 1) and (column 2, column 3) of a row form the double2 elements 2 row + g of the FFTs g = 0, 1
// (two real columns per complex FFT, as in the row kernels).  Thread (j, g) = (tid >> 1, tid & 1) runs butterfly j of FFT g, so
// a warp's shared-memory accesses interleave the two FFTs and stay conflict-free in the padded layout 2 pad(p) + g.  The head
// stage of the forward FFT reads the raw strip through the Makhoul permutation; everything else happens in place; the
// result goes back as TMA box stores.  Persistent CTAs (two per SM at n = 2048, running in different phases) take the
// strips round-robin.  Columns beyond M are zero-filled by the loads and clipped by the stores.
struct ColPipeArgs {
    int N, M;
    const double2 *tw, *mk;
    const double *cos_k, *cos_c;   // cos(pi k / M), k < N ; cos(pi c / N), c < M      (phase_unwrap.py:109, swapped on purpose)
    UwScalars* sc;
    // <r, z> of the PCG without another pass over the arrays: the DCT-II is orthogonal up to the weights
    // sum_n x[n] y[n] = X[0] Y[0] / 4n + sum_{k>0} X[k] Y[k] / 2n per axis, and this kernel holds both dctn(r) (before the
    // Poisson scale) and dctn(z) (after it) in registers.  partial = one value per CTA (null: not wanted).
    double* partial;
};

template <int LOGN>
__global__ void __launch_bounds__(1 << (LOGN - 2), LOGN >= 12 ? 1 : 1 << (12 - LOGN) > 16 ? 16 : 1 << (12 - LOGN))
k_poisson_cols_pipe(const ColPipeArgs a, const __grid_constant__ CUtensorMap tmap) {
    if (a.sc->done) return;
    constexpr int n = 1 << LOGN, T = n >> 3, half = n >> 1;
    constexpr int BOXR = n < 256 ? n : 256;                 // rows per TMA box
    extern __shared__ __align__(128) unsigned char dp_smem[];
    __shared__ unsigned long long mbar;
    __shared__ double red[32];
    double2* const buf = reinterpret_cast<double2*>(dp_smem);
    double2* const t8 = buf + 2 * (n + n / 8);
    const int tid = threadIdx.x, g = tid & 1, j = tid >> 1;
    double dot = 0.0;
    const double wi0 = 0.25 / (double)n, wi1 = 0.5 / (double)n, wj0 = 0.25 / (double)a.M, wj1 = 0.5 / (double)a.M;
    const int units = (a.M + 3) >> 2;
    if (tid == 0) {
        dp_mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    dp_tw_fill<LOGN>(t8, a.tw, tid, 2 * T);
    const double2 mk0 = a.mk[j];
    __syncthreads();
    auto idx = [g](int p) { return 2 * dp_pad(p) + g; };
    auto sync = [] { __syncthreads(); };
    auto noop = [] {};
    unsigned parity = 0;
    const double inv = 1.0 / (double)n;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int c0 = 4 * unit;
        if (tid == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous strip's stores have read the buffer
            dp_mbar_expect_tx(&mbar, (unsigned)(n * 4 * sizeof(double)));
#pragma unroll 1
            for (int b = 0; b < n / BOXR; ++b) dp_tma_load_2d(dp_smem + (size_t)b * BOXR * 32, &tmap, &mbar, c0, b * BOXR);
        }
        dp_mbar_wait(&mbar, parity);
        parity ^= 1u;
        // ---- forward FFT of the permuted columns: v[p] = x[2p] (p < n/2), x[2(n-1-p)+1] (p >= n/2)
        dp_fft<LOGN>(buf, t8, j, [&](int m) -> double2 {
            const int p = j + m * T;
            const int src = m < 4 ? 2 * p : 2 * (n - 1 - p) + 1;
            return buf[2 * src + g];
        }, sync, noop, idx);
        // ---- DCT-II epilogue, Poisson scale, DCT-III prologue on the pairs (k, n - k)
        {
            const int ca = c0 + 2 * g, cb = ca + 1;              // the two columns of this FFT (axis-1 frequencies J)
            const double cra = ca < a.M ? a.cos_c[ca] : 0.0, crb = cb < a.M ? a.cos_c[cb] : 0.0;
            const double wja = ca < a.M ? (ca == 0 ? wj0 : wj1) : 0.0, wjb = cb < a.M ? wj1 : 0.0;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int k = j + q * T;                         // q = 4: k = n/2 (thread j = 0 only)
                if (q == 4 && j != 0) break;
                const int kn = k ? n - k : 0;
                const double2 zk = buf[idx(k)], zn = buf[idx(kn)];
                const double2 wk = q == 4 ? make_double2(0.70710678118654752440084436210485, -0.70710678118654752440084436210485)
                                          : zmul(mk0, dp_rot16(q));           // e^{-i pi k/2n}
                const double2 wn = make_double2(-wk.y, -wk.x);             // e^{-i pi (n-k)/2n} = -i conj(wk)
                double ra_k = wk.x * (zk.x + zn.x) - wk.y * (zk.y - zn.y);
                double rb_k = wk.x * (zk.y + zn.y) + wk.y * (zk.x - zn.x);
                const double ck = __ldg(a.cos_k + k);
                {
                    const double ua = ra_k, ub = rb_k;
                    ra_k /= (k == 0 && ca == 0) ? 1.0 : 2.0 * (ck + cra - 2.0);
                    rb_k /= 2.0 * (ck + crb - 2.0);
                    dot = fma(k == 0 ? wi0 : wi1, fma(wja * ua, ra_k, wjb * ub * rb_k), dot);
                }
                double ra_n = 0.0, rb_n = 0.0;
                if (k != 0) {
                    if (k == half) {
                        ra_n = ra_k; rb_n = rb_k;
                    } else {      // forward coefficients at n - k: Z and its conjugate partner swap roles
                        ra_n = wn.x * (zn.x + zk.x) - wn.y * (zn.y - zk.y);
                        rb_n = wn.x * (zn.y + zk.y) + wn.y * (zn.x - zk.x);
                        const double cn = __ldg(a.cos_k + kn);
                        const double ua = ra_n, ub = rb_n;
                        ra_n /= 2.0 * (cn + cra - 2.0);
                        rb_n /= 2.0 * (cn + crb - 2.0);
                        dot = fma(wi1, fma(wja * ua, ra_n, wjb * ub * rb_n), dot);
                    }
                }
                {
                    const double re_a = 0.5 * (wk.x * ra_k - wk.y * ra_n), im_a = 0.5 * (-wk.y * ra_k - wk.x * ra_n);
                    const double re_b = 0.5 * (wk.x * rb_k - wk.y * rb_n), im_b = 0.5 * (-wk.y * rb_k - wk.x * rb_n);
                    buf[idx(k)] = make_double2(re_a - im_b, -im_a - re_b);
                }
                if (k != 0 && k != half) {
                    const double re_a = 0.5 * (wn.x * ra_n - wn.y * ra_k), im_a = 0.5 * (-wn.y * ra_n - wn.x * ra_k);
                    const double re_b = 0.5 * (wn.x * rb_n - wn.y * rb_k), im_b = 0.5 * (-wn.y * rb_n - wn.x * rb_k);
                    buf[idx(kn)] = make_double2(re_a - im_b, -im_a - re_b);
                }
            }
        }
        __syncthreads();
        // ---- inverse transform (forward FFT of the conjugated input), in place
        dp_fft<LOGN>(buf, t8, j, [&](int m) -> double2 { return buf[idx(j + m * T)]; }, sync, noop, idx);
        // ---- x[r] = Re / -Im of F[perm(r)] / n, back into the row-major strip, then out by TMA
        double2 f[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = j + q * T;
            f[q] = buf[idx((r & 1) ? n - 1 - (r >> 1) : (r >> 1))];
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) buf[2 * (j + q * T) + g] = make_double2(f[q].x * inv, -f[q].y * inv);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes visible to the TMA store
        __syncthreads();
        if (tid == 0) {
#pragma unroll 1
            for (int b = 0; b < n / BOXR; ++b) dp_tma_store_2d(&tmap, dp_smem + (size_t)b * BOXR * 32, c0, b * BOXR);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // stores complete before the CTA retires
    if (a.partial) {
        const double s = block_sum(dot, red);
        if (tid == 0) a.partial[blockIdx.x] = s;
        UwScalars* sc = a.sc;
        finish_reduction(&sc->ticket[kTicketBeta], a.partial, gridDim.x, red, [sc](double t) { sc_beta(sc, t); });
    }
}

__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, int rows, int cols,
                            const UwScalars* sc) {
    if (sc->done) return;
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------
// PCG pieces
// ------------------------------------------------------------------------------------------
struct SetupArgs {
    const double *psi, *dx, *dy, *weight;   // psi (N,M) or dx (N,M-1) & dy (N-1,M); weight (N,M) or null
    double *wwx, *wwy, *r, *phi, *partial;
    UwScalars* sc;
    int N, M, kmax;
};

__device__ __forceinline__ double edge_w(const double* w, size_t i, size_t j) {
    if (!w) return 1.0;
    const double a = w[i] * w[i], b = w[j] * w[j];
    return fmin(a, b);    // phase_unwrap.py:166-167
}

// weighted right-hand side r = A^T W^T W b, edge weights, phi = 0        (phase_unwrap.py:154-176)
constexpr int kUwRows = 32;     // rows per CTA of the stencil kernels (4 rows x 64 columns per step)

// PSI: wrapped differences of psi, else the given gradients dx, dy; WEIGHT: weighted.  The switches are template parameters
// so that a pixel's 5 + 5 (4 + 5) loads sit in one basic block and are issued back to back (with runtime switches every load
// stayed inside its branch: the kernel ran at 2.4 TB/s).
template <bool PSI, bool WEIGHT>
__global__ void __launch_bounds__(256) k_uw_setup(const SetupArgs a) {
    __shared__ double red[32];
    const int N = a.N, M = a.M;
    double rsq = 0.0;
    const int tiles_x = (M + 63) >> 6, n_sub = tiles_x * ((N + 3) >> 2);     // sub-tiles of 4 rows x 64 columns, round-robin
    for (int t = blockIdx.x + blockIdx.y * gridDim.x; t < n_sub; t += gridDim.x * gridDim.y) {
        const int by = t / tiles_x, bx = t - by * tiles_x;
        const int c = bx * 64 + (threadIdx.x & 63);
        const int r = by * 4 + (threadIdx.x >> 6);
        if (r < N && c < M) {
            const size_t i = (size_t)r * M + c;
            const bool hr = c < M - 1, hl = c > 0, hd = r < N - 1, hu = r > 0;     // right / left / down / up neighbour exists
            // ---- loads (neighbour indices clamped to the pixel itself where there is no neighbour)
            double pc = 0.0, pr = 0.0, pl = 0.0, pd = 0.0, pu = 0.0;               // psi, or the four gradients in pr, pl, pd, pu
            if (PSI) {
                pc = a.psi[i];
                pr = a.psi[hr ? i + 1 : i];
                pl = a.psi[hl ? i - 1 : i];
                pd = a.psi[hd ? i + M : i];
                pu = a.psi[hu ? i - M : i];
            } else {
                const size_t ix = (size_t)r * (M - 1) + c;
                pr = hr ? a.dx[ix] : 0.0;
                pl = hl ? a.dx[ix - 1] : 0.0;
                pd = hd ? a.dy[i] : 0.0;
                pu = hu ? a.dy[i - M] : 0.0;
            }
            double wc = 1.0, wr = 1.0, wl = 1.0, wd = 1.0, wu = 1.0;
            if (WEIGHT) {
                wc = a.weight[i];
                wr = a.weight[hr ? i + 1 : i];
                wl = a.weight[hl ? i - 1 : i];
                wd = a.weight[hd ? i + M : i];
                wu = a.weight[hu ? i - M : i];
            }
            // ---- wrapped differences and edge weights min(w_a^2, w_b^2)   (phase_unwrap.py:154-176)
            const double bxr = wrap_pi_d(PSI ? pr - pc : pr), bxl = wrap_pi_d(PSI ? pc - pl : pl);
            const double byd = wrap_pi_d(PSI ? pd - pc : pd), byu = wrap_pi_d(PSI ? pc - pu : pu);
            const double c2 = wc * wc;
            const double er = WEIGHT ? fmin(c2, wr * wr) : 1.0, el = WEIGHT ? fmin(wl * wl, c2) : 1.0;
            const double ed = WEIGHT ? fmin(c2, wd * wd) : 1.0, eu = WEIGHT ? fmin(wu * wu, c2) : 1.0;
            double v = 0.0;
            if (hr) {
                a.wwx[(size_t)r * (M - 1) + c] = er;
                v += er * bxr;
            }
            if (hl) v -= el * bxl;
            if (hd) {
                a.wwy[i] = ed;
                v += ed * byd;
            }
            if (hu) v -= eu * byu;
            a.r[i] = v;
            a.phi[i] = 0.0;
            rsq = fma(v, v, rsq);
        }
    }
    const double s = block_sum(rsq, red);
    if (threadIdx.x == 0) a.partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
    UwScalars* sc = a.sc;
    const int kmax = a.kmax;
    finish_reduction(&sc->ticket[kTicketInit], a.partial, gridDim.x * gridDim.y, red, [sc, kmax](double t) {
        sc->r0sq = t;
        sc->rsq = t;
        sc->rz = sc->rz_prev = sc->pqp = sc->alpha = sc->beta = 0.0;
        sc->k = 0;
        sc->kmax = kmax;
        sc->done = (t == 0.0);        // `while not all(rk == 0)`: nothing to do
    });
}

// p = z + beta p  (first iteration: p = z)
__global__ void __launch_bounds__(256) k_uw_update_p(const double* __restrict__ z, double* __restrict__ p, size_t n,
                                                     const UwScalars* sc) {
    if (sc->done) return;
    const double beta = sc->beta;
    const bool first = sc->k == 1;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        p[i] = first ? z[i] : fma(beta, p[i], z[i]);
}

// q = Q p = A^T W^T W A p (phase_unwrap.py:118-132) and partial sums of <p, q>
__global__ void __launch_bounds__(256) k_uw_apply_q(const double* __restrict__ p, const double* __restrict__ wwx,
                                                    const double* __restrict__ wwy, double* __restrict__ q,
                                                    double* __restrict__ partial, int N, int M, UwScalars* sc) {
    if (sc->done) return;
    __shared__ double red[32];
    // sub-tiles of 4 rows x 64 columns, taken round-robin by a grid that fills the machine exactly once (a grid of one CTA
    // per 32 x 64 pixels was 1.73 waves at 2048^2: the second wave ran 73 % full)
    const int tiles_x = (M + 63) >> 6, n_sub = tiles_x * ((N + 3) >> 2);
    double pq = 0.0;
    for (int t = blockIdx.x + blockIdx.y * gridDim.x; t < n_sub; t += gridDim.x * gridDim.y) {
        const int by = t / tiles_x, bx = t - by * tiles_x;
        const int c = bx * 64 + (threadIdx.x & 63);
        const int r = by * 4 + (threadIdx.x >> 6);
        if (r < N && c < M) {
            const size_t i = (size_t)r * M + c, ix = (size_t)r * (M - 1) + c;
            const bool hr = c < M - 1, hl = c > 0, hd = r < N - 1, hu = r > 0;
            // all nine loads first (clamped to the pixel itself where a neighbour is missing), then the arithmetic
            const double pc = p[i], pr = p[hr ? i + 1 : i], pl = p[hl ? i - 1 : i], pd = p[hd ? i + M : i], pu = p[hu ? i - M : i];
            const double wr = hr ? wwx[ix] : 0.0, wl = hl ? wwx[ix - 1] : 0.0, wd = hd ? wwy[i] : 0.0, wu = hu ? wwy[i - M] : 0.0;
            double v = 0.0;
            if (hr) v += wr * (pr - pc);
            if (hl) v -= wl * (pc - pl);
            if (hd) v += wd * (pd - pc);
            if (hu) v -= wu * (pc - pu);
            q[i] = v;
            pq = fma(pc, v, pq);
        }
    }
    const double s = block_sum(pq, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
    finish_reduction(&sc->ticket[kTicketAlpha], partial, gridDim.x * gridDim.y, red, [sc](double t) {
        sc->pqp = t;
        sc->alpha = sc->rz / t;                                 // phase_unwrap.py:201
    });
}

// phi += alpha p ; r -= alpha q ; partial sums of r^2
__global__ void __launch_bounds__(256) k_uw_update_xr(double* __restrict__ phi, double* __restrict__ r,
                                                      const double* __restrict__ p, const double* __restrict__ q,
                                                      double* __restrict__ partial, size_t n, UwScalars* sc) {
    if (sc->done) return;
    __shared__ double red[32];
    const double alpha = sc->alpha;
    double rs = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        phi[i] = fma(alpha, p[i], phi[i]);
        const double v = fma(-alpha, q[i], r[i]);
        r[i] = v;
        rs = fma(v, v, rs);
    }
    const double s = block_sum(rs, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
    finish_reduction(&sc->ticket[kTicketStop], partial, gridDim.x, red, [sc](double t) {
        sc->rsq = t;
        // :206  k >= kmax or |r| < 1e-9 |r0| ; and the loop condition `not all(r == 0)`
        if (sc->k >= sc->kmax || sqrt(t) < 1e-9 * sqrt(sc->r0sq) || t == 0.0) sc->done = 1;
    });
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static bool is_pow2(int n) { return n >= 2 && (n & (n - 1)) == 0; }

struct AxisTables {
    double2 *tw, *mk;
    double* ct;
    AxisPlan bs;          // bs.L != 0: any-length axis handled by the Bluestein transform
    int n, pow2;
};

constexpr int kMaxBluesteinAxis = kMaxFftLen / 2;    // L = pow2 >= 2n - 1 must fit shared memory

struct UwPlan {
    int N, M;
    double *r, *z, *t, *p, *q, *wwx, *wwy, *partial;
    AxisTables axN, axM;
    double *cosI, *cosJ;      // cos(pi I / M), I < N ; cos(pi J / N), J < M   (phase_unwrap.py:109, swapped on purpose)
    UwScalars* sc;
    int npart;
    CUtensorMap tmap_z;       // z as a 2-D tensor (M, N) for the pipelined column stage; valid if cols_pipe
    int cols_pipe;
    int fuse_p;               // the last inverse row pass also updates the search direction p (needs cols_pipe: beta is known)
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*PFN_uwEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_uwEncodeTiled uw_load_encode_tiled() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return reinterpret_cast<PFN_uwEncodeTiled>(fn);
}

static size_t carve_unwrap(UwPlan& u, void* ws, size_t ws_bytes, int N, int M) {
    Arena a(ws, ws_bytes);
    const size_t nm = (size_t)N * M;
    u.N = N; u.M = M;
    u.r = a.take<double>(nm); u.z = a.take<double>(nm); u.t = a.take<double>(nm);
    u.p = a.take<double>(nm); u.q = a.take<double>(nm);
    u.wwx = a.take<double>(nm); u.wwy = a.take<double>(nm);
    // per-CTA partial sums: the stencil kernels use one CTA per kUwRows x 64 pixels, the row kernels one value per row
    const size_t tiles = (size_t)ceil_div(M, 64) * ceil_div(N, kUwRows);
    u.npart = (int)(tiles > 16384 ? tiles : 16384);
    u.partial = a.take<double>(u.npart);
    for (AxisTables* ax : {&u.axN, &u.axM}) {
        const int n = ax == &u.axN ? N : M;
        ax->n = n; ax->pow2 = is_pow2(n) && n <= kMaxFftLen;
        ax->tw = a.take<double2>(n + 1);
        ax->mk = a.take<double2>(n);
        ax->bs.P = n; ax->bs.L = 0; ax->bs.chirp = ax->bs.bhat = ax->bs.tw = nullptr;
        if (!ax->pow2 && n <= kMaxBluesteinAxis) {
            ax->bs.L = bs_pow2_at_least(2 * n - 1);
            ax->bs.chirp = a.take<double2>(n);
            ax->bs.bhat = a.take<double2>(ax->bs.L);
            ax->bs.tw = a.take<double2>(ax->bs.L);
        }
        ax->ct = a.take<double>(ax->pow2 || ax->bs.L ? 1 : 4 * (size_t)n);
    }
    u.cosI = a.take<double>(N);
    u.cosJ = a.take<double>(M);
    u.sc = a.take<UwScalars>(1);
    u.cols_pipe = 0;
    u.fuse_p = 0;
    return a.off;
}

static bool g_dct_pipe = true;      // pipelined row / column kernels (gpa_set_dct_pipeline)

// describe u.z for the TMA boxes of k_poisson_cols_pipe (strips of 4 columns x 256 rows); leaves cols_pipe = 0 when the
// shape does not qualify (the one-CTA-per-strip kernel or the transposing path then runs)
static void plan_cols_pipe(UwPlan& u) {
    u.cols_pipe = 0;
    const int N = u.N, M = u.M;
    if (!g_dct_pipe || !u.axN.pow2 || N < 256 || N > 4096 || M % 2 != 0 || M < 4) return;
    static PFN_uwEncodeTiled encode = uw_load_encode_tiled();
    if (encode == nullptr) return;
    std::memset(&u.tmap_z, 0, sizeof(u.tmap_z));
    const cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)M * sizeof(double)};
    const cuuint32_t box[2] = {4, 256};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(&u.tmap_z, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)u.z, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    u.cols_pipe = cr == CUDA_SUCCESS;
    u.fuse_p = u.cols_pipe && u.axM.pow2 && M >= 256 && M <= 4096;       // the row pass along axis 1 is a pipelined kernel too
}

template <int LOGN>
static int launch_cols_pipe(const UwPlan& u, cudaStream_t st) {
    constexpr int n = 1 << LOGN, threads = n / 4;
    const size_t smem = dp_cols_smem_bytes(n);
    auto kern = k_poisson_cols_pipe<LOGN>;
    GPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static int per_sm = 0, sms = 0;
    if (!per_sm) {
        int dev = 0, v = 0;
        GPA_CHECK_CUDA(cudaGetDevice(&dev));
        GPA_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        GPA_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, threads, smem));
        per_sm = v > 0 ? v : 1;
    }
    ColPipeArgs c;
    c.N = u.N; c.M = u.M; c.tw = u.axN.tw; c.mk = u.axN.mk; c.cos_k = u.cosI; c.cos_c = u.cosJ; c.sc = u.sc;
    c.partial = u.partial;
    const int units = (u.M + 3) / 4, slots = sms * per_sm;
    kern<<<units < slots ? units : slots, threads, smem, st>>>(c, u.tmap_z);
    return GPA_OK;
}

template <int INVERSE, int LOGN, int MODE = 0>
static int launch_rows_pipe(const DctArgs& a, cudaStream_t st) {
    if (INVERSE && MODE == 0 && a.dot_with) return launch_rows_pipe<INVERSE, LOGN, 1>(a, st);
    if (INVERSE && MODE == 0 && a.p_update) return launch_rows_pipe<INVERSE, LOGN, 2>(a, st);
    auto kern = INVERSE ? k_idct2_rows_pipe<LOGN, MODE> : k_dct2_rows_pipe<LOGN>;
    constexpr int n = 1 << LOGN, threads = n / 8;
    const size_t smem = dp_rows_smem_bytes(n);
    GPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static int per_sm[2] = {0, 0}, sms = 0;        // occupancy of this instantiation (same on every B200)
    if (!per_sm[INVERSE]) {
        int dev = 0, v = 0;
        GPA_CHECK_CUDA(cudaGetDevice(&dev));
        GPA_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        GPA_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, threads, smem));
        per_sm[INVERSE] = v > 0 ? v : 1;
    }
    const int pairs = (a.rows + 1) / 2;
    const int slots = sms * per_sm[INVERSE];
    kern<<<pairs < slots ? pairs : slots, threads, smem, st>>>(a);
    return GPA_OK;
}

template <int INVERSE>
static int launch_rows(const AxisTables& ax, DctArgs a, cudaStream_t st) {
    a.tw = ax.tw; a.mk = ax.mk; a.ct = ax.ct; a.n = ax.n; a.bs = ax.bs;
    if (g_dct_pipe && ax.pow2 && !ax.bs.L && ax.n >= 256 && ax.n <= 4096 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0) {
        switch (ax.n) {
            case 256: return launch_rows_pipe<INVERSE, 8>(a, st);
            case 512: return launch_rows_pipe<INVERSE, 9>(a, st);
            case 1024: return launch_rows_pipe<INVERSE, 10>(a, st);
            case 2048: return launch_rows_pipe<INVERSE, 11>(a, st);
            default: return launch_rows_pipe<INVERSE, 12>(a, st);
        }
    }
    if (ax.pow2 || ax.bs.L) {
        // one radix-8 butterfly per thread, two for the longest transforms (at most 512 threads)
        const int len = ax.bs.L ? ax.bs.L : ax.n;
        int threads, per;
        fft_launch_shape(len, threads, per);
        const size_t smem = (size_t)len * sizeof(double2);
        const int ctas = (a.rows + 1) / 2;           // two rows per FFT
        auto go = [&](auto kern) -> int {
            GPA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
            kern<<<ctas, threads, smem, st>>>(a);
            return GPA_OK;
        };
        if (per <= 1) return INVERSE ? go(k_idct2_rows_pow2<1>) : go(k_dct2_rows_pow2<1>);
        return INVERSE ? go(k_idct2_rows_pow2<2>) : go(k_dct2_rows_pow2<2>);
    }
    GPA_REQUIRE((size_t)ax.n * sizeof(double) <= 200 * 1024, "axis of length %d is too long for the direct DCT", ax.n);
    const size_t smem = (size_t)ax.n * sizeof(double);
    GPA_CHECK_CUDA(cudaFuncSetAttribute(k_dct2_rows_direct<INVERSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_dct2_rows_direct<INVERSE><<<a.rows, 256, smem, st>>>(a);
    return GPA_OK;
}

static void transpose(const double* in, double* out, int rows, int cols, const UwScalars* sc, cudaStream_t st) {
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
    k_transpose<<<grid, dim3(32, 8), 0, st>>>(in, out, rows, cols, sc);
}

// z = idctn(dctn(r) / scale); partial[0..N) = row sums of r * z           (phase_unwrap.py:95-103)
static int poisson_solve(const UwPlan& u, cudaStream_t st) {
    const int N = u.N, M = u.M;
    int rc;
    DctArgs a;
    std::memset(&a, 0, sizeof(a));
    a.sc = u.sc; a.dimN = N; a.dimM = M;
    KernelTimer timer("uw_poisson_solve", st);
    a.in = u.r; a.out = u.z; a.rows = N;                                   // rows along axis 1
    if ((rc = launch_rows<0>(u.axM, a, st))) return rc;
    if (u.cols_pipe) {                    // pipelined fused column stage (TMA strips of 4 columns)
        switch (N) {
            case 256: rc = launch_cols_pipe<8>(u, st); break;
            case 512: rc = launch_cols_pipe<9>(u, st); break;
            case 1024: rc = launch_cols_pipe<10>(u, st); break;
            case 2048: rc = launch_cols_pipe<11>(u, st); break;
            default: rc = launch_cols_pipe<12>(u, st); break;
        }
        if (rc) return rc;
        a.in = u.z; a.out = u.t; a.rows = N;                               // <r, z> came out of the column stage
        if (u.fuse_p) a.p_update = u.p;                                    // p = z_k + beta p instead of t = z_k
        return launch_rows<1>(u.axM, a, st);                               // t = z_k
    }
    if (u.axN.pow2 && M % 2 == 0) {       // fused column stage: z <- idct_0(dct_0(z) / scale) in one pass over the array
        ColArgs c;
        c.z = u.z; c.N = N; c.M = M; c.tw = u.axN.tw; c.mk = u.axN.mk; c.cos_k = u.cosI; c.cos_c = u.cosJ; c.sc = u.sc;
        const int maxb = N / 8 > 512 ? 2 : 1;
        c.tpf = N / (8 * maxb) < 32 ? 32 : N / (8 * maxb);
        int cw = 8;
        while (cw > 2 && ((size_t)(cw / 2) * (N + 1) * sizeof(double2) > 200 * 1024 || (cw / 2) * c.tpf > (maxb == 1 ? 1024 : 512))) cw /= 2;
        c.CW = cw;
        const size_t smem = (size_t)(cw / 2) * (N + 1) * sizeof(double2);
        const int threads = (cw / 2) * c.tpf;
        if (maxb == 1) {
            GPA_CHECK_CUDA(cudaFuncSetAttribute(k_poisson_cols<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
            k_poisson_cols<1><<<ceil_div(M, cw), threads, smem, st>>>(c);
        } else {
            GPA_CHECK_CUDA(cudaFuncSetAttribute(k_poisson_cols<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
            k_poisson_cols<2><<<ceil_div(M, cw), threads, smem, st>>>(c);
        }
        a.in = u.z; a.out = u.t; a.rows = N; a.dot_with = u.r; a.partial = u.partial;
        return launch_rows<1>(u.axM, a, st);                               // t = z_k, partial = <r, z> rows
    }
    transpose(u.z, u.t, N, M, u.sc, st);                                   // t: (M, N)
    a.in = u.t; a.out = u.z; a.rows = M; a.fuse_scale = 1;                 // rows along axis 0, then / scale
    a.cos_col = u.cosI; a.cos_row = u.cosJ;
    if ((rc = launch_rows<0>(u.axN, a, st))) return rc;
    a.fuse_scale = 0;
    a.in = u.z; a.out = u.t; a.rows = M;                                   // inverse along axis 0
    if ((rc = launch_rows<1>(u.axN, a, st))) return rc;
    transpose(u.t, u.z, M, N, u.sc, st);                                   // z: (N, M)
    a.in = u.z; a.out = u.t; a.rows = N; a.dot_with = u.r; a.partial = u.partial;
    if ((rc = launch_rows<1>(u.axM, a, st))) return rc;                    // t = z_k, partial = <r, z> rows
    return GPA_OK;
}

// twiddle / Makhoul / Bluestein tables of both axes and the cosines of the Poisson scale (a few us; the workspace is
// the caller's and may have been reused by other entry points since the last call, so they are rebuilt per call)
static int build_uw_tables(const UwPlan& u, cudaStream_t st) {
    for (const AxisTables* ax : {&u.axN, &u.axM}) {
        const int mode = ax->pow2 ? 2 : (ax->bs.L ? 1 : 0);
        const int cnt = mode ? ax->n : 4 * ax->n;
        k_uw_tables<<<ceil_div(cnt, 256), 256, 0, st>>>(ax->tw, ax->mk, ax->ct, ax->n, mode);
        if (ax->bs.L) {       // chirp, FFT_L twiddles and the transformed chirp of the Bluestein convolution
            k_bs_tables<<<ceil_div(ax->bs.L, 256), 256, 0, st>>>(ax->bs.chirp, ax->bs.tw, ax->bs.P, ax->bs.L);
            int threads, per;
            fft_launch_shape(ax->bs.L, threads, per);
            const size_t smem = (size_t)ax->bs.L * sizeof(double2);
            if (per <= 1) {
                GPA_CHECK_CUDA(cudaFuncSetAttribute(k_bs_prep<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
                k_bs_prep<1><<<1, threads, smem, st>>>(ax->bs);
            } else {
                GPA_CHECK_CUDA(cudaFuncSetAttribute(k_bs_prep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
                k_bs_prep<2><<<1, threads, smem, st>>>(ax->bs);
            }
        }
    }
    k_uw_cos_table<<<ceil_div(u.N, 256), 256, 0, st>>>(u.cosI, u.N, u.M);
    k_uw_cos_table<<<ceil_div(u.M, 256), 256, 0, st>>>(u.cosJ, u.M, u.N);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

// scale[I][J] = 2 (cos(pi I / M) + cos(pi J / N) - 2), [0][0] = 1        (phase_unwrap.py:106-115, N / M swapped as there)
__global__ void k_uw_poisson_scale(double* __restrict__ scale, int N, int M) {
    const size_t n = (size_t)N * M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        scale[i] = poisson_scale((int)(i / M), (int)(i % M), N, M);
}

__global__ void k_uw_divide(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = a[i] / b[i];
}

}  // namespace gpa

using namespace gpa;

// scipy.fft.dctn / idctn (type 2, norm=None) of an (N, M) float64 array: the transform pair of solvePoisson
// (phase_unwrap.py:81-103), exposed so that the helper functions of the reference have device mirrors.
extern "C" int gpa_dctn(const double* in, int N, int M, int inverse, double* out, void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(in && out && ws && N >= 2 && M >= 2, "bad argument");
    UwPlan u;
    const size_t need = carve_unwrap(u, ws, ws_bytes, N, M);
    if (need > ws_bytes) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = build_uw_tables(u, st);
    if (rc) return rc;
    GPA_CHECK_CUDA(cudaMemsetAsync(u.sc, 0, sizeof(UwScalars), st));
    DctArgs a;
    std::memset(&a, 0, sizeof(a));
    a.sc = u.sc; a.dimN = N; a.dimM = M;
    a.in = in; a.out = u.z; a.rows = N;                                     // along axis 1
    if ((rc = inverse ? launch_rows<1>(u.axM, a, st) : launch_rows<0>(u.axM, a, st))) return rc;
    transpose(u.z, u.t, N, M, u.sc, st);
    a.in = u.t; a.out = u.z; a.rows = M;                                    // along axis 0
    if ((rc = inverse ? launch_rows<1>(u.axN, a, st) : launch_rows<0>(u.axN, a, st))) return rc;
    transpose(u.z, out, M, N, u.sc, st);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

/* K2 row transforms: pipelined kernels (bulk-copy prefetch, persistent CTAs; default) or the one-CTA-per-row-pair kernels. */
extern "C" int gpa_set_dct_pipeline(int on) {
    g_dct_pipe = on != 0;
    return GPA_OK;
}

extern "C" int gpa_poisson_scale(int N, int M, double* scale, void* stream) {
    GPA_REQUIRE(scale && N >= 1 && M >= 1, "bad argument");
    size_t blocks = ((size_t)N * M + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_uw_poisson_scale<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(scale, N, M);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_divide_f64(const double* a, const double* b, double* out, size_t n, void* stream) {
    GPA_REQUIRE(a && b && out, "null pointer argument");
    if (n == 0) return GPA_OK;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_uw_divide<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, n);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

// q = (A^T)(W^T)(W)(A) p for edge weights wwx (N, M-1), wwy (N-1, M)             (applyQ, phase_unwrap.py:118-132)
extern "C" int gpa_apply_q(const double* p, const double* wwx, const double* wwy, int N, int M, double* q, void* ws,
                           size_t ws_bytes, void* stream) {
    GPA_REQUIRE(p && wwx && wwy && q && ws && N >= 2 && M >= 2, "bad argument");
    dim3 g2(ceil_div(M, 64), ceil_div(N, kUwRows));
    const size_t need = 512 + (size_t)g2.x * g2.y * sizeof(double);
    if (need > ws_bytes) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    UwScalars* sc = static_cast<UwScalars*>(ws);
    double* partial = reinterpret_cast<double*>(static_cast<char*>(ws) + 512);
    GPA_CHECK_CUDA(cudaMemsetAsync(sc, 0, sizeof(UwScalars), st));
    k_uw_apply_q<<<g2, 256, 0, st>>>(p, wwx, wwy, q, partial, N, M, sc);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_unwrap_workspace_bytes(int N, int M, size_t* bytes) {
    GPA_REQUIRE(bytes && N >= 2 && M >= 2, "frame must be at least 2x2");
    UwPlan u;
    *bytes = carve_unwrap(u, nullptr, 0, N, M) + 512;
    return GPA_OK;
}

extern "C" int gpa_unwrap_pcg(const double* psi, const double* dx, const double* dy, const double* weight,
                              int N, int M, int kmax, double* phi, int* iterations /*host, may be null*/,
                              void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(phi && ws, "null pointer argument");
    GPA_REQUIRE((psi != nullptr) != (dx != nullptr || dy != nullptr), "pass either psi or (dx, dy)");
    GPA_REQUIRE(psi || (dx && dy), "both dx and dy are needed");
    GPA_REQUIRE(N >= 2 && M >= 2, "frame must be at least 2x2 (got %dx%d)", N, M);
    UwPlan u;
    const size_t need = carve_unwrap(u, ws, ws_bytes, N, M);
    if (need > ws_bytes) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t nm = (size_t)N * M;
    plan_cols_pipe(u);
    {
        int rc = build_uw_tables(u, st);
        if (rc) return rc;
    }
    dim3 g2(ceil_div(M, 64), ceil_div(N, kUwRows));
    const int n2 = g2.x * g2.y;
    GPA_REQUIRE(n2 <= u.npart && N <= u.npart && M <= u.npart, "frame too large for the reduction scratch");
    // streaming / stencil kernels: grid-stride loops under a grid that fills the machine exactly once (8 CTAs of 256 threads per SM)
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        GPA_CHECK_CUDA(cudaGetDevice(&dev));
        GPA_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int wave = sms * 8;
    const int g1 = (int)((nm + 255) / 256 < (size_t)wave ? (nm + 255) / 256 : wave);
    const int n_sub = ceil_div(M, 64) * ceil_div(N, 4);
    const int gq = n_sub < wave ? n_sub : wave;
    {
        SetupArgs s;
        s.psi = psi; s.dx = dx; s.dy = dy; s.weight = weight;
        s.wwx = u.wwx; s.wwy = u.wwy; s.r = u.r; s.phi = phi; s.partial = u.partial; s.N = N; s.M = M;
        s.sc = u.sc; s.kmax = kmax;
        KernelTimer timer("uw_setup", st);
        GPA_CHECK_CUDA(cudaMemsetAsync(u.sc, 0, sizeof(UwScalars), st));     // done = 0, tickets = 0
        if (psi) {
            if (weight) k_uw_setup<true, true><<<gq, 256, 0, st>>>(s);
            else k_uw_setup<true, false><<<gq, 256, 0, st>>>(s);
        } else {
            if (weight) k_uw_setup<false, true><<<gq, 256, 0, st>>>(s);
            else k_uw_setup<false, false><<<gq, 256, 0, st>>>(s);
        }
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    // The reference always runs at least one iteration (k is tested after the update), so kmax <= 1
    // behaves like kmax = 1.
    const int iters = kmax < 1 ? 1 : kmax;
    int done_host = 0;
    for (int k = 0; k < iters; ++k) {
        int rc = poisson_solve(u, st);                       // t = z
        if (rc) return rc;
        {
            KernelTimer timer("uw_vector_ops", st);
            // (fusing p = z + beta p into the stencil kernel was measured SLOWER on B200: 0.115 vs 0.097 ms per iteration at
            // 2048^2 — the stencil is LSU- / latency-bound, not DRAM-bound, and the fused form doubles its loads)
            if (!u.fuse_p) k_uw_update_p<<<g1, 256, 0, st>>>(u.t, u.p, nm, u.sc);
            k_uw_apply_q<<<gq, 256, 0, st>>>(u.p, u.wwx, u.wwy, u.q, u.partial, N, M, u.sc);
            k_uw_update_xr<<<g1, 256, 0, st>>>(phi, u.r, u.p, u.q, u.partial, nm, u.sc);
        }
        GPA_CHECK_CUDA(cudaGetLastError());
        // stop enqueueing once converged — only for long runs: a host synchronisation stalls the caller's pipeline (the
        // adaptive chain calls this with kmax = 10 twice per frame), and converged iterations exit at once anyway
        if ((k & 7) == 7 && k + 1 < iters && iters > 16) {
            GPA_CHECK_CUDA(cudaMemcpyAsync(&done_host, &u.sc->done, sizeof(int), cudaMemcpyDeviceToHost, st));
            GPA_CHECK_CUDA(cudaStreamSynchronize(st));
            if (done_host) break;
        }
    }
    if (iterations) {
        GPA_CHECK_CUDA(cudaMemcpyAsync(iterations, &u.sc->k, sizeof(int), cudaMemcpyDeviceToHost, st));
        GPA_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    return GPA_OK;
}
