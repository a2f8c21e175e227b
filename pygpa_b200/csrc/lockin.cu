// K1 — spatial lock-in and adaptive (windowed-Fourier-ridge) sweep for sm_100a.
//
// Reference semantics: pyGPA/geometric_phase_analysis.py:48-76 (optGPA), 763-813
// (wfr2_grad_opt), pyGPA/cuGPA.py:41-87.  See include/gpa_b200.h for the contract and
// DESIGN.md for the derivation.  Structure:
//
//   k_build_phasors   fp64 range-reduced carrier tables  e^{2 pi i w x'}  (tiny)
//   k_pass1           plane_iy = G_y * (img . e^{2 pi i wy y})      one launch per plane chunk
//   k_pass2<ARGMAX>   for every candidate row ix: sf = G_x * (e^{2 pi i wx x} . plane_iy),
//                     running arg-max of |sf|^2 in registers, one 64-bit atomicMax per pixel
//   k_pass2<STORE>    same filter, writes sf (fixed-reference lock-in)
//   k_finalize        re-evaluates the winner at the pixel and its 4 neighbours -> lock-in
//                     re-referenced to kref, phase gradient, k-index
//
// Both filter kernels share one register-blocked FIR core: each thread owns P consecutive
// outputs along the filter axis, keeps a P-deep rotating window of complex samples in
// registers and issues P packed FFMA2 (fma.rn.f32x2: real tap x complex sample) per tap;
// taps come from the kernel-parameter constant bank through uniform registers.
#include "common.cuh"

namespace gpa {

constexpr int kMaxTaps = 446;      // 2R+1 <= kMaxTaps  (sigma <= 49 at 4.5 sigma); param space budget
constexpr int kP = 16;             // outputs per thread along the filter axis
constexpr int kWarps = 8;          // warps per CTA
constexpr int kTile = kP * kWarps; // outputs per CTA along the filter axis (128)
constexpr int kLanes = 32;         // outputs per CTA across the filter axis

struct TapTable {
    float2 g[kMaxTaps + 2];        // (tap, tap): packed operand of FFMA2; zero-filled past 2R+1
};

struct WList {
    double w[224];
};

// ---------------------------------------------------------------------------------------------
// carrier tables
// ---------------------------------------------------------------------------------------------
// table[i][r] = exp(2 pi i w[i] * ((r - shift) mod period)),  r in [0, len)
__global__ void k_build_phasors(float2* __restrict__ table, double* __restrict__ w_out,
                                const __grid_constant__ WList wl, int n_w, int len, int shift,
                                int period) {
    const int i = blockIdx.y;
    if (i >= n_w) return;
    const double w = wl.w[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) w_out[i] = w;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < len; r += gridDim.x * blockDim.x) {
        int xs = (r - shift) % period;
        if (xs < 0) xs += period;
        table[(size_t)i * len + r] = phasor_turns(w * (double)xs);
    }
}

// ---------------------------------------------------------------------------------------------
// asynchronous global -> shared copies (LDGSTS): tile fills with every row in flight at once
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// register-blocked FIR core
// ---------------------------------------------------------------------------------------------
// acc[p] = sum_{d<T} g[d] * sample(p + d),  p < P.   `load(j)` returns sample j.  Taps and samples
// are fetched kAhead steps before they are consumed (software pipeline in registers), so load(j)
// is called for j up to T + P - 1 + kAhead and taps.g is read up to index T - 1 + kAhead: both
// must be readable (the callers pad their tiles / the tap table is zero-filled).
constexpr int kAhead = 2;

// acc[p] += sum_{d<cnt} taps.g[off + d] * sample(p + d)
template <int P, typename Load>
__device__ __forceinline__ void fir_phase(float2 (&acc)[P], const TapTable& taps, int off, int cnt, Load load) {
    static_assert(P % kAhead == 0, "P must be a multiple of the prefetch depth");
    float2 win[P], gq[kAhead], sq[kAhead];
#pragma unroll
    for (int p = 0; p < P; ++p) win[p] = load(p);
#pragma unroll
    for (int a = 0; a < kAhead; ++a) {
        gq[a] = taps.g[off + a];
        sq[a] = load(P + a);
    }
    int d0 = 0;
    for (; d0 + P <= cnt; d0 += P) {
#pragma unroll
        for (int u = 0; u < P; ++u) {
            const float2 g = gq[u % kAhead];
            const float2 s = sq[u % kAhead];
            gq[u % kAhead] = taps.g[off + d0 + u + kAhead];
            sq[u % kAhead] = load(d0 + u + P + kAhead);
#pragma unroll
            for (int p = 0; p < P; ++p) acc[p] = __ffma2_rn(g, win[(u + p) % P], acc[p]);
            win[u] = s;
        }
    }
    const int rem = cnt - d0;
#pragma unroll
    for (int u = 0; u < P - 1; ++u) {
        if (u < rem) {  // warp-uniform
            const float2 g = gq[u % kAhead];
            const float2 s = sq[u % kAhead];
            gq[u % kAhead] = taps.g[off + d0 + u + kAhead];
            sq[u % kAhead] = load(d0 + u + P + kAhead);
#pragma unroll
            for (int p = 0; p < P; ++p) acc[p] = __ffma2_rn(g, win[(u + p) % P], acc[p]);
            win[u] = s;
        }
    }
}

template <int P, typename Load>
__device__ __forceinline__ void fir_block(float2 (&acc)[P], const TapTable& taps, int T, Load load) {
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = make_float2(0.f, 0.f);
    fir_phase<P>(acc, taps, 0, T, load);
}

// ---------------------------------------------------------------------------------------------
// pass 1: filter along axis 1 (contiguous) of the demodulated real image
// ---------------------------------------------------------------------------------------------
struct Pass1Params {
    const float* img;     // (N, M)
    const float2* phy;    // [n_planes][M] carrier along axis 1
    float2* planes;       // [chunk][n_alloc][pitch]
    size_t plane_stride;  // elements
    int N, M, pitch, n_rows_filled /* N + 2Rx */, Rx, Ry, T /* 2Ry+1 */, plane0 /* global index of chunk plane 0 */;
};

// CTA: 32 padded rows (one per lane) x kTile output columns (warp w owns columns [w*P, w*P+P)).
// smem: demodulated samples s[j][lane], j in [0, kTile + T), row pitch 33 float2 (conflict-free
// reads across lanes; the transposing fill is 2-way conflicted, once per tile).
__global__ void __launch_bounds__(kWarps * 32, 2)
k_pass1(const Pass1Params prm, const __grid_constant__ TapTable taps) {
    extern __shared__ float2 smem[];
    constexpr int SP = 33;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32;            // first padded row of the tile
    const int y0 = blockIdx.y * kTile;         // first output column
    const int pl = blockIdx.z;                 // plane within the chunk
    const int T = prm.T, M = prm.M, N = prm.N;
    const int n_samp = kTile + T + kAhead; // T-1 halo + prefetch slack
    const float2* __restrict__ phy = prm.phy + (size_t)(prm.plane0 + pl) * M;

    int cbase = (y0 - prm.Ry) % M;
    if (cbase < 0) cbase += M;
    for (int rr = warp; rr < 32; rr += kWarps) {
        int xs = (r0 + rr - prm.Rx) % N;
        if (xs < 0) xs += N;
        const float* __restrict__ row = prm.img + (size_t)xs * M;
        for (int j = lane; j < n_samp; j += 32) {
            int c = cbase + j;
            if (c >= M) c %= M;
            const float v = __ldg(row + c);
            const float2 ph = __ldg(phy + c);
            smem[j * SP + rr] = make_float2(v * ph.x, v * ph.y);
        }
    }
    __syncthreads();

    const float2* col = smem + (warp * kP) * SP + lane;
    float2 acc[kP];
    fir_block<kP>(acc, taps, T, [&](int j) { return col[j * SP]; });

    const int r = r0 + lane;
    const int y = y0 + warp * kP;
    if (r < prm.n_rows_filled) {
        float2* out = prm.planes + (size_t)pl * prm.plane_stride + (size_t)r * prm.pitch + y;
#pragma unroll
        for (int p = 0; p < kP; p += 2) {
            if (y + p + 1 < prm.pitch) {
                *reinterpret_cast<float4*>(out + p) = make_float4(acc[p].x, acc[p].y, acc[p + 1].x, acc[p + 1].y);
            } else if (y + p < prm.pitch) {
                out[p] = acc[p];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2: filter along axis 0 with per-candidate demodulation; arg-max or store
// ---------------------------------------------------------------------------------------------
struct Pass2Params {
    const float2* planes;   // [chunk][n_alloc][pitch]
    size_t plane_stride;
    const float2* phx;      // [n_rows][n_alloc] carrier along axis 0, indexed by PADDED row
    unsigned long long* key;  // ARGMAX: (N, M)
    void* out;                // STORE:  (N, M) float2 or double2
    int out_f64;
    int N, M, pitch, n_alloc, T /* 2Rx+1 */;
    int plane0;             // global plane index of chunk plane 0
    int n_cand;             // candidates per plane (grid: n_rows, list: 1)
    int row_c, row_p;       // phasor row  = c*row_c + plane*row_p
    int idx_c, idx_p;       // flat index  = c*idx_c + plane*idx_p
};

enum { kArgmax = 0, kStore = 1 };

// CTA: kTile output rows (warp w owns rows [w*P, w*P+P)) x 32 columns (one per lane).
// smem: the plane tile [kTile + T][32] complex, loaded once and reused by every candidate.
template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 2)
k_pass2(const Pass2Params prm, const __grid_constant__ TapTable taps) {
    extern __shared__ float2 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = blockIdx.x * kLanes;
    const int x0 = blockIdx.y * kTile;
    const int pl = blockIdx.z;
    const int plane = prm.plane0 + pl;
    const int T = prm.T;
    const int n_samp = kTile + T + kAhead;

    {   // tile fill: rows are 256 B (16 copies of 16 B), all in flight; pitch/n_alloc padding keeps it in bounds
        const float2* __restrict__ src = prm.planes + (size_t)pl * prm.plane_stride + (size_t)x0 * prm.pitch + y0;
        for (int i = threadIdx.x; i < n_samp * (kLanes / 2); i += kWarps * 32) {
            const int j = i / (kLanes / 2), c = 2 * (i % (kLanes / 2));
            cp_async16(smem + j * kLanes + c, src + (size_t)j * prm.pitch + c);
        }
        cp_async_commit();
        cp_async_wait_all();
    }
    __syncthreads();

    const float2* col = smem + (warp * kP) * kLanes + lane;
    float best[kP];
    unsigned bidx[kP / 2];   // winning candidate per output, two 16-bit fields per register
#pragma unroll
    for (int p = 0; p < kP; ++p) best[p] = 0.f;
#pragma unroll
    for (int p = 0; p < kP / 2; ++p) bidx[p] = 0u;

    for (int c = 0; c < prm.n_cand; ++c) {
        const float2* __restrict__ ph = prm.phx + (size_t)(c * prm.row_c + plane * prm.row_p) * prm.n_alloc +
                                        x0 + warp * kP;
        float2 acc[kP];
        fir_block<kP>(acc, taps, T, [&](int j) { return cmul(col[j * kLanes], __ldg(ph + j)); });
        if (MODE == kArgmax) {
            const unsigned c2 = (unsigned)c * 0x10001u;
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const float a2 = fmaf(acc[p].x, acc[p].x, acc[p].y * acc[p].y);
                if (a2 > best[p]) {   // strict: the earlier candidate keeps exact ties
                    best[p] = a2;
                    const unsigned keep = (p & 1) ? 0x0000FFFFu : 0xFFFF0000u;
                    bidx[p / 2] = (bidx[p / 2] & keep) | (c2 & ~keep);
                }
            }
        } else {
            const int y = y0 + lane;
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const int x = x0 + warp * kP + p;
                if (x < prm.N && y < prm.M) {
                    if (prm.out_f64) static_cast<double2*>(prm.out)[(size_t)x * prm.M + y] = make_double2(acc[p].x, acc[p].y);
                    else static_cast<float2*>(prm.out)[(size_t)x * prm.M + y] = acc[p];
                }
            }
        }
    }

    if (MODE == kArgmax) {
        const int y = y0 + lane;
        if (y < prm.M) {
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const int x = x0 + warp * kP + p;
                if (x < prm.N && best[p] > 0.f) {
                    const unsigned cwin = (bidx[p / 2] >> ((p & 1) * 16)) & 0xFFFFu;
                    const unsigned idx = cwin * (unsigned)prm.idx_c + (unsigned)(plane * prm.idx_p);
                    const unsigned long long k =
                        ((unsigned long long)__float_as_uint(best[p]) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
                    atomicMax(prm.key + (size_t)x * prm.M + y, k);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2, sequential acceptance (wfr4): geometric_phase_analysis.py:839-862
// ---------------------------------------------------------------------------------------------
// The candidates form an ORDERED list and a pixel accepts candidate i only if |sf_i| is strictly
// larger than what it holds AND k_i lies within 2 sqrt(2) dk of the k it currently holds.  The rule
// is order dependent per pixel, so one CTA owns a pixel tile and walks the planes of the chunk in
// list order; the (amplitude, held index) state lives in registers and is carried across chunks in
// `key` (same packing as the arg-max sweeps, but plain loads/stores: no other CTA touches the tile).
// The neighbourhood test is a host-built K x K byte table (the reference's float64 expression,
// evaluated once per PAIR of list entries instead of once per pixel and candidate).
__global__ void k_fill_u64(unsigned long long* __restrict__ dst, unsigned long long v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

struct SeqParams {
    const float2* planes;
    size_t plane_stride;
    const float2* phx;
    const unsigned char* allowed;   // [K][K]: allowed[held * K + candidate]
    unsigned long long* key;
    int N, M, pitch, n_alloc, T, plane0, count, K;
};

__global__ void __launch_bounds__(kWarps * 32, 1)
k_pass2_seq(const SeqParams prm, const __grid_constant__ TapTable taps) {
    extern __shared__ float2 smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = blockIdx.x * kLanes;
    const int x0 = blockIdx.y * kTile;
    const int T = prm.T;
    const int n_samp = kTile + T + kAhead;
    const int y = y0 + lane;
    float best[kP];
    int held[kP];
#pragma unroll
    for (int p = 0; p < kP; ++p) {
        const int x = x0 + warp * kP + p;
        best[p] = 0.f;
        held[p] = 0;
        if (x < prm.N && y < prm.M) {
            const unsigned long long k = prm.key[(size_t)x * prm.M + y];
            best[p] = __uint_as_float((unsigned)(k >> 32));
            held[p] = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
        }
    }
    const float2* col = smem + (warp * kP) * kLanes + lane;
    for (int pl = 0; pl < prm.count; ++pl) {
        __syncthreads();      // the previous plane's tile is no longer read
        const float2* __restrict__ src = prm.planes + (size_t)pl * prm.plane_stride + (size_t)x0 * prm.pitch + y0;
        for (int i = threadIdx.x; i < n_samp * (kLanes / 2); i += kWarps * 32) {
            const int j = i / (kLanes / 2), c = 2 * (i % (kLanes / 2));
            cp_async16(smem + j * kLanes + c, src + (size_t)j * prm.pitch + c);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        const int cand = prm.plane0 + pl;
        const float2* __restrict__ ph = prm.phx + (size_t)cand * prm.n_alloc + x0 + warp * kP;
        float2 acc[kP];
        fir_block<kP>(acc, taps, T, [&](int j) { return cmul(col[j * kLanes], __ldg(ph + j)); });
        const unsigned char* __restrict__ ok = prm.allowed + cand;
#pragma unroll
        for (int p = 0; p < kP; ++p) {
            const float a2 = fmaf(acc[p].x, acc[p].x, acc[p].y * acc[p].y);
            if (a2 > best[p] && __ldg(ok + (size_t)held[p] * prm.K)) {
                best[p] = a2;
                held[p] = cand;
            }
        }
    }
    if (y < prm.M) {
#pragma unroll
        for (int p = 0; p < kP; ++p) {
            const int x = x0 + warp * kP + p;
            if (x < prm.N)
                prm.key[(size_t)x * prm.M + y] = ((unsigned long long)__float_as_uint(best[p]) << 32) |
                                                 (unsigned long long)(0xFFFFFFFFu - (unsigned)held[p]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// multirate arg-max sweep
// ---------------------------------------------------------------------------------------------
// The Gaussian factorises, G_sigma = G_a * G_b with sigma_a^2 + sigma_b^2 = sigma^2, and after G_a
// the signal is band-limited, so it can be decimated by S per axis and G_b applied as an S-fold
// interpolator (aliasing ~ exp(-2 pi^2 sigma_a^2 sigma_b^2 / (sigma^2 S^2)) < 1e-7 for the strides the
// host picks).  Per candidate the cost falls from T = 2R+1 full-rate taps to ~T_a/S^2 + W/S + W with
// W = 11 coarse taps:
//   k_mr_pass1  P1[wy](x', my)    = sum_y' G_a(S my - y') img(x', y') e^{2 pi i wy y'}       (per plane)
//   k_mr_pass2  P2[wx,wy](mx, my) = sum_x' G_a(S mx - x') e^{2 pi i wx x'} P1[wy](x', my)    (per candidate)
//   k_mr_interp sf(x, y) = sum_my S G_b(y - S my) sum_mx S G_b(x - S mx) P2(mx, my), |sf|^2, arg-max
// Only the arg-max DECISION uses these amplitudes; k_finalize recomputes the winner with the direct
// form, so lock-in, gradient and w keep the direct path's accuracy.
constexpr int kMrW = 12;      // coarse taps per output (11 used, padded to 12)
constexpr int kMrHL = 5;      // coarse samples to the left of an output's own cell
constexpr int kMrTX = 64;     // k_mr_interp tile: rows
constexpr int kMrTY = 128;    //                   columns
constexpr int kPmB = 8;       // bound blocks for the pruning: kPmB x kPmB coarse cells

struct MrPass1Params {
    const float* img;
    const float2* phy;
    float2* p1;            // [chunk][n_alloc][pitch_d]
    size_t plane_stride;
    int N, M, Md, pitch_d, n_rows_filled, Rax, Ray, J /* taps per phase */, plane0, pstep, count, planes_per_cta;
};

// decimating version of k_pass1: lane = padded row, warp w owns decimated outputs [w*P, w*P+P).
// The image tile is plane independent, so it is staged ONCE per CTA as raw float samples
// (transposed, pitch 33) and the CTA loops over `planes_per_cta` planes; per plane only the
// carrier of the tile columns is staged (double buffered) and applied on the fly (2 FMUL/sample).
template <int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
k_mr_pass1(const MrPass1Params prm, const __grid_constant__ TapTable taps) {
    extern __shared__ float smem_f[];
    constexpr int SP = 33;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32;
    const int m0 = blockIdx.y * (WARPS * kP);          // first decimated output column
    const int M = prm.M, N = prm.N, J = prm.J;
    const int n_samp = S * (WARPS * kP + J + kAhead + 1);
    float* const tile = smem_f;                                              // [n_samp][SP] float
    float2* const car = reinterpret_cast<float2*>(smem_f + (size_t)n_samp * SP + (n_samp * SP & 1));   // [2][n_samp]
    int cbase = (S * m0 - prm.Ray) % M;
    if (cbase < 0) cbase += M;
    for (int rr = warp; rr < 32; rr += WARPS) {
        int xs = (r0 + rr - prm.Rax) % N;
        if (xs < 0) xs += N;
        const float* __restrict__ row = prm.img + (size_t)xs * M;
        for (int j = lane; j < n_samp; j += 32) {
            int c = cbase + j;
            if (c >= M) c %= M;
            cp_async4(tile + j * SP + rr, row + c);       // transposing fill, every copy in flight at once
        }
    }
    cp_async_commit();
    const int pl0 = blockIdx.z * prm.planes_per_cta;
    const int pl1 = min(pl0 + prm.planes_per_cta, prm.count);
    auto stage_carrier = [&](int pl, int slot) {
        const float2* __restrict__ phy = prm.phy + (size_t)(prm.plane0 + pl * prm.pstep) * M;
        float2* dst = car + slot * n_samp;
        for (int j = threadIdx.x; j < n_samp; j += WARPS * 32) {
            int c = cbase + j;
            if (c >= M) c %= M;
            dst[j] = __ldg(phy + c);
        }
    };
    stage_carrier(pl0, 0);
    cp_async_wait_all();
    __syncthreads();
    const float* col = tile + (S * warp * kP) * SP + lane;
    const int r = r0 + lane;
    const int m = m0 + warp * kP;
    int slot = 0;
    for (int pl = pl0; pl < pl1; ++pl, slot ^= 1) {
        if (pl + 1 < pl1) stage_carrier(pl + 1, slot ^ 1);
        const float2* ph = car + slot * n_samp + S * warp * kP;
        float2 acc[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) acc[p] = make_float2(0.f, 0.f);
        for (int q = 0; q < S; ++q) {
            const float* colq = col + q * SP;
            const float2* phq = ph + q;
            fir_phase<kP>(acc, taps, q * J, J, [&](int j) {
                const float v = colq[j * (S * SP)];
                const float2 c = phq[j * S];
                return make_float2(v * c.x, v * c.y);
            });
        }
        if (r < prm.n_rows_filled) {
            float2* out = prm.p1 + (size_t)pl * prm.plane_stride + (size_t)r * prm.pitch_d + m;
#pragma unroll
            for (int p = 0; p < kP; p += 2) {
                if (m + p + 1 < prm.pitch_d) *reinterpret_cast<float4*>(out + p) = make_float4(acc[p].x, acc[p].y, acc[p + 1].x, acc[p + 1].y);
                else if (m + p < prm.pitch_d) out[p] = acc[p];
            }
        }
        __syncthreads();
    }
}

struct MrPass2Params {
    const float2* p1;      // [chunk][n_alloc][pitch_d]
    size_t plane_stride;
    const float2* phx;     // [n_rows][n_alloc], padded-row carrier
    float2* p2;            // [chunk][n_cand][Nd][Md]
    float* pmax;           // [chunk][n_cand][nbx][nby]: max |P2|^2 over blocks of kPmB x kPmB coarse cells
    int Nd, Md, pitch_d, n_alloc, J, plane0, pstep, n_cand, row_c, row_p, nbx, nby;   // nbx, nby: ALLOCATED block grid
};

// decimating version of k_pass2: lane = decimated column, warp w owns decimated rows [w*P, w*P+P);
// the result of every candidate goes to HBM (coarse grid: 1/S^2 of a frame per candidate).
// The shared-memory plane tile (S (WARPS P + J) rows) allows one CTA per SM, so the CTA carries
// GROUPS independent warp groups that share the tile and split the candidates between them
// (named barriers per group): twice the resident warps for the same shared memory.
template <int S, int WARPS, int GROUPS>
__global__ void __launch_bounds__(GROUPS * WARPS * 32, 1)
k_mr_pass2(const MrPass2Params prm, const __grid_constant__ TapTable taps) {
    extern __shared__ float2 smem[];
    constexpr int GT = WARPS * 32;                        // threads per group
    const int group = threadIdx.x / GT, tig = threadIdx.x % GT;
    const int lane = tig & 31, warp = tig >> 5;
    const int my0 = blockIdx.x * kLanes;
    const int mx0 = blockIdx.y * (WARPS * kP);
    const int pl = blockIdx.z;
    const int plane = prm.plane0 + pl * prm.pstep;
    const int J = prm.J;
    const int n_samp = S * (WARPS * kP + J + kAhead + 1);
    {   // plane tile: rows of 32 float2 = 256 B, all copies in flight at once (cp.async, 16 B each: with a
        // load + store per row the fill was one DRAM round trip per row and warp, ~15 % of the CTA's life)
        const float2* __restrict__ src = prm.p1 + (size_t)pl * prm.plane_stride + (size_t)(S * mx0) * prm.pitch_d + my0;
        for (int i = threadIdx.x; i < n_samp * (kLanes / 2); i += GROUPS * WARPS * 32) {
            const int j = i / (kLanes / 2), c = 2 * (i % (kLanes / 2));
            cp_async16(smem + j * kLanes + c, src + (size_t)j * prm.pitch_d + c);
        }
        cp_async_commit();
    }
    const float2* col = smem + (S * warp * kP) * kLanes + lane;
    const int my = my0 + lane;
    // carrier of the tile rows, staged per candidate in shared memory (double buffered per group):
    // cheap 32-bit addressing in the FIR loop instead of 64-bit global address arithmetic per sample
    float2* const sph = smem + (size_t)n_samp * kLanes + (size_t)group * 2 * n_samp;     // [2][n_samp]
    auto stage_carrier = [&](int c, int slot) {
        const float2* __restrict__ ph = prm.phx + (size_t)(c * prm.row_c + plane * prm.row_p) * prm.n_alloc + S * mx0;
        float2* dst = sph + slot * n_samp;
        for (int j = tig; j < n_samp; j += GT) cp_async8(dst + j, ph + j);
        cp_async_commit();
    };
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(GT) : "memory"); };
    if (group < prm.n_cand) stage_carrier(group, 0);
    cp_async_wait_all();
    __syncthreads();
    int slot = 0;
    for (int c = group; c < prm.n_cand; c += GROUPS, slot ^= 1) {
        if (c + GROUPS < prm.n_cand) stage_carrier(c + GROUPS, slot ^ 1);
        const float2* ph = sph + slot * n_samp + S * warp * kP;
        float2 acc[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) acc[p] = make_float2(0.f, 0.f);
        for (int q = 0; q < S; ++q) {
            const float2* colq = col + q * kLanes;
            const float2* phq = ph + q;
            fir_phase<kP>(acc, taps, q * J, J, [&](int j) { return cmul(colq[j * (S * kLanes)], phq[j * S]); });
        }
        float2* out = prm.p2 + ((size_t)pl * prm.n_cand + c) * prm.Nd * prm.Md;
        float a2max[kP / kPmB];
#pragma unroll
        for (int hb = 0; hb < kP / kPmB; ++hb) a2max[hb] = 0.f;
        if (my < prm.Md) {
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const int mx = mx0 + warp * kP + p;
                if (mx < prm.Nd) {
                    out[(size_t)mx * prm.Md + my] = acc[p];
                    a2max[p / kPmB] = fmaxf(a2max[p / kPmB], fmaf(acc[p].x, acc[p].x, acc[p].y * acc[p].y));
                }
            }
        }
        // block maxima (kPmB x kPmB coarse cells) for the interpolation kernel's branch and bound
#pragma unroll
        for (int hb = 0; hb < kP / kPmB; ++hb) {
#pragma unroll
            for (int o = kPmB / 2; o > 0; o >>= 1) a2max[hb] = fmaxf(a2max[hb], __shfl_xor_sync(0xffffffffu, a2max[hb], o));
            if ((lane & (kPmB - 1)) == 0 && prm.pmax != nullptr)
                prm.pmax[(((size_t)pl * prm.n_cand + c) * prm.nbx + ((mx0 + warp * kP) / kPmB + hb)) * prm.nby +
                         (my0 + lane) / kPmB] = a2max[hb];
        }
        cp_async_wait_all();
        group_sync();     // this group's next carrier is complete; the current one is no longer read
    }
}

// Statically scheduled variant of k_mr_pass2 for JT taps per phase (JT a compile-time multiple of 4):
// the JT tap pairs of a phase sit in uniform registers, every sample is demodulated once and applied
// to all the outputs it reaches with compile-time tap / accumulator indices — no rotating window, no
// register moves, no tail branches.  acc[p] += g_q[k - p] * sample_q[k], 0 <= k - p < JT.
template <int S, int WARPS, int GROUPS, int JT>
__global__ void __launch_bounds__(GROUPS * WARPS * 32, 1)
k_mr_pass2s(const MrPass2Params prm, const __grid_constant__ TapTable taps) {
    extern __shared__ float2 smem[];
    constexpr int GT = WARPS * 32;
    const int group = threadIdx.x / GT, tig = threadIdx.x % GT;
    const int lane = tig & 31, warp = tig >> 5;
    const int my0 = blockIdx.x * kLanes;
    const int mx0 = blockIdx.y * (WARPS * kP);
    const int pl = blockIdx.z;
    const int plane = prm.plane0 + pl * prm.pstep;
    constexpr int n_samp = S * (WARPS * kP + JT);
    {   // plane tile: rows of 32 float2 = 256 B, all copies in flight at once (cp.async, 16 B each: with a
        // load + store per row the fill was one DRAM round trip per row and warp, ~15 % of the CTA's life)
        const float2* __restrict__ src = prm.p1 + (size_t)pl * prm.plane_stride + (size_t)(S * mx0) * prm.pitch_d + my0;
        for (int i = threadIdx.x; i < n_samp * (kLanes / 2); i += GROUPS * WARPS * 32) {
            const int j = i / (kLanes / 2), c = 2 * (i % (kLanes / 2));
            cp_async16(smem + j * kLanes + c, src + (size_t)j * prm.pitch_d + c);
        }
        cp_async_commit();
    }
    const float2* col = smem + (S * warp * kP) * kLanes + lane;
    const int my = my0 + lane;
    float2* const sph = smem + (size_t)n_samp * kLanes + (size_t)group * 2 * n_samp;     // [2][n_samp]
    auto stage_carrier = [&](int c, int slot) {
        const float2* __restrict__ ph = prm.phx + (size_t)(c * prm.row_c + plane * prm.row_p) * prm.n_alloc + S * mx0;
        float2* dst = sph + slot * n_samp;
        for (int j = tig; j < n_samp; j += GT) cp_async8(dst + j, ph + j);
        cp_async_commit();
    };
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(GT) : "memory"); };
    if (group < prm.n_cand) stage_carrier(group, 0);
    cp_async_wait_all();
    __syncthreads();
    int slot = 0;
    for (int c = group; c < prm.n_cand; c += GROUPS, slot ^= 1) {
        if (c + GROUPS < prm.n_cand) stage_carrier(c + GROUPS, slot ^ 1);
        const float2* ph = sph + slot * n_samp + S * warp * kP;
        float2 acc[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) acc[p] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int q = 0; q < S; ++q) {
            const float2* colq = col + q * kLanes;
            const float2* phq = ph + q;
            float2 g[JT];
#pragma unroll
            for (int j = 0; j < JT; ++j) g[j] = taps.g[q * JT + j];
#pragma unroll
            for (int k = 0; k < kP + JT - 1; ++k) {
                const float2 smp = cmul(colq[k * (S * kLanes)], phq[k * S]);
#pragma unroll
                for (int p = 0; p < kP; ++p)
                    if (k - p >= 0 && k - p < JT) acc[p] = __ffma2_rn(g[k - p], smp, acc[p]);
            }
        }
        float2* out = prm.p2 + ((size_t)pl * prm.n_cand + c) * prm.Nd * prm.Md;
        float a2max[kP / kPmB];
#pragma unroll
        for (int hb = 0; hb < kP / kPmB; ++hb) a2max[hb] = 0.f;
        if (my < prm.Md) {
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const int mx = mx0 + warp * kP + p;
                if (mx < prm.Nd) {
                    out[(size_t)mx * prm.Md + my] = acc[p];
                    a2max[p / kPmB] = fmaxf(a2max[p / kPmB], fmaf(acc[p].x, acc[p].x, acc[p].y * acc[p].y));
                }
            }
        }
#pragma unroll
        for (int hb = 0; hb < kP / kPmB; ++hb) {
#pragma unroll
            for (int o = kPmB / 2; o > 0; o >>= 1) a2max[hb] = fmaxf(a2max[hb], __shfl_xor_sync(0xffffffffu, a2max[hb], o));
            if ((lane & (kPmB - 1)) == 0 && prm.pmax != nullptr)
                prm.pmax[(((size_t)pl * prm.n_cand + c) * prm.nbx + ((mx0 + warp * kP) / kPmB + hb)) * prm.nby +
                         (my0 + lane) / kPmB] = a2max[hb];
        }
        cp_async_wait_all();
        group_sync();
    }
}

// ---------------------------------------------------------------------------------------------
// split pass 2: one shared anchor stage per plane + a coarse-rate stage per candidate
// ---------------------------------------------------------------------------------------------
// The candidates of one plane differ only in the axis-0 carrier, wx = wx0 + dw with |dw| far below
// the decimated band, so the full-rate part of the decimating filter can be shared.  Factorise
// G_a = G_1 * G_2 (sigma_a^2 = sigma_1^2 + sigma_2^2) and demodulate by the ANCHOR wx0 only:
//   stage A (per plane, k_mr_pass2 / k_mr_pass2s with a two-row carrier table)
//       A(e) = sum_x' G_1(S (e - H) - x') P1(t(x')) e^{2 pi i wx0 t(x')},   e in [0, Nd + 2H)
//   stage B (per candidate, k_mr_pass2b, coarse rate)
//       P2(mx) = c e^{2 pi i (dw - delta) S mx} sum_j h[j] e^{2 pi i delta S (mx + j - H)} A(mx + j)
//       h[j] = S G_2(S (j - H)),  delta = dw sigma_a^2 / sigma_2^2,  c = exp(2 pi^2 dw^2 sigma_a^2 sigma_1^2 / sigma_2^2)
// In the frequency domain G_1(f) G_2(f + delta) c = G_a(f + dw): the product of the anchor-centred
// G_1 and the shifted G_2 IS the candidate-centred G_a, so P2 equals the single-stage result up to
// the truncation of the factors and the aliasing of the coarse-rate G_2 (both below the existing
// 4.5 sigma truncation error for the parameters the host picks, pygpa_b200/_taps.py).
// Frame border: the reference demodulates by the carrier of the WRAPPED index t, which differs from the
// linear-phase ramp e^{2 pi i dw x'} by the constant J = e^{+-2 pi i dw N} on the rows that wrapped.
// Stage A therefore keeps the wrapped rows' contribution apart (A_edge; non-zero only within
// ceil(R_1/S) coarse rows of the frame edge) and stage B adds it back multiplied by J.
struct SplitTabParams {
    float2* phx1;      // [2][n_alloc]: anchor carrier masked to the frame body / to the wrapped halo rows
    float2* carB;      // [n_cand][NdE]
    float2* derotB;    // [n_cand][Nd]
    float2* jB;        // [n_cand][2]
    const double* wx_d;
    double wx0, ratio /* sigma_a^2 / sigma_2^2 */, cexp /* 2 pi^2 sigma_a^2 sigma_1^2 / sigma_2^2 */;
    int n_cand, N, S, H, Nd, NdE, n_alloc, Rtot;
};

__global__ void k_build_split_tables(const SplitTabParams p) {
    const int c = blockIdx.y;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, tstep = gridDim.x * blockDim.x;
    if (c == p.n_cand) {
        for (int r = t0; r < p.n_alloc; r += tstep) {
            const int xu = r - p.Rtot;
            int t = xu % p.N;
            if (t < 0) t += p.N;
            const float2 ph = phasor_turns(p.wx0 * (double)t);
            const bool body = xu >= 0 && xu < p.N;
            const float2 zero = make_float2(0.f, 0.f);
            p.phx1[r] = body ? ph : zero;
            p.phx1[p.n_alloc + r] = body ? zero : ph;
        }
        return;
    }
    const double dw = p.wx_d[c] - p.wx0;
    const double delta = dw * p.ratio;
    const float cs = (float)exp(p.cexp * dw * dw);
    for (int e = t0; e < p.NdE; e += tstep) p.carB[(size_t)c * p.NdE + e] = phasor_turns(delta * (double)(p.S * (e - p.H)));
    for (int mx = t0; mx < p.Nd; mx += tstep) {
        const float2 d = phasor_turns((dw - delta) * (double)(p.S * mx));
        p.derotB[(size_t)c * p.Nd + mx] = make_float2(cs * d.x, cs * d.y);
    }
    if (t0 == 0) {
        p.jB[2 * c] = phasor_turns(dw * (double)p.N);
        p.jB[2 * c + 1] = phasor_turns(-dw * (double)p.N);
    }
}

struct MrPass2bParams {
    const float2* A;        // [chunk][2][NdE][Md]: body / edge parts of the anchor stage
    const float2* carB;
    const float2* derotB;
    const float2* jB;
    float2* p2;             // [chunk][n_cand][Nd][Md]
    float* pmax;            // [chunk][n_cand][nbx][nby]
    int Nd, Md, NdE, H, EB /* coarse rows next to the frame edge that A_edge reaches */, n_cand, nbx, nby;
};

// CTA: kWarps * kP coarse output rows x 32 coarse columns (lane = column) of one plane; the A tile is
// staged once and every candidate of the plane streams through: per-candidate carrier / de-rotation
// rows by cp.async (double buffered, one barrier per candidate), JB-tap FIR along the rows with the
// taps in registers and compile-time tap / accumulator indices (as k_mr_pass2s).
template <int JB>
__global__ void __launch_bounds__(kWarps * 32, 2)
k_mr_pass2b(const MrPass2bParams prm, const __grid_constant__ TapTable taps) {
    constexpr int TO = kWarps * kP;            // output rows per CTA
    constexpr int TR = TO + JB - 1;            // A rows per CTA
    constexpr int NT = kWarps * 32;
    extern __shared__ float2 smem[];
    float2* const tB = smem;                   // [TR][32]
    float2* const tE = tB + TR * kLanes;       // [TR][32]
    float2* const scar = tE + TR * kLanes;     // [2][TR]
    float2* const sder = scar + 2 * TR;        // [2][TO]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int my0 = blockIdx.x * kLanes;
    const int mx0 = blockIdx.y * TO;
    const int pl = blockIdx.z;
    const int Nd = prm.Nd, Md = prm.Md, NdE = prm.NdE;
    const float2 zero = make_float2(0.f, 0.f);
    const int lo_end = prm.H + prm.EB, hi_begin = Nd + prm.H - prm.EB;     // A_edge is zero on rows [lo_end, hi_begin)
    {
        const float2* __restrict__ Ab = prm.A + (size_t)pl * 2 * NdE * Md;
        const float2* __restrict__ Ae = Ab + (size_t)NdE * Md;
        const bool cta_edge = mx0 < lo_end || mx0 + TR > hi_begin;
        for (int i = threadIdx.x; i < TR * kLanes; i += NT) {
            const int e = mx0 + i / kLanes, col = my0 + i % kLanes;
            if (e < NdE && col < Md) {
                cp_async8(tB + i, Ab + (size_t)e * Md + col);
                if (cta_edge) cp_async8(tE + i, Ae + (size_t)e * Md + col);
            } else {
                tB[i] = zero;
                if (cta_edge) tE[i] = zero;
            }
        }
        cp_async_commit();
    }
    auto stage = [&](int c, int slot) {
        for (int j = threadIdx.x; j < TR + TO; j += NT) {
            if (j < TR) {
                const int e = mx0 + j;
                if (e < NdE) cp_async8(scar + slot * TR + j, prm.carB + (size_t)c * NdE + e);
                else scar[slot * TR + j] = zero;
            } else {
                const int mx = mx0 + j - TR;
                if (mx < Nd) cp_async8(sder + slot * TO + j - TR, prm.derotB + (size_t)c * Nd + mx);
                else sder[slot * TO + j - TR] = zero;
            }
        }
        cp_async_commit();
    };
    stage(0, 0);
    cp_async_wait_all();
    __syncthreads();
    const int e0 = mx0 + warp * kP;                          // first A row of this warp
    const bool edge = e0 < lo_end || e0 + kP + JB - 1 > hi_begin;   // warp-uniform
    const int e_mid = prm.H + Nd / 2;                        // rows below wrapped downwards (J_lo), the others upwards (J_hi)
    const float2* colB = tB + (warp * kP) * kLanes + lane;
    const float2* colE = tE + (warp * kP) * kLanes + lane;
    const int my = my0 + lane;
    float2 g[JB];
#pragma unroll
    for (int j = 0; j < JB; ++j) g[j] = taps.g[j];
    for (int c = 0; c < prm.n_cand; ++c) {
        const int slot = c & 1;
        if (c + 1 < prm.n_cand) stage(c + 1, slot ^ 1);
        const float2* car = scar + slot * TR + warp * kP;
        float2 acc[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) acc[p] = zero;
        if (!edge) {
#pragma unroll
            for (int k = 0; k < kP + JB - 1; ++k) {
                const float2 smp = cmul(colB[k * kLanes], car[k]);
#pragma unroll
                for (int p = 0; p < kP; ++p)
                    if (k - p >= 0 && k - p < JB) acc[p] = __ffma2_rn(g[k - p], smp, acc[p]);
            }
        } else {
            const float2 jlo = __ldg(prm.jB + 2 * c), jhi = __ldg(prm.jB + 2 * c + 1);
#pragma unroll
            for (int k = 0; k < kP + JB - 1; ++k) {
                const float2 jj = (e0 + k < e_mid) ? jlo : jhi;
                const float2 ed = cmul(colE[k * kLanes], jj);
                const float2 bd = colB[k * kLanes];
                const float2 smp = cmul(make_float2(bd.x + ed.x, bd.y + ed.y), car[k]);
#pragma unroll
                for (int p = 0; p < kP; ++p)
                    if (k - p >= 0 && k - p < JB) acc[p] = __ffma2_rn(g[k - p], smp, acc[p]);
            }
        }
        const float2* der = sder + slot * TO + warp * kP;
        float2* out = prm.p2 + ((size_t)pl * prm.n_cand + c) * Nd * Md;
        float a2max[kP / kPmB];
#pragma unroll
        for (int hb = 0; hb < kP / kPmB; ++hb) a2max[hb] = 0.f;
        if (mx0 + warp * kP + kP <= Nd) {      // whole row block inside the grid (warp-uniform): no per-row guards
            if (my < Md) {
                float2* o = out + (size_t)(mx0 + warp * kP) * Md + my;
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    const float2 v = cmul(acc[p], der[p]);
                    o[(size_t)p * Md] = v;
                    a2max[p / kPmB] = fmaxf(a2max[p / kPmB], fmaf(v.x, v.x, v.y * v.y));
                }
            }
        } else if (my < Md) {
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const int mx = mx0 + warp * kP + p;
                if (mx < Nd) {
                    const float2 v = cmul(acc[p], der[p]);
                    out[(size_t)mx * Md + my] = v;
                    a2max[p / kPmB] = fmaxf(a2max[p / kPmB], fmaf(v.x, v.x, v.y * v.y));
                }
            }
        }
#pragma unroll
        for (int hb = 0; hb < kP / kPmB; ++hb) {
#pragma unroll
            for (int o = kPmB / 2; o > 0; o >>= 1) a2max[hb] = fmaxf(a2max[hb], __shfl_xor_sync(0xffffffffu, a2max[hb], o));
            if ((lane & (kPmB - 1)) == 0)
                prm.pmax[(((size_t)pl * prm.n_cand + c) * prm.nbx + ((mx0 + warp * kP) / kPmB + hb)) * prm.nby +
                         (my0 + lane) / kPmB] = a2max[hb];
        }
        cp_async_wait_all();
        __syncthreads();      // the next candidate's rows are complete; this one's are no longer read
    }
}

// 16 consecutive fine outputs from kP/S + kMrW - 1 coarse samples; tb[phi * kMrW + w] are the
// interpolation taps (S G_b(phi + S (HL - w)), zero outside the truncation radius)
template <int S, int Q>
__device__ __forceinline__ void interp_block(float2 (&acc)[Q], const float2 (&smp)[Q / S + kMrW - 2],
                                             const TapTable& taps, int tb) {
    // Q consecutive fine outputs from Q/S + 10 coarse samples.  With Rb <= 5 S (enforced by plan_mr)
    // the distance phi + S (HL - w) exceeds Rb for w = 11 (every phase) and for w = 0 unless
    // phi = 0: those taps are identically zero and are not issued.
    static_assert(Q % S == 0, "block must hold whole coarse cells");
#pragma unroll
    for (int p = 0; p < Q; ++p) acc[p] = make_float2(0.f, 0.f);
#pragma unroll
    for (int p = 0; p < Q; p += S) acc[p] = __ffma2_rn(taps.g[tb], smp[p / S], acc[p]);
#pragma unroll
    for (int w = 1; w < kMrW - 1; ++w) {
#pragma unroll
        for (int p = 0; p < Q; ++p) acc[p] = __ffma2_rn(taps.g[tb + (p % S) * kMrW + w], smp[p / S + w], acc[p]);
    }
}

struct MrInterpParams {
    const float2* p2;      // [chunk][n_cand][Nd][Md]
    const float* pmax;     // [chunk][n_cand][nbx][nby] block maxima of |P2|^2 (k_mr_pass2)
    const unsigned short* perm;   // [tiles][count] plane order per tile (k_mr_order); used when prune != 0
    unsigned long long* key;
    int N, M, Nd, Md, plane0, pstep, n_cand, idx_c, idx_p, nbx, nby, nbx_alloc, nby_alloc, count, prune;   // nbx, nby: logical (wrap) block grid
};

constexpr int kMaxPruneCand = 2048;   // candidates per plane that the survivor list can hold

// Per tile of k_mr_interp: order the planes of the chunk by how large their best candidate can get
// inside the tile (max over candidates and over the tile's coarse-window blocks of pmax), most
// promising first.  CTA (tile, z) of k_mr_interp then handles plane perm[tile][z], so the z = 0 wave
// already records near-final winners in `key` and every later CTA prunes against tight thresholds.
template <int S>
__global__ void __launch_bounds__(256) k_mr_order(const float* __restrict__ pmax, int n_cand, int count, int nbx, int nby,
                                                  int nbx_alloc, int nby_alloc, unsigned short* __restrict__ perm) {
    constexpr int CX = kMrTX / S + kMrW - 2, CY = kMrTY / S + kMrW - 2;
    __shared__ float bound[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = blockIdx.x * kMrTY, x0 = blockIdx.y * kMrTX;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int r_lo = x0 / S - kMrHL, c_lo = y0 / S - kMrHL;
    const int bx0 = (r_lo >= 0 ? r_lo : r_lo - (kPmB - 1)) / kPmB, bx1 = (r_lo + CX - 1 >= 0 ? r_lo + CX - 1 : r_lo + CX - kPmB) / kPmB;
    const int by0 = (c_lo >= 0 ? c_lo : c_lo - (kPmB - 1)) / kPmB, by1 = (c_lo + CY - 1 >= 0 ? c_lo + CY - 1 : c_lo + CY - kPmB) / kPmB;
    // one warp per plane, lanes over candidates
    for (int pl = warp; pl < count; pl += 8) {
        float m = 0.f;
        for (int c = lane; c < n_cand; c += 32) {
            const float* __restrict__ pm = pmax + ((size_t)pl * n_cand + c) * nbx_alloc * nby_alloc;
            for (int bx = bx0; bx <= bx1; ++bx) {
                int wx = bx % nbx;
                if (wx < 0) wx += nbx;
                for (int by = by0; by <= by1; ++by) {
                    int wy = by % nby;
                    if (wy < 0) wy += nby;
                    m = fmaxf(m, __ldg(pm + wx * nby_alloc + wy));
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) bound[pl] = m;
    }
    __syncthreads();
    for (int pl = threadIdx.x; pl < count; pl += blockDim.x) {
        const float b = bound[pl];
        int rank = 0;
        for (int q = 0; q < count; ++q) rank += (bound[q] > b) || (bound[q] == b && q < pl);
        perm[(size_t)tile * count + rank] = (unsigned short)pl;
    }
}


// CTA = kMrTX x kMrTY fine pixels of one plane; all candidate rows of the plane stream through:
//   coarse tile -> smem (cp.async, double buffered), interpolate along x into smem (transposed),
//   interpolate along y in registers, |sf|^2, running arg-max (2 x 16 outputs per thread), one
//   atomicMax per pixel.  IB = bits per packed winner index (8 when n_cand <= 256, else 16).
template <int S, int IB>
__global__ void __launch_bounds__(256, 2)
k_mr_interp(const MrInterpParams prm, const __grid_constant__ TapTable taps) {
    constexpr int CX = kMrTX / S + kMrW - 2;        // coarse rows / columns held per candidate
    constexpr int CY = kMrTY / S + kMrW - 2;
    constexpr int NS = kP / S + kMrW - 2;           // coarse samples per 16 outputs
    constexpr int P3P = kMrTX + 1;                  // pitch of the x-interpolated tile [cy][x]
    constexpr int PER = (CX * CY + 255) / 256;
    constexpr int IPR = 32 / IB;                    // indices per register
    constexpr unsigned IMASK = (1u << IB) - 1u;
    constexpr int Q3 = S < 4 ? 8 : S;               // outputs per x-interpolation task (whole cells)
    constexpr int NS3 = Q3 / S + kMrW - 2;
    constexpr int N3 = CY * (kMrTX / Q3);           // x-interpolation tasks per candidate
    extern __shared__ float2 smem[];
    // smem: two coarse tiles [CX][CY] (cp.async targets), two x-interpolated tiles [CY][P3P]
    float2* const p3t0 = smem + 2 * CX * CY;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = blockIdx.x * kMrTY, x0 = blockIdx.y * kMrTX;
    // with pruning every tile visits the planes in its own order, most promising first (k_mr_order)
    const int pl = prm.prune ? (int)prm.perm[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * prm.count + blockIdx.z]
                             : (int)blockIdx.z;
    const int plane = prm.plane0 + pl * prm.pstep;
    const int Nd = prm.Nd, Md = prm.Md;
    // ---- branch and bound: a candidate whose |P2|^2 block maxima over this tile's coarse window
    // stay below the smallest winning |sf|^2 already recorded for the tile's pixels cannot win
    // anywhere in the tile (the interpolation taps are non-negative and sum to <= 1), so it is
    // dropped before any work is spent on it.  The thresholds come from `key`, which only grows.
    constexpr int SBX = kMrTX / S / kPmB, SBY = kMrTY / S / kPmB;   // bound blocks inside the tile (2 x 4 at S = 4)
    static_assert(SBX >= 1 && SBY >= 1 && kMrTX / S % kPmB == 0 && kMrTY / S % kPmB == 0, "tile must hold whole bound blocks");
    static_assert(kMrHL <= kPmB, "the interpolation halo must stay within one neighbouring bound block");
    __shared__ int s_blk[SBX][SBY];       // smallest recorded winner (float bits, >= 0) per bound block of the tile
    __shared__ int s_cnt;
    __shared__ unsigned short s_list[kMaxPruneCand];
    __shared__ unsigned s_mask[kMaxPruneCand];      // per survivor: bound blocks of the tile in which it can still win
    static_assert(SBX * SBY <= 32, "one mask bit per bound block of the tile");
    int n_live = prm.n_cand;
    const bool prune = prm.prune != 0;
    if (prune) {
        constexpr int BPX = kPmB * S;             // pixels per bound-block edge
        if (threadIdx.x < SBX * SBY) s_blk[threadIdx.x / SBY][threadIdx.x % SBY] = 0x7f7fffff;   // FLT_MAX
        __syncthreads();
        // row-major (coalesced) sweep over the tile's keys: thread = (column, row parity)
        const int col = threadIdx.x % kMrTY, r0 = threadIdx.x / kMrTY;
        constexpr int RSTEP = 256 / kMrTY;
        static_assert(256 % kMrTY == 0 && BPX % RSTEP == 0, "a thread's rows must not straddle bound blocks");
        float tmin[SBX];
#pragma unroll
        for (int i = 0; i < SBX; ++i) tmin[i] = 3.4028234e38f;
        const int y = y0 + col;
#pragma unroll
        for (int e = 0; e < kMrTX / RSTEP; ++e) {
            const int x = x0 + r0 + RSTEP * e;
            if (x < prm.N && y < prm.M)
                tmin[(RSTEP * e) / BPX] = fminf(tmin[(RSTEP * e) / BPX],
                                               __uint_as_float((unsigned)(prm.key[(size_t)x * prm.M + y] >> 32)));   // only grows: any value read is a valid bound
        }
        constexpr int SPAN = BPX < 32 ? BPX : 32;     // lanes of a warp that share a bound-block column
#pragma unroll
        for (int i = 0; i < SBX; ++i) {
#pragma unroll
            for (int o = SPAN / 2; o > 0; o >>= 1) tmin[i] = fminf(tmin[i], __shfl_xor_sync(0xffffffffu, tmin[i], o));
            if (lane % SPAN == 0) atomicMin(&s_blk[i][col / BPX], __float_as_int(tmin[i]));
        }
        __syncthreads();
        if (warp == 0) {
            float thr[SBX][SBY];
#pragma unroll
            for (int i = 0; i < SBX; ++i)
#pragma unroll
                for (int j = 0; j < SBY; ++j) thr[i][j] = __int_as_float(s_blk[i][j]);
            const int bx0 = (x0 / S) / kPmB - 1, by0 = (y0 / S) / kPmB - 1;      // window: one block around the tile's blocks
            int cnt = 0;
            for (int base = 0; base < prm.n_cand; base += 32) {
                const int c = base + lane;
                bool keep = false;
                unsigned bits = 0u;
                if (c < prm.n_cand) {
                    const float* __restrict__ pm = prm.pmax + ((size_t)pl * prm.n_cand + c) * prm.nbx_alloc * prm.nby_alloc;
                    float m[SBX + 2][SBY + 2];
#pragma unroll
                    for (int i = 0; i < SBX + 2; ++i) {
                        int wx = (bx0 + i) % prm.nbx;
                        if (wx < 0) wx += prm.nbx;
#pragma unroll
                        for (int j = 0; j < SBY + 2; ++j) {
                            int wy = (by0 + j) % prm.nby;
                            if (wy < 0) wy += prm.nby;
                            m[i][j] = __ldg(pm + wx * prm.nby_alloc + wy);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < SBX; ++i)
#pragma unroll
                        for (int j = 0; j < SBY; ++j) {
                            float mm = 0.f;
#pragma unroll
                            for (int di = 0; di < 3; ++di)
#pragma unroll
                                for (int dj = 0; dj < 3; ++dj) mm = fmaxf(mm, m[i + di][j + dj]);
                            if (!(mm * 1.0002f < thr[i][j])) bits |= 1u << (i * SBY + j);
                        }
                    keep = bits != 0u;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
                    s_list[pos] = (unsigned short)c;
                    s_mask[pos] = bits;
                }
                cnt += __popc(bal);
            }
            if (lane == 0) s_cnt = cnt;
        }
        __syncthreads();
        n_live = s_cnt;
    }
    auto cand_of = [&](int i) -> int { return prune ? (int)s_list[i] : i; };
    // this thread's share of the coarse tile: fixed (row, col) offsets, wrapped once
    int off[PER];
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int t = threadIdx.x + e * 256;
        int i = x0 / S - kMrHL + t / CY, j = y0 / S - kMrHL + t % CY;
        i %= Nd; if (i < 0) i += Nd;
        j %= Md; if (j < 0) j += Md;
        off[e] = t < CX * CY ? i * Md + j : -1;
    }
    float best[2][kP];
    unsigned bidx[2][kP / IPR];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int p = 0; p < kP; ++p) best[h][p] = 0.f;
#pragma unroll
        for (int p = 0; p < kP / IPR; ++p) bidx[h][p] = 0u;
    }
    // Bound blocks touched by this thread's two output regions (rows [32h, 32h+32) x columns [16 warp, 16 warp + 16)):
    // a surviving candidate is interpolated along y only in the regions where its block bound can still win,
    // and along x only for the tasks such a region reads (exact: it cannot win or tie anywhere else).
    constexpr int BPXc = kPmB * S;
    auto region_mask = [&](int h, int w) -> unsigned {
        unsigned m = 0u;
        for (int i = (32 * h) / BPXc; i <= (32 * h + 31) / BPXc; ++i)
            for (int j = (kP * w) / BPXc; j <= (kP * w + kP - 1) / BPXc; ++j) m |= 1u << (i * SBY + j);
        return m;
    };
    // Column block of this warp's two regions: staggered by half a tile between h = 0 and h = 1, because a
    // candidate is usually alive in neighbouring blocks — the stagger spreads its regions over more warps
    // (fewer warps waiting at the per-candidate barrier for the ones that own two live regions).
    constexpr int NCB = kMrTY / kP;
    static_assert(NCB == 8, "one column block per warp");
    const int wcol[2] = {warp, (warp + NCB / 2) % NCB};
    unsigned regmask[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) regmask[h] = region_mask(h, wcol[h]);
    constexpr int PER3 = (N3 + 255) / 256;
    unsigned taskmask[PER3];       // regions (h, w) read p3t rows cy in [w kP/S, w kP/S + NS) and x in [32h, 32h+32)
#pragma unroll
    for (int e = 0; e < PER3; ++e) {
        const int t = threadIdx.x + e * 256;
        unsigned m = 0u;
        if (t < N3) {
            const int cy = t % CY, xb = t / CY;
            const int h = (xb * Q3) / 32;
            for (int w = 0; w < kMrTY / kP; ++w)
                if (cy >= w * (kP / S) && cy < w * (kP / S) + NS) m |= region_mask(h, w);
        }
        taskmask[e] = m;
    }
    const float2* __restrict__ src = prm.p2 + (size_t)pl * prm.n_cand * Nd * Md;
    auto fetch = [&](int i) {
        if (i < n_live) {
            const float2* __restrict__ g = src + (size_t)cand_of(i) * Nd * Md;
            float2* dst = smem + (i & 1) * CX * CY;
#pragma unroll
            for (int e = 0; e < PER; ++e)
                if (off[e] >= 0) cp_async8(dst + threadIdx.x + e * 256, g + off[e]);
        }
        cp_async_commit();
    };
    // along x: p3t[cy][x] = sum_w tbx[x % S][w] p2c[x / S + w][cy], tasks of Q3 outputs
    auto interp_x = [&](int c) {
        const float2* p2c = smem + (c & 1) * CX * CY;
        float2* p3t = p3t0 + (c & 1) * CY * P3P;
        const unsigned live = prune ? s_mask[c] : 0xffffffffu;
#pragma unroll
        for (int e = 0; e < PER3; ++e) {
            const int t = threadIdx.x + e * 256;
            if (t >= N3 || !(live & taskmask[e])) continue;
            const int cy = t % CY, xb = t / CY;
            float2 smp[NS3], acc[Q3];
#pragma unroll
            for (int i = 0; i < NS3; ++i) smp[i] = p2c[(xb * (Q3 / S) + i) * CY + cy];
            interp_block<S, Q3>(acc, smp, taps, 0);
#pragma unroll
            for (int p = 0; p < Q3; ++p) p3t[cy * P3P + xb * Q3 + p] = acc[p];
        }
    };
    // Software pipeline over candidates, ONE barrier per candidate: in phase c every thread
    // interpolates candidate c+1 along x (into the other p3t buffer) and candidate c along y (+ arg-max),
    // while cp.async brings in the coarse tile of candidate c+2.
    fetch(0);
    fetch(1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    if (n_live > 0) interp_x(0);
    for (int i = 0; i < n_live; ++i) {
        const int c = cand_of(i);
        cp_async_wait_all();
        __syncthreads();          // tile i+1 landed, p3t[i] complete, buffers of phase i-1 released
        fetch(i + 2);
        if (i + 1 < n_live) interp_x(i + 1);
        // ---- along y in registers + arg-max: thread = (x = lane + 32 h, 16 columns of block wcol[h])
        const float2* p3t = p3t0 + (i & 1) * CY * P3P;
        const unsigned cr = (unsigned)c * (IB == 8 ? 0x01010101u : 0x00010001u);
        const unsigned live = prune ? s_mask[i] : 0xffffffffu;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!(live & regmask[h])) continue;       // warp-uniform
            float2 smp[NS], acc[kP];
#pragma unroll
            for (int i = 0; i < NS; ++i) smp[i] = p3t[(wcol[h] * (kP / S) + i) * P3P + lane + 32 * h];
            interp_block<S, kP>(acc, smp, taps, S * kMrW);
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const float a2 = fmaf(acc[p].x, acc[p].x, acc[p].y * acc[p].y);
                if (a2 > best[h][p]) {
                    best[h][p] = a2;
                    const unsigned field = IMASK << ((p % IPR) * IB);
                    bidx[h][p / IPR] = (bidx[h][p / IPR] & ~field) | (cr & field);
                }
            }
        }
    }
    cp_async_wait_all();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int x = x0 + lane + 32 * h;
#pragma unroll
        for (int p = 0; p < kP; ++p) {
            const int y = y0 + wcol[h] * kP + p;
            if (x < prm.N && y < prm.M && best[h][p] > 0.f) {
                const unsigned cwin = (bidx[h][p / IPR] >> ((p % IPR) * IB)) & IMASK;
                const unsigned idx = cwin * (unsigned)prm.idx_c + (unsigned)(plane * prm.idx_p);
                const unsigned long long k =
                    ((unsigned long long)__float_as_uint(best[h][p]) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
                atomicMax(prm.key + (size_t)x * prm.M + y, k);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// finalize: winner's lock-in, phase gradient, k-index
// ---------------------------------------------------------------------------------------------
struct FinalizeParams {
    const float2* planes;
    size_t plane_stride;
    const float2* phx;
    const double* wx_rows;   // device copies of the candidate axes
    const double* wy_planes;
    const unsigned long long* key;
    void* lockin;  // (N, M) complex, float2 or double2
    void* grad;    // (N, M, 2) or null
    void* w;       // (2, N, M) winning k-vector or null
    int* kidx;     // may be null
    double kref_x, kref_y;
    int N, M, pitch, n_alloc, T, Rx;
    int plane0, plane_begin, plane_end;
    int list_mode, n_planes;
    int grad_mode;
    double w0x, w0y;         // 'w' of pixels that never accepted a candidate (0 for the arg-max sweeps, klist[0] for wfr4)
};

__device__ __forceinline__ double wrap_to_pi(double v) {
    // (v + pi) mod 2 pi - pi with a non-negative modulo: mathtools.py:72-75
    const double two_pi = 6.283185307179586476925286766559;
    double t = (v + 3.141592653589793238462643383279) / two_pi;
    t -= floor(t);
    return t * two_pi - 3.141592653589793238462643383279;
}

__device__ __forceinline__ double neg_arg_conj(float2 a, float2 b) {
    // -arg(a * conj(b)) = phi(a) - phi(b) (mod 2 pi) with phi = -angle
    const float re = fmaf(a.x, b.x, a.y * b.y);
    const float im = fmaf(a.y, b.x, -a.x * b.y);
    return -(double)atan2f(im, re);
}

template <typename T2>
struct real_of;
template <>
struct real_of<float2> { using type = float; };
template <>
struct real_of<double2> { using type = double; };

// Shared tail of the finalize kernels: re-reference the winner to kref, phase gradient, w, k-index.
template <typename T2>
__device__ __forceinline__ void finalize_store(const FinalizeParams& prm, size_t pix, int x, int y, unsigned idx, int row,
                                               int plane, float2 s_0, float2 s_m, float2 s_p, float2 s_ym, float2 s_yp) {
    using R = typename real_of<T2>::type;
    T2* const o_lockin = static_cast<T2*>(prm.lockin);
    R* const o_grad = static_cast<R*>(prm.grad);
    R* const o_w = static_cast<R*>(prm.w);
    const size_t npix = (size_t)prm.N * prm.M;
    const int N = prm.N, M = prm.M;
    const bool want_grad = o_grad != nullptr && prm.grad_mode != GPA_GRAD_NONE;
    const double dkx = prm.wx_rows[row] - prm.kref_x;
    const double dky = prm.wy_planes[plane] - prm.kref_y;
    const float2 rot = phasor_turns(-(dkx * (double)x + dky * (double)y));
    {
        const float2 v = cmul(s_0, rot);
        T2 o;
        o.x = v.x;
        o.y = v.y;
        o_lockin[pix] = o;
    }
    if (o_w) {
        o_w[pix] = (R)prm.wx_rows[row];
        o_w[npix + pix] = (R)prm.wy_planes[plane];
    }
    if (prm.kidx) prm.kidx[pix] = (int)idx;
    if (want_grad) {
        const double four_pi = 12.566370614359172953850573533118;
        double g0, g1;
        if (prm.grad_mode == GPA_GRAD_CENTRAL) {
            // np.gradient: central inside, one-sided (x2 after the final doubling) at the frame edge
            double d0, d1;
            if (x == 0) d0 = 2.0 * neg_arg_conj(s_p, s_0);
            else if (x == N - 1) d0 = 2.0 * neg_arg_conj(s_0, s_m);
            else d0 = neg_arg_conj(s_p, s_m);
            if (y == 0) d1 = 2.0 * neg_arg_conj(s_yp, s_0);
            else if (y == M - 1) d1 = 2.0 * neg_arg_conj(s_0, s_ym);
            else d1 = neg_arg_conj(s_yp, s_ym);
            g0 = 0.5 * wrap_to_pi(d0 + four_pi * dkx);
            g1 = 0.5 * wrap_to_pi(d1 + four_pi * dky);
        } else {
            // cuGPA.py:58-62 grad='diff': forward difference, NaN past the end
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            g0 = (x == N - 1) ? nan : 0.5 * wrap_to_pi(2.0 * neg_arg_conj(s_p, s_0) + four_pi * dkx);
            g1 = (y == M - 1) ? nan : 0.5 * wrap_to_pi(2.0 * neg_arg_conj(s_yp, s_0) + four_pi * dky);
        }
        o_grad[2 * pix] = (R)g0;
        o_grad[2 * pix + 1] = (R)g1;
    }
}

template <typename T2>   // float2: c64 / f32 outputs, double2: c128 / f64 outputs (the reference's dtypes)
__global__ void __launch_bounds__(256)
k_finalize(const FinalizeParams prm, const __grid_constant__ TapTable taps) {
    using R = typename real_of<T2>::type;
    T2* const o_lockin = static_cast<T2*>(prm.lockin);
    R* const o_grad = static_cast<R*>(prm.grad);
    R* const o_w = static_cast<R*>(prm.w);
    const size_t npix = (size_t)prm.N * prm.M;
    const int y = blockIdx.x * 32 + (threadIdx.x & 31);
    const int x = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= prm.N || y >= prm.M) return;
    const size_t pix = (size_t)x * prm.M + y;
    const unsigned long long k = prm.key[pix];
    if ((k >> 32) == 0ull) {   // nothing ever exceeded |0|: geometric_phase_analysis.py:806 keeps the zeros
        T2 z;
        z.x = 0;
        z.y = 0;
        o_lockin[pix] = z;
        if (o_grad) {
            o_grad[2 * pix] = 0;
            o_grad[2 * pix + 1] = 0;
        }
        if (o_w) {
            o_w[pix] = (R)prm.w0x;
            o_w[npix + pix] = (R)prm.w0y;
        }
        if (prm.kidx) prm.kidx[pix] = -1;
        return;
    }
    const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
    int plane, row;
    if (prm.list_mode) {
        plane = (int)idx;
        row = plane;
    } else {
        plane = (int)(idx % (unsigned)prm.n_planes);
        row = (int)(idx / (unsigned)prm.n_planes);
    }
    if (plane < prm.plane_begin || plane >= prm.plane_end) return;

    const int N = prm.N, M = prm.M, T = prm.T;
    const float2* __restrict__ A = prm.planes + (size_t)(plane - prm.plane0) * prm.plane_stride;
    const float2* __restrict__ ph = prm.phx + (size_t)row * prm.n_alloc;
    const int ym = max(y - 1, 0), yp = min(y + 1, M - 1);
    const bool want_grad = o_grad != nullptr && prm.grad_mode != GPA_GRAD_NONE;

    // padded row r holds frame row r - Rx; S(x + e) = sum_d g[d] b(x + e + d), e in {-1,0,1}
    float2 s_m = make_float2(0.f, 0.f), s_0 = s_m, s_p = s_m, s_ym = s_m, s_yp = s_m;
    const int r_last = N - 1 + 2 * prm.Rx;
    for (int j = 0; j < T + 2; ++j) {
        const int r = x - 1 + j;
        if (r < 0 || r > r_last) continue;
        const float2 c = __ldg(ph + r);
        const float2 b0 = cmul(__ldg(A + (size_t)r * prm.pitch + y), c);
        if (j >= 1 && j <= T) {
            const float g = taps.g[j - 1].x;
            s_0.x = fmaf(g, b0.x, s_0.x);
            s_0.y = fmaf(g, b0.y, s_0.y);
            if (want_grad) {
                const float2 bm = cmul(__ldg(A + (size_t)r * prm.pitch + ym), c);
                const float2 bp = cmul(__ldg(A + (size_t)r * prm.pitch + yp), c);
                s_ym.x = fmaf(g, bm.x, s_ym.x);
                s_ym.y = fmaf(g, bm.y, s_ym.y);
                s_yp.x = fmaf(g, bp.x, s_yp.x);
                s_yp.y = fmaf(g, bp.y, s_yp.y);
            }
        }
        if (want_grad) {
            if (j < T) {
                const float g = taps.g[j].x;
                s_m.x = fmaf(g, b0.x, s_m.x);
                s_m.y = fmaf(g, b0.y, s_m.y);
            }
            if (j >= 2) {
                const float g = taps.g[j - 2].x;
                s_p.x = fmaf(g, b0.x, s_p.x);
                s_p.y = fmaf(g, b0.y, s_p.y);
            }
        }
    }

    finalize_store<T2>(prm, pix, x, y, idx, row, plane, s_0, s_m, s_p, s_ym, s_yp);
}

// Multirate twin of k_finalize: the winner's sf at the pixel and its four neighbours is interpolated
// from the candidate's coarse grid P2 (still resident after gpa_sweep_argmax_mr) instead of being
// re-filtered from full-resolution planes, which removes the extra full-rate pass 1.
struct MrFinalizeParams {
    FinalizeParams f;      // planes / phx unused
    const float2* p2;      // [chunk][n_cand][Nd][Md]
    int Nd, Md, n_cand, S, pstep;
};

// winner of pixel (x, y), known to belong to one of this call's planes
template <int S, typename T2>
__device__ __forceinline__ void mr_finalize_pixel(const MrFinalizeParams& mp, const TapTable& taps, int x, int y, unsigned idx,
                                                  int plane, int row, int cand) {
    const FinalizeParams& prm = mp.f;
    const size_t pix = (size_t)x * prm.M + y;
    const int Nd = mp.Nd, Md = mp.Md;
    const float2* __restrict__ P = mp.p2 + ((size_t)((plane - prm.plane0) / mp.pstep) * mp.n_cand + cand) * Nd * Md;
    // Fine positions x-1, x, x+1 and y-1, y, y+1 in UNWRAPPED coordinates (the coarse grid is circular
    // like the frame; the reference never uses the values beyond the frame edge, they are ignored).
    // The three positions span at most two adjacent coarse cells, so a 12 x 12 coarse window holds
    // every sample: row i <-> coarse row cx0 - HL + i, column j <-> cy0 - HL + j.
    auto fdiv = [](int a, int b) { return (a >= 0 ? a : a - b + 1) / b; };
    int offx[3], phx_[3], offy[3];
    const int cx0 = fdiv(x - 1, S), cy0 = fdiv(y - 1, S);
    float gy[3][kMrW];          // y taps of the three y positions aligned to the 12-column window
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int cx = fdiv(x - 1 + e, S), cy = fdiv(y - 1 + e, S);
        offx[e] = cx - cx0;
        phx_[e] = x - 1 + e - S * cx;
        offy[e] = cy - cy0;
        const int phy_ = y - 1 + e - S * cy;
#pragma unroll
        for (int j = 0; j < kMrW; ++j) {
            const int v = j - offy[e];
            gy[e][j] = (v >= 0 && v < kMrW - 1) ? taps.g[S * kMrW + phy_ * kMrW + v].x : 0.f;
        }
    }
    int colj[kMrW];
#pragma unroll
    for (int j = 0; j < kMrW; ++j) {
        int c = (cy0 - kMrHL + j) % Md;
        colj[j] = c < 0 ? c + Md : c;
    }
    float2 s_xm = make_float2(0.f, 0.f), s_0 = s_xm, s_xp = s_xm, s_ym = s_xm, s_yp = s_xm;
#pragma unroll 1
    for (int i = 0; i < kMrW; ++i) {
        int r = (cx0 - kMrHL + i) % Nd;
        if (r < 0) r += Nd;
        const float2* __restrict__ prow = P + (size_t)r * Md;
        float2 rv[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) rv[d] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kMrW; ++j) {
            const float2 smp = __ldg(prow + colj[j]);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                rv[d].x = fmaf(gy[d][j], smp.x, rv[d].x);
                rv[d].y = fmaf(gy[d][j], smp.y, rv[d].y);
            }
        }
        // x taps are warp-uniform (a warp shares x)
        float gx[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int w = i - offx[e];
            gx[e] = (w >= 0 && w < kMrW - 1) ? taps.g[phx_[e] * kMrW + w].x : 0.f;
        }
        s_xm.x = fmaf(gx[0], rv[1].x, s_xm.x); s_xm.y = fmaf(gx[0], rv[1].y, s_xm.y);
        s_0.x = fmaf(gx[1], rv[1].x, s_0.x);   s_0.y = fmaf(gx[1], rv[1].y, s_0.y);
        s_xp.x = fmaf(gx[2], rv[1].x, s_xp.x); s_xp.y = fmaf(gx[2], rv[1].y, s_xp.y);
        s_ym.x = fmaf(gx[1], rv[0].x, s_ym.x); s_ym.y = fmaf(gx[1], rv[0].y, s_ym.y);
        s_yp.x = fmaf(gx[1], rv[2].x, s_yp.x); s_yp.y = fmaf(gx[1], rv[2].y, s_yp.y);
    }
    finalize_store<T2>(prm, pix, x, y, idx, row, plane, s_0, s_xm, s_xp, s_ym, s_yp);
}

// A warp owns 128 consecutive pixels of one frame row.  When the planes are sharded over GPUs only a
// fraction of them has its winner in this call's planes, finely interleaved (neighbouring pixels win in
// neighbouring planes, which belong to different ranks), so the warp first compacts the pixels it has to
// work on (ballot + prefix) and then processes them 32 at a time: the per-rank finalize time scales
// with the rank's share instead of staying that of the whole frame.  All pixels of a warp share x, so
// the x taps stay warp-uniform.  SPAN = pixels per warp: 32 when every plane is this call's (nothing to
// compact; the small patch keeps the gathers of a CTA in L1), 128 for a share of the planes.
template <int S, typename T2, int kFinSpan>
__global__ void __launch_bounds__(256)
k_mr_finalize(const MrFinalizeParams mp, const __grid_constant__ TapTable taps) {
    const FinalizeParams& prm = mp.f;
    using R = typename real_of<T2>::type;
    __shared__ unsigned char s_list[8][kFinSpan];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.y * 8 + warp;
    const int yb = blockIdx.x * kFinSpan;
    if (x >= prm.N) return;
    const size_t npix = (size_t)prm.N * prm.M;
    int count = 0;
#pragma unroll
    for (int j = 0; j < kFinSpan / 32; ++j) {
        const int y = yb + 32 * j + lane;
        bool own = false;
        if (y < prm.M) {
            const size_t pix = (size_t)x * prm.M + y;
            const unsigned long long k = prm.key[pix];
            if ((k >> 32) == 0ull) {      // nothing ever exceeded |0|: geometric_phase_analysis.py:806 keeps the zeros
                T2 z;
                z.x = 0;
                z.y = 0;
                static_cast<T2*>(prm.lockin)[pix] = z;
                if (prm.grad) {
                    static_cast<R*>(prm.grad)[2 * pix] = 0;
                    static_cast<R*>(prm.grad)[2 * pix + 1] = 0;
                }
                if (prm.w) {
                    static_cast<R*>(prm.w)[pix] = (R)prm.w0x;
                    static_cast<R*>(prm.w)[npix + pix] = (R)prm.w0y;
                }
                if (prm.kidx) prm.kidx[pix] = -1;
            } else {
                const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
                const int plane = prm.list_mode ? (int)idx : (int)(idx % (unsigned)prm.n_planes);
                own = plane >= prm.plane_begin && plane < prm.plane_end && (plane - prm.plane_begin) % mp.pstep == 0;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, own);
        if (own) s_list[warp][count + __popc(mask & ((1u << lane) - 1u))] = (unsigned char)(32 * j + lane);
        count += __popc(mask);
    }
    __syncwarp();
    for (int t = lane; t < count; t += 32) {
        const int y = yb + s_list[warp][t];
        const unsigned long long k = prm.key[(size_t)x * prm.M + y];
        const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
        int plane, row, cand;
        if (prm.list_mode) {
            plane = (int)idx; row = plane; cand = 0;
        } else {
            plane = (int)(idx % (unsigned)prm.n_planes);
            row = (int)(idx / (unsigned)prm.n_planes);
            cand = row;
        }
        mr_finalize_pixel<S, T2>(mp, taps, x, y, idx, plane, row, cand);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Geometry {
    int N, M, n_rows, n_planes, Rx, Ry, Tx, Ty;
    int pitch, n_alloc, chunk;
    size_t plane_stride;
    // workspace carve-up
    double *wx_d, *wy_d;
    float2 *phx, *phy, *planes;
};

static int plan(Geometry& g, int N, int M, int n_rows, int n_planes, int Rx, int Ry) {
    GPA_REQUIRE(N >= 2 && M >= 2, "frame must be at least 2x2 (got %dx%d)", N, M);
    GPA_REQUIRE(n_rows >= 1 && n_planes >= 1, "empty candidate set");
    GPA_REQUIRE(Rx >= 0 && Ry >= 0 && 2 * Rx + 1 <= kMaxTaps && 2 * Ry + 1 <= kMaxTaps,
                "filter radius out of range (Rx=%d Ry=%d, at most %d taps)", Rx, Ry, kMaxTaps);
    GPA_REQUIRE(2 * Rx + 1 <= N && 2 * Ry + 1 <= M,
                "truncated filter (2R+1 = %d x %d) must fit the frame (%d x %d)", 2 * Rx + 1, 2 * Ry + 1, N, M);
    g.N = N; g.M = M; g.n_rows = n_rows; g.n_planes = n_planes;
    g.Rx = Rx; g.Ry = Ry; g.Tx = 2 * Rx + 1; g.Ty = 2 * Ry + 1;
    g.pitch = (int)align_up((size_t)M, 32);
    g.n_alloc = ceil_div(N, kTile) * kTile + g.Tx + kAhead;   // every pass-2 tile reads this many rows past x0
    g.plane_stride = (size_t)g.n_alloc * g.pitch;
    return GPA_OK;
}

static size_t carve(Geometry& g, void* ws, size_t ws_bytes, int chunk) {
    Arena a(ws, ws_bytes);
    g.wx_d = a.take<double>(g.n_rows);
    g.wy_d = a.take<double>(g.n_planes);
    g.phx = a.take<float2>((size_t)g.n_rows * g.n_alloc);
    g.phy = a.take<float2>((size_t)g.n_planes * g.M);
    g.planes = a.take<float2>((size_t)chunk * g.plane_stride);
    g.chunk = chunk;
    return a.off;
}

// largest number of resident planes that fits ws_bytes
static int fit_chunk(Geometry& g, void* ws, size_t ws_bytes, int want) {
    size_t fixed = carve(g, nullptr, 0, 0);
    size_t per_plane = g.plane_stride * sizeof(float2);
    if (ws_bytes < fixed + per_plane + 256) return 0;
    size_t c = (ws_bytes - fixed - 256) / per_plane;
    int chunk = (int)(c < (size_t)want ? c : (size_t)want);
    carve(g, ws, ws_bytes, chunk);
    return chunk;
}

static int fill_taps(TapTable& t, const float* taps, int R) {
    GPA_REQUIRE(taps != nullptr, "taps pointer is null");
    std::memset(&t, 0, sizeof(t));
    for (int i = 0; i < 2 * R + 1; ++i) t.g[i] = make_float2(taps[i], taps[i]);
    return GPA_OK;
}

static int build_tables(const Geometry& g, const double* wx_rows, const double* wy_planes, cudaStream_t st) {
    const int per = (int)(sizeof(WList) / sizeof(double));
    for (int b = 0; b < g.n_rows; b += per) {
        WList wl;
        int n = g.n_rows - b < per ? g.n_rows - b : per;
        std::memcpy(wl.w, wx_rows + b, n * sizeof(double));
        dim3 grid(ceil_div(g.n_alloc, 256), n);
        k_build_phasors<<<grid, 256, 0, st>>>(g.phx + (size_t)b * g.n_alloc, g.wx_d + b, wl, n, g.n_alloc, g.Rx, g.N);
    }
    for (int b = 0; b < g.n_planes; b += per) {
        WList wl;
        int n = g.n_planes - b < per ? g.n_planes - b : per;
        std::memcpy(wl.w, wy_planes + b, n * sizeof(double));
        dim3 grid(ceil_div(g.M, 256), n);
        k_build_phasors<<<grid, 256, 0, st>>>(g.phy + (size_t)b * g.M, g.wy_d + b, wl, n, g.M, 0, g.M);
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

static int launch_pass1(const Geometry& g, const float* img, const TapTable& ty, int plane0, int count, cudaStream_t st) {
    Pass1Params p;
    p.img = img; p.phy = g.phy; p.planes = g.planes; p.plane_stride = g.plane_stride;
    p.N = g.N; p.M = g.M; p.pitch = g.pitch; p.n_rows_filled = g.N + 2 * g.Rx;
    p.Rx = g.Rx; p.Ry = g.Ry; p.T = g.Ty; p.plane0 = plane0;
    const size_t smem = (size_t)(kTile + g.Ty + kAhead) * 33 * sizeof(float2);
    static bool attr_set = false;
    if (!attr_set) {
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid(ceil_div(p.n_rows_filled, 32), ceil_div(g.pitch, kTile), count);
    KernelTimer timer("k_pass1", st);
    k_pass1<<<grid, kWarps * 32, smem, st>>>(p, ty);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

template <int MODE>
static int launch_pass2(const Geometry& g, const TapTable& tx, int plane0, int count, int cand_mode,
                        unsigned long long* key, void* out, int out_f64, cudaStream_t st) {
    Pass2Params p;
    p.planes = g.planes; p.plane_stride = g.plane_stride; p.phx = g.phx; p.key = key; p.out = out; p.out_f64 = out_f64;
    p.N = g.N; p.M = g.M; p.pitch = g.pitch; p.n_alloc = g.n_alloc; p.T = g.Tx; p.plane0 = plane0;
    if (cand_mode == GPA_CAND_GRID) {
        p.n_cand = g.n_rows; p.row_c = 1; p.row_p = 0; p.idx_c = g.n_planes; p.idx_p = 1;
    } else {
        p.n_cand = 1; p.row_c = 0; p.row_p = 1; p.idx_c = 0; p.idx_p = 1;
    }
    const size_t smem = (size_t)(kTile + g.Tx + kAhead) * kLanes * sizeof(float2);
    static bool attr_set = false;
    if (!attr_set) {
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_pass2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid(g.pitch / kLanes, ceil_div(g.N, kTile), count);
    KernelTimer timer(MODE == kArgmax ? "k_pass2_argmax" : "k_pass2_store", st);
    k_pass2<MODE><<<grid, kWarps * 32, smem, st>>>(p, tx);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

static int check_common(const float* img, const double* wx_rows, const double* wy_planes, int n_rows,
                        int n_planes, int cand_mode, int plane_begin, int plane_end, void* ws) {
    GPA_REQUIRE(img && wx_rows && wy_planes && ws, "null pointer argument");
    GPA_REQUIRE(cand_mode == GPA_CAND_GRID || cand_mode == GPA_CAND_LIST, "bad cand_mode %d", cand_mode);
    GPA_REQUIRE(cand_mode == GPA_CAND_GRID || n_rows == n_planes, "list mode needs n_rows == n_planes");
    GPA_REQUIRE(0 <= plane_begin && plane_begin <= plane_end && plane_end <= n_planes,
                "bad plane range [%d, %d) of %d", plane_begin, plane_end, n_planes);
    GPA_REQUIRE(cand_mode == GPA_CAND_LIST || (long long)n_rows * n_planes < 0x7fffffffLL, "too many candidates");
    GPA_REQUIRE(cand_mode == GPA_CAND_LIST || n_rows <= 65535, "grid mode supports at most 65535 rows (got %d)", n_rows);
    return GPA_OK;
}


// ---------------------------------------------------------------------------------------------
// multirate host side
// ---------------------------------------------------------------------------------------------
static int g_prune_enabled = 1;

struct MrGeometry {
    int N, M, S, Nd, Md, pitch_d, n_rows, n_planes, Rax, Ray, Rb, Jx, Jy, txd, warps2, n_alloc, n_rows_filled, n_cand;
    int nbx_alloc, nby_alloc, can_prune;
    // split pass 2 (R1 > 0): stage-A radius / taps per phase, stage-B coarse radius, extended coarse rows, row shift of P1
    int R1, J1, H, NdE, row_shift;
    size_t plane_stride, p2_stride, pm_stride, a_stride;   // elements per plane
    double *wx_d, *wy_d;
    float2 *phx, *phy, *p1, *p2;
    float2 *a_st, *phx1, *carB, *derotB, *jB;    // split pass 2: stage-A output and the carrier tables
    float* pmax;
    unsigned short* perm;
    int chunk;
};

static int plan_mr(MrGeometry& g, int N, int M, int n_rows, int n_planes, int cand_mode, int S, int Rax, int Ray, int Rb,
                   int R1 = 0, int H = 0) {
    GPA_REQUIRE(S == 2 || S == 4 || S == 8, "multirate stride must be 2, 4 or 8 (got %d)", S);
    GPA_REQUIRE(N % S == 0 && M % S == 0, "frame (%d x %d) is not divisible by the stride %d", N, M, S);
    GPA_REQUIRE(n_rows >= 1 && n_planes >= 1, "empty candidate set");
    GPA_REQUIRE(2 * Rax + 1 <= N && 2 * Ray + 1 <= M, "decimation filter must fit the frame");
    GPA_REQUIRE(Rb >= 1 && Rb <= kMrHL * S,
                "interpolation radius %d does not fit the %d-tap window at stride %d", Rb, kMrW, S);
    GPA_REQUIRE(N / S >= kMrW && M / S >= kMrW, "frame too small for the multirate sweep");
    g.N = N; g.M = M; g.S = S; g.Nd = N / S; g.Md = M / S; g.n_rows = n_rows; g.n_planes = n_planes;
    g.Rax = Rax; g.Ray = Ray; g.Rb = Rb;
    g.Jx = ceil_div(2 * Rax + 1, S); g.Jy = ceil_div(2 * Ray + 1, S);
    GPA_REQUIRE(S * g.Jx + kAhead <= kMaxTaps && S * g.Jy + kAhead <= kMaxTaps, "decimation filter too long");
    g.warps2 = S == 8 ? 4 : 8;
    g.txd = g.warps2 * kP;
    g.pitch_d = (int)align_up((size_t)g.Md, 32);
    g.n_alloc = S * (ceil_div(g.Nd, g.txd) * g.txd + g.Jx + kAhead + 1);
    g.n_rows_filled = N + S * g.Jx;
    g.row_shift = Rax;
    g.n_cand = cand_mode == GPA_CAND_GRID ? n_rows : 1;
    g.R1 = R1; g.H = H; g.J1 = 0; g.NdE = 0; g.a_stride = 0;
    if (R1 > 0) {
        // split pass 2: P1 carries a halo of R1 + S H rows each side; stage A produces Nd + 2H coarse rows
        GPA_REQUIRE(cand_mode == GPA_CAND_GRID, "the split pass 2 needs a candidate grid");
        GPA_REQUIRE(H >= 1 && 2 * H + 1 <= 23, "split pass 2: coarse radius %d out of range", H);
        GPA_REQUIRE(g.Nd >= 2 * (ceil_div(R1, S) + 1) && R1 + S * (H + 1) <= N,
                    "split pass 2: stage-A radius %d does not fit the frame", R1);
        g.J1 = ceil_div(2 * R1 + 1, S);
        GPA_REQUIRE(S * g.J1 + kAhead <= kMaxTaps, "split pass 2: stage-A filter too long");
        g.NdE = g.Nd + 2 * H;
        g.row_shift = R1 + S * H;
        g.n_alloc = S * (ceil_div(g.NdE, g.txd) * g.txd + g.J1 + kAhead + 1);
        g.n_rows_filled = S * g.NdE + S * g.J1;
        g.a_stride = (size_t)2 * g.NdE * g.Md;
    }
    g.plane_stride = (size_t)g.n_alloc * g.pitch_d;
    g.p2_stride = (size_t)g.n_cand * g.Nd * g.Md;
    g.nbx_alloc = ceil_div(g.Nd, kWarps * kP) * (kWarps * kP / kPmB);
    g.nby_alloc = g.pitch_d / kPmB;
    g.pm_stride = (size_t)g.n_cand * g.nbx_alloc * g.nby_alloc;
    g.can_prune = g.Nd % kPmB == 0 && g.Md % kPmB == 0 && g.n_cand <= kMaxPruneCand && g.n_planes <= 256;
    return GPA_OK;
}

static size_t carve_mr(MrGeometry& g, void* ws, size_t ws_bytes, int chunk) {
    Arena a(ws, ws_bytes);
    g.wx_d = a.take<double>(g.n_rows);
    g.wy_d = a.take<double>(g.n_planes);
    g.phx = a.take<float2>((size_t)g.n_rows * g.n_alloc);
    g.phy = a.take<float2>((size_t)g.n_planes * g.M);
    g.p1 = a.take<float2>((size_t)chunk * g.plane_stride);
    g.p2 = a.take<float2>((size_t)chunk * g.p2_stride);
    g.pmax = a.take<float>((size_t)chunk * g.pm_stride);
    g.perm = a.take<unsigned short>((size_t)chunk * ceil_div(g.N, kMrTX) * ceil_div(g.M, kMrTY));
    if (g.R1 > 0) {
        g.a_st = a.take<float2>((size_t)chunk * g.a_stride);
        g.phx1 = a.take<float2>((size_t)2 * g.n_alloc);
        g.carB = a.take<float2>((size_t)g.n_cand * g.NdE);
        g.derotB = a.take<float2>((size_t)g.n_cand * g.Nd);
        g.jB = a.take<float2>((size_t)2 * g.n_cand);
    }
    g.chunk = chunk;
    return a.off;
}

static int fit_chunk_mr(MrGeometry& g, void* ws, size_t ws_bytes, int want) {
    const size_t fixed = carve_mr(g, nullptr, 0, 0);
    const size_t per_plane = (g.plane_stride + g.p2_stride + g.a_stride) * sizeof(float2) + g.pm_stride * sizeof(float) +
                             (size_t)ceil_div(g.N, kMrTX) * ceil_div(g.M, kMrTY) * sizeof(unsigned short) + 1024;
    if (ws_bytes < fixed + per_plane + 512) return 0;
    size_t c = (ws_bytes - fixed - 512) / per_plane;
    const int chunk = (int)(c < (size_t)want ? c : (size_t)want);
    carve_mr(g, ws, ws_bytes, chunk);
    return chunk;
}

// polyphase table of a decimating filter: g[q * J + j] = taps[q + S j]
static int fill_polyphase(TapTable& t, const float* taps, int R, int S, int J) {
    GPA_REQUIRE(taps != nullptr, "taps pointer is null");
    std::memset(&t, 0, sizeof(t));
    const int T = 2 * R + 1;
    for (int q = 0; q < S; ++q)
        for (int j = 0; j < J; ++j) {
            const int i = q + S * j;
            const float v = i < T ? taps[i] : 0.f;
            t.g[q * J + j] = make_float2(v, v);
        }
    return GPA_OK;
}

// interpolation tables: [phi * W + w] = S * gb[phi + S (HL - w) + Rb] (x table, then y table)
static int fill_interp(TapTable& t, const float* bx, const float* by, int Rb, int S) {
    GPA_REQUIRE(bx && by, "taps pointer is null");
    std::memset(&t, 0, sizeof(t));
    for (int ax = 0; ax < 2; ++ax) {
        const float* b = ax ? by : bx;
        for (int phi = 0; phi < S; ++phi)
            for (int w = 0; w < kMrW; ++w) {
                const int d = phi + S * (kMrHL - w);
                const float v = (d >= -Rb && d <= Rb) ? (float)S * b[d + Rb] : 0.f;
                t.g[ax * S * kMrW + phi * kMrW + w] = make_float2(v, v);
            }
    }
    return GPA_OK;
}

template <int S>
static int launch_mr(const MrGeometry& g, const float* img, const TapTable& ty, const TapTable& tx, const TapTable& tb,
                     const TapTable& t2, int plane0, int pstep, int count, int cand_mode, unsigned long long* key,
                     cudaStream_t st) {
    {   // stage 1
        MrPass1Params p;
        p.img = img; p.phy = g.phy; p.p1 = g.p1; p.plane_stride = g.plane_stride;
        p.N = g.N; p.M = g.M; p.Md = g.Md; p.pitch_d = g.pitch_d; p.n_rows_filled = g.n_rows_filled;
        p.Rax = g.row_shift; p.Ray = g.Ray; p.J = g.Jy; p.plane0 = plane0; p.pstep = pstep;
        constexpr int W1 = 8;
        const size_t n_samp1 = (size_t)S * (W1 * kP + g.Jy + kAhead + 1);
        const size_t smem = (n_samp1 * 33 + 1) * sizeof(float) + 2 * n_samp1 * sizeof(float2);
        GPA_REQUIRE(smem <= 227 * 1024, "decimation filter too long for shared memory (%zu bytes)", smem);
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_pass1<S, W1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        // planes per CTA: amortise the tile fill but keep >= ~4 waves of CTAs
        const int tiles = ceil_div(g.n_rows_filled, 32) * ceil_div(g.pitch_d, W1 * kP);
        int ppc = 1;
        while (ppc < 8 && (long long)tiles * ceil_div(count, ppc * 2) >= 4 * 296) ppc *= 2;
        p.count = count; p.planes_per_cta = ppc;
        dim3 grid(ceil_div(g.n_rows_filled, 32), ceil_div(g.pitch_d, W1 * kP), ceil_div(count, ppc));
        KernelTimer timer("k_mr_pass1", st);
        k_mr_pass1<S, W1><<<grid, W1 * 32, smem, st>>>(p, ty);
    }
    {   // stage 2 (split: the anchor stage A, tx then holds the G_1 polyphase taps)
        const bool split = g.R1 > 0;
        const int Jx = split ? g.J1 : g.Jx;
        MrPass2Params p;
        p.p1 = g.p1; p.plane_stride = g.plane_stride; p.phx = g.phx; p.p2 = g.p2; p.pmax = g.pmax;
        p.nbx = g.nbx_alloc; p.nby = g.nby_alloc;
        p.Nd = g.Nd; p.Md = g.Md; p.pitch_d = g.pitch_d; p.n_alloc = g.n_alloc; p.J = Jx; p.plane0 = plane0; p.pstep = pstep;
        p.n_cand = g.n_cand;
        if (cand_mode == GPA_CAND_GRID) { p.row_c = 1; p.row_p = 0; } else { p.row_c = 0; p.row_p = 1; }
        if (split) {   // two "candidates": the body-masked and the halo-masked anchor carrier; output A[chunk][2][NdE][Md]
            p.phx = g.phx1; p.p2 = g.a_st; p.pmax = nullptr; p.Nd = g.NdE; p.n_cand = 2; p.row_c = 1; p.row_p = 0;
        }
        constexpr int W2 = S == 8 ? 4 : 8;
        constexpr int G2 = 2;
        const size_t n_samp2 = (size_t)S * (W2 * kP + Jx + kAhead + 1);
        const size_t smem = n_samp2 * (kLanes + 2 * G2) * sizeof(float2);   // plane tile + 2 carrier buffers per group
        GPA_REQUIRE(smem <= 227 * 1024, "decimation filter too long for shared memory (%zu bytes)", smem);
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_pass2<S, W2, G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        dim3 grid(g.pitch_d / kLanes, ceil_div(p.Nd, W2 * kP), count);
        KernelTimer timer(split ? "k_mr_pass2a" : "k_mr_pass2", st);
        bool launched = false;
        if constexpr (S >= 4) {   // statically scheduled variant for the common filter lengths (even taps per phase)
            const size_t ns = (size_t)S * (W2 * kP + Jx);
            const size_t smem_s = ns * (kLanes + 2 * G2) * sizeof(float2);
#define GPA_P2S(JTV)                                                                                                   \
    case JTV:                                                                                                          \
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_pass2s<S, W2, G2, JTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            227 * 1024));                                                              \
        k_mr_pass2s<S, W2, G2, JTV><<<grid, G2 * W2 * 32, smem_s, st>>>(p, tx);                                        \
        launched = true;                                                                                               \
        break;
            if (smem_s <= 227 * 1024) {
                switch (Jx) {
                    GPA_P2S(18) GPA_P2S(20) GPA_P2S(22) GPA_P2S(24) GPA_P2S(26) GPA_P2S(28) GPA_P2S(30) GPA_P2S(32)
                    default: break;
                }
            }
#undef GPA_P2S
        }
        if (!launched) k_mr_pass2<S, W2, G2><<<grid, G2 * W2 * 32, smem, st>>>(p, tx);
    }
    if (g.R1 > 0) {   // stage B: per candidate, at the coarse rate
        MrPass2bParams p;
        p.A = g.a_st; p.carB = g.carB; p.derotB = g.derotB; p.jB = g.jB; p.p2 = g.p2; p.pmax = g.pmax;
        p.Nd = g.Nd; p.Md = g.Md; p.NdE = g.NdE; p.H = g.H; p.EB = ceil_div(g.R1, S) + 1; p.n_cand = g.n_cand;
        p.nbx = g.nbx_alloc; p.nby = g.nby_alloc;
        const int JB = 2 * g.H + 1;
        const size_t smem = (size_t)(2 * (kWarps * kP + JB - 1) * kLanes + 2 * (kWarps * kP + JB - 1) + 2 * kWarps * kP) * sizeof(float2);
        dim3 grid(g.pitch_d / kLanes, ceil_div(g.Nd, kWarps * kP), count);
        KernelTimer timer("k_mr_pass2b", st);
#define GPA_P2B(JBV)                                                                                                    \
    case JBV:                                                                                                           \
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_pass2b<JBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
        k_mr_pass2b<JBV><<<grid, kWarps * 32, smem, st>>>(p, t2);                                                       \
        break;
        switch (JB) {
            GPA_P2B(13) GPA_P2B(15) GPA_P2B(17) GPA_P2B(19) GPA_P2B(21) GPA_P2B(23)
            default:
                set_error("split pass 2: unsupported coarse tap count %d", JB);
                return GPA_ERR_INVALID;
        }
#undef GPA_P2B
    }
    {   // stages 3 + 4 + arg-max
        MrInterpParams p;
        p.p2 = g.p2; p.pmax = g.pmax; p.key = key; p.N = g.N; p.M = g.M; p.Nd = g.Nd; p.Md = g.Md; p.plane0 = plane0; p.pstep = pstep; p.n_cand = g.n_cand;
        p.nbx = g.Nd / kPmB; p.nby = g.Md / kPmB; p.nbx_alloc = g.nbx_alloc; p.nby_alloc = g.nby_alloc; p.count = count;
        p.prune = g.can_prune && g_prune_enabled;
        p.perm = g.perm;
        if (p.prune) {
            dim3 tg(ceil_div(g.M, kMrTY), ceil_div(g.N, kMrTX));
            KernelTimer timer("k_mr_order", st);
            k_mr_order<S><<<tg, 256, 0, st>>>(g.pmax, g.n_cand, count, p.nbx, p.nby, g.nbx_alloc, g.nby_alloc, g.perm);
        }
        if (cand_mode == GPA_CAND_GRID) { p.idx_c = g.n_planes; p.idx_p = 1; } else { p.idx_c = 0; p.idx_p = 1; }
        constexpr int CX = kMrTX / S + kMrW - 2, CY = kMrTY / S + kMrW - 2;
        const size_t smem = (size_t)(2 * CX * CY + 2 * CY * (kMrTX + 1)) * sizeof(float2);
        dim3 grid(ceil_div(g.M, kMrTY), ceil_div(g.N, kMrTX), count);
        KernelTimer timer("k_mr_interp", st);
        if (g.n_cand <= 256) {
            GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_interp<S, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            k_mr_interp<S, 8><<<grid, 256, smem, st>>>(p, tb);
        } else {
            GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_interp<S, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            k_mr_interp<S, 16><<<grid, 256, smem, st>>>(p, tb);
        }
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_set_pruning(int on) {
    g_prune_enabled = on != 0;
    return GPA_OK;
}

extern "C" int gpa_sweep_mr_workspace_bytes(int N, int M, int n_rows, int n_planes, int cand_mode, int S, int Rax,
                                            int Ray, int Rb, int R1x, int H2x, int planes_in_flight, size_t* bytes) {
    MrGeometry g;
    int rc = plan_mr(g, N, M, n_rows, n_planes, cand_mode, S, Rax, Ray, Rb, R1x, H2x);
    if (rc) return rc;
    GPA_REQUIRE(bytes != nullptr, "bytes is null");
    GPA_REQUIRE(planes_in_flight >= 1 && planes_in_flight <= n_planes, "planes_in_flight out of range");
    *bytes = carve_mr(g, nullptr, 0, planes_in_flight) + (size_t)planes_in_flight * 1024 + 4096;
    return GPA_OK;
}

extern "C" int gpa_sweep_argmax_mr(const float* img, int N, int M, const double* wx_rows, int n_rows,
                                   const double* wy_planes, int n_planes, int cand_mode, int plane_begin,
                                   int plane_end, int plane_step, int S, const float* taps_ax, int Rax, const float* taps_ay, int Ray,
                                   const float* taps_bx, const float* taps_by, int Rb, const float* taps_1x, int R1x,
                                   const float* taps_2x, int H2x, double sigma_a, double sigma_1,
                                   unsigned long long* key, void* ws, size_t ws_bytes, void* stream) {
    MrGeometry g;
    int rc = plan_mr(g, N, M, n_rows, n_planes, cand_mode, S, Rax, Ray, Rb, R1x, H2x);
    if (rc) return rc;
    GPA_REQUIRE(R1x == 0 || (taps_1x && taps_2x && sigma_1 > 0.0 && sigma_1 < sigma_a),
                "split pass 2 needs both tap sets and 0 < sigma_1 < sigma_a");
    if ((rc = check_common(img, wx_rows, wy_planes, n_rows, n_planes, cand_mode, plane_begin, plane_end, ws))) return rc;
    GPA_REQUIRE(key != nullptr, "key is null");
    GPA_REQUIRE(plane_step >= 1, "plane_step must be >= 1");
    if (plane_begin == plane_end) return GPA_OK;
    const int total = ceil_div(plane_end - plane_begin, plane_step);     // planes begin, begin+step, ... < end
    const int chunk = fit_chunk_mr(g, ws, ws_bytes, total);
    if (chunk < 1) {
        set_error("workspace too small (%zu bytes)", ws_bytes);
        return GPA_ERR_WORKSPACE;
    }
    TapTable tx, ty, tb, t2;
    const bool split = g.R1 > 0;
    if ((rc = split ? fill_polyphase(tx, taps_1x, R1x, S, g.J1) : fill_polyphase(tx, taps_ax, Rax, S, g.Jx)) ||
        (rc = fill_polyphase(ty, taps_ay, Ray, S, g.Jy)) || (rc = fill_interp(tb, taps_bx, taps_by, Rb, S)))
        return rc;
    std::memset(&t2, 0, sizeof(t2));
    if (split)
        for (int j = 0; j < 2 * H2x + 1; ++j) t2.g[j] = make_float2(taps_2x[j], taps_2x[j]);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {   // carrier tables: same kernel as the direct path, padded-row layout of the decimating stage
        Geometry t;
        t.N = N; t.M = M; t.n_rows = n_rows; t.n_planes = n_planes; t.Rx = g.row_shift; t.n_alloc = g.n_alloc;
        t.wx_d = g.wx_d; t.wy_d = g.wy_d; t.phx = g.phx; t.phy = g.phy;
        if ((rc = build_tables(t, wx_rows, wy_planes, st))) return rc;
    }
    if (split) {   // anchor = the middle row of the candidate grid; per-candidate coarse carriers relative to it
        SplitTabParams tp;
        tp.phx1 = g.phx1; tp.carB = g.carB; tp.derotB = g.derotB; tp.jB = g.jB; tp.wx_d = g.wx_d;
        tp.wx0 = wx_rows[n_rows / 2];
        const double s2sq = sigma_a * sigma_a - sigma_1 * sigma_1;
        tp.ratio = sigma_a * sigma_a / s2sq;
        tp.cexp = 2.0 * 9.869604401089358 * sigma_a * sigma_a * sigma_1 * sigma_1 / s2sq;
        tp.n_cand = g.n_cand; tp.N = N; tp.S = S; tp.H = g.H; tp.Nd = g.Nd; tp.NdE = g.NdE; tp.n_alloc = g.n_alloc;
        tp.Rtot = g.row_shift;
        dim3 grid(ceil_div(g.n_alloc, 256) < 8 ? ceil_div(g.n_alloc, 256) : 8, g.n_cand + 1);
        k_build_split_tables<<<grid, 256, 0, st>>>(tp);
        GPA_CHECK_CUDA(cudaGetLastError());
    }
    for (int i0 = 0; i0 < total; i0 += chunk) {
        const int cnt = total - i0 < chunk ? total - i0 : chunk;
        const int p0 = plane_begin + i0 * plane_step;
        if (S == 2) rc = launch_mr<2>(g, img, ty, tx, tb, t2, p0, plane_step, cnt, cand_mode, key, st);
        else if (S == 4) rc = launch_mr<4>(g, img, ty, tx, tb, t2, p0, plane_step, cnt, cand_mode, key, st);
        else rc = launch_mr<8>(g, img, ty, tx, tb, t2, p0, plane_step, cnt, cand_mode, key, st);
        if (rc) return rc;
    }
    return GPA_OK;
}

// Finalize from the coarse grids left in the workspace by gpa_sweep_argmax_mr.  Requires that call to
// have covered exactly [plane_begin, plane_end) with all its planes resident (same ws, same arguments).
extern "C" int gpa_sweep_finalize_mr(int N, int M, const double* wx_rows, int n_rows, const double* wy_planes,
                                     int n_planes, int cand_mode, int plane_begin, int plane_end, int plane_step,
                                     int S, int Rax, int Ray, const float* taps_bx, const float* taps_by, int Rb,
                                     int R1x, int H2x, const unsigned long long* key, double kref_x, double kref_y,
                                     int grad_mode, int out_f64, void* lockin, void* grad, void* w, int* kidx, void* ws,
                                     size_t ws_bytes, void* stream) {
    MrGeometry g;
    int rc = plan_mr(g, N, M, n_rows, n_planes, cand_mode, S, Rax, Ray, Rb, R1x, H2x);
    if (rc) return rc;
    GPA_REQUIRE(wx_rows && wy_planes && ws && key && lockin, "null pointer argument");
    GPA_REQUIRE(0 <= plane_begin && plane_begin <= plane_end && plane_end <= n_planes, "bad plane range");
    GPA_REQUIRE(grad_mode == GPA_GRAD_CENTRAL || grad_mode == GPA_GRAD_FORWARD || grad_mode == GPA_GRAD_NONE,
                "bad grad_mode %d", grad_mode);
    GPA_REQUIRE(grad_mode == GPA_GRAD_NONE || grad != nullptr, "grad is null but a gradient was requested");
    GPA_REQUIRE(plane_step >= 1, "plane_step must be >= 1");
    if (plane_begin == plane_end) return GPA_OK;
    const int total = ceil_div(plane_end - plane_begin, plane_step);
    const int chunk = fit_chunk_mr(g, ws, ws_bytes, total);
    if (chunk != total) {
        set_error("gpa_sweep_finalize_mr needs every plane of the range resident (workspace holds %d of %d)", chunk, total);
        return GPA_ERR_WORKSPACE;
    }
    TapTable tb;
    if ((rc = fill_interp(tb, taps_bx, taps_by, Rb, S))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MrFinalizeParams mp;
    std::memset(&mp, 0, sizeof(mp));
    FinalizeParams& f = mp.f;
    f.wx_rows = g.wx_d; f.wy_planes = g.wy_d; f.key = key;
    f.lockin = lockin; f.grad = grad_mode == GPA_GRAD_NONE ? nullptr : grad; f.w = w; f.kidx = kidx;
    f.kref_x = kref_x; f.kref_y = kref_y; f.N = N; f.M = M;
    f.plane0 = plane_begin; f.plane_begin = plane_begin; f.plane_end = plane_end;
    f.list_mode = cand_mode == GPA_CAND_LIST; f.n_planes = n_planes; f.grad_mode = grad_mode;
    mp.p2 = g.p2; mp.Nd = g.Nd; mp.Md = g.Md; mp.n_cand = g.n_cand; mp.S = S; mp.pstep = plane_step;
    // pixels per warp: no compaction needed when every plane is ours, wider spans for smaller shares
    const bool all_planes = plane_begin == 0 && plane_end == n_planes && plane_step == 1;
    const int span = all_planes ? 32 : 128;      // measured on 2 GPUs (C3): 1.67 ms without compaction, 1.70 at 64, 1.49 at 128
    dim3 grid(ceil_div(M, span), ceil_div(N, 8));
    KernelTimer timer("k_mr_finalize", st);
#define GPA_MRFIN2(SS, TT)                                                                  \
    do {                                                                                    \
        if (span == 32) k_mr_finalize<SS, TT, 32><<<grid, 256, 0, st>>>(mp, tb);            \
        else k_mr_finalize<SS, TT, 128><<<grid, 256, 0, st>>>(mp, tb);                      \
    } while (0)
#define GPA_MRFIN(SS)                          \
    if (out_f64) GPA_MRFIN2(SS, double2);      \
    else GPA_MRFIN2(SS, float2)
    if (S == 2) { GPA_MRFIN(2); } else if (S == 4) { GPA_MRFIN(4); } else { GPA_MRFIN(8); }
#undef GPA_MRFIN2
#undef GPA_MRFIN
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

extern "C" int gpa_lockin_workspace_bytes(int N, int M, int n_rows, int n_planes, int Rx, int Ry,
                                          int planes_in_flight, size_t* bytes) {
    Geometry g;
    int rc = plan(g, N, M, n_rows, n_planes, Rx, Ry);
    if (rc) return rc;
    GPA_REQUIRE(bytes != nullptr, "bytes is null");
    GPA_REQUIRE(planes_in_flight >= 1 && planes_in_flight <= n_planes, "planes_in_flight out of range");
    *bytes = carve(g, nullptr, 0, planes_in_flight) + 256;
    return GPA_OK;
}

extern "C" int gpa_lockin_fixed(const float* img, int N, int M, double kx, double ky,
                                const float* taps_x, int Rx, const float* taps_y, int Ry,
                                int out_f64, void* out, void* ws, size_t ws_bytes, void* stream) {
    Geometry g;
    int rc = plan(g, N, M, 1, 1, Rx, Ry);
    if (rc) return rc;
    GPA_REQUIRE(img && out && ws, "null pointer argument");
    if (fit_chunk(g, ws, ws_bytes, 1) < 1) {
        set_error("workspace too small (%zu bytes)", ws_bytes);
        return GPA_ERR_WORKSPACE;
    }
    TapTable tx, ty;
    if ((rc = fill_taps(tx, taps_x, Rx)) || (rc = fill_taps(ty, taps_y, Ry))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = build_tables(g, &kx, &ky, st))) return rc;
    if ((rc = launch_pass1(g, img, ty, 0, 1, st))) return rc;
    return launch_pass2<kStore>(g, tx, 0, 1, GPA_CAND_GRID, nullptr, out, out_f64, st);
}

extern "C" int gpa_sweep_argmax(const float* img, int N, int M, const double* wx_rows, int n_rows,
                                const double* wy_planes, int n_planes, int cand_mode, int plane_begin,
                                int plane_end, const float* taps_x, int Rx, const float* taps_y, int Ry,
                                unsigned long long* key, void* ws, size_t ws_bytes, void* stream) {
    Geometry g;
    int rc = plan(g, N, M, n_rows, n_planes, Rx, Ry);
    if (rc) return rc;
    if ((rc = check_common(img, wx_rows, wy_planes, n_rows, n_planes, cand_mode, plane_begin, plane_end, ws))) return rc;
    GPA_REQUIRE(key != nullptr, "key is null");
    if (plane_begin == plane_end) return GPA_OK;
    const int chunk = fit_chunk(g, ws, ws_bytes, plane_end - plane_begin);
    if (chunk < 1) {
        set_error("workspace too small (%zu bytes)", ws_bytes);
        return GPA_ERR_WORKSPACE;
    }
    TapTable tx, ty;
    if ((rc = fill_taps(tx, taps_x, Rx)) || (rc = fill_taps(ty, taps_y, Ry))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = build_tables(g, wx_rows, wy_planes, st))) return rc;
    for (int p0 = plane_begin; p0 < plane_end; p0 += chunk) {
        const int cnt = plane_end - p0 < chunk ? plane_end - p0 : chunk;
        if ((rc = launch_pass1(g, img, ty, p0, cnt, st))) return rc;
        if ((rc = launch_pass2<kArgmax>(g, tx, p0, cnt, cand_mode, key, nullptr, 0, st))) return rc;
    }
    return GPA_OK;
}

// Direct-form finalize of planes [plane_begin, plane_end) in workspace-sized chunks (pass 1 is redone
// per chunk unless the planes of a single chunk are still resident).
static int finalize_direct(Geometry& g, const float* img, const double* wx_rows, const double* wy_planes, int cand_mode,
                           int plane_begin, int plane_end, int planes_valid, const float* taps_x, const float* taps_y,
                           const unsigned long long* key, double kref_x, double kref_y, int grad_mode, int out_f64,
                           void* lockin, void* grad, void* w, int* kidx, double w0x, double w0y, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
    int rc;
    const int chunk = fit_chunk(g, ws, ws_bytes, plane_end - plane_begin);
    if (chunk < 1) {
        set_error("workspace too small (%zu bytes)", ws_bytes);
        return GPA_ERR_WORKSPACE;
    }
    const bool reuse = planes_valid && chunk == plane_end - plane_begin;
    TapTable tx, ty;
    if ((rc = fill_taps(tx, taps_x, g.Rx)) || (rc = fill_taps(ty, taps_y, g.Ry))) return rc;
    if (!reuse && (rc = build_tables(g, wx_rows, wy_planes, st))) return rc;
    for (int p0 = plane_begin; p0 < plane_end; p0 += chunk) {
        const int cnt = plane_end - p0 < chunk ? plane_end - p0 : chunk;
        if (!reuse && (rc = launch_pass1(g, img, ty, p0, cnt, st))) return rc;
        FinalizeParams f;
        f.planes = g.planes; f.plane_stride = g.plane_stride; f.phx = g.phx;
        f.wx_rows = g.wx_d; f.wy_planes = g.wy_d; f.key = key;
        f.lockin = lockin; f.grad = grad_mode == GPA_GRAD_NONE ? nullptr : grad; f.w = w; f.kidx = kidx;
        f.kref_x = kref_x; f.kref_y = kref_y;
        f.N = g.N; f.M = g.M; f.pitch = g.pitch; f.n_alloc = g.n_alloc; f.T = g.Tx; f.Rx = g.Rx;
        f.plane0 = p0; f.plane_begin = p0; f.plane_end = p0 + cnt;
        f.list_mode = cand_mode == GPA_CAND_LIST; f.n_planes = g.n_planes; f.grad_mode = grad_mode;
        f.w0x = w0x; f.w0y = w0y;
        dim3 grid(ceil_div(g.M, 32), ceil_div(g.N, 8));
        KernelTimer timer("k_finalize", st);
        if (out_f64) k_finalize<double2><<<grid, 256, 0, st>>>(f, tx);
        else k_finalize<float2><<<grid, 256, 0, st>>>(f, tx);
        GPA_CHECK_CUDA(cudaGetLastError());
    }
    return GPA_OK;
}

extern "C" int gpa_sweep_finalize(const float* img, int N, int M, const double* wx_rows, int n_rows,
                                  const double* wy_planes, int n_planes, int cand_mode, int plane_begin,
                                  int plane_end, int planes_valid, const float* taps_x, int Rx,
                                  const float* taps_y, int Ry, const unsigned long long* key, double kref_x,
                                  double kref_y, int grad_mode, int out_f64, void* lockin, void* grad, void* w,
                                  int* kidx, void* ws, size_t ws_bytes, void* stream) {
    Geometry g;
    int rc = plan(g, N, M, n_rows, n_planes, Rx, Ry);
    if (rc) return rc;
    if ((rc = check_common(img, wx_rows, wy_planes, n_rows, n_planes, cand_mode, plane_begin, plane_end, ws))) return rc;
    GPA_REQUIRE(key && lockin, "null output pointer");
    GPA_REQUIRE(grad_mode == GPA_GRAD_CENTRAL || grad_mode == GPA_GRAD_FORWARD || grad_mode == GPA_GRAD_NONE,
                "bad grad_mode %d", grad_mode);
    GPA_REQUIRE(grad_mode == GPA_GRAD_NONE || grad != nullptr, "grad is null but a gradient was requested");
    if (plane_begin == plane_end) return GPA_OK;
    return finalize_direct(g, img, wx_rows, wy_planes, cand_mode, plane_begin, plane_end, planes_valid, taps_x, taps_y,
                           key, kref_x, kref_y, grad_mode, out_f64, lockin, grad, w, kidx, 0.0, 0.0, ws, ws_bytes,
                           static_cast<cudaStream_t>(stream));
}

// wfr4 (geometric_phase_analysis.py:839-862): ordered k-list with the neighbourhood acceptance rule.
// klist_x/klist_y: the K list entries (host); allowed: DEVICE K x K byte table, allowed[held*K + cand]
// (host-evaluated `norm(k_held - k_cand) < 2 sqrt(2) dk`).  Outputs as gpa_wfr_sweep in list mode;
// pixels that never accept a candidate keep lockin = 0 and w = klist[0] (the reference's initial state).
extern "C" int gpa_wfr4_sweep(const float* img, int N, int M, const double* klist_x, const double* klist_y, int K,
                              const unsigned char* allowed, const float* taps_x, int Rx, const float* taps_y, int Ry,
                              double kref_x, double kref_y, int out_f64, unsigned long long* key, void* lockin,
                              void* w, int* kidx, void* ws, size_t ws_bytes, void* stream) {
    Geometry g;
    int rc = plan(g, N, M, K, K, Rx, Ry);
    if (rc) return rc;
    if ((rc = check_common(img, klist_x, klist_y, K, K, GPA_CAND_LIST, 0, K, ws))) return rc;
    GPA_REQUIRE(allowed && key && lockin, "null pointer argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // initial state: amplitude 0, holding list entry 0
    k_fill_u64<<<ceil_div(N * M, 256 * 4) < 1184 ? ceil_div(N * M, 256 * 4) : 1184, 256, 0, st>>>(key, 0x00000000FFFFFFFFull,
                                                                                                  (size_t)N * M);
    const int chunk = fit_chunk(g, ws, ws_bytes, K);
    if (chunk < 1) {
        set_error("workspace too small (%zu bytes)", ws_bytes);
        return GPA_ERR_WORKSPACE;
    }
    TapTable tx, ty;
    if ((rc = fill_taps(tx, taps_x, Rx)) || (rc = fill_taps(ty, taps_y, Ry))) return rc;
    if ((rc = build_tables(g, klist_x, klist_y, st))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_pass2_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    for (int p0 = 0; p0 < K; p0 += chunk) {
        const int cnt = K - p0 < chunk ? K - p0 : chunk;
        if ((rc = launch_pass1(g, img, ty, p0, cnt, st))) return rc;
        SeqParams p;
        p.planes = g.planes; p.plane_stride = g.plane_stride; p.phx = g.phx; p.allowed = allowed; p.key = key;
        p.N = N; p.M = M; p.pitch = g.pitch; p.n_alloc = g.n_alloc; p.T = g.Tx; p.plane0 = p0; p.count = cnt; p.K = K;
        const size_t smem = (size_t)(kTile + g.Tx + kAhead) * kLanes * sizeof(float2);
        dim3 grid(g.pitch / kLanes, ceil_div(N, kTile), 1);
        KernelTimer timer("k_pass2_seq", st);
        k_pass2_seq<<<grid, kWarps * 32, smem, st>>>(p, tx);
        GPA_CHECK_CUDA(cudaGetLastError());
    }
    return finalize_direct(g, img, klist_x, klist_y, GPA_CAND_LIST, 0, K, chunk == K, taps_x, taps_y, key, kref_x, kref_y,
                           GPA_GRAD_NONE, out_f64, lockin, nullptr, w, kidx, klist_x[0], klist_y[0], ws, ws_bytes, st);
}

extern "C" int gpa_wfr_sweep(const float* img, int N, int M, const double* wx_rows, int n_rows,
                             const double* wy_planes, int n_planes, int cand_mode, const float* taps_x, int Rx,
                             const float* taps_y, int Ry, double kref_x, double kref_y, int grad_mode,
                             int out_f64, unsigned long long* key, void* lockin, void* grad, void* w, int* kidx,
                             void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(key != nullptr && N > 0 && M > 0, "key is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GPA_CHECK_CUDA(cudaMemsetAsync(key, 0, (size_t)N * M * sizeof(unsigned long long), st));
    int rc = gpa_sweep_argmax(img, N, M, wx_rows, n_rows, wy_planes, n_planes, cand_mode, 0, n_planes, taps_x,
                              Rx, taps_y, Ry, key, ws, ws_bytes, st);
    if (rc) return rc;
    return gpa_sweep_finalize(img, N, M, wx_rows, n_rows, wy_planes, n_planes, cand_mode, 0, n_planes, 1, taps_x,
                              Rx, taps_y, Ry, key, kref_x, kref_y, grad_mode, out_f64, lockin, grad, w, kidx, ws,
                              ws_bytes, st);
}
