// K1 direct-form kernels: k_pass1, k_pass2<argmax/store>, k_pass2_seq (wfr4) — part of lockin.cu (single translation unit; included inside namespace gpa).
#pragma once

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

