// K1 shared pieces: constants, tap table, carrier tables, cp.async helpers, register-blocked FIR core — part of lockin.cu (single translation unit; included inside namespace gpa).
#pragma once


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

