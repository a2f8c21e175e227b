// K1 finalize kernels: k_finalize (direct form), k_mr_finalize (from the coarse grids) — part of lockin.cu (single translation unit; included inside namespace gpa).
#pragma once

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
    // owner-writes of the k-grid sharded sweep (n_dst > 0): pixel (x, y) is stored to the arrays of destination
    // x / dst_rows (local or peer-mapped over NVLink); lockin / grad above are then unused
    int n_dst, dst_rows, write_zero;
    void* lockin_dst[GPA_MAX_PEERS];
    void* grad_dst[GPA_MAX_PEERS];
};

template <typename T2>
__device__ __forceinline__ T2* lockin_of(const FinalizeParams& prm, int x) {
    return static_cast<T2*>(prm.n_dst > 0 ? prm.lockin_dst[prm.n_dst > 1 ? x / prm.dst_rows : 0] : prm.lockin);
}
template <typename R>
__device__ __forceinline__ R* grad_of(const FinalizeParams& prm, int x) {
    if (prm.grad_mode == GPA_GRAD_NONE) return nullptr;
    return static_cast<R*>(prm.n_dst > 0 ? prm.grad_dst[prm.n_dst > 1 ? x / prm.dst_rows : 0] : prm.grad);
}

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
    T2* const o_lockin = lockin_of<T2>(prm, x);
    R* const o_grad = grad_of<R>(prm, x);
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
// tap(i) = entry i of the interpolation table (x table, then y table): the kernel-parameter constant bank when the
// index is warp-uniform, a shared-memory copy when it varies per lane
template <int S, typename T2, typename Tap>
__device__ __forceinline__ void mr_finalize_pixel(const MrFinalizeParams& mp, Tap tap, int x, int y, unsigned idx,
                                                  int plane, int row, int cand) {
    const FinalizeParams& prm = mp.f;
    const size_t pix = (size_t)x * prm.M + y;
    const int Nd = mp.Nd, Md = mp.Md;
    const float2* __restrict__ P = mp.p2 + ((size_t)((plane - prm.plane0) / mp.pstep) * mp.n_cand + cand) * Nd * Md;
    // Fine positions x-1, x, x+1 and y-1, y, y+1 in UNWRAPPED coordinates (the coarse grid is circular
    // like the frame; the reference never uses the values beyond the frame edge, they are ignored).
    // The three positions span at most two adjacent coarse cells, so a 12 x 12 coarse window holds
    // every sample: row i <-> coarse row cx0 - HL + i, column j <-> cy0 - HL + j.
    // One pass over the window with packed FFMA2 (real tap x complex sample), 2 MACs per sample instead of 3:
    //   a_i  = sum_j gy_c[j] P[i][j]   -> s(x-1, y), s(x+1, y) = sum_i gx_{m,p}[i] a_i
    //   C[j] = sum_i gx_c[i] P[i][j]   -> s(x, y-1), s(x, y), s(x, y+1) = sum_j gy_{m,c,p}[j] C[j]
    constexpr int LOG = S == 2 ? 1 : (S == 4 ? 2 : 3);
    static_assert((1 << LOG) == S, "stride must be 2, 4 or 8");
    int offx[3], phx_[3], offy[3], phy_[3];
    const int cx0 = (x - 1) >> LOG, cy0 = (y - 1) >> LOG;          // floor division (arithmetic shift)
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int cx = (x - 1 + e) >> LOG, cy = (y - 1 + e) >> LOG;
        offx[e] = cx - cx0;
        phx_[e] = x - 1 + e - S * cx;
        offy[e] = cy - cy0;
        phy_[e] = y - 1 + e - S * cy;
    }
    auto ytap = [&](int e, int j) -> float {       // y tap of position e aligned to the 12-column window
        const int v = j - offy[e];
        return (v >= 0 && v < kMrW - 1) ? tap(S * kMrW + phy_[e] * kMrW + v) : 0.f;
    };
    float2 gyc[kMrW];
    int colj[kMrW];
#pragma unroll
    for (int j = 0; j < kMrW; ++j) {
        const float g = ytap(1, j);
        gyc[j] = make_float2(g, g);
        int c = cy0 - kMrHL + j;                    // in [-6, Md + 5]: one conditional wrap (Md >= 12)
        if (c < 0) c += Md;
        else if (c >= Md) c -= Md;
        colj[j] = c;
    }
    float2 C[kMrW];
#pragma unroll
    for (int j = 0; j < kMrW; ++j) C[j] = make_float2(0.f, 0.f);
    float2 s_xm = make_float2(0.f, 0.f), s_0 = s_xm, s_xp = s_xm, s_ym = s_xm, s_yp = s_xm;
#pragma unroll 1
    for (int i = 0; i < kMrW; ++i) {
        int r = cx0 - kMrHL + i;
        if (r < 0) r += Nd;
        else if (r >= Nd) r -= Nd;
        const float2* __restrict__ prow = P + (size_t)r * Md;
        float gx[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int w = i - offx[e];
            gx[e] = (w >= 0 && w < kMrW - 1) ? tap(phx_[e] * kMrW + w) : 0.f;
        }
        const float2 gxc = make_float2(gx[1], gx[1]);
        float2 smp[kMrW];
#pragma unroll
        for (int j = 0; j < kMrW; ++j) smp[j] = __ldg(prow + colj[j]);
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kMrW; ++j) {
            a = __ffma2_rn(gyc[j], smp[j], a);
            C[j] = __ffma2_rn(gxc, smp[j], C[j]);
        }
        s_xm = __ffma2_rn(make_float2(gx[0], gx[0]), a, s_xm);
        s_xp = __ffma2_rn(make_float2(gx[2], gx[2]), a, s_xp);
    }
#pragma unroll
    for (int j = 0; j < kMrW; ++j) {
        const float gm = ytap(0, j), gp = ytap(2, j);
        s_0 = __ffma2_rn(gyc[j], C[j], s_0);
        s_ym = __ffma2_rn(make_float2(gm, gm), C[j], s_ym);
        s_yp = __ffma2_rn(make_float2(gp, gp), C[j], s_yp);
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
        mr_finalize_pixel<S, T2>(mp, [&](int i) { return taps.g[i].x; }, x, y, idx, plane, row, cand);
    }
}


// Sharded twin of k_mr_finalize (k-grid split over GPUs): a rank owns only the pixels whose winner lies in its
// planes — 1/W of the frame, finely interleaved.  The CTA compacts the owned pixels of an FX x 64 tile (FX = 16, or 32 for shares below 1/6) into a
// shared list (ballot + one shared atomic per warp) and all 256 threads then work through the list, so the
// time scales with the rank's share; the tile is compact in both axes, which keeps the 12 x 12 coarse windows
// of its pixels in L1.  x varies per lane here, so the interpolation taps come from shared memory.  The same
// sequence of FMAs per pixel as k_mr_finalize: bit-identical results.  Stores go to the destination arrays
// of the pixel (FinalizeParams::lockin_dst / grad_dst, possibly peer memory).
template <int S, typename T2, int FX>
__global__ void __launch_bounds__(256)
k_mr_finalize_sharded(const MrFinalizeParams mp, const __grid_constant__ TapTable taps) {
    const FinalizeParams& prm = mp.f;
    using R = typename real_of<T2>::type;
    constexpr int FY = 64;
    __shared__ unsigned short s_list[FX * FY];
    __shared__ int s_cnt;
    __shared__ float s_tap[2 * S * kMrW];
    const int lane = threadIdx.x & 31;
    const int x0 = blockIdx.y * FX, y0 = blockIdx.x * FY;
    if (threadIdx.x == 0) s_cnt = 0;
    for (int i = threadIdx.x; i < 2 * S * kMrW; i += 256) s_tap[i] = taps.g[i].x;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < FX * FY / 256; ++j) {
        const int p = threadIdx.x + 256 * j;
        const int x = x0 + p / FY, y = y0 + p % FY;
        bool own = false;
        if (x < prm.N && y < prm.M) {
            const size_t pix = (size_t)x * prm.M + y;
            const unsigned long long k = prm.key[pix];
            if ((k >> 32) == 0ull) {      // nothing ever exceeded |0|: geometric_phase_analysis.py:806 keeps the zeros
                if (prm.write_zero) {
                    T2 z;
                    z.x = 0;
                    z.y = 0;
                    lockin_of<T2>(prm, x)[pix] = z;
                    R* const g = grad_of<R>(prm, x);
                    if (g) {
                        g[2 * pix] = 0;
                        g[2 * pix + 1] = 0;
                    }
                }
            } else {
                const unsigned idx = 0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull);
                const int plane = prm.list_mode ? (int)idx : (int)(idx % (unsigned)prm.n_planes);
                own = plane >= prm.plane_begin && plane < prm.plane_end && (plane - prm.plane_begin) % mp.pstep == 0;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, own);
        int base = 0;
        if (lane == 0 && mask) base = atomicAdd(&s_cnt, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (own) s_list[base + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)p;
    }
    __syncthreads();
    const int count = s_cnt;
    for (int t = threadIdx.x; t < count; t += 256) {
        const int p = s_list[t];
        const int x = x0 + p / FY, y = y0 + p % FY;
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
        mr_finalize_pixel<S, T2>(mp, [&](int i) { return s_tap[i]; }, x, y, idx, plane, row, cand);
    }
}
