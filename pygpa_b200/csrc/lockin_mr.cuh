// K1 multirate kernels: k_mr_pass1, k_mr_pass2(s), split pass 2 (k_build_split_tables, k_mr_pass2b), k_mr_order, k_mr_interp — part of lockin.cu (single translation unit; included inside namespace gpa).
#pragma once

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
// ANCHOR (split pass 1): the two "planes" of the launch are the anchor plane prm.plane0 demodulated with its
// carrier masked to the frame body (pl = 0) and to the columns that wrapped around the frame edge (pl = 1);
// prm.Ray is then the column shift R_1 + S H and prm.Md / pitch_d describe the extended coarse axis.
template <int S, int WARPS, bool ANCHOR = false>
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
        const float2* __restrict__ phy = prm.phy + (size_t)(ANCHOR ? prm.plane0 : prm.plane0 + pl * prm.pstep) * M;
        float2* dst = car + slot * n_samp;
        for (int j = threadIdx.x; j < n_samp; j += WARPS * 32) {
            int c = cbase + j;
            if (c >= M) c %= M;
            float2 v = __ldg(phy + c);
            if constexpr (ANCHOR) {
                const int yu = S * m0 - prm.Ray + j;                 // unwrapped frame column of sample j
                const bool body = yu >= 0 && yu < M;
                if (body != (pl == 0)) v = make_float2(0.f, 0.f);
            }
            dst[j] = v;
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
        // ANCHOR, halo part (pl = 1): its carrier is zero on every frame-body column, so a CTA whose samples all lie inside
        // the frame (four in five at C3) has nothing to filter and stores zeros
        const bool skip = ANCHOR && pl == 1 && S * m0 - prm.Ray >= 0 && S * m0 - prm.Ray + n_samp <= M;
        for (int q = 0; q < S && !skip; ++q) {
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

// Split pass 1, coarse-rate stage (the y twin of k_mr_pass2b): all planes of a peak differ only in wy, so ONE
// anchor plane A_y(r, e) (k_mr_pass1<.., ANCHOR>, body / wrapped-column parts) is filtered at the full rate and
// every plane costs a JB-tap filter along the coarse axis:
//   P1[wy](r, my) = c e^{2 pi i (dw - delta) S my} sum_j h[j] e^{2 pi i delta S (my + j - H)} (A_body + J A_edge)(r, my + j)
// CTA: 32 padded rows (lane = row) x kWarps kP coarse output columns (warp w owns columns [w kP, w kP + kP)); the A tile is
// staged once (transposed, pitch 33) and the CTA loops over its share of the chunk's planes.
struct MrPass1bParams {
    const float2* A;        // [2][n_rows][pitch_e]
    size_t a_part;          // elements between the body and the edge part
    const float2* carB;     // [n_planes][MdE]   indexed by GLOBAL plane
    const float2* derotB;   // [n_planes][Md]
    const float2* jB;       // [n_planes][2]
    float2* p1;             // [chunk][n_alloc][pitch_d]
    size_t plane_stride;
    int n_rows, Md, MdE, pitch_d, pitch_e, H, EB, plane0, pstep, count, planes_per_cta;
};

template <int JB>
__global__ void __launch_bounds__(kWarps * 32, 2)
k_mr_pass1b(const MrPass1bParams prm, const __grid_constant__ TapTable taps) {
    constexpr int TO = kWarps * kP;            // output columns per CTA
    constexpr int TR = TO + JB - 1;            // A columns per CTA
    constexpr int NT = kWarps * 32;
    constexpr int SP = 33;
    extern __shared__ float2 smem[];
    float2* const tB = smem;                   // [TR][SP]
    float2* const tE = tB + TR * SP;           // [TR][SP]
    float2* const scar = tE + TR * SP;         // [2][TR]
    float2* const sder = scar + 2 * TR;        // [2][TO]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32;
    const int m0 = blockIdx.y * TO;
    const int Md = prm.Md, MdE = prm.MdE;
    const float2 zero = make_float2(0.f, 0.f);
    const int lo_end = prm.H + prm.EB, hi_begin = Md + prm.H - prm.EB;     // A_edge is zero on columns [lo_end, hi_begin)
    {
        const bool cta_edge = m0 < lo_end || m0 + TR > hi_begin;
        for (int rr = warp; rr < 32; rr += kWarps) {
            const int row = r0 + rr;
            const float2* __restrict__ Ab = prm.A + (size_t)row * prm.pitch_e + m0;
            const float2* __restrict__ Ae = Ab + prm.a_part;
            for (int j = lane; j < TR; j += 32) {
                if (row < prm.n_rows && m0 + j < MdE) {
                    cp_async8(tB + j * SP + rr, Ab + j);
                    if (cta_edge) cp_async8(tE + j * SP + rr, Ae + j);
                } else {
                    tB[j * SP + rr] = zero;
                    if (cta_edge) tE[j * SP + rr] = zero;
                }
            }
        }
        cp_async_commit();
    }
    const int pl0 = blockIdx.z * prm.planes_per_cta;
    const int pl1 = min(pl0 + prm.planes_per_cta, prm.count);
    auto stage = [&](int pl, int slot) {
        const int plane = prm.plane0 + pl * prm.pstep;
        for (int j = threadIdx.x; j < TR + TO; j += NT) {
            if (j < TR) {
                const int e = m0 + j;
                if (e < MdE) cp_async8(scar + slot * TR + j, prm.carB + (size_t)plane * MdE + e);
                else scar[slot * TR + j] = zero;
            } else {
                const int my = m0 + j - TR;
                if (my < Md) cp_async8(sder + slot * TO + j - TR, prm.derotB + (size_t)plane * Md + my);
                else sder[slot * TO + j - TR] = zero;
            }
        }
        cp_async_commit();
    };
    if (pl0 < pl1) stage(pl0, 0);
    cp_async_wait_all();
    __syncthreads();
    const int e0 = m0 + warp * kP;                                   // first A column of this warp
    const bool edge = e0 < lo_end || e0 + kP + JB - 1 > hi_begin;   // warp-uniform
    const int e_mid = prm.H + Md / 2;
    const float2* colB = tB + (warp * kP) * SP + lane;
    const float2* colE = tE + (warp * kP) * SP + lane;
    const int r = r0 + lane;
    const int m = m0 + warp * kP;
    float2 g[JB];
#pragma unroll
    for (int j = 0; j < JB; ++j) g[j] = taps.g[j];
    int slot = 0;
    for (int pl = pl0; pl < pl1; ++pl, slot ^= 1) {
        if (pl + 1 < pl1) stage(pl + 1, slot ^ 1);
        const float2* car = scar + slot * TR + warp * kP;
        float2 acc[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) acc[p] = zero;
        if (!edge) {
#pragma unroll
            for (int k = 0; k < kP + JB - 1; ++k) {
                const float2 smp = cmul(colB[k * SP], car[k]);
#pragma unroll
                for (int p = 0; p < kP; ++p)
                    if (k - p >= 0 && k - p < JB) acc[p] = __ffma2_rn(g[k - p], smp, acc[p]);
            }
        } else {
            const int plane = prm.plane0 + pl * prm.pstep;
            const float2 jlo = __ldg(prm.jB + 2 * plane), jhi = __ldg(prm.jB + 2 * plane + 1);
#pragma unroll
            for (int k = 0; k < kP + JB - 1; ++k) {
                const float2 jj = (e0 + k < e_mid) ? jlo : jhi;
                const float2 ed = cmul(colE[k * SP], jj);
                const float2 bd = colB[k * SP];
                const float2 smp = cmul(make_float2(bd.x + ed.x, bd.y + ed.y), car[k]);
#pragma unroll
                for (int p = 0; p < kP; ++p)
                    if (k - p >= 0 && k - p < JB) acc[p] = __ffma2_rn(g[k - p], smp, acc[p]);
            }
        }
        const float2* der = sder + slot * TO + warp * kP;
        if (r < prm.n_rows) {
            float2* out = prm.p1 + (size_t)pl * prm.plane_stride + (size_t)r * prm.pitch_d + m;
#pragma unroll
            for (int p = 0; p < kP; p += 2) {
                const float2 v0 = cmul(acc[p], der[p]), v1 = cmul(acc[p + 1], der[p + 1]);
                if (m + p + 1 < prm.pitch_d) *reinterpret_cast<float4*>(out + p) = make_float4(v0.x, v0.y, v1.x, v1.y);
                else if (m + p < prm.pitch_d) out[p] = v0;
            }
        }
        cp_async_wait_all();
        __syncthreads();      // the next plane's rows are complete; this one's are no longer read
    }
}

struct MrPass2Params {
    const float2* p1;      // [chunk][n_alloc][pitch_d]
    size_t plane_stride;
    const float2* phx;     // [n_rows][n_alloc], padded-row carrier
    float2* p2;            // [chunk][n_cand][Nd][Md]
    float* pmax;           // [chunk][nbx][nby][n_cand]: max |P2|^2 over blocks of kPmB x kPmB coarse cells (candidate-minor)
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
        // Anchor stage of the split (pmax == nullptr): the halo-masked carrier is zero on every row of the frame body, so the
        // group that owns it has nothing to filter in four CTAs out of five and leaves the SM to the other group.
        unsigned live = 1u;
        if (prm.pmax == nullptr) {
            unsigned nz = 0u;
            for (int j = tig; j < n_samp; j += GT) {
                const float2 v = sph[slot * n_samp + j];
                nz |= (v.x != 0.f) | (v.y != 0.f);
            }
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(live) : "r"(nz), "r"(group + 1), "n"(GT) : "memory");
        }
        for (int q = 0; q < S && live; ++q) {
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
                prm.pmax[(((size_t)pl * prm.nbx + ((mx0 + warp * kP) / kPmB + hb)) * prm.nby + (my0 + lane) / kPmB) * prm.n_cand + c] =
                    a2max[hb];      // candidate-minor: the survivor search / plane ordering read it with lanes over candidates
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
        // Anchor stage of the split (pmax == nullptr): the halo-masked carrier is zero on every row of the frame body, so the
        // group that owns it has nothing to filter in four CTAs out of five and leaves the SM to the other group.
        unsigned live = 1u;
        if (prm.pmax == nullptr) {
            unsigned nz = 0u;
            for (int j = tig; j < n_samp; j += GT) {
                const float2 v = sph[slot * n_samp + j];
                nz |= (v.x != 0.f) | (v.y != 0.f);
            }
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(live) : "r"(nz), "r"(group + 1), "n"(GT) : "memory");
        }
#pragma unroll 1
        for (int q = 0; q < S && live; ++q) {
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
                prm.pmax[(((size_t)pl * prm.nbx + ((mx0 + warp * kP) / kPmB + hb)) * prm.nby + (my0 + lane) / kPmB) * prm.n_cand + c] =
                    a2max[hb];      // candidate-minor: the survivor search / plane ordering read it with lanes over candidates
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
    int c_split;            // CTAs per (tile, plane): each takes a contiguous share of the candidates (small plane counts: fills the SMs)
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
    const int pl = blockIdx.z / prm.c_split;
    const int part = blockIdx.z % prm.c_split;
    const int c_begin = (prm.n_cand * part) / prm.c_split, c_end = (prm.n_cand * (part + 1)) / prm.c_split;
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
    if (c_begin < c_end) stage(c_begin, 0);
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
    for (int c = c_begin; c < c_end; ++c) {
        const int slot = (c - c_begin) & 1;
        if (c + 1 < c_end) stage(c + 1, slot ^ 1);
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
                prm.pmax[(((size_t)pl * prm.nbx + ((mx0 + warp * kP) / kPmB + hb)) * prm.nby + (my0 + lane) / kPmB) * prm.n_cand + c] =
                    a2max[hb];      // candidate-minor: the survivor search / plane ordering read it with lanes over candidates
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
    // Threshold gossip between the ranks of a k-grid sharded sweep (n_hint > 0): hint[r][block] = (epoch << 32) | float bits of
    // a LOWER bound of the final winner's |sf|^2 over the pixels of a bound block, as found by any rank so far.
    // hint[0] is this rank's array (read here), hint[1..] are the peers' (peer-mapped; every improvement is pushed to all).
    unsigned long long* hint[GPA_MAX_PEERS];
    int n_hint;
    unsigned epoch;
    // Two-phase sharded sweep (phase != 0): best_all[r * tiles + tile] = largest bound any plane of rank r reaches in the tile
    // (k_mr_order, exchanged over peer memory).  Phase 1 runs only the z = 0 CTA (the rank's most promising plane) of the tiles
    // where this rank holds the GLOBALLY most promising plane — unpruned, exactly the CTA a single GPU would run first — and
    // publishes its block bounds to every rank; phase 2 runs everything else against those thresholds.
    const float* best_all;
    int phase, rank, world, z0;
    int use_tma;           // the coarse tiles of interior CTAs come as one TMA box per candidate (tmap = P2 as a 3-D tensor)
};

// ---- TMA / mbarrier helpers (cp.async.bulk.tensor: SASS UTMALDG) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, unsigned long long* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_max_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.sys.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int kMaxPruneCand = 2048;   // candidates per plane that the survivor list can hold

// Per tile of k_mr_interp: order the planes of the chunk by how large their best candidate can get
// inside the tile (max over candidates and over the tile's coarse-window blocks of pmax), most
// promising first.  CTA (tile, z) of k_mr_interp then handles plane perm[tile][z], so the z = 0 wave
// already records near-final winners in `key` and every later CTA prunes against tight thresholds.
struct OrderShare {            // two-phase sharded sweep: where to publish the tile's best bound (row `rank` of every rank's table)
    float* best[GPA_MAX_PEERS];
    int n, rank;
};

template <int S>
__global__ void __launch_bounds__(256) k_mr_order(const float* __restrict__ pmax, int n_cand, int count, int nbx, int nby,
                                                  int nbx_alloc, int nby_alloc, unsigned short* __restrict__ perm,
                                                  const OrderShare share) {
    constexpr int CX = kMrTX / S + kMrW - 2, CY = kMrTY / S + kMrW - 2;
    __shared__ float bound[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = blockIdx.x * kMrTY, x0 = blockIdx.y * kMrTX;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int r_lo = x0 / S - kMrHL, c_lo = y0 / S - kMrHL;
    const int bx0 = (r_lo >= 0 ? r_lo : r_lo - (kPmB - 1)) / kPmB, bx1 = (r_lo + CX - 1 >= 0 ? r_lo + CX - 1 : r_lo + CX - kPmB) / kPmB;
    const int by0 = (c_lo >= 0 ? c_lo : c_lo - (kPmB - 1)) / kPmB, by1 = (c_lo + CY - 1 >= 0 ? c_lo + CY - 1 : c_lo + CY - kPmB) / kPmB;
    // wrapped offsets of the window's bound blocks, once per CTA (they used to be two integer divisions per loaded value)
    __shared__ int s_off[64];
    const int nwx = bx1 - bx0 + 1, nwy = by1 - by0 + 1, nw = nwx * nwy;      // <= (SBX + 2) (SBY + 2) <= 60
    if ((int)threadIdx.x < nw) {
        int wx = (bx0 + (int)threadIdx.x / nwy) % nbx, wy = (by0 + (int)threadIdx.x % nwy) % nby;
        if (wx < 0) wx += nbx;
        if (wy < 0) wy += nby;
        s_off[threadIdx.x] = (wx * nby_alloc + wy) * n_cand;
    }
    __syncthreads();
    // one warp per plane, lanes over candidates
    for (int pl = warp; pl < count; pl += 8) {
        float m = 0.f;
        for (int c = lane; c < n_cand; c += 32) {
            const float* __restrict__ pm = pmax + (size_t)pl * n_cand * nbx_alloc * nby_alloc + c;      // [plane][bx][by][cand]
            for (int w = 0; w < nw; ++w) m = fmaxf(m, __ldg(pm + s_off[w]));
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
        if (rank == 0 && share.n > 0) {        // this plane leads the tile: tell every rank how promising it is
            const size_t slot = (size_t)share.rank * (gridDim.x * gridDim.y) + tile;
            for (int r = 0; r < share.n; ++r) share.best[r][slot] = b;
        }
    }
}


// CTA = kMrTX x kMrTY fine pixels of one plane; all candidate rows of the plane stream through:
//   coarse tile -> smem (cp.async, double buffered), interpolate along x into smem (transposed),
//   interpolate along y in registers, |sf|^2, running arg-max (2 x 16 outputs per thread), one
//   atomicMax per pixel.  IB = bits per packed winner index (8 when n_cand <= 256, else 16).
template <int S, int IB>
__global__ void __launch_bounds__(256, 2)
k_mr_interp(const MrInterpParams prm, const __grid_constant__ TapTable taps, const __grid_constant__ CUtensorMap tmap) {
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
    extern __shared__ __align__(128) float2 smem[];
    // smem: two coarse tiles [CX][CY] (cp.async / TMA targets, 128-byte aligned), two x-interpolated tiles [CY][P3P]
    // TMA boxes must start on a 16-byte boundary of the global tensor: the window's first column c_lo = y0 / S - 5 is odd, so
    // the box starts one column earlier and is CYB = CY + 2 columns wide (inner extent a multiple of 16 bytes)
    constexpr int CYB = CY + 2;
    constexpr int CTILE = (CX * CYB * 8 + 127) / 128 * 16;     // float2 elements between the two coarse buffers
    float2* const p3t0 = smem + 2 * CTILE;
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = blockIdx.x * kMrTY, x0 = blockIdx.y * kMrTX;
    if (prm.phase != 0 && blockIdx.z == 0) {
        const int tile = blockIdx.y * gridDim.x + blockIdx.x, tiles = gridDim.x * gridDim.y;
        const float mine = prm.best_all[(size_t)prm.rank * tiles + tile];
        bool global_best = true;
        for (int r = 0; r < prm.world; ++r) {
            const float v = prm.best_all[(size_t)r * tiles + tile];
            if (v > mine || (v == mine && r < prm.rank)) global_best = false;
        }
        if ((prm.phase == 1) != global_best) return;      // phase 1: only the global best; phase 2: every other first CTA
    }
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
    __shared__ int s_cnt, s_needboot;
    __shared__ int s_boot[SBX * SBY];     // bootstrap: per bound block the candidate with the largest bound (bound bits | candidate)
    __shared__ int s_new[SBX][SBY];       // block minima of this plane's best so far
    __shared__ int s_woff[(SBX + 2) * (SBY + 2)];
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
        if (prm.n_hint > 0) {
            // another rank may already know a better lower bound of the final winners of these blocks (it owns planes
            // closer to the local optimum): any such bound is valid for every pixel of the block, so the pruning stays exact
            if (threadIdx.x < SBX * SBY) {
                const int i = threadIdx.x / SBY, j = threadIdx.x % SBY;
                const int gbx = x0 / BPX + i, gby = y0 / BPX + j;
                if (gbx < prm.nbx && gby < prm.nby) {
                    const unsigned long long h = ld_relaxed_sys_u64(prm.hint[0] + (size_t)gbx * prm.nby + gby);
                    if ((unsigned)(h >> 32) == prm.epoch) s_blk[i][j] = max(s_blk[i][j], (int)(unsigned)(h & 0xffffffffull));
                }
            }
            __syncthreads();
        }
        // bootstrap (below) is needed when a bound block of the tile has pixels but no recorded winner yet: the tile's
        // first plane, which would otherwise sweep all candidates unpruned
        if (threadIdx.x == 0) {
            int nb = 0;
            for (int i = 0; i < SBX; ++i)
                for (int j = 0; j < SBY; ++j) nb |= s_blk[i][j] == 0;
            s_needboot = nb;
        }
        if (threadIdx.x < SBX * SBY) s_boot[threadIdx.x] = 0;
        if (threadIdx.x < (SBX + 2) * (SBY + 2)) {     // element offsets of the window's bound blocks in pmax[plane][bx][by][cand]
            const int bx0 = (x0 / S) / kPmB - 1, by0 = (y0 / S) / kPmB - 1;
            int wx = (bx0 + (int)threadIdx.x / (SBY + 2)) % prm.nbx, wy = (by0 + (int)threadIdx.x % (SBY + 2)) % prm.nby;
            if (wx < 0) wx += prm.nbx;
            if (wy < 0) wy += prm.nby;
            s_woff[threadIdx.x] = (wx * prm.nby_alloc + wy) * prm.n_cand;
        }
        __syncthreads();
    }
    const bool boot = prune && s_needboot != 0;
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
    // Interior tiles (no circular wrap inside the coarse window): ONE thread issues one TMA box [CX][CY] per candidate and
    // the tile's arrival is tracked by an mbarrier; tiles on the frame border keep the per-element cp.async gather (wrap).
    const int r_lo = x0 / S - kMrHL, c_lo = y0 / S - kMrHL;
    const bool tma = prm.use_tma && r_lo >= 0 && c_lo >= 1 && ((c_lo - 1) & 1) == 0 && r_lo + CX <= Nd && c_lo - 1 + CYB <= Md;     // CTA-uniform
    const int cpitch = tma ? CYB : CY, coff = tma ? 1 : 0;     // layout of the staged coarse tile
    if (tma && threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int gb = 0;        // tiles staged by earlier passes: buffer choice and mbarrier parity continue across the passes
    auto fetch = [&](int i) {
        if (tma) {
            if (i < n_live && threadIdx.x == 0) {
                mbar_expect_tx(&s_bar[(gb + i) & 1], CX * CYB * 8);
                tma_load_3d(smem + ((gb + i) & 1) * CTILE, &tmap, &s_bar[(gb + i) & 1], 2 * (c_lo - 1), r_lo, pl * prm.n_cand + cand_of(i));
            }
            return;
        }
        if (i < n_live) {
            const float2* __restrict__ g = src + (size_t)cand_of(i) * Nd * Md;
            float2* dst = smem + ((gb + i) & 1) * CTILE;
#pragma unroll
            for (int e = 0; e < PER; ++e)
                if (off[e] >= 0) cp_async8(dst + threadIdx.x + e * 256, g + off[e]);
        }
        cp_async_commit();
    };
    auto landed = [&](int i) {          // tile i is in shared memory (for this thread; the CTA barrier that follows covers the rest)
        if (tma) {
            if (i < n_live) mbar_wait(&s_bar[(gb + i) & 1], (unsigned)((gb + i) >> 1) & 1u);
        } else {
            cp_async_wait_all();
        }
    };
    // along x: p3t[cy][x] = sum_w tbx[x % S][w] p2c[x / S + w][cy], tasks of Q3 outputs
    auto interp_x = [&](int c) {
        const float2* p2c = smem + ((gb + c) & 1) * CTILE;
        float2* p3t = p3t0 + ((gb + c) & 1) * CY * P3P;
        const unsigned live = prune ? s_mask[c] : 0xffffffffu;
#pragma unroll
        for (int e = 0; e < PER3; ++e) {
            const int t = threadIdx.x + e * 256;
            if (t >= N3 || !(live & taskmask[e])) continue;
            const int cy = t % CY, xb = t / CY;
            float2 smp[NS3], acc[Q3];
#pragma unroll
            for (int i = 0; i < NS3; ++i) smp[i] = p2c[(xb * (Q3 / S) + i) * cpitch + cy + coff];
            interp_block<S, Q3>(acc, smp, taps, 0);
#pragma unroll
            for (int p = 0; p < Q3; ++p) p3t[cy * P3P + xb * Q3 + p] = acc[p];
        }
    };
    // block minima of this plane's best so far -> s_new (all threads; ends with a barrier)
    auto block_min_of_best = [&]() {
        constexpr int BPX = kPmB * S;
        __syncthreads();                    // every warp is past its last use of the lists
        if (threadIdx.x < SBX * SBY) s_new[threadIdx.x / SBY][threadIdx.x % SBY] = 0x7f7fffff;
        __syncthreads();
        constexpr int SPANX = BPX < 32 ? BPX : 32;       // lanes (rows) of a warp inside one bound-block row
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int x = x0 + lane + 32 * h;
            float mn = 3.4028234e38f;
#pragma unroll
            for (int p = 0; p < kP; ++p) {
                const int y = y0 + wcol[h] * kP + p;
                if (x < prm.N && y < prm.M) mn = fminf(mn, best[h][p]);
            }
#pragma unroll
            for (int o = SPANX / 2; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            if (lane % SPANX == 0) atomicMin(&s_new[(lane + 32 * h) / BPX][(wcol[h] * kP) / BPX], __float_as_int(mn));
        }
        __syncthreads();
    };
    // Two passes when the tile has no recorded winners yet (its first plane): pass 0 sweeps only the candidate with the
    // largest bound of every bound block (<= SBX SBY candidates), its results become the block thresholds, and pass 1 is
    // the regular pruned sweep against them — instead of all candidates unpruned (a first-plane CTA cost 2.6 x an
    // average one; on a k-grid sharded over 8 GPUs one plane in five is a first plane).  Exact like the pruning itself:
    // the thresholds are amplitudes of candidates of this very plane.
    __syncthreads();              // mbarrier initialisation visible before the first TMA / wait
    for (int pass = boot ? 0 : 1; pass < 2; ++pass) {
    if (prune) {
        // survivor search, all 8 warps: warp w tests candidates [32 w, 32 w + 32), [32 w + 256, ...) and writes its survivors
        // in candidate order through a CTA-wide prefix over the per-(round, warp) counts — the list comes out in exactly
        // the order the serial search of warp 0 produced (it took ~10 % of a pruned CTA's life)
        __shared__ int s_wcnt[(kMaxPruneCand + 255) / 256][8];
        float thr[SBX][SBY];
#pragma unroll
        for (int i = 0; i < SBX; ++i)
#pragma unroll
            for (int j = 0; j < SBY; ++j) thr[i][j] = __int_as_float(s_blk[i][j]);
        // window: one block around the tile's blocks; its wrapped offsets were computed once per CTA (s_woff)
        constexpr int MAXR = (kMaxPruneCand + 255) / 256;
        unsigned keep_bits[MAXR];      // per round: this lane's mask (0 = dropped)
        const int rounds = (prm.n_cand + 255) / 256;
#pragma unroll
        for (int rd = 0; rd < MAXR; ++rd) {
            keep_bits[rd] = 0u;
            if (rd < rounds) {
                const int c = rd * 256 + warp * 32 + lane;
                unsigned bits = 0u;
                float bnd[SBX * SBY];
#pragma unroll
                for (int b = 0; b < SBX * SBY; ++b) bnd[b] = 0.f;
                if (c < prm.n_cand) {
                    const float* __restrict__ pm = prm.pmax + (size_t)pl * prm.n_cand * prm.nbx_alloc * prm.nby_alloc + c;   // [plane][bx][by][cand]
                    float m[SBX + 2][SBY + 2];
#pragma unroll
                    for (int i = 0; i < SBX + 2; ++i)
#pragma unroll
                        for (int j = 0; j < SBY + 2; ++j) m[i][j] = __ldg(pm + s_woff[i * (SBY + 2) + j]);
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
                            bnd[i * SBY + j] = mm;
                        }
                }
                if (pass == 0) {      // bootstrap pass: only the best-bounded candidate of every block is wanted
#pragma unroll
                    for (int b = 0; b < SBX * SBY; ++b) {
                        const unsigned k32 = c < prm.n_cand ? ((__float_as_uint(bnd[b]) & 0xfffff800u) | (unsigned)c) : 0u;
                        const unsigned best_k = __reduce_max_sync(0xffffffffu, k32);
                        if (lane == 0) atomicMax(&s_boot[b], (int)(best_k & 0x7fffffffu));
                    }
                    bits = 0u;
                }
                keep_bits[rd] = bits;
                const unsigned bal = __ballot_sync(0xffffffffu, bits != 0u);
                if (lane == 0) s_wcnt[rd][warp] = __popc(bal);
            }
        }
        __syncthreads();
        int cnt = 0;
#pragma unroll
        for (int rd = 0; rd < MAXR; ++rd) {
            if (rd < rounds) {
                int before = cnt;
                for (int w = 0; w < 8; ++w) {
                    const int n_w = s_wcnt[rd][w];
                    if (w < warp) before += n_w;
                    cnt += n_w;
                }
                const bool keep = keep_bits[rd] != 0u;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int pos = before + __popc(bal & ((1u << lane) - 1u));
                    s_list[pos] = (unsigned short)(rd * 256 + warp * 32 + lane);
                    s_mask[pos] = keep_bits[rd];
                }
            }
        }
        if (threadIdx.x == 0) {
            if (pass == 0) {      // the (distinct) bootstrap candidates, alive everywhere
                cnt = 0;
                for (int b = 0; b < SBX * SBY; ++b) {
                    const unsigned short cb = (unsigned short)(s_boot[b] & 0x7ff);
                    bool dup = false;
                    for (int k = 0; k < cnt; ++k) dup |= s_list[k] == cb;
                    if (!dup) {
                        s_list[cnt] = cb;
                        s_mask[cnt] = 0xffffffffu;
                        ++cnt;
                    }
                }
            }
            s_cnt = cnt;
        }
        __syncthreads();
        n_live = s_cnt;
    }
    // Software pipeline over candidates, ONE barrier per candidate: in phase c every thread
    // interpolates candidate c+1 along x (into the other p3t buffer) and candidate c along y (+ arg-max),
    // while cp.async brings in the coarse tile of candidate c+2.
    fetch(0);
    fetch(1);
    if (tma) landed(0);
    else asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    if (n_live > 0) interp_x(0);
    for (int i = 0; i < n_live; ++i) {
        const int c = cand_of(i);
        landed(i + 1);
        __syncthreads();          // tile i+1 landed, p3t[i] complete, buffers of phase i-1 released
        fetch(i + 2);
        if (i + 1 < n_live) interp_x(i + 1);
        // ---- along y in registers + arg-max: thread = (x = lane + 32 h, 16 columns of block wcol[h])
        const float2* p3t = p3t0 + ((gb + i) & 1) * CY * P3P;
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
                    const unsigned field = IMASK << ((p % IPR) * IB);      // a constant once the loop is unrolled
                    // bit-field insert as ONE lop3 ((a & ~c) | (b & c)) — the compiler emitted three per output
                    asm("lop3.b32 %0, %0, %1, %2, 0xD8;" : "+r"(bidx[h][p / IPR]) : "r"(cr), "r"(field));
                }
            }
        }
    }
    cp_async_wait_all();
    gb += n_live;
    if (pass == 0) {              // the bootstrap candidates' amplitudes are the thresholds of the real sweep
        block_min_of_best();
        if (threadIdx.x < SBX * SBY) {
            const int i = threadIdx.x / SBY, j = threadIdx.x % SBY;
            if (s_new[i][j] != 0x7f7fffff) s_blk[i][j] = max(s_blk[i][j] == 0x7f7fffff ? 0 : s_blk[i][j], s_new[i][j]);
        }
        __syncthreads();
        // ... and nothing else: the running winners start again from zero, so that pass 1 meets the candidates in
        // ascending order and an exact tie goes to the lowest index, as in the reference's strict `>` loop
        // (geometric_phase_analysis.py:806).  Keeping the bootstrap winners flipped a handful of exactly tied pixels
        // per 2048^2 frame (caught by bench.py's N-GPU vs 1-GPU key comparison).
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int p = 0; p < kP; ++p) best[h][p] = 0.f;
#pragma unroll
            for (int p = 0; p < kP / IPR; ++p) bidx[h][p] = 0u;
        }
    }
    }       // pass
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
    if (prune && prm.n_hint > 0) {
        // Publish what this CTA learned: min over a bound block of max(old winner, this plane's best) >= max(old block
        // minimum, block minimum of this plane's best) is a lower bound of the final winners of the block.
        constexpr int BPX = kPmB * S;
        block_min_of_best();
        if (threadIdx.x < SBX * SBY) {
            const int i = threadIdx.x / SBY, j = threadIdx.x % SBY;
            const int gbx = x0 / BPX + i, gby = y0 / BPX + j;
            const int nb = max(s_blk[i][j] == 0x7f7fffff ? 0 : s_blk[i][j], s_new[i][j] == 0x7f7fffff ? 0 : s_new[i][j]);
            if (gbx < prm.nbx && gby < prm.nby && nb > 0) {
                const size_t blk = (size_t)gbx * prm.nby + gby;
                const unsigned long long mine = ((unsigned long long)prm.epoch << 32) | (unsigned)nb;
                if (ld_relaxed_sys_u64(prm.hint[0] + blk) < mine) {      // nobody has published a better bound for this frame yet
                    for (int r = 0; r < prm.n_hint; ++r) red_max_sys_u64(prm.hint[r] + blk, mine);
                }
            }
        }
    }
}

