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
// The multirate form (default for candidate grids) replaces pass 2 by k_mr_pass1 -> anchor stage ->
// k_mr_pass2b -> k_mr_order -> k_mr_interp -> k_mr_finalize.
//
// Both filter kernels share one register-blocked FIR core: each thread owns P consecutive
// outputs along the filter axis, keeps a P-deep rotating window of complex samples in
// registers and issues P packed FFMA2 (fma.rn.f32x2: real tap x complex sample) per tap;
// taps come from the kernel-parameter constant bank through uniform registers.
//
// One translation unit; the device code lives in four headers included below inside namespace gpa
// (lockin_core.cuh: constants, carrier tables, cp.async, FIR core; lockin_direct.cuh: direct-form
// kernels; lockin_mr.cuh: multirate kernels; lockin_finalize.cuh: finalize kernels); this file holds
// the host side (geometry, workspace carve-up, launches) and the extern "C" entry points.
#include "common.cuh"

namespace gpa {

#include "lockin_core.cuh"
#include "lockin_direct.cuh"
#include "lockin_mr.cuh"
#include "lockin_finalize.cuh"

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
    GPA_CHECK_CUDA(cudaFuncSetAttribute(k_pass1, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
    GPA_CHECK_CUDA(cudaFuncSetAttribute(k_pass2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
static int g_tma_enabled = 1;

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled load_encode_tiled() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(fn);
}

// threshold gossip of a k-grid sharded sweep: armed by gpa_sweep_arm_gossip, consumed by the next gpa_sweep_argmax_mr
struct GossipArm {
    unsigned long long* hint[GPA_MAX_PEERS];
    int n = 0;
    unsigned epoch = 0;
    // two-phase sweep (optional): tables of the tiles' best bounds, flag slots of the two barriers
    bool two_phase = false;
    unsigned long long epoch64 = 0;
    int rank = 0;
    float* best[GPA_MAX_PEERS];
    void* flag_a[GPA_MAX_PEERS];
    void* flag_b[GPA_MAX_PEERS];
    const unsigned long long *wait_a = nullptr, *wait_b = nullptr;
    int* status = nullptr;
    double timeout_s = 20.0;
};
static thread_local GossipArm g_gossip;

struct MrGeometry {
    int N, M, S, Nd, Md, pitch_d, n_rows, n_planes, Rax, Ray, Rb, Jx, Jy, txd, warps2, n_alloc, n_rows_filled, n_cand;
    int nbx_alloc, nby_alloc, can_prune;
    // split pass 2 (R1 > 0): stage-A radius / taps per phase, stage-B coarse radius, extended coarse rows, row shift of P1
    int R1, J1, H, NdE, row_shift;
    // split pass 1 (R1y > 0): the same along axis 1, one anchor plane per call
    int R1y, J1y, Hy, MdE, pitch_e, col_shift, rows_a;
    size_t plane_stride, p2_stride, pm_stride, a_stride;   // elements per plane
    double *wx_d, *wy_d;
    float2 *phx, *phy, *p1, *p2;
    float2 *a_st, *phx1, *carB, *derotB, *jB;    // split pass 2: stage-A output and the carrier tables
    float2 *a_y, *carBy, *derotBy, *jBy;         // split pass 1: anchor plane (body / edge parts) and the carrier tables
    float* pmax;
    unsigned short* perm;
    int chunk;
};

static int plan_mr(MrGeometry& g, int N, int M, int n_rows, int n_planes, int cand_mode, int S, int Rax, int Ray, int Rb,
                   int R1 = 0, int H = 0, int R1y = 0, int Hy = 0) {
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
    g.R1y = R1y; g.Hy = Hy; g.J1y = 0; g.MdE = 0; g.pitch_e = 0; g.col_shift = Ray; g.rows_a = 0;
    if (R1y > 0) {
        GPA_REQUIRE(Hy >= 1 && 2 * Hy + 1 <= 23, "split pass 1: coarse radius %d out of range", Hy);
        GPA_REQUIRE(g.Md >= 2 * (ceil_div(R1y, S) + 1) && R1y + S * (Hy + 1) <= M,
                    "split pass 1: stage-A radius %d does not fit the frame", R1y);
        g.J1y = ceil_div(2 * R1y + 1, S);
        GPA_REQUIRE(S * g.J1y + kAhead <= kMaxTaps, "split pass 1: stage-A filter too long");
        g.MdE = g.Md + 2 * Hy;
        g.pitch_e = (int)align_up((size_t)g.MdE, 32);
        g.col_shift = R1y + S * Hy;
        g.rows_a = (int)align_up((size_t)g.n_rows_filled, 32);
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
    if (g.R1y > 0) {
        g.a_y = a.take<float2>((size_t)2 * g.rows_a * g.pitch_e);
        g.carBy = a.take<float2>((size_t)g.n_planes * g.MdE);
        g.derotBy = a.take<float2>((size_t)g.n_planes * g.Md);
        g.jBy = a.take<float2>((size_t)2 * g.n_planes);
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

// Split pass 1, anchor stage (once per call): the anchor plane n_planes / 2 filtered by G_1 at the full rate,
// decimated, with its carrier masked to the frame body (part 0) and to the wrapped columns (part 1).
template <int S, int W1>
static int launch_anchor_y_w(const MrGeometry& g, const float* img, const TapTable& t1y, cudaStream_t st) {
    MrPass1Params p;
    p.img = img; p.phy = g.phy; p.p1 = g.a_y; p.plane_stride = (size_t)g.rows_a * g.pitch_e;
    p.N = g.N; p.M = g.M; p.Md = g.MdE; p.pitch_d = g.pitch_e; p.n_rows_filled = g.n_rows_filled;
    p.Rax = g.row_shift; p.Ray = g.col_shift; p.J = g.J1y; p.plane0 = g.n_planes / 2; p.pstep = 0;
    p.count = 2; p.planes_per_cta = 2;
    const size_t n_samp1 = (size_t)S * (W1 * kP + g.J1y + kAhead + 1);
    const size_t smem = (n_samp1 * 33 + 1) * sizeof(float) + 2 * n_samp1 * sizeof(float2);
    GPA_REQUIRE(smem <= 227 * 1024, "split pass 1: stage-A filter too long for shared memory (%zu bytes)", smem);
    GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_pass1<S, W1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dim3 grid(ceil_div(g.n_rows_filled, 32), ceil_div(g.pitch_e, W1 * kP), 1);
    KernelTimer timer("k_mr_pass1a", st);
    k_mr_pass1<S, W1, true><<<grid, W1 * 32, smem, st>>>(p, t1y);
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

// The anchor stage is ONE plane per peak: a latency-bound launch of a few hundred CTAs (two per SM), replicated on every
// rank of a sharded sweep.  The column tile W1 * 16 is chosen among 8 / 9 / 10 / 12 warps so that the grid needs the fewest
// waves (C3: 69 x 5 = 345 CTAs = 2 waves with 8 warps, 69 x 4 = 276 = 1 wave with 9: 61 -> 3x us per peak).
template <int S>
static int launch_anchor_y(const MrGeometry& g, const float* img, const TapTable& t1y, cudaStream_t st) {
    const int rows = ceil_div(g.n_rows_filled, 32), slots = 2 * 148;
    int best_w = 8;
    long best_cost = -1;
    for (int w : {8, 9, 10, 12}) {
        const long ctas = (long)rows * ceil_div(g.pitch_e, w * kP);
        const long cost = ((ctas + slots - 1) / slots) * 100000L * w + ctas;     // waves x CTA length, then fewer CTAs
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_w = w; }
    }
    switch (best_w) {
        case 9: return launch_anchor_y_w<S, 9>(g, img, t1y, st);
        case 10: return launch_anchor_y_w<S, 10>(g, img, t1y, st);
        case 12: return launch_anchor_y_w<S, 12>(g, img, t1y, st);
        default: return launch_anchor_y_w<S, 8>(g, img, t1y, st);
    }
}

template <int S>
static int launch_mr(const MrGeometry& g, const float* img, const TapTable& ty, const TapTable& tx, const TapTable& tb,
                     const TapTable& t2, const TapTable& t2y, int plane0, int pstep, int count, int cand_mode,
                     unsigned long long* key, cudaStream_t st, bool whole_share) {
    if (g.R1y > 0) {   // stage 1, split: coarse-rate stage per plane from the anchor plane (launch_anchor_y ran once per call)
        MrPass1bParams p;
        p.A = g.a_y; p.a_part = (size_t)g.rows_a * g.pitch_e; p.carB = g.carBy; p.derotB = g.derotBy; p.jB = g.jBy;
        p.p1 = g.p1; p.plane_stride = g.plane_stride;
        p.n_rows = g.n_rows_filled; p.Md = g.Md; p.MdE = g.MdE; p.pitch_d = g.pitch_d; p.pitch_e = g.pitch_e;
        p.H = g.Hy; p.EB = ceil_div(g.R1y, S) + 1; p.plane0 = plane0; p.pstep = pstep; p.count = count;
        const int JB = 2 * g.Hy + 1;
        const int TR = kWarps * kP + JB - 1;
        const size_t smem = (size_t)(2 * TR * 33 + 2 * TR + 2 * kWarps * kP) * sizeof(float2);
        const int tiles = ceil_div(g.n_rows_filled, 32) * ceil_div(g.pitch_d, kWarps * kP);
        int zs = ceil_div(2 * 296, tiles);                 // >= ~2 waves of CTAs
        if (zs > count) zs = count;
        if (zs < 1) zs = 1;
        p.planes_per_cta = ceil_div(count, zs);
        dim3 grid(ceil_div(g.n_rows_filled, 32), ceil_div(g.pitch_d, kWarps * kP), ceil_div(count, p.planes_per_cta));
        KernelTimer timer("k_mr_pass1b", st);
#define GPA_P1B(JBV)                                                                                                    \
    case JBV:                                                                                                           \
        GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_pass1b<JBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
        k_mr_pass1b<JBV><<<grid, kWarps * 32, smem, st>>>(p, t2y);                                                      \
        break;
        switch (JB) {
            GPA_P1B(13) GPA_P1B(15) GPA_P1B(17) GPA_P1B(19) GPA_P1B(21) GPA_P1B(23)
            default:
                set_error("split pass 1: unsupported coarse tap count %d", JB);
                return GPA_ERR_INVALID;
        }
#undef GPA_P1B
    } else {   // stage 1
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
        // a small share of the planes (k-grid sharded over GPUs): one CTA per tile takes all of them, so the image tile
        // fill (about as expensive as one plane's filter) is paid once instead of once per plane
        if (count <= 8 && tiles >= 200) ppc = count;
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
        // few planes (a rank's share of a sharded sweep): split the candidates of a (tile, plane) over several CTAs so that
        // the grid is >= ~4 waves of the 2 x 148 CTA slots instead of e.g. 320 CTAs = 1.08 waves running as 2
        const int tiles2b = (g.pitch_d / kLanes) * ceil_div(g.Nd, kWarps * kP);
        int c_split = ceil_div(4 * 296, tiles2b * count);
        if (c_split > g.n_cand / 8) c_split = g.n_cand / 8;      // keep >= 8 candidates per staged tile
        if (c_split < 1) c_split = 1;
        p.c_split = c_split;
        dim3 grid(g.pitch_d / kLanes, ceil_div(g.Nd, kWarps * kP), count * c_split);
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
        p.n_hint = p.prune ? g_gossip.n : 0;
        p.epoch = g_gossip.epoch;
        for (int r = 0; r < GPA_MAX_PEERS; ++r) p.hint[r] = r < g_gossip.n ? g_gossip.hint[r] : nullptr;
        // two-phase sharded sweep: only when this launch covers the rank's whole share (one chunk)
        const bool two_phase = p.prune && p.n_hint > 1 && g_gossip.two_phase && whole_share;
        p.phase = 0; p.best_all = nullptr; p.rank = g_gossip.rank; p.world = g_gossip.n; p.z0 = 0;
        if (p.prune) {
            dim3 tg(ceil_div(g.M, kMrTY), ceil_div(g.N, kMrTX));
            OrderShare share;
            share.n = two_phase ? g_gossip.n : 0;
            share.rank = g_gossip.rank;
            for (int r = 0; r < GPA_MAX_PEERS; ++r) share.best[r] = r < share.n ? g_gossip.best[r] : nullptr;
            KernelTimer timer("k_mr_order", st);
            k_mr_order<S><<<tg, 256, 0, st>>>(g.pmax, g.n_cand, count, p.nbx, p.nby, g.nbx_alloc, g.nby_alloc, g.perm, share);
        }
        if (cand_mode == GPA_CAND_GRID) { p.idx_c = g.n_planes; p.idx_p = 1; } else { p.idx_c = 0; p.idx_p = 1; }
        constexpr int CX = kMrTX / S + kMrW - 2, CY = kMrTY / S + kMrW - 2;
        constexpr size_t ctile = ((size_t)CX * (CY + 2) * 8 + 127) / 128 * 128;  // bytes per coarse buffer (TMA box: CY + 2 columns, 128-byte aligned)
        const size_t smem = 2 * ctile + (size_t)(2 * CY * (kMrTX + 1)) * sizeof(float2);
        // P2 as a 3-D tensor (2 Md floats, Nd, candidates x planes) for the TMA box loads of the interior tiles
        CUtensorMap tmap;
        std::memset(&tmap, 0, sizeof(tmap));
        p.use_tma = 0;
        if (g_tma_enabled && (size_t)g.Md * 8 % 16 == 0 && (long long)count * g.n_cand < (1LL << 31)) {
            static PFN_encodeTiled encode = load_encode_tiled();
            if (encode != nullptr) {
                const cuuint64_t dims[3] = {(cuuint64_t)2 * g.Md, (cuuint64_t)g.Nd, (cuuint64_t)count * g.n_cand};
                const cuuint64_t strides[2] = {(cuuint64_t)g.Md * 8, (cuuint64_t)g.Nd * g.Md * 8};
                const cuuint32_t box[3] = {2 * (CY + 2), CX, 1};
                const cuuint32_t estr[3] = {1, 1, 1};
                const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)g.p2, dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                p.use_tma = cr == CUDA_SUCCESS;
            }
        }
        dim3 grid(ceil_div(g.M, kMrTY), ceil_div(g.N, kMrTX), count);
        auto launch_interp = [&](dim3 gr) -> int {
            KernelTimer timer("k_mr_interp", st);
            if (g.n_cand <= 256) {
                GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_interp<S, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                k_mr_interp<S, 8><<<gr, 256, smem, st>>>(p, tb, tmap);
            } else {
                GPA_CHECK_CUDA(cudaFuncSetAttribute(k_mr_interp<S, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                k_mr_interp<S, 16><<<gr, 256, smem, st>>>(p, tb, tmap);
            }
            return GPA_OK;
        };
        int rc = GPA_OK;
        if (two_phase) {
            // barrier 1: every rank has published the best bounds of its tiles (k_mr_order above)
            if ((rc = gpa_peer_signal(g_gossip.flag_a, g_gossip.n, g_gossip.epoch64, st)) ||
                (rc = gpa_peer_wait(g_gossip.wait_a, g_gossip.n, g_gossip.epoch64, g_gossip.timeout_s, g_gossip.status, st)))
                return rc;
            p.best_all = g_gossip.best[0];          // entry 0 is this rank's own table
            p.phase = 1;
            if ((rc = launch_interp(dim3(grid.x, grid.y, 1)))) return rc;
            // barrier 2: the globally most promising plane of every tile has been swept and its bounds published
            if ((rc = gpa_peer_signal(g_gossip.flag_b, g_gossip.n, g_gossip.epoch64, st)) ||
                (rc = gpa_peer_wait(g_gossip.wait_b, g_gossip.n, g_gossip.epoch64, g_gossip.timeout_s, g_gossip.status, st)))
                return rc;
            p.phase = 2;
        }
        if ((rc = launch_interp(grid))) return rc;
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_sweep_arm_gossip(void* const* hint_ptrs, int n_ranks, unsigned int epoch) {
    GPA_REQUIRE(n_ranks >= 0 && n_ranks <= GPA_MAX_PEERS && (n_ranks == 0 || hint_ptrs), "bad argument");
    for (int r = 0; r < n_ranks; ++r) {
        GPA_REQUIRE(hint_ptrs[r] != nullptr, "null hint array");
        g_gossip.hint[r] = static_cast<unsigned long long*>(hint_ptrs[r]);
    }
    g_gossip.n = n_ranks;
    g_gossip.epoch = epoch;
    g_gossip.two_phase = false;
    return GPA_OK;
}

extern "C" int gpa_sweep_arm_two_phase(void* const* best_ptrs, void* const* flag_slots_a, void* const* flag_slots_b,
                                       const unsigned long long* wait_a, const unsigned long long* wait_b, int rank,
                                       unsigned long long epoch, double timeout_s, int* status) {
    GPA_REQUIRE(g_gossip.n > 1, "gpa_sweep_arm_two_phase follows gpa_sweep_arm_gossip with n_ranks > 1");
    GPA_REQUIRE(best_ptrs && flag_slots_a && flag_slots_b && wait_a && wait_b && status, "null pointer argument");
    GPA_REQUIRE(rank >= 0 && rank < g_gossip.n && timeout_s > 0, "bad rank / timeout");
    for (int r = 0; r < g_gossip.n; ++r) {
        GPA_REQUIRE(best_ptrs[r] && flag_slots_a[r] && flag_slots_b[r], "null table / flag slot");
        g_gossip.best[r] = static_cast<float*>(best_ptrs[r]);
        g_gossip.flag_a[r] = flag_slots_a[r];
        g_gossip.flag_b[r] = flag_slots_b[r];
    }
    g_gossip.wait_a = wait_a; g_gossip.wait_b = wait_b; g_gossip.rank = rank; g_gossip.epoch64 = epoch;
    g_gossip.timeout_s = timeout_s; g_gossip.status = status;
    g_gossip.two_phase = true;
    return GPA_OK;
}

extern "C" int gpa_set_tma(int on) {
    g_tma_enabled = on != 0;
    return GPA_OK;
}

extern "C" int gpa_set_pruning(int on) {
    g_prune_enabled = on != 0;
    return GPA_OK;
}

extern "C" int gpa_sweep_mr_workspace_bytes(int N, int M, int n_rows, int n_planes, int cand_mode, int S, int Rax,
                                            int Ray, int Rb, int R1x, int H2x, int R1y, int H2y, int planes_in_flight,
                                            size_t* bytes) {
    MrGeometry g;
    int rc = plan_mr(g, N, M, n_rows, n_planes, cand_mode, S, Rax, Ray, Rb, R1x, H2x, R1y, H2y);
    if (rc) return rc;
    GPA_REQUIRE(bytes != nullptr, "bytes is null");
    GPA_REQUIRE(planes_in_flight >= 1 && planes_in_flight <= n_planes, "planes_in_flight out of range");
    *bytes = carve_mr(g, nullptr, 0, planes_in_flight) + (size_t)planes_in_flight * 1024 + 8192;
    return GPA_OK;
}

extern "C" int gpa_sweep_argmax_mr(const float* img, int N, int M, const double* wx_rows, int n_rows,
                                   const double* wy_planes, int n_planes, int cand_mode, int plane_begin,
                                   int plane_end, int plane_step, int S, const float* taps_ax, int Rax, const float* taps_ay, int Ray,
                                   const float* taps_bx, const float* taps_by, int Rb, const float* taps_1x, int R1x,
                                   const float* taps_2x, int H2x, double sigma_a, double sigma_1,
                                   const float* taps_1y, int R1y, const float* taps_2y, int H2y, double sigma_1y,
                                   unsigned long long* key, void* ws, size_t ws_bytes, void* stream) {
    MrGeometry g;
    int rc = plan_mr(g, N, M, n_rows, n_planes, cand_mode, S, Rax, Ray, Rb, R1x, H2x, R1y, H2y);
    if (rc) return rc;
    GPA_REQUIRE(R1y == 0 || (taps_1y && taps_2y && sigma_1y > 0.0 && sigma_1y < sigma_a),
                "split pass 1 needs both tap sets and 0 < sigma_1y < sigma_a");
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
    TapTable t1y, t2y;
    std::memset(&t2y, 0, sizeof(t2y));
    if (g.R1y > 0) {
        if ((rc = fill_polyphase(t1y, taps_1y, R1y, S, g.J1y))) return rc;
        for (int j = 0; j < 2 * H2y + 1; ++j) t2y.g[j] = make_float2(taps_2y[j], taps_2y[j]);
    }
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
    if (g.R1y > 0) {   // split pass 1: per-plane coarse carriers relative to the anchor plane, then the anchor stage
        SplitTabParams tp;
        tp.phx1 = nullptr; tp.carB = g.carBy; tp.derotB = g.derotBy; tp.jB = g.jBy; tp.wx_d = g.wy_d;
        tp.wx0 = wy_planes[n_planes / 2];
        const double s2sq = sigma_a * sigma_a - sigma_1y * sigma_1y;
        tp.ratio = sigma_a * sigma_a / s2sq;
        tp.cexp = 2.0 * 9.869604401089358 * sigma_a * sigma_a * sigma_1y * sigma_1y / s2sq;
        tp.n_cand = n_planes; tp.N = M; tp.S = S; tp.H = g.Hy; tp.Nd = g.Md; tp.NdE = g.MdE; tp.n_alloc = 0;
        tp.Rtot = g.col_shift;
        dim3 grid(ceil_div(g.MdE, 256) < 8 ? ceil_div(g.MdE, 256) : 8, n_planes);     // no (n_cand + 1)-th block: phx1 is not built
        k_build_split_tables<<<grid, 256, 0, st>>>(tp);
        GPA_CHECK_CUDA(cudaGetLastError());
        if (S == 2) rc = launch_anchor_y<2>(g, img, t1y, st);
        else if (S == 4) rc = launch_anchor_y<4>(g, img, t1y, st);
        else rc = launch_anchor_y<8>(g, img, t1y, st);
        if (rc) return rc;
    }
    for (int i0 = 0; i0 < total; i0 += chunk) {
        const int cnt = total - i0 < chunk ? total - i0 : chunk;
        const int p0 = plane_begin + i0 * plane_step;
        const bool whole = cnt == total;
        if (S == 2) rc = launch_mr<2>(g, img, ty, tx, tb, t2, t2y, p0, plane_step, cnt, cand_mode, key, st, whole);
        else if (S == 4) rc = launch_mr<4>(g, img, ty, tx, tb, t2, t2y, p0, plane_step, cnt, cand_mode, key, st, whole);
        else rc = launch_mr<8>(g, img, ty, tx, tb, t2, t2y, p0, plane_step, cnt, cand_mode, key, st, whole);
        if (rc) break;
    }
    g_gossip.n = 0;      // one call per arming
    g_gossip.two_phase = false;
    return rc;
}

static int finalize_mr(int N, int M, const double* wx_rows, int n_rows, const double* wy_planes, int n_planes, int cand_mode,
                       int plane_begin, int plane_end, int plane_step, int S, int Rax, int Ray, const float* taps_bx,
                       const float* taps_by, int Rb, int R1x, int H2x, int R1y, int H2y, const unsigned long long* key, double kref_x,
                       double kref_y, int grad_mode, int out_f64, void* lockin, void* grad, void* w, int* kidx,
                       void* const* lockin_dst, void* const* grad_dst, int n_dst, int dst_rows, int write_zero, void* ws,
                       size_t ws_bytes, void* stream) {
    MrGeometry g;
    int rc = plan_mr(g, N, M, n_rows, n_planes, cand_mode, S, Rax, Ray, Rb, R1x, H2x, R1y, H2y);
    if (rc) return rc;
    GPA_REQUIRE(wx_rows && wy_planes && ws && key && (lockin || n_dst > 0), "null pointer argument");
    GPA_REQUIRE(0 <= plane_begin && plane_begin <= plane_end && plane_end <= n_planes, "bad plane range");
    GPA_REQUIRE(grad_mode == GPA_GRAD_CENTRAL || grad_mode == GPA_GRAD_FORWARD || grad_mode == GPA_GRAD_NONE,
                "bad grad_mode %d", grad_mode);
    GPA_REQUIRE(grad_mode == GPA_GRAD_NONE || grad != nullptr || n_dst > 0, "grad is null but a gradient was requested");
    GPA_REQUIRE(plane_step >= 1, "plane_step must be >= 1");
    GPA_REQUIRE(n_dst >= 0 && n_dst <= GPA_MAX_PEERS, "n_dst out of range");
    if (n_dst > 0) {
        GPA_REQUIRE(lockin_dst && (grad_mode == GPA_GRAD_NONE || grad_dst), "null destination table");
        GPA_REQUIRE(n_dst == 1 || (dst_rows >= 1 && (long long)dst_rows * n_dst >= N), "dst_rows * n_dst must cover the frame");
        for (int d = 0; d < n_dst; ++d)
            GPA_REQUIRE(lockin_dst[d] && (grad_mode == GPA_GRAD_NONE || grad_dst[d]), "null destination pointer");
    }
    const bool nothing = plane_begin == plane_end;
    if (nothing && !(n_dst > 0 && write_zero)) return GPA_OK;
    const int total = nothing ? 0 : ceil_div(plane_end - plane_begin, plane_step);
    if (!nothing) {
        const int chunk = fit_chunk_mr(g, ws, ws_bytes, total);
        if (chunk != total) {
            set_error("gpa_sweep_finalize_mr needs every plane of the range resident (workspace holds %d of %d)", chunk, total);
            return GPA_ERR_WORKSPACE;
        }
    } else {
        carve_mr(g, ws, ws_bytes, 0);
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
    f.n_dst = n_dst; f.dst_rows = dst_rows > 0 ? dst_rows : N; f.write_zero = write_zero;
    for (int d = 0; d < n_dst; ++d) {
        f.lockin_dst[d] = lockin_dst[d];
        f.grad_dst[d] = grad_mode == GPA_GRAD_NONE ? nullptr : grad_dst[d];
    }
    mp.p2 = g.p2; mp.Nd = g.Nd; mp.Md = g.Md; mp.n_cand = g.n_cand; mp.S = S; mp.pstep = plane_step;
    KernelTimer timer("k_mr_finalize", st);
    if (n_dst > 0) {   // a rank's share of the planes: CTA-level compaction of the owned pixels, owner-writes
        // tile rows per CTA: ~256 owned pixels per 256-thread CTA (a rank owns about total / n_planes of the pixels)
        // (measured on a C3 share: 32 rows beat 16 from a third of the planes down — 0.62 -> 0.58 ms at 1/4 — and 64 rows lose
        // again at 1/8, 0.45 -> 0.53 ms: too few CTAs)
        const int fx = (!nothing && (long long)total * 3 <= n_planes) ? 32 : 16;
        dim3 grid(ceil_div(M, 64), ceil_div(N, fx));
#define GPA_MRFINS2(SS, TT)                                                                       \
        do {                                                                                          \
            if (fx == 32) k_mr_finalize_sharded<SS, TT, 32><<<grid, 256, 0, st>>>(mp, tb);            \
            else k_mr_finalize_sharded<SS, TT, 16><<<grid, 256, 0, st>>>(mp, tb);                     \
        } while (0)
#define GPA_MRFINS(SS)                          \
        if (out_f64) GPA_MRFINS2(SS, double2);      \
        else GPA_MRFINS2(SS, float2)
        if (S == 2) { GPA_MRFINS(2); } else if (S == 4) { GPA_MRFINS(4); } else { GPA_MRFINS(8); }
#undef GPA_MRFINS2
#undef GPA_MRFINS
        GPA_CHECK_CUDA(cudaGetLastError());
        return GPA_OK;
    }
    // pixels per warp: no compaction needed when every plane is ours, wider spans for smaller shares
    const bool all_planes = plane_begin == 0 && plane_end == n_planes && plane_step == 1;
    const int span = all_planes ? 32 : 128;      // measured on 2 GPUs (C3): 1.67 ms without compaction, 1.70 at 64, 1.49 at 128
    dim3 grid(ceil_div(M, span), ceil_div(N, 8));
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

// Finalize from the coarse grids left in the workspace by gpa_sweep_argmax_mr.  Requires that call to
// have covered exactly [plane_begin, plane_end) with all its planes resident (same ws, same arguments).
extern "C" int gpa_sweep_finalize_mr(int N, int M, const double* wx_rows, int n_rows, const double* wy_planes,
                                     int n_planes, int cand_mode, int plane_begin, int plane_end, int plane_step,
                                     int S, int Rax, int Ray, const float* taps_bx, const float* taps_by, int Rb,
                                     int R1x, int H2x, int R1y, int H2y, const unsigned long long* key, double kref_x, double kref_y,
                                     int grad_mode, int out_f64, void* lockin, void* grad, void* w, int* kidx, void* ws,
                                     size_t ws_bytes, void* stream) {
    GPA_REQUIRE(lockin != nullptr, "null pointer argument");
    return finalize_mr(N, M, wx_rows, n_rows, wy_planes, n_planes, cand_mode, plane_begin, plane_end, plane_step, S, Rax, Ray,
                       taps_bx, taps_by, Rb, R1x, H2x, R1y, H2y, key, kref_x, kref_y, grad_mode, out_f64, lockin, grad, w, kidx,
                       nullptr, nullptr, 0, 0, 0, ws, ws_bytes, stream);
}

// Owner-writes variant for a rank's share of the planes (k-grid sharded over GPUs, peer.cu).
extern "C" int gpa_sweep_finalize_mr_sharded(int N, int M, const double* wx_rows, int n_rows, const double* wy_planes,
                                             int n_planes, int cand_mode, int plane_begin, int plane_end, int plane_step,
                                             int S, int Rax, int Ray, const float* taps_bx, const float* taps_by, int Rb,
                                             int R1x, int H2x, int R1y, int H2y, const unsigned long long* key, double kref_x, double kref_y,
                                             int grad_mode, int out_f64, void* const* lockin_dst, void* const* grad_dst,
                                             int n_dst, int dst_rows, int write_zero, void* ws, size_t ws_bytes,
                                             void* stream) {
    GPA_REQUIRE(n_dst >= 1, "n_dst must be >= 1");
    return finalize_mr(N, M, wx_rows, n_rows, wy_planes, n_planes, cand_mode, plane_begin, plane_end, plane_step, S, Rax, Ray,
                       taps_bx, taps_by, Rb, R1x, H2x, R1y, H2y, key, kref_x, kref_y, grad_mode, out_f64, nullptr, nullptr, nullptr,
                       nullptr, lockin_dst, grad_dst, n_dst, dst_rows, write_zero, ws, ws_bytes, stream);
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
        std::memset(&f, 0, sizeof(f));
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
    GPA_CHECK_CUDA(cudaFuncSetAttribute(k_pass2_seq, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
