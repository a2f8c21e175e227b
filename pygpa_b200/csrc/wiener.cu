// K8 — gaussian_deconvolve: Wiener deconvolution of the displacement field with the lock-in
// Gaussian (float64), sm_100a.  SURVEY 8(f) row 3.
//
// Reference semantics: gaussian_deconvolve (pyGPA/geometric_phase_analysis.py:892-904): reflect-pad
// every plane by 2 dr, skimage.restoration.wiener with the Gaussian PSF and balance = 5000, crop.
// For that PSF the Wiener filter has a closed form on the DFT grid of the padded (P0 x P1) frame,
//     W(f) = H / (H^2 + balance L^2),  H = exp(-2 pi^2 sigma^2 |f|^2),  L = 4 - 2 cos 2 pi fx - 2 cos 2 pi fy,
// so the whole operation is  crop( IDFT2( W . DFT2( pad(u) ) ) ).
//
// The padded sizes (N + 4 dr) are not powers of two, so every 1-D DFT of length P runs as a
// Bluestein chirp-z transform inside shared memory: x[n] c[n] -> zero-pad to L >= 2P-1 (a power of
// two) -> FFT_L -> times the precomputed FFT of the chirp -> inverse FFT_L -> times c[k], with
// c[n] = exp(-i pi n^2 / P) and the radix-8 Stockham FFT of K2 (fft_device.cuh).  2-D = rows,
// transpose, rows; W is applied in the store of the second forward pass, the crop and the real part
// in the store of the last inverse pass.
#include "common.cuh"
#include "fft_device.cuh"

namespace gpa {

constexpr int kMaxBluesteinL = 8192;       // 128 KB of shared memory per row

struct BsArgs {
    const double2* in;       // rows of length P (stride in_stride elements)
    double2* out;            // rows of length P
    double* out_real;        // last pass: real part of the cropped window
    int rows, row0;          // rows processed, first input row
    int in_stride, out_stride;
    int inverse;             // 1: x = conj(DFT(conj X)) / P
    int apply_w;             // forward pass 2: multiply by W(k0 = k, k1 = row) (transposed layout)
    double sigma2, balance;  // 2 pi^2 sigma^2
    int P0, P1;              // padded frame (for the frequencies of W)
    int crop_lo, crop_n;     // last pass: columns [crop_lo, crop_lo + crop_n) -> out_real row of length crop_n
};

template <int MAXB>
__global__ void __launch_bounds__(512, 1) k_bs_rows(const AxisPlan ax, const BsArgs a) {
    extern __shared__ double2 buf[];
    const int P = ax.P, L = ax.L;
    const int row = a.row0 + blockIdx.x;
    const double2* __restrict__ x = a.in + (size_t)row * a.in_stride;
    for (int n = threadIdx.x; n < L; n += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (n < P) {
            v = x[n];
            if (a.inverse) v.y = -v.y;
            v = zmul(v, __ldg(ax.chirp + n));
        }
        buf[n] = v;
    }
    __syncthreads();
    fft_pow2<MAXB>(buf, L, ax.tw);
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        const double2 y = zmul(buf[k], __ldg(ax.bhat + k));
        buf[k] = make_double2(y.x, -y.y);                 // conj: the second forward FFT then inverts
    }
    __syncthreads();
    fft_pow2<MAXB>(buf, L, ax.tw);
    const double inv_l = 1.0 / (double)L;
    const double scale = a.inverse ? inv_l / (double)P : inv_l;
    if (a.out_real) {
        double* __restrict__ o = a.out_real + (size_t)blockIdx.x * a.crop_n;
        for (int j = threadIdx.x; j < a.crop_n; j += blockDim.x) {
            const int k = a.crop_lo + j;
            const double2 conv = make_double2(buf[k].x, -buf[k].y);
            const double2 v = zmul(conv, __ldg(ax.chirp + k));
            o[j] = v.x * scale;                           // real part (the conj of the inverse does not touch it)
        }
        return;
    }
    double2* __restrict__ o = a.out + (size_t)blockIdx.x * a.out_stride;
    for (int k = threadIdx.x; k < P; k += blockDim.x) {
        const double2 conv = make_double2(buf[k].x, -buf[k].y);
        double2 v = zmul(conv, __ldg(ax.chirp + k));
        v.x *= scale;
        v.y *= a.inverse ? -scale : scale;
        if (a.apply_w) {
            // transposed layout: this row is frequency index k1 = row of axis 1, k runs over axis 0
            const int k0 = k <= P / 2 ? k : k - P;                       // np.fft.fftfreq ordering
            const int k1 = row <= a.P1 / 2 ? row : row - a.P1;
            const double fx = (double)k0 / (double)a.P0, fy = (double)k1 / (double)a.P1;
            const double h = exp(-a.sigma2 * (fx * fx + fy * fy));
            const double lap = 4.0 - 2.0 * cospi(2.0 * fx) - 2.0 * cospi(2.0 * fy);
            const double w = h / (h * h + a.balance * lap * lap);
            v.x *= w;
            v.y *= w;
        }
        o[k] = v;
    }
}

// reflect padding (np.pad mode='reflect': no edge repeat) of one real plane into a complex frame
__global__ void __launch_bounds__(256) k_reflect_pad(const double* __restrict__ src, int N, int M, int pad, double2* __restrict__ dst) {
    const int P1 = M + 2 * pad;
    const size_t total = (size_t)(N + 2 * pad) * P1;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        int r = (int)(i / P1) - pad, c = (int)(i % P1) - pad;
        r = r < 0 ? -r : (r >= N ? 2 * (N - 1) - r : r);
        c = c < 0 ? -c : (c >= M ? 2 * (M - 1) - c : c);
        dst[i] = make_double2(src[(size_t)r * M + c], 0.0);
    }
}

__global__ void k_transpose_z(const double2* __restrict__ in, double2* __restrict__ out, int rows, int cols) {
    __shared__ double2 tile[32][33];
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

static size_t carve_wiener(AxisPlan (&ax)[2], double2*& A, double2*& B, void* ws, size_t ws_bytes, int P0, int P1) {
    Arena a(ws, ws_bytes);
    A = a.take<double2>((size_t)P0 * P1);
    B = a.take<double2>((size_t)P0 * P1);
    const int Ps[2] = {P0, P1};
    for (int i = 0; i < 2; ++i) {
        ax[i].P = Ps[i];
        ax[i].L = bs_pow2_at_least(2 * Ps[i] - 1);
        ax[i].chirp = a.take<double2>(Ps[i]);
        ax[i].bhat = a.take<double2>(ax[i].L);
        ax[i].tw = a.take<double2>(ax[i].L);
    }
    return a.off;
}

template <typename K1, typename K2>
static int launch_by_len(int L, K1 one, K2 two) {      // one / two radix-8 butterflies per thread
    const int eighth = L / 8 > 0 ? L / 8 : 1;
    const int threads = eighth < 32 ? 32 : (eighth > 512 ? 512 : eighth);
    return (eighth + threads - 1) / threads <= 1 ? one(threads) : two(threads);
}

static int bs_rows(const AxisPlan& ax, const BsArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)ax.L * sizeof(double2);
    return launch_by_len(
        ax.L,
        [&](int threads) -> int {
            GPA_CHECK_CUDA(cudaFuncSetAttribute(k_bs_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
            k_bs_rows<1><<<a.rows, threads, smem, st>>>(ax, a);
            return GPA_OK;
        },
        [&](int threads) -> int {
            GPA_CHECK_CUDA(cudaFuncSetAttribute(k_bs_rows<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
            k_bs_rows<2><<<a.rows, threads, smem, st>>>(ax, a);
            return GPA_OK;
        });
}

}  // namespace gpa

using namespace gpa;

extern "C" int gpa_deconvolve_workspace_bytes(int N, int M, int dr, size_t* bytes) {
    GPA_REQUIRE(bytes && N >= 2 && M >= 2 && dr >= 0, "bad argument");
    const int P0 = N + 4 * dr, P1 = M + 4 * dr;
    GPA_REQUIRE(2 * dr <= N - 1 && 2 * dr <= M - 1, "reflect padding of %d does not fit a %d x %d frame", 2 * dr, N, M);
    GPA_REQUIRE(bs_pow2_at_least(2 * P0 - 1) <= kMaxBluesteinL && bs_pow2_at_least(2 * P1 - 1) <= kMaxBluesteinL,
                "padded frame %d x %d exceeds the in-shared-memory transform (at most %d samples per axis)", P0, P1,
                kMaxBluesteinL / 2);
    AxisPlan ax[2];
    double2 *A, *B;
    *bytes = carve_wiener(ax, A, B, nullptr, 0, P0, P1) + 1024;
    return GPA_OK;
}

extern "C" int gpa_gaussian_deconvolve(const double* data /*(planes,N,M)*/, int planes, int N, int M, double sigma, int dr,
                                       double balance, double* out /*(planes,N,M)*/, void* ws, size_t ws_bytes, void* stream) {
    GPA_REQUIRE(data && out && ws && planes >= 1, "bad argument");
    size_t need = 0;
    int rc = gpa_deconvolve_workspace_bytes(N, M, dr, &need);
    if (rc) return rc;
    if (ws_bytes < need) {
        set_error("workspace too small (%zu < %zu)", ws_bytes, need);
        return GPA_ERR_WORKSPACE;
    }
    const int pad = 2 * dr, P0 = N + 2 * pad, P1 = M + 2 * pad;
    AxisPlan ax[2];
    double2 *A, *B;
    carve_wiener(ax, A, B, ws, ws_bytes, P0, P1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < 2; ++i) {
        k_bs_tables<<<ceil_div(ax[i].L, 256), 256, 0, st>>>(ax[i].chirp, ax[i].tw, ax[i].P, ax[i].L);
        const size_t smem = (size_t)ax[i].L * sizeof(double2);
        const AxisPlan axi = ax[i];
        rc = launch_by_len(
            axi.L,
            [&](int threads) -> int {
                GPA_CHECK_CUDA(cudaFuncSetAttribute(k_bs_prep<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
                k_bs_prep<1><<<1, threads, smem, st>>>(axi);
                return GPA_OK;
            },
            [&](int threads) -> int {
                GPA_CHECK_CUDA(cudaFuncSetAttribute(k_bs_prep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 136 * 1024));
                k_bs_prep<2><<<1, threads, smem, st>>>(axi);
                return GPA_OK;
            });
        if (rc) return rc;
    }
    const double pi = 3.141592653589793238462643383279;
    KernelTimer timer("gaussian_deconvolve", st);
    for (int pl = 0; pl < planes; ++pl) {
        size_t blocks = ((size_t)P0 * P1 + 1023) / 1024;
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_reflect_pad<<<(unsigned)blocks, 256, 0, st>>>(data + (size_t)pl * N * M, N, M, pad, A);          // A: (P0, P1)
        BsArgs a;
        std::memset(&a, 0, sizeof(a));
        a.P0 = P0; a.P1 = P1; a.sigma2 = 2.0 * pi * pi * sigma * sigma; a.balance = balance;
        a.in = A; a.out = B; a.rows = P0; a.in_stride = P1; a.out_stride = P1;                             // DFT along axis 1
        if ((rc = bs_rows(ax[1], a, st))) return rc;
        k_transpose_z<<<dim3(ceil_div(P1, 32), ceil_div(P0, 32)), dim3(32, 8), 0, st>>>(B, A, P0, P1);     // A: (P1, P0)
        a.in = A; a.out = B; a.rows = P1; a.in_stride = P0; a.out_stride = P0; a.apply_w = 1;              // DFT along axis 0, times W
        if ((rc = bs_rows(ax[0], a, st))) return rc;
        a.apply_w = 0; a.inverse = 1;
        a.in = B; a.out = A;                                                                               // inverse along axis 0
        if ((rc = bs_rows(ax[0], a, st))) return rc;
        k_transpose_z<<<dim3(ceil_div(P0, 32), ceil_div(P1, 32)), dim3(32, 8), 0, st>>>(A, B, P1, P0);     // B: (P0, P1)
        a.in = B; a.out = nullptr; a.out_real = out + (size_t)pl * N * M;                                  // inverse along axis 1,
        a.rows = N; a.row0 = pad; a.in_stride = P1; a.crop_lo = pad; a.crop_n = M;                         // cropped rows / columns
        if ((rc = bs_rows(ax[1], a, st))) return rc;
    }
    GPA_CHECK_CUDA(cudaGetLastError());
    return GPA_OK;
}
