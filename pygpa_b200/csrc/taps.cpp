// Host-side helpers of the C ABI: the filter taps and the multirate plan, so that a non-Python host
// (or the CuPy stub in INTEGRATION.md) does not have to re-derive them.  Pure C++, no CUDA.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/gpa_b200.h"

namespace gpa {
void set_error(const char* fmt, ...);
}

// taps[d + R] = real-space kernel of the reference's Fourier-domain Gaussian
// exp(-2 pi^2 sigma^2 f^2) (scipy.ndimage.fourier_gaussian; geometric_phase_analysis.py:44,75) on a
// circular axis of length n, for |d| <= R: (1/n) sum_k H(f_k) cos(2 pi k d / n).
extern "C" int gpa_gaussian_taps(int n, double sigma, int R, float* taps) {
    if (n < 1 || R < 0 || 2 * R + 1 > n || !taps || !(sigma >= 0.0)) {
        gpa::set_error("gpa_gaussian_taps: bad argument (n=%d, R=%d, sigma=%g)", n, R, sigma);
        return GPA_ERR_INVALID;
    }
    const double pi = 3.141592653589793238462643383279;
    std::vector<double> h(n);
    for (int k = 0; k < n; ++k) {
        const double f = (k <= (n - 1) / 2 ? k : k - n) / (double)n;     // np.fft.fftfreq
        h[k] = std::exp(-2.0 * pi * pi * sigma * sigma * f * f);
    }
    // the transfer function is negligible beyond |f| ~ 1.4/sigma: only those terms are summed
    int kmax = n / 2;
    for (int k = 1; k <= n / 2; ++k)
        if (h[k] < 1e-40) { kmax = k; break; }
    for (int d = 0; d <= R; ++d) {
        double acc = h[0];
        for (int k = 1; k <= kmax && k < n; ++k) {
            const double c = std::cos(2.0 * pi * (double)((long long)k * d % n) / n);
            acc += h[k] * c;
            if (n - k != k && n - k > kmax) acc += h[n - k] * c;          // mirror frequency (cos is even)
        }
        const float v = (float)(acc / n);
        taps[R + d] = v;
        taps[R - d] = v;
    }
    return GPA_OK;
}

extern "C" int gpa_default_radius(int n, double sigma, double trunc) {
    int r = (int)std::ceil(trunc * sigma);
    const int cap = (n - 1) / 2;
    if (r > cap) r = cap;
    return r < 0 ? 0 : r;
}

// Shared-memory budget of the decimating pass-2 kernels (lockin.cu, launch_mr): S (W2 kP + J + kAhead + 1) fine rows
// x (32 columns + two carrier buffers per warp group) of float2 must fit 227 KB.
static bool pass2_tile_fits(int s, int j) {
    const int w2 = s == 8 ? 4 : 8;
    return (long long)s * (w2 * 16 + j + 3) * (32 + 4) * 8 <= 227 * 1024;
}

// Parameters of the multirate sweep (same rule as pygpa_b200/_taps.py): returns the stride (2, 4 or 8)
// or 0 when the multirate form does not apply (use the direct form).
extern "C" int gpa_multirate_plan(int N, int M, double sigma, double* sigma_a, double* sigma_b, int* Ra, int* Rb) {
    const double trunc = 4.5;
    for (int s = 8; s >= 2; s /= 2) {
        if (N % s || M % s || N / s < 12 || M / s < 12) continue;
        double c = std::sqrt(0.2) * sigma / s;
        if (c > 1.1) c = 1.1;
        if (c < 1.0) continue;
        const double sb = c * s, sa = std::sqrt(sigma * sigma - sb * sb);
        const int rb = (int)std::ceil(trunc * sb);
        int ra = (int)std::ceil(trunc * sa);
        {   // even number of taps per phase (statically scheduled pass-2 kernels); extra taps widen the radius
            const int j0 = (2 * ra + 1 + s - 1) / s, jt = 2 * ((j0 + 1) / 2);
            const int ra2 = (s * jt - 1) / 2;
            if (ra2 > ra) ra = ra2;
        }
        if (rb > 5 * s) continue;
        if (2 * ra + 1 > (N < M ? N : M) || s * ((2 * ra + 1 + s - 1) / s) + 2 > 446) continue;
        if (!pass2_tile_fits(s, (2 * ra + 1 + s - 1) / s)) continue;     // large sigma: try the next smaller stride
        if (sigma_a) *sigma_a = sa;
        if (sigma_b) *sigma_b = sb;
        if (Ra) *Ra = ra;
        if (Rb) *Rb = rb;
        return s;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// split pass 2 planner (same search as pygpa_b200/_taps.py: _split_plan / split_taps)
// ---------------------------------------------------------------------------------------------
namespace {
// max_f | c G_1t(f) h_t(S (f + delta)) - G_a(f + dw) |  over nf frequencies in [-1/2, 1/2)
double split_error(int s, double sigma_a, double sigma_1, int r1, int h, double dw, int nf = 2048) {
    const double pi = 3.141592653589793238462643383279;
    const double s2sq = sigma_a * sigma_a - sigma_1 * sigma_1, sigma_2 = std::sqrt(s2sq);
    const double delta = dw * sigma_a * sigma_a / s2sq;
    const double log_c = 2.0 * pi * pi * dw * dw * sigma_a * sigma_a * sigma_1 * sigma_1 / s2sq;
    if (log_c > 2.0794415416798357) return HUGE_VAL;   // c > 8: the band leaves the anchor stage attenuated by 1/c and
                                                        // fp32 rounding noise would come back amplified by c
    const double c = std::exp(log_c);
    std::vector<double> g1(r1 + 1), h2(h + 1);
    for (int d = 0; d <= r1; ++d) g1[d] = std::exp(-(double)d * d / (2.0 * sigma_1 * sigma_1)) / (sigma_1 * std::sqrt(2.0 * pi));
    for (int m = 0; m <= h; ++m) h2[m] = s * std::exp(-(double)(s * m) * (s * m) / (2.0 * s2sq)) / (sigma_2 * std::sqrt(2.0 * pi));
    double worst = 0.0;
    for (int i = 0; i < nf; ++i) {
        const double f = (double)(i - nf / 2) / nf;
        double G1 = g1[0], H2 = h2[0];                       // both filters are even: cosine sums
        for (int d = 1; d <= r1; ++d) G1 += 2.0 * g1[d] * std::cos(2.0 * pi * f * d);
        for (int m = 1; m <= h; ++m) H2 += 2.0 * h2[m] * std::cos(2.0 * pi * s * (f + delta) * m);
        const double e = std::fabs(c * G1 * H2 - std::exp(-2.0 * pi * pi * sigma_a * sigma_a * (f + dw) * (f + dw)));
        if (e > worst) worst = e;
    }
    return worst;
}
}  // namespace

// Plan of the split pass 2 for an axis of length n, multirate stride `stride`, decimation sigma_a and the
// candidate axis wx_rows (anchor = wx_rows[n_rows / 2]).  Returns 1 and fills R1 (stage-A radius, 2 R1 + 1 fine
// taps in taps_1), H (stage-B coarse radius, 2 H + 1 taps S G_2(S m) in taps_2) and sigma_1, or returns 0 when
// no factorisation meets the 1.3e-6 transfer-function tolerance with at most 23 coarse taps (then pass
// R1x = 0 to gpa_sweep_argmax_mr).  taps_1 must hold 446 floats, taps_2 23.
extern "C" int gpa_split_plan(int n, int stride, double sigma_a, const double* wx_rows, int n_rows, int* R1, int* H,
                              double* sigma_1_out, float* taps_1, float* taps_2) {
    if (n < 1 || !(stride == 2 || stride == 4 || stride == 8) || !(sigma_a > 0.0) || !wx_rows || n_rows < 1 || !R1 || !H ||
        !sigma_1_out || !taps_1 || !taps_2) {
        gpa::set_error("gpa_split_plan: bad argument");
        return GPA_ERR_INVALID;
    }
    if (n_rows < 8) return 0;
    const double pi = 3.141592653589793238462643383279;
    const int s = stride;
    double dw_max = 0.0;
    for (int i = 0; i < n_rows; ++i) dw_max = std::fmax(dw_max, std::fabs(wx_rows[i] - wx_rows[n_rows / 2]));
    {   // quantised upwards exactly as the Python planner does (shared cache key there)
        const double q = std::exp2(std::floor(std::log2(std::fmax(dw_max, 1e-12))) - 4.0);
        dw_max = std::ceil(dw_max / q) * q;
    }
    struct Best { double err, s1, s2; int r1; };
    auto best_for = [&](int h) {
        Best b{-1.0, 0.0, 0.0, 0};
        const double top = std::fmin(h / 4.4, 2.4) + 1e-9;
        for (int step = 0;; ++step) {
            const double s2c = 1.35 + 0.05 * step;         // np.arange(1.35, top, 0.05)
            if (!(s2c < top)) break;
            const double sigma_2 = s2c * s;
            if (sigma_2 >= 0.98 * sigma_a) break;
            const double sigma_1 = std::sqrt(sigma_a * sigma_a - sigma_2 * sigma_2);
            int r1 = (int)std::ceil(6.0 * sigma_1);
            int j1 = (2 * r1 + 1 + s - 1) / s;
            j1 += j1 & 1;
            if (j1 < 18) j1 = 18;
            r1 = (s * j1 - 1) / 2;
            if (r1 + s * (h + 1) > n || n / s <= 2 * ((r1 + s - 1) / s + 1) || s * j1 + 2 > 446 || !pass2_tile_fits(s, j1)) continue;
            const double err = std::fmax(split_error(s, sigma_a, sigma_1, r1, h, dw_max), split_error(s, sigma_a, sigma_1, r1, h, 0.5 * dw_max));
            if (b.err < 0.0 || err < b.err) b = Best{err, sigma_1, sigma_2, r1};
        }
        return b;
    };
    auto emit = [&](int h, const Best& b) {
        *R1 = b.r1; *H = h; *sigma_1_out = b.s1;
        for (int d = -b.r1; d <= b.r1; ++d)
            taps_1[d + b.r1] = (float)(std::exp(-(double)d * d / (2.0 * b.s1 * b.s1)) / (b.s1 * std::sqrt(2.0 * pi)));
        for (int m = -h; m <= h; ++m)
            taps_2[m + h] = (float)(s * std::exp(-(double)(s * m) * (s * m) / (2.0 * b.s2 * b.s2)) / (b.s2 * std::sqrt(2.0 * pi)));
        return 1;
    };
    // the longest filter first: if even 23 coarse taps miss the tolerance nothing shorter is tried
    const Best longest = best_for(11);
    if (longest.err < 0.0 || longest.err > 1.3e-6) return 0;
    for (int h = 6; h <= 10; ++h) {
        const Best b = best_for(h);
        if (b.err >= 0.0 && b.err <= 1.3e-6) return emit(h, b);
    }
    return emit(11, longest);
}
