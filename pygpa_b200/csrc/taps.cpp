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
        if (sigma_a) *sigma_a = sa;
        if (sigma_b) *sigma_b = sb;
        if (Ra) *Ra = ra;
        if (Rb) *Rb = rb;
        return s;
    }
    return 0;
}
