/*
 * gpa_b200.h — C ABI of libgpa_b200.so: the B200 (sm_100a) implementation of pyGPA's
 * adaptive-GPA hot path.  Plain pointers and sizes only; no torch / CuPy types.
 *
 * The reference (TAdeJong/pyGPA) is pure Python and has no FFI: its boundary for this
 * path is a set of module-level functions taking and returning NumPy arrays.  Each entry
 * point below names the reference function(s) whose arithmetic it replaces
 * (paths relative to the reference checkout).  the modules under pygpa_b200/ bind these with ctypes
 * and re-exports the reference signatures; INTEGRATION.md shows the stub a pyGPA
 * maintainer would add.
 *
 * Conventions
 *   - images are row-major (N, M) with axis 0 = "x" (pyGPA's convention); k-vectors in
 *     cycles/pixel, component 0 multiplies the axis-0 index.
 *   - every pointer is a DEVICE pointer unless the parameter comment says "host".
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls return
 *     after enqueueing (small host arrays are staged before return, so the caller may
 *     free them immediately).
 *   - return value: 0 on success, negative gpaStatus otherwise; gpa_last_error() gives the
 *     message of the last failure on the calling thread.
 *   - no hidden device allocation: scratch comes from the caller through (ws, ws_bytes),
 *     sized by the matching *_workspace_bytes().
 */
#ifndef GPA_B200_H
#define GPA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GPA_OK = 0,
    GPA_ERR_INVALID = -1,      /* bad argument                      */
    GPA_ERR_WORKSPACE = -2,    /* workspace too small               */
    GPA_ERR_CUDA = -3,         /* CUDA runtime error                */
    GPA_ERR_UNSUPPORTED = -4   /* valid request this build cannot serve */
} gpaStatus;

/* candidate layout of a sweep */
#define GPA_CAND_GRID 0   /* candidates = rows x planes, flat index ix*n_planes + iy (np.arange x np.arange,
                             geometric_phase_analysis.py:803-804: wx outer, wy inner)                        */
#define GPA_CAND_LIST 1   /* candidate i = (wx_rows[i], wy_planes[i]), flat index i (wfr3, :647-666)          */

/* phase-gradient flavour of wfr2_grad_opt */
#define GPA_GRAD_CENTRAL 0 /* np.gradient (geometric_phase_analysis.py:807; cuGPA.py:63-65)  */
#define GPA_GRAD_FORWARD 1 /* grad='diff' (cuGPA.py:58-62): forward difference, NaN in the last row/column */
#define GPA_GRAD_NONE 2    /* optwfr2 / wfr2 / wfr3: no gradient output */

const char* gpa_last_error(void);
int gpa_version(void);          /* 10000*major + 100*minor + patch */
int gpa_device_sm_count(void);  /* of the current device; <0 on error */

/* Per-kernel timing for benchmarks: when enabled every kernel launch of this library is
 * bracketed by CUDA events on its stream.  gpa_profile_read sums the recorded durations of
 * the kernel called `kernel` ("k_pass1", "k_pass2_argmax", "k_finalize", ...), synchronising
 * on them, and optionally clears the record. */
int gpa_profile_enable(int on);
int gpa_profile_read(const char* kernel, double* total_ms, int* launches, int reset);

/* Measured FP32 FMA-pipe peak of the current device in TFLOP/s (register-only fma.rn.f32x2 loop, best of 5,
 * CUDA events; synchronises).  bench.py uses it as the denominator of the K1 roofline.  ws: >= 1.5 MB. */
int gpa_fp32_peak_tflops(void* ws, size_t ws_bytes, double* tflops /*host, out*/, void* stream);

/* Staging helper for the host binding: the reference takes float64 images
 * (cp.asarray(image), cuGPA.py:52); the kernels read float32. */
int gpa_cast_f64_to_f32(const double* in, float* out, size_t n, void* stream);

/* Decode merged arg-max keys (see gpa_sweep_argmax): kidx[i] = flat candidate index, -1 where nothing won. */
int gpa_key_to_kidx(const unsigned long long* key, int* kidx, size_t n, void* stream);

/* Glue of extract_displacement_field (geometric_phase_analysis.py:922-926) kept on the device:
 * phases = angle(lockin), weights = |lockin| * (mask + eps) with mask = 1 on the interior
 * [border, N-border) x [border, M-border).  lockin is float2 (is_f64 = 0) or double2. */
int gpa_phase_weight(const void* lockin, int is_f64, int N, int M, int border, double eps,
                     double* phases, double* weights, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1 — spatial lock-in and the adaptive (windowed-Fourier-ridge) sweep.
 *
 * Replaces: GPA / optGPA / vecGPA (geometric_phase_analysis.py:20-89), cuGPA.cuGPA
 * (cuGPA.py:11-38) for the fixed reference; wfr2_grad_opt / optwfr2 / wfr2 / wfr3 /
 * wfr2_only_lockin (geometric_phase_analysis.py:583-813) and cuGPA.wfr2_grad_opt,
 * wfr2_grad_single, wfr2_only_lockin, wfr2_only_grad (cuGPA.py:41-202) for the sweep.
 *
 * The reference's FFT low-pass is a circular convolution with a periodised Gaussian; here
 * it is a separable, truncated, real-tap circular convolution of the demodulated samples:
 *   plane_iy(x', y) = sum_d taps_y[d+Ry] * img(x', y+d) * exp(2 pi i wy (y+d))        (pass 1)
 *   sf(x, y)        = sum_d taps_x[d+Rx] * exp(2 pi i wx (x+d)) * plane_iy(x+d, y)    (pass 2)
 * indices wrapped into the frame BEFORE the carrier is evaluated (as the reference does).
 * The caller supplies the taps (host float arrays of 2R+1 entries; the Python host derives
 * them from scipy's exact transfer function, see pygpa_b200/_taps.py).
 * ------------------------------------------------------------------------------------------ */

/* Host-side helpers (no CUDA): the taps of the reference's Gaussian low-pass on a circular axis of
 * length n (real-space kernel of exp(-2 pi^2 sigma^2 f^2), |d| <= R), the default truncation radius
 * min(ceil(trunc * sigma), (n-1)/2), and the multirate plan: returns the stride (2, 4, 8; 0 = use the
 * direct form) and sigma_a, sigma_b, Ra = ceil(4.5 sigma_a), Rb = ceil(4.5 sigma_b); the taps for
 * gpa_sweep_argmax_mr are then gpa_gaussian_taps(N, sigma_a, Ra), (M, sigma_a, Ra), (N, sigma_b, Rb),
 * (M, sigma_b, Rb). */
int gpa_gaussian_taps(int n, double sigma, int R, float* taps /*host, 2R+1*/);
int gpa_default_radius(int n, double sigma, double trunc);
int gpa_multirate_plan(int N, int M, double sigma, double* sigma_a, double* sigma_b, int* Ra, int* Rb);

/* Scratch for a sweep / fixed lock-in over `n_rows` axis-0 carriers and `n_planes` first-pass
 * planes, keeping `planes_in_flight` (1..n_planes) planes resident at a time.  Keeping all
 * planes resident avoids recomputing pass 1 in gpa_sweep_finalize. */
int gpa_lockin_workspace_bytes(int N, int M, int n_rows, int n_planes, int Rx, int Ry,
                               int planes_in_flight, size_t* bytes);

/* Output precision: the arithmetic is fp32 either way; out_f64 != 0 widens on store so the
 * host binding can hand back the reference's float64 / complex128 arrays without a CPU pass. */

/* Fixed-reference lock-in, one k-vector: out[x*M+y] = (re, im) of the complex lock-in signal
 * (float2, or double2 when out_f64).
 * Reference: optGPA (geometric_phase_analysis.py:48-76), cuGPA (cuGPA.py:11-38). */
int gpa_lockin_fixed(const float* img, int N, int M, double kx, double ky,
                     const float* taps_x /*host*/, int Rx, const float* taps_y /*host*/, int Ry,
                     int out_f64, void* out, void* ws, size_t ws_bytes, void* stream);

/* Running arg-max of the sweep over planes [plane_begin, plane_end):
 *   key[x*M+y] = max(key, (float_bits(|sf|^2) << 32) | (0xFFFFFFFF - flat_index))   (64-bit atomic max)
 * so the largest amplitude wins and, on an exact tie, the LOWEST flat index — the
 * reference's strict-'>' first-wins rule (geometric_phase_analysis.py:806).  Candidates with
 * |sf| == 0 never win.  The caller zero-initialises key; keys from disjoint plane ranges
 * (other calls, other GPUs) merge with a plain integer max. */
int gpa_sweep_argmax(const float* img, int N, int M,
                     const double* wx_rows /*host*/, int n_rows,
                     const double* wy_planes /*host*/, int n_planes, int cand_mode,
                     int plane_begin, int plane_end,
                     const float* taps_x /*host*/, int Rx, const float* taps_y /*host*/, int Ry,
                     unsigned long long* key, void* ws, size_t ws_bytes, void* stream);

/* Multirate form of gpa_sweep_argmax (same key, same merge rules).  The Gaussian is factorised,
 * G_sigma = G_a * G_b (sigma_a^2 + sigma_b^2 = sigma^2): G_a is applied with decimation by
 * `stride` (2, 4 or 8) per axis and G_b as a `stride`-fold interpolator, which cuts the taps per
 * candidate about five-fold; the aliasing this introduces is below 1e-7 of the signal for the
 * strides the host chooses (pygpa_b200/_taps.py), far inside the near-tie budget.  Only the
 * arg-max decision uses these amplitudes: gpa_sweep_finalize recomputes the winner in the direct
 * form.  taps_a*: decimation filter (2 Ra + 1 taps per axis); taps_b*: interpolation filter
 * (2 Rb + 1 taps, not yet multiplied by the stride).  N and M must be multiples of stride.
 * The call covers planes plane_begin, plane_begin + plane_step, ... < plane_end: an interleaved
 * share (plane_step = number of ranks) gives every rank planes near the centre of the grid, which
 * keeps the exact pruning effective when the k-grid is sharded. */
/* The multirate arg-max drops, per 64 x 128 pixel tile, every candidate whose coarse-grid amplitude
 * bound cannot beat the winners already recorded in `key` (exact branch and bound: results are
 * bit-identical with it on or off; it only changes how much work is done).  On by default. */
int gpa_set_pruning(int on);
/* Coarse-tile staging of k_mr_interp: TMA box loads (cp.async.bulk.tensor + mbarrier) for the tiles whose coarse window does
 * not wrap around the frame (default on), per-element cp.async gathers otherwise / when off.  Results are identical. */
int gpa_set_tma(int on);

/* Split pass 2 (R1x > 0; candidate grids only).  All candidates of a plane share their full-rate
 * filtering: G_a = G_1 * G_2 (sigma_a^2 = sigma_1^2 + sigma_2^2), the plane is demodulated by the
 * ANCHOR wx0 = wx_rows[n_rows / 2] and filtered by G_1 once (decimating), and every candidate
 * wx = wx0 + dw then costs a (2 H2x + 1)-tap filter at the COARSE rate: demodulate the coarse samples by
 * delta = dw sigma_a^2 / sigma_2^2, apply G_2, scale by exp(2 pi^2 dw^2 sigma_a^2 sigma_1^2 / sigma_2^2) and
 * rotate back — exactly G_a centred on wx, because G_1(f) G_2(f + delta) is that Gaussian.  The rows that
 * wrapped around the frame edge (whose carrier phase jumps by dw N, as the reference's does) are carried
 * separately through the anchor stage.  taps_1x: G_1 (2 R1x + 1 fine taps); taps_2x: S G_2(S m), |m| <= H2x.
 * R1x = 0 selects the single-stage pass 2 (taps_1x, taps_2x, sigma_a, sigma_1 ignored). */
/* Host-side planner of the split (no CUDA; same search as pygpa_b200/_taps.py): picks the shortest coarse
 * filter (13 ... 23 taps) whose worst-case transfer-function error against the candidate-centred G_a, over
 * all input frequencies and the widest dw of wx_rows, stays below 1.3e-6 (the error of truncating G_a at
 * 4.5 sigma) with a re-amplification c <= 8 (fp32 rounding noise of the anchor stage is multiplied by it).  Returns 1 and fills R1x, H2x, sigma_1, taps_1x (capacity 446 floats) and taps_2x (23 floats);
 * returns 0 when the candidate axis is too wide or too short for one shared anchor (use R1x = 0). */
int gpa_split_plan(int n, int stride, double sigma_a, const double* wx_rows /*host*/, int n_rows,
                   int* R1x, int* H2x, double* sigma_1, float* taps_1x /*host*/, float* taps_2x /*host*/);
int gpa_sweep_mr_workspace_bytes(int N, int M, int n_rows, int n_planes, int cand_mode, int stride,
                                 int Rax, int Ray, int Rb, int R1x, int H2x, int R1y, int H2y, int planes_in_flight,
                                 size_t* bytes);
int gpa_sweep_argmax_mr(const float* img, int N, int M,
                        const double* wx_rows /*host*/, int n_rows,
                        const double* wy_planes /*host*/, int n_planes, int cand_mode,
                        int plane_begin, int plane_end, int plane_step, int stride,
                        const float* taps_ax /*host*/, int Rax, const float* taps_ay /*host*/, int Ray,
                        const float* taps_bx /*host*/, const float* taps_by /*host*/, int Rb,
                        const float* taps_1x /*host*/, int R1x, const float* taps_2x /*host*/, int H2x,
                        double sigma_a, double sigma_1,
                        const float* taps_1y /*host*/, int R1y, const float* taps_2y /*host*/, int H2y, double sigma_1y,
                        unsigned long long* key, void* ws, size_t ws_bytes, void* stream);

/* gpa_sweep_finalize computed from the coarse grids gpa_sweep_argmax_mr left in ws (same ws, same
 * geometry arguments, every plane of [plane_begin, plane_end) resident, nothing else enqueued on ws in
 * between): the winner and its four neighbours are interpolated instead of re-filtered.  Outputs
 * and conventions as gpa_sweep_finalize.  Returns GPA_ERR_WORKSPACE if the range was chunked. */
int gpa_sweep_finalize_mr(int N, int M, const double* wx_rows /*host*/, int n_rows,
                          const double* wy_planes /*host*/, int n_planes, int cand_mode,
                          int plane_begin, int plane_end, int plane_step, int stride, int Rax, int Ray,
                          const float* taps_bx /*host*/, const float* taps_by /*host*/, int Rb,
                          int R1x, int H2x, int R1y, int H2y,
                          const unsigned long long* key, double kref_x, double kref_y, int grad_mode,
                          int out_f64, void* lockin, void* grad, void* w, int* kidx,
                          void* ws, size_t ws_bytes, void* stream);

/* For every pixel whose winning candidate (decoded from key) lies in planes
 * [plane_begin, plane_end): recompute that candidate's lock-in at the pixel and its four
 * neighbours and write
 *   lockin[x*M+y]   = sf * exp(-2 pi i ((wx-kref_x) x + (wy-kref_y) y))          (complex, :808)
 *   grad[(x*M+y)*2] = wrapToPi(2 (grad(-angle sf) + 2 pi (k - kref)))/2          (2 reals, :807-812)
 *   w[c*N*M+x*M+y]  = winning k-vector component c                               ((2,N,M), :809,811)
 *   kidx[x*M+y]     = flat candidate index, -1 where no candidate ever won
 * Other pixels are left untouched.  lockin/grad/w are float (c64) or, with out_f64, double
 * (c128) arrays.  grad may be NULL with grad_mode GPA_GRAD_NONE; w and kidx may be NULL.  planes_valid != 0 promises that ws still holds the planes written by
 * gpa_sweep_argmax for exactly this plane range (all resident); otherwise pass 1 is redone. */
int gpa_sweep_finalize(const float* img, int N, int M,
                       const double* wx_rows /*host*/, int n_rows,
                       const double* wy_planes /*host*/, int n_planes, int cand_mode,
                       int plane_begin, int plane_end, int planes_valid,
                       const float* taps_x /*host*/, int Rx, const float* taps_y /*host*/, int Ry,
                       const unsigned long long* key, double kref_x, double kref_y, int grad_mode,
                       int out_f64, void* lockin, void* grad, void* w, int* kidx,
                       void* ws, size_t ws_bytes, void* stream);

/* One-GPU convenience: zero key, arg-max over all planes, finalize.  key is scratch+output
 * ([N*M] u64).  Equivalent of one call of wfr2_grad_opt (cuGPA.py:41-87) minus host copies. */
int gpa_wfr_sweep(const float* img, int N, int M,
                  const double* wx_rows /*host*/, int n_rows,
                  const double* wy_planes /*host*/, int n_planes, int cand_mode,
                  const float* taps_x /*host*/, int Rx, const float* taps_y /*host*/, int Ry,
                  double kref_x, double kref_y, int grad_mode, int out_f64,
                  unsigned long long* key, void* lockin, void* grad, void* w, int* kidx,
                  void* ws, size_t ws_bytes, void* stream);

/* wfr4 (geometric_phase_analysis.py:839-862): sweep over an ORDERED k-list in which a pixel accepts
 * candidate i only if |sf_i| is strictly larger than the amplitude it holds AND k_i lies within
 * 2 sqrt(2) dk of the k-vector it currently holds (initially klist[0]).  The rule is sequential per
 * pixel; pixels are independent, so one CTA walks the list for its tile (k_pass2_seq).
 *   klist_x, klist_y  host, K entries each (kvec[0] / kvec[1] of the list)
 *   allowed           DEVICE, K*K bytes: allowed[held*K + cand] = (norm(k_held - k_cand) < 2 sqrt(2) dk),
 *                     evaluated by the caller with the reference's float64 expression (:854)
 * Outputs as gpa_wfr_sweep with GPA_CAND_LIST and GPA_GRAD_NONE; pixels that never accept a candidate
 * keep lockin = 0, kidx = -1 and w = klist[0] (the reference's initial state, :850-851). */
int gpa_wfr4_sweep(const float* img, int N, int M,
                   const double* klist_x /*host*/, const double* klist_y /*host*/, int K,
                   const unsigned char* allowed /*device*/,
                   const float* taps_x /*host*/, int Rx, const float* taps_y /*host*/, int Ry,
                   double kref_x, double kref_y, int out_f64,
                   unsigned long long* key, void* lockin, void* w, int* kidx,
                   void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3 — per-pixel phase -> displacement least squares (float64 in, float64 out).
 *
 * Replaces: myweighed_lstsq (geometric_phase_analysis.py:97-113, numba + LAPACK gelsd per
 * pixel), the three branches of reconstruct_u_inv (:157-193) and the wrapped-difference /
 * least-squares stage of reconstruct_u_inv_from_phases (:228-237).
 * ------------------------------------------------------------------------------------------ */

/* where the right-hand side b (d planes) comes from */
#define GPA_LSQ_SRC_PLAIN 0     /* b = src (d,N,M)                                   solve grid (N, M)   */
#define GPA_LSQ_SRC_DIFF1 1     /* b = wrapToPi(diff(src, axis=2)), src (d,N,M)  :234 solve grid (N, M-1) */
#define GPA_LSQ_SRC_DIFF0 2     /* b = wrapToPi(diff(src, axis=1))               :235 solve grid (N-1, M) */
#define GPA_LSQ_SRC_PREDIFF0 3  /* b = wrapToPi(src[...,0])[:, :, :-1], src (d,N,M,2) :229,231  (N, M-1)  */
#define GPA_LSQ_SRC_PREDIFF1 4  /* b = wrapToPi(src[...,1])[:, :-1]              :230,232        (N-1, M)  */

#define GPA_LSQ_WEIGHTED 0      /* per pixel argmin |w (K x - b)|, K = 2 pi kvecs, minimum norm when rank deficient */
#define GPA_LSQ_MATRIX 1        /* x = matrix (2,d) . b : the unweighted pinv(K) (:185) and the two-k inverse (:191) */

int gpa_lstsq_workspace_bytes(int d, size_t* bytes);

/* out (2, n, m) on the solve grid of src_kind.  w is (d, wn, wm) with wn >= n, wm >= m and is
 * indexed with the solve-grid indices, exactly like the reference indexes `w[:, i, j]`.
 * subtract_mean: subtract each plane's mean first (reconstruct_u_inv, :182); needs ws. */
int gpa_lstsq_u(const double* src, int src_kind, const double* w, int wn, int wm,
                const double* kvecs /*host (d,2)*/, int d, int N, int M,
                int solver, const double* matrix /*host (2,d) or NULL*/, int subtract_mean,
                double* out, void* ws, size_t ws_bytes, void* stream);

/* out[i] = sqrt(sum_k w[k*n + i]^2): np.linalg.norm(weights, axis=0) (:240). */
int gpa_norm_axis0(const double* w, int d, size_t n, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5 — local lattice properties from the sweep's phase-gradient maps (float64), SURVEY 8(f) row 1.
 *
 * Replaces: phasegradient2J / phasegradient2Jac (property_extract.py:69-101, 55-66) and
 * props_from_Jac / props_from_J (property_extract.py:137-178, 218-219).
 * ------------------------------------------------------------------------------------------ */

/* Per pixel: b_i[c] = grads[order[i], x, y, c] - sub[i][c] (wrapped to [-pi, pi) when do_wrap),
 * then for c = 0, 1 the weighted least squares argmin | weights[i] (K_i . x - b_i[c]) | with the
 * minimum-norm rule of K3; J[x, y, a, c] = x_a / nmperpixel (+ 1 on the diagonal when add_identity:
 * phasegradient2Jac).  The host supplies what property_extract.py:78-94 derives from the k-vectors:
 * K = 2 pi (kvecs[order] + dks), sub = 2 pi dks (calc_diff_from_isotropic), order (sort != 0).
 *   grads (d, N, M, 2) device, weights (d, wn, wm) device (wn >= N, wm >= M; NOT re-ordered, as in
 *   the reference), K / sub (d, 2) host, order d host (NULL = identity), J (N, M, 2, 2) device. */
int gpa_phasegradient_to_j(const double* grads, const double* weights, int wn, int wm,
                           const double* K /*host*/, const double* sub /*host or NULL*/,
                           const int* order /*host or NULL*/, int do_wrap, int d, int N, int M,
                           double nmperpixel, int add_identity, double* J, void* stream);

/* props_from_Jac on npix 2x2 Jacobians (row-major, + identity when add_identity: props_from_J):
 * props[0] = lattice angle + refangle [deg], props[1] = anisotropy angle mod 180 [deg] (+ 90 when
 * diff), props[2] = refscale * smaller (diff: larger) singular value, props[3] = s0 / s1; props is
 * (4, npix).  The SVD reproduces LAPACK dgesdd's sign conventions, on which the reference's
 * formulas depend (pygpa_b200/csrc/props_device.cuh). */
int gpa_props_from_jac(const double* jac, size_t npix, double refangle, double refscale, int diff,
                       int add_identity, double* props, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6 — device pieces of the k-vector refinement loop (float64), SURVEY 8(f) row 2.
 *
 * Replaces, inside iterate_GPA (geometric_phase_analysis.py:116-154): np.angle / np.abs on the
 * cropped lock-in (:134-139), sqrt(w / w.max()) (:141), and mathtools.fit_plane (mathtools.py:30-47,
 * scipy.optimize.least_squares(loss='huber')) behind fit_delta_k (:92-94).
 * ------------------------------------------------------------------------------------------ */

/* phases = angle(lockin)[edge:N-edge, edge:M-edge], amp = abs(...) likewise (contiguous
 * (N-2 edge, M-2 edge) arrays), *amp_max = max(amp) (device scalar, reset by the call). */
int gpa_lockin_phase_amp(const void* lockin, int is_f64, int N, int M, int edge,
                         double* phases, double* amp, double* amp_max /*device*/, void* stream);

/* out = sqrt(amp / *amp_max): the unwrap weight of iterate_GPA. */
int gpa_weight_sqrt_norm(const double* amp, const double* amp_max /*device*/, size_t n, double* out, void* stream);

int gpa_fit_plane_workspace_bytes(size_t* bytes);

/* Huber-loss plane fit img[x, y] ~ theta[0] x + theta[1] y + theta[2] (x, y array indices,
 * f_scale = 1 in the reference) by iteratively reweighted least squares on the device: each step is
 * one streaming reduction, the 3x3 solve runs in the kernel.  Stops when the plane moves by less than
 * tol anywhere on the frame or after max_iter steps.  theta (3 doubles) and iterations are HOST
 * outputs (the caller's k-vector update is host arithmetic); the call synchronises the stream. */
int gpa_fit_plane_huber(const double* img, int n, int m, double f_scale, int max_iter, double tol,
                        double* theta /*host*/, int* iterations /*host or NULL*/,
                        void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2 — weighted least-squares phase unwrapping (Ghiglia-Romero PCG, DCT Poisson preconditioner),
 * float64 throughout.
 *
 * Replaces: phase_unwrap (phase_unwrap.py:141-208) when `psi` is given, phase_unwrap_prediff
 * (:282-350) when (`dx`, `dy`) are given: dx (N, M-1) = differences along axis 1, dy (N-1, M)
 * along axis 0; both are wrapped to [-pi, pi) first, like the reference does.  weight (N, M) or
 * NULL (unweighted).  Stops after kmax iterations or when |r| < 1e-9 |r0| (at least one
 * iteration, as the reference).  *iterations (host, optional) receives the iteration count and
 * forces a stream synchronisation.
 * ------------------------------------------------------------------------------------------ */
/* K2 transforms of the PCG: 1 (default) = pipelined kernels (persistent CTAs; rows prefetched by cp.async.bulk, 4-column strips
 * by TMA boxes; <r, z> from the DCT coefficients in the column stage; p = z + beta p fused into the last inverse pass),
 * 0 = one CTA per pair of rows / per strip of 8 columns.  The two agree to rounding (1e-13 on PCG iterates). */
int gpa_set_dct_pipeline(int on);

/* Device mirrors of the reference's solver helpers, all float64:
 *   gpa_dctn           scipy.fft.dctn / idctn (type 2, norm=None) of an (N, M) array — the transform pair of solvePoisson
 *                      and solvePoisson_precomped (phase_unwrap.py:81-103); ws as gpa_unwrap_workspace_bytes(N, M)
 *   gpa_poisson_scale  precomp_Poissonscaling (phase_unwrap.py:106-115): 2 (cos(pi I/M) + cos(pi J/N) - 2), [0,0] = 1
 *   gpa_divide_f64     out = a / b elementwise (dctn(rho) / scale)
 *   gpa_apply_q        applyQ (phase_unwrap.py:118-132); ws >= 512 + 8 ceil(M/64) ceil(N/32) bytes */
int gpa_dctn(const double* in, int N, int M, int inverse, double* out, void* ws, size_t ws_bytes, void* stream);
int gpa_poisson_scale(int N, int M, double* scale, void* stream);
int gpa_divide_f64(const double* a, const double* b, double* out, size_t n, void* stream);
int gpa_apply_q(const double* p, const double* wwx, const double* wwy, int N, int M, double* q,
                void* ws, size_t ws_bytes, void* stream);
int gpa_unwrap_workspace_bytes(int N, int M, size_t* bytes);
int gpa_unwrap_pcg(const double* psi, const double* dx, const double* dy, const double* weight,
                   int N, int M, int kmax, double* phi, int* iterations /*host or NULL*/,
                   void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4 — Lawler-Fujita: invert the displacement field and resample (float64).
 *
 * Replaces: invert_u_overlap (geometric_phase_analysis.py:262-300; invert_u :248-259 is the
 * edge = 0 case up to its `- edge` typo) and undistort_image (:935-974), i.e. 73 calls of
 * scipy.ndimage.map_coordinates(order=3) — mode='nearest' for the fixed-point inversion,
 * mode='constant', cval=0 for the final resample.  The cubic-spline prefilter of u runs once
 * instead of 72 times and the fixed-point loop of a pixel runs in registers.
 * ------------------------------------------------------------------------------------------ */
int gpa_lawler_workspace_bytes(int N, int M, int edge, size_t* bytes);

/* out (2, N+2 edge, M+2 edge): u_it <- (scale u)(r), then `iters` times u_it <- (scale u)(r + u_it),
 * r on the grid [-edge, N+edge) x [-edge, M+edge), scipy mode='nearest'.  u is (2, N, M). */
int gpa_invert_u(const double* u, int N, int M, double scale, int iters, int edge, double* out,
                 void* ws, size_t ws_bytes, void* stream);

/* invert_u(us, iters, edge) of the reference (geometric_phase_analysis.py:248-259): out (2, N, M), u_it <- (scale u)(r),
 * then `iters` times u_it <- (scale u)(r - edge + u_it) — the `- edge` enters the iterations only, as there. */
int gpa_invert_u_plain(const double* u, int N, int M, double scale, int iters, int edge, double* out,
                       void* ws, size_t ws_bytes, void* stream);

/* out (N, M) = cubic-spline resampling of img at (r + u_inv[0], c + u_inv[1]), 0 outside the frame
 * (scipy mode='constant', cval=0). */
int gpa_resample_image(const double* img, int N, int M, const double* u_inv, double* out,
                       void* ws, size_t ws_bytes, void* stream);

/* undistort_image(deformed, u): gpa_invert_u(u, scale=-1, iters, edge=0) then gpa_resample_image. */
int gpa_undistort_image(const double* img, const double* u, int N, int M, int iters, double* out,
                        void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7 — unit-cell averaging (float64), SURVEY 8(f) row 4.
 *
 * Replaces unit_cell_average (unit_cell_averaging.py:132-205, numba loop with add_to_position
 * :208-217) and expand_unitcell (:234-249, scipy map_coordinates).  The host passes what
 * calc_ucell_parameters (:45-53) derives from the two k-vectors: ks (2,2) row-major, its inverse,
 * the lower cell corner rmin and the zoomed cell array size (rs0, rs1).
 * ------------------------------------------------------------------------------------------ */
int gpa_uc_workspace_bytes(int rs0, int rs1, size_t* bytes);

/* out (rs0, rs1) = drizzle average of img (N, M) [displaced by u (2, N, M), may be NULL] over the unit
 * cell, zoom z; NaN pixels of img are skipped, cells nothing landed on are NaN (0 / 0). */
int gpa_uc_average(const double* img, const double* u, int N, int M,
                   const double* ks /*host*/, const double* kinv /*host*/, const double* rmin /*host*/,
                   double z, int rs0, int rs1, double* out, void* ws, size_t ws_bytes, void* stream);

/* out (H, W) = cubic-spline (order 3, mode 'constant', cval 0) samples of the NaN-cleared cell image
 * ucell (n, m) at the folded coordinates of r / z2 + u (u (2, H, W), or NULL with the scalar u_const). */
int gpa_uc_expand(const double* ucell, int n, int m, int H, int W, const double* u, double u_const,
                  double z2, const double* ks /*host*/, const double* kinv /*host*/,
                  const double* rmin /*host*/, double z, double* out,
                  void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K8 — gaussian_deconvolve (float64), SURVEY 8(f) row 3.
 *
 * Replaces gaussian_deconvolve (geometric_phase_analysis.py:892-904): every (N, M) plane of `data` is
 * reflect-padded by 2 dr, filtered with skimage's Wiener filter for the Gaussian PSF of std sigma
 * (W = H / (H^2 + balance L^2) on the DFT grid of the padded frame) and cropped back.  The padded
 * axes (N + 4 dr, M + 4 dr) may have any length up to 4096 (Bluestein transforms in shared memory).
 * ------------------------------------------------------------------------------------------ */
int gpa_deconvolve_workspace_bytes(int N, int M, int dr, size_t* bytes);
int gpa_gaussian_deconvolve(const double* data, int planes, int N, int M, double sigma, int dr,
                            double balance, double* out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1e — k-grid sharding over the GPUs of one NVSwitch box (the reference's
 * cuGPA.py is single-device, so this has no reference twin).  One process per GPU; every rank
 * gpa_peer_alloc's one arena, the host binding exchanges the 64-byte IPC handles (torch.distributed)
 * and gpa_peer_open's the arenas of the other ranks, after which the kernels address peer HBM
 * directly over NVLink.  Per peak and frame:
 *     gpa_sweep_argmax_mr (local planes)  ->  gpa_peer_signal / gpa_peer_wait  ->  gpa_key_merge
 *     ->  gpa_peer_signal / gpa_peer_wait  ->  gpa_sweep_finalize_mr_sharded (owner-writes)
 *     ->  gpa_peer_signal to the destination rank(s), which gpa_peer_wait.
 * Flags carry a monotonically increasing epoch (one per frame), so they are never reset.
 * ------------------------------------------------------------------------------------------ */
#define GPA_MAX_PEERS 8
#define GPA_PEER_HANDLE_BYTES 64
/* cudaMalloc + zero + cudaIpcGetMemHandle / cudaIpcOpenMemHandle / cudaIpcCloseMemHandle / cudaFree.
 * handle: host buffer of GPA_PEER_HANDLE_BYTES. */
int gpa_peer_alloc(size_t bytes, void** dev_ptr /*host, out*/, unsigned char* handle /*host, out*/);
int gpa_peer_open(const unsigned char* handle /*host*/, void** dev_ptr /*host, out*/);
int gpa_peer_close(void* dev_ptr);
int gpa_peer_free(void* dev_ptr);
/* target_slots: host array of n_targets device addresses (the caller's own slot in every target's flag
 * array, peer-mapped).  Everything enqueued on `stream` before the signal is visible to a rank that
 * has waited for this epoch. */
int gpa_peer_signal(void* const* target_slots /*host*/, int n_targets, unsigned long long epoch, void* stream);
/* Blocks the STREAM (not the host) until flags[0..n_sources) >= epoch; gives up after timeout_s and
 * sets *status (device int) to 1 + the index of the missing source. */
int gpa_peer_wait(const unsigned long long* flags, int n_sources, unsigned long long epoch, double timeout_s,
                  int* status, void* stream);
/* dst <- src on `stream`; either side may be local, peer-mapped or page-locked host memory (cudaMemcpyDefault). */
int gpa_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
/* Page-lock an existing host range (a shared-memory segment every rank maps) so that each GPU can DMA its
 * share of the results into it over its own PCIe link. */
int gpa_host_register(void* host_ptr /*host*/, size_t bytes);
int gpa_host_unregister(void* host_ptr /*host*/);
/* In-place max-with-index all-reduce of the packed arg-max keys (see gpa_sweep_argmax): key_ptrs is the
 * host array of the `world` ranks' key buffers (same layout, n_keys each, own buffer at [rank]); this
 * rank reduces its 1/world slice of every buffer and stores the result into all of them.  The caller
 * brackets it with signal / wait pairs (all ranks have finished their local arg-max before, all ranks
 * have merged their slice after). */
int gpa_key_merge(void* const* key_ptrs /*host*/, int world, int rank, size_t n_keys, void* stream);
/* w[i] = wx_dev[row], w[comp_stride + i] = wy_dev[plane] of the winner packed in key[i] (0 where nothing
 * won): the 'w' output of wfr2_grad_opt from merged keys. */
int gpa_key_to_w(const unsigned long long* key, size_t n, size_t comp_stride, const double* wx_dev,
                 const double* wy_dev, int n_planes, int list_mode, int out_f64, void* w, void* stream);
/* Threshold gossip for the exact pruning of a sharded sweep.  A rank only sweeps its own planes, so the winners it has
 * recorded are weaker pruning thresholds than a single GPU would have.  hint_ptrs: host array of n_ranks device arrays
 * of (N / B) x (M / B) uint64 (B = 8 * stride pixels per bound block), entry 0 this rank's own, the rest the peers'
 * (peer-mapped), zero-initialised once.  The NEXT gpa_sweep_argmax_mr call on this thread then (a) raises its per-block
 * thresholds to the bounds found there for `epoch` and (b) pushes every bound it improves to all n_ranks arrays
 * (64-bit max over NVLink, tagged with the epoch: stale frames never match, nothing is ever reset).  Any published
 * value is a lower bound of the block's final winners, so results stay bit-identical to the unpruned sweep. */
int gpa_sweep_arm_gossip(void* const* hint_ptrs /*host*/, int n_ranks, unsigned int epoch);
/* Two-phase variant (call right after gpa_sweep_arm_gossip, same ordering of the ranks: entry 0 = this rank).  The next
 * gpa_sweep_argmax_mr then (1) publishes, per tile of k_mr_interp, the largest bound any of its planes reaches into row
 * `rank` of every rank's table best_ptrs[r] (float [n_ranks][tiles]), (2) after a flag barrier sweeps ONLY the tiles where
 * it holds the globally most promising plane — the one CTA per tile a single GPU would run first, unpruned — and
 * (3) after a second flag barrier everything else, against the thresholds phase (2) published through the gossip arrays.
 * Without it every rank starts every tile unpruned: W unpruned plane sweeps per tile instead of one.
 * flag_slots_a / _b: this rank's slot in every rank's flag array for the two barriers; wait_a / _b: this rank's own
 * n_ranks slots (device); epoch as gpa_peer_signal; *status, timeout_s as gpa_peer_wait.  Applies only when the call
 * sweeps its planes in one resident chunk. */
int gpa_sweep_arm_two_phase(void* const* best_ptrs /*host*/, void* const* flag_slots_a /*host*/, void* const* flag_slots_b /*host*/,
                            const unsigned long long* wait_a, const unsigned long long* wait_b, int rank,
                            unsigned long long epoch, double timeout_s, int* status);
/* gpa_sweep_finalize_mr for one rank's share of the planes with owner-writes: pixel (x, y) of a winner
 * this rank owns is stored to lockin_dst[x / dst_rows] and grad_dst[x / dst_rows] (host arrays of n_dst
 * device pointers to full (N, M) / (N, M, 2) arrays, local or peer-mapped; all pixels go to entry 0 when
 * n_dst == 1).  Pixels nothing won are zero-filled by the rank that passes write_zero != 0.  Results are
 * bit-identical to the single-GPU gpa_sweep_finalize_mr. */
int gpa_sweep_finalize_mr_sharded(int N, int M, const double* wx_rows /*host*/, int n_rows,
                                  const double* wy_planes /*host*/, int n_planes, int cand_mode,
                                  int plane_begin, int plane_end, int plane_step, int stride, int Rax, int Ray,
                                  const float* taps_bx /*host*/, const float* taps_by /*host*/, int Rb,
                                  int R1x, int H2x, int R1y, int H2y,
                                  const unsigned long long* key, double kref_x, double kref_y, int grad_mode,
                                  int out_f64, void* const* lockin_dst /*host*/, void* const* grad_dst /*host*/,
                                  int n_dst, int dst_rows, int write_zero,
                                  void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPA_B200_H */
