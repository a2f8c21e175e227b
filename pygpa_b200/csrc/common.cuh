// Shared helpers for libgpa_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/gpa_b200.h"

namespace gpa {

void set_error(const char* fmt, ...);

#define GPA_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            gpa::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                           __FILE__, __LINE__);                                           \
            return GPA_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define GPA_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            gpa::set_error(__VA_ARGS__);                                                  \
            return GPA_ERR_INVALID;                                                       \
        }                                                                                 \
    } while (0)

// Optional per-kernel timing (CUDA events on the launch stream), off by default.
// bench.py turns it on to report the dominant kernel's live duration.
struct KernelTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    bool on;
    KernelTimer(const char* name, cudaStream_t stream);
    ~KernelTimer();
    const char* name;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Bump allocator over the caller's workspace.
struct Arena {
    char* base;
    size_t size, off;
    Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size; }
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// exp(2 pi i t) for a phase given in TURNS, evaluated with double range reduction so the
// result is accurate to fp32 rounding even when |t| is thousands of turns.
__device__ __forceinline__ float2 phasor_turns(double t) {
    double f = t - rint(t);                      // [-0.5, 0.5]
    float s, c;
    sincospif(2.0f * static_cast<float>(f), &s, &c);
    return make_float2(c, s);
}

}  // namespace gpa
