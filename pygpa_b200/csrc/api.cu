// Error reporting and library-level queries of the C ABI (include/gpa_b200.h).
#include "common.cuh"

#include <mutex>
#include <vector>

namespace gpa {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- per-kernel timing ------------------------------------------------------------------
struct TimedLaunch {
    const char* name;
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<TimedLaunch> g_prof;

KernelTimer::KernelTimer(const char* n, cudaStream_t stream) : st(stream), on(g_prof_on), name(n) {
    if (!on) return;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
}
KernelTimer::~KernelTimer() {
    if (!on) return;
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({name, e0, e1});
}
}  // namespace gpa

extern "C" int gpa_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(gpa::g_prof_mu);
    gpa::g_prof_on = on != 0;
    return GPA_OK;
}

extern "C" int gpa_profile_read(const char* kernel, double* total_ms, int* launches, int reset) {
    GPA_REQUIRE(kernel && total_ms && launches, "null pointer argument");
    std::lock_guard<std::mutex> lk(gpa::g_prof_mu);
    double tot = 0;
    int n = 0;
    for (auto& t : gpa::g_prof) {
        if (std::strcmp(t.name, kernel) != 0) continue;
        GPA_CHECK_CUDA(cudaEventSynchronize(t.e1));
        float ms = 0;
        GPA_CHECK_CUDA(cudaEventElapsedTime(&ms, t.e0, t.e1));
        tot += ms;
        ++n;
    }
    *total_ms = tot;
    *launches = n;
    if (reset) {
        for (auto& t : gpa::g_prof) {
            cudaEventDestroy(t.e0);
            cudaEventDestroy(t.e1);
        }
        gpa::g_prof.clear();
    }
    return GPA_OK;
}

extern "C" const char* gpa_last_error(void) { return gpa::g_err; }

extern "C" int gpa_version(void) { return 100; }  // 0.1.0

extern "C" int gpa_device_sm_count(void) {
    int dev = 0, n = 0;
    GPA_CHECK_CUDA(cudaGetDevice(&dev));
    GPA_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}
