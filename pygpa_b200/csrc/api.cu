// Error reporting and library-level queries of the C ABI (include/gpa_b200.h).
#include "common.cuh"

namespace gpa {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gpa

extern "C" const char* gpa_last_error(void) { return gpa::g_err; }

extern "C" int gpa_version(void) { return 100; }  // 0.1.0

extern "C" int gpa_device_sm_count(void) {
    int dev = 0, n = 0;
    GPA_CHECK_CUDA(cudaGetDevice(&dev));
    GPA_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}
