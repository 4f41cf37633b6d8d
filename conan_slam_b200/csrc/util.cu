// util.cu — error reporting and library-level entry points of libcslam.so.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cslam {
static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace cslam

extern "C" {
unsigned long long cslam_kernel_launches(void) { return cslam::g_launches.load(); }
const char* cslam_last_error(void) { return cslam::g_last_error; }
int cslam_version(void) { return CSLAM_VERSION; }
int cslam_device_count(int* count) {
    CSLAM_REQUIRE(count != nullptr, CSLAM_ERR_BAD_ARG, "count is null");
    *count = 0;
    CSLAM_CUDA(cudaGetDeviceCount(count));
    return CSLAM_OK;
}
}
